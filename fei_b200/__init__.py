"""fei_b200 — sm_100a scan engine behind Fei's Memdir search/filter and Memorychain validation.

Layout:
  csrc/            CUDA kernels + C ABI (built in-tree into libfeiscan.so, see include/feiscan.h)
  _abi.py          ctypes binding (the only device boundary; no torch)
  regexc/          Python `re` pattern -> UTF-8 byte DFA compiler (host side of the scan)
  program.py       SearchQuery / MemoryFilter -> predicate program blob
  corpus.py        packed corpus handle (pack once, scan many)
  memdir_tools/    mirror of the reference's memdir_tools API for the hot path
"""
__version__ = "0.1.0"
