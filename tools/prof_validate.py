#!/usr/bin/env python3
"""Driver for ncu captures / timing of MemoryChain.validate_chain() on reference-shaped Python block objects:
tools/prof_validate.py [blocks].  The blocks are the deterministic synthetic chain (hashes fetched from the device generator);
constructing the chain marshals them into typed columns and builds the canonical JSON on the GPU (k_json_size / k_json_write),
validate_chain() then runs k_prepare_links / k_sha256_validate on resident data."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fei_b200 import _abi, synth
from fei_b200.memdir_tools import memorychain as mc
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
_abi.init(); lib = _abi.lib()
ch = C.c_void_p(); _abi.check(lib.fei_chain_create(C.byref(ch)))
_abi.check(lib.fei_chain_synth(ch, bench.CHAIN_SEED, 0, n, -1))
hh = np.zeros(64 * n, dtype=np.uint8); moff = np.zeros(n + 1, dtype=np.uint64)
_abi.check(lib.fei_chain_fetch(ch, 0, n, None, 0, _abi.ptr(moff), _abi.ptr(hh), None))
lib.fei_chain_destroy(ch)
hashes = hh.tobytes().decode()
blocks = []
for i, sp in enumerate(synth.chain_specs(bench.CHAIN_SEED, 0, n)):
    b = mc.MemoryBlock(sp["index"], sp["timestamp"], sp["memory_data"], "0" if i == 0 else hashes[64 * i - 64:64 * i], sp["responsible_node"], sp["proposer_node"])
    b.hash = hashes[64 * i:64 * i + 64]
    blocks.append(b)
t0 = time.perf_counter(); chain = mc.MemoryChain(blocks=blocks); t1 = time.perf_counter()
ok = chain.validate_chain(); t2 = time.perf_counter()
ok2 = chain.validate_chain(); t3 = time.perf_counter()
print(f"{n} blocks: construction {1e3 * (t1 - t0):.1f} ms, first validate {1e3 * (t2 - t1):.2f} ms ({ok}), resident validate {1e3 * (t3 - t2):.2f} ms ({ok2})")
