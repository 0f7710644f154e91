#!/usr/bin/env python3
"""Tiny driver for ncu captures of the chain kernels: tools/prof_chain.py [blocks]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fei_b200 import _abi
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
_abi.init(); lib = _abi.lib()
ch = C.c_void_p(); _abi.check(lib.fei_chain_create(C.byref(ch)))
_abi.check(lib.fei_chain_synth(ch, bench.CHAIN_SEED, 0, n, -1))
fb, kind, kms = C.c_int64(), C.c_int32(), C.c_float()
ms = []
for _ in range(6):
    _abi.check(lib.fei_chain_validate(ch, C.byref(fb), C.byref(kind), None, C.byref(kms))); ms.append(kms.value)
print("chain", n, fb.value, kind.value, ["%.4f" % m for m in ms], "blocks/s %.3e" % ((n - 1) / (min(ms) * 1e-3)))
