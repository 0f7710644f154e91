#!/usr/bin/env python3
"""Where does MemoryChain.validate_chain() on Python block objects spend its time?  tools/prof_validate.py [blocks]
(FEI_DEBUG_TIMING=1 adds the library's own split: JSON serialisation / H2D + padding / kernel)"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fei_b200 import _abi, synth
from fei_b200.memdir_tools import memorychain as mc
from oracle import chain_oracle as co
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
blocks = []
for ob in co.build_chain(synth.chain_specs(bench.CHAIN_SEED, 0, n)):
    b = mc.MemoryBlock(ob.index, ob.timestamp, ob.memory_data, ob.previous_hash, ob.responsible_node, ob.proposer_node)
    b.nonce = ob.nonce; b.hash = ob.hash
    blocks.append(b)
ch = mc.MemoryChain(blocks=blocks)
assert ch.validate_chain()
for rep in range(3):
    t0 = time.perf_counter(); nat = mc.chain_columns_native(blocks); t1 = time.perf_counter()
    cols, stored = mc.chain_columns(blocks); hb, ho = mc._str_blob(stored); t2 = time.perf_counter()
    arr = mc._cols_struct(nat[0] if nat else cols)
    fb, kind = C.c_int64(-1), C.c_int32(0)
    t3 = time.perf_counter()
    _abi.check(_abi.lib().fei_chain_validate_cols(arr, _abi.ptr(hb), _abi.ptr(ho), n, 0, C.byref(fb), C.byref(kind), None, None, 0, None)); t4 = time.perf_counter()
    t5 = time.perf_counter(); ok = ch.validate_chain(); t6 = time.perf_counter()
    print("marshal native %.1f ms | python %.1f ms | C call %.1f ms | validate_chain() total %.1f ms = %.3g blocks/s" %
          ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t4 - t3) * 1e3, (t6 - t5) * 1e3, n / (t6 - t5)))
