#!/usr/bin/env python3
"""Tiny driver for ncu launch lists of the raw-ingest path: tools/prof_load.py [records] [loads]
Builds `records` synthetic files (header + '---' + body, as on disk) in host memory and runs fei_corpus_load_raw `loads` times."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("TZ", "UTC")
import numpy as np
from fei_b200 import _abi
from fei_b200.corpus import Corpus
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
loads = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lib = _abi.lib()
src = Corpus().synth(bench.SEED, 0, n)
host = src.fetch(0, n)
hdr, ho, body, bo = host["hdr"].tobytes(), host["hdr_off"], host["body"].tobytes(), host["body_off"]
parts = []
for i in range(n):
    parts.append(hdr[int(ho[i]):int(ho[i + 1])]); parts.append(b"---\n"); parts.append(body[int(bo[i]):int(bo[i + 1])])
raw = np.frombuffer(b"".join(parts), dtype=np.uint8).copy()
raw_off = np.zeros(n + 1, dtype=np.uint64)
np.cumsum((ho[1:] - ho[:-1]) + 4 + (bo[1:] - bo[:-1]), out=raw_off[1:])
arrays = {"n": n, "global_base": 0, "raw": raw, "raw_off": raw_off, "ts": host["ts"], "wall": host["wall"], "flags8": host["flags8"],
          "fsb": (host["fsb"] & np.uint32(0x00FFFFFF))}
lib.fei_host_register(raw.ctypes.data, raw.nbytes)
src.close()
c = Corpus()
st = np.zeros(3, dtype=np.float32)
for k in range(loads):
    t0 = time.perf_counter()
    assert c.load_raw(arrays).all()
    _abi.check(lib.fei_corpus_last_load_timing(c.handle, _abi.ptr(st)))
    print(f"load {k}: wall {1e3 * (time.perf_counter() - t0):.1f} ms, text H2D {st[0]:.1f} ms ({st[2]:.1f} GB/s), pack stage {st[1]:.1f} ms", flush=True)
