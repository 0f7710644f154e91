#!/usr/bin/env python3
"""Multi-GPU parity check (run under torchrun, one rank per GPU; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py

Every rank scans its record range, the hit lists are all-gathered through libfeiscan's NCCL path
(fei_comm_allgather_hits, both the sparse list format and the dense mask format), and rank 0 compares the
gathered global lists with the oracle on the whole corpus.  A sharded chain is validated with the 8-byte
min-reduce (fei_comm_allreduce_first_bad)."""
import ctypes as C
import os
import re
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
os.environ.setdefault("TZ", "UTC")


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from fei_b200 import _abi, shard, synth
    from fei_b200.corpus import Corpus
    from fei_b200.program import C_BODY, Cond, ProgramBuilder
    from fei_b200.regexc import Pattern
    from oracle import chain_oracle as co, memdir_oracle as mo
    lib = _abi.lib()
    _abi.init(local)
    idbuf = np.zeros(_abi.NCCL_ID_BYTES, dtype=np.uint8)
    if rank == 0:
        _abi.check(lib.fei_comm_unique_id(_abi.ptr(idbuf)))
    t = torch.from_numpy(idbuf).cuda(); dist.broadcast(t, 0); idbuf = t.cpu().numpy()
    _abi.check(lib.fei_comm_init(_abi.ptr(idbuf), world, rank))

    n = 3001                                               # uneven shards on purpose
    a, b = shard.shard_ranges(n, world)[rank]
    corpus = Corpus().synth(0xFE1, a, b - a)
    for name, pats in (("sparse", ["zebra", r"kubernetes.*docker.*terraform", "rust.*python.*go"]), ("dense", ["python", "docker|kubernetes", "react", "e"])):
        pb = ProgramBuilder()
        for p in pats:
            pb.add_query([Cond(C_BODY, pattern=Pattern("regex", p, re.IGNORECASE))])
        prog = pb.build(); nq = len(pats)
        corpus.scan_count(prog, nq)
        bufs = [np.zeros(n, dtype=np.uint64) for _ in range(nq)]
        ptrs = (C.c_void_p * 32)(*[x.ctypes.data for x in bufs])
        cap = np.zeros(32, dtype=np.uint64); cap[:nq] = n
        tot = np.zeros(32, dtype=np.uint64); counts = np.zeros(world * nq, dtype=np.uint64)
        _abi.check(lib.fei_comm_allgather_hits(corpus.handle, nq, ptrs, _abi.ptr(cap), _abi.ptr(tot), _abi.ptr(counts)))
        if rank == 0:
            recs = [synth.record(0xFE1, i) for i in range(n)]
            mems = [mo.make_memory(r["filename"], r["folder"], r["status"], synth.file_text(r), True) for r in recs]
            for q, p in enumerate(pats):
                want = mo.run_search(mems, [{"field": "content", "operator": "matches", "value": p}])
                got = bufs[q][:int(tot[q])].tolist()
                assert got == want, (name, p, len(got), len(want))
            print(f"[multi_gpu_check] {name} all-gatherv over {world} ranks: ok ({[int(x) for x in tot[:nq]]} hits)", flush=True)

    # range-sharded chain with a one-block halo; first failure = min over ranks
    nb, bad_at = 4000, 2777
    ch = C.c_void_p(); _abi.check(lib.fei_chain_create(C.byref(ch)))
    lo, hi = shard.shard_ranges(nb, world)[rank]
    _abi.check(lib.fei_chain_synth(ch, 0xC4A1, lo, hi - lo, bad_at))
    fb, kind = C.c_int64(), C.c_int32()
    _abi.check(lib.fei_chain_validate(ch, C.byref(fb), C.byref(kind), None, None))
    _abi.check(lib.fei_comm_allreduce_first_bad(C.byref(fb), C.byref(kind)))
    assert (fb.value, kind.value) == (bad_at, 1), (fb.value, kind.value)
    if rank == 0:
        print(f"[multi_gpu_check] sharded chain: first bad block {fb.value} kind {kind.value}: ok", flush=True)
    lib.fei_chain_destroy(ch)
    lib.fei_comm_destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
