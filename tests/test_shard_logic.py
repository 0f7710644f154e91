"""CPU suite for the N>1 host logic: range sharding, global indices, rank-order concatenation, and the
hit all-gatherv protocol over a world_size-2 gloo group (the oracle stands in for the per-rank scan)."""
import os

import numpy as np
import pytest

from fei_b200 import shard


def test_shard_ranges_cover_and_order():
    for n in (0, 1, 7, 8, 9, 1000, 10_000_001):
        for w in (1, 2, 4, 8):
            r = shard.shard_ranges(n, w)
            assert len(r) == w and r[0][0] == 0 and r[-1][1] == n
            assert all(a <= b for a, b in r) and all(r[k][1] == r[k + 1][0] for k in range(w - 1))
    off = np.concatenate([[0], np.cumsum(np.random.default_rng(1).integers(300, 7000, size=5000))]).astype(np.uint64)
    r = shard.shard_ranges_by_bytes(off, 8)
    sizes = [int(off[b] - off[a]) for a, b in r]
    assert r[0][0] == 0 and r[-1][1] == 5000 and max(sizes) - min(sizes) < 2 * 7000
    assert shard.chain_shard_ranges(100, 4) == [(0, 25), (24, 50), (49, 75), (74, 100)]
    assert shard.merge_first_bad([(-1, 0), (57, 1), (30, 2), (-1, 0)]) == (30, 2)
    assert shard.merge_first_bad([(-1, 0)]) == (-1, 0)


def _worker(rank, world, port, q):
    import re
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fei_b200 import synth
    from oracle import memdir_oracle as mo
    n = 600
    a, b = shard.shard_ranges(n, world)[rank]
    recs = [synth.record(0xFE1, i) for i in range(a, b)]
    mems = [mo.make_memory(r["filename"], r["folder"], r["status"], synth.file_text(r), True) for r in recs]
    pats = ["python", "docker|kubernetes", "rust", "zzz-no-hit"]
    local = [np.array([a + i for i in mo.run_search(mems, [{"field": "content", "operator": "matches", "value": p}])], dtype=np.uint64) for p in pats]
    got = shard.gather_hit_lists(dist, local)
    if rank == 0:
        q.put([g.tolist() for g in got])
    dist.destroy_process_group()


def test_gloo_world2_gather_equals_single_shard():
    import socket
    import torch.multiprocessing as mp
    from fei_b200 import synth
    from oracle import memdir_oracle as mo
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    recs = [synth.record(0xFE1, i) for i in range(600)]
    mems = [mo.make_memory(r["filename"], r["folder"], r["status"], synth.file_text(r), True) for r in recs]
    for k, pat in enumerate(["python", "docker|kubernetes", "rust", "zzz-no-hit"]):
        assert got[k] == mo.run_search(mems, [{"field": "content", "operator": "matches", "value": pat}]), pat
    assert got[3] == []
    shard.concat_in_rank_order([[np.array(g[:len(g) // 2], dtype=np.uint64) for g in got], [np.array(g[len(g) // 2:], dtype=np.uint64) for g in got]])
