"""Range sharding of a corpus / chain over ranks and the host side of the hit all-gatherv.

Records are independent, so the scan shards by contiguous record range: rank r owns
[start_r, end_r); every rank reports GLOBAL indices, so the rank-order concatenation of the per-rank
ordered hit lists is the global listing order and no sort is needed.  On GPUs the exchange is one NCCL
all-gatherv inside libfeiscan (fei_comm_allgather_hits); `gather_hit_lists` below is the same protocol
over a torch.distributed process group (counts all-gather, then one padded all-gather), used where the
lists already live on the host (tests on gloo, tooling).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_ranges(n_total: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, near-equal record ranges in rank order (ceil split, trailing ranks may be empty)."""
    per = -(-n_total // world) if world > 0 else n_total
    return [(min(n_total, r * per), min(n_total, (r + 1) * per)) for r in range(world)]


def shard_ranges_by_bytes(offsets: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Ranges balanced on the byte prefix sum (offsets[n+1]) for skewed record sizes."""
    n = len(offsets) - 1
    total = int(offsets[n])
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(offsets, total * r // world, side="left")))
    cuts.append(n)
    cuts = np.maximum.accumulate(np.clip(cuts, 0, n))
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]


def chain_shard_ranges(n_blocks: int, world: int) -> List[Tuple[int, int]]:
    """Chain shards overlap by one block: position 0 of a shard is only the predecessor of position 1
    (exactly how the genesis block is treated, memorychain.py:604)."""
    base = shard_ranges(n_blocks, world)
    return [(max(0, a - 1), b) for a, b in base]


def merge_first_bad(verdicts: Sequence[Tuple[int, int]]) -> Tuple[int, int]:
    """min over ranks of (index, kind) with -1 = no failure: the reference reports the first failing block."""
    bad = [(i, k) for i, k in verdicts if i >= 0]
    return min(bad) if bad else (-1, 0)


def concat_in_rank_order(per_rank: Sequence[Sequence[np.ndarray]]) -> List[np.ndarray]:
    """per_rank[r][q] -> gathered[q]; asserts the global-order invariant."""
    nq = len(per_rank[0])
    out = []
    for q in range(nq):
        cat = np.concatenate([np.asarray(per_rank[r][q], dtype=np.uint64) for r in range(len(per_rank))])
        assert np.all(cat[1:] > cat[:-1]) if cat.size > 1 else True, "rank-order concatenation must be strictly increasing"
        out.append(cat)
    return out


def gather_hit_lists(dist, local: Sequence[np.ndarray]) -> List[np.ndarray]:
    """All-gatherv of per-query ordered hit lists over a torch.distributed group (same protocol as the NCCL path)."""
    import torch
    world = dist.get_world_size()
    nq = len(local)
    counts = torch.tensor([len(x) for x in local], dtype=torch.int64)
    all_counts = [torch.zeros(nq, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(all_counts, counts)
    width = int(max(int(c.max()) if nq else 0 for c in all_counts)) if world else 0
    pad = torch.zeros((nq, max(width, 1)), dtype=torch.int64)
    for q, x in enumerate(local):
        if len(x):
            pad[q, :len(x)] = torch.from_numpy(np.asarray(x, dtype=np.uint64).astype(np.int64))
    bufs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    out = []
    for q in range(nq):
        out.append(np.concatenate([bufs[r][q, :int(all_counts[r][q])].numpy().astype(np.uint64) for r in range(world)]))
    return out
