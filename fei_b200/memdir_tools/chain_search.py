"""Substring / tag search over the memories stored in a Memorychain, on the GPU.

Replaces the loops of the reference's MemorychainConnector.search_memories and search_by_tag
(fei/tools/memorychain_connector.py:273-362): for every block but the genesis one(s),
    query.lower() in headers.get("Subject", "").lower()  |  ... in headers.get("Tags", "").lower()  |  ... in content.lower()
and  tag.lower() in [t.strip() for t in headers.get("Tags", "").lower().split(",")].

The three searched values of every block are packed once as three records of a corpus (value = record body: header
parsing plays no part, a value may hold newlines or '---'); a search is one content scan, the verdicts are OR-ed per block.
"""
from __future__ import annotations

from typing import Any, Dict, List, Sequence

import numpy as np

from ..corpus import Corpus
from ..program import C_BODY, Cond, ProgramBuilder
from ..regexc import Pattern

_FIELDS = ("Subject", "Tags", "content")


def _memory_of(block: Any) -> Dict[str, Any]:
    return block["memory_data"] if isinstance(block, dict) else block.memory_data


class ChainMemorySearch:
    """Pack once (per chain state), search many times."""

    def __init__(self, chain: Sequence[Any]):
        self.memories = [_memory_of(b) for b in chain]
        n = len(self.memories)
        vals: List[bytes] = []
        self.genesis = np.zeros(n, dtype=bool)
        for i, m in enumerate(self.memories):
            headers = m.get("headers", {})
            trio = (headers.get("Subject", ""), headers.get("Tags", ""), m.get("content", ""))
            for name, v in zip(_FIELDS, trio):
                if not isinstance(v, str):
                    raise NotImplementedError(f"block {i}: {name} is a {type(v).__name__}; the reference calls .lower() on it (str only)")
                vals.append(v.encode("utf-8", "surrogatepass"))
            self.genesis[i] = m.get("metadata", {}).get("unique_id", "") == "genesis"
        k = 3 * n
        body_off = np.zeros(k + 1, dtype=np.uint64)
        if k:
            np.cumsum(np.fromiter(map(len, vals), dtype=np.int64, count=k), out=body_off[1:])
        blob = b"".join(vals)
        arrays = {"n": k, "global_base": 0, "hdr": np.zeros(1, dtype=np.uint8), "hdr_off": np.zeros(k + 1, dtype=np.uint64),
                  "body": np.frombuffer(blob, dtype=np.uint8).copy() if blob else np.zeros(1, dtype=np.uint8), "body_off": body_off,
                  "ts": np.zeros(max(1, k), dtype=np.int64), "wall": np.zeros(max(1, k), dtype=np.int64), "flags8": np.zeros(max(1, k), dtype=np.uint64),
                  "fsb": np.zeros(max(1, k), dtype=np.uint32)}
        self.corpus = Corpus().load(arrays)
        self.n = n

    def close(self) -> None:
        self.corpus.close()

    def _field_hits(self, pattern: Pattern) -> np.ndarray:
        """bool[n, 3]: which of (Subject, Tags, content) of each block satisfies the pattern."""
        if self.n == 0:
            return np.zeros((0, 3), dtype=bool)
        pb = ProgramBuilder()
        pb.add_query([Cond(C_BODY, pattern=pattern)])
        return (self.corpus.scan_masks(pb.build()) & 1).astype(bool).reshape(self.n, 3)

    def search_memories(self, query: str, search_content: bool = True, search_subject: bool = True, search_tags: bool = True) -> List[Dict[str, Any]]:
        hits = self._field_hits(Pattern("contains", query.lower()))
        match = np.zeros(self.n, dtype=bool)
        if search_subject:
            match |= hits[:, 0]
        if search_tags:
            match |= hits[:, 1]
        if search_content:
            match |= hits[:, 2]
        match &= ~self.genesis
        return [self.memories[i] for i in np.nonzero(match)[0].tolist()]

    def search_by_tag(self, tag: str) -> List[Dict[str, Any]]:
        if tag.startswith("#"):
            tag = tag[1:]
        hits = self._field_hits(Pattern("has_tag", tag.lower()))
        return [self.memories[i] for i in np.nonzero(hits[:, 1] & ~self.genesis)[0].tolist()]
