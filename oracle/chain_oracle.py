"""TEST INFRASTRUCTURE — CPU restatement of the reference's Memorychain link-hash path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; it is the checker for the CUDA path, never the thing measured or shipped.

Each function follows /root/reference/memdir_tools/memorychain.py:
  hashed_fields  <- the dict literal in MemoryBlock.calculate_hash            (:117-128)
  block_text     <- json.dumps(<that dict>, sort_keys=True)                   (:117-128)
  block_hash     <- hashlib.sha256(block_string.encode()).hexdigest()         (:130)
  validate       <- MemoryChain.validate_chain, first failure + which check   (:596-618)
json / hashlib are CPython stdlib, exactly what the reference calls.  Pinned by
tests/golden/chain_kats.json (produced by importing the reference, tests/golden/make_golden.py).
The plain-C twin (chain_oracle.c -> libchain_oracle.so) is exposed through `c_*` helpers.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import json
import os
from typing import Any, Dict, List, Optional, Sequence, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
C_LIB_PATH = os.path.join(_HERE, "libchain_oracle.so")


class Block:
    """Minimal stand-in for the reference's MemoryBlock attributes that validation reads."""

    __slots__ = ("index", "timestamp", "memory_data", "previous_hash", "responsible_node", "proposer_node",
                 "nonce", "task_state", "difficulty", "solver_node", "hash")

    def __init__(self, index, timestamp, memory_data, previous_hash, responsible_node, proposer_node,
                 nonce=0, hash=None):
        self.index = index
        self.timestamp = timestamp
        self.memory_data = memory_data
        self.previous_hash = previous_hash
        self.responsible_node = responsible_node
        self.proposer_node = proposer_node
        self.nonce = nonce
        self.difficulty = memory_data.get("task_difficulty", "medium")     # reference :102
        self.task_state = memory_data.get("task_state", "proposed")        # reference :104
        self.solver_node = None                                            # reference :105
        self.hash = block_hash(self) if hash is None else hash


def hashed_fields(b: Any) -> Dict[str, Any]:
    return {
        "index": b.index,
        "timestamp": b.timestamp,
        "memory_id": b.memory_data.get("metadata", {}).get("unique_id", ""),
        "previous_hash": b.previous_hash,
        "responsible_node": b.responsible_node,
        "proposer_node": b.proposer_node,
        "task_state": getattr(b, "task_state", None),
        "difficulty": getattr(b, "difficulty", None),
        "solver_node": getattr(b, "solver_node", None),
        "nonce": b.nonce,
    }


def block_text(b: Any) -> str:
    return json.dumps(hashed_fields(b), sort_keys=True)


def block_hash(b: Any) -> str:
    return hashlib.sha256(block_text(b).encode()).hexdigest()


def validate(chain: Sequence[Any]) -> Tuple[bool, int, int]:
    """(ok, first_bad, kind): kind 1 = invalid hash, 2 = broken link; (True, -1, 0) when valid."""
    for i in range(1, len(chain)):
        cur, prev = chain[i], chain[i - 1]
        if cur.hash != block_hash(cur):
            return False, i, 1
        if cur.previous_hash != prev.hash:
            return False, i, 2
    return True, -1, 0


def mine(b: Any, difficulty: int = 2) -> None:
    """MemoryBlock.mine_block (memorychain.py:132-143): bump the nonce until the hash has `difficulty` leading zeros."""
    target = "0" * difficulty
    while b.hash[:difficulty] != target:
        b.nonce += 1
        b.hash = block_hash(b)


def build_chain(specs: Sequence[Dict[str, Any]]) -> List[Block]:
    """Link a list of {index,timestamp,memory_data,responsible_node,proposer_node[,nonce]} dicts."""
    out: List[Block] = []
    prev = "0"
    for s in specs:
        b = Block(s["index"], s["timestamp"], s["memory_data"], prev, s["responsible_node"], s["proposer_node"],
                  nonce=s.get("nonce", 0))
        out.append(b)
        prev = b.hash
    return out


# --------------------------------------------------------------------------- C twin
class _CVal(C.Structure):
    _fields_ = [("kind", C.c_int), ("s", C.c_char_p), ("slen", C.c_uint64), ("i", C.c_int64), ("f", C.c_double)]


_clib: Optional[C.CDLL] = None


def c_lib() -> C.CDLL:
    global _clib
    if _clib is None:
        l = C.CDLL(C_LIB_PATH)
        l.co_sha256_hex.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p]
        l.co_block_json.restype = C.c_void_p
        l.co_block_json.argtypes = [C.POINTER(_CVal), C.POINTER(C.c_uint64)]
        l.co_free.argtypes = [C.c_void_p]
        l.co_validate.restype = C.c_int
        l.co_validate.argtypes = [C.POINTER(_CVal), C.POINTER(C.c_char_p), C.POINTER(C.c_uint64), C.c_uint64,
                                  C.POINTER(C.c_int64), C.POINTER(C.c_int)]
        _clib = l
    return _clib


_ORDER = ("difficulty", "index", "memory_id", "nonce", "previous_hash", "proposer_node", "responsible_node",
          "solver_node", "task_state", "timestamp")


def _cvals(b: Any, keep: list) -> List[_CVal]:
    f = hashed_fields(b)
    out = []
    for k in _ORDER:
        v = f[k]
        cv = _CVal()
        if v is None:
            cv.kind = 0
        elif v is True:
            cv.kind = 4
        elif v is False:
            cv.kind = 5
        elif isinstance(v, str):
            raw = v.encode("utf-8", "surrogatepass")
            keep.append(raw)
            cv.kind, cv.s, cv.slen = 1, raw, len(raw)
        elif isinstance(v, int):
            cv.kind, cv.i = 2, v
        elif isinstance(v, float):
            cv.kind, cv.f = 3, v
        else:
            raise TypeError(f"unsupported field type for the C oracle: {type(v)}")
        out.append(cv)
    return out


def c_sha256_hex(data: bytes) -> str:
    out = C.create_string_buffer(65)
    c_lib().co_sha256_hex(data, len(data), out)
    return out.value.decode()


def c_block_text(b: Any) -> bytes:
    keep: list = []
    vals = (_CVal * 10)(*_cvals(b, keep))
    n = C.c_uint64()
    p = c_lib().co_block_json(vals, C.byref(n))
    try:
        return C.string_at(p, n.value)
    finally:
        c_lib().co_free(p)


def c_validate(chain: Sequence[Any]) -> Tuple[bool, int, int]:
    keep: list = []
    flat: List[_CVal] = []
    for b in chain:
        flat.extend(_cvals(b, keep))
    vals = (_CVal * len(flat))(*flat)
    hashes = [b.hash.encode() for b in chain]
    harr = (C.c_char_p * len(chain))(*hashes)
    hlen = (C.c_uint64 * len(chain))(*[len(h) for h in hashes])
    fb, kind = C.c_int64(), C.c_int()
    ok = c_lib().co_validate(vals, harr, hlen, len(chain), C.byref(fb), C.byref(kind))
    return bool(ok), fb.value, kind.value


# --------------------------------------------------------------------------- chain-memory search ("next" row 4)
def search_chain_memories(chain: Sequence[Any], query: str, search_content: bool = True, search_subject: bool = True,
                          search_tags: bool = True) -> List[int]:
    """Indices of the blocks MemorychainConnector.search_memories returns (fei/tools/memorychain_connector.py:273-324)."""
    q = query.lower()
    out = []
    for i, block in enumerate(chain):
        memory = block["memory_data"] if isinstance(block, dict) else block.memory_data
        if memory.get("metadata", {}).get("unique_id", "") == "genesis":          # :304-305
            continue
        headers = memory.get("headers", {})
        found = search_subject and q in headers.get("Subject", "").lower()         # :310
        found = found or (search_tags and q in headers.get("Tags", "").lower())    # :314
        found = found or (search_content and q in memory.get("content", "").lower())   # :318
        if found:
            out.append(i)
    return out


def search_chain_by_tag(chain: Sequence[Any], tag: str) -> List[int]:
    """Indices of the blocks MemorychainConnector.search_by_tag returns (memorychain_connector.py:326-362)."""
    if tag.startswith("#"):
        tag = tag[1:]
    tag = tag.lower()
    out = []
    for i, block in enumerate(chain):
        memory = block["memory_data"] if isinstance(block, dict) else block.memory_data
        tags = [t.strip() for t in memory.get("headers", {}).get("Tags", "").lower().split(",")]
        if memory.get("metadata", {}).get("unique_id", "") == "genesis":
            continue
        if tag in tags:
            out.append(i)
    return out
