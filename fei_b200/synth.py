"""Python face of the deterministic synthetic generators (fei_b200/csrc/synth.cuh, host build).

Record / block ``i`` is a pure function of ``(seed, i)``; the GPU generator
(``fei_corpus_synth``) produces byte-identical records, so fixtures made here can be
checked against corpora generated straight into HBM.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Any, Dict, List, Tuple

import numpy as np

from . import _abi

FOLDERS = ["", ".Projects/Python", ".Projects/AI", ".ToDoLater/Learning"]
STATUSES = ["cur", "new", "tmp"]
HOSTNAME = "b200node"
TASK_STATES = ["proposed", "accepted", "in_progress", "solution_proposed", "completed", "rejected"]
DIFFICULTIES = ["easy", "medium", "hard", "very_hard", "extreme"]
RESPONSIBLE = "3f0c9a52-7d41-4e8b-9c1d-5a6b7c8d9e0f"
PROPOSER = "b7e1d2c3-4a5f-4b6c-8d7e-0f1a2b3c4d5e"


def record(seed: int, i: int) -> Dict[str, Any]:
    l = _abi.lib()
    hl, bl = C.c_uint32(), C.c_uint32()
    hdr = np.zeros(512, dtype=np.uint8)
    body = np.zeros(16384, dtype=np.uint8)
    ts = C.c_int64(); uid = C.create_string_buffer(8); fl = C.create_string_buffer(4)
    nf, st, fo = C.c_uint8(), C.c_uint8(), C.c_uint8()
    _abi.check(l.fei_synth_record_host(seed, i, _abi.ptr(hdr), hdr.size, C.byref(hl), _abi.ptr(body), body.size, C.byref(bl),
                                       C.byref(ts), uid, fl, C.byref(nf), C.byref(st), C.byref(fo)))
    flags = fl.raw[:nf.value].decode()
    uid_s = uid.raw.decode()
    return {
        "hdr": bytes(hdr[:hl.value]), "body": bytes(body[:bl.value]),
        "ts": ts.value, "uid": uid_s, "flags": flags, "status": STATUSES[st.value], "folder": FOLDERS[fo.value],
        "status_id": st.value, "folder_id": fo.value,
        "filename": f"{ts.value}.{uid_s}.{HOSTNAME}:2,{flags}",
    }


def file_text(rec: Dict[str, Any]) -> str:
    """The on-disk file content the packed pieces came from (create_memory_content, utils.py:129-132)."""
    return rec["hdr"].decode() + "---\n" + rec["body"].decode()


def flags8(flags: str) -> int:
    b = flags.encode("ascii")
    if len(b) > 7:
        raise NotImplementedError("more than 7 flag letters")
    v = 0
    for k, ch in enumerate(b):
        v |= ch << (8 * k)
    return v | (len(b) << 56)


def corpus_arrays(seed: int, first: int, n: int) -> Dict[str, np.ndarray]:
    """Canonical host arrays (include/feiscan.h fei_corpus_host) for records [first, first+n)."""
    recs = [record(seed, first + k) for k in range(n)]
    return arrays_from_records(recs, global_base=first)


def arrays_from_records(recs: List[Dict[str, Any]], global_base: int = 0) -> Dict[str, Any]:
    n = len(recs)

    def blob(key):
        parts = [r[key] for r in recs]
        off = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum([len(p) for p in parts], out=off[1:])
        data = np.frombuffer(b"".join(parts), dtype=np.uint8).copy() if off[n] else np.zeros(1, dtype=np.uint8)
        return data, off

    hdr, hdr_off = blob("hdr")
    body, body_off = blob("body")
    parts = [r["filename"].encode("utf-8", "surrogateescape") for r in recs]
    name_off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum([len(p) for p in parts], out=name_off[1:])
    name = np.frombuffer(b"".join(parts), dtype=np.uint8).copy() if n else np.zeros(1, dtype=np.uint8)
    ts = np.array([r["ts"] for r in recs], dtype=np.int64)
    wall = np.array([r.get("wall", r["ts"]) for r in recs], dtype=np.int64)
    f8 = np.array([flags8(r["flags"]) for r in recs], dtype=np.uint64)
    fsb = np.array([(r["folder_id"] & 0xFFFF) | (r["status_id"] << 16) | (r.get("bits", 0) << 24) for r in recs], dtype=np.uint32)
    return {"n": n, "global_base": global_base, "hdr": hdr, "hdr_off": hdr_off, "body": body, "body_off": body_off,
            "name": name, "name_off": name_off, "ts": ts, "wall": wall, "flags8": f8, "fsb": fsb, "records": recs}


def write_memdir(base: str, recs: List[Dict[str, Any]]) -> None:
    """Materialise records as a Maildir-style tree the reference can read (utils.py:32-41, :134-151)."""
    for folder in FOLDERS + [".Trash", ".Archive"]:
        for st in STATUSES:
            os.makedirs(os.path.join(base, folder, st) if folder else os.path.join(base, st), exist_ok=True)
    for r in recs:
        d = os.path.join(base, r["folder"], r["status"]) if r["folder"] else os.path.join(base, r["status"])
        raw = r.get("raw")
        with open(os.path.join(d, r["filename"]), "wb") as f:
            f.write(raw if raw is not None else file_text(r).encode("utf-8"))


def write_memdir_native(base: str, seed: int, first: int, n: int, threads: int = 0) -> None:
    """The same tree as write_memdir([record(seed, first + k) for k in range(n)]), written by native threads
    (fei_synth_write_tree): a million files in seconds."""
    for folder in FOLDERS + [".Trash", ".Archive"]:
        for st in STATUSES:
            os.makedirs(os.path.join(base, folder, st) if folder else os.path.join(base, st), exist_ok=True)
    if not threads:
        threads = max(1, min(32, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)))
    _abi.check(_abi.lib().fei_synth_write_tree(os.fsencode(base), HOSTNAME.encode(), seed, first, n, threads))


def block(seed: int, i: int) -> Dict[str, Any]:
    l = _abi.lib()
    ts = C.c_double(); mid = C.create_string_buffer(8)
    tstate, diff, is_task = C.c_uint8(), C.c_uint8(), C.c_uint8()
    _abi.check(l.fei_synth_block_host(seed, i, C.byref(ts), mid, C.byref(tstate), C.byref(diff), C.byref(is_task)))
    md: Dict[str, Any] = {"metadata": {"unique_id": mid.raw.decode()}}
    if is_task.value:
        md.update({"type": "task", "task_state": TASK_STATES[tstate.value], "task_difficulty": DIFFICULTIES[diff.value]})
    return {"index": i, "timestamp": ts.value, "memory_data": md, "responsible_node": RESPONSIBLE, "proposer_node": PROPOSER}


def chain_specs(seed: int, first: int, n: int) -> List[Dict[str, Any]]:
    return [block(seed, first + k) for k in range(n)]
