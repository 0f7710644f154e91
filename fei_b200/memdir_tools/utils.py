"""Host-side Memdir helpers with the reference's names and behaviour (memdir_tools/utils.py).

Only what the scan path and its callers need: locating the Maildir-style tree, the file-name
grammar, record splitting, and the rename-based mutations the filter actions perform.  The
per-record *matching* never happens here — records are packed (fei_b200.packer) and matched
on the GPU.
"""
from __future__ import annotations

import os
import re
import socket
import time
import uuid
from datetime import datetime
from typing import Any, Dict, List, Optional, Tuple

MEMDIR_BASE = os.path.join(os.getcwd(), "Memdir")           # bound at import, like the reference (utils.py:16)
STANDARD_FOLDERS = ["cur", "new", "tmp"]
SPECIAL_FOLDERS = [".Trash", ".ToDoLater", ".Projects", ".Archive"]
FLAGS = {"S": "Seen", "R": "Replied", "F": "Flagged", "P": "Priority"}

FILENAME_RE = re.compile(r"(\d+)\.([a-z0-9]+)\.([^:]+):2,([A-Z]*)")     # utils.py:81 (prefix match)


def set_memdir_base(path: str) -> None:
    global MEMDIR_BASE
    MEMDIR_BASE = path


def ensure_memdir_structure() -> None:
    for st in STANDARD_FOLDERS:
        os.makedirs(os.path.join(MEMDIR_BASE, st), exist_ok=True)
    for sp in SPECIAL_FOLDERS:
        for st in STANDARD_FOLDERS:
            os.makedirs(os.path.join(MEMDIR_BASE, sp, st), exist_ok=True)


def get_memdir_folders() -> List[str]:
    """Every directory that directly contains a cur/new/tmp child, in os.walk order (utils.py:43-57)."""
    found = []
    for root, dirs, _ in os.walk(MEMDIR_BASE):
        if any(st in dirs for st in STANDARD_FOLDERS):
            rel = os.path.relpath(root, MEMDIR_BASE)
            found.append("" if rel == "." else rel)
    return found


def generate_memory_filename(flags: str = "") -> str:
    keep = "".join(f for f in flags if f in FLAGS)
    return f"{int(time.time())}.{uuid.uuid4().hex[:8]}.{socket.gethostname()}:2,{keep}"


def parse_memory_filename(filename: str) -> Dict[str, Any]:
    m = FILENAME_RE.match(filename)
    if not m:
        raise ValueError(f"Invalid memory filename: {filename}")
    ts, uid, host, flags = m.groups()
    return {"timestamp": int(ts), "unique_id": uid, "hostname": host, "flags": list(flags),
            "date": datetime.fromtimestamp(int(ts))}


def parse_memory_content(content: str) -> Tuple[Dict[str, str], str]:
    head, sep, rest = content.partition("---")                # first '---' anywhere (utils.py:105)
    if not sep:
        return {}, content.strip()
    headers: Dict[str, str] = {}
    for line in head.strip().split("\n"):
        key, colon, value = line.partition(":")
        if colon:
            headers[key.strip()] = value.strip()
    return headers, rest.strip()


def create_memory_content(headers: Dict[str, str], body: str) -> str:
    return "\n".join(f"{k}: {v}" for k, v in headers.items()) + f"\n---\n{body}"


def get_memory_path(folder: str, status: str = "new") -> str:
    if status not in STANDARD_FOLDERS:
        raise ValueError(f"Invalid status: {status}. Must be one of {STANDARD_FOLDERS}")
    return os.path.join(MEMDIR_BASE, folder, status) if folder else os.path.join(MEMDIR_BASE, status)


def save_memory(folder: str, content: str, headers: Optional[Dict[str, str]] = None, flags: str = "") -> str:
    ensure_memdir_structure()
    tmp_dir = get_memory_path(folder, "tmp")
    os.makedirs(tmp_dir, exist_ok=True)
    filename = generate_memory_filename(flags)
    headers = {} if headers is None else headers
    headers.setdefault("Date", datetime.now().isoformat())
    headers.setdefault("Subject", f"Memory {filename.split('.')[1]}")
    tmp_path = os.path.join(tmp_dir, filename)
    with open(tmp_path, "w") as f:
        f.write(create_memory_content(headers, content))
    os.makedirs(get_memory_path(folder, "new"), exist_ok=True)
    os.rename(tmp_path, os.path.join(get_memory_path(folder, "new"), filename))      # atomic tmp -> new
    return filename


def move_memory(filename: str, source_folder: str, target_folder: str, source_status: str = "new",
                target_status: str = "cur", new_flags: Optional[str] = None) -> bool:
    src = os.path.join(get_memory_path(source_folder, source_status), filename)
    if not os.path.exists(src):
        return False
    if new_flags is not None:
        parse_memory_filename(filename)
        parts = filename.split(":2,")
        if len(parts) == 2:
            filename = f"{parts[0]}:2,{new_flags}"
    dst = os.path.join(get_memory_path(target_folder, target_status), filename)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    os.rename(src, dst)
    return True


def update_memory_flags(filename: str, folder: str, status: str, flags: str) -> bool:
    try:
        parse_memory_filename(filename)
    except ValueError:
        return False
    if len(filename.split(":2,")) != 2:
        return False
    return move_memory(filename, folder, folder, status, status, flags)


def list_memories(folder: str, status: str = "cur", include_content: bool = False) -> List[Dict[str, Any]]:
    """Directory listing in the reference's order (utils.py:202-253): host-side materialisation only."""
    from .. import packer
    seg = packer.read_segment(MEMDIR_BASE, folder, status)
    return [packer.memory_dict(r, include_content) for r in seg]


def search_memories(query: str, folders: List[str] = None, statuses: List[str] = None,
                    headers_only: bool = False) -> List[Dict[str, Any]]:
    """The reference's simple substring search (utils.py:299-352): `query.lower()` in any header value (lower-cased),
    else -- unless `headers_only` -- in the content.  One GPU pass: an "any header value" slot (one automaton run per value
    of the record's headers dict) and a content condition as two queries, unioned per (folder, status) in listing order.
    Hits lose `content` for a 100-character `content_preview`, exactly as there."""
    import numpy as np
    from .. import packer
    from ..program import C_BODY, C_SLOT, Cond, ProgramBuilder
    from ..regexc import Pattern
    low = query.lower()
    pm = packer.packed()
    with pm.lock:                                        # ranges, scan and materialisation from one corpus state
        ranges = pm.ranges(folders, statuses)
        pm.report_skipped(folders, statuses)
        pb = ProgramBuilder()
        pb.add_query([Cond(C_SLOT, pattern=Pattern("contains", low), field="", mode=2)])
        if not headers_only:
            pb.add_query([Cond(C_BODY, pattern=Pattern("contains", low))])
        nq = 1 if headers_only else 2
        per_query = pm.scan_hits(pb.build(), nq)
        hits = per_query[0] if nq == 1 else np.union1d(per_query[0], per_query[1])       # sorted = listing order inside a segment
        results = []
        for a, b in ranges:
            lo, hi = np.searchsorted(hits, a), np.searchsorted(hits, b)
            for memory in pm.materialize(hits[lo:hi], True):
                if not headers_only:
                    content = memory["content"]
                    memory["content_preview"] = content[:100] + "..." if len(content) > 100 else content
                    del memory["content"]
                results.append(memory)
    return results
