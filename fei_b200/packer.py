"""Pack an on-disk Memdir into the canonical arrays of include/feiscan.h (the one-time ingest).

Replaces the per-query walk of utils.list_memories (memdir_tools/utils.py:202-253): files are read
once, in exactly the reference's listing order — folders in os.walk order (utils.py:48), statuses
cur/new/tmp, inside a directory os.listdir order stably sorted by filename timestamp, newest first
(utils.py:220,251) — decoded like `open(path, "r")` (UTF-8 strict, universal newlines; undecodable
files are reported and skipped, utils.py:247-248), split at the first '---' (utils.py:105), body
.strip()ped (utils.py:120).  Header *parsing* is left to the GPU (k_head); the host only parses the
headers of records it has to materialise as result dicts.
"""
from __future__ import annotations

import calendar
import os
import re
import threading
from datetime import datetime
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .memdir_tools import utils as U

REC_NO_SEPARATOR, REC_NONASCII, REC_HAS_SIGMA, REC_HAS_IDOT = 1, 2, 4, 8
_LIST_RE = re.compile(r"\d+\.[a-z0-9]+\.[^:]+:2,[A-Z]*")          # utils.py:223


def read_segment(base: str, folder: str, status: str) -> List[Dict[str, Any]]:
    path = os.path.join(base, folder, status) if folder else os.path.join(base, status)
    if not os.path.exists(path):
        return []
    out = []
    for name in os.listdir(path):
        try:
            if not _LIST_RE.match(name):
                continue
            m = U.FILENAME_RE.match(name)
            if not m:
                raise ValueError(f"Invalid memory filename: {name}")
            with open(os.path.join(path, name), "r") as f:
                text = f.read()
            head, sep, rest = text.partition("---")
            ts = int(m.group(1))
            out.append({
                "filename": name, "folder": folder, "status": status, "ts": ts,
                "uid": m.group(2), "host": m.group(3), "flags": m.group(4),
                "uid_span": (m.start(2), m.end(2)), "host_span": (m.start(3), m.end(3)),
                "hdr_text": head if sep else "", "body_text": (rest if sep else text).strip(), "has_sep": bool(sep),
                "date": datetime.fromtimestamp(ts),
            })
        except Exception as e:
            print(f"Error processing {name}: {e}")
    out.sort(key=lambda r: r["ts"], reverse=True)
    return out


def read_segment_raw(base: str, folder: str, status: str, file_cache: Optional[Dict[Tuple, bytes]] = None) -> List[Dict[str, Any]]:
    """Like read_segment but leaves the file *text* work (decode, newline folding, '---' split, strip) to the
    GPU ingest kernels (fei_corpus_load_raw): files are read as bytes; only the file-name grammar and the
    listing order are handled here.  `file_cache` maps (inode, size, mtime_ns) to content: a memory file is
    immutable, and moves / flag changes are renames (utils.py:255-297, :354-388), so an incremental re-pack
    re-reads only files it has never seen."""
    path = os.path.join(base, folder, status) if folder else os.path.join(base, status)
    if not os.path.exists(path):
        return []
    out = []
    for name in os.listdir(path):
        try:
            if not _LIST_RE.match(name):
                continue
            m = U.FILENAME_RE.match(name)
            if not m:
                raise ValueError(f"Invalid memory filename: {name}")
            full = os.path.join(path, name)
            raw = None
            if file_cache is not None:
                st = os.stat(full)
                key = (st.st_ino, st.st_size, st.st_mtime_ns)
                raw = file_cache.get(key)
            if raw is None:
                with open(full, "rb") as f:
                    raw = f.read()
                if file_cache is not None:
                    file_cache[key] = raw
            ts = int(m.group(1))
            out.append({"filename": name, "folder": folder, "status": status, "ts": ts, "uid": m.group(2), "host": m.group(3),
                        "flags": m.group(4), "uid_span": (m.start(2), m.end(2)), "host_span": (m.start(3), m.end(3)),
                        "raw": raw, "date": datetime.fromtimestamp(ts)})
        except Exception as e:
            print(f"Error processing {name}: {e}")
    out.sort(key=lambda r: r["ts"], reverse=True)
    return out


def _ensure_text(rec: Dict[str, Any]) -> None:
    """Host-side text of one record (hits only): what open(path, "r").read() + parse_memory_content see."""
    if "hdr_text" in rec:
        return
    text = rec["raw"].decode("utf-8").replace("\r\n", "\n").replace("\r", "\n")
    head, sep, rest = text.partition("---")
    rec["hdr_text"] = head if sep else ""
    rec["body_text"] = (rest if sep else text).strip()
    rec["has_sep"] = bool(sep)


def memory_dict(rec: Dict[str, Any], include_content: bool) -> Dict[str, Any]:
    """The dict list_memories yields (utils.py:234-243); headers parsed on the host for materialisation."""
    _ensure_text(rec)
    headers = rec.get("_headers")                       # parsed once per record; callers get their own copy (they may mutate it)
    if headers is None:
        headers = {}
        if rec["has_sep"]:
            for line in rec["hdr_text"].strip().split("\n"):
                k, colon, v = line.partition(":")
                if colon:
                    headers[k.strip()] = v.strip()
        rec["_headers"] = headers
    mem = {"filename": rec["filename"], "folder": rec["folder"], "status": rec["status"], "headers": dict(headers),
           "metadata": {"timestamp": rec["ts"], "unique_id": rec["uid"], "hostname": rec["host"], "flags": list(rec["flags"]),
                        "date": rec["date"]}}
    if include_content:
        mem["content"] = rec["body_text"]
    return mem


def _flags8(flags: str, name: str) -> int:
    if len(flags) > 7:
        raise NotImplementedError(f"{name}: more than 7 flag letters are not supported by the packed layout")
    v = len(flags) << 56
    for k, ch in enumerate(flags):
        v |= ord(ch) << (8 * k)
    return v


def arrays_from_segments(recs: Sequence[Dict[str, Any]], folder_ids: Dict[str, int], global_base: int = 0) -> Dict[str, Any]:
    n = len(recs)
    hdr_parts = [r["hdr_text"].encode("utf-8") for r in recs]
    body_parts = [r["body_text"].encode("utf-8") for r in recs]
    name_parts = [os.fsencode(r["filename"]) for r in recs]

    def blob(parts):
        off = np.zeros(n + 1, dtype=np.uint64)
        if n:
            np.cumsum(np.fromiter(map(len, parts), dtype=np.int64, count=n), out=off[1:])
        data = np.frombuffer(b"".join(parts), dtype=np.uint8).copy() if n and off[n] else np.zeros(1, dtype=np.uint8)
        return data, off

    hdr, hdr_off = blob(hdr_parts)
    body, body_off = blob(body_parts)
    name, name_off = blob(name_parts)
    spans = np.zeros((max(n, 1), 4), dtype=np.uint16)
    bits = np.zeros(max(n, 1), dtype=np.uint32)
    for i, r in enumerate(recs):
        fname = r["filename"]
        # spans are character offsets; convert to byte offsets when the name has non-ASCII characters
        (a0, a1), (b0, b1) = r["uid_span"], r["host_span"]
        if len(name_parts[i]) != len(fname):
            a0, a1, b0, b1 = (len(os.fsencode(fname[:x])) for x in (a0, a1, b0, b1))
        spans[i] = (a0, a1 - a0, b0, b1 - b0)
        b = 0 if r["has_sep"] else REC_NO_SEPARATOR
        text = r["hdr_text"] + r["body_text"]
        if not text.isascii():
            b |= REC_NONASCII
            if "Σ" in text:
                b |= REC_HAS_SIGMA
            if "İ" in text:
                b |= REC_HAS_IDOT
        bits[i] = b
    ts = np.array([r["ts"] for r in recs], dtype=np.int64)
    wall = np.array([calendar.timegm(r["date"].timetuple()) for r in recs], dtype=np.int64)
    f8 = np.array([_flags8(r["flags"], r["filename"]) for r in recs], dtype=np.uint64)
    fsb = np.array([(folder_ids[r["folder"]] & 0xFFFF) | (U.STANDARD_FOLDERS.index(r["status"]) << 16) | (int(bits[i]) << 24)
                    for i, r in enumerate(recs)], dtype=np.uint32)
    return {"n": n, "global_base": global_base, "hdr": hdr, "hdr_off": hdr_off, "body": body, "body_off": body_off,
            "name": name, "name_off": name_off, "name_spans": spans.reshape(-1), "ts": ts, "wall": wall, "flags8": f8, "fsb": fsb,
            "rec_bits": bits[:n].astype(np.uint8)}


def raw_arrays_from_segments(recs: Sequence[Dict[str, Any]], folder_ids: Dict[str, int], global_base: int = 0) -> Dict[str, Any]:
    """Host arrays for fei_corpus_load_raw: raw file bytes + meta columns + names (no text processing)."""
    n = len(recs)
    name_parts = [os.fsencode(r["filename"]) for r in recs]

    def blob(parts):
        off = np.zeros(n + 1, dtype=np.uint64)
        if n:
            np.cumsum(np.fromiter(map(len, parts), dtype=np.int64, count=n), out=off[1:])
        data = np.frombuffer(b"".join(parts), dtype=np.uint8).copy() if n and off[n] else np.zeros(1, dtype=np.uint8)
        return data, off

    raw, raw_off = blob([r["raw"] for r in recs])
    name, name_off = blob(name_parts)
    spans = np.zeros((max(n, 1), 4), dtype=np.uint16)
    for i, r in enumerate(recs):
        fname = r["filename"]
        (a0, a1), (b0, b1) = r["uid_span"], r["host_span"]
        if len(name_parts[i]) != len(fname):
            a0, a1, b0, b1 = (len(os.fsencode(fname[:x])) for x in (a0, a1, b0, b1))
        spans[i] = (a0, a1 - a0, b0, b1 - b0)
    ts = np.array([r["ts"] for r in recs], dtype=np.int64)
    wall = np.array([calendar.timegm(r["date"].timetuple()) for r in recs], dtype=np.int64)
    f8 = np.array([_flags8(r["flags"], r["filename"]) for r in recs], dtype=np.uint64)
    fsb = np.array([(folder_ids[r["folder"]] & 0xFFFF) | (U.STANDARD_FOLDERS.index(r["status"]) << 16) for r in recs], dtype=np.uint32)
    return {"n": n, "global_base": global_base, "raw": raw, "raw_off": raw_off, "name": name, "name_off": name_off,
            "name_spans": spans.reshape(-1), "ts": ts, "wall": wall, "flags8": f8, "fsb": fsb}


class PackedMemdir:
    """Host bookkeeping for one packed tree: records in listing order + the device corpus."""

    def __init__(self, base: str):
        self.base = base
        self.folders: List[str] = []
        self.segments: Dict[Tuple[str, str], Tuple[int, int]] = {}
        self.recs: List[Dict[str, Any]] = []
        self.arrays: Dict[str, Any] = {}
        self.corpus = None
        self.signature: Tuple = ()
        self.seg_cache: Dict[Tuple[str, str], Tuple[int, List[Dict[str, Any]]]] = {}   # (folder, status) -> (dir mtime, records)
        self.file_cache: Dict[Tuple, bytes] = {}
        self.files_read = 0          # files whose content was read from disk by the last build (incremental-sync telemetry)
        self.bad: Dict[Tuple[str, str], List[str]] = {}      # undecodable files per directory, reported on every listing
        self.lock = threading.RLock()                        # one query at a time plans aux columns / scans this corpus state
        self._field_values: Dict[str, Tuple[np.ndarray, np.ndarray, List[str]]] = {}
        self.parsed_dates: Dict[str, Tuple[Any, bool]] = {}  # header value -> (dateutil result | None, depends on today's date)
        self._aux_next = 0

    @staticmethod
    def tree_signature(base: str) -> Tuple:
        sig = []
        for root, dirs, _ in os.walk(base):
            for st in U.STANDARD_FOLDERS:
                if st in dirs:
                    p = os.path.join(root, st)
                    s = os.stat(p)
                    sig.append((p, s.st_mtime_ns))
        return tuple(sig)

    def build(self, upload: bool = True, gpu_text: Optional[bool] = None) -> "PackedMemdir":
        """gpu_text (default on; FEI_PACK_HOST_TEXT=1 turns it off): decode / newline folding / '---' split / strip run in the
        ingest kernels (fei_corpus_load_raw) instead of Python."""
        if gpu_text is None:
            gpu_text = os.environ.get("FEI_PACK_HOST_TEXT", "0") != "1"
        if gpu_text and upload:
            return self._build_raw()
        self.signature = self.tree_signature(self.base)
        self.folders = []
        for root, dirs, _ in os.walk(self.base):
            if any(st in dirs for st in U.STANDARD_FOLDERS):
                rel = os.path.relpath(root, self.base)
                self.folders.append("" if rel == "." else rel)
        if len(self.folders) > 65535:
            raise NotImplementedError("more than 65535 folders")
        self.folder_ids = {f: i for i, f in enumerate(self.folders)}
        self.recs = []
        self.segments = {}
        for folder in self.folders:
            for st in U.STANDARD_FOLDERS:
                seg = read_segment(self.base, folder, st)
                self.segments[(folder, st)] = (len(self.recs), len(self.recs) + len(seg))
                self.recs.extend(seg)
        self.arrays = arrays_from_segments(self.recs, self.folder_ids)
        if upload:
            from .corpus import Corpus
            self.corpus = Corpus().load(self.arrays)
            self._field_values = {}
        return self

    def _build_raw(self) -> "PackedMemdir":
        from .corpus import Corpus
        self.signature = self.tree_signature(self.base)
        self.folders = []
        for root, dirs, _ in os.walk(self.base):
            if any(st in dirs for st in U.STANDARD_FOLDERS):
                rel = os.path.relpath(root, self.base)
                self.folders.append("" if rel == "." else rel)
        if len(self.folders) > 65535:
            raise NotImplementedError("more than 65535 folders")
        self.folder_ids = {f: i for i, f in enumerate(self.folders)}
        segs: Dict[Tuple[str, str], List[Dict[str, Any]]] = {}
        self.bad = {}
        skipped_raw = set()
        before = len(self.file_cache)
        new_seg_cache = {}
        for folder in self.folders:
            for st in U.STANDARD_FOLDERS:
                path = os.path.join(self.base, folder, st) if folder else os.path.join(self.base, st)
                mt = os.stat(path).st_mtime_ns if os.path.exists(path) else -1
                cached = self.seg_cache.get((folder, st))
                if cached is not None and cached[0] == mt:
                    segs[(folder, st)] = list(cached[1])                 # directory untouched since the last pack
                else:
                    segs[(folder, st)] = read_segment_raw(self.base, folder, st, self.file_cache)
                new_seg_cache[(folder, st)] = (mt, segs[(folder, st)])
        self.files_read = len(self.file_cache) - before
        corpus = Corpus()
        while True:
            recs = [r for folder in self.folders for st in U.STANDARD_FOLDERS for r in segs[(folder, st)]]
            arrays = raw_arrays_from_segments(recs, self.folder_ids)
            valid = corpus.load_raw(arrays)
            if valid.all():
                break
            for r, ok in zip(recs, valid.tolist()):            # undecodable files: reported and skipped (utils.py:247-248)
                if not ok:
                    try:
                        r["raw"].decode("utf-8")
                        msg = "invalid UTF-8"
                    except UnicodeDecodeError as e:
                        msg = str(e)
                    self.bad.setdefault((r["folder"], r["status"]), []).append(f"Error processing {r['filename']}: {msg}")
                    skipped_raw.add(id(r["raw"]))
                    segs[(r["folder"], r["status"])].remove(r)
        self.recs = recs
        self.seg_cache = {k: (mt, list(segs[k])) for k, (mt, _r) in new_seg_cache.items()}
        live = {id(r["raw"]) for r in recs} | skipped_raw
        self.file_cache = {k: v for k, v in self.file_cache.items() if id(v) in live}   # forget deleted files
        self.segments = {}
        pos = 0
        for folder in self.folders:
            for st in U.STANDARD_FOLDERS:
                k = len(segs[(folder, st)])
                self.segments[(folder, st)] = (pos, pos + k)
                pos += k
        self.corpus = corpus
        self._field_values = {}
        fsb = corpus.fetch_meta()["fsb"] if corpus.n else np.zeros(0, dtype=np.uint32)
        arrays["rec_bits"] = (fsb >> 24).astype(np.uint8)
        self.arrays = arrays
        return self

    # ---- per-record header values for conditions only Python can judge (search.py:126-130)
    def header_values(self, field: str) -> Tuple[np.ndarray, np.ndarray, List[str]]:
        """(present[n], inv[n], distinct): the value _get_field_value would read for `field` (first key whose lower() equals
        field.lower(), last line of that exact key), record by record, as an index into the list of distinct values
        (len(distinct) for records without the header).  Extracted by the GPU once per packed corpus state."""
        key = field.lower()
        got = self._field_values.get(key)
        if got is None:
            from .program import C_SLOT, Cond, ProgramBuilder
            from .regexc import Pattern
            pb = ProgramBuilder()
            pb.add_query([Cond(C_SLOT, pattern=Pattern("regex", "", re.IGNORECASE), field=field, mode=0)])
            present, off, blob = self.corpus.slot_values(pb.build())
            n = self.corpus.n
            raw = blob.tobytes()
            index: Dict[bytes, int] = {}
            inv = np.empty(n, dtype=np.int64)
            o = off.astype(np.int64)
            for i in range(n):
                if not present[i]:
                    inv[i] = -1
                    continue
                inv[i] = index.setdefault(raw[o[i]:o[i + 1]], len(index))
            distinct = [b.decode("utf-8") for b in index]
            inv[inv < 0] = len(distinct)
            got = self._field_values[key] = (present, inv, distinct)
        return got

    def sigma_in(self, ranges: Sequence[Tuple[int, int]]) -> bool:
        """Does a record of these index ranges hold U+03A3 (the one character whose str.lower() the automata do not model)?"""
        bits = self.arrays.get("rec_bits")
        if bits is None or not len(bits):
            return False
        return any(bool((bits[a:b] & REC_HAS_SIGMA).any()) for a, b in ranges)

    def begin_query(self) -> None:
        self._aux_next = 0

    def new_aux(self, verdicts: np.ndarray) -> int:
        """Uploads one per-record verdict column for the query being compiled; returns its index (C_RECBITS.which)."""
        from .program import MAX_AUX
        if self._aux_next >= MAX_AUX:
            raise NotImplementedError(f"more than {MAX_AUX} host-judged header conditions in one query")
        k = self._aux_next
        self._aux_next += 1
        self.corpus.set_aux(k, verdicts)
        return k

    def report_skipped(self, folders: Optional[Sequence[str]], statuses: Optional[Sequence[str]]) -> None:
        """The reference prints `Error processing <file>: <error>` each time a directory is listed (utils.py:247-248)."""
        for f in (self.folders if folders is None else folders):
            for st in (U.STANDARD_FOLDERS if statuses is None else statuses):
                for msg in self.bad.get((f, st), []):
                    print(msg)

    def ranges(self, folders: Optional[Sequence[str]], statuses: Optional[Sequence[str]]) -> List[Tuple[int, int]]:
        """Index ranges of the requested (folder, status) pairs in the caller's order (search.py:361-363)."""
        if folders is None:
            folders = self.folders
        if statuses is None:
            statuses = U.STANDARD_FOLDERS
        out = []
        for f in folders:
            for st in statuses:
                if st not in U.STANDARD_FOLDERS:
                    raise ValueError(f"Invalid status: {st}. Must be one of {U.STANDARD_FOLDERS}")
                r = self.segments.get((f, st))
                if r and r[1] > r[0]:
                    out.append(r)
        return out


_cache: Dict[str, PackedMemdir] = {}


def packed(base: Optional[str] = None) -> PackedMemdir:
    """The packed corpus for a tree, rebuilt when any cur/new/tmp directory changed."""
    base = base or U.MEMDIR_BASE
    pm = _cache.get(base)
    if pm is None:
        pm = PackedMemdir(base).build()
        _cache[base] = pm
    elif pm.signature != PackedMemdir.tree_signature(base):
        # incremental sync: unchanged directories and already-seen files come from the host cache; the packed
        # arrays are rebuilt and re-uploaded (H2D + tiling are cheap next to file I/O)
        old = pm.corpus
        pm.build()
        if old is not None and old is not pm.corpus:
            old.close()
    return pm
