"""TEST INFRASTRUCTURE — CPU restatement of the reference's Memdir search / filter path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; it is the checker for the CUDA scan (fei_b200/csrc/scan.cu), never the
thing measured or shipped.

Each function follows /root/reference/memdir_tools:
  split_filename     <- utils.parse_memory_filename                (utils.py:74-95)
  split_content      <- utils.parse_memory_content                 (utils.py:97-120)
  memdir_folders     <- utils.get_memdir_folders                   (utils.py:43-57)
  read_folder        <- utils.list_memories                        (utils.py:202-253)
  lookup             <- search._get_field_value                    (search.py:97-139)
  holds              <- search._compare_values                     (search.py:141-242)
  record_matches     <- search._memory_matches_query               (search.py:244-335)
  run_search         <- search.search_memories                     (search.py:337-390)
  filter_accepts     <- filter.MemoryFilter.matches                (filter.py:67-109)
  run_filters        <- filter.FilterManager.process_memories, dry run   (filter.py:188-261)
  legacy_search      <- utils.search_memories (substring search)   (utils.py:299-352)
  folder_stats       <- folders.MemdirFolderManager.get_folder_stats   (folders.py:216-318)
  matches_criteria   <- archiver.MemoryArchiver._memory_matches_criteria   (archiver.py:128-181)
`re`, `datetime` are CPython stdlib; `dateutil` (unpinned third-party, 2.9.0.post0 here) is what the
reference imports (search.py:12).  Pinned by tests/golden/memdir_golden.json, produced by running the
unmodified reference on the same inputs (tests/golden/make_golden_memdir.py).
"""
from __future__ import annotations

import os
import re
from datetime import datetime, timedelta
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import dateutil.parser

_NAME_RE = re.compile(r"(\d+)\.([a-z0-9]+)\.([^:]+):2,([A-Z]*)")
STATUS_DIRS = ("cur", "new", "tmp")
_DATE_HEADERS = ("date", "due", "created", "modified", "deleteddate")


# ----------------------------------------------------------------------------- records
def split_filename(name: str) -> Dict[str, Any]:
    m = _NAME_RE.match(name)
    if m is None:
        raise ValueError(f"Invalid memory filename: {name}")
    ts, uid, host, flags = m.groups()
    return {"timestamp": int(ts), "unique_id": uid, "hostname": host, "flags": list(flags),
            "date": datetime.fromtimestamp(int(ts))}


def split_content(text: str) -> Tuple[Dict[str, str], str]:
    head, sep, rest = text.partition("---")
    if not sep:
        return {}, text.strip()
    headers: Dict[str, str] = {}
    for line in head.strip().split("\n"):
        k, colon, v = line.partition(":")
        if colon:
            headers[k.strip()] = v.strip()
    return headers, rest.strip()


def make_memory(filename: str, folder: str, status: str, text: str, include_content: bool) -> Dict[str, Any]:
    headers, body = split_content(text)
    mem = {"filename": filename, "folder": folder, "status": status, "headers": headers, "metadata": split_filename(filename)}
    if include_content:
        mem["content"] = body
    return mem


def memdir_folders(base: str) -> List[str]:
    out = []
    for root, dirs, _files in os.walk(base):
        if any(d in dirs for d in STATUS_DIRS):
            rel = os.path.relpath(root, base)
            out.append("" if rel == "." else rel)
    return out


def read_folder(base: str, folder: str, status: str, include_content: bool) -> List[Dict[str, Any]]:
    path = os.path.join(base, folder, status) if folder else os.path.join(base, status)
    if not os.path.exists(path):
        return []
    found = []
    for name in os.listdir(path):
        try:
            if not re.match(r"\d+\.[a-z0-9]+\.[^:]+:2,[A-Z]*", name):
                continue
            with open(os.path.join(path, name), "r") as f:
                text = f.read()
            found.append(make_memory(name, folder, status, text, include_content))
        except Exception as e:                      # unreadable / undecodable file: reported and skipped
            print(f"Error processing {name}: {e}")
    found.sort(key=lambda m: m["metadata"]["timestamp"], reverse=True)
    return found


# ----------------------------------------------------------------------------- search
def lookup(mem: Dict[str, Any], field: str) -> Any:
    low = field.lower()
    meta = mem["metadata"]
    direct = {"content": lambda: mem.get("content", ""), "flags": lambda: "".join(meta["flags"]),
              "date": lambda: meta["date"], "id": lambda: meta["unique_id"], "filename": lambda: mem["filename"],
              "folder": lambda: mem["folder"], "status": lambda: mem["status"], "maildir_status": lambda: mem["status"]}
    if low in direct:
        return direct[low]()
    if field == "Status" or low in ("status_value", "state"):
        return mem["headers"].get("Status", "")
    for key, val in mem["headers"].items():
        if key.lower() == low:
            if low in _DATE_HEADERS:
                try:
                    return dateutil.parser.parse(val)
                except (ValueError, TypeError):
                    return val
            return val
    for key, val in meta.items():
        if key.lower() == low:
            return val
    return None


def _as_datetime_operand(v1: Any, v2: Any):
    """(ok, v2') — a datetime field makes the operand go through dateutil; failure poisons the test."""
    if isinstance(v1, datetime) and not isinstance(v2, datetime):
        try:
            return True, dateutil.parser.parse(str(v2))
        except (ValueError, TypeError):
            return False, v2
    return True, v2


def _relative_now(v2: Any) -> Any:
    if not (isinstance(v2, str) and v2.startswith("now")):
        return v2
    now = datetime.now()
    if v2 == "now":
        return now
    m = re.match(r"now([+-])(\d+)([dwmy])", v2)
    if not m:
        return v2
    sign, num, unit = m.groups()
    n = int(num) * (-1 if sign == "-" else 1)
    days = {"d": n, "w": 7 * n, "m": 30 * n, "y": 365 * n}[unit]
    return now + timedelta(days=days)


def holds(v1: Any, op: str, v2: Any) -> bool:
    if v1 is None:
        return False
    if op == "contains":
        return str(v2).lower() in str(v1).lower()
    if op == "matches":
        try:
            return bool(re.search(str(v2), str(v1), re.IGNORECASE))
        except re.error:
            return False
    if op == "startswith":
        return str(v1).lower().startswith(str(v2).lower())
    if op == "endswith":
        return str(v1).lower().endswith(str(v2).lower())
    if op == "has_tag":
        return str(v2).lower() in [t.strip() for t in str(v1).lower().split(",")]
    if op in ("=", "!="):
        ok, v2 = _as_datetime_operand(v1, v2)
        if not ok:
            return False
        if isinstance(v1, str) and isinstance(v2, str):
            same = v1.lower() == v2.lower()
        else:
            same = v1 == v2
        return same if op == "=" else not same
    if op in (">", "<", ">=", "<="):
        ok, v2 = _as_datetime_operand(v1, v2)
        if not ok:
            return False
        v2 = _relative_now(v2)
        if op == ">":
            return v1 > v2
        if op == "<":
            return v1 < v2
        if op == ">=":
            return v1 >= v2
        return v1 <= v2
    if op == "has_flag":
        return str(v2).upper() in str(v1)
    return False


def record_matches(mem: Dict[str, Any], conditions: Sequence[Dict[str, Any]]) -> bool:
    if not conditions:
        return True
    # "keyword" conditions (Subject contains X paired with content contains X) are tested first,
    # everything else afterwards in list order; the verdict is the AND of all of them.
    def is_keyword(c):
        return c["field"] == "Subject" and c["operator"] == "contains" and any(
            o["field"] == "content" and o["operator"] == "contains" and o["value"] == c["value"] for o in conditions)
    first = [c for c in conditions if is_keyword(c)]
    rest = [c for c in conditions if not is_keyword(c)]
    for c in first:
        if not holds(lookup(mem, c["field"]), c["operator"], c["value"]):
            return False
    for c in rest:
        f = c["field"]
        val = mem["headers"].get("Status", "") if f in ("Status", "status_value", "state") else lookup(mem, f)
        if not holds(val, c["operator"], c["value"]):
            return False
    return True


def run_search(memories: Iterable[Dict[str, Any]], conditions: Sequence[Dict[str, Any]]) -> List[int]:
    """Indices (listing order) of the matching memories."""
    return [i for i, m in enumerate(memories) if record_matches(m, conditions)]


def listing(base: str, folders: Optional[List[str]], statuses: Optional[List[str]], include_content: bool) -> List[Dict[str, Any]]:
    """All memories in the reference's listing order (search.py:353-367)."""
    if folders is None:
        folders = memdir_folders(base)
    if statuses is None:
        statuses = list(STATUS_DIRS)
    out: List[Dict[str, Any]] = []
    for folder in folders:
        for status in statuses:
            out.extend(read_folder(base, folder, status, include_content))
    return out


# ----------------------------------------------------------------------------- filters
def filter_accepts(mem: Dict[str, Any], conditions: Sequence[Dict[str, Any]]) -> bool:
    for c in conditions:
        field, pattern, negate = c["field"], c["pattern"], c["negate"]
        if field == "content":
            value = mem.get("content", "")
        elif field in mem["headers"]:
            value = mem["headers"][field]
        elif field == "flags":
            value = "".join(mem["metadata"]["flags"])
        elif field in mem["metadata"]:
            value = str(mem["metadata"][field])
        else:
            value = None
        if value is None:
            if negate:
                continue
            return False
        hit = re.search(pattern, value, re.IGNORECASE) is not None
        if hit == bool(negate):
            return False
    return True


def run_filters(memories: Sequence[Dict[str, Any]], filters: Sequence[Dict[str, Any]]) -> Dict[str, Any]:
    """Dry-run statistics of FilterManager.process_memories (filters: {name, conditions, actions})."""
    stats = {"total_memories": len(memories), "filters_applied": 0, "actions_taken": 0, "memories_modified": 0, "details": []}
    touched = set()
    for mem in memories:
        applied = []
        for flt in filters:
            if filter_accepts(mem, flt["conditions"]):
                stats["filters_applied"] += 1
                acts = [f"Would {a['type']}" for a in flt["actions"]]
                stats["actions_taken"] += len(acts)
                if acts:
                    touched.add(mem["metadata"]["unique_id"])
                    applied.append({"filter": flt["name"], "actions": acts})
        if applied:
            stats["details"].append({"memory_id": mem["metadata"]["unique_id"],
                                     "subject": mem["headers"].get("Subject", "No subject"), "filters_applied": applied})
    stats["memories_modified"] = len(touched)
    return stats


def legacy_search(memories: Sequence[Dict[str, Any]], query: str, headers_only: bool = False) -> List[Dict[str, Any]]:
    """memdir_tools.utils.search_memories (utils.py:299-352) on an already listed sequence (include_content=True):
    case-insensitive substring in any header value, else in the content; hits lose `content` for a 100-character preview."""
    out = []
    for memory in memories:
        memory = dict(memory)
        found = False
        for _key, value in memory["headers"].items():
            if query.lower() in value.lower():
                found = True
                break
        if not found and not headers_only and "content" in memory:
            if query.lower() in memory["content"].lower():
                found = True
        if found:
            if "content" in memory and not headers_only:
                c = memory["content"]
                memory["content_preview"] = c[:100] + "..." if len(c) > 100 else c
                del memory["content"]
            out.append(memory)
    return out


def folder_stats(base: str, folder_path: str, include_subfolders: bool = False) -> Dict[str, Any]:
    """MemdirFolderManager.get_folder_stats (folders.py:216-318) on the tree at `base`."""
    folder_path = folder_path.replace("\\", "/").strip("/")
    all_folders = memdir_folders(base)
    todo = []
    if include_subfolders:
        for folder in all_folders:
            if folder == folder_path or (folder.startswith(folder_path) and folder != folder_path):
                todo.append(folder)
    elif folder_path in all_folders or folder_path == "":
        todo = [folder_path]
    stats: Dict[str, Any] = {"folder": folder_path or "Inbox", "total_memories": 0, "memory_counts": {"cur": 0, "new": 0, "tmp": 0},
                             "flag_counts": {"S": 0, "R": 0, "F": 0, "P": 0}, "tags": {}, "subfolders": [], "newest_memory": None, "oldest_memory": None}
    for folder in todo:
        sub = {"folder": folder or "Inbox", "memory_counts": {"cur": 0, "new": 0, "tmp": 0}, "total_memories": 0}
        for status in ("cur", "new", "tmp"):
            memories = read_folder(base, folder, status, False)
            sub["memory_counts"][status] = len(memories); sub["total_memories"] += len(memories)
            stats["total_memories"] += len(memories); stats["memory_counts"][status] += len(memories)
            for memory in memories:
                for flag in memory["metadata"]["flags"]:
                    if flag in stats["flag_counts"]:
                        stats["flag_counts"][flag] += 1
                if "Tags" in memory["headers"]:
                    for tag in [t.strip() for t in memory["headers"]["Tags"].split(",")]:
                        stats["tags"][tag] = stats["tags"].get(tag, 0) + 1
                d = memory["metadata"]["date"]
                if stats["newest_memory"] is None or d > stats["newest_memory"]["date"]:
                    stats["newest_memory"] = {"id": memory["metadata"]["unique_id"], "subject": memory["headers"].get("Subject", "No subject"), "date": d}
                if stats["oldest_memory"] is None or d < stats["oldest_memory"]["date"]:
                    stats["oldest_memory"] = {"id": memory["metadata"]["unique_id"], "subject": memory["headers"].get("Subject", "No subject"), "date": d}
        if include_subfolders and folder != folder_path:
            stats["subfolders"].append(sub)
    return stats


def matches_criteria(memory: Dict[str, Any], criteria: Dict[str, Any], now=None) -> bool:
    """MemoryArchiver._memory_matches_criteria (archiver.py:128-181), restated: every criterion must hold.
    Order of interpretation per key: age family, tag family, a header of exactly that name, the file-name flags, else fail.
    `now` pins datetime.now() for tests."""
    from datetime import datetime
    now = now or datetime.now()
    headers, meta = memory["headers"], memory["metadata"]

    def text_rule(value: str, pattern: Any) -> bool:
        return bool(re.search(pattern, value, re.IGNORECASE)) if isinstance(pattern, str) else value == pattern

    for key, pattern in criteria.items():
        if key in ("age", "min_age", "max_age"):
            days = (now - meta["date"]).days
            ok = days <= pattern if key == "max_age" else days >= pattern
        elif key in ("tag", "tags"):
            have = [piece.strip() for piece in headers.get("Tags", "").lower().split(",")]
            ok = any(t in have for t in pattern) if isinstance(pattern, list) else pattern.lower() in have
        elif key in headers:
            ok = text_rule(headers[key], pattern)
        elif key == "flags":
            ok = text_rule("".join(meta["flags"]), pattern)
        else:
            ok = False
        if not ok:
            return False
    return True
