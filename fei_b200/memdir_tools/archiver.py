"""Criteria scans of the reference's MemoryArchiver (memdir_tools/archiver.py:113-181, :306-381) on the packed corpus.

`_memory_matches_criteria` is a per-record predicate over age, tags, header fields and flags; here a criteria dict becomes GPU
conditions (one query per criterion alternative, combined on the host: AND over the criteria keys, OR inside a tag list),
evaluated for every record in one pass.  `cleanup_memories` keeps the reference's statistics and file actions.  Archiving by
age, retention policies and trash emptying are file moves driven by the same age predicate and stay with the reference."""
from __future__ import annotations

import calendar
import os
import re
from datetime import datetime, timedelta
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from .. import packer
from ..program import C_CONST, C_DATE_CMP, C_FLAGS, C_SLOT, CMP, Cond, ProgramBuilder, const
from ..regexc import Pattern
from . import utils as U


def _micros(dt: datetime) -> int:
    return calendar.timegm(dt.timetuple()) * 1000000 + dt.microsecond


def compile_criteria(criteria: Dict[str, Any], now: datetime) -> List[List[List[Cond]]]:
    """criteria -> [per key: [alternative: [conditions]]]; a record matches when every key has a matching alternative."""
    out: List[List[List[Cond]]] = []
    for key, pattern in criteria.items():
        if key in ("age", "min_age", "max_age"):
            if isinstance(pattern, bool) or not isinstance(pattern, int):
                raise NotImplementedError(f"criteria {key!r}: only integer day counts are supported on the GPU")
            # (now - date).days < p  <=>  date > now - p days;  .days > p  <=>  date <= now - (p + 1) days   (timedelta.days floors)
            if key in ("age", "min_age"):
                out.append([[Cond(C_DATE_CMP, op=CMP["<="], i64=_micros(now - timedelta(days=pattern)))]])
            else:
                out.append([[Cond(C_DATE_CMP, op=CMP[">"], i64=_micros(now - timedelta(days=pattern + 1)))]])
        elif key in ("tag", "tags"):
            items = pattern if isinstance(pattern, list) else [pattern.lower()]
            alts = []
            for item in items:
                if not isinstance(item, str):
                    continue                                   # `item in memory_tags`: a non-str never equals a tag
                alts.append([Cond(C_SLOT, pattern=Pattern("has_tag", item), field="Tags", mode=1, empty_if_missing=True)])
            out.append(alts or [[const(False)]])
        else:
            # a header with exactly this key decides; only without one does `flags` mean the file-name flags (archiver.py:158-176)
            if isinstance(pattern, str):
                hdr = Cond(C_SLOT, pattern=Pattern("regex", pattern, re.IGNORECASE), field=key, mode=1, if_missing=2 if key == "flags" else 0)
                out.append([[hdr, Cond(C_FLAGS, pattern=Pattern("regex", pattern, re.IGNORECASE))] if key == "flags" else [hdr]])
            else:
                out.append([[const(False)]])                   # a header value (str) never equals a non-str, nor does the flags string
    return out


class MemoryArchiver:
    def __init__(self):
        self.archive_age = 90
        self.trash_age = 30
        self.archive_folder = ".Archive"
        self.trash_folder = ".Trash"
        self.archive_rules: List[Dict[str, Any]] = []
        self.cleanup_rules: List[Dict[str, Any]] = []
        self.retention_policies: Dict[str, Any] = {}
        self.tag_based_archiving: Dict[str, Any] = {}

    def add_cleanup_rule(self, criteria: Dict[str, Any], action: str = "trash") -> None:
        self.cleanup_rules.append({"criteria": criteria, "action": action})

    # ---- one GPU pass for any number of criteria dicts
    def match_criteria(self, criteria_list: Sequence[Dict[str, Any]], now: Optional[datetime] = None) -> np.ndarray:
        """bool[len(criteria_list), n]: which packed records match which criteria dict."""
        now = now or datetime.now()
        pm = packer.packed()
        with pm.lock:
            plans = [compile_criteria(c, now) for c in criteria_list]
            queries: List[List[Cond]] = []
            for plan in plans:
                for alts in plan:
                    queries.extend(alts)
            out = np.ones((len(plans), pm.n), dtype=bool)
            masks = np.zeros((0, pm.n), dtype=np.uint32)
            for lo in range(0, len(queries), 32):
                pb = ProgramBuilder()
                for q in queries[lo:lo + 32]:
                    pb.add_query(q)
                masks = np.vstack([masks, pm.scan_masks(pb.build())[None, :]])
            qi = 0
            for r, plan in enumerate(plans):
                for alts in plan:
                    hit = np.zeros(pm.n, dtype=bool)
                    for _ in alts:
                        hit |= (masks[qi // 32] >> np.uint32(qi % 32)) & np.uint32(1) != 0
                        qi += 1
                    out[r] &= hit
        return out

    def _memory_matches_criteria(self, memory: Dict[str, Any], criteria: Dict[str, Any]) -> bool:
        """Single-record form (host): the same rules as compile_criteria, evaluated directly."""
        for key, pattern in criteria.items():
            if key in ("age", "min_age", "max_age"):
                age = (datetime.now() - memory["metadata"]["date"]).days
                if (key != "max_age" and age < pattern) or (key == "max_age" and age > pattern):
                    return False
            elif key in ("tag", "tags"):
                tags = [t.strip() for t in memory["headers"].get("Tags", "").lower().split(",")]
                wanted = pattern if isinstance(pattern, list) else [pattern.lower()]
                if not any(w in tags for w in wanted):
                    return False
            elif key in memory["headers"] or key == "flags":
                value = memory["headers"][key] if key in memory["headers"] else "".join(memory["metadata"]["flags"])
                if not (re.search(pattern, value, re.IGNORECASE) if isinstance(pattern, str) else value == pattern):
                    return False
            else:
                return False
        return True

    def cleanup_memories(self, dry_run: bool = False) -> Dict[str, Any]:
        """archiver.py:306-381: every rule over every folder outside the trash, status cur only."""
        stats: Dict[str, Any] = {"trashed": 0, "deleted": 0, "details": []}
        if not self.cleanup_rules:
            self.add_cleanup_rule({"status": "completed|done"}, "trash")
            self.add_cleanup_rule({"status": "obsolete|deprecated"}, "trash")
        pm = packer.packed()
        matched = self.match_criteria([r["criteria"] for r in self.cleanup_rules])
        gone = np.zeros(pm.n, dtype=bool)               # acted upon by an earlier rule (real runs re-list after every rule)
        for rule, row in zip(self.cleanup_rules, matched):
            action = rule["action"]
            for folder in U.get_memdir_folders():
                if folder.startswith(self.trash_folder):
                    continue
                pm.report_skipped([folder], ["cur"])
                lo, hi = pm.segments.get((folder, "cur"), (0, 0))
                for i in (np.nonzero(row[lo:hi] & ~gone[lo:hi])[0] + lo).tolist():
                    memory = pm.materialize([i], True)[0]
                    if action == "trash":
                        if not dry_run:
                            U.move_memory(memory["filename"], folder, self.trash_folder, "cur", "cur")
                            gone[i] = True
                        stats["trashed"] += 1
                        stats["details"].append({"memory_id": memory["metadata"]["unique_id"], "subject": memory["headers"].get("Subject", "No subject"),
                                                 "action": "Moved to trash" if not dry_run else "Would move to trash"})
                    elif action == "delete":
                        if not dry_run:
                            path = os.path.join(U.MEMDIR_BASE, folder, "cur", memory["filename"])
                            if os.path.exists(path):
                                os.remove(path)
                            gone[i] = True
                        stats["deleted"] += 1
                        stats["details"].append({"memory_id": memory["metadata"]["unique_id"], "subject": memory["headers"].get("Subject", "No subject"),
                                                 "action": "Deleted permanently" if not dry_run else "Would delete permanently"})
        return stats
