"""memdir_tools.filter on the GPU: same public names and behaviour as the reference
(memdir_tools/filter.py): ``MemoryFilter``, ``FilterManager``, ``create_default_filters``,
``run_filters`` — plus ``apply_filters``, the name BASELINE.json's north_star uses (absent in the
reference; it returns exactly ``process_memories``' statistics).

The memory x filter match matrix (filter.py:229-233 -> MemoryFilter.matches :67-109) is one pass of
the scan kernels: every filter is a query of the predicate program, every condition an output bit of
its field's DFA.  Actions (renames) stay on the host and only touch matched records.
"""
from __future__ import annotations

import re
from typing import Any, Dict, List, Optional, Sequence, Union

import numpy as np

from .. import packer
from ..program import C_BODY, C_FLAGS, C_NAME, C_SLOT, NAME_DATE_STR, NAME_HOST, NAME_ID, NAME_TS_STR, Cond, ProgramBuilder, const
from ..regexc import Pattern
from . import utils as U
from .utils import move_memory, update_memory_flags, STANDARD_FOLDERS  # noqa: F401  (reference re-exports)


def _compile_filter_conditions(conditions: Sequence[Dict[str, Any]]) -> List[Cond]:
    out: List[Cond] = []
    for c in conditions:
        field, pattern, negate = c["field"], c["pattern"], bool(c["negate"])
        re.compile(pattern, re.IGNORECASE)                   # re.error propagates, as in the reference (filter.py:105)
        pat = Pattern("regex", pattern, re.IGNORECASE)
        if field == "content":
            out.append(Cond(C_BODY, pattern=pat, negate=negate))
        elif field == "flags":                               # a header literally called "flags" wins (filter.py:90-93)
            out.append(Cond(C_SLOT, pattern=pat, negate=negate, field=field, mode=1, if_missing=2))
            out.append(Cond(C_FLAGS, pattern=pat, negate=negate))
        elif field in ("unique_id", "hostname"):             # header first, else str(metadata[field]) (filter.py:94-95)
            out.append(Cond(C_SLOT, pattern=pat, negate=negate, field=field, mode=1, if_missing=2))
            out.append(Cond(C_NAME, pattern=pat, negate=negate, which=NAME_ID if field == "unique_id" else NAME_HOST))
        elif field in ("timestamp", "date"):                 # str(metadata[field]): decimal digits / "YYYY-MM-DD HH:MM:SS" (filter.py:94-95)
            out.append(Cond(C_SLOT, pattern=pat, negate=negate, field=field, mode=1, if_missing=2))
            out.append(Cond(C_NAME, pattern=pat, negate=negate, which=NAME_TS_STR if field == "timestamp" else NAME_DATE_STR))
        else:                                                # missing field: passes iff negated (filter.py:97-102)
            out.append(Cond(C_SLOT, pattern=pat, negate=negate, field=field, mode=1, if_missing=1 if negate else 0))
    return out


class MemoryFilter:
    """Filter for memories based on conditions (reference filter.py:20-173)."""

    def __init__(self, name: str):
        self.name = name
        self.conditions: List[Dict[str, Any]] = []
        self.actions: List[Dict[str, Any]] = []

    def add_condition(self, field: str, pattern: str, negate: bool = False) -> "MemoryFilter":
        self.conditions.append({"field": field, "pattern": pattern, "negate": negate})
        return self

    def add_action(self, action_type: str, **params) -> "MemoryFilter":
        self.actions.append({"type": action_type, **params})
        return self

    def matches(self, memory: Dict[str, Any]) -> bool:
        """One record through the same kernels (a 1-record packed corpus)."""
        if not self.conditions:
            return True
        from ..corpus import Corpus
        meta = memory["metadata"]
        flags = "".join(meta["flags"])
        rec = {"filename": f"{meta['timestamp']}.{meta['unique_id']}.{meta['hostname']}:2,{flags}", "folder": memory.get("folder", ""),
               "status": memory.get("status", "new") if memory.get("status") in U.STANDARD_FOLDERS else "new", "ts": int(meta["timestamp"]),
               "uid": meta["unique_id"], "host": meta["hostname"], "flags": flags, "date": meta["date"], "has_sep": True,
               "hdr_text": "".join(f"{k}: {v}\n" for k, v in memory["headers"].items()), "body_text": memory.get("content", "")}
        a = len(str(meta["timestamp"])) + 1
        rec["uid_span"] = (a, a + len(meta["unique_id"]))
        rec["host_span"] = (rec["uid_span"][1] + 1, rec["uid_span"][1] + 1 + len(meta["hostname"]))
        arrays = packer.arrays_from_segments([rec], {rec["folder"]: 0})
        pb = ProgramBuilder()
        pb.add_query(_compile_filter_conditions(self.conditions))
        c = Corpus().load(arrays)
        try:
            return bool(c.scan_masks(pb.build())[0] & 1)
        finally:
            c.close()

    def apply_actions(self, memory: Dict[str, Any]) -> List[str]:
        msgs: List[str] = []
        for action in self.actions:
            kind = action["type"]
            if kind == "move":
                tf, ts = action.get("target_folder", ""), action.get("target_status", "cur")
                if move_memory(memory["filename"], memory["folder"], tf, memory["status"], ts):
                    msgs.append(f"Moved to {tf or 'Inbox'}/{ts}")
            elif kind == "flag":
                flags, mode = action.get("flags", ""), action.get("mode", "add")
                cur = "".join(memory["metadata"]["flags"])
                if mode == "add":
                    new = "".join(sorted(set(cur + flags)))
                elif mode == "remove":
                    new = "".join(f for f in cur if f not in flags)
                else:
                    new = flags
                if update_memory_flags(memory["filename"], memory["folder"], memory["status"], new):
                    msgs.append(f"Flags updated from '{cur}' to '{new}'")
            elif kind == "copy":
                msgs.append(f"Would copy to {action.get('target_folder', '') or 'Inbox'}")
        return msgs


class FilterManager:
    """Manager for memory filters (reference filter.py:175-261)."""

    def __init__(self):
        self.filters: List[MemoryFilter] = []

    def add_filter(self, filter_obj: MemoryFilter) -> None:
        self.filters.append(filter_obj)

    def match_matrix(self, pm: "packer.PackedMemdir") -> np.ndarray:
        """mask[i] bit f = filter f accepts record i; 32 filters per pass."""
        n = pm.n
        masks = np.zeros((max(1, (len(self.filters) + 31) // 32), n), dtype=np.uint32)
        for g in range(0, len(self.filters), 32):
            pb = ProgramBuilder()
            for f in self.filters[g:g + 32]:
                conds = _compile_filter_conditions(f.conditions)
                pb.add_query(conds if conds else [const(True)])
            masks[g // 32] = pm.scan_masks(pb.build())
        return masks

    def process_memories(self, folders: List[str] = None, statuses: List[str] = None, dry_run: bool = False) -> Dict[str, Any]:
        if statuses is None:
            statuses = ["new"]
        pm = packer.packed()
        with pm.lock:                                        # the match matrix and the matched records come from one corpus state
            ranges = pm.ranges(folders, statuses)
            pm.report_skipped(folders, statuses)
            masks = self.match_matrix(pm) if self.filters and pm.n else np.zeros((1, pm.n), dtype=np.uint32)
            idx = np.concatenate([np.arange(a, b) for a, b in ranges]) if ranges else np.zeros(0, dtype=np.int64)
            any_hit = np.zeros(pm.n, dtype=bool)
            for row in masks:
                any_hit |= row != 0
            matched = idx[any_hit[idx]] if idx.size else idx
            memories = pm.materialize(matched, True)         # the reference gathers every memory before it acts (filter.py:213-216)
        stats = {"total_memories": int(idx.size), "filters_applied": 0, "actions_taken": 0, "memories_modified": 0, "details": []}
        modified = set()
        for i, memory in zip(matched.tolist(), memories):
            applied = []
            for fi, f in enumerate(self.filters):
                if not (int(masks[fi // 32, i]) >> (fi % 32)) & 1:
                    continue
                stats["filters_applied"] += 1
                actions = f.apply_actions(memory) if not dry_run else [f"Would {a['type']}" for a in f.actions]
                stats["actions_taken"] += len(actions)
                if actions:
                    modified.add(memory["metadata"]["unique_id"])
                    applied.append({"filter": f.name, "actions": actions})
            if applied:
                stats["details"].append({"memory_id": memory["metadata"]["unique_id"],
                                         "subject": memory["headers"].get("Subject", "No subject"), "filters_applied": applied})
        stats["memories_modified"] = len(modified)
        return stats


def create_default_filters() -> FilterManager:
    """The reference's six default filters (filter.py:263-309)."""
    m = FilterManager()
    m.add_filter(MemoryFilter("Python Content").add_condition("Tags", r"python").add_condition("content", r"python|django|flask", negate=True)
                 .add_action("move", target_folder=".Projects/Python", target_status="cur").add_action("flag", flags="P", mode="add"))
    m.add_filter(MemoryFilter("AI Content").add_condition("Tags", r"ai|machine[- ]learning|neural|llm")
                 .add_action("move", target_folder=".Projects/AI", target_status="cur"))
    m.add_filter(MemoryFilter("Learning Content").add_condition("Tags", r"books|reading|learning").add_condition("Subject", r"books|read|learning")
                 .add_action("move", target_folder=".ToDoLater/Learning", target_status="cur"))
    m.add_filter(MemoryFilter("High Priority").add_condition("Priority", r"high").add_action("flag", flags="FP", mode="add"))
    m.add_filter(MemoryFilter("Completed Items").add_condition("Status", r"completed|done|archived")
                 .add_action("move", target_folder=".Archive/2023", target_status="cur").add_action("flag", flags="S", mode="add"))
    m.add_filter(MemoryFilter("Trash Items").add_condition("Tags", r"trash|delete|remove")
                 .add_action("move", target_folder=".Trash", target_status="cur"))
    return m


def _print_stats(stats: Dict[str, Any]) -> None:
    print(f"Processed {stats['total_memories']} memories")
    print(f"Applied {stats['filters_applied']} filters")
    print(f"Took {stats['actions_taken']} actions")
    print(f"Modified {stats['memories_modified']} memories")
    if stats["details"]:
        print("\nDetails:")
        for d in stats["details"]:
            print(f"- {d['subject']} ({d['memory_id']})")
            for fa in d["filters_applied"]:
                print(f"  - {fa['filter']}: {', '.join(fa['actions'])}")


def run_filters(dry_run: bool = False) -> None:
    """Run the default filters on new memories; prints a summary and returns None (filter.py:311-328)."""
    _print_stats(create_default_filters().process_memories(statuses=["new"], dry_run=dry_run))


def apply_filters(filters: Union[FilterManager, Sequence[MemoryFilter], None] = None, folders: Optional[List[str]] = None,
                  statuses: Optional[List[str]] = None, dry_run: bool = False) -> Dict[str, Any]:
    """Name used by BASELINE.json's north_star; not present in the reference.  Returns process_memories' statistics."""
    if filters is None:
        mgr = create_default_filters()
    elif isinstance(filters, FilterManager):
        mgr = filters
    else:
        mgr = FilterManager()
        for f in filters:
            mgr.add_filter(f)
    return mgr.process_memories(folders=folders, statuses=statuses, dry_run=dry_run)
