"""Memdir part of make_golden.py: run the unmodified reference's search / filter code over a
seeded on-disk Memdir (synthetic records + adversarial files) and record what it returns."""
from __future__ import annotations

import base64
import contextlib
import io
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
SEED, N_SYNTH = 0xFE1, 400

# (folder, status, filename, raw bytes) — parser and decode quirks
ADVERSARIAL = [
    ("", "cur", "1700000500.adv00001.hostA:2,FS", b"Subject: Beta --- split here\nTags: x\n---\nbody after second sep"),
    ("", "cur", "1700000500.adv00002.hostA:2,", b"Tags: a,b\ntags: lower,c\nTags: final,python\n---\ndup keys"),
    ("", "new", "1700000501.adv00003.hostB:2,S", "  Subject  :  spaced out  \n\tTags\t:\tpython , rust\t\n---\n\u00dcn\u00efc\u00f6d\u00e9 K \u017f body\nwith KELVIN \u212a and long s \u017f\n".encode()),
    (".Projects/AI", "cur", "1700000502.adv00004.hostB:2,RSP", b"NoColonLine\nTags: python\n: emptykey\nKey:\n---\n"),
    (".Projects/AI", "new", "1700000503.adv00005.host-c:2,P", "Tags:\u00a0python\u2003\n---\nnbsp and em-space around the tag".encode()),
    ("", "cur", "1700000504.adv00006.hostA:2,F", b"Subject: x\r\nStatus: done\r\nStatus: ACTIVE\r\n---\r\ncrlf file\r\nsecond line react\r\n"),
    ("", "cur", "1700000505.adv00007.hostA:2,", b"no separator at all: just text mentioning python and docker then kubernetes"),
    (".ToDoLater/Learning", "cur", "1700000506.adv00008.hostA:2,SF", "Tags: Python,\u212aelvin\nPriority: HIGH\n---\nkelvin sign tag".encode()),
    ("", "tmp", "1700000507.adv00009.hostA:2,", "Subject: caf\u00e9 R\u00c9SUM\u00c9\n---\ncaf\u00e9 r\u00e9sum\u00e9 \U0001F409 dragon\nline2 react".encode()),
    ("", "cur", "1700000508.adv00010.hostA:2,F", b"Tags: a\x0bb,\x1cpython\x1f\n---\nodd whitespace controls"),
    ("", "cur", "1700000509.adv00011.hostA:2,", b"Subject: broken utf8\n---\nbad byte \xff\xfe here python"),          # skipped by the reference
    ("", "cur", "1700000510.adv00012.hostA:2,FRSP", b"Subject: all flags\nTags: python,learning\nDue: 2030-05-05\n---\n\n\n  padded body  \n\n"),
    ("", "cur", "not-a-memory.txt", b"ignored: name does not match"),
    ("", "cur", "1700000511.adv00013.hostA:2,Fjunk", b"Subject: trailing junk after flags\n---\nflags F then lowercase junk"),
    (".Projects/Python", "new", "1700000512.adv00014.hostA:2,", b"---\n---\nbody starts with a separator"),
    ("", "new", "1700000513.adv00015.hostA:2,", b"Subject: empty body\nTags: python\n---"),
    # round 2: values only dateutil can judge, a timestamp header, and the two characters whose str.lower() is not one-to-one
    ("", "cur", "1700000514.adv00016.hostA:2,S", b"Subject: someday due\nDue: someday\nModified: 2024-02-30\nTags: python\n---\nunparseable dates stay strings"),
    ("", "cur", "1700000515.adv00017.hostA:2,", b"Subject: aware created\nCreated: 2031-01-01T10:00:00+02:00\nDue: 2030-05-05 00:00:00\n---\ntz-aware Created header"),
    (".Projects/AI", "cur", "1700000516.adv00018.hostA:2,F", b"Subject: partial due\nDue: May 5\nModified: 2024-02-10\nCreated: 10 Jan 2031 08:00\n---\nyear comes from today"),
    ("", "new", "1700000517.adv00019.hostA:2,", b"Subject: header named timestamp\nTimestamp: 12345\nDUE: 2030-05-05\ndue: 1999-01-01\n---\nthe header wins over the metadata"),
    ("", "cur", "1700000518.adv00020.hostA:2,P", "Subject: \u0130stanbul \u03a3\u039f\u03a6\u0399\u0391\nTags: \u0130zmir, x\n---\n\u0130 in the body, and \u039f\u0394\u03a5\u03a3\u03a3\u0395\u03a5\u03a3 too\nreact".encode()),
]

QUERIES = [
    # (name, [(field, op, value)...], include_content, folders, statuses)
    ("cfg1_single_regex", [("content", "matches", r"kubernetes.*docker|docker.*kubernetes")], True, None, None),
    ("cfg2_multi_field", [("Tags", "has_tag", "python"), ("flags", "has_flag", "F"), ("date", ">", "2023-11-14 22:14:00"), ("content", "matches", r"react|angular")], True, None, None),
    ("tag_python", [("Tags", "has_tag", "python")], False, None, None),
    ("tag_PYTHON_upper", [("tags", "has_tag", "PYTHON")], False, None, None),
    ("tag_final_dupkeys", [("Tags", "has_tag", "final")], False, None, None),
    ("tag_lower_dupkeys", [("tags", "has_tag", "lower")], False, None, None),
    ("flags_FS", [("flags", "has_flag", "FS")], False, None, None),
    ("flags_SF", [("flags", "has_flag", "SF")], False, None, None),
    ("flags_eq_fs", [("flags", "=", "fs")], False, None, None),
    ("id_eq_upper", [("id", "=", "ADV00001")], False, None, None),
    ("filename_contains", [("filename", "contains", "hostb")], False, None, None),
    ("content_without_flag", [("content", "contains", "python")], False, None, None),
    ("content_with_flag", [("content", "contains", "python")], True, None, None),
    ("content_matches_empty_nocontent", [("content", "matches", "")], False, None, None),
    ("keyword_both", [("Subject", "contains", "review"), ("content", "contains", "review")], True, None, None),
    ("status_hdr_active", [("Status", "=", "active")], False, None, None),
    ("state_done", [("state", "=", "done")], False, None, None),
    ("status_maildir_new", [("status", "=", "new")], False, None, None),
    ("folder_contains_ai", [("folder", "contains", "ai")], False, None, None),
    ("priority_gt_str", [("priority", ">", "=high")], False, None, None),
    ("subject_startswith", [("Subject", "startswith", "research")], False, None, None),
    ("author_endswith", [("author", "endswith", "ng")], False, None, None),
    ("tags_contains_empty", [("Tags", "contains", "")], False, None, None),
    ("nope_field", [("nope", "contains", "")], False, None, None),
    ("empty_conditions", [], False, None, None),
    ("date_lt", [("date", "<", "2023-11-14 22:14:30")], False, None, None),
    ("date_eq", [("date", "=", "2023-11-14 22:21:40")], False, None, None),
    ("date_gt_now", [("date", ">", "now-7d")], False, None, None),
    ("date_garbage_ne", [("date", "!=", "garbage")], False, None, None),
    ("kelvin_b", [("content", "matches", r"\bk")], True, None, None),
    ("long_s", [("content", "matches", "s")], True, None, None),
    ("multiline", [("content", "matches", r"(?m)^line2")], True, None, None),
    ("bad_regex", [("content", "matches", r"(unclosed")], True, None, None),
    ("only_new_root", [("Tags", "has_tag", "python")], False, [""], ["new"]),
    ("folders_reordered", [("flags", "has_flag", "F")], False, [".Projects/AI", ""], ["new", "cur"]),
    ("unique_id_meta", [("unique_id", "=", "adv00004")], False, None, None),
    ("hostname_meta", [("hostname", "contains", "host-c")], False, None, None),
    ("unknown_op", [("Tags", "fuzzy", "x")], False, None, None),
    ("subject_ne", [("Subject", "!=", "x")], False, None, None),
    # round 2 -- date-like headers through dateutil (search.py:126-130, :166-234)
    ("due_eq", [("Due", "=", "2030-05-05")], False, None, None),
    ("due_ne_garbage", [("Due", "!=", "garbage")], False, None, None),
    ("due_gt_now", [("Due", ">", "now")], False, [".Projects/AI", ".Projects/Python"], None),       # every Due there parses: parse("now") fails -> False
    ("due_gt_date", [("due", ">", "2023-12-01")], False, None, None),
    ("due_le_date", [("DUE", "<=", "2023-12-10 00:00")], False, None, None),
    ("due_contains_year", [("Due", "contains", "2030")], False, None, None),
    ("due_eq_someday", [("Due", "=", "SOMEDAY")], False, None, None),
    ("due_matches", [("Due", "matches", "^2030-05-05 00")], False, None, None),
    ("due_ne_date", [("Due", "!=", "2030-05-05")], False, None, None),
    ("created_eq_aware", [("Created", "=", "2031-01-01T08:00:00+00:00")], False, None, None),
    ("created_ne_naive", [("Created", "!=", "2031-01-01 08:00")], False, None, None),
    ("modified_startswith", [("Modified", "startswith", "2024-02")], False, None, None),
    ("modified_lt", [("modified", "<", "2024-02-20")], False, None, None),
    ("deleteddate_absent", [("DeletedDate", "!=", "x")], False, None, None),
    # metadata timestamp (search.py:134-137) and text operators on the date field (search.py:148-163)
    ("ts_contains", [("timestamp", "contains", "17000005")], False, None, None),
    ("ts_eq_int", [("timestamp", "=", 1700000504)], False, None, None),
    ("ts_eq_str", [("timestamp", "=", "1700000504")], False, None, None),
    ("ts_eq_hdr", [("timestamp", "=", "12345")], False, None, None),
    ("ts_ne_str", [("timestamp", "!=", "x")], False, None, None),
    ("ts_matches", [("timestamp", "matches", "^1700000[01]")], False, None, None),
    ("date_contains", [("date", "contains", "2023-11-14 22:2")], False, None, None),
    ("date_matches", [("date", "matches", r"^2023-11-1[45] 2")], False, None, None),
    ("date_has_flag", [("date", "has_flag", "22:21")], False, None, None),
    ("date_endswith", [("date", "endswith", ":40")], False, None, None),
    # U+0130 lowers to two characters; a needle without sigma cannot see which sigma U+03A3 lowers to
    ("idot_contains_dotted", [("Subject", "contains", "i\u0307stanbul")], False, None, None),
    ("idot_contains_plain", [("Subject", "contains", "istanbul")], False, None, None),
    ("idot_startswith_i", [("Subject", "startswith", "i")], False, None, None),
    ("idot_tag", [("Tags", "has_tag", "\u0130zmir")], False, None, None),
    ("idot_content", [("content", "contains", "i\u0307 in")], True, None, None),
    ("sigma_free_needle", [("Subject", "contains", "\u03bf\u03c6\u03b9\u03b1")], False, None, None),
    # capital sigma lowers to the final form only after a cased letter and before a non-cased one (case-ignorables skipped)
    ("sigma_medial_at_word_start", [("Subject", "contains", "\u03c3\u03bf\u03c6\u03b9\u03b1")], False, None, None),
    ("sigma_final_needle_start", [("Subject", "contains", "\u03c2\u03bf\u03c6")], False, None, None),
    ("sigma_final_in_body", [("content", "contains", "\u03b5\u03c5\u03c2")], True, None, None),
    ("sigma_medial_not_final", [("content", "contains", "\u03b5\u03c5\u03c3")], True, None, None),
    ("sigma_double", [("content", "endswith", "\u03c3\u03c3\u03b5\u03c5\u03c2 too\nreact")], True, None, None),
]

RAISING = [
    ("priority_gt_now", [("Priority", ">", "now-1d")], "TypeError"),
    ("date_gt_aware", [("date", ">", "2023-11-14T22:14:00+02:00")], "TypeError"),
    ("created_gt_naive", [("Created", ">", "2020-01-01")], "TypeError"),
    ("due_gt_now_unparseable", [("Due", ">", "now")], "TypeError"),
    ("ts_gt_str", [("timestamp", ">", "5")], "TypeError"),
    ("due_gt_now_minus", [("Subject", "contains", "someday"), ("Due", ">", "now-1d")], "TypeError"),
]

SORTED = [
    ("sort_subject", [("Tags", "has_tag", "python")], "Subject", False, None, 0),
    ("sort_subject_rev_page", [("Tags", "has_tag", "python")], "subject", True, 5, 2),
    ("sort_date_meta", [("flags", "has_flag", "F")], "date", False, 10, 0),
    ("sort_mixed_due", [("Tags", "has_tag", "python")], "due", False, None, 0),
    ("offset_only", [("flags", "has_flag", "S")], None, False, None, 3),
]

PARSE = ["subject:python tags:learning", "content:/regex pattern/ date>2023-01-01", "priority:high status!=completed", "#python +FS review notes",
         "date>=2020-01-01", "sort:date limit:5 python", "tags:a,b, c with_content", 'Subject:"quoted phrase" state:active status:new', "+X flags:F",
         "status_value=done id=abc", "Priority<=low", 'plain words "and a phrase"', "x:/a/ y:// z:/", "status:pending status:cur"]

FILTERS_EXTRA = [
    {"name": "neg-missing", "conditions": [("Nope", "x", True)], "actions": [{"type": "copy", "target_folder": ".Archive"}]},
    {"name": "flags-f", "conditions": [("flags", "f", False)], "actions": [{"type": "flag", "flags": "P", "mode": "add"}]},
    {"name": "lowercase-key", "conditions": [("tags", "python", False)], "actions": [{"type": "move", "target_folder": ".Trash"}]},
    {"name": "uid", "conditions": [("unique_id", "^adv", False), ("content", "python", True)], "actions": []},
    {"name": "empty", "conditions": [], "actions": [{"type": "move", "target_folder": ".Trash"}]},
    {"name": "meta-ts", "conditions": [("timestamp", "^17000005", False)], "actions": [{"type": "flag", "flags": "S", "mode": "add"}]},
    {"name": "meta-date", "conditions": [("date", "2023-11-14 22:2", False), ("content", "python", True)], "actions": []},
]


# memdir_tools.utils.search_memories (the legacy substring search, utils.py:299-352): (query, folders, statuses, headers_only)
LEGACY = [
    ("python", None, None, False), ("python", None, None, True), ("", None, None, True), ("KELVIN", None, None, False),
    ("learning", [""], ["cur", "new"], False), ("beta", None, ["new", "cur"], True), ("no such words anywhere", None, None, False),
    ("high", [".Projects", ""], None, True), ("caf\u00e9", None, None, False), (":", None, None, True),
]


# MemoryArchiver._memory_matches_criteria (archiver.py:128-181) over the whole listing
CRITERIA = [
    {"tags": "python"}, {"tag": ["python", "rust"]}, {"tags": ["Python"]}, {"tags": "PYTHON"}, {"tags": [" python"]}, {"tags": []},
    {"Status": "active|done"}, {"status": "completed|done"}, {"Priority": "^high$", "tags": "python"}, {"Subject": 7},
    {"flags": "f"}, {"flags": "^FS?$"}, {"flags": ["F"]}, {"age": 100}, {"min_age": 100000}, {"max_age": 100000}, {"max_age": 10},
    {"age": 100, "tags": ["learning", "ai"], "Status": "."}, {"Nope": "x"}, {},
]


# MemdirFolderManager.get_folder_stats (folders.py:216-318): (folder_path, include_subfolders)
FOLDER_STATS = [("", False), ("", True), (".Projects", False), (".Projects", True), (".Archive", True), ("no-such-folder", False), (".Proj", True),
                ("/.Projects/", False)]


def build_corpus(base: str):
    sys.path.insert(0, REPO)
    from fei_b200 import synth
    recs = [synth.record(SEED, i) for i in range(N_SYNTH)]
    synth.write_memdir(base, recs)
    for folder, status, name, raw in ADVERSARIAL:
        d = os.path.join(base, folder, status) if folder else os.path.join(base, status)
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, name), "wb") as f:
            f.write(raw)


def key_of(m):
    return [m["folder"], m["status"], m["filename"]]


def make_memdir(scratch: str, import_reference):
    ru, rs, rf, rm = import_reference(scratch)
    base = os.path.join(scratch, "Memdir")
    build_corpus(base)
    assert ru.MEMDIR_BASE == base, (ru.MEMDIR_BASE, base)
    out = {"generator": "tests/golden/make_golden.py memdir", "seed": SEED, "n_synth": N_SYNTH,
           "adversarial": [[f, s, n, base64.b64encode(raw).decode()] for f, s, n, raw in ADVERSARIAL],
           "folders": ru.get_memdir_folders(), "queries": [], "raising": [], "sorted": [], "parse": [], "filters": []}

    def run(conds, include_content, folders, statuses, sort=None, rev=False, limit=None, offset=0):
        q = rs.SearchQuery()
        for f, op, v in conds:
            q.add_condition(f, op, v)
        q.with_content(include_content)
        if sort:
            q.set_sort(sort, rev)
        if limit is not None or offset:
            q.set_pagination(limit, offset)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            res = rs.search_memories(q, folders, statuses)
        return res, buf.getvalue()

    listing, printed = run([], True, None, None)
    out["listing"] = [key_of(m) for m in listing]
    out["listing_printed"] = printed
    for name, conds, inc, folders, statuses in QUERIES:
        res, printed = run(conds, inc, folders, statuses)
        out["queries"].append({"name": name, "conditions": conds, "include_content": inc, "folders": folders, "statuses": statuses,
                               "result": [key_of(m) for m in res], "has_content_key": ["content" in m for m in res][:3]})
    for name, conds, exc in RAISING:
        try:
            run(conds, False, None, None)
            got = None
        except Exception as e:
            got = type(e).__name__
        assert got == exc, (name, got)
        out["raising"].append({"name": name, "conditions": conds, "exception": got})
    for name, conds, sort, rev, limit, offset in SORTED:
        res, printed = run(conds, False, None, None, sort, rev, limit, offset)
        out["sorted"].append({"name": name, "conditions": conds, "sort": sort, "reverse": rev, "limit": limit, "offset": offset,
                              "result": [key_of(m) for m in res], "printed": printed})
    for s in PARSE:
        q = rs.parse_search_args(s)
        out["parse"].append({"input": s, "conditions": q.conditions, "sort_by": q.sort_by, "sort_reverse": q.sort_reverse,
                             "limit": q.limit, "offset": q.offset, "include_content": q.include_content})
    out["legacy"] = []
    for query, folders, statuses, headers_only in LEGACY:
        res = ru.search_memories(query, folders, statuses, headers_only)
        out["legacy"].append({"query": query, "folders": folders, "statuses": statuses, "headers_only": headers_only,
                              "result": [key_of(m) for m in res], "previews": [m.get("content_preview") for m in res][:4],
                              "has_content_key": ["content" in m for m in res][:4]})
    # filters: default set + extras, dry run, several status selections
    for statuses in (None, ["cur", "new", "tmp"], ["cur"]):
        mgr = rf.create_default_filters()
        for fx in FILTERS_EXTRA:
            f = rf.MemoryFilter(fx["name"])
            for fld, pat, neg in fx["conditions"]:
                f.add_condition(fld, pat, neg)
            for a in fx["actions"]:
                f.add_action(a["type"], **{k: v for k, v in a.items() if k != "type"})
            mgr.add_filter(f)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            stats = mgr.process_memories(statuses=statuses, dry_run=True)
        out["filters"].append({"statuses": statuses, "stats": stats})
    import memdir_tools.archiver as ra
    arch = ra.MemoryArchiver()
    out["criteria"] = [{"criteria": c, "result": [key_of(m) for m in listing if arch._memory_matches_criteria(m, c)]} for c in CRITERIA]
    arch2 = ra.MemoryArchiver()
    arch2.add_cleanup_rule({"tags": "python", "Priority": "high"}, "trash")
    arch2.add_cleanup_rule({"flags": "P"}, "delete")
    st = arch2.cleanup_memories(dry_run=True)
    out["cleanup_dry_run"] = {"trashed": st["trashed"], "deleted": st["deleted"], "details": st["details"]}
    st = ra.MemoryArchiver().cleanup_memories(dry_run=True)                  # the two default rules
    out["cleanup_default_dry_run"] = {"trashed": st["trashed"], "deleted": st["deleted"]}
    # folder statistics last: the manager's constructor creates the special folders, which changes get_memdir_folders()
    import memdir_tools.folders as rfo
    out["folder_stats"] = []
    for folder_path, include_sub in FOLDER_STATS:
        st = rfo.MemdirFolderManager().get_folder_stats(folder_path, include_sub)
        out["folder_stats"].append({"folder_path": folder_path, "include_subfolders": include_sub, "stats": st,
                                    "tag_order": list(st["tags"].keys())})
    out["folders_after_manager"] = ru.get_memdir_folders()
    # public signatures of the entry points the drop-in keeps (checked against fei_b200.memdir_tools on every CPU run)
    import inspect
    sig = lambda f: [[p.name, int(p.kind), None if p.default is inspect.Parameter.empty else repr(p.default)] for p in inspect.signature(f).parameters.values()]
    out["signatures"] = {
        "search.search_memories": sig(rs.search_memories), "search.parse_search_args": sig(rs.parse_search_args),
        "search.SearchQuery.add_condition": sig(rs.SearchQuery.add_condition), "search.SearchQuery.set_sort": sig(rs.SearchQuery.set_sort),
        "search.SearchQuery.set_pagination": sig(rs.SearchQuery.set_pagination), "search.SearchQuery.with_content": sig(rs.SearchQuery.with_content),
        "filter.MemoryFilter.__init__": sig(rf.MemoryFilter.__init__), "filter.MemoryFilter.add_condition": sig(rf.MemoryFilter.add_condition),
        "filter.MemoryFilter.add_action": sig(rf.MemoryFilter.add_action), "filter.MemoryFilter.matches": sig(rf.MemoryFilter.matches),
        "filter.MemoryFilter.apply_actions": sig(rf.MemoryFilter.apply_actions), "filter.FilterManager.add_filter": sig(rf.FilterManager.add_filter),
        "filter.FilterManager.process_memories": sig(rf.FilterManager.process_memories), "filter.create_default_filters": sig(rf.create_default_filters),
        "filter.run_filters": sig(rf.run_filters),
        "memorychain.MemoryBlock.__init__": sig(rm.MemoryBlock.__init__), "memorychain.MemoryBlock.calculate_hash": sig(rm.MemoryBlock.calculate_hash),
        "memorychain.MemoryBlock.mine_block": sig(rm.MemoryBlock.mine_block), "memorychain.MemoryBlock.to_dict": sig(rm.MemoryBlock.to_dict),
        "memorychain.MemoryBlock.from_dict": sig(rm.MemoryBlock.from_dict), "memorychain.MemoryChain.validate_chain": sig(rm.MemoryChain.validate_chain),
        "memorychain.MemoryChain.receive_chain_update": sig(rm.MemoryChain.receive_chain_update),
        "utils.parse_memory_filename": sig(ru.parse_memory_filename), "utils.parse_memory_content": sig(ru.parse_memory_content),
        "utils.list_memories": sig(ru.list_memories), "utils.move_memory": sig(ru.move_memory), "utils.update_memory_flags": sig(ru.update_memory_flags),
        "utils.save_memory": sig(ru.save_memory), "utils.get_memdir_folders": sig(ru.get_memdir_folders),
    }
    with open(os.path.join(HERE, "memdir_golden.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True, default=str)
    print("wrote memdir_golden.json:", len(out["listing"]), "memories listed,", len(out["queries"]), "queries,", len(out["filters"]), "filter runs")
