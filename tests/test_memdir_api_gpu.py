"""GPU parity of the reference-shaped API (search_memories / FilterManager.process_memories / apply_filters)
on an on-disk Memdir: vs the oracle on the same tree (exact order) and vs the reference-generated goldens."""
import contextlib
import io

import pytest

from oracle import memdir_oracle as mo
from tests.memdir_util import build_tree, key_of, same_modulo_ties

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api(gpu, tmp_path_factory):
    base = str(tmp_path_factory.mktemp("memdir_api") / "Memdir")
    g = build_tree(base)
    from fei_b200.memdir_tools import utils as U
    U.set_memdir_base(base)
    return base, g


def _query(conds, include_content, sort=None, rev=False, limit=None, offset=0):
    from fei_b200.memdir_tools.search import SearchQuery
    q = SearchQuery()
    for f, op, v in conds:
        q.add_condition(f, op, v)
    q.with_content(include_content)
    if sort:
        q.set_sort(sort, rev)
    if limit is not None or offset:
        q.set_pagination(limit, offset)
    return q


def test_search_memories_matches_oracle_and_reference(api):
    from fei_b200.memdir_tools.search import search_memories
    base, g = api
    for q in g["queries"]:
        with contextlib.redirect_stdout(io.StringIO()):
            res = search_memories(_query(q["conditions"], q["include_content"]), q["folders"], q["statuses"])
            mems = mo.listing(base, q["folders"], q["statuses"], q["include_content"])
        want = [mems[i] for i in mo.run_search(mems, [{"field": f, "operator": op, "value": v} for f, op, v in q["conditions"]])]
        assert [key_of(m) for m in res] == [key_of(m) for m in want], q["name"]
        assert same_modulo_ties([key_of(m) for m in res], q["result"]), q["name"]
        for a, b in zip(res[:5], want[:5]):                      # result dict shape (utils.py:234-243)
            assert a["headers"] == b["headers"] and a["metadata"] == b["metadata"] and a.get("content") == b.get("content")


def test_raising_queries_raise_typeerror(api):
    from fei_b200.memdir_tools.search import search_memories
    base, g = api
    for q in g["raising"]:
        with pytest.raises(TypeError):
            search_memories(_query(q["conditions"], False))


def test_raising_query_raises_the_first_reached_records_error(api):
    """Mixed raising kinds under one condition: the reference raises what the FIRST record reaching it raises."""
    from fei_b200.memdir_tools.search import search_memories
    base, g = api
    conds = [("Created", ">", "2020-01-01")]
    mems = mo.listing(base, None, None, False)
    with pytest.raises(TypeError) as want:
        mo.run_search(mems, [{"field": f, "operator": op, "value": v} for f, op, v in conds])
    with pytest.raises(TypeError) as got:
        search_memories(_query(conds, False))
    assert str(got.value) == str(want.value)


def test_capital_sigma_lowers_by_context(api):
    """str.lower() turns U+03A3 into a final or a medial sigma depending on its neighbours; the automata of sigma-bearing needles
    carry that rule (regexc._add_sigma_exact), so nothing is refused and nothing is guessed."""
    from fei_b200.memdir_tools.search import search_memories
    base, g = api
    mems = mo.listing(base, None, None, True)
    for field, needle in (("Subject", "σοφια"), ("Subject", "ςοφια"), ("content", "ευς"), ("content", "ευσ"), ("content", "σσ"), ("content", "σς")):
        for op in ("contains", "endswith", "startswith", "="):
            conds = [(field, op, needle)]
            want = [key_of(mems[i]) for i in mo.run_search(mems, [{"field": f, "operator": o, "value": v} for f, o, v in conds])]
            got = [key_of(m) for m in search_memories(_query(conds, True))]
            assert got == want, (field, op, needle)


def test_sort_and_pagination(api):
    from fei_b200.memdir_tools.search import search_memories
    base, g = api
    for q in g["sorted"]:
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            res = search_memories(_query(q["conditions"], False, q["sort"], q["reverse"], q["limit"], q["offset"]))
        assert len(res) == len(q["result"]), q["name"]
        assert ("Unable to sort" in buf.getvalue()) == ("Unable to sort" in q["printed"]), q["name"]
        if "Unable to sort" in q["printed"]:                 # fell back to a global newest-first sort (search.py:379-382)
            from tests.memdir_util import ts_of
            got = [key_of(m) for m in res]
            assert [ts_of(k) for k in got] == [ts_of(k) for k in q["result"]], q["name"]
            assert sorted(map(tuple, got)) == sorted(map(tuple, q["result"])) or q["limit"] is not None, q["name"]
        elif q["sort"] is None:
            assert same_modulo_ties([key_of(m) for m in res], q["result"]) or q["offset"] or q["limit"], q["name"]
        else:
            assert sorted(map(tuple, map(key_of, res))) == sorted(map(tuple, q["result"])) or q["limit"] is not None, q["name"]


def test_process_memories_and_apply_filters(api):
    from fei_b200.memdir_tools import filter as F
    from tests.golden.make_golden_memdir import FILTERS_EXTRA
    base, g = api
    for run in g["filters"]:
        mgr = F.create_default_filters()
        for fx in FILTERS_EXTRA:
            f = F.MemoryFilter(fx["name"])
            for fld, pat, neg in fx["conditions"]:
                f.add_condition(fld, pat, neg)
            for a in fx["actions"]:
                f.add_action(a["type"], **{k: v for k, v in a.items() if k != "type"})
            mgr.add_filter(f)
        with contextlib.redirect_stdout(io.StringIO()):
            stats = mgr.process_memories(statuses=run["statuses"], dry_run=True)
            again = F.apply_filters(mgr, statuses=run["statuses"], dry_run=True)
        want = run["stats"]
        assert stats == again
        for k in ("total_memories", "filters_applied", "actions_taken", "memories_modified"):
            assert stats[k] == want[k], (run["statuses"], k)
        by_id = {d["memory_id"]: d for d in want["details"]}
        assert sorted(d["memory_id"] for d in stats["details"]) == sorted(by_id)
        for d in stats["details"]:
            assert d["filters_applied"] == by_id[d["memory_id"]]["filters_applied"] and d["subject"] == by_id[d["memory_id"]]["subject"]


def test_single_record_matches_and_bad_regex(api):
    import re
    from fei_b200.memdir_tools import filter as F
    base, g = api
    with contextlib.redirect_stdout(io.StringIO()):
        mems = mo.listing(base, [""], ["cur"], True)[:40]
    flt = F.MemoryFilter("t").add_condition("Tags", "python").add_condition("content", "django", negate=True)
    conds = [{"field": "Tags", "pattern": "python", "negate": False}, {"field": "content", "pattern": "django", "negate": True}]
    for m in mems:
        assert flt.matches(m) == mo.filter_accepts(m, conds)
    with pytest.raises(re.error):
        F.MemoryFilter("bad").add_condition("Tags", "(unclosed").matches(mems[0])


def test_real_filter_actions_move_and_flag(gpu, tmp_path):
    """Not a dry run: matched records are renamed exactly like the reference's apply_actions (filter.py:111-173)."""
    import os
    from fei_b200 import synth
    from fei_b200.memdir_tools import filter as F, utils as U
    base = str(tmp_path / "Memdir")
    recs = [synth.record(5, i) for i in range(60)]
    for r in recs:
        r["status"], r["status_id"] = "new", 1
    synth.write_memdir(base, recs)
    old = U.MEMDIR_BASE
    U.set_memdir_base(base)
    try:
        mgr = F.FilterManager()
        mgr.add_filter(F.MemoryFilter("hp").add_condition("Priority", "high").add_action("flag", flags="FP", mode="add"))
        with contextlib.redirect_stdout(io.StringIO()):
            mems = mo.listing(base, None, ["new"], True)
        want = [m for m in mems if mo.filter_accepts(m, [{"field": "Priority", "pattern": "high", "negate": False}])]
        stats = mgr.process_memories(dry_run=False)
        assert stats["filters_applied"] == len(want)
        for m in want:
            new_flags = "".join(sorted(set("".join(m["metadata"]["flags"]) + "FP")))
            name = m["filename"].split(":2,")[0] + ":2," + new_flags
            d = os.path.join(base, m["folder"], "new") if m["folder"] else os.path.join(base, "new")
            assert os.path.exists(os.path.join(d, name)), name
    finally:
        U.set_memdir_base(old)


def test_incremental_sync_after_flag_and_move(gpu, tmp_path, capsys):
    """Renames (flag updates, moves) and new files invalidate only the touched directories; file contents already
    seen are not read again; results stay identical to the oracle's fresh listing; skipped files are reported on
    every listing like the reference does."""
    import os
    from fei_b200 import packer, synth
    from fei_b200.memdir_tools import utils as U
    from fei_b200.memdir_tools.search import search_memories
    base = str(tmp_path / "Memdir")
    recs = [synth.record(21, i) for i in range(200)]
    synth.write_memdir(base, recs)
    with open(os.path.join(base, "cur", "1700009999.badbad01.host:2,"), "wb") as f:
        f.write(b"Subject: x\n---\n\xff\xfe")
    old = U.MEMDIR_BASE
    U.set_memdir_base(base)
    try:
        q = _query([("Tags", "has_tag", "python")], False)
        r1 = search_memories(q)
        pm = packer.packed()
        assert pm.files_read == 201
        assert "Error processing 1700009999.badbad01.host:2,:" in capsys.readouterr().out
        # flag update (rename inside one directory) + a move across folders + one brand-new memory
        m = r1[0]
        assert U.update_memory_flags(m["filename"], m["folder"], m["status"], "FRS")
        m2 = r1[1]
        assert U.move_memory(m2["filename"], m2["folder"], ".Archive", m2["status"], "cur")
        U.save_memory(".Projects/AI", "fresh body about python", {"Tags": "python,new", "Subject": "fresh"}, "P")
        r2 = search_memories(q)
        pm2 = packer.packed()
        assert pm2 is pm and pm.files_read == 1                      # only the new file's content was read
        assert "Error processing 1700009999.badbad01.host:2,:" in capsys.readouterr().out
        mems = mo.listing(base, None, None, False)
        capsys.readouterr()
        want = [mems[i] for i in mo.run_search(mems, [{"field": "Tags", "operator": "has_tag", "value": "python"}])]
        assert [key_of(x) for x in r2] == [key_of(x) for x in want]
        assert len(r2) == len(r1) + 1
        got_flags = [x for x in r2 if x["metadata"]["unique_id"] == m["metadata"]["unique_id"]][0]["metadata"]["flags"]
        assert got_flags == list("FRS")
    finally:
        U.set_memdir_base(old)


def _oracle_keys(base, conds, include_content=False):
    mems = mo.listing(base, None, None, include_content)
    return [key_of(mems[i]) for i in mo.run_search(mems, [{"field": f, "operator": op, "value": v} for f, op, v in conds])]


def test_in_place_rewrite_delete_and_bad_file_reporting(gpu, tmp_path, capsys):
    """The reference rewrites memory files in place (folders.py:575, archiver.py:591): the directory's mtime does not move, the
    packed corpus must still notice.  Deleted files disappear; an undecodable file in an untouched directory keeps being
    reported on every listing (utils.py:247-248) without being read again."""
    import os, time
    from fei_b200 import packer, synth
    from fei_b200.memdir_tools import utils as U
    from fei_b200.memdir_tools.search import search_memories
    base = str(tmp_path / "Memdir")
    synth.write_memdir(base, [synth.record(33, i) for i in range(300)])
    with open(os.path.join(base, ".Projects/AI", "cur", "1700009999.badbad02.host:2,"), "wb") as f:
        f.write(b"Subject: x\n---\n\xff\xfe")
    old = U.MEMDIR_BASE
    U.set_memdir_base(base)
    try:
        conds = [("Tags", "has_tag", "zznewtag")]
        assert search_memories(_query(conds, False)) == []
        pm = packer.packed()
        assert "badbad02" in capsys.readouterr().out
        victim = search_memories(_query([("flags", "has_flag", "F")], False))[0]
        path = os.path.join(U.get_memory_path(victim["folder"], victim["status"]), victim["filename"])
        text = open(path).read()
        with open(path, "w") as f:                                        # same name, same directory entry: an in-place rewrite
            f.write(text.replace("Tags: ", "Tags: zznewtag,", 1))
        got = search_memories(_query(conds, False))
        assert [key_of(m) for m in got] == [key_of(victim)] == _oracle_keys(base, conds)
        assert pm.files_read == 1 and pm.windows_packed == 1 and packer.packed() is pm
        assert "badbad02" in capsys.readouterr().out                    # reported again, though its directory was not touched
        os.remove(path)
        assert search_memories(_query(conds, False)) == []
        assert pm.files_read == 0
        every = search_memories(_query([], False))
        capsys.readouterr()
        assert [key_of(m) for m in every] == _oracle_keys(base, [])
        for env in ("0", "1"):                                           # the same through the no-inotify path (periodic native re-listing)
            os.environ["FEI_INOTIFY"] = env; os.environ["FEI_REVALIDATE_S"] = "0"
            packer.drop()
            search_memories(_query(conds, False))
            p2 = os.path.join(base, "cur", [n for n in os.listdir(os.path.join(base, "cur"))][0])
            t2 = open(p2).read()
            with open(p2, "w") as f:
                f.write(t2.replace("---", "Tags: zznewtag\n---", 1) if "Tags: " not in t2 else t2.replace("Tags: ", "Tags: zznewtag,", 1))
            assert [key_of(m) for m in search_memories(_query(conds, False))] == _oracle_keys(base, conds) != []
            with open(p2, "w") as f:
                f.write(t2)
    finally:
        os.environ.pop("FEI_INOTIFY", None); os.environ.pop("FEI_REVALIDATE_S", None)
        packer.drop()
        U.set_memdir_base(old)


def test_cold_pack_without_stat_pass_equals_default(gpu, tmp_path, capsys, monkeypatch):
    """FEI_COLD_ARENA=1: names-only listing + open/fstat/read/close into an arena + fei_corpus_load_raw_spans.  Same listing, same
    hits, same report of an undecodable file, and the incremental sync that follows works on the keys taken from the open files."""
    import os
    from fei_b200 import packer, synth
    from fei_b200.memdir_tools import utils as U
    from fei_b200.memdir_tools.search import search_memories
    base = str(tmp_path / "Memdir")
    synth.write_memdir(base, [synth.record(35, i) for i in range(5000)])
    with open(os.path.join(base, ".Projects/AI", "cur", "1700009999.badbad03.host:2,"), "wb") as f:
        f.write(b"Subject: x\n---\n\xff\xfe")
    old = U.MEMDIR_BASE
    U.set_memdir_base(base)
    try:
        conds = [("content", "matches", "python|rust"), ("flags", "has_flag", "S")]
        want = _oracle_keys(base, conds, True)
        results = {}
        for arena in ("0", "tiny-chunks", "one-buffer", "1"):
            monkeypatch.setenv("FEI_COLD_ARENA", "1" if arena == "1" else "0")
            monkeypatch.setenv("FEI_COLD_CHUNKS", "0" if arena == "one-buffer" else "1")
            monkeypatch.setattr(packer, "COLD_CHUNK_BYTES", 50_000 if arena == "tiny-chunks" else 256 << 20)   # many chunks per directory, slots reused
            packer.drop()
            got = search_memories(_query(conds, True))
            pm = packer.packed()
            assert "badbad03" in capsys.readouterr().out
            path_name = pm.timing["cold_path"]
            assert path_name.startswith("names-only" if arena == "1" else "listing with stat")
            assert ("reused host buffers" in path_name) == (arena in ("0", "tiny-chunks")) and ("one exact buffer" in path_name) == (arena == "one-buffer")
            assert [key_of(m) for m in got] == want != []
            results[arena] = [(m["filename"], m["headers"], m["content"]) for m in got]
        assert results["0"] == results["1"] == results["tiny-chunks"] == results["one-buffer"]
        monkeypatch.setenv("FEI_COLD_ARENA", "1"); monkeypatch.setenv("FEI_COLD_CHUNKS", "1")
        victim = search_memories(_query([("flags", "has_flag", "F")], False))[0]          # still on the arena-packed corpus
        path = os.path.join(U.get_memory_path(victim["folder"], victim["status"]), victim["filename"])
        text = open(path).read()
        with open(path, "w") as f:
            f.write(text.replace("---", "Tags: zzarena\n---", 1) if "Tags: " not in text else text.replace("Tags: ", "Tags: zzarena,", 1))
        got = search_memories(_query([("Tags", "has_tag", "zzarena")], False))
        assert [key_of(m) for m in got] == [key_of(victim)]
        assert pm.files_read == 1 and packer.packed() is pm
    finally:
        packer.drop()
        U.set_memdir_base(old)


def test_one_changed_file_in_100k_repacks_one_window(gpu, tmp_path, capsys):
    """Incremental sync at scale: 100 k files packed once; one new file, one rewritten file and one flag change later at most one
    4096-record window is (re)packed and exactly the files whose content is new are read.  Results equal a fresh listing by the oracle."""
    import os
    from fei_b200 import packer, synth
    from fei_b200.memdir_tools import utils as U
    from fei_b200.memdir_tools.search import search_memories
    base = str(tmp_path / "Memdir")
    synth.write_memdir_native(base, 0xFE1, 0, 100_000)
    old = U.MEMDIR_BASE
    U.set_memdir_base(base)
    try:
        conds = [("Tags", "has_tag", "kubernetes"), ("flags", "has_flag", "P"), ("content", "matches", "terraform")]
        r1 = search_memories(_query(conds, True))
        pm = packer.packed()
        assert pm.files_read == 100_000 and pm.full_packs == 1 and pm.n == 100_000
        U.save_memory(".Projects/AI", "terraform notes", {"Tags": "kubernetes,new", "Subject": "fresh"}, "P")
        r2 = search_memories(_query(conds, True))
        assert pm.files_read == 1 and pm.windows_packed == 1 and pm.full_packs == 1 and len(r2) == len(r1) + 1
        m = r1[len(r1) // 2]
        assert U.update_memory_flags(m["filename"], m["folder"], m["status"], "FS")          # drops P: leaves the result
        r3 = search_memories(_query(conds, True))
        assert pm.files_read == 0 and pm.windows_packed == 1 and pm.full_packs == 1 and len(r3) == len(r2) - 1
        capsys.readouterr()
        assert [key_of(x) for x in r3] == _oracle_keys(base, conds, True)
        want = {tuple(key_of(m)): m for m in mo.listing(base, None, None, True)}
        capsys.readouterr()
        assert all(x["content"] == want[tuple(key_of(x))]["content"] and x["headers"] == want[tuple(key_of(x))]["headers"] for x in r3[:200])
    finally:
        packer.drop()
        U.set_memdir_base(old)


def test_snapshot_restore_and_resync(gpu, tmp_path, capsys):
    """save_snapshot / from_snapshot: the restored corpus answers like the packed one, and the first sync after the restore reads
    only what changed on disk since the snapshot."""
    import os
    from fei_b200 import packer, synth
    from fei_b200.memdir_tools import utils as U
    from fei_b200.memdir_tools.search import search_memories
    base = str(tmp_path / "Memdir")
    synth.write_memdir_native(base, 5, 0, 6000)
    old = U.MEMDIR_BASE
    U.set_memdir_base(base)
    try:
        conds = [("content", "matches", "docker|rust"), ("Tags", "has_tag", "python")]
        r1 = search_memories(_query(conds, True))
        snap = str(tmp_path / "snap")
        packer.packed().save_snapshot(snap)
        packer.drop()
        U.save_memory("", "rust and python after the snapshot", {"Tags": "python", "Subject": "late"}, "")
        pm = packer.PackedMemdir.from_snapshot(base, snap)
        assert pm.snapshot_gbs is not None and pm.n == 6000
        packer._cache[base] = pm
        r2 = search_memories(_query(conds, True))
        assert pm.files_read == 1 and pm.full_packs == 0
        capsys.readouterr()
        assert [key_of(x) for x in r2] == _oracle_keys(base, conds, True) and len(r2) == len(r1) + 1
        assert [(x["headers"], x["content"]) for x in r2 if x["headers"].get("Subject") != "late"] == [(x["headers"], x["content"]) for x in r1]
    finally:
        packer.drop()
        U.set_memdir_base(old)


def test_legacy_substring_search(api):
    """memdir_tools.utils.search_memories (utils.py:299-352): any header value / content substring, previews."""
    from fei_b200.memdir_tools import utils as U
    base, g = api
    assert g["legacy"]
    for case in g["legacy"]:
        with contextlib.redirect_stdout(io.StringIO()):
            res = U.search_memories(case["query"], case["folders"], case["statuses"], case["headers_only"])
            mems = mo.listing(base, case["folders"], case["statuses"], True)
        want = mo.legacy_search(mems, case["query"], case["headers_only"])
        assert [key_of(m) for m in res] == [key_of(m) for m in want], case
        assert same_modulo_ties([key_of(m) for m in res], case["result"]), case
        for a, b in zip(res, want):
            assert a.get("content_preview") == b.get("content_preview") and ("content" in a) == ("content" in b) and a["headers"] == b["headers"]
        assert [("content" in m) for m in res][:4] == case["has_content_key"], case


def test_folder_stats(api):
    """MemdirFolderManager.get_folder_stats (folders.py:216-318): counts from the packed segments, flag counts and the tag
    statistics (fei_corpus_token_histogram) from the GPU, newest / oldest from the packed wall-clock column."""
    from fei_b200.memdir_tools.folders import MemdirFolderManager
    base, g = api
    assert g["folder_stats"]
    for case in g["folder_stats"]:
        with contextlib.redirect_stdout(io.StringIO()):
            got = MemdirFolderManager().get_folder_stats(case["folder_path"], case["include_subfolders"])
            want = mo.folder_stats(base, case["folder_path"], case["include_subfolders"])
        assert got == want, case["folder_path"]                                       # same tree: exact, including dict order
        assert list(got["tags"].keys()) == list(want["tags"].keys()), case["folder_path"]
        ref = case["stats"]                                                           # the reference on its own copy of the tree
        for k in ("folder", "total_memories", "memory_counts", "flag_counts"):
            assert got[k] == ref[k], (case["folder_path"], k)
        by_name = lambda subs: sorted(subs, key=lambda x: x["folder"])                # folder order = os.walk order: differs between boxes
        assert by_name(got["subfolders"]) == by_name(ref["subfolders"]), case["folder_path"]
        assert got["tags"] == ref["tags"]                                             # order may differ on timestamp ties (listdir order)
        for k in ("newest_memory", "oldest_memory"):
            assert (got[k] is None) == (ref[k] is None) and (got[k] is None or str(got[k]["date"]) == ref[k]["date"])


def test_archiver_criteria_and_cleanup(api):
    """MemoryArchiver._memory_matches_criteria / cleanup_memories (archiver.py:128-181, :306-381) through match_criteria on the GPU."""
    import numpy as np
    from fei_b200 import packer
    from fei_b200.memdir_tools.archiver import MemoryArchiver
    base, g = api
    arch = MemoryArchiver()
    crits = [c["criteria"] for c in g["criteria"]]
    with contextlib.redirect_stdout(io.StringIO()):
        rows = arch.match_criteria(crits)
        pm = packer.packed()
        mems = mo.listing(base, None, None, True)
    order = pm.ranges(None, None)
    idx = np.concatenate([np.arange(a, b) for a, b in order])
    assert idx.size == len(mems)
    for c, row in zip(g["criteria"], rows):
        got = [key_of(mems[k]) for k in np.nonzero(row[idx])[0].tolist()]
        want = [key_of(m) for m in mems if mo.matches_criteria(m, c["criteria"])]
        assert got == want, c["criteria"]
        assert same_modulo_ties(got, c["result"]), c["criteria"]
        for m in mems[:40]:                                   # the single-record host form agrees too
            assert arch._memory_matches_criteria(m, c["criteria"]) == mo.matches_criteria(m, c["criteria"]), c["criteria"]
    a2 = MemoryArchiver()
    a2.add_cleanup_rule({"tags": "python", "Priority": "high"}, "trash")
    a2.add_cleanup_rule({"flags": "P"}, "delete")
    with contextlib.redirect_stdout(io.StringIO()):
        st = a2.cleanup_memories(dry_run=True)
    ref = g["cleanup_dry_run"]
    assert (st["trashed"], st["deleted"]) == (ref["trashed"], ref["deleted"])
    key = lambda d: (d["action"], d["memory_id"], d["subject"])
    assert sorted(map(key, st["details"])) == sorted(map(key, ref["details"]))
    with contextlib.redirect_stdout(io.StringIO()):
        st = MemoryArchiver().cleanup_memories(dry_run=True)
    assert (st["trashed"], st["deleted"]) == (g["cleanup_default_dry_run"]["trashed"], g["cleanup_default_dry_run"]["deleted"])
