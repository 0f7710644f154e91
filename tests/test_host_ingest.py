"""Host side of the cold pack (no GPU): native directory listing and the two read paths of fei_b200.packer give the same files.

Reference behaviour being mirrored: utils.list_memories (memdir_tools/utils.py:202-253) lists a cur/new/tmp directory, keeps the
names matching the Maildir grammar, reads every file and sorts newest first."""
import os

import numpy as np
import pytest

from fei_b200 import packer, synth


@pytest.fixture(scope="module")
def tree(tmp_path_factory):
    base = str(tmp_path_factory.mktemp("memdir"))
    synth.write_memdir_native(base, 0xFE1, 0, 3000, 4)
    d = os.path.join(base, "cur")
    with open(os.path.join(d, "1700000001.abc.host:2,S"), "wb") as f:          # CRLF + non-ASCII content, read verbatim
        f.write("Subject: x\r\n---\r\nbödy\r\n".encode())
    with open(os.path.join(d, "not-a-memory.txt"), "wb") as f:
        f.write(b"ignored")
    open(os.path.join(d, "1700000002.empty.host:2,"), "wb").close()
    return base


def test_names_only_listing_matches_stat_listing(tree):
    d = os.path.join(tree, "cur")
    a, b = packer.list_dir(d, True), packer.list_dir(d, False)
    assert a.n == b.n > 100 and a.names == b.names and np.array_equal(a.name_off, b.name_off)
    for k in ("ts", "wall", "flags8", "ino"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert np.array_equal(a.spans, b.spans)
    assert (b.size == 0).all() and (b.mtime_ns == -1).all()
    assert (np.diff(a.ts) <= 0).all()                                             # newest first (utils.py:251)


def test_packed_read_equals_listed_read(tree):
    d = os.path.join(tree, "cur")
    a = packer.list_dir(d, True)
    raw, off, problems = packer.read_files(d, a)
    assert not problems
    b = packer.list_dir(d, False)
    arena = packer._Arena(64 << 20)
    try:
        begin, ln, err = packer.read_dir_packed(d, b, arena)
        assert not err.any()
        assert int(arena.cursor.value) == int(off[-1]) == int(ln.sum())
        for i in range(a.n):
            want = raw[int(off[i]):int(off[i + 1])].tobytes()
            assert arena.buf[int(begin[i]):int(begin[i] + ln[i])].tobytes() == want, a.name(i)
            with open(os.path.join(d, a.name(i)), "rb") as f:
                assert f.read() == want
        # size / inode / mtime now come from the open files and equal what stat reports
        assert np.array_equal(b.size, a.size) and np.array_equal(b.ino, a.ino) and np.array_equal(b.mtime_ns, a.mtime_ns)
        spans = sorted(zip(begin.tolist(), ln.tolist()))
        assert all(s0 + l0 <= s1 for (s0, l0), (s1, _l1) in zip(spans, spans[1:]))      # no two files overlap in the arena
    finally:
        arena.close()


def test_packed_read_reports_vanished_oversized_and_full_arena(tree, monkeypatch):
    import errno
    d = os.path.join(tree, "new")
    b = packer.list_dir(d, False)
    assert b.n > 10
    victim = os.path.join(d, b.name(3))
    keep = open(victim, "rb").read()
    os.unlink(victim)
    monkeypatch.setattr(packer, "MAX_BODY", 100)                                  # almost every synthetic file is larger than that
    arena = packer._Arena(1 << 20)
    try:
        begin, ln, err = packer.read_dir_packed(d, b, arena)
        assert err[3] == errno.ENOENT
        sizes = np.array([os.stat(os.path.join(d, b.name(i))).st_size if i != 3 else 0 for i in range(b.n)])
        big = sizes > 100
        assert big.any() and (err[big] == errno.EFBIG).all() and np.array_equal(ln[big], sizes[big].astype(np.uint64))
        ok = (~big) & (np.arange(b.n) != 3)
        assert not err[ok].any()
    finally:
        arena.close()
        with open(victim, "wb") as f:
            f.write(keep)
    monkeypatch.setattr(packer, "MAX_BODY", 32 << 20)
    b = packer.list_dir(d, False)
    tiny = packer._Arena(4096)
    try:
        begin, ln, err = packer.read_dir_packed(d, b, tiny)
        assert (err == errno.ENOMEM).any() and not ((err != 0) & (err != errno.ENOMEM)).any()
        for i in np.nonzero(err == 0)[0].tolist():
            with open(os.path.join(d, b.name(i)), "rb") as f:
                assert tiny.buf[int(begin[i]):int(begin[i] + ln[i])].tobytes() == f.read()
    finally:
        tiny.close()


class _RecordingCorpus:
    """Stands in for fei_b200.corpus.Corpus on a box without a GPU: keeps what fei_corpus_stage_text would have been given."""

    def __init__(self):
        self.text = None

    def stage_text(self, total, piece, offset):
        if self.text is None or len(self.text) != total:
            self.text = np.full(total, 0xEE, dtype=np.uint8)
        self.text[offset:offset + len(piece)] = piece


def _cold_read(tree, monkeypatch, chunk_bytes):
    from fei_b200.memdir_tools import utils as U
    monkeypatch.setattr(packer, "COLD_CHUNK_BYTES", chunk_bytes)
    pm = packer.PackedMemdir(tree)
    pm._set_folders(pm._walk())
    order = pm._order()
    segs = {}
    corpus = _RecordingCorpus()
    raw, begin, ln, err, staged = pm._cold_read_listed(order, segs, corpus)
    return pm, order, segs, corpus, raw, begin, ln, err, staged


@pytest.mark.parametrize("chunk_bytes", [10_000, 1 << 20, 256 << 20])
def test_chunked_cold_read_stages_every_file_at_its_offset(tree, monkeypatch, chunk_bytes):
    """The default cold read (listing with stat, chunks into reused buffers, each chunk staged at its offset in the device text):
    whatever the chunk size, the staged text is the files back to back in listing order and (begin, len) point at them."""
    pm, order, segs, corpus, raw, begin, ln, err, staged = _cold_read(tree, monkeypatch, chunk_bytes)
    assert staged and raw is None and not err.any()
    assert "reused host buffers" in pm.timing["cold_path"]
    j = 0
    for key in order:
        L = segs[key].listing
        for i in range(L.n):
            with open(os.path.join(pm._dir(*key), L.name(i)), "rb") as f:
                want = f.read()
            assert int(ln[j]) == len(want) and corpus.text[int(begin[j]):int(begin[j] + ln[j])].tobytes() == want, (key, L.name(i))
            j += 1
    assert j == len(begin) > 3000 and int(begin[-1] + ln[-1]) == len(corpus.text)
    assert not (corpus.text == 0xEE).all()


def test_chunked_cold_read_falls_back_when_a_file_changes_under_the_listing(tree, monkeypatch):
    """A file that shrank between the listing and the read: the chunked path gives up and the one-buffer path re-reads it."""
    real_list = packer.list_dir
    victim = {}

    def list_then_truncate(path, want_stat=True):
        L = real_list(path, want_stat)
        if path.endswith(os.path.join("new")) and L.n and not victim:
            victim["path"] = os.path.join(path, L.name(0)); victim["keep"] = open(victim["path"], "rb").read()
            with open(victim["path"], "wb") as f:
                f.write(victim["keep"][:10])
        return L
    monkeypatch.setattr(packer, "list_dir", list_then_truncate)
    try:
        pm, order, segs, corpus, raw, begin, ln, err, staged = _cold_read(tree, monkeypatch, 1 << 20)
        assert victim and "one exact buffer" in pm.timing["cold_path"]
        assert raw is not None
        j = 0
        for key in order:
            L = segs[key].listing
            for i in range(L.n):
                with open(os.path.join(pm._dir(*key), L.name(i)), "rb") as f:
                    assert raw[int(begin[j]):int(begin[j] + ln[j])].tobytes() == f.read()
                j += 1
    finally:
        if victim:
            with open(victim["path"], "wb") as f:
                f.write(victim["keep"])


def test_native_listing_agrees_with_the_oracle_on_adversarial_file_names(tmp_path, capsys):
    """fei_dir_list's file-name grammar (memdir_host.cpp parse_name; names it cannot judge go to Python) against the oracle's
    read_folder (utils.py:216-251 restated): which names are listed, in which order, and timestamp / unique_id / hostname / flags."""
    import random
    from oracle import memdir_oracle as mo
    rng = random.Random(77)
    d = tmp_path / "Memdir" / "cur"
    d.mkdir(parents=True)
    fixed = ["1700000000.abc.host:2,", "1700000000.abc.host:2,FRS", "1700000001.a1b2.h.o.s.t:2,S", "1700000002.abc.host:2,Fjunk", "1700000003.abc.host:2,FSPRXYZQ",
             "1700000004.ABC.host:2,S", "1700000005.abc.:2,S", "1700000006.abc.host:3,S", "1700000007..host:2,S", ".1700000008.abc.host:2,S", "x1700000009.abc.host:2,S",
             "1700000010.abc.host:2", "1700000011.abc.ho:st:2,S", "1700000012.abc.host:2,S:2,T", "0.a.b:2,", "00000000000000000012.abc.host:2,S", "17000000130000000000000.abc.host:2,S",
             "1700000014.abc.hôst:2,S", "١٧٠٠٠٠٠٠١٥.abc.host:2,S", "1700000016.abc.host :2,S", "1700000017.abc.host:2,s", "1700000018.ab_c.host:2,S",
             "1700000019.abc.host:2,SSSSSSS", "1700000019.abd.host:2,", "1700000019.abb.host:2,", "999999999999.abc.host:2,S", "not-a-memory.txt", "1700000020.abc.host:2,F\nx"]
    names = set(fixed)
    while len(names) < 400:
        ts = rng.choice(["17%08d" % rng.randrange(10 ** 8), str(rng.randrange(10 ** rng.randrange(1, 12))), "1700000100"])
        uid = "".join(rng.choice("abc019xyzAB_.") for _ in range(rng.randrange(0, 6)))
        host = "".join(rng.choice("hostä:.-2,") for _ in range(rng.randrange(0, 6)))
        tail = rng.choice([":2,", ":2,S", ":2,FRS", ":2,Fx", ":2", ":1,S", "", ":2,SS:2,F"])
        n = f"{ts}.{uid}.{host}{tail}"
        if "/" not in n and "\0" not in n and 0 < len(n.encode()) < 200:
            names.add(n)
    for k, n in enumerate(sorted(names)):
        (d / n).write_bytes(b"Subject: %d\n---\nbody %d\n" % (k, k))
    want = mo.read_folder(str(tmp_path / "Memdir"), "", "cur", False)
    printed = capsys.readouterr().out
    L = packer.list_dir(str(d), True)
    got = [L.name(i) for i in range(L.n)]
    # year > 9999 etc.: the oracle prints "Error processing" and skips; the native listing reports the same names as bad
    def reported(messages):
        return sorted(n for n in names if any(m.startswith(f"Error processing {n}: ") for m in messages))
    bad_native = reported(L.bad)
    bad_oracle = reported(printed.splitlines())
    unpackable = [n for n in bad_native if n not in bad_oracle]                  # > 7 flag letters: refused by the packed layout, loudly
    assert all(len(mo.split_filename(n)["flags"]) > 7 for n in unpackable), unpackable
    assert [n for n in bad_native if n not in unpackable] == bad_oracle
    want_names = [m["filename"] for m in want if m["filename"] not in unpackable]
    # ties in timestamp keep os.listdir order in both (stable sorts over the same readdir order)
    assert got == want_names
    for i, m in enumerate(x for x in want if x["filename"] not in unpackable):
        nb = L.name_bytes(i); sp = L.spans[i]; f8 = int(L.flags8[i])
        assert int(L.ts[i]) == m["metadata"]["timestamp"], m["filename"]
        assert os.fsdecode(nb[int(sp[0]):int(sp[0] + sp[1])]) == m["metadata"]["unique_id"], m["filename"]
        assert os.fsdecode(nb[int(sp[2]):int(sp[2] + sp[3])]) == m["metadata"]["hostname"], m["filename"]
        assert [chr((f8 >> (8 * k)) & 0xFF) for k in range(f8 >> 56)] == m["metadata"]["flags"], m["filename"]
    assert len(got) > 50 and len(names) - len(got) > 100          # both sides of the grammar are exercised


_WALL_CHECK = r"""
import calendar, os, sys
from datetime import datetime
sys.path.insert(0, sys.argv[1])
from fei_b200 import packer
d = sys.argv[2]
L = packer.list_dir(d, True)
assert L.n > 500, L.n
bad = 0
for i in range(L.n):
    ts = int(L.ts[i])
    want = calendar.timegm(datetime.fromtimestamp(ts).timetuple())           # naive local wall clock, as utils.py:94 builds it
    if int(L.wall[i]) != want:
        bad += 1
        print("MISMATCH", ts, int(L.wall[i]), want)
print("checked", L.n, "bad", bad)
sys.exit(1 if bad else 0)
"""


@pytest.mark.parametrize("tz", ["Europe/Prague", "America/New_York", "Australia/Lord_Howe", "UTC", "Asia/Kolkata"])
def test_listing_wall_clock_equals_datetime_fromtimestamp_across_dst(tmp_path, tz):
    """The listing's `wall` column (per-day UTC-offset cache in memdir_host.cpp) against datetime.fromtimestamp under zones with DST
    (including Lord Howe's half-hour shift): timestamps packed around every transition of 2023-2025, on both sides, to the second."""
    import subprocess, sys
    if not os.path.exists(os.path.join("/usr/share/zoneinfo", tz)):
        pytest.skip("no tzdata for " + tz)
    d = tmp_path / "cur"
    d.mkdir()
    stamps = set()
    for year in (2023, 2024, 2025):
        for month, day in ((3, 10), (3, 12), (3, 26), (3, 31), (4, 2), (4, 7), (10, 1), (10, 6), (10, 27), (10, 29), (11, 3), (11, 5), (1, 1), (7, 1), (12, 31)):
            base = int(__import__("calendar").timegm((year, month, day, 0, 0, 0)))
            for h in range(-14, 40):
                for delta in (-1, 0, 1, 1799, 1800, 1801, 3599):
                    stamps.add(base + 3600 * h + delta)
    stamps.update((0, 1, 59, 86399, 86400, 2 ** 31 - 1, 2 ** 31, 4102444800, 253402300799 - 86400 * 30))
    for k, ts in enumerate(sorted(stamps)):
        (d / f"{ts}.u{k}.host:2,S").write_bytes(b"x")
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", _WALL_CHECK, repo, str(d)], env=dict(os.environ, TZ=tz), capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
