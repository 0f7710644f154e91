"""Differential fuzzing of the pattern compiler against CPython's re / str semantics (CPU suite).
Random regexes from a small grammar (literals, classes, escapes, groups, alternation, quantifiers, anchors,
inline flags) x random strings over an adversarial alphabet."""
import random
import re

import pytest

from fei_b200.regexc import Pattern, PatternTooLarge, compile_patterns

ALPHABET = list("abkKsS \n.-_,/09") + ["K", "ſ", "é", "É", "ß", "日", "\U0001F409", "\t", " "]
ATOMS = ["a", "b", "k", "s", "K", r"\.", ".", r"\w", r"\W", r"\d", r"\s", r"\S", "[ab]", "[^a]", "[a-k]", "[k-s]", r"[\w-]", "é", "ß", "日", " ", "-", ",",
         r"\b", r"\B", "^", "$", r"\A", r"\Z", "[é日]", r"[^\W\d]", "/"]


def rand_regex(rng, depth=0):
    n = rng.randint(1, 4)
    parts = []
    for _ in range(n):
        r = rng.random()
        if depth < 2 and r < 0.2:
            inner = rand_regex(rng, depth + 1)
            parts.append(rng.choice(["(%s)", "(?:%s)", "(?i:%s)", "(?-i:%s)", "(?s:%s)", "(?m:%s)"]) % inner)
        elif depth < 2 and r < 0.3:
            parts.append("(?:%s|%s)" % (rand_regex(rng, depth + 1), rand_regex(rng, depth + 1)))
        else:
            parts.append(rng.choice(ATOMS))
        if parts[-1] not in ("^", "$", r"\b", r"\B", r"\A", r"\Z") and rng.random() < 0.3:
            parts[-1] += rng.choice(["*", "+", "?", "{2}", "{1,2}", "{0,1}", "*?", "+?"])
    return "".join(parts)


def rand_text(rng):
    return "".join(rng.choice(ALPHABET) for _ in range(rng.randint(0, 12)))


def test_random_regexes_agree_with_re():
    rng = random.Random(20240921)
    checked = 0
    for _ in range(220):
        p = rand_regex(rng)
        try:
            rx = re.compile(p, re.IGNORECASE)
        except re.error:
            continue
        try:
            d = compile_patterns([Pattern("regex", p, re.IGNORECASE)])
            ds = compile_patterns([Pattern("regex", p, re.IGNORECASE)], sticky=True)
        except PatternTooLarge:
            continue
        for _ in range(40):
            t = rand_text(rng)
            want = rx.search(t) is not None
            raw = t.encode("utf-8")
            assert bool(d.run(raw) & 1) == want, (p, t)
            assert bool(ds.run(raw) & 1) == want, ("sticky", p, t)
            checked += 1
    assert checked > 3000


def test_random_unions_agree_with_re():
    rng = random.Random(7)
    for _ in range(25):
        pats = []
        while len(pats) < 6:
            p = rand_regex(rng)
            try:
                re.compile(p, re.IGNORECASE); pats.append(p)
            except re.error:
                pass
        try:
            d = compile_patterns([Pattern("regex", p, re.IGNORECASE) for p in pats])
        except PatternTooLarge:
            continue
        for _ in range(60):
            t = rand_text(rng)
            m = d.run(t.encode("utf-8"))
            for k, p in enumerate(pats):
                assert bool(m >> k & 1) == (re.search(p, t, re.IGNORECASE) is not None), (p, t, pats)


def test_random_string_operators_agree_with_str():
    rng = random.Random(11)
    for _ in range(100):
        needle = "".join(rng.choice(ALPHABET[:14] + ["é", "ß", "日"]) for _ in range(rng.randint(0, 4)))
        nl = needle.lower()
        d = compile_patterns([Pattern("contains", nl), Pattern("startswith", nl), Pattern("endswith", nl), Pattern("equals", nl), Pattern("has_tag", nl),
                              Pattern("exact_contains", needle), Pattern("cmp_gt", needle), Pattern("cmp_le", needle), Pattern("exact_equals", needle)])
        for _ in range(40):
            v = rand_text(rng)
            m = d.run(v.encode("utf-8"))
            vl = v.lower()
            want = [nl in vl, vl.startswith(nl), vl.endswith(nl), vl == nl, nl in [x.strip() for x in vl.split(",")], needle in v, v > needle, v <= needle, v == needle]
            got = [bool(m >> k & 1) for k in range(9)]
            assert got == want, (needle, v, got, want)
