"""CPU suite for the pattern -> byte-DFA compiler: the compiled automaton (simulated on the
host, tests only) must agree with CPython's re / str semantics the reference relies on
(search.py:141-242, filter.py:105) on adversarial values."""
import random
import re

import pytest

from fei_b200.regexc import Pattern, compile_patterns

VALUES = [
    "", "a", "K", "k", "\u212a", "s", "S", "\u017f", "ſtraße", "STRASSE", "python", "Python3", "xpythonx", "py\nthon",
    "docker and kubernetes", "kubernetes\ndocker", "Docker ... Kubernetes", "line1\nline2", "line2", "a.b", "vue.js", "vueXjs",
    "foo bar", "foobar", "foo_bar", "foo-bar", " foo ", "bar foo", "FOO", "ﬁ", "é", "É", "éa", "naïve café", "日本語テキスト",
    "Ünïcödé wörd", "x\u00a0y", "tab\there", "a,b,c", "a, b ,c", "python,learning", " python ,x", "PYTHON", "py", "pythons",
    "1234", "v1.2.3", "2024-01-01", "aaa", "aaaa", "ab" * 10, "\U0001F409 dragon", "end\n", "\nstart", "mid\n\nmid",
    "react", "Angular", "reactangular", "re act", "ci/cd", "CI/CD pipeline", "ui/ux", "node.js", "nodeXjs", "big data", "bigdata",
    "ǅ", "ǆ", "Σ", "σ", "ς", "ΑΣ", "İ", "i̇", "ß", "ẞ", "\u0345", "\x85", "\x1c", "a\x0bb", "word1 word2", "_under_", "٣",
]

REGEXES = [
    r"python", r"docker|kubernetes", r"kubernetes.*docker|docker.*kubernetes", r"react|angular", r"vue\.js", r"node\.js",
    r"ci/cd", r"big data", r"\bk", r"k\b", r"\Bk", r"^line2", r"(?m)^line2", r"line1$", r"(?m)line1$", r"\Aa", r"a\Z", r"^$", r"",
    r"\w+", r"\d+", r"\s", r"[a-c]+x", r"[^a-z]", r"fo+\s?bar", r"(foo|bar){2}", r"a{3}", r"a{2,3}$", r"^a{0,2}$", r"(ab)+$",
    r"s", r"ſ", r"ss", r"ß", r"é", r"[é]", r"(?a)\w+é", r"(?a:\bfoo\b)", r"\bfoo\b", r"foo\b.", r"x.y", r"(?s)x.y", r"py.thon",
    r"(?s)py.thon", r"\.", r"[.]js", r"日本", r"\bword\d\b", r"^\s*foo\s*$", r"σ", r"Σ", r"ς", r"i", r"İ", r"ǆ", r"(?i:K)", r"(?-i:K)",
    r"[\W_]+", r"^[^,]+,[^,]+$", r"(?x) p y # comment", r"a|", r"(|a)b", r"\d{4}-\d{2}-\d{2}", r"v\d+(\.\d+)*$", r"[\d\.]+$",
    r"\x85", r"[\x1c-\x1f]", r"\bab", r"(a|ab)(c|bcd)?$", r"\b", r"\B", r"$", r"^", r"\Z", r"dragon$", r"\U0001F409",
]


def _expect_regex(p, v):
    return bool(re.search(p, v, re.IGNORECASE))


def test_regex_search_semantics_individually():
    for p in REGEXES:
        d = compile_patterns([Pattern("regex", p, re.IGNORECASE)])
        for v in VALUES:
            got = bool(d.run(v.encode("utf-8")) & 1)
            assert got == _expect_regex(p, v), (p, v)


def test_regex_union_is_bitwise_equal_to_individual_runs():
    rng = random.Random(3)
    for _ in range(6):
        pats = rng.sample(REGEXES, 12)
        d = compile_patterns([Pattern("regex", p, re.IGNORECASE) for p in pats])
        for v in VALUES:
            m = d.run(v.encode("utf-8"))
            for k, p in enumerate(pats):
                assert bool(m >> k & 1) == _expect_regex(p, v), (p, v)


def test_unsupported_constructs_raise():
    for p in [r"(\w)\1", r"(?<=has) x", r"(?=x)", r"(?!x)y", r"(?>a*)a", r"a*+a", r"(a)?(?(1)b|c)"]:
        with pytest.raises(NotImplementedError):
            compile_patterns([Pattern("regex", p, re.IGNORECASE)])
    with pytest.raises(re.error):
        compile_patterns([Pattern("regex", r"(unclosed", re.IGNORECASE)])


LOWER_OK = lambda s: True               # capital sigma and dotted capital I included: both are modelled exactly


def test_lowercase_string_operators():
    needles = ["", "a", "k", "K", "python", "PYTHON", "s", "ſ", "é", "É", "b ,c", "foo", " foo", "bar", "c", "js", "σ", "ß", "1.2", "\u212a"]
    for nd in needles:
        nl = nd.lower()
        d = compile_patterns([Pattern("contains", nl), Pattern("startswith", nl), Pattern("endswith", nl), Pattern("equals", nl), Pattern("has_tag", nl)])
        for v in VALUES:
            if not LOWER_OK(v):
                continue
            m = d.run(v.encode("utf-8"))
            vl = v.lower()
            assert bool(m & 1) == (nl in vl), ("contains", nd, v)
            assert bool(m & 2) == vl.startswith(nl), ("startswith", nd, v)
            assert bool(m & 4) == vl.endswith(nl), ("endswith", nd, v)
            assert bool(m & 8) == (vl == nl), ("equals", nd, v)
            assert bool(m & 16) == (nl in [t.strip() for t in vl.split(",")]), ("has_tag", nd, v)


def test_dotted_capital_i_lowers_to_two_characters():
    """'İ'.lower() == 'i' + U+0307: one input character advances the needle by two, may end a match inside the expansion
    (needle ends in 'i') or start one at the combining dot."""
    import random
    rnd = random.Random(7)
    alpha = ["i", "\u0307", "\u0130", "I", "a", "n", ",", " ", "x", "K"]
    for _ in range(250):
        nl = "".join(rnd.choice(alpha) for _ in range(rnd.randint(0, 3))).lower()
        d = compile_patterns([Pattern("contains", nl), Pattern("startswith", nl), Pattern("endswith", nl), Pattern("equals", nl), Pattern("has_tag", nl)])
        for _ in range(30):
            v = "".join(rnd.choice(alpha) for _ in range(rnd.randint(0, 6)))
            m = d.run(v.encode("utf-8"))
            vl = v.lower()
            assert bool(m & 1) == (nl in vl), ("contains", nl, v)
            assert bool(m & 2) == vl.startswith(nl), ("startswith", nl, v)
            assert bool(m & 4) == vl.endswith(nl), ("endswith", nl, v)
            assert bool(m & 8) == (vl == nl), ("equals", nl, v)
            assert bool(m & 16) == (nl in [t.strip() for t in vl.split(",")]), ("has_tag", nl, v)


def test_capital_sigma_final_form_rule():
    """'Σ'.lower() is 'ς' after a cased letter and before a non-cased one (case-ignorable characters skipped), else 'σ':
    needles that hold a sigma get the product automaton of regexc._add_sigma_exact."""
    import random
    rnd = random.Random(11)
    alpha = ["σ", "ς", "Σ", "a", "A", "'", "\u0307", ".", " ", ",", "\u0130", "ο", "Ο", "1", "i"]
    for kind in ("contains", "startswith", "endswith", "equals", "has_tag"):
        for _ in range(40):
            nl = "".join(rnd.choice(alpha) for _ in range(rnd.randint(1, 4))).lower()
            if "σ" not in nl and "ς" not in nl:
                nl += rnd.choice("σς")
            d = compile_patterns([Pattern(kind, nl)])
            for _ in range(40):
                v = "".join(rnd.choice(alpha) for _ in range(rnd.randint(0, 8)))
                vl = v.lower()
                want = {"contains": nl in vl, "startswith": vl.startswith(nl), "endswith": vl.endswith(nl), "equals": vl == nl,
                        "has_tag": nl in [t.strip() for t in vl.split(",")]}[kind]
                assert bool(d.run(v.encode("utf-8")) & 1) == want, (kind, nl, v)


def test_exact_and_ordering_operators():
    needles = ["", "F", "FS", "SF", "a", "high", "=high", "b", "aa", "é", "日本", "python", "Z", "~"]
    for nd in needles:
        d = compile_patterns([Pattern("exact_contains", nd), Pattern("cmp_gt", nd), Pattern("cmp_ge", nd), Pattern("cmp_lt", nd), Pattern("cmp_le", nd)])
        for v in VALUES + ["FS", "SF", "FRS", "high", "medium", "low"]:
            m = d.run(v.encode("utf-8"))
            assert bool(m & 1) == (nd in v), ("in", nd, v)
            assert bool(m & 2) == (v > nd), (">", nd, v)
            assert bool(m & 4) == (v >= nd), (">=", nd, v)
            assert bool(m & 8) == (v < nd), ("<", nd, v)
            assert bool(m & 16) == (v <= nd), ("<=", nd, v)


def test_batch_of_32_baseline_patterns_is_compact():
    pats = ["python", "docker|kubernetes", "neural networks", "react", "angular", "rust", "django", "flask", "terraform", "ansible",
            "microservices", "big data", "ci/cd", "git", "aws|azure|gcp", "spring boot", r"vue\.js", r"node\.js", "devops", "security",
            "blockchain", "testing", "databases", "algorithms", "cloud computing", "mobile development", "computer vision",
            "reinforcement learning", "ui/ux", "web development", "data structures", "machine learning"]
    d = compile_patterns([Pattern("regex", p, re.IGNORECASE) for p in pats])
    assert d.n_states < 1000
    text = "Notes on Machine Learning, big DATA and ci/cd; also Vue.js / node.js with KUBERNETES + gcp. digital"
    m = d.run(text.encode())
    for k, p in enumerate(pats):
        assert bool(m >> k & 1) == bool(re.search(p, text, re.IGNORECASE)), p
