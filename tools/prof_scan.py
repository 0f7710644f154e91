#!/usr/bin/env python3
"""Tiny driver for ncu captures of one scan kernel: tools/prof_scan.py [batch32|single|nomatch|cfg2|hdr|hdr3] [entries]"""
import os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("TZ", "UTC")
from fei_b200.corpus import Corpus
from fei_b200.program import C_BODY, C_DATE_CMP, C_FLAGS, C_SLOT, CMP, Cond, ProgramBuilder, content_batch_program
from fei_b200.regexc import Pattern
import bench

what = sys.argv[1] if len(sys.argv) > 1 else "batch32"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
c = Corpus().synth(bench.SEED, 0, n)
if what == "batch32":
    prog, nq = content_batch_program([Pattern("regex", p, re.IGNORECASE) for p in bench.BATCH32]), 32
elif what in ("single", "nomatch"):
    pat = r"kubernetes.*docker|docker.*kubernetes" if what == "single" else r"quagga.*zebra|zebra.*quagga"
    pb = ProgramBuilder(); pb.add_query([Cond(C_BODY, pattern=Pattern("regex", pat, re.IGNORECASE))]); prog, nq = pb.build(), 1
elif what == "half":
    pb = ProgramBuilder()
    pb.add_query([Cond(C_FLAGS, pattern=Pattern("exact_contains", "F")), Cond(C_BODY, pattern=Pattern("regex", r"quagga.*zebra|zebra.*quagga", re.IGNORECASE))])
    prog, nq = pb.build(), 1
elif what == "hdr":
    pb = ProgramBuilder()
    pb.add_query([Cond(C_SLOT, pattern=Pattern("has_tag", "python"), field="Tags", mode=0)])
    prog, nq = pb.build(), 1
elif what == "hdr3":
    pb = ProgramBuilder()
    pb.add_query([Cond(C_SLOT, pattern=Pattern("has_tag", "python"), field="Tags", mode=0)])
    pb.add_query([Cond(C_SLOT, pattern=Pattern("contains", "learning"), field="Subject", mode=0)])
    pb.add_query([Cond(C_SLOT, pattern=Pattern("equals", "high"), field="Priority", mode=0), Cond(C_FLAGS, pattern=Pattern("exact_contains", "F"))])
    prog, nq = pb.build(), 3
else:
    pb = ProgramBuilder()
    pb.add_query([Cond(C_FLAGS, pattern=Pattern("exact_contains", "F")), Cond(C_DATE_CMP, op=CMP[">"], i64=(1700000000 + n // 8) * 1000000),
                  Cond(C_SLOT, pattern=Pattern("has_tag", "python"), field="Tags", mode=0),
                  Cond(C_BODY, pattern=Pattern("regex", r"react|angular", re.IGNORECASE))])
    prog, nq = pb.build(), 1
for _ in range(4):
    cnt = c.scan_count(prog, nq)
print(what, n, [int(x) for x in cnt[:4]], c.timing())
