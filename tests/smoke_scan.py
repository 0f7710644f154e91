"""One small invocation of the scan path on cuda:0, checked against the oracle (called by __graft_entry__.smoke)."""
import re

import numpy as np


def run() -> None:
    from fei_b200 import synth
    from fei_b200.corpus import Corpus
    from fei_b200.program import C_BODY, C_FLAGS, C_SLOT, Cond, ProgramBuilder
    from fei_b200.regexc import Pattern
    from oracle import memdir_oracle as mo

    n = 1500
    c = Corpus().synth(0xFE1, 0, n)                       # generated + tiled on the GPU
    recs = [synth.record(0xFE1, i) for i in range(n)]      # same records from the host generator
    mems = [mo.make_memory(r["filename"], r["folder"], r["status"], synth.file_text(r), True) for r in recs]
    pb = ProgramBuilder()
    pats = ["python", "docker|kubernetes", r"kubernetes.*docker|docker.*kubernetes", r"vue\.js"]
    for p in pats:
        pb.add_query([Cond(C_BODY, pattern=Pattern("regex", p, re.IGNORECASE))])
    pb.add_query([Cond(C_SLOT, pattern=Pattern("has_tag", "python"), field="Tags"),
                  Cond(C_FLAGS, pattern=Pattern("exact_contains", "F")),
                  Cond(C_BODY, pattern=Pattern("regex", "react|angular", re.IGNORECASE))])
    hits = c.scan_hits(pb.build(), len(pats) + 1)
    for q, p in enumerate(pats):
        want = mo.run_search(mems, [{"field": "content", "operator": "matches", "value": p}])
        assert hits[q].tolist() == want, p
    want = mo.run_search(mems, [{"field": "Tags", "operator": "has_tag", "value": "python"}, {"field": "flags", "operator": "has_flag", "value": "F"},
                                {"field": "content", "operator": "matches", "value": "react|angular"}])
    assert hits[len(pats)].tolist() == want
    c.close()
