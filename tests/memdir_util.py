"""Rebuild the golden Memdir tree (seeded synthetic records + adversarial raw files) in a temp dir."""
import base64
import os

from tests.conftest import load_golden


def build_tree(base):
    from fei_b200 import synth
    g = load_golden("memdir_golden.json")
    recs = [synth.record(g["seed"], i) for i in range(g["n_synth"])]
    synth.write_memdir(base, recs)
    for folder, status, name, raw64 in g["adversarial"]:
        d = os.path.join(base, folder, status) if folder else os.path.join(base, status)
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, name), "wb") as f:
            f.write(base64.b64decode(raw64))
    return g


def key_of(m):
    return [m["folder"], m["status"], m["filename"]]


def ts_of(key):
    return int(key[2].split(".")[0])


def same_modulo_ties(a, b):
    """Equal up to what is file-system specific: the order in which os.walk / os.listdir enumerate
    directories and files.  Inside one (folder, status) directory the reference's order is newest
    timestamp first with ties in listdir order; across directories it is os.walk order.  So: same
    multiset, and inside every (folder, status) group the same timestamp sequence, groups contiguous."""
    if sorted(map(tuple, a)) != sorted(map(tuple, b)):
        return False

    def groups(lst):
        out, seen = {}, []
        for k in lst:
            g = (k[0], k[1])
            if g not in out:
                out[g] = []
                seen.append(g)
            elif seen[-1] != g:
                return None                     # a (folder, status) group must be contiguous
            out[g].append(ts_of(k))
        return out
    ga, gb = groups(a), groups(b)
    return ga is not None and gb is not None and ga == gb
