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
    """Equal as ordered lists up to permutations among entries of one (folder, status, timestamp) group:
    os.listdir order is file-system specific and the reference's sort is stable on timestamp only."""
    if len(a) != len(b):
        return False
    norm = lambda lst: [(k[0], k[1], ts_of(k)) for k in lst]
    return norm(a) == norm(b) and sorted(map(tuple, a)) == sorted(map(tuple, b))
