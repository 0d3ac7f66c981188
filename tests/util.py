"""Shared helpers for the parity tests."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kat.json")


def load_golden():
    with open(GOLDEN) as f:
        return json.load(f)


def csr(lists):
    off = np.zeros(len(lists) + 1, dtype=np.int64)
    if lists:
        np.cumsum([len(x) for x in lists], out=off[1:])
    tok = np.array([t for x in lists for t in x], dtype=np.int32)
    return tok, off


def bits(a):
    a = np.asarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def as_tuples(matches, with_cost=False):
    """Structured match array -> list of exact (bit-level) tuples."""
    out = []
    for m in matches:
        t = [int(m["s_id"]), int(bits(m["score"])), int(bits(m["penalty"])), int(m["max_subseq"]), int(m["length"])]
        if with_cost:
            t.append(int(bits(m["cost"])))
        out.append(tuple(t))
    return out


def fix_params(p):
    p = dict(p)
    if "costs" in p:
        p["costs"] = tuple(p["costs"])
    return p
