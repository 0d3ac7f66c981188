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


# match(Tokens) parameter sets run through the live reference (it has no no_perfect argument)
REALTEXT_PARAM_SETS = [
    dict(fuzzy=0.8, n=5, ml=3, mr=0.3),                       # CLI defaults
    dict(fuzzy=0.5, n=2, ml=3, mr=0.3),                       # the reference's own tm2 test (test/test.cc:217-221)
    dict(fuzzy=0.7, n=1, ml=3),                               # BASELINE config 2 parameters
    dict(fuzzy=0.5, n=1, ml=3),                               # config 3
    dict(fuzzy=0.7, n=10, ml=3, idf=1.0, contrast=0.5),       # config 5
    dict(fuzzy=0.4, n=4, ml=2, idf=0.7, costs=(1, 0, 1), contrast=0.5, reduce=1, buffer=8),
    dict(fuzzy=0.4, n=4, ml=2, costs=(0.5, 1.5, 1.2)),
    dict(fuzzy=0.3, n=0, ml=2),
]


def load_realtext():
    """tests/golden/realtext.npz (made by tests/golden/make_realtext.py from the live reference)."""
    d = np.load(os.path.join(os.path.dirname(GOLDEN), "realtext.npz"))
    expected = []
    for k in range(int(d["n_param_sets"])):
        cnt = d["cnt_%d" % k]
        off = np.concatenate([[0], np.cumsum(cnt)])
        rows = list(zip(d["s_id_%d" % k].tolist(), d["score_%d" % k].tolist(), d["penalty_%d" % k].tolist(),
                        d["lm_%d" % k].tolist(), d["len_%d" % k].tolist()))
        expected.append([rows[off[i]:off[i + 1]] for i in range(len(cnt))])
    return d["tm_tok"], d["tm_off"], int(d["vocab_size"]), d["q_tok"], d["q_off"], expected


def first_seen_ids(tm_tokens, q_tokens):
    """Relabels word ids in first-seen order over the TM, the order in which VocabIndexer::addWords hands them out
    (reference src/vocab_indexer.cc:37-50); query words the TM does not hold get ids beyond the vocabulary.
    Returns (tm, queries, vocab_size). subsequence() observes the walk order inside a suffix-array range, which
    follows the word ids, so comparisons with the live reference need the reference's own numbering."""
    ids = {}
    tm = np.empty(len(tm_tokens), dtype=np.int32)
    for i, t in enumerate(np.asarray(tm_tokens).tolist()):
        tm[i] = ids.setdefault(t, len(ids) + 2)
    vocab_size = len(ids) + 2
    q = np.array([ids.get(t, vocab_size + 7) for t in np.asarray(q_tokens).tolist()], dtype=np.int32)
    return tm, q, vocab_size


def load_subseq_golden():
    """tests/golden/subseq.json (made by tests/golden/make_subseq.py from the live reference)."""
    with open(os.path.join(os.path.dirname(GOLDEN), "subseq.json")) as f:
        return json.load(f)
