"""Generates tests/golden/subseq.json by running FuzzyMatch::subsequence of the UNMODIFIED reference
(oracle/_ref/libfm_ref.so, built from /root/reference by oracle/Makefile). Run in the build container only:

    make -C oracle all && python tests/golden/make_subseq.py

The reference's test suite holds no subsequence() expectations (test/test.cc never calls it), so the vectors are
outputs of the reference itself: seeded random TMs whose word ids are assigned in first-seen order (what
VocabIndexer::addWords does, src/vocab_indexer.cc:37-50 -- the walk order inside a suffix-array range, which
subsequence() observes, depends on it), a repetitive tiny-vocabulary TM (many equally good sentences, long ranges),
and a TM with exact copies of the queries for no_perfect. expected[q] = [found, s_id, score bits, max_subseq, position]
where position is recovered from the text the reference appends to Match::id.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fuzzy_match_b200 import synth  # noqa: E402
from oracle import binding as ob  # noqa: E402
from tests.util import first_seen_ids  # noqa: E402

PARAMS = [dict(n=1), dict(n=5, no_perfect=True), dict(n=3, ml=2, mr=0.0, idf_weighting=True),
          dict(n=50, ml=1, mr=0.5, no_perfect=True, idf_weighting=True), dict(n=2, ml=4, mr=0.0)]


def position_of(text, pattern, length):
    """First position at which the detokenised sub-sequence (decimal ids joined by blanks) occurs in the pattern; the
    reference tries equal-weight sub-sequences by ascending position, so the first occurrence is the one it used."""
    want = [int(x) for x in text.split()]
    assert len(want) == length
    for i in range(len(pattern) - length + 1):
        if list(pattern[i:i + length]) == want:
            return i
    raise AssertionError("sub-sequence text not found in the pattern")


def main():
    cases = []
    for seed, (n_sent, vocab, n_q) in enumerate([(400, 8, 60), (600, 40, 60), (800, 500, 60), (500, 5000, 40)]):
        tm, off, _ = synth.make_tm(n_sent, vocab=vocab, seed=900 + seed)
        q, qo = synth.make_queries(tm, off, n_q, vocab=vocab, seed=950 + seed)
        tm, q, V = first_seen_ids(tm, q)
        R = ob.RefIndex(tm, off)
        for p in PARAMS:
            rec, texts = R.subsequence_batch(q, qo, **p)
            exp = []
            for i in range(n_q):
                if not rec[i]["found"]:
                    exp.append([0, 0, 0, 0, 0])
                    continue
                pos = position_of(texts[i], q[qo[i]:qo[i + 1]], int(rec[i]["max_subseq"]))
                exp.append([1, int(rec[i]["s_id"]), int(rec[i]["score"].view(np.uint32)), int(rec[i]["max_subseq"]), pos])
            cases.append(dict(name="subseq_seed%d_%s" % (seed, "_".join("%s%s" % (k, v) for k, v in sorted(p.items()))),
                              vocab_size=V, tm=[tm[off[i]:off[i + 1]].tolist() for i in range(n_sent)],
                              queries=[q[qo[i]:qo[i + 1]].tolist() for i in range(n_q)], params=p, expected=exp))
    # share the TM / query lists between the cases of one seed to keep the file small
    out = dict(generator="tests/golden/make_subseq.py", reference="src/fuzzy_match.cc:238-365 through oracle/ref_driver.cc", tms=[], cases=[])
    for c in cases:
        key = (c["tm"], c["queries"])
        for k, t in enumerate(out["tms"]):
            if (t["tm"], t["queries"]) == key:
                break
        else:
            k = len(out["tms"])
            out["tms"].append(dict(vocab_size=c["vocab_size"], tm=c["tm"], queries=c["queries"]))
        out["cases"].append(dict(name=c["name"], tm=k, params=c["params"], expected=c["expected"]))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "subseq.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote %s: %d cases over %d TMs, %d bytes" % (path, len(out["cases"]), len(out["tms"]), os.path.getsize(path)))


if __name__ == "__main__":
    main()
