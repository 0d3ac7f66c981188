"""Generates tests/golden/realtext.npz: a realistic-text differential fixture.

Run in the build container only (needs /root/reference and oracle/_ref/libfm_ref.so):

    make -C oracle all && python tests/golden/make_realtext.py

TM      = the 20 000 already tokenised Europarl sentences of the reference's test/data/tm2.en.gz,
          whitespace-split, words replaced by vocabulary ids in first-occurrence order from 2
          (src/vocab_indexer.cc:37-50) -- only the ids are stored, not the text.
queries = the 100 queries of test/data/test-tm2.en (unseen words become distinct out-of-vocabulary ids) plus
          400 TM sentences perturbed like the synthetic workload (5 % delete / replace / insert per token).
expected = what the UNMODIFIED reference (fuzzy::FuzzyMatch::match(Tokens), oracle/_ref) returns for each
          parameter set in PARAM_SETS, floats as uint32 bit patterns.
The expected scores of test/data/test-tm2 itself are not usable here: they depend on the OpenNMT tokenizer's
case / number / placeholder features (out of scope).
"""
import gzip
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fuzzy_match_b200 import synth  # noqa: E402
from oracle import binding as ob  # noqa: E402
from tests.util import REALTEXT_PARAM_SETS as PARAM_SETS  # noqa: E402

DATA = "/root/reference/test/data"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "realtext.npz")

def main():
    voc = {}
    tm_ids = []
    with gzip.open(os.path.join(DATA, "tm2.en.gz"), "rt", encoding="utf-8") as f:
        for line in f:
            tm_ids.append([voc.setdefault(w, len(voc) + 2) for w in line.split()])
    V = len(voc) + 2
    tm_off = np.zeros(len(tm_ids) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in tm_ids], out=tm_off[1:])
    tm_tok = np.array([t for s in tm_ids for t in s], dtype=np.int32)
    q_ids = []
    n_oov = 0
    with open(os.path.join(DATA, "test-tm2.en"), encoding="utf-8") as f:
        for line in f:
            ids = []
            for w in line.split():
                if w in voc:
                    ids.append(voc[w])
                else:  # any id >= vocab_size is "unknown" at the boundary
                    ids.append(V + n_oov)
                    n_oov += 1
            q_ids.append(ids)
    pq, pqo = synth.make_queries(tm_tok, tm_off, 400, vocab=V - 2, seed=777, frac_random=0.0)
    q_ids += [pq[pqo[i]:pqo[i + 1]].tolist() for i in range(400)]
    q_off = np.zeros(len(q_ids) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in q_ids], out=q_off[1:])
    q_tok = np.array([t for s in q_ids for t in s], dtype=np.int32)

    assert ob.ref_available(), "build oracle/_ref first (make -C oracle all)"
    R = ob.RefIndex(tm_tok, tm_off, max_tokens=300)
    out = dict(tm_tok=tm_tok, tm_off=tm_off, q_tok=q_tok, q_off=q_off, vocab_size=np.int64(V), n_param_sets=np.int64(len(PARAM_SETS)))
    cap = 256
    for k, params in enumerate(PARAM_SETS):
        res, cnt = R.match_batch(q_tok, q_off, cap=cap, nthreads=os.cpu_count() or 1, **params)
        assert cnt.max() <= cap
        flat = np.concatenate([r for r in res]) if len(res) else np.zeros(0, dtype=ob.REF_MATCH_DTYPE)
        out["cnt_%d" % k] = cnt.astype(np.int32)
        out["s_id_%d" % k] = flat["s_id"].astype(np.uint32)
        out["score_%d" % k] = flat["score"].view(np.uint32)
        out["penalty_%d" % k] = flat["penalty"].view(np.uint32)
        out["lm_%d" % k] = flat["max_subseq"].astype(np.int32)
        out["len_%d" % k] = flat["length"].astype(np.int32)
        print("param set %d: %d of %d queries matched, %d matches" % (k, int((cnt > 0).sum()), len(cnt), len(flat)))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes; vocab", V, "tokens", len(tm_tok), "queries", len(q_ids), "oov", n_oov)


if __name__ == "__main__":
    main()
