"""One parity case in a fresh process, so that the library's environment switches (read once per
process) take effect:  python -m tests.gpu_env_case <case>
Exit code 0 = bit-identical to the oracle. Used by tests/test_gpu_scale.py::test_env_gated_paths."""
import sys

import numpy as np

import fuzzy_match_b200 as fmb
from fuzzy_match_b200 import synth
from oracle import binding as ob


def same(index, oracle, q, qo, cap, **params):
    out, cnt = index.match_batch(q, qo, cap=cap, **params)
    ro, oc = oracle.match_batch(q, qo, cap=cap, nthreads=8, **params)
    if not (cnt == oc).all():
        print("counts differ at", np.nonzero(cnt != oc)[0][:10], params)
        return False
    for i in range(len(oc)):
        if out[i, :min(cnt[i], cap)].tobytes() != ro[i].tobytes():
            print("query", i, "differs", params)
            return False
    return True


def many_candidates():
    """Tiny vocabulary: hundreds to thousands of scored candidates per query (sort / replay tiers)."""
    tm, off, V = synth.make_tm(3000, vocab=40, len_lo=1, len_hi=30, seed=61)
    q, qo = synth.make_queries(tm, off, 200, vocab=40, seed=62, len_lo=1, len_hi=30)
    index, oracle = fmb.Index(tm, off, V), ob.OracleIndex(tm, off, V)
    ok = True
    for params in (dict(fuzzy=0.5, n=3, ml=2), dict(fuzzy=0.2, n=0, ml=1), dict(fuzzy=0.6, n=2, ml=3, costs=(1, 0, 1)),
                   dict(fuzzy=0.3, n=5, ml=2, idf=1.0, contrast=0.5, buffer=20)):
        ok &= same(index, oracle, q, qo, 3000, **params)
    return ok


def all_scoring_paths():
    """Short and long patterns, unit / general costs, IDF: whichever DP kernels the environment selects."""
    tm, off, V = synth.make_tm(6000, vocab=900, len_lo=1, len_hi=70, seed=71)
    q, qo = synth.make_queries(tm, off, 500, vocab=900, seed=72, len_lo=1, len_hi=70)
    index, oracle = fmb.Index(tm, off, V), ob.OracleIndex(tm, off, V)
    ok = True
    for params in (dict(fuzzy=0.5, n=4, ml=2), dict(fuzzy=0.4, n=3, ml=3, idf=1.0), dict(fuzzy=0.3, n=3, ml=2, costs=(1, 0, 1)),
                   dict(fuzzy=0.4, n=4, ml=2, costs=(0.5, 1.5, 1.2)), dict(fuzzy=0.4, n=4, ml=2, costs=(2.5, 2.5, 2.5)),
                   dict(fuzzy=0.6, n=2, ml=2, no_perfect=True)):
        ok &= same(index, oracle, q, qo, 16, **params)
    return ok


def short_patterns():
    """Sentences of at most 40 words (no wide signatures): the thread-per-query prepare kernel, patterns of up to 60
    words and repeated words for the warp kernel it hands over to."""
    tm, off, V = synth.make_tm(8000, vocab=300, len_lo=1, len_hi=40, seed=81)
    q, qo = synth.make_queries(tm, off, 1500, vocab=300, seed=82, len_lo=1, len_hi=60)
    index, oracle = fmb.Index(tm, off, V), ob.OracleIndex(tm, off, V)
    ok = True
    for params in (dict(fuzzy=0.5, n=3, ml=2), dict(fuzzy=0.7, n=1, ml=3), dict(fuzzy=0.4, n=4, ml=3, mr=0.3, idf=1.0)):
        ok &= same(index, oracle, q, qo, 8, **params)
    return ok


CASES = {"many_candidates": many_candidates, "all_scoring_paths": all_scoring_paths, "short_patterns": short_patterns}

if __name__ == "__main__":
    ob.build()
    ok = CASES[sys.argv[1]]()
    print("identical" if ok else "MISMATCH")
    sys.exit(0 if ok else 1)
