"""Generates tests/golden/kat.json by running the UNMODIFIED reference (oracle/_ref/libfm_ref.so,
built from /root/reference by oracle/Makefile). Run in the build container only:

    make -C oracle all && python tests/golden/make_golden.py

Cases: the reference's tokenizer-free gtest known answers (test/test.cc:223-262, 273-303, 337-632,
inputs transcribed as word ids), the order-dependence vectors Q1/Q2 of SURVEY.md section 3.1, and
seeded random TMs under many parameter sets. Floats are stored as uint32 bit patterns.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fuzzy_match_b200 import synth  # noqa: E402
from oracle import binding as ob  # noqa: E402


def words_case(name, tm, queries, src, **params):
    voc = {}

    def ids(s):
        return [voc.setdefault(w, len(voc) + 2) for w in s.split()]

    tm_ids = [ids(s) for s in tm]
    unk = 1000
    q_ids = [[voc.get(w, unk + i) for i, w in enumerate(q.split())] for q in queries]
    return dict(name=name, source=src, vocab_size=2000, max_tokens=params.pop("max_tokens", 300), tm=tm_ids,
                queries=q_ids, params=params)


def csr(lists):
    off = np.zeros(len(lists) + 1, dtype=np.int64)
    np.cumsum([len(x) for x in lists], out=off[1:])
    tok = np.array([t for x in lists for t in x], dtype=np.int32)
    return tok, off


def uncsr(tok, off):
    return [tok[off[i]:off[i + 1]].tolist() for i in range(len(off) - 1)]


def run_ref(case):
    tok, off = csr(case["tm"])
    q, qo = csr(case["queries"])
    if "tm_real" in case:  # Sentence API: real tokens + penalty tokens
        blob, ioff = synth.itok_table()
        R = ob.RefIndex(tok, off, max_tokens=case["max_tokens"], real=np.array(case["tm_real"], dtype=np.int32),
                        gaps=np.array(case["tm_gaps"], dtype=np.int32), itok_blob=blob, itok_off=ioff)
        res, cnt = R.match_batch_real(q, np.array(case["q_real"], dtype=np.int32), np.array(case["q_gaps"], dtype=np.int32), qo,
                                      cap=64, **case["params"])
    else:
        R = ob.RefIndex(tok, off, max_tokens=case["max_tokens"])
        res, cnt = R.match_batch(q, qo, cap=64, **case["params"])
    case["expected"] = [[[int(m["s_id"]), int(m["score"].view(np.uint32)), int(m["penalty"].view(np.uint32)),
                          int(m["max_subseq"]), int(m["length"])] for m in r] for r in res]
    assert all(c <= 64 for c in cnt)
    return case


def main():
    cases = []
    T = "test/test.cc"
    small = ["single", "two words", "three kind words"]
    cases.append(words_case("small_sentence_matches", small, ["single", "two words", "three kind words"], T + ":223-262",
                            fuzzy=1.0, n=1, ml=3))
    cases.append(words_case("max_tokens_in_pattern", small, ["three kind words", "two words"], T + ":273-303",
                            fuzzy=1.0, n=1, ml=3, max_tokens=2))
    cases.append(words_case("lcs_cost", ["a b c", "a b c d e x x x", "x x a b c d e f x x x x x"], ["a b c d e f"],
                            T + ":337-375", fuzzy=0.0, n=10, ml=3, mr=0.5, costs=(1, 0, 1)))
    pr = ["a b c d e", "a b c d e f", "a b c d e f g"]
    cases.append(words_case("pre_reject", pr, ["a b c", "a b c d e f g h i j k l"], T + ":377-418", fuzzy=0.5, n=10, ml=0))
    cases.append(words_case("idf_weight_1", ["a b c", "a b d", "d d d d d", "d e", "c"], ["a b c d"], T + ":420-452",
                            fuzzy=0.0, n=10, ml=0, idf=1.0, costs=(1, 0, 1)))
    idf2 = ["a b c e", "a b e d", "d d d d d", "d e", "c"]
    cases.append(words_case("idf_weight_2_lcs", idf2, ["a b c d"], T + ":454-486", fuzzy=0.0, n=10, ml=0, idf=1.0, costs=(1, 0, 1)))
    cases.append(words_case("idf_weight_2_unit", idf2, ["a b c d"], T + ":487-507", fuzzy=0.0, n=10, ml=0, idf=1.0))
    con = ["a b c d", "b c d", "d e f"]
    cases.append(words_case("contrastive_reduce_mean", con, ["a b c d e f"], T + ":509-548", fuzzy=0.0, n=10, ml=0, contrast=1.0))
    cases.append(words_case("contrastive_reduce_max", con, ["a b c d e f"], T + ":550-590", fuzzy=0.0, n=10, ml=0, contrast=1.0, reduce=1))
    cases.append(words_case("contrastive_buffer", ["a b c d e", "b c d e", "c d e f", "d e f g", "h i j"],
                            ["a b c d e f g h i j"], T + ":592-632", fuzzy=0.0, n=3, ml=0, costs=(1, 0, 1), contrast=1.0,
                            reduce=1, buffer=10))
    # the running cost bound makes the result depend on candidate order
    P = " ".join("p%d" % i for i in range(30))
    s0 = " ".join(["p%d" % i for i in range(20)] + ["x%d" % i for i in range(20, 30)])
    s1 = " ".join(["p%d" % i for i in range(30)] + ["y%d" % i for i in range(15)])
    for n in (1, 2):
        cases.append(words_case("Q1_order_dependence_N%d" % n, [s0, s1], [P], "SURVEY.md 3.1 Q1", fuzzy=0.5, n=n, ml=3))
    # Q2: the early exit ignores column 0 (delete cost 0)
    for n in (1, 2, 3):
        cases.append(words_case("Q2_column0_N%d" % n, ["q a b z c d", "a b c d", "a b z c d"], ["a b c d"], "SURVEY.md 3.1 Q2",
                                fuzzy=0.5, n=n, ml=2, costs=(1, 0, 1)))
    # edge cases: empty pattern, single-token patterns, all-unknown pattern, repeated words
    cases.append(dict(name="edge_patterns", source="edge cases", vocab_size=50, max_tokens=300,
                      tm=[[2], [2, 2, 2], [3, 4], [5, 6, 7, 8], [2, 3, 4, 5, 6, 7, 8, 9], [9, 9, 9, 9]],
                      queries=[[], [2], [3], [40], [40, 41, 42], [2, 2], [2, 2, 2, 2], [9, 9], [3, 4, 5], [2, 3, 4, 5, 6, 7, 8, 9]],
                      params=dict(fuzzy=0.3, n=0, ml=1)))
    # seeded random TMs (small vocabulary => many candidates, duplicates and ties)
    param_sets = [
        dict(fuzzy=0.8, n=1, ml=3, mr=0.3),
        dict(fuzzy=0.7, n=1, ml=3),
        dict(fuzzy=0.5, n=5, ml=2),
        dict(fuzzy=0.3, n=0, ml=2),
        dict(fuzzy=0.5, n=10, ml=3, idf=1.0, contrast=0.5),
        dict(fuzzy=0.4, n=4, ml=2, idf=0.7, costs=(1, 0, 1), contrast=0.5, reduce=1, buffer=8),
        dict(fuzzy=0.4, n=4, ml=2, costs=(0.5, 1.5, 1.2)),
        dict(fuzzy=0.4, n=3, ml=2, costs=(0.4, 0.3, 1.2), idf=2.0),
        dict(fuzzy=0.6, n=2, ml=0, costs=(1, 0, 1)),
        dict(fuzzy=0.0, n=3, ml=4, mr=0.2, buffer=1),
    ]
    tm, off, V = synth.make_tm(400, vocab=60, len_lo=1, len_hi=20, seed=11)
    q, qo = synth.make_queries(tm, off, 40, vocab=60, seed=12, len_lo=1, len_hi=20)
    for i, ps in enumerate(param_sets):
        cases.append(dict(name="random_small_%d" % i, source="synth seed 11/12", vocab_size=V, max_tokens=300,
                          tm=uncsr(tm, off), queries=uncsr(q, qo), params=ps))
    # long sentences (multi-row-per-lane DP on the GPU), cap lowered so some TM sentences are dropped
    tm, off, V = synth.make_tm(60, vocab=300, len_lo=40, len_hi=120, seed=21)
    q, qo = synth.make_queries(tm, off, 12, vocab=300, seed=22, len_lo=40, len_hi=120)
    cases.append(dict(name="random_long", source="synth seed 21/22", vocab_size=V, max_tokens=100, tm=uncsr(tm, off),
                      queries=uncsr(q, qo), params=dict(fuzzy=0.5, n=3, ml=3)))
    # Sentence API (real tokens, case class, penalty tokens): match(const Sentence&, const Tokens&, ...)
    real_params = [dict(fuzzy=0.5, n=5, ml=2), dict(fuzzy=0.3, n=20, ml=2, idf=1.0, costs=(1, 0, 1)),
                   dict(fuzzy=0.4, n=4, ml=3, contrast=0.5, costs=(0.5, 1.5, 1.2)), dict(fuzzy=0.6, n=2, ml=2, no_perfect=True),
                   dict(fuzzy=0.7, n=1, ml=3, mr=0.3)]
    tm, off, V = synth.make_tm(300, vocab=50, len_lo=1, len_hi=18, seed=31)
    q, qo = synth.make_queries(tm, off, 30, vocab=50, seed=32, len_lo=1, len_hi=18)
    real, gaps = synth.make_real(tm, off, 33)
    qreal, qgaps = synth.make_real(q, qo, 34)
    for i, ps in enumerate(real_params):
        cases.append(dict(name="sentence_api_%d" % i, source="synth seed 31-34, include/fuzzy/fuzzy_match.hh:53,70-82",
                          vocab_size=V, max_tokens=300, tm=uncsr(tm, off), queries=uncsr(q, qo), tm_real=real.tolist(),
                          tm_gaps=gaps.tolist(), q_real=qreal.tolist(), q_gaps=qgaps.tolist(), params=ps))
    out = [run_ref(c) for c in cases]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "cases")
    for c in out[:17]:
        print(c["name"], [[(m[0], round(float(np.uint32(m[1]).view(np.float32)), 4)) for m in r] for r in c["expected"]])


if __name__ == "__main__":
    main()
