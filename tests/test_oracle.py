"""CPU tests: the C restatement (oracle/fm_oracle.c) against (a) the committed golden vectors that
were produced by the unmodified reference and (b) the reference itself when oracle/_ref exists."""
import numpy as np
import pytest

from fuzzy_match_b200 import synth
from oracle import binding as ob
from tests.util import (REALTEXT_PARAM_SETS, as_tuples, csr, first_seen_ids, fix_params, load_golden, load_realtext,
                        load_subseq_golden)

CASES = load_golden()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_golden(case):
    tok, off = csr(case["tm"])
    q, qo = csr(case["queries"])
    O = ob.OracleIndex(tok, off, case["vocab_size"], max_tokens=case["max_tokens"])
    if "tm_real" in case:
        blob, ioff = synth.itok_table()
        O.set_real(case["tm_real"], case["tm_gaps"], off, blob, ioff)
        res, cnt = O.match_batch_real(q, case["q_real"], case["q_gaps"], qo, cap=64, **fix_params(case["params"]))
    else:
        res, cnt = O.match_batch(q, qo, cap=64, **fix_params(case["params"]))
    got = [as_tuples(r) for r in res]
    want = [[tuple(m) for m in r] for r in case["expected"]]
    assert got == want


def test_reference_known_answers_are_in_golden():
    """The gtest expectations themselves (test/test.cc), independent of how the golden file was made."""
    by = {c["name"]: c for c in CASES}
    f = lambda b: float(np.uint32(b).view(np.float32))
    lcs = by["lcs_cost"]["expected"][0]
    assert [m[0] for m in lcs] == [2, 1, 0]
    assert abs(f(lcs[0][1]) - 1.0) < 1e-3 and abs(f(lcs[1][1]) - 5 / 6) < 1e-3 and abs(f(lcs[2][1]) - 0.5) < 1e-3
    assert [len(r) for r in by["pre_reject"]["expected"]] == [2, 2]
    for name in ("idf_weight_1", "idf_weight_2_lcs", "idf_weight_2_unit"):
        r = by[name]["expected"][0]
        assert [m[0] for m in r] == [0, 1]
        assert abs(f(r[0][1]) - 0.6706515) < 1e-4 and abs(f(r[1][1]) - 0.6076691) < 1e-4
    r = by["contrastive_reduce_mean"]["expected"][0]
    assert [m[0] for m in r] == [0, 2, 1]
    assert [round(f(m[1]) - f(m[2]), 3) for m in r] == [0.667, 0.5, 0.125]
    r = by["contrastive_reduce_max"]["expected"][0]
    assert [round(f(m[1]) - f(m[2]), 3) for m in r] == [0.667, 0.5, -0.25]
    assert [m[0] for m in by["contrastive_buffer"]["expected"][0]] == [0, 3, 4]
    assert [[m[0] for m in q] for q in by["small_sentence_matches"]["expected"]] == [[0], [1], [2]]
    assert [len(q) for q in by["max_tokens_in_pattern"]["expected"]] == [0, 1]
    assert [m[0] for m in by["Q1_order_dependence_N1"]["expected"][0]] == [1]
    assert [m[0] for m in by["Q1_order_dependence_N2"]["expected"][0]] == [0, 1]


PARAM_SETS = [
    dict(fuzzy=0.8, n=1, ml=3, mr=0.3),
    dict(fuzzy=0.5, n=5, ml=2),
    dict(fuzzy=0.3, n=0, ml=2),
    dict(fuzzy=0.5, n=10, ml=3, idf=1.0, contrast=0.5),
    dict(fuzzy=0.4, n=4, ml=2, idf=0.7, costs=(1, 0, 1), contrast=0.5, reduce=1, buffer=8),
    dict(fuzzy=0.4, n=4, ml=2, costs=(0.5, 1.5, 1.2)),
]


@pytest.mark.skipif(not ob.ref_available(), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("params", PARAM_SETS, ids=[str(i) for i in range(len(PARAM_SETS))])
def test_oracle_matches_live_reference(params):
    tm, off, V = synth.make_tm(5000, vocab=800, seed=31)
    q, qo = synth.make_queries(tm, off, 300, vocab=800, seed=32)
    O = ob.OracleIndex(tm, off, V)
    R = ob.RefIndex(tm, off)
    ro, co = O.match_batch(q, qo, cap=32, **params)
    rr, cr = R.match_batch(q, qo, cap=32, **params)
    assert (co == cr).all()
    for a, b in zip(ro, rr):
        assert as_tuples(a) == as_tuples(b)


def test_oracle_threads_agree():
    tm, off, V = synth.make_tm(3000, vocab=500, seed=41)
    q, qo = synth.make_queries(tm, off, 200, vocab=500, seed=42)
    O = ob.OracleIndex(tm, off, V)
    r1, c1 = O.match_batch(q, qo, cap=8, nthreads=1, fuzzy=0.5, n=3)
    r4, c4 = O.match_batch(q, qo, cap=8, nthreads=4, fuzzy=0.5, n=3)
    assert (c1 == c4).all() and all(as_tuples(a, True) == as_tuples(b, True) for a, b in zip(r1, r4))


def test_oracle_rejects_bad_tm_tokens():
    with pytest.raises(ValueError):
        ob.OracleIndex(np.array([2, 1, 3], dtype=np.int32), np.array([0, 3], dtype=np.int64), 10)


@pytest.mark.skipif(not ob.ref_available(), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("params", [dict(fuzzy=0.5, n=5, ml=2), dict(fuzzy=0.3, n=0, ml=2, idf=1.0, costs=(1, 0, 1)),
                                    dict(fuzzy=0.4, n=4, ml=3, contrast=0.5, costs=(0.5, 1.5, 1.2))], ids=["0", "1", "2"])
def test_oracle_matches_live_reference_sentence_api(params):
    """Real tokens, case class and penalty tokens through add_tm(id, Sentence, Tokens) / match(Sentence, ...)."""
    blob, ioff = synth.itok_table()
    tm, off, V = synth.make_tm(3000, vocab=300, len_lo=1, len_hi=20, seed=71)
    q, qo = synth.make_queries(tm, off, 300, vocab=300, seed=72, len_lo=1, len_hi=20)
    real, gaps = synth.make_real(tm, off, 73)
    qreal, qgaps = synth.make_real(q, qo, 74)
    O = ob.OracleIndex(tm, off, V)
    O.set_real(real, gaps, off, blob, ioff)
    R = ob.RefIndex(tm, off, real=real, gaps=gaps, itok_blob=blob, itok_off=ioff)
    ro, co = O.match_batch_real(q, qreal, qgaps, qo, cap=32, **params)
    rr, cr = R.match_batch_real(q, qreal, qgaps, qo, cap=32, **params)
    assert (co == cr).all()
    for a, b in zip(ro, rr):
        assert as_tuples(a) == as_tuples(b)


COST_SETS = [(1, 1, 1), (1, 0, 1), (0.5, 1.5, 1.2), (0.4, 0.3, 1.2), (1.7, 1, 0.4), (1, 1, 2.5), (0.4, 0.4, 1.0)]


@pytest.mark.skipif(not ob.ref_available(), reason="oracle/_ref not built (no /root/reference here)")
def test_oracle_matches_live_reference_randomised():
    """Seeded sweep over TM shapes and every match() parameter (ml -1/0/1, mr above 1, fuzzy outside [0,1],
    N=0, buffer 0/1/large, idf, contrastive, no_perfect, junk query ids). The reference is driven through
    the Sentence overload with real == norm so that no_perfect is reachable."""
    blob, ioff = synth.itok_table()
    rng = np.random.default_rng(20251017)
    for trial in range(30):
        vocab = int(rng.choice([6, 20, 80, 600]))
        hi = int(rng.choice([3, 8, 20, 45]))
        n_sent = int(rng.choice([1, 40, 700, 2500]))
        tm, off, V = synth.make_tm(n_sent, vocab=vocab, len_lo=0, len_hi=hi, seed=1000 + trial)
        q, qo = synth.make_queries(tm, off, 60, vocab=vocab, seed=2000 + trial, len_lo=0, len_hi=hi)
        q = q.copy()
        if len(q):
            q[rng.integers(0, len(q), size=max(1, len(q) // 15))] = rng.choice([-3, 0, 1, V, V + 7, 2**31 - 1])
        max_tokens = int(rng.choice([300, hi, max(1, hi // 2)]))
        params = dict(fuzzy=float(rng.choice([-0.2, 0.0, 0.3, 0.5, 0.7, 0.9, 1.0, 1.1])), n=int(rng.choice([0, 1, 2, 5])),
                      ml=int(rng.choice([-1, 0, 1, 2, 3, 5])), mr=float(rng.choice([0.0, 0.3, 0.9, 1.5])),
                      idf=float(rng.choice([0.0, 0.0, 0.5, 2.0])), costs=COST_SETS[int(rng.integers(0, len(COST_SETS)))],
                      contrast=float(rng.choice([0.0, 0.0, 0.5, 1.0])), reduce=int(rng.integers(0, 2)),
                      buffer=int(rng.choice([-1, 0, 1, 3, 50])), no_perfect=bool(rng.integers(0, 2)))
        if params["idf"] and n_sent == 1:
            params["idf"] = 0.0  # log(1) == 0: the reference divides the idf weight by zero
        O = ob.OracleIndex(tm, off, V, max_tokens=max_tokens)
        real = (tm.astype(np.int64) * 2).astype(np.int32)
        R = ob.RefIndex(tm, off, max_tokens=max_tokens, real=real, gaps=np.zeros(len(tm) + len(off) - 1, dtype=np.int32),
                        itok_blob=blob, itok_off=ioff)
        qreal = (q.astype(np.int64) * 2 % (2**31)).astype(np.int32)
        ro, oc = O.match_batch(q, qo, cap=4096, **params)
        rr, cr = R.match_batch_real(q, qreal, np.zeros(len(q) + len(qo) - 1, dtype=np.int32), qo, cap=4096, **params)
        assert (oc == cr).all(), (trial, params)
        assert all(as_tuples(a) == as_tuples(b) for a, b in zip(ro, rr)), (trial, params)


def test_oracle_matches_reference_on_real_text():
    """Realistic text (the reference's test/data/tm2.en.gz as word ids + its test-tm2.en queries,
    tests/golden/realtext.npz): the restatement against what the unmodified reference returned."""
    tm, off, V, q, qo, expected = load_realtext()
    O = ob.OracleIndex(tm, off, V)
    for params, want in zip(REALTEXT_PARAM_SETS, expected):
        res, cnt = O.match_batch(q, qo, cap=256, nthreads=8, **params)
        assert [as_tuples(r) for r in res] == want, params


# ---- subsequence() (src/fuzzy_match.cc:238-365): the reference's tests never call it, so the pin is the
# reference itself -- committed outputs (tests/golden/subseq.json) and, where oracle/_ref exists, live runs.
SUBSEQ = load_subseq_golden()


def subseq_tuples(rec):
    return [[int(r["found"]), int(r["s_id"]), int(r["score"].view(np.uint32)), int(r["length"]), int(r["position"])] if r["found"]
            else [0, 0, 0, 0, 0] for r in rec]


@pytest.mark.parametrize("case", SUBSEQ["cases"], ids=[c["name"] for c in SUBSEQ["cases"]])
def test_oracle_subsequence_matches_golden(case):
    t = SUBSEQ["tms"][case["tm"]]
    tok, off = csr(t["tm"])
    q, qo = csr(t["queries"])
    O = ob.OracleIndex(tok, off, t["vocab_size"])
    assert subseq_tuples(O.subsequence_batch(q, qo, **case["params"])) == case["expected"]


@pytest.mark.skipif(not ob.ref_available(), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("seed,vocab", [(0, 8), (1, 30), (2, 2000), (3, 12)])
def test_oracle_subsequence_matches_live_reference(seed, vocab):
    tm, off, _ = synth.make_tm(3000, vocab=vocab, seed=seed)
    q, qo = synth.make_queries(tm, off, 300, vocab=vocab, seed=seed + 100)
    tm, q, V = first_seen_ids(tm, q)
    O, R = ob.OracleIndex(tm, off, V), ob.RefIndex(tm, off)
    for kw in (dict(n=1), dict(n=5, no_perfect=True), dict(n=3, ml=2, mr=0.0, idf_weighting=True),
               dict(n=50, ml=1, mr=0.5, no_perfect=True, idf_weighting=True), dict(n=0), dict(n=2, ml=40, mr=0.0)):
        a = O.subsequence_batch(q, qo, **kw)
        b, texts = R.subsequence_batch(q, qo, **kw)
        assert (a["found"] == b["found"]).all()
        for i in np.nonzero(a["found"])[0]:
            assert a[i]["s_id"] == b[i]["s_id"] and a[i]["score"].tobytes() == b[i]["score"].tobytes() and a[i]["length"] == b[i]["max_subseq"]
            lo = qo[i] + a[i]["position"]
            assert " ".join(str(x) for x in q[lo:lo + a[i]["length"]]) == texts[i]  # what the reference appends to Match::id


# ---- match() into a non-empty result vector (src/fuzzy_match.cc:626-679): earlier entries count against
# number_of_matches and the contrastive rerank penalises the candidates against them
PRIOR_PARAM_PAIRS = [(dict(fuzzy=0.5, n=3, ml=2), dict(fuzzy=0.4, n=6, ml=2, contrast=0.5)),
                     (dict(fuzzy=0.5, n=2, ml=2, contrast=0.3), dict(fuzzy=0.3, n=5, ml=2, contrast=0.8, reduce=1, buffer=10)),
                     (dict(fuzzy=0.5, n=2, ml=2), dict(fuzzy=0.3, n=4, ml=2)),
                     (dict(fuzzy=0.5, n=4, ml=2), dict(fuzzy=0.3, n=3, ml=2, contrast=0.5)),
                     (dict(fuzzy=0.8, n=0, ml=2), dict(fuzzy=0.75, n=0, ml=3, contrast=0.5, idf=1.0))]


@pytest.mark.skipif(not ob.ref_available(), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("seed,vocab", [(0, 30), (1, 200), (2, 12)])
def test_oracle_prior_matches_vs_live_reference(seed, vocab):
    tm, off, V = synth.make_tm(3000, vocab=vocab, seed=seed)
    q1, q1o = synth.make_queries(tm, off, 100, vocab=vocab, seed=seed + 100)
    q2, q2o = synth.make_queries(tm, off, 100, vocab=vocab, seed=seed + 100, p_sub=0.2)  # same sources, perturbed differently
    O, R = ob.OracleIndex(tm, off, V), ob.RefIndex(tm, off)
    both = 0
    for p1, p2 in PRIOR_PARAM_PAIRS:
        pri, pcnt, out, cnt = R.match_batch_twice(q1, q1o, q2, q2o, p1, p2, cap=64)
        poff = np.zeros(len(pcnt) + 1, dtype=np.int64)
        np.cumsum(pcnt, out=poff[1:])
        psid = np.array([m["s_id"] for r in pri for m in r], dtype=np.uint32)
        o, ocnt = O.match_batch_prior(q2, q2o, psid, poff, cap=64, **p2)
        assert (cnt == ocnt).all()
        assert [as_tuples(r) for r in out] == [as_tuples(r) for r in o]
        both += int(((pcnt > 0) & (cnt > 0)).sum())
    assert both > 100  # the case under test occurs
