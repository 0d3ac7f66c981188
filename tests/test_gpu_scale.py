"""GPU parity at BASELINE.json scale and on the rarely taken code paths (run with -m gpu on the B200 box).

Every BASELINE configuration is run on a 1M-sentence TM through the C ABI and compared bit for bit with the
C restatement on a sample of >= 1000 queries (>= 300 for the 200-300 token patterns), and with the
reference itself (oracle/_ref/libfm_ref.so, the reference's own sources; it travels to the box as a built
library) where the reference's Tokens API can express the call. Realistic text (the reference's
test/data/tm2.en.gz as word ids, tests/golden/realtext.npz) is compared with results the unmodified
reference produced in the build container. Environment-gated kernels run in fresh subprocesses.
"""
import os
import subprocess
import sys
import time

import numpy as np
import pytest

import fuzzy_match_b200 as fmb
from fuzzy_match_b200 import synth
from oracle import binding as ob
from tests.util import REALTEXT_PARAM_SETS, as_tuples, load_realtext

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NT = os.cpu_count() or 1


def compare(index, checker, q, qo, cap, with_cost=True, **params):
    out, cnt = index.match_batch(q, qo, cap=cap, **params)
    ro, oc = checker.match_batch(q, qo, cap=cap, nthreads=NT, **params)
    assert (cnt == oc).all(), "match counts differ at queries %s" % np.nonzero(cnt != oc)[0][:10]
    got = [as_tuples(out[i, :min(cnt[i], cap)], with_cost) for i in range(len(cnt))]
    want = [as_tuples(r, with_cost) for r in ro]
    bad = [i for i in range(len(want)) if got[i] != want[i]]
    assert not bad, "query %d: gpu %s != checker %s" % (bad[0], got[bad[0]], want[bad[0]])
    return cnt


@pytest.fixture(scope="module")
def tm1m():
    """BASELINE.json configs[1] TM: 1M sentences, Zipf(1) over 50k words, len U[5,25]."""
    tm, off, V = synth.make_tm(1000000, seed=1234)
    q, qo = synth.make_queries(tm, off, 20000, seed=5678)
    return dict(tm=tm, off=off, V=V, q=q, qo=qo, index=fmb.Index(tm, off, V), oracle=ob.OracleIndex(tm, off, V))


@pytest.fixture(scope="module")
def ref1m(tm1m):
    """The reference itself on the same 1M-sentence TM (add_tm x 1M + sort: ~15 s)."""
    if not ob.ref_available():
        pytest.skip("oracle/_ref/libfm_ref.so not built (needs /root/reference at build time)")
    return ob.RefIndex(tm1m["tm"], tm1m["off"])


def sample(d, n):
    return d["q"][:d["qo"][n]], d["qo"][:n + 1]


def test_config2_full_size_properties(tm1m):
    """configs[1]: size-independent properties at full size -- every unperturbed TM sentence finds itself
    with score 1.0, results do not depend on how the batch is split -- and a 2000-query sample against the oracle."""
    tm, off, index = tm1m["tm"], tm1m["off"], tm1m["index"]
    ids = np.arange(0, 1000000, 997)[:1000]
    qo = np.zeros(len(ids) + 1, dtype=np.int64)
    np.cumsum(off[ids + 1] - off[ids], out=qo[1:])
    q = np.concatenate([tm[off[i]:off[i + 1]] for i in ids])
    out, cnt = index.match_batch(q, qo, cap=1, fuzzy=0.7, n=1, ml=3)
    assert (cnt == 1).all() and (out["score"][:, 0] == 1.0).all()
    for k in range(0, 1000, 50):  # the match is an identical sentence with the smallest s_id
        sid = int(out["s_id"][k, 0])
        assert sid <= ids[k] and np.array_equal(index.sentence(sid), tm[off[ids[k]]:off[ids[k] + 1]])
    q2, qo2 = tm1m["q"], tm1m["qo"]
    a, ca = index.match_batch(q2, qo2, cap=1, fuzzy=0.7, n=1, ml=3)
    half = 10000
    b1, c1 = index.match_batch(q2[:qo2[half]], qo2[:half + 1], cap=1, fuzzy=0.7, n=1, ml=3)
    b2, c2 = index.match_batch(q2[qo2[half]:], qo2[half:] - qo2[half], cap=1, fuzzy=0.7, n=1, ml=3)
    assert (np.concatenate([c1, c2]) == ca).all()
    assert np.concatenate([b1, b2]).tobytes() == a.tobytes()
    compare(index, tm1m["oracle"], *sample(tm1m, 2000), cap=1, fuzzy=0.7, n=1, ml=3)


def test_config2_vs_reference_itself(tm1m, ref1m):
    """configs[1] and the CLI defaults against the unmodified reference on the box (closes the chain
    GPU == restatement == reference inside the -m gpu run)."""
    q, qo = sample(tm1m, 1500)
    cnt = compare(tm1m["index"], ref1m, q, qo, cap=1, with_cost=False, fuzzy=0.7, n=1, ml=3)
    assert (cnt > 0).mean() > 0.5
    compare(tm1m["index"], ref1m, q, qo, cap=5, with_cost=False, fuzzy=0.8, n=5, ml=3, mr=0.3)


def test_config3_shape_1m(tm1m):
    """configs[2] parameters (f=0.5, ml=3, n=1) on one GPU's worth of TM."""
    cnt = compare(tm1m["index"], tm1m["oracle"], *sample(tm1m, 1500), cap=1, fuzzy=0.5, n=1, ml=3)
    assert (cnt > 0).mean() > 0.6


def test_config5_contrastive_idf_1m(tm1m, ref1m):
    """configs[4]: n=10, contrast 0.5 (mean), idf-penalty 1.0, f=0.7 -- against the restatement and the reference."""
    q, qo = sample(tm1m, 1200)
    params = dict(fuzzy=0.7, n=10, ml=3, idf=1.0, contrast=0.5)
    compare(tm1m["index"], tm1m["oracle"], q, qo, cap=10, **params)
    compare(tm1m["index"], ref1m, q, qo, cap=10, with_cost=False, **params)
    compare(tm1m["index"], tm1m["oracle"], q, qo, cap=10, fuzzy=0.5, n=10, ml=3, idf=1.0, contrast=0.5, reduce=1, buffer=30)


def test_config4_long_patterns_1m():
    """configs[3]: 980k short + 20k sentences of 200-300 tokens, queries = perturbed long sentences
    (plus short ones in the same batch), max_tokens_in_pattern=300, f=0.7."""
    tm, off, V = synth.make_tm(1000000, seed=1234, n_long=20000)
    src = np.arange(980000, 1000000)
    ql, qlo = synth.make_queries(tm, off, 320, seed=72, source_ids=src, frac_random=0.1, len_lo=200, len_hi=300)
    qs, qso = synth.make_queries(tm, off, 700, seed=73)
    q, qo = np.concatenate([ql, qs]), np.concatenate([qlo, qso[1:] + qlo[-1]])
    index, oracle = fmb.Index(tm, off, V), ob.OracleIndex(tm, off, V)
    cnt = compare(index, oracle, q, qo, cap=1, fuzzy=0.7, n=1, ml=3)
    assert (cnt[:320] > 0).mean() > 0.7
    compare(index, oracle, q, qo, cap=4, fuzzy=0.5, n=4, ml=3, idf=1.0)
    compare(index, oracle, ql, qlo, cap=3, fuzzy=0.6, n=3, ml=3, costs=(1, 0, 1))
    compare(index, oracle, ql, qlo, cap=2, fuzzy=0.7, n=2, ml=3, mr=0.3, costs=(2, 2, 2))


@pytest.fixture(scope="module")
def realtext():
    tm, off, V, q, qo, expected = load_realtext()
    return dict(tm=tm, off=off, V=V, q=q, qo=qo, expected=expected, index=fmb.Index(tm, off, V))


@pytest.mark.parametrize("k", range(len(REALTEXT_PARAM_SETS)))
def test_realtext_vs_reference_golden(realtext, k):
    """Europarl sentences (real phrase structure, stop-word trigrams shared by thousands of sentences):
    ids, scores, penalties, match lengths identical to what the unmodified reference returned."""
    params = REALTEXT_PARAM_SETS[k]
    t = time.perf_counter()
    out, cnt = realtext["index"].match_batch(realtext["q"], realtext["qo"], cap=256, **params)
    dt = time.perf_counter() - t
    got = [as_tuples(out[i, :cnt[i]]) for i in range(len(cnt))]
    assert got == realtext["expected"][k]
    print("realtext %s: %d queries in %.2f ms" % (params, len(cnt), dt * 1e3))


def test_realtext_all_parameter_sets_vs_oracle(realtext):
    from tests.test_gpu_parity import PARAM_SETS
    oracle = ob.OracleIndex(realtext["tm"], realtext["off"], realtext["V"])
    for params in PARAM_SETS:
        compare(realtext["index"], oracle, realtext["q"], realtext["qo"], cap=512, **params)


def run_case(case, env):
    e = dict(os.environ)
    e.update(env)
    out = subprocess.run([sys.executable, "-m", "tests.gpu_env_case", case], cwd=ROOT, env=e, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "identical" in out.stdout, (env, out.stdout[-2000:], out.stderr[-2000:])


@pytest.mark.parametrize("env", [
    {"FM_HEAVY_SMEM": "64", "FM_WARP_MAX": "32"},   # CTA radix sort for > 64 candidates, CTA replay above 32
    {"FM_WARP_MAX": "32"},
    {"FM_HEAVY_SMEM": "64"},
], ids=["radix+cta", "cta", "radix"])
def test_env_gated_replay_tiers(env):
    run_case("many_candidates", env)


@pytest.mark.parametrize("env", [
    {"FM_SCORE_WARP_ONLY": "1"},   # every pair through the float wavefront kernel (also p <= 32)
    {"FM_SCORE_FLOAT_ONLY": "1"},  # no bit-parallel DP: float kernels for unit costs too
    {},
], ids=["wavefront-only", "float-only", "default"])
def test_env_gated_scoring_paths(env):
    run_case("all_scoring_paths", env)


def test_env_gated_prepare_warp_kernel():
    """The warp-per-query prepare kernel for every query (the default on an index without long sentences is the
    thread-per-query kernel, with the warp kernel for the patterns it leaves over)."""
    run_case("many_candidates", {"FM_PREPARE_WARP_ONLY": "1"})
    run_case("short_patterns", {"FM_PREPARE_WARP_ONLY": "1"})
    run_case("short_patterns", {})


def test_more_than_24576_candidates_per_query():
    """Vocabulary of 8 words: ~55k scored candidates per query -- the CTA radix sort in global memory
    (lists beyond the shared-memory sort) without any environment switch."""
    tm, off, V = synth.make_tm(60000, vocab=8, len_lo=4, len_hi=24, seed=601)
    q, qo = synth.make_queries(tm, off, 12, vocab=8, seed=602, len_lo=6, len_hi=20)
    index, oracle = fmb.Index(tm, off, V), ob.OracleIndex(tm, off, V)
    index.set_profiling(True)
    compare(index, oracle, q, qo, cap=8, fuzzy=0.2, n=5, ml=1)
    assert index.profile()["n_survivors"] > 12 * 24576
    compare(index, oracle, q, qo, cap=64, fuzzy=0.3, n=0, ml=2, costs=(1, 0, 1))


def test_more_than_a_million_candidates_per_query():
    """Three-word vocabulary, 1.15M sentences: > 2^20 scored candidates for one query (the pair sort
    that carries the record index beside the key)."""
    tm, off, V = synth.make_tm(1150000, vocab=3, len_lo=6, len_hi=10, seed=611)
    q, qo = synth.make_queries(tm, off, 2, vocab=3, seed=612, len_lo=8, len_hi=8, frac_random=1.0)
    index, oracle = fmb.Index(tm, off, V), ob.OracleIndex(tm, off, V)
    index.set_profiling(True)
    compare(index, oracle, q, qo, cap=8, fuzzy=0.0, n=3, ml=2)
    assert index.profile()["n_survivors"] > 2 * (1 << 20)


def test_sharded_tm_over_nccl_two_gpus():
    """Sentence-id shards on two GPUs, one process per GPU over NCCL, against the unsharded oracle."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29731", os.path.join(ROOT, "tools", "check_sharded_nccl.py")],
                         cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "identical to oracle: False" not in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("identical to oracle: True") >= 6  # three of them contrastive
