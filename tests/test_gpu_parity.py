"""GPU parity tests (run with -m gpu on the B200 box). Everything goes through the C ABI
(fm_index_create / fm_match_batch / ...) and is compared bit for bit with the CPU oracle and with
the committed golden vectors produced by the unmodified reference."""
import numpy as np
import pytest

import fuzzy_match_b200 as fmb
from fuzzy_match_b200 import synth
from oracle import binding as ob
from tests.util import as_tuples, csr, first_seen_ids, fix_params, load_golden, load_subseq_golden

pytestmark = pytest.mark.gpu
CASES = load_golden()


def gpu_results(index, q, qo, cap, **params):
    out, cnt = index.match_batch(q, qo, cap=cap, **params)
    return [as_tuples(out[i, :min(cnt[i], cap)], True) for i in range(len(cnt))], cnt


def assert_same(index, oracle, q, qo, cap=32, **params):
    got, gcnt = gpu_results(index, q, qo, cap, **params)
    ro, ocnt = oracle.match_batch(q, qo, cap=cap, **params)
    want = [as_tuples(r, True) for r in ro]
    assert (gcnt == ocnt).all(), "match counts differ at queries %s" % np.nonzero(gcnt != ocnt)[0][:10]
    bad = [i for i in range(len(want)) if got[i] != want[i]]
    assert not bad, "query %d: gpu %s != oracle %s" % (bad[0], got[bad[0]], want[bad[0]])


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_golden_vectors(case):
    tok, off = csr(case["tm"])
    q, qo = csr(case["queries"])
    index = fmb.Index(tok, off, case["vocab_size"], max_tokens=case["max_tokens"])
    if "tm_real" in case:  # Sentence API: real tokens, case class, penalty tokens
        from fuzzy_match_b200 import capi
        index.set_real(case["tm_real"], case["tm_gaps"], off)
        out, cnt = index.match_batch_real(q, case["q_real"], case["q_gaps"], qo, capi.itok_distance_table(synth.ITOKS), cap=64,
                                          **fix_params(case["params"]))
    else:
        out, cnt = index.match_batch(q, qo, cap=64, **fix_params(case["params"]))
    got = [as_tuples(out[i, :cnt[i]]) for i in range(len(cnt))]
    want = [[tuple(m) for m in r] for r in case["expected"]]
    assert got == want


PARAM_SETS = [
    dict(fuzzy=0.8, n=1, ml=3, mr=0.3),
    dict(fuzzy=0.7, n=1, ml=3),
    dict(fuzzy=0.5, n=5, ml=2),
    dict(fuzzy=0.3, n=0, ml=2),
    dict(fuzzy=0.5, n=10, ml=3, idf=1.0, contrast=0.5),
    dict(fuzzy=0.4, n=4, ml=2, idf=0.7, costs=(1, 0, 1), contrast=0.5, reduce=1, buffer=8),
    dict(fuzzy=0.4, n=4, ml=2, costs=(0.5, 1.5, 1.2)),
    dict(fuzzy=0.4, n=3, ml=2, costs=(0.4, 0.3, 1.2), idf=2.0),
    dict(fuzzy=0.6, n=2, ml=0, costs=(1, 0, 1), no_perfect=True),
    dict(fuzzy=0.0, n=3, ml=4, mr=0.2, buffer=1),
]


@pytest.fixture(scope="module")
def medium():
    tm, off, V = synth.make_tm(20000, vocab=5000, seed=51)
    q, qo = synth.make_queries(tm, off, 1500, vocab=5000, seed=52)
    return fmb.Index(tm, off, V), ob.OracleIndex(tm, off, V), q, qo


@pytest.mark.parametrize("params", PARAM_SETS, ids=[str(i) for i in range(len(PARAM_SETS))])
def test_random_tm_vs_oracle(medium, params):
    index, oracle, q, qo = medium
    assert_same(index, oracle, q, qo, cap=48, **params)


def test_index_metadata(medium):
    index, oracle, _, _ = medium
    assert index.num_sentences == oracle.num_sentences
    assert (index.sfreq() == oracle.sfreq).all()
    assert (index.kept_sources() == np.arange(index.num_sentences)).all()


def test_small_vocab_many_duplicates():
    """Tiny vocabulary: every query has thousands of candidates, heavy dedup and ties."""
    tm, off, V = synth.make_tm(3000, vocab=40, len_lo=1, len_hi=30, seed=61)
    q, qo = synth.make_queries(tm, off, 200, vocab=40, seed=62, len_lo=1, len_hi=30)
    index, oracle = fmb.Index(tm, off, V), ob.OracleIndex(tm, off, V)
    for params in (dict(fuzzy=0.5, n=3, ml=2), dict(fuzzy=0.2, n=0, ml=1), dict(fuzzy=0.6, n=2, ml=3, costs=(1, 0, 1))):
        assert_same(index, oracle, q, qo, cap=3000, **params)


def test_long_patterns():
    """200-300 token sentences: several DP columns per lane, multi-word coverage masks."""
    tm, off, V = synth.make_tm(1500, vocab=3000, seed=71, n_long=300)
    src = np.arange(1200, 1500)
    q, qo = synth.make_queries(tm, off, 120, vocab=3000, seed=72, source_ids=src, frac_random=0.1, len_lo=200, len_hi=300)
    index, oracle = fmb.Index(tm, off, V), ob.OracleIndex(tm, off, V)
    for params in (dict(fuzzy=0.7, n=1, ml=3), dict(fuzzy=0.4, n=5, ml=3, idf=1.0), dict(fuzzy=0.5, n=3, ml=2, costs=(1, 0, 1))):
        assert_same(index, oracle, q, qo, cap=16, **params)


def test_max_tokens_cap_and_dropped_sentences():
    tm, off, V = synth.make_tm(800, vocab=300, len_lo=1, len_hi=40, seed=81)
    q, qo = synth.make_queries(tm, off, 150, vocab=300, seed=82, len_lo=1, len_hi=40)
    index, oracle = fmb.Index(tm, off, V, max_tokens=25), ob.OracleIndex(tm, off, V, max_tokens=25)
    assert index.num_sentences == oracle.num_sentences < 800
    assert_same(index, oracle, q, qo, cap=8, fuzzy=0.5, n=4, ml=2)


def test_workspace_regrowth_and_reuse():
    """A batch whose worklists overflow the initial workspace must regrow and give the same answer,
    and a later small batch on the same index must still be right."""
    tm, off, V = synth.make_tm(60000, vocab=30, len_lo=8, len_hi=12, seed=91)
    q, qo = synth.make_queries(tm, off, 64, vocab=30, seed=92, len_lo=8, len_hi=12)
    index, oracle = fmb.Index(tm, off, V), ob.OracleIndex(tm, off, V)
    index.set_profiling(True)
    assert_same(index, oracle, q, qo, cap=8, fuzzy=0.45, n=4, ml=2)
    assert index.profile()["n_elements"] > 0
    assert_same(index, oracle, q[:qo[3]], qo[:4], cap=8, fuzzy=0.45, n=4, ml=2)


def test_python_mirror_of_reference_api():
    """The reference's own Tokens-API test (test/test.cc:337-375, lcs_cost) against the mirror class."""
    fm = fmb.FuzzyMatch(max_tokens_in_pattern=300)
    for s in ("a b c", "a b c d e x x x", "x x a b c d e f x x x x x"):
        fm.add_tm("", s.split(), sort=False)
    fm.sort()
    matches = []
    assert fm.match("a b c d e f".split(), 0, 10, matches, 3, 0.5, 0, fmb.EditCosts(1, 0, 1))
    assert [m.s_id for m in matches] == [2, 1, 0]
    assert abs(matches[0].score - 1.0) < 1e-3 and abs(matches[1].score - 5 / 6) < 1e-3 and abs(matches[2].score - 0.5) < 1e-3
    # match() appends (src/fuzzy_match.cc:670-679) and returns False on an empty / over-long pattern
    assert fm.match([], 0.5, 1, []) is False
    assert fm.match(["a"] * 301, 0.5, 1, []) is False


def test_errors_are_loud():
    with pytest.raises(fmb.FuzzyMatchError):
        fmb.Index(np.array([2, 1, 3], dtype=np.int32), np.array([0, 3], dtype=np.int64), 10)
    with pytest.raises(fmb.FuzzyMatchError):
        fmb.Index(np.array([2, 3], dtype=np.int32), np.array([0, 2], dtype=np.int64), 10, max_tokens=5000)


def test_sharded_tm_shards_one_gpu():
    """The sharded path without the collective: fm_shard_accept_device per shard (the shard's own candidate loop,
    accepted records in one block per shard) + fm_merge_accepted_device over the blocks must equal the unsharded
    oracle bit for bit -- two and three sentence-id shards on one GPU, global IDF; a block that is too small is
    reported and the rerun with the reported capacity is complete."""
    import torch
    from fuzzy_match_b200 import capi, sharded
    tm, off, V = synth.make_tm(6000, vocab=700, len_lo=0, len_hi=30, seed=101)
    q, qo = synth.make_queries(tm, off, 400, vocab=700, seed=102, len_lo=1, len_hi=30)
    oracle = ob.OracleIndex(tm, off, V, max_tokens=28)
    n_sent = len(off) - 1
    dev = torch.device("cuda", 0)
    d_tok = torch.as_tensor(q, device=dev)
    d_off = torch.as_tensor(qo.astype(np.int32), device=dev)
    n_q, n_tok, cap = len(qo) - 1, int(qo[-1]), 8
    st = torch.cuda.current_stream(dev).cuda_stream
    for n_shards in (2, 3):
        shards, base = [], 0
        for r in range(n_shards):
            lo, hi = sharded.shard_range(n_sent, r, n_shards)
            ix = fmb.Index(tm[off[lo]:off[hi]], off[lo:hi + 1] - off[lo], V, max_tokens=28, s_id_base=base)
            base += ix.num_sentences
            shards.append(ix)
        sf = sum(ix.sfreq().astype(np.int64) for ix in shards).astype(np.uint32)
        assert (sf == oracle.sfreq).all() and base == oracle.num_sentences
        for ix in shards:
            ix.set_idf_stats(sf, base)
        grew = 0
        for params in (dict(fuzzy=0.5, n=4, ml=2), dict(fuzzy=0.4, n=3, ml=3, idf=1.0, costs=(1, 0, 1)), dict(fuzzy=0.6, n=1, ml=3, mr=0.3),
                       dict(fuzzy=0.3, n=0, ml=2), dict(fuzzy=0.5, n=2, ml=2, no_perfect=True)):
            p = capi.Params.make(**params)
            capacity = 64
            for attempt in range(8):
                blocks = [torch.zeros(capi.wire_block_bytes(n_q, capacity), dtype=torch.uint8, device=dev) for _ in shards]
                for ix, blk in zip(shards, blocks):
                    ix.shard_accept_device(d_tok.data_ptr(), d_off.data_ptr(), n_q, n_tok, capacity, blk.data_ptr(), stream=st, params=p)
                d_out = torch.zeros(n_q * cap * 24, dtype=torch.uint8, device=dev)
                d_cnt = torch.zeros(n_q, dtype=torch.int32, device=dev)
                need = shards[0].merge_accepted_device([b.data_ptr() for b in blocks], capacity * n_shards, d_off.data_ptr(), n_q,
                                                       d_out.data_ptr(), d_cnt.data_ptr(), cap, stream=st, params=p)
                torch.cuda.synchronize()
                if need == 0:
                    break
                assert need > capacity
                capacity, grew = need, grew + 1
            out = d_out.cpu().numpy().view(capi.MATCH_DTYPE).reshape(n_q, cap)
            cnt = d_cnt.cpu().numpy()
            ro, oc = oracle.match_batch(q, qo, cap=cap, **params)
            assert (cnt == oc).all()
            assert [as_tuples(out[i, :min(cnt[i], cap)], True) for i in range(n_q)] == [as_tuples(r, True) for r in ro]
            # the blocks carry only what the shard's own loop accepted: far fewer records than scored candidates
            hdr = blocks[0][:16].cpu().numpy().view(np.int32)
            assert hdr[0] == 0 and hdr[1] == n_q and hdr[2] == capacity and 0 < hdr[3] <= capacity
        assert grew >= 1  # (64 records are not enough for 400 queries: the first attempt reports the size it needs)
        for ix in shards:
            ix.close()


def test_device_resident_api_matches_host_api(medium):
    import torch
    from fuzzy_match_b200 import capi
    index, _, q, qo = medium
    dev = torch.device("cuda", 0)
    d_tok = torch.as_tensor(q, device=dev)
    d_off = torch.as_tensor(qo.astype(np.int32), device=dev)
    n_q, cap = len(qo) - 1, 4
    d_out = torch.zeros(n_q * cap * 24, dtype=torch.uint8, device=dev)
    d_cnt = torch.zeros(n_q, dtype=torch.int32, device=dev)
    stream = torch.cuda.Stream(dev)
    with torch.cuda.stream(stream):
        index.match_batch_device(d_tok.data_ptr(), d_off.data_ptr(), n_q, int(qo[-1]), d_out.data_ptr(), d_cnt.data_ptr(), cap,
                                 stream=stream.cuda_stream, fuzzy=0.5, n=4, ml=2)
    torch.cuda.synchronize()
    out, cnt = index.match_batch(q, qo, cap=cap, fuzzy=0.5, n=4, ml=2)
    assert (d_cnt.cpu().numpy() == cnt).all()
    got = d_out.cpu().numpy().view(capi.MATCH_DTYPE).reshape(n_q, cap)
    for i in range(n_q):
        assert got[i, :cnt[i]].tobytes() == out[i, :cnt[i]].tobytes()


def test_cpp_adapter_runs_reference_gtests():
    """tests/cpp/test_adapter.cc: the reference's Tokens-API gtest cases against the C++ adapter
    (fuzzy_match_b200/cpp/fuzzy_match_b200.hh) on top of the C ABI."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "test_adapter")
    if not os.path.exists(exe):
        import __graft_entry__ as g
        g.build()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "all adapter tests passed" in out.stdout, out.stdout + out.stderr


def test_index_save_load_roundtrip(tmp_path, medium):
    """fm_index_save / fm_index_load: the reloaded index answers bit-identically and keeps its tables."""
    index, _, q, qo = medium
    path = tmp_path / "tm.fmb"
    index.save(path)
    again = fmb.Index.load(path, index.vocab_size)
    assert again.num_sentences == index.num_sentences and again.num_suffixes == index.num_suffixes
    assert (again.sfreq() == index.sfreq()).all() and (again.kept_sources() == index.kept_sources()).all()
    assert np.array_equal(again.sentence(17), index.sentence(17))
    for params in (dict(fuzzy=0.5, n=4, ml=2), dict(fuzzy=0.4, n=3, ml=3, idf=1.0)):
        a, ca = index.match_batch(q, qo, cap=4, **params)
        b, cb = again.match_batch(q, qo, cap=4, **params)
        assert (ca == cb).all() and a.tobytes() == b.tobytes()
    with pytest.raises(fmb.FuzzyMatchError):
        fmb.Index.load(tmp_path / "missing.fmb", 10)


def test_concurrent_host_threads(medium):
    """fm_match_batch is re-entrant on one shared index (like the reference's const match(), called by the
    CLI's N worker threads, cli/src/FuzzyMatch-cli.cc:139-147): 4 threads, different parameters."""
    import threading
    index, oracle, q, qo = medium
    params = [dict(fuzzy=0.5, n=4, ml=2), dict(fuzzy=0.7, n=1, ml=3), dict(fuzzy=0.4, n=3, ml=3, idf=1.0), dict(fuzzy=0.6, n=2, ml=2, costs=(1, 0, 1))]
    results = [None] * 4

    def work(k):
        for _ in range(3):
            results[k] = index.match_batch(q, qo, cap=4, **params[k])

    threads = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for k in range(4):
        out, cnt = results[k]
        ro, oc = oracle.match_batch(q, qo, cap=4, **params[k])
        assert (cnt == oc).all()
        assert all(out[i, :cnt[i]].tobytes() == ro[i].tobytes() for i in range(len(oc)))


def test_gpu_suffix_sort_matches_host_sort(monkeypatch):
    """The GPU-built index (prefix-doubling suffix sort, fm_sort.cu) and the host-thread build
    (FM_HOST_SORT=1, comparator sort like SuffixArray::sort) must answer identically, including on a TM
    full of repeated sentences (true ties in the suffix order)."""
    tm, off, V = synth.make_tm(4000, vocab=25, len_lo=1, len_hi=40, seed=301)
    tm = np.concatenate([tm, tm[:off[500]]])              # 500 duplicated sentences
    off = np.concatenate([off, off[-1] + off[1:501]])
    q, qo = synth.make_queries(tm, off, 300, vocab=25, seed=302, len_lo=1, len_hi=40)
    gpu_built = fmb.Index(tm, off, V)
    monkeypatch.setenv("FM_HOST_SORT", "1")
    host_built = fmb.Index(tm, off, V)
    monkeypatch.delenv("FM_HOST_SORT")
    oracle = ob.OracleIndex(tm, off, V)
    for params in (dict(fuzzy=0.5, n=5, ml=2), dict(fuzzy=0.3, n=0, ml=3, idf=1.0), dict(fuzzy=0.6, n=3, ml=1, costs=(1, 0, 1))):
        a, ca = gpu_built.match_batch(q, qo, cap=64, **params)
        b, cb = host_built.match_batch(q, qo, cap=64, **params)
        assert (ca == cb).all() and a.tobytes() == b.tobytes()
        ro, oc = oracle.match_batch(q, qo, cap=64, **params)
        assert (ca == oc).all() and all(a[i, :min(ca[i], 64)].tobytes() == ro[i].tobytes() for i in range(len(oc)))
    # subsequence() walks ranges in suffix order and stops early: the order among identical suffixes (by sentence id) shows
    for kw in (dict(n=1), dict(n=3, ml=1, mr=0.0, no_perfect=True)):
        a, b, o = gpu_built.subsequence_batch(q, qo, **kw), host_built.subsequence_batch(q, qo, **kw), oracle.subsequence_batch(q, qo, **kw)
        assert a.tobytes() == b.tobytes() == o.tobytes()


def test_sentence_api_vs_oracle(tmp_path):
    """Real tokens / case class / penalty tokens (fm_index_set_real + fm_match_batch_real) against the oracle,
    incl. long patterns, and after a save / load round trip of the index."""
    from fuzzy_match_b200 import capi
    blob, ioff = synth.itok_table()
    dist = capi.itok_distance_table(synth.ITOKS)
    tm, off, V = synth.make_tm(4000, vocab=300, len_lo=0, len_hi=60, seed=401)
    q, qo = synth.make_queries(tm, off, 400, vocab=300, seed=402, len_lo=1, len_hi=60)
    real, gaps = synth.make_real(tm, off, 403)
    qreal, qgaps = synth.make_real(q, qo, 404)
    index, oracle = fmb.Index(tm, off, V, max_tokens=50), ob.OracleIndex(tm, off, V, max_tokens=50)
    index.set_real(real, gaps, off)
    oracle.set_real(real, gaps, off, blob, ioff)
    path = tmp_path / "real.fmb"
    index.save(path)
    for ix in (index, fmb.Index.load(path, V)):
        for params in (dict(fuzzy=0.5, n=5, ml=2), dict(fuzzy=0.3, n=8, ml=2, idf=1.0, costs=(1, 0, 1)),
                       dict(fuzzy=0.4, n=4, ml=3, contrast=0.5, costs=(0.5, 1.5, 1.2)), dict(fuzzy=0.6, n=2, ml=2, no_perfect=True)):
            out, cnt = ix.match_batch_real(q, qreal, qgaps, qo, dist, cap=16, **params)
            ro, oc = oracle.match_batch_real(q, qreal, qgaps, qo, cap=16, **params)
            assert (cnt == oc).all()
            assert [as_tuples(out[i, :cnt[i]], True) for i in range(len(oc))] == [as_tuples(r, True) for r in ro]
    # the plain call on the same index still ignores the real side
    out, cnt = index.match_batch(q, qo, cap=4, fuzzy=0.5, n=4, ml=2)
    ro, oc = ob.OracleIndex(tm, off, V, max_tokens=50).match_batch(q, qo, cap=4, fuzzy=0.5, n=4, ml=2)
    assert (cnt == oc).all() and all(out[i, :cnt[i]].tobytes() == ro[i].tobytes() for i in range(len(oc)))


def test_cli_config1_output_equals_oracle(tmp_path):
    """BASELINE.json configs[0]: 1K-sentence TM, 100 queries, f=0.8, n=1, CLI defaults ml=3 mr=0.3 -- the
    streaming driver's stdout must equal the oracle's "score\\tid" lines (9 significant digits, ids = 1-based
    corpus line numbers, empty line when nothing matches; cli/src/FuzzyMatch-cli.cc:219-233)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "fuzzy_match_b200", "fm_cli")
    if not os.path.exists(exe):
        import __graft_entry__ as g
        g.build()
    tm, off, V = synth.make_tm(1000, vocab=400, seed=1234)
    q, qo = synth.make_queries(tm, off, 100, vocab=400, seed=5678)
    text = lambda tok, o, i: " ".join("w%d" % t for t in tok[o[i]:o[i + 1]])
    corpus = tmp_path / "tm.txt"
    corpus.write_text("".join(text(tm, off, i) + "\n" for i in range(1000)))
    queries = "".join(text(q, qo, i) + "\n" for i in range(100))
    for extra, params in (([], dict(fuzzy=0.8, n=1, ml=3, mr=0.3)),
                          (["-n", "3", "-f", "0.5", "--ml", "2", "--mr", "0", "-I", "1", "--batch", "7"], dict(fuzzy=0.5, n=3, ml=2, idf=1.0))):
        args = [exe, "-c", str(corpus), "-f", "0.8", "-n", "1"] + extra
        out = subprocess.run(args, input=queries, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr
        ro, oc = ob.OracleIndex(tm, off, V).match_batch(q, qo, cap=3, **params)
        want = ["\t".join("%.9g\t%d" % (float(m["score"]), int(m["s_id"]) + 1) for m in r) for r in ro]
        assert out.stdout.split("\n")[:-1] == want
        assert "NMATCH\t%d\t/\t100" % int((oc > 0).sum()) in out.stderr


# cost sets whose normaliser (include/fuzzy/costs.hh:33-47) stays positive: with e.g. (1, 0, 0) the reference
# divides by zero and crashes, so there is nothing to be compatible with
COST_SETS = [(1, 1, 1), (1, 0, 1), (0.5, 1.5, 1.2), (0.4, 0.3, 1.2), (1.7, 1, 0.4), (1, 1, 2.5), (0.4, 0.4, 1.0)]


def test_randomised_parameters_vs_oracle():
    """Seeded sweep over TM shapes and every match() parameter (incl. ml 0/1, mr above 1, fuzzy above 1 and
    below 0, N=0, buffer 0/1/large, zero and fractional costs, idf, contrastive, no_perfect)."""
    rng = np.random.default_rng(20251017)
    for trial in range(40):
        vocab = int(rng.choice([6, 20, 80, 600]))
        hi = int(rng.choice([3, 8, 20, 45]))
        n_sent = int(rng.choice([1, 40, 700, 2500]))
        tm, off, V = synth.make_tm(n_sent, vocab=vocab, len_lo=0, len_hi=hi, seed=1000 + trial)
        q, qo = synth.make_queries(tm, off, 60, vocab=vocab, seed=2000 + trial, len_lo=0, len_hi=hi)
        q = q.copy()
        if len(q):
            q[rng.integers(0, len(q), size=max(1, len(q) // 15))] = rng.choice([-3, 0, 1, V, V + 7, 2**31 - 1])  # junk ids
        max_tokens = int(rng.choice([300, hi, max(1, hi // 2)]))
        params = dict(fuzzy=float(rng.choice([-0.2, 0.0, 0.3, 0.5, 0.7, 0.9, 1.0, 1.1])), n=int(rng.choice([0, 1, 2, 5])),
                      ml=int(rng.choice([-1, 0, 1, 2, 3, 5])), mr=float(rng.choice([0.0, 0.3, 0.9, 1.5])),
                      idf=float(rng.choice([0.0, 0.0, 0.5, 2.0])),
                      costs=COST_SETS[int(rng.integers(0, len(COST_SETS)))],
                      contrast=float(rng.choice([0.0, 0.0, 0.5, 1.0])), reduce=int(rng.integers(0, 2)),
                      buffer=int(rng.choice([-1, 0, 1, 3, 50])), no_perfect=bool(rng.integers(0, 2)))
        if params["idf"] and n_sent == 1:
            params["idf"] = 0.0  # log(1) == 0 divides the idf weight by zero in the reference (NaN costs)
        try:
            index = fmb.Index(tm, off, V, max_tokens=max_tokens)
        except fmb.FuzzyMatchError:
            raise
        oracle = ob.OracleIndex(tm, off, V, max_tokens=max_tokens)
        cap = 4096
        out, cnt = index.match_batch(q, qo, cap=cap, **params)
        ro, oc = oracle.match_batch(q, qo, cap=cap, **params)
        assert (cnt == oc).all(), (trial, params)
        for i in range(len(oc)):
            assert out[i, :cnt[i]].tobytes() == ro[i].tobytes(), (trial, i, params)


def test_maximum_sentence_length():
    """max_tokens_in_pattern at the hard cap (1023): 32 DP columns per lane, 32-word coverage masks."""
    tm, off, V = synth.make_tm(60, vocab=900, seed=501, n_long=40, long_lo=700, long_hi=1023)
    q, qo = synth.make_queries(tm, off, 16, vocab=900, seed=502, source_ids=np.arange(20, 60), frac_random=0.0, len_lo=700, len_hi=1023)
    keep = np.diff(qo) <= 1023
    index, oracle = fmb.Index(tm, off, V, max_tokens=1023), ob.OracleIndex(tm, off, V, max_tokens=1023)
    for params in (dict(fuzzy=0.6, n=2, ml=3), dict(fuzzy=0.5, n=3, ml=3, idf=1.0, costs=(1, 0, 1))):
        out, cnt = index.match_batch(q, qo, cap=4, **params)
        ro, oc = oracle.match_batch(q, qo, cap=4, **params)
        assert (cnt == oc).all() and keep.any()
        assert all(out[i, :cnt[i]].tobytes() == ro[i].tobytes() for i in range(len(oc)))


def test_degenerate_inputs():
    empty = fmb.Index(np.zeros(0, dtype=np.int32), np.zeros(1, dtype=np.int64), 10)
    out, cnt = empty.match_batch(np.array([2, 3, 4], dtype=np.int32), np.array([0, 3], dtype=np.int64), cap=2, fuzzy=0.5, n=2)
    assert empty.num_sentences == 0 and cnt.tolist() == [0]
    dropped = fmb.Index(np.array([2, 3, 4, 5], dtype=np.int32), np.array([0, 0, 4], dtype=np.int64), 10, max_tokens=3)
    assert dropped.num_sentences == 0
    index = fmb.Index(np.array([2, 3, 4], dtype=np.int32), np.array([0, 3], dtype=np.int64), 10)
    out, cnt = index.match_batch(np.zeros(0, dtype=np.int32), np.zeros(1, dtype=np.int64), cap=1, fuzzy=0.5, n=1)
    assert len(cnt) == 0
    out, cnt = index.match_batch(np.array([2, 3, 4], dtype=np.int32), np.array([0, 0, 3, 3], dtype=np.int64), cap=1, fuzzy=0.5, n=1)
    assert cnt.tolist() == [0, 1, 0] and out[1, 0]["score"] == 1.0


def test_submit_wait_pipeline(medium):
    """fm_match_batch_submit / fm_ticket_wait: several batches in flight on one index (each ticket owns a
    workspace) give the same bytes as the synchronous call, in any wait order; same for device buffers."""
    import torch
    from fuzzy_match_b200 import capi
    index, _, q, qo = medium
    params = capi.Params.make(fuzzy=0.5, n=4, ml=2)
    cap, n_q = 4, len(qo) - 1
    parts = [(0, 500), (500, 1100), (1100, n_q)]
    want, wcnt = index.match_batch(q, qo, cap=cap, params=params)
    bufs, tickets = [], []
    for a, b in parts:
        pq = np.ascontiguousarray(q[qo[a]:qo[b]])
        po = np.ascontiguousarray(qo[a:b + 1] - qo[a])
        out = np.zeros((b - a, cap), dtype=capi.MATCH_DTYPE)
        cnt = np.zeros(b - a, dtype=np.int32)
        bufs.append((pq, po, out, cnt))
        tickets.append(index.submit(pq, po, out, cnt, cap, params))
    for k in (1, 0, 2):
        index.wait(tickets[k])
    for (a, b), (_, _, out, cnt) in zip(parts, bufs):
        assert (cnt == wcnt[a:b]).all() and out.tobytes() == want[a:b].tobytes()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(dev)
    dbufs, tickets = [], []
    for a, b in parts:
        d_tok = torch.as_tensor(q[qo[a]:qo[b]], device=dev)
        d_off = torch.as_tensor((qo[a:b + 1] - qo[a]).astype(np.int32), device=dev)
        d_out = torch.zeros((b - a) * cap * 24, dtype=torch.uint8, device=dev)
        d_cnt = torch.zeros(b - a, dtype=torch.int32, device=dev)
        dbufs.append((d_tok, d_off, d_out, d_cnt))
        tickets.append(index.submit_device(d_tok.data_ptr(), d_off.data_ptr(), b - a, int(qo[b] - qo[a]), d_out.data_ptr(),
                                           d_cnt.data_ptr(), cap, stream.cuda_stream, params))
    for t in tickets:
        index.wait(t)
    for (a, b), (_, _, d_out, d_cnt) in zip(parts, dbufs):
        cnt = d_cnt.cpu().numpy()
        got = d_out.cpu().numpy().view(capi.MATCH_DTYPE).reshape(b - a, cap)
        assert (cnt == wcnt[a:b]).all()
        assert all(got[i, :cnt[i]].tobytes() == want[a + i, :cnt[i]].tobytes() for i in range(b - a))


# ---- subsequence() (reference src/fuzzy_match.cc:238-365) through fm_subsequence_batch
SUBSEQ = load_subseq_golden()


def subseq_tuples(rec, with_cost=False):
    return [([int(r["found"]), int(r["s_id"]), int(r["score"].view(np.uint32)), int(r["length"]), int(r["position"])]
             + ([int(r["cost"].view(np.uint32))] if with_cost else [])) if r["found"] else [0, 0, 0, 0, 0] + ([0] if with_cost else [])
            for r in rec]


@pytest.mark.parametrize("case", SUBSEQ["cases"], ids=[c["name"] for c in SUBSEQ["cases"]])
def test_subsequence_golden_vectors(case):
    t = SUBSEQ["tms"][case["tm"]]
    tok, off = csr(t["tm"])
    q, qo = csr(t["queries"])
    index = fmb.Index(tok, off, t["vocab_size"])
    assert subseq_tuples(index.subsequence_batch(q, qo, **case["params"])) == case["expected"]


SUBSEQ_PARAMS = [dict(n=1), dict(n=5, no_perfect=True), dict(n=3, ml=2, mr=0.0, idf_weighting=True),
                 dict(n=50, ml=1, mr=0.5, no_perfect=True, idf_weighting=True), dict(n=0), dict(n=2, ml=40, mr=0.0),
                 dict(n=4000, ml=1, mr=0.0)]


@pytest.mark.parametrize("n_sent,vocab,n_long", [(3000, 8, 0), (20000, 300, 0), (50000, 20000, 0), (4000, 50, 40)])
def test_subsequence_vs_oracle(n_sent, vocab, n_long):
    tm, off, _ = synth.make_tm(n_sent, vocab=vocab, seed=n_sent + vocab, n_long=n_long)
    src = np.arange(n_sent - n_long, n_sent) if n_long else None
    q, qo = synth.make_queries(tm, off, 400, vocab=vocab, seed=n_sent + vocab + 1, source_ids=src,
                               **(dict(len_lo=200, len_hi=300) if n_long else {}))
    tm, q, V = first_seen_ids(tm, q)
    index, oracle = fmb.Index(tm, off, V, max_tokens=400), ob.OracleIndex(tm, off, V, max_tokens=400)  # perturbed long patterns may exceed 300
    for kw in SUBSEQ_PARAMS:
        got = subseq_tuples(index.subsequence_batch(q, qo, **kw), True)
        want = subseq_tuples(oracle.subsequence_batch(q, qo, **kw), True)
        bad = [i for i in range(len(want)) if got[i] != want[i]]
        assert not bad, "%s query %d: gpu %s != oracle %s" % (kw, bad[0], got[bad[0]], want[bad[0]])


def test_subsequence_degenerate_inputs():
    tm, off, V = synth.make_tm(500, vocab=40, seed=3)
    index, oracle = fmb.Index(tm, off, V), ob.OracleIndex(tm, off, V)
    q, qo = csr([[], [5], [V + 3, V + 4, V + 5], [1, 1, 1, 1], tm[off[7]:off[8]].tolist(), list(range(2, 12)) * 30])
    for kw in (dict(n=1), dict(n=1, ml=0, mr=0.0), dict(n=3, ml=1, mr=0.0, no_perfect=True)):
        assert subseq_tuples(index.subsequence_batch(q, qo, **kw), True) == subseq_tuples(oracle.subsequence_batch(q, qo, **kw), True)
    assert len(index.subsequence_batch(*csr([]), n=1)) == 0
    with pytest.raises(fmb.FuzzyMatchError):  # longer than max_tokens_in_pattern: refused, never silently unmatched
        index.subsequence_batch(*csr([list(range(2, 12)) * 40]), n=1)


def test_match_into_nonempty_result_vectors():
    """fm_match_batch_prior: entries already in `matches` (src/fuzzy_match.cc:626-679) against the oracle, whose prior
    handling is pinned to the live reference in tests/test_oracle.py."""
    from tests.test_oracle import PRIOR_PARAM_PAIRS
    for seed, vocab in [(0, 30), (1, 200), (2, 12), (3, 5000)]:
        tm, off, V = synth.make_tm(20000 if vocab == 5000 else 3000, vocab=vocab, seed=seed)
        q1, q1o = synth.make_queries(tm, off, 300, vocab=vocab, seed=seed + 100)
        q2, q2o = synth.make_queries(tm, off, 300, vocab=vocab, seed=seed + 100, p_sub=0.2)
        index, oracle = fmb.Index(tm, off, V), ob.OracleIndex(tm, off, V)
        for p1, p2 in PRIOR_PARAM_PAIRS:
            first, fcnt = index.match_batch(q1, q1o, cap=64, **p1)
            poff = np.zeros(len(fcnt) + 1, dtype=np.int64)
            np.cumsum(np.minimum(fcnt, 64), out=poff[1:])
            psid = np.concatenate([first[i, :min(fcnt[i], 64)]["s_id"] for i in range(len(fcnt))] + [np.zeros(0, np.uint32)]).astype(np.uint32)
            got, gcnt = index.match_batch_prior(q2, q2o, psid, poff, cap=64, **p2)
            want, wcnt = oracle.match_batch_prior(q2, q2o, psid, poff, cap=64, **p2)
            assert (gcnt == wcnt).all()
            for i in range(len(wcnt)):
                assert as_tuples(got[i, :min(gcnt[i], 64)], True) == as_tuples(want[i], True), (p2, i)
    with pytest.raises(fmb.FuzzyMatchError):  # a sentence id the index does not hold
        index.match_batch_prior(q2[:q2o[1]], q2o[:2], np.array([10 ** 9], dtype=np.uint32), np.array([0, 1]), fuzzy=0.5, n=2, contrast=0.5)
