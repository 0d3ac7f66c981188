"""ctypes loaders for the two CPU checkers (TEST INFRASTRUCTURE ONLY).

  OracleIndex  -> oracle/libfm_oracle.so   plain-C restatement (oracle/fm_oracle.c)
  RefIndex     -> oracle/_ref/libfm_ref.so the reference's own sources behind oracle/ref_driver.cc

Both take the TM and the queries as int32 CSR (tokens, int64 offsets) and return, per query, a list
of (s_id, score, penalty, max_subseq, length) in the order the reference returns its matches.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libfm_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libfm_ref.so")


def build(verbose=False):
    """Compile the restatement and, when /root/reference is present, the reference itself."""
    out = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout, out.stderr)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed")


class Params(C.Structure):
    _fields_ = [("fuzzy", C.c_float), ("number_of_matches", C.c_int32), ("no_perfect", C.c_int32),
                ("min_subseq_length", C.c_int32), ("min_subseq_ratio", C.c_float),
                ("vocab_idf_penalty", C.c_float), ("insert_cost", C.c_float), ("delete_cost", C.c_float),
                ("replace_cost", C.c_float), ("contrastive_factor", C.c_float),
                ("contrast_reduce", C.c_int32), ("contrast_buffer", C.c_int32)]


class RefParams(C.Structure):
    _fields_ = [("fuzzy", C.c_float), ("number_of_matches", C.c_int32),
                ("min_subseq_length", C.c_int32), ("min_subseq_ratio", C.c_float),
                ("vocab_idf_penalty", C.c_float), ("insert_cost", C.c_float), ("delete_cost", C.c_float),
                ("replace_cost", C.c_float), ("contrastive_factor", C.c_float),
                ("contrast_reduce", C.c_int32), ("contrast_buffer", C.c_int32)]


MATCH_DTYPE = np.dtype([("s_id", np.uint32), ("score", np.float32), ("penalty", np.float32),
                        ("max_subseq", np.int32), ("length", np.int32), ("cost", np.float32)])
REF_MATCH_DTYPE = np.dtype([("s_id", np.uint32), ("score", np.float32), ("penalty", np.float32),
                            ("max_subseq", np.int32), ("length", np.int32)])
CAND_DTYPE = np.dtype([("s_id", np.uint32), ("longest_match", np.int32), ("s_length", np.int32),
                       ("cover", np.int32), ("rejected", np.int32), ("cost_full", np.float32),
                       ("rowmin_max", np.float32), ("accepted", np.int32)])
COUNTER_NAMES = ["queries", "equal_range_calls", "probes", "elements_walked", "candidates",
                 "candidate_tokens", "dp_pairs", "dp_tokens", "dp_cells", "pattern_tokens", "matches_out"]


def make_params(cls, fuzzy=0.7, n=1, ml=2, mr=0.0, idf=0.0, costs=(1.0, 1.0, 1.0), contrast=0.0,
                reduce=0, buffer=-1, no_perfect=False):
    p = cls()
    p.fuzzy, p.number_of_matches, p.min_subseq_length, p.min_subseq_ratio = fuzzy, n, ml, mr
    p.vocab_idf_penalty = idf
    p.insert_cost, p.delete_cost, p.replace_cost = costs
    p.contrastive_factor, p.contrast_reduce, p.contrast_buffer = contrast, reduce, buffer
    if cls is Params:
        p.no_perfect = int(no_perfect)
    elif no_perfect:
        raise ValueError("match(Tokens) of the reference has no no_perfect argument")
    return p


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _csr(tokens, off):
    return (np.ascontiguousarray(tokens, dtype=np.int32), np.ascontiguousarray(off, dtype=np.int64))


def _split(out, cnt, cap):
    res = []
    for q in range(len(cnt)):
        k = min(int(cnt[q]), cap)
        res.append(out[q * cap:q * cap + k].copy())
    return res


class OracleIndex:
    def __init__(self, tokens, off, vocab_size, max_tokens=300, sfreq_global=None, n_sent_global=0):
        self.lib = lib = C.CDLL(ORACLE_SO)
        lib.fmo_index_create.restype = C.c_void_p
        lib.fmo_index_num_sentences.restype = C.c_int64
        lib.fmo_index_num_suffixes.restype = C.c_int64
        lib.fmo_index_sfreq.restype = C.POINTER(C.c_uint32)
        lib.fmo_index_kept.restype = C.POINTER(C.c_int64)
        lib.fmo_match_debug.restype = C.c_int64
        tokens, off = _csr(tokens, off)
        sf = None if sfreq_global is None else np.ascontiguousarray(sfreq_global, dtype=np.uint32)
        self.vocab_size = int(vocab_size)
        h = lib.fmo_index_create(_ptr(tokens), _ptr(off), C.c_int64(len(off) - 1), C.c_int32(vocab_size),
                                 C.c_int32(max_tokens), None if sf is None else _ptr(sf), C.c_int64(n_sent_global))
        if not h:
            raise ValueError("fmo_index_create failed (TM token outside [2, vocab_size))")
        self.h = C.c_void_p(h)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.fmo_index_destroy(self.h)
            self.h = None

    @property
    def num_sentences(self):
        return self.lib.fmo_index_num_sentences(self.h)

    @property
    def sfreq(self):
        return np.ctypeslib.as_array(self.lib.fmo_index_sfreq(self.h), shape=(self.vocab_size,)).copy()

    def match_batch(self, q_tokens, q_off, cap=16, nthreads=1, counters=False, **kw):
        q_tokens, q_off = _csr(q_tokens, q_off)
        n_q = len(q_off) - 1
        p = make_params(Params, **kw)
        out = np.zeros(n_q * cap, dtype=MATCH_DTYPE)
        cnt = np.zeros(n_q, dtype=np.int32)
        ct = (C.c_int64 * len(COUNTER_NAMES))()
        self.lib.fmo_match_batch(self.h, _ptr(q_tokens), _ptr(q_off), C.c_int64(n_q), C.byref(p), C.c_int(nthreads),
                                 C.c_int64(cap), _ptr(out), _ptr(cnt), ct if counters else None)
        res = _split(out, cnt, cap)
        if counters:
            return res, cnt, dict(zip(COUNTER_NAMES, list(ct)))
        return res, cnt

    def match_batch_prior(self, q_tokens, q_off, prior_sid, prior_off, cap=16, **kw):
        """fmo_match_batch_prior: match() into result vectors that already hold the sentences prior_sid (CSR per query)."""
        q_tokens, q_off = _csr(q_tokens, q_off)
        prior_sid = np.ascontiguousarray(prior_sid, dtype=np.uint32)
        prior_off = np.ascontiguousarray(prior_off, dtype=np.int64)
        n_q = len(q_off) - 1
        p = make_params(Params, **kw)
        out = np.zeros(n_q * cap, dtype=MATCH_DTYPE)
        cnt = np.zeros(n_q, dtype=np.int32)
        self.lib.fmo_match_batch_prior(self.h, _ptr(q_tokens), _ptr(q_off), C.c_int64(n_q), C.byref(p), _ptr(prior_sid), _ptr(prior_off),
                                       C.c_int64(cap), _ptr(out), _ptr(cnt))
        return _split(out, cnt, cap), cnt

    def set_real(self, real, gaps, off, itok_blob, itok_off):
        """Attach real tokens / penalty tokens (Sentence API) to the indexed sentences."""
        real = np.ascontiguousarray(real, dtype=np.int32)
        gaps = np.ascontiguousarray(gaps, dtype=np.int32)
        off = np.ascontiguousarray(off, dtype=np.int64)
        self._itok = (np.ascontiguousarray(itok_blob, dtype=np.uint8), np.ascontiguousarray(itok_off, dtype=np.int32))
        rc = self.lib.fmo_index_set_real(self.h, _ptr(real), _ptr(gaps), _ptr(off), _ptr(self._itok[0]), _ptr(self._itok[1]),
                                         C.c_int32(len(itok_off) - 1))
        if rc:
            raise ValueError("fmo_index_set_real failed (itok id out of range)")

    def match_batch_real(self, q_tokens, q_real, q_gaps, q_off, cap=16, nthreads=1, **kw):
        q_tokens, q_off = _csr(q_tokens, q_off)
        q_real = np.ascontiguousarray(q_real, dtype=np.int32)
        q_gaps = np.ascontiguousarray(q_gaps, dtype=np.int32)
        n_q = len(q_off) - 1
        p = make_params(Params, **kw)
        out = np.zeros(n_q * cap, dtype=MATCH_DTYPE)
        cnt = np.zeros(n_q, dtype=np.int32)
        self.lib.fmo_match_batch_real(self.h, _ptr(q_tokens), _ptr(q_real), _ptr(q_gaps), _ptr(q_off), C.c_int64(n_q), C.byref(p),
                                      C.c_int(nthreads), C.c_int64(cap), _ptr(out), _ptr(cnt), None)
        return _split(out, cnt, cap), cnt

    def match_debug(self, pattern, cap=64, dbg_cap=4096, **kw):
        pattern = np.ascontiguousarray(pattern, dtype=np.int32)
        p = make_params(Params, **kw)
        out = np.zeros(cap, dtype=MATCH_DTYPE)
        dbg = np.zeros(dbg_cap, dtype=CAND_DTYPE)
        dbg_n = C.c_int64(0)
        n = self.lib.fmo_match_debug(self.h, _ptr(pattern), C.c_int64(len(pattern)), C.byref(p), C.c_int64(cap),
                                     _ptr(out), _ptr(dbg), C.c_int64(dbg_cap), C.byref(dbg_n))
        return out[:min(n, cap)].copy(), dbg[:dbg_n.value].copy()

    def subsequence_batch(self, q_tokens, q_off, n=1, no_perfect=False, ml=3, mr=0.3, idf_weighting=False):
        """fmo_subsequence_batch: FuzzyMatch::subsequence per pattern -> structured array (found, s_id, score, cost, position, length)."""
        q_tokens, q_off = _csr(q_tokens, q_off)
        n_q = len(q_off) - 1
        out = np.zeros(n_q, dtype=SUBSEQ_DTYPE)
        self.lib.fmo_subsequence_batch(self.h, _ptr(q_tokens), _ptr(q_off), C.c_int64(n_q), C.c_int32(n), C.c_int32(int(no_perfect)),
                                       C.c_int32(ml), C.c_float(mr), C.c_int32(int(idf_weighting)), _ptr(out))
        return out

    def equal_range(self, ngram):
        ngram = np.ascontiguousarray(ngram, dtype=np.int32)
        lo, hi = C.c_int64(0), C.c_int64(0)
        self.lib.fmo_equal_range(self.h, _ptr(ngram), C.c_int64(len(ngram)), C.byref(lo), C.byref(hi))
        return lo.value, hi.value


SUBSEQ_DTYPE = np.dtype([("s_id", np.uint32), ("score", np.float32), ("cost", np.float32), ("position", np.int32), ("length", np.int32),
                         ("found", np.int32)])
REF_SUBSEQ_DTYPE = np.dtype([("s_id", np.uint32), ("score", np.float32), ("max_subseq", np.int32), ("found", np.int32)])


def _scrub_uninitialised_penalty(res):
    """FuzzyMatch::Match::penalty is never initialised (include/fuzzy/fuzzy_match.hh:34-40); the first match
    picked by the contrastive rerank therefore carries stack garbage (observed: a denormal). The reference's
    own tests rely on it reading as 0 (test/test.cc:537), so report 0."""
    for r in res:
        r["penalty"][~(np.abs(r["penalty"]) >= 1e-30)] = 0.0
    return res


def ref_available():
    return os.path.exists(REF_SO)


class RefIndex:
    """The unmodified reference (fuzzy::FuzzyMatch) behind oracle/ref_driver.cc. With real / gaps / itok
    table it is built through add_tm(id, Sentence, Tokens) and queried through match(Sentence, Tokens, ...)."""

    def __init__(self, tokens, off, max_tokens=300, real=None, gaps=None, itok_blob=None, itok_off=None):
        self.lib = lib = C.CDLL(REF_SO)
        lib.fmref_create.restype = C.c_void_p
        lib.fmref_match_batch.restype = C.c_double
        lib.fmref_match_batch_real.restype = C.c_double
        tokens, off = _csr(tokens, off)
        self.h = C.c_void_p(lib.fmref_create(C.c_int(max_tokens)))
        if real is None:
            lib.fmref_add_tm(self.h, _ptr(tokens), _ptr(off), C.c_int64(len(off) - 1))
        else:
            real = np.ascontiguousarray(real, dtype=np.int32)
            gaps = np.ascontiguousarray(gaps, dtype=np.int32)
            self._itok = (np.ascontiguousarray(itok_blob, dtype=np.uint8), np.ascontiguousarray(itok_off, dtype=np.int32))
            lib.fmref_add_tm_real(self.h, _ptr(tokens), _ptr(real), _ptr(gaps), _ptr(off), C.c_int64(len(off) - 1),
                                  _ptr(self._itok[0]), _ptr(self._itok[1]))
        lib.fmref_sort(self.h)
        self.last_seconds = 0.0

    def match_batch_real(self, q_tokens, q_real, q_gaps, q_off, cap=16, nthreads=1, no_perfect=False, **kw):
        q_tokens, q_off = _csr(q_tokens, q_off)
        q_real = np.ascontiguousarray(q_real, dtype=np.int32)
        q_gaps = np.ascontiguousarray(q_gaps, dtype=np.int32)
        n_q = len(q_off) - 1
        p = make_params(RefParams, **kw)
        out = np.zeros(n_q * cap, dtype=REF_MATCH_DTYPE)
        cnt = np.zeros(n_q, dtype=np.int32)
        self.last_seconds = self.lib.fmref_match_batch_real(self.h, _ptr(q_tokens), _ptr(q_real), _ptr(q_gaps), _ptr(q_off),
                                                            C.c_int64(n_q), C.byref(p), C.c_int(int(no_perfect)), C.c_int(nthreads),
                                                            C.c_int64(cap), _ptr(out), _ptr(cnt), _ptr(self._itok[0]), _ptr(self._itok[1]))
        return _scrub_uninitialised_penalty(_split(out, cnt, cap)), cnt

    def match_batch_twice(self, q1_tokens, q1_off, q2_tokens, q2_off, params1, params2, cap=16):
        """Per query: matches = []; match(pattern 1, **params1, matches); match(pattern 2, **params2, matches).
        Returns (prior lists, prior counts, appended lists, appended counts)."""
        q1_tokens, q1_off = _csr(q1_tokens, q1_off)
        q2_tokens, q2_off = _csr(q2_tokens, q2_off)
        n_q = len(q1_off) - 1
        assert len(q2_off) - 1 == n_q
        p1, p2 = make_params(RefParams, **params1), make_params(RefParams, **params2)
        pri, out = np.zeros(n_q * cap, dtype=REF_MATCH_DTYPE), np.zeros(n_q * cap, dtype=REF_MATCH_DTYPE)
        pcnt, cnt = np.zeros(n_q, dtype=np.int32), np.zeros(n_q, dtype=np.int32)
        self.lib.fmref_match_batch_twice(self.h, _ptr(q1_tokens), _ptr(q1_off), _ptr(q2_tokens), _ptr(q2_off), C.c_int64(n_q), C.byref(p1),
                                         C.byref(p2), C.c_int64(cap), _ptr(pri), _ptr(pcnt), _ptr(out), _ptr(cnt))
        return _split(pri, pcnt, cap), pcnt, _split(out, cnt, cap), cnt

    def subsequence_batch(self, q_tokens, q_off, n=1, no_perfect=False, ml=3, mr=0.3, idf_weighting=False):
        """FuzzyMatch::subsequence(string, ...) of the reference per pattern -> (records, texts): texts[q] is the
        detokenised sub-sequence the reference appends to Match::id (decimal token strings joined by blanks)."""
        q_tokens, q_off = _csr(q_tokens, q_off)
        n_q = len(q_off) - 1
        out = np.zeros(n_q, dtype=REF_SUBSEQ_DTYPE)
        stride = 16 * int(max(1, np.diff(q_off).max() if n_q else 1)) + 16
        text = np.zeros(n_q * stride, dtype=np.uint8)
        self.lib.fmref_subsequence_batch(self.h, _ptr(q_tokens), _ptr(q_off), C.c_int64(n_q), C.c_int32(n), C.c_int32(int(no_perfect)),
                                         C.c_int32(ml), C.c_float(mr), C.c_int32(int(idf_weighting)), _ptr(out), _ptr(text), C.c_int64(stride))
        texts = [bytes(text[q * stride:(q + 1) * stride]).split(b"\0")[0].decode() for q in range(n_q)]
        return out, texts

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.fmref_destroy(self.h)
            self.h = None

    def match_batch(self, q_tokens, q_off, cap=16, nthreads=1, **kw):
        q_tokens, q_off = _csr(q_tokens, q_off)
        n_q = len(q_off) - 1
        p = make_params(RefParams, **kw)
        out = np.zeros(n_q * cap, dtype=REF_MATCH_DTYPE)
        cnt = np.zeros(n_q, dtype=np.int32)
        self.last_seconds = self.lib.fmref_match_batch(self.h, _ptr(q_tokens), _ptr(q_off), C.c_int64(n_q), C.byref(p),
                                                       C.c_int(nthreads), C.c_int64(cap), _ptr(out), _ptr(cnt))
        return _scrub_uninitialised_penalty(_split(out, cnt, cap)), cnt
