"""ctypes binding of the C ABI declared in include/fuzzy_match_b200.h.

This is the binding a maintainer of the reference would write for the hot path;
nothing here computes anything -- all work happens in libfm_b200.so (hand-written sm_100a CUDA).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class FuzzyMatchError(RuntimeError):
    pass


def library_path():
    # FM_B200_LIB selects another build of the same library (kernel tuning experiments)
    return os.environ.get("FM_B200_LIB") or os.path.join(_HERE, "libfm_b200.so")


def build_library(verbose=False):
    """Compile csrc/*.cu for sm_100a into libfm_b200.so (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout, out.stderr)
    if out.returncode != 0:
        raise FuzzyMatchError("building libfm_b200.so failed")
    return library_path()


class Params(C.Structure):
    """fm_params: the arguments of FuzzyMatch::match after the pattern (fuzzy_match.hh:59-82)."""
    _fields_ = [("fuzzy", C.c_float), ("number_of_matches", C.c_int32), ("no_perfect", C.c_int32),
                ("min_subseq_length", C.c_int32), ("min_subseq_ratio", C.c_float),
                ("vocab_idf_penalty", C.c_float), ("insert_cost", C.c_float), ("delete_cost", C.c_float),
                ("replace_cost", C.c_float), ("contrastive_factor", C.c_float),
                ("contrast_reduce", C.c_int32), ("contrast_buffer", C.c_int32)]

    @classmethod
    def make(cls, fuzzy=0.7, n=1, ml=2, mr=0.0, idf=0.0, costs=(1.0, 1.0, 1.0), contrast=0.0, reduce=0, buffer=-1,
             no_perfect=False):
        p = cls()
        p.fuzzy, p.number_of_matches, p.no_perfect = fuzzy, n, int(no_perfect)
        p.min_subseq_length, p.min_subseq_ratio, p.vocab_idf_penalty = ml, mr, idf
        p.insert_cost, p.delete_cost, p.replace_cost = costs
        p.contrastive_factor, p.contrast_reduce, p.contrast_buffer = contrast, int(reduce), buffer
        return p


class Profile(C.Structure):
    _fields_ = [("ms_prepare", C.c_float), ("ms_search", C.c_float), ("ms_gather", C.c_float), ("ms_scan", C.c_float),
                ("ms_score", C.c_float), ("ms_replay", C.c_float), ("ms_total", C.c_float),
                ("n_queries", C.c_int64), ("n_query_tokens", C.c_int64), ("n_slices", C.c_int64),
                ("n_elements", C.c_int64), ("n_survivors", C.c_int64), ("n_matches", C.c_int64),
                ("launches", C.c_int32), ("retries", C.c_int32), ("n_stage2", C.c_int64), ("n_verified", C.c_int64),
                ("ms_walk", C.c_float), ("ms_verify", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


MATCH_DTYPE = np.dtype([("s_id", np.uint32), ("score", np.float32), ("penalty", np.float32),
                        ("max_subseq", np.int32), ("length", np.int32), ("cost", np.float32)])
SUBSEQ_DTYPE = np.dtype([("s_id", np.uint32), ("score", np.float32), ("cost", np.float32), ("position", np.int32),
                         ("length", np.int32), ("found", np.int32)])
WIRE_DTYPE = np.dtype([("s_id", np.uint32), ("lm_len", np.uint32), ("cost", np.float32), ("rowmin_max", np.float32)])

EXPORTS = ["fm_index_create", "fm_index_destroy", "fm_index_save", "fm_index_load", "fm_index_num_sentences", "fm_index_num_suffixes",
           "fm_index_max_tokens_in_pattern", "fm_index_device_bytes", "fm_index_kept_sources", "fm_index_sfreq",
           "fm_index_sentence", "fm_index_set_idf_stats", "fm_index_set_real", "fm_match_batch", "fm_match_batch_real", "fm_match_batch_device",
           "fm_subsequence_batch", "fm_match_batch_prior",
           "fm_match_batch_submit", "fm_match_batch_device_submit", "fm_ticket_wait", "fm_wire_block_bytes", "fm_shard_accept_device",
           "fm_merge_accepted_device", "fm_comm_unique_id", "fm_comm_create", "fm_comm_destroy", "fm_match_batch_sharded_device", "fm_match_batch_sharded_submit",
           "fm_comm_last_gather_bytes", "fm_comm_block_capacity", "fm_set_profiling", "fm_get_profile", "fm_last_error", "fm_version"]


def load_library():
    """Loads libfm_b200.so; raises (never falls back) when it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise FuzzyMatchError("%s not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)" % path)
    lib = C.CDLL(path)
    lib.fm_last_error.restype = C.c_char_p
    lib.fm_version.restype = C.c_char_p
    for name in ("fm_index_num_sentences", "fm_index_num_suffixes", "fm_index_device_bytes"):
        getattr(lib, name).restype = C.c_int64
        getattr(lib, name).argtypes = [C.c_void_p]
    lib.fm_index_max_tokens_in_pattern.restype = C.c_int32
    lib.fm_index_max_tokens_in_pattern.argtypes = [C.c_void_p]
    lib.fm_index_destroy.restype = None
    lib.fm_index_destroy.argtypes = [C.c_void_p]
    lib.fm_index_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                    C.c_int64, C.c_int, C.POINTER(C.c_void_p)]
    lib.fm_index_save.argtypes = [C.c_void_p, C.c_char_p]
    lib.fm_index_load.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    lib.fm_index_kept_sources.argtypes = [C.c_void_p, C.c_void_p]
    lib.fm_index_sfreq.argtypes = [C.c_void_p, C.c_void_p]
    lib.fm_index_set_idf_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    lib.fm_index_sentence.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int32)]
    lib.fm_match_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(Params), C.c_int64,
                                   C.c_void_p, C.c_void_p]
    lib.fm_match_batch_prior.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(Params), C.c_void_p, C.c_void_p,
                                         C.c_int64, C.c_void_p, C.c_void_p]
    lib.fm_subsequence_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                         C.c_int32, C.c_void_p]
    lib.fm_index_set_real.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    lib.fm_match_batch_real.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(Params),
                                        C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]
    lib.fm_match_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.POINTER(Params),
                                          C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.fm_match_batch_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(Params), C.c_int64,
                                          C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.fm_match_batch_device_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.POINTER(Params),
                                                 C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.fm_ticket_wait.argtypes = [C.c_void_p]
    lib.fm_wire_block_bytes.restype = C.c_int64
    lib.fm_wire_block_bytes.argtypes = [C.c_int64, C.c_int64]
    lib.fm_shard_accept_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.POINTER(Params), C.c_int64,
                                           C.c_void_p, C.c_void_p]
    lib.fm_merge_accepted_device.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.c_int64, C.c_void_p, C.c_int64, C.POINTER(Params),
                                             C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p]
    lib.fm_comm_unique_id.argtypes = [C.c_void_p]
    lib.fm_comm_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    lib.fm_comm_destroy.argtypes = [C.c_void_p]
    lib.fm_comm_destroy.restype = None
    lib.fm_match_batch_sharded_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.POINTER(Params),
                                                  C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.fm_match_batch_sharded_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.POINTER(Params),
                                                  C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.fm_comm_last_gather_bytes.restype = C.c_int64
    lib.fm_comm_last_gather_bytes.argtypes = [C.c_void_p]
    lib.fm_comm_block_capacity.restype = C.c_int64
    lib.fm_comm_block_capacity.argtypes = [C.c_void_p]
    lib.fm_set_profiling.argtypes = [C.c_void_p, C.c_int]
    lib.fm_get_profile.argtypes = [C.c_void_p, C.POINTER(Profile)]
    _LIB = lib
    return lib


def _check(lib, rc):
    if rc != 0:
        raise FuzzyMatchError("fuzzy_match_b200 error %d: %s" % (rc, lib.fm_last_error().decode()))


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def itok_distance_table(itoks):
    """Pairwise _edit_distance_char (reference include/fuzzy/edit_distance.hxx:7-35) of penalty-token byte
    strings; row / column 0 (the empty string) hold the lengths. Host-side table, a few entries."""
    k = len(itoks)
    dist = np.zeros((k, k), dtype=np.int32)
    for a in range(k):
        for b in range(k):
            s1, s2 = itoks[a], itoks[b]
            if not s1 or not s2:
                dist[a, b] = len(s1) + len(s2)
                continue
            prev = list(range(len(s2) + 1))
            for i in range(1, len(s1) + 1):
                cur = [i] + [0] * len(s2)
                for j in range(1, len(s2) + 1):
                    cur[j] = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (0 if s1[i - 1] == s2[j - 1] else 1))
                prev = cur
            dist[a, b] = prev[len(s2)]
    return dist


class Index:
    """One device-resident index (fm_index) and the calls that run batches against it."""

    def __init__(self, tokens, sent_off, vocab_size, max_tokens=300, sfreq_global=None, n_sent_global=0, s_id_base=0,
                 device=0):
        self.lib = load_library()
        tokens = np.ascontiguousarray(tokens, dtype=np.int32)
        sent_off = np.ascontiguousarray(sent_off, dtype=np.int64)
        sf = None if sfreq_global is None else np.ascontiguousarray(sfreq_global, dtype=np.uint32)
        h = C.c_void_p()
        _check(self.lib, self.lib.fm_index_create(_ptr(tokens), _ptr(sent_off), len(sent_off) - 1, int(vocab_size),
                                                  int(max_tokens), _ptr(sf), int(n_sent_global), int(s_id_base),
                                                  int(device), C.byref(h)))
        self.h = h
        self.vocab_size = int(vocab_size)
        self.device = int(device)

    @classmethod
    def load(cls, path, vocab_size, device=0):
        """fm_index_load: an index written by save()."""
        self = cls.__new__(cls)
        self.lib = load_library()
        h = C.c_void_p()
        _check(self.lib, self.lib.fm_index_load(str(path).encode(), int(device), C.byref(h)))
        self.h, self.vocab_size, self.device = h, int(vocab_size), int(device)
        return self

    def save(self, path):
        _check(self.lib, self.lib.fm_index_save(self.h, str(path).encode()))

    def close(self):
        if getattr(self, "h", None):
            self.lib.fm_index_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    num_sentences = property(lambda self: self.lib.fm_index_num_sentences(self.h))
    num_suffixes = property(lambda self: self.lib.fm_index_num_suffixes(self.h))
    max_tokens_in_pattern = property(lambda self: self.lib.fm_index_max_tokens_in_pattern(self.h))
    device_bytes = property(lambda self: self.lib.fm_index_device_bytes(self.h))

    def kept_sources(self):
        out = np.zeros(self.num_sentences, dtype=np.int64)
        _check(self.lib, self.lib.fm_index_kept_sources(self.h, _ptr(out)))
        return out

    def sfreq(self):
        out = np.zeros(self.vocab_size, dtype=np.uint32)
        _check(self.lib, self.lib.fm_index_sfreq(self.h, _ptr(out)))
        return out

    def set_idf_stats(self, sfreq_global, n_sent_global):
        sf = np.ascontiguousarray(sfreq_global, dtype=np.uint32)
        assert len(sf) == self.vocab_size
        _check(self.lib, self.lib.fm_index_set_idf_stats(self.h, _ptr(sf), int(n_sent_global)))

    def sentence(self, s_id):
        p, n = C.POINTER(C.c_int32)(), C.c_int32()
        _check(self.lib, self.lib.fm_index_sentence(self.h, int(s_id), C.byref(p), C.byref(n)))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def match_batch(self, q_tokens, q_off, cap=None, params=None, out=None, cnt=None, **kw):
        """Host buffers in, host buffers out (fm_match_batch). Returns (matches[n_q, cap], counts[n_q]).
        out / cnt may be preallocated (e.g. views of pinned memory)."""
        p = params if params is not None else Params.make(**kw)
        if cap is None:
            cap = max(1, p.number_of_matches)
        q_tokens = np.ascontiguousarray(q_tokens, dtype=np.int32)
        q_off = np.ascontiguousarray(q_off, dtype=np.int64)
        n_q = len(q_off) - 1
        if out is None:
            out = np.zeros((n_q, cap), dtype=MATCH_DTYPE)
        if cnt is None:
            cnt = np.zeros(n_q, dtype=np.int32)
        assert out.dtype == MATCH_DTYPE and out.size == n_q * cap and out.flags.c_contiguous and len(cnt) == n_q
        _check(self.lib, self.lib.fm_match_batch(self.h, _ptr(q_tokens), _ptr(q_off), n_q, C.byref(p), cap, _ptr(out),
                                                 _ptr(cnt)))
        return out, cnt

    def match_batch_prior(self, q_tokens, q_off, prior_sid, prior_off, cap=None, params=None, **kw):
        """fm_match_batch_prior: match() into result vectors that already hold the sentences prior_sid (CSR per query)."""
        p = params if params is not None else Params.make(**kw)
        if cap is None:
            cap = max(1, p.number_of_matches)
        q_tokens = np.ascontiguousarray(q_tokens, dtype=np.int32)
        q_off = np.ascontiguousarray(q_off, dtype=np.int64)
        prior_sid = np.ascontiguousarray(prior_sid, dtype=np.uint32)
        prior_off = np.ascontiguousarray(prior_off, dtype=np.int64)
        n_q = len(q_off) - 1
        assert len(prior_off) == n_q + 1
        out = np.zeros((n_q, cap), dtype=MATCH_DTYPE)
        cnt = np.zeros(n_q, dtype=np.int32)
        _check(self.lib, self.lib.fm_match_batch_prior(self.h, _ptr(q_tokens), _ptr(q_off), n_q, C.byref(p), _ptr(prior_sid),
                                                       _ptr(prior_off), cap, _ptr(out), _ptr(cnt)))
        return out, cnt

    def subsequence_batch(self, q_tokens, q_off, n=1, no_perfect=False, ml=3, mr=0.3, idf_weighting=False):
        """fm_subsequence_batch: FuzzyMatch::subsequence per pattern -> records[n_q] of SUBSEQ_DTYPE."""
        q_tokens = np.ascontiguousarray(q_tokens, dtype=np.int32)
        q_off = np.ascontiguousarray(q_off, dtype=np.int64)
        n_q = len(q_off) - 1
        out = np.zeros(n_q, dtype=SUBSEQ_DTYPE)
        _check(self.lib, self.lib.fm_subsequence_batch(self.h, _ptr(q_tokens), _ptr(q_off), n_q, int(n), int(no_perfect), int(ml),
                                                       float(mr), int(idf_weighting), _ptr(out)))
        return out

    def set_real(self, real, gaps, sent_off):
        """fm_index_set_real: real tokens ((form id << 1) | case class) and gap penalty-token ids of the TM."""
        real = np.ascontiguousarray(real, dtype=np.int32)
        gaps = np.ascontiguousarray(gaps, dtype=np.int32)
        sent_off = np.ascontiguousarray(sent_off, dtype=np.int64)
        _check(self.lib, self.lib.fm_index_set_real(self.h, _ptr(real), _ptr(gaps), _ptr(sent_off), len(sent_off) - 1))

    def match_batch_real(self, q_tokens, q_real, q_gaps, q_off, itok_dist, cap=None, params=None, **kw):
        """fm_match_batch_real: match(Sentence real, Tokens pattern, ...) for a batch (host buffers)."""
        p = params if params is not None else Params.make(**kw)
        if cap is None:
            cap = max(1, p.number_of_matches)
        q_tokens = np.ascontiguousarray(q_tokens, dtype=np.int32)
        q_real = np.ascontiguousarray(q_real, dtype=np.int32)
        q_gaps = np.ascontiguousarray(q_gaps, dtype=np.int32)
        q_off = np.ascontiguousarray(q_off, dtype=np.int64)
        dist = np.ascontiguousarray(itok_dist, dtype=np.int32)
        assert dist.ndim == 2 and dist.shape[0] == dist.shape[1]
        n_q = len(q_off) - 1
        out = np.zeros((n_q, cap), dtype=MATCH_DTYPE)
        cnt = np.zeros(n_q, dtype=np.int32)
        _check(self.lib, self.lib.fm_match_batch_real(self.h, _ptr(q_tokens), _ptr(q_real), _ptr(q_gaps), _ptr(q_off), n_q,
                                                      C.byref(p), _ptr(dist), dist.shape[0], cap, _ptr(out), _ptr(cnt)))
        return out, cnt

    def match_batch_device(self, d_q_tokens, d_q_off, n_q, n_tok, d_out, d_out_count, cap, stream=0, params=None, **kw):
        """Raw device pointers (ints) in and out (fm_match_batch_device)."""
        p = params if params is not None else Params.make(**kw)
        _check(self.lib, self.lib.fm_match_batch_device(self.h, d_q_tokens, d_q_off, n_q, n_tok, C.byref(p), cap, d_out,
                                                        d_out_count, stream))

    def submit(self, q_tokens, q_off, out, cnt, cap, params):
        """fm_match_batch_submit: returns a ticket; the arrays must stay alive and untouched until wait(ticket)."""
        assert q_tokens.dtype == np.int32 and q_off.dtype == np.int64 and out.dtype == MATCH_DTYPE and cnt.dtype == np.int32
        t = C.c_void_p()
        _check(self.lib, self.lib.fm_match_batch_submit(self.h, _ptr(q_tokens), _ptr(q_off), len(q_off) - 1, C.byref(params), cap,
                                                        _ptr(out), _ptr(cnt), C.byref(t)))
        return t

    def submit_device(self, d_q_tokens, d_q_off, n_q, n_tok, d_out, d_out_count, cap, stream, params):
        t = C.c_void_p()
        _check(self.lib, self.lib.fm_match_batch_device_submit(self.h, d_q_tokens, d_q_off, n_q, n_tok, C.byref(params), cap, d_out,
                                                               d_out_count, stream, C.byref(t)))
        return t

    def wait(self, ticket):
        _check(self.lib, self.lib.fm_ticket_wait(ticket))

    def shard_accept_device(self, d_q_tokens, d_q_off, n_q, n_tok, capacity, d_block, stream=0, params=None, **kw):
        """fm_shard_accept_device: this shard's accepted records into d_block (fm_wire_block_bytes(n_q, capacity) device bytes)."""
        p = params if params is not None else Params.make(**kw)
        _check(self.lib, self.lib.fm_shard_accept_device(self.h, d_q_tokens, d_q_off, n_q, n_tok, C.byref(p), capacity, d_block, stream))

    def merge_accepted_device(self, d_blocks, total_capacity, d_q_off, n_q, d_out, d_out_count, cap, stream=0, params=None, **kw):
        """fm_merge_accepted_device over the blocks of all shards; returns the block capacity a rerun needs (0: complete)."""
        p = params if params is not None else Params.make(**kw)
        k = len(d_blocks)
        blocks = (C.c_void_p * k)(*d_blocks)
        need = C.c_int64(0)
        _check(self.lib, self.lib.fm_merge_accepted_device(self.h, k, blocks, total_capacity, d_q_off, n_q, C.byref(p), cap, d_out,
                                                           d_out_count, C.byref(need), stream))
        return need.value

    def submit_sharded_device(self, comm, d_q_tokens, d_q_off, n_q, n_tok, d_out, d_out_count, cap, stream, params):
        """fm_match_batch_sharded_submit: returns a ticket for wait(); collective -- every rank in the same order."""
        t = C.c_void_p()
        _check(self.lib, self.lib.fm_match_batch_sharded_submit(self.h, comm, d_q_tokens, d_q_off, n_q, n_tok, C.byref(params), cap, d_out,
                                                                d_out_count, stream, C.byref(t)))
        return t

    def match_batch_sharded_device(self, comm, d_q_tokens, d_q_off, n_q, n_tok, d_out, d_out_count, cap, stream=0, params=None, **kw):
        p = params if params is not None else Params.make(**kw)
        _check(self.lib, self.lib.fm_match_batch_sharded_device(self.h, comm, d_q_tokens, d_q_off, n_q, n_tok, C.byref(p), cap, d_out,
                                                                d_out_count, stream))

    def set_profiling(self, enabled=True):
        _check(self.lib, self.lib.fm_set_profiling(self.h, int(enabled)))

    def profile(self):
        p = Profile()
        _check(self.lib, self.lib.fm_get_profile(self.h, C.byref(p)))
        return p.as_dict()


def wire_block_bytes(n_q, capacity):
    return int(load_library().fm_wire_block_bytes(n_q, capacity))


def comm_unique_id():
    """fm_comm_unique_id: 128 bytes that rank 0 hands to every rank."""
    lib = load_library()
    buf = (C.c_char * 128)()
    _check(lib, lib.fm_comm_unique_id(buf))
    return bytes(buf)


class Comm:
    """fm_comm: the NCCL communicator of a sharded TM (one rank per GPU)."""

    def __init__(self, unique_id, rank, world, device):
        self.lib = load_library()
        self.h = C.c_void_p()
        assert len(unique_id) == 128
        _check(self.lib, self.lib.fm_comm_create(unique_id, rank, world, device, C.byref(self.h)))

    @property
    def last_gather_bytes(self):
        return int(self.lib.fm_comm_last_gather_bytes(self.h))

    @property
    def block_capacity(self):
        return int(self.lib.fm_comm_block_capacity(self.h))

    def close(self):
        if self.h:
            self.lib.fm_comm_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
