"""Deterministic synthetic translation memories and query sets.

Counter-based splitmix64, so every draw is a pure function of (seed, stream, index) and the same
arrays come out on any machine / numpy version: vocabulary of V word ids drawn Zipf(s=1) by inverse
CDF, sentence length uniform in [len_lo, len_hi]; queries are 80 % perturbed TM sentences
(per token 5 % delete / 5 % replace / 5 % insert-after) and 20 % fresh random sentences.
Word ids start at 2 (0 = sentence separator, 1 = unknown; reference src/vocab_indexer.cc:10-11).
"""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def _uniform(seed, stream, n, start=0):
    """n doubles in [0,1) from (seed, stream, start..start+n)."""
    with np.errstate(over="ignore"):
        base = _splitmix64(np.array([seed * 1000003 + stream], dtype=np.uint64))[0]
        idx = np.arange(start, start + n, dtype=np.uint64)
        r = _splitmix64(idx * np.uint64(0xD1342543DE82EF95) + base)
    return (r >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


class _Zipf:
    def __init__(self, vocab):
        w = 1.0 / np.arange(1, vocab + 1, dtype=np.float64)
        self.cdf = np.cumsum(w / w.sum())
        self.cdf[-1] = 1.0

    def draw(self, u):
        return (np.searchsorted(self.cdf, u, side="right").astype(np.int32) + 2).astype(np.int32)


def make_tm(n_sent, vocab=50000, len_lo=5, len_hi=25, seed=1234, n_long=0, long_lo=200, long_hi=300):
    """Returns (tokens int32, off int64, vocab_size). The last n_long sentences are long ones."""
    z = _Zipf(vocab)
    lens = (len_lo + np.floor(_uniform(seed, 1, n_sent) * (len_hi - len_lo + 1))).astype(np.int64)
    if n_long:
        lens[n_sent - n_long:] = (long_lo + np.floor(_uniform(seed, 3, n_long) * (long_hi - long_lo + 1))).astype(np.int64)
    off = np.zeros(n_sent + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    tokens = z.draw(_uniform(seed, 2, int(off[-1])))
    return tokens, off, vocab + 2


def make_queries(tm_tokens, tm_off, n_q, vocab=50000, seed=5678, len_lo=5, len_hi=25, frac_random=0.2,
                 p_del=0.05, p_sub=0.05, p_ins=0.05, source_ids=None):
    """Returns (q_tokens int32, q_off int64). source_ids restricts the perturbed sources."""
    z = _Zipf(vocab)
    n_sent = len(tm_off) - 1
    kind = _uniform(seed, 1, n_q)
    pool = np.arange(n_sent, dtype=np.int64) if source_ids is None else np.asarray(source_ids, dtype=np.int64)
    src = pool[np.minimum((_uniform(seed, 2, n_q) * len(pool)).astype(np.int64), len(pool) - 1)]
    rnd_len = (len_lo + np.floor(_uniform(seed, 3, n_q) * (len_hi - len_lo + 1))).astype(np.int64)
    is_rnd = kind < frac_random
    base_len = np.where(is_rnd, rnd_len, tm_off[src + 1] - tm_off[src])
    boff = np.zeros(n_q + 1, dtype=np.int64)
    np.cumsum(base_len, out=boff[1:])
    total = int(boff[-1])
    qid = np.repeat(np.arange(n_q, dtype=np.int64), base_len)
    within = np.arange(total, dtype=np.int64) - boff[qid]
    # base tokens: random draw or copy of the source sentence
    rnd_tok = z.draw(_uniform(seed, 4, total))
    src_pos = np.where(is_rnd[qid], 0, tm_off[src[qid]] + within)
    base_tok = np.where(is_rnd[qid], rnd_tok, tm_tokens[src_pos])
    # perturbation (perturbed sources only)
    u = _uniform(seed, 5, total)
    pert = ~is_rnd[qid]
    dele = pert & (u < p_del)
    sub = pert & (u >= p_del) & (u < p_del + p_sub)
    ins = pert & (u >= p_del + p_sub) & (u < p_del + p_sub + p_ins)
    sub_tok = z.draw(_uniform(seed, 6, total))
    ins_tok = z.draw(_uniform(seed, 7, total))
    tok = np.where(sub, sub_tok, base_tok)
    emit = np.where(dele, 0, np.where(ins, 2, 1)).astype(np.int64)
    eoff = np.zeros(total + 1, dtype=np.int64)
    np.cumsum(emit, out=eoff[1:])
    out = np.empty(int(eoff[-1]), dtype=np.int32)
    keep = emit >= 1
    out[eoff[:-1][keep]] = tok[keep]
    out[eoff[:-1][ins] + 1] = ins_tok[ins]
    q_len = np.add.reduceat(emit, boff[:-1][base_len > 0]) if total else np.zeros(0, dtype=np.int64)
    full_len = np.zeros(n_q, dtype=np.int64)
    full_len[base_len > 0] = q_len
    q_off = np.zeros(n_q + 1, dtype=np.int64)
    np.cumsum(full_len, out=q_off[1:])
    return out, q_off


ITOKS = [b"", b" ", b",", b".", b"T", b", ", b"...", b"(", b")"]  # id 0 = no penalty token


def itok_table(itoks=ITOKS):
    """(blob uint8, off int32[K+1]) of the penalty-token strings; id 0 is the empty string."""
    off = np.zeros(len(itoks) + 1, dtype=np.int32)
    np.cumsum([len(x) for x in itoks], out=off[1:])
    return np.frombuffer(b"".join(itoks) + b"\0", dtype=np.uint8).copy(), off


def make_real(tokens, off, seed, p_variant=0.15, p_case=0.1, p_itok=0.15, n_itok=len(ITOKS)):
    """Synthetic real tokens and penalty tokens for the Sentence API: real[k] = (form id << 1) | case class,
    where the form id is word*4 + variant; gaps = n+1 itok ids per sentence at off[s] + s."""
    n_tok, n_sent = len(tokens), len(off) - 1
    u = _uniform(seed, 11, n_tok)
    variant = (u < p_variant).astype(np.int32)
    cls = (_uniform(seed, 12, n_tok) < p_case).astype(np.int32)
    real = (((tokens.astype(np.int64) * 4 + variant) << 1) | cls).astype(np.int32)
    n_gap = n_tok + n_sent
    ug = _uniform(seed, 13, n_gap)
    pick = (1 + np.floor(_uniform(seed, 14, n_gap) * (n_itok - 1))).astype(np.int32)
    gaps = np.where(ug < p_itok, pick, 0).astype(np.int32)
    return real, gaps
