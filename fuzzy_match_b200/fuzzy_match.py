"""Host-side mirror of the reference's fuzzy::FuzzyMatch for the pre-tokenised path.

Same names, argument meaning, defaults and error behaviour as include/fuzzy/fuzzy_match.hh:17-119
(add_tm(id, Tokens, sort) :52, sort() :57, match(Tokens, fuzzy, N, matches, ml=2, mr=0, idf=0,
EditCosts(), contrast=0, reduce=MEAN, buffer=-1) :59-69, subsequence(pattern, N, no_perfect, matches,
ml=3, mr=0.3, idf_weighting=False) :96-102 behind its tokenizer, max_tokens_in_pattern() :119), so the
reference's own Tokens-API tests read the same against this class. The vocabulary (string -> id,
reference src/vocab_indexer.cc) stays on the host; everything match() computes runs on the GPU
through the C ABI. The tokenizer front-end (match(std::string)) is out of scope.
"""
import enum
from dataclasses import dataclass

import numpy as np

from . import capi

DEFAULT_MAX_TOKENS_IN_PATTERN = 300  # include/fuzzy/suffix_array_index.hh:15


class ContrastReduce(enum.IntEnum):  # include/fuzzy/fuzzy_match.hh:15
    MEAN = 0
    MAX = 1


@dataclass(frozen=True)
class EditCosts:  # include/fuzzy/costs.hh:7-29
    insert_cost: float = 1.0
    delete_cost: float = 1.0
    replace_cost: float = 1.0


@dataclass
class Match:  # include/fuzzy/fuzzy_match.hh:32-46
    score: float
    penalty: float
    max_subseq: int
    s_id: int
    id: str
    length: int
    s: np.ndarray  # word ids of the matched TM sentence


class FuzzyMatch:
    SENTENCE_SEPARATOR = 0  # src/vocab_indexer.cc:10-11
    VOCAB_UNK = 1

    def __init__(self, pt=0, max_tokens_in_pattern=DEFAULT_MAX_TOKENS_IN_PATTERN, device=0):
        if pt != 0:
            raise NotImplementedError("penalty tokens need the tokenizer front-end, which is out of scope")
        self._max_tokens = int(max_tokens_in_pattern)
        self._device = device
        self._vocab = {}
        self._ids = []
        self._sentences = []
        self._index = None

    # -- build side -------------------------------------------------------------------------
    def add_tm(self, id, norm, sort=True):
        """add_tm(id, Tokens, sort): sentences that are empty or longer than the cap are ignored
        (src/suffix_array_index.cc:16) but the call still returns True (src/fuzzy_match.cc:196-203)."""
        norm = list(norm)
        if norm and len(norm) <= self._max_tokens:
            self._sentences.append([self._vocab.setdefault(w, len(self._vocab) + 2) for w in norm])
            self._ids.append(id)
            self._index = None
        if sort:
            self.sort()
        return True

    def sort(self):
        if self._index is not None:
            return
        off = np.zeros(len(self._sentences) + 1, dtype=np.int64)
        if self._sentences:
            np.cumsum([len(s) for s in self._sentences], out=off[1:])
        tok = np.fromiter((t for s in self._sentences for t in s), dtype=np.int32, count=int(off[-1]))
        self._index = capi.Index(tok, off, len(self._vocab) + 2, max_tokens=self._max_tokens, device=self._device)

    def max_tokens_in_pattern(self):
        return self._max_tokens

    # -- query side -------------------------------------------------------------------------
    def _wids(self, pattern):
        return [self._vocab.get(w, self.VOCAB_UNK) for w in pattern]

    def match(self, pattern, fuzzy, number_of_matches, matches, min_subseq_length=2, min_subseq_ratio=0.0,
              vocab_idf_penalty=0.0, edit_costs=EditCosts(), contrastive_factor=0.0, reduce=ContrastReduce.MEAN,
              contrast_buffer=-1, no_perfect=False):
        """Appends to `matches` and returns len(matches) > 0, like the reference: entries already in `matches` count
        against number_of_matches and take part in the contrastive penalties (src/fuzzy_match.cc:626-679)."""
        self.match_batch([pattern], fuzzy, number_of_matches, [matches], min_subseq_length, min_subseq_ratio,
                         vocab_idf_penalty, edit_costs, contrastive_factor, reduce, contrast_buffer, no_perfect)
        return len(matches) > 0

    def match_batch(self, patterns, fuzzy, number_of_matches, matches_out, min_subseq_length=2, min_subseq_ratio=0.0,
                    vocab_idf_penalty=0.0, edit_costs=EditCosts(), contrastive_factor=0.0, reduce=ContrastReduce.MEAN,
                    contrast_buffer=-1, no_perfect=False):
        """The batched front-end: many patterns through one launch sequence."""
        if self._index is None:
            self.sort()
        wids = [self._wids(p) for p in patterns]
        q_off = np.zeros(len(wids) + 1, dtype=np.int64)
        if wids:
            np.cumsum([len(w) for w in wids], out=q_off[1:])
        q_tok = np.fromiter((t for w in wids for t in w), dtype=np.int32, count=int(q_off[-1]))
        params = capi.Params.make(fuzzy=fuzzy, n=number_of_matches, ml=min_subseq_length, mr=min_subseq_ratio,
                                  idf=vocab_idf_penalty,
                                  costs=(edit_costs.insert_cost, edit_costs.delete_cost, edit_costs.replace_cost),
                                  contrast=contrastive_factor, reduce=int(reduce), buffer=contrast_buffer,
                                  no_perfect=no_perfect)
        cap = max(1, number_of_matches) if number_of_matches > 0 else 64
        prior_off = np.zeros(len(wids) + 1, dtype=np.int64)
        np.cumsum([len(m) for m in matches_out], out=prior_off[1:])
        prior_sid = np.fromiter((m.s_id for ms in matches_out for m in ms), dtype=np.uint32, count=int(prior_off[-1]))
        while True:
            if prior_off[-1]:
                out, cnt = self._index.match_batch_prior(q_tok, q_off, prior_sid, prior_off, cap=cap, params=params)
            else:
                out, cnt = self._index.match_batch(q_tok, q_off, cap=cap, params=params)
            if len(cnt) == 0 or cnt.max() <= cap:
                break
            cap = int(cnt.max())  # number_of_matches == 0 returns everything: rerun with room for it
        for q, dst in enumerate(matches_out):
            for m in out[q, :cnt[q]]:
                sid = int(m["s_id"])
                dst.append(Match(score=float(m["score"]), penalty=float(m["penalty"]), max_subseq=int(m["max_subseq"]),
                                 s_id=sid, id=self._ids[sid], length=int(m["length"]),
                                 s=np.asarray(self._sentences[sid], dtype=np.int32)))
        return [len(m) > 0 for m in matches_out]

    def subsequence(self, pattern, number_of_matches, no_perfect, matches, min_subseq_length=3, min_subseq_ratio=0.3,
                    idf_weighting=False):
        """subsequence() for a tokenised pattern: appends at most one Match (fuzzy_match.cc:360-363) whose id is
        "<tm id>\t<the sub-sequence, tokens joined by blanks>" and returns whether one was found."""
        return self.subsequence_batch([pattern], number_of_matches, no_perfect, [matches], min_subseq_length, min_subseq_ratio,
                                      idf_weighting)[0]

    def subsequence_batch(self, patterns, number_of_matches, no_perfect, matches_out, min_subseq_length=3, min_subseq_ratio=0.3,
                          idf_weighting=False):
        if self._index is None:
            self.sort()
        wids = [self._wids(p) for p in patterns]
        q_off = np.zeros(len(wids) + 1, dtype=np.int64)
        if wids:
            np.cumsum([len(w) for w in wids], out=q_off[1:])
        q_tok = np.fromiter((t for w in wids for t in w), dtype=np.int32, count=int(q_off[-1]))
        rec = self._index.subsequence_batch(q_tok, q_off, n=number_of_matches, no_perfect=no_perfect, ml=min_subseq_length,
                                            mr=min_subseq_ratio, idf_weighting=idf_weighting)
        found = []
        for q, dst in enumerate(matches_out):
            r = rec[q]
            found.append(bool(r["found"]))
            if r["found"]:
                sid, pos, n = int(r["s_id"]), int(r["position"]), int(r["length"])
                dst.append(Match(score=float(r["score"]), penalty=0.0, max_subseq=n, s_id=sid,
                                 id=self._ids[sid] + "\t" + " ".join(list(patterns[q])[pos:pos + n]), length=0,
                                 s=np.asarray(self._sentences[sid], dtype=np.int32)))
        return found
