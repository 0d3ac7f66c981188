// fuzzy_match_b200.hh -- header-only C++ adapter: the reference's fuzzy::FuzzyMatch interface for the
// pre-tokenised path, on top of the C ABI (include/fuzzy_match_b200.h, libfm_b200.so).
//
// Mirrors include/fuzzy/fuzzy_match.hh:17-119 of SYSTRAN/fuzzy-match: same class, nested Match,
// EditCosts (include/fuzzy/costs.hh:7-29), ContrastReduce, Tokens, same argument order, defaults and
// return conventions (match() APPENDS to `matches` and returns matches.size() > 0; empty or
// over-long patterns return false; add_tm silently ignores empty / over-long sentences), so a caller
// of add_tm(id, Tokens) / sort() / match(Tokens, ...) recompiles against this header unchanged.
// The vocabulary (string -> id; reference src/vocab_indexer.cc) lives here on the host; everything
// match() computes runs on the GPU. Extra: match_batch() feeds many patterns through one launch
// sequence. The tokenizer front-end (match(std::string), penalty tokens) is out of scope.
//
// Define FUZZY_MATCH_B200_NAMESPACE before including to put the classes elsewhere than `fuzzy`.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/fuzzy_match_b200.h"

#ifndef FUZZY_MATCH_B200_NAMESPACE
#define FUZZY_MATCH_B200_NAMESPACE fuzzy
#endif

namespace FUZZY_MATCH_B200_NAMESPACE {

typedef std::vector<std::string> Tokens;
constexpr size_t DEFAULT_MAX_TOKENS_IN_PATTERN = 300;  // include/fuzzy/suffix_array_index.hh:15

enum class ContrastReduce { MEAN, MAX };

struct EditCosts {
  const float insert_cost;
  const float delete_cost;
  const float replace_cost;
  EditCosts() : insert_cost(1), delete_cost(1), replace_cost(1) {}
  EditCosts(float insert_cost, float delete_cost, float replace_cost)
      : insert_cost(insert_cost), delete_cost(delete_cost), replace_cost(replace_cost) {}
};

// include/fuzzy/sentence.hh:24-48: the real (surface) tokens of a sentence plus the penalty tokens
// (itoks) that sit in the gaps between them; gap i is in front of token i, gap size() is trailing.
class Sentence {
public:
  Sentence() {}
  Sentence(const Tokens& s) : _tokens(s) {}
  operator Tokens() const { return _tokens; }
  const std::string& operator[](size_t i) const { return _tokens[i]; }
  void push_back(const std::string& token) { _tokens.push_back(token); }
  void reserve(size_t n) { _tokens.reserve(n); }
  size_t size() const { return _tokens.size(); }
  bool empty() const { return _tokens.empty(); }
  void set_itok(size_t idx, const std::string& itok) { _itoks[idx] += itok; }
  const std::unordered_map<size_t, std::string>& itoks() const { return _itoks; }

private:
  Tokens _tokens;
  std::unordered_map<size_t, std::string> _itoks;
};

class FuzzyMatch {
public:
  enum penalty_token { pt_none = 0, pt_tag = 1 << 0, pt_pct = 1 << 1, pt_sep = 1 << 2, pt_jnr = 1 << 3, pt_nbr = 1 << 4, pt_cas = 1 << 5 };

  struct Match {
    Match(const unsigned* seq, int length) : length(length), s(seq) {}
    Match() {}
    float score = 0;
    float penalty = 0;
    int max_subseq = 0;
    unsigned s_id = 0;
    std::string id;
    int length = 0;
    const unsigned* s = nullptr;  // borrowed from the index, valid for the life of the FuzzyMatch
  };

  explicit FuzzyMatch(int pt = penalty_token::pt_none, size_t max_tokens_in_pattern = DEFAULT_MAX_TOKENS_IN_PATTERN, int device = 0)
      : _max_tokens(max_tokens_in_pattern), _device(device) {
    if (pt != pt_none) throw std::invalid_argument("penalty tokens need the tokenizer front-end (out of scope)");
    _forms.push_back(std::string(1, '\0'));  // 0 = sentence separator, 1 = unknown (src/vocab_indexer.cc:10-19)
    _forms.push_back("\xEF\xBD\x9Funk\xEF\xBD\xA0");
  }
  ~FuzzyMatch() { fm_index_destroy(_index); }
  FuzzyMatch(const FuzzyMatch&) = delete;
  FuzzyMatch& operator=(const FuzzyMatch&) = delete;

  // add_tm(id, Tokens, sort): src/fuzzy_match.cc:196-203 + src/suffix_array_index.cc:10-30
  bool add_tm(const std::string& id, const Tokens& norm, bool sort = true) { return add_tm(id, Sentence(norm), norm, sort); }

  // add_tm(id, Sentence, Tokens, sort): src/fuzzy_match.cc:205-211; kept iff the real sentence is
  // non-empty and the normalised one is not longer than the cap (src/suffix_array_index.cc:16)
  bool add_tm(const std::string& id, const Sentence& source, const Tokens& norm, bool sort = true) {
    if (!source.empty() && norm.size() <= _max_tokens) {
      if (source.size() != norm.size()) throw std::invalid_argument("real and normalised sentences differ in length");
      for (const auto& w : norm) _tm_tokens.push_back((int32_t)add_word(w));
      append_real(source, _tm_real, _tm_gaps);
      if (!source.itoks().empty() || static_cast<Tokens>(source) != norm) _tm_plain = false;
      _tm_off.push_back((int64_t)_tm_tokens.size());
      _ids.push_back(id);
      _dirty = true;
    }
    if (sort) this->sort();
    return true;
  }

  void sort() {
    if (!_dirty && _index) return;
    fm_index_destroy(_index);
    _index = nullptr;
    check(fm_index_create(_tm_tokens.data(), _tm_off.data(), (int64_t)_tm_off.size() - 1, (int32_t)_forms.size(),
                          (int32_t)_max_tokens, nullptr, 0, 0, _device, &_index));
    _real_uploaded = false;
    _dirty = false;
  }

  size_t max_tokens_in_pattern() const { return _max_tokens; }

  // match(Tokens, ...): include/fuzzy/fuzzy_match.hh:59-69
  bool match(const Tokens& pattern, float fuzzy, unsigned number_of_matches, std::vector<Match>& matches,
             int min_subseq_length = 2, float min_subseq_ratio = 0, float vocab_idf_penalty = 0,
             const EditCosts& edit_costs = EditCosts(), float contrastive_factor = 0,
             ContrastReduce reduce = ContrastReduce::MEAN, int contrast_buffer = -1, bool no_perfect = false) const {
    std::vector<std::vector<Match>> out(1);
    if (!matches.empty() && _tm_plain) {
      // entries already in `matches` count against number_of_matches and take part in the contrastive penalties
      // (src/fuzzy_match.cc:626-679): their sentence ids go with the call (fm_match_batch_prior)
      std::vector<uint32_t> prior;
      for (const Match& m : matches) prior.push_back(m.s_id);
      run_batch({pattern}, nullptr, fuzzy, number_of_matches, out, min_subseq_length, min_subseq_ratio, vocab_idf_penalty, edit_costs,
                contrastive_factor, reduce, contrast_buffer, no_perfect, &prior);
      matches.insert(matches.end(), out[0].begin(), out[0].end());
      return true;
    }
    match_batch({pattern}, fuzzy, number_of_matches, out, min_subseq_length, min_subseq_ratio, vocab_idf_penalty, edit_costs,
                contrastive_factor, reduce, contrast_buffer, no_perfect);
    append_results(matches, out[0], number_of_matches, contrastive_factor);
    return matches.size() > 0;
  }

  // match(Sentence real, Tokens pattern, ..., no_perfect, ...): include/fuzzy/fuzzy_match.hh:70-82 --
  // adds the real-token / case / penalty-token terms of _edit_distance (src/edit_distance.cc:19-62).
  bool match(const Sentence& real, const Tokens& pattern, float fuzzy, unsigned number_of_matches, bool no_perfect,
             std::vector<Match>& matches, int min_subseq_length = 3, float min_subseq_ratio = 0.3f, float vocab_idf_penalty = 0,
             const EditCosts& edit_costs = EditCosts(), float contrastive_factor = 0,
             ContrastReduce reduce = ContrastReduce::MEAN, int contrast_buffer = -1) const {
    std::vector<std::vector<Match>> out(1);
    const std::vector<Sentence> reals(1, real);
    run_batch({pattern}, &reals, fuzzy, number_of_matches, out, min_subseq_length, min_subseq_ratio, vocab_idf_penalty, edit_costs,
              contrastive_factor, reduce, contrast_buffer, no_perfect);
    append_results(matches, out[0], number_of_matches, contrastive_factor);
    return matches.size() > 0;
  }

  // The batched front-end: out[i] receives the matches of patterns[i] (appended).
  void match_batch(const std::vector<Tokens>& patterns, float fuzzy, unsigned number_of_matches,
                   std::vector<std::vector<Match>>& out, int min_subseq_length = 2, float min_subseq_ratio = 0,
                   float vocab_idf_penalty = 0, const EditCosts& edit_costs = EditCosts(), float contrastive_factor = 0,
                   ContrastReduce reduce = ContrastReduce::MEAN, int contrast_buffer = -1, bool no_perfect = false) const {
    run_batch(patterns, nullptr, fuzzy, number_of_matches, out, min_subseq_length, min_subseq_ratio, vocab_idf_penalty, edit_costs,
              contrastive_factor, reduce, contrast_buffer, no_perfect);
  }

  // subsequence(sentence, ...): include/fuzzy/fuzzy_match.hh:96-102, src/fuzzy_match.cc:238-365, behind its tokenizer --
  // the pattern arrives tokenised. Appends at most one Match; Match::id = "<tm id>\t<sub-sequence>" with the tokens
  // of the sub-sequence joined by blanks (what detokenize gives for pt_none).
  bool subsequence(const Tokens& pattern, unsigned number_of_matches, bool no_perfect, std::vector<Match>& matches,
                   int min_subseq_length = 3, float min_subseq_ratio = 0.3f, bool idf_weighting = false) const {
    if (!_index || _dirty) throw std::logic_error("FuzzyMatch::sort() must be called before subsequence()");
    std::vector<int32_t> q_tok;
    for (const auto& w : pattern) {
      auto it = _form2index.find(w);
      q_tok.push_back(it == _form2index.end() ? 1 : (int32_t)it->second);
    }
    const int64_t q_off[2] = {0, (int64_t)q_tok.size()};
    fm_subseq r;
    check(fm_subsequence_batch(_index, q_tok.data(), q_off, 1, (int32_t)number_of_matches, no_perfect, min_subseq_length, min_subseq_ratio,
                               idf_weighting, &r));
    if (r.found != 1) return false;
    const int32_t* toks = nullptr;
    int32_t len = 0;
    check(fm_index_sentence(_index, r.s_id, &toks, &len));
    Match m;  // the reference leaves length / s unset here (best_match is default-constructed, :295)
    m.score = r.score; m.max_subseq = r.length; m.s_id = r.s_id; m.id = _ids[r.s_id] + "\t";
    for (int32_t k = 0; k < r.length; k++) m.id += (k ? " " : "") + pattern[(size_t)(r.position + k)];
    matches.push_back(m);
    return true;
  }

private:
  // The reference APPENDS to `matches` and stops at number_of_matches entries in total, so entries that are
  // already there shorten what a call adds (src/fuzzy_match.cc:670-679). Its contrastive rerank also penalises
  // the candidates against entries that were there before the call (:634-652): match(Tokens) hands them to the
  // library (above); the Sentence overload has no such entry point and refuses instead of answering differently.
  static void append_results(std::vector<Match>& matches, const std::vector<Match>& found, unsigned number_of_matches,
                             float contrastive_factor) {
    if (contrastive_factor > 0 && !matches.empty())
      throw std::logic_error("contrastive match() into a non-empty result vector is not supported: clear it first");
    for (const Match& m : found) {
      if (number_of_matches != 0 && matches.size() >= number_of_matches) break;
      matches.push_back(m);
    }
  }

  // reals == nullptr: the real sentence of every pattern is the pattern itself (Tokens overload)
  void run_batch(const std::vector<Tokens>& patterns, const std::vector<Sentence>* reals, float fuzzy, unsigned number_of_matches,
                 std::vector<std::vector<Match>>& out, int min_subseq_length, float min_subseq_ratio, float vocab_idf_penalty,
                 const EditCosts& edit_costs, float contrastive_factor, ContrastReduce reduce, int contrast_buffer,
                 bool no_perfect, const std::vector<uint32_t>* prior = nullptr) const {
    if (!_index || _dirty) throw std::logic_error("FuzzyMatch::sort() must be called before match()");
    std::vector<int32_t> q_tok, q_real, q_gaps;
    std::vector<int64_t> q_off(1, 0);
    for (size_t k = 0; k < patterns.size(); k++) {
      const Tokens& p = patterns[k];
      for (const auto& w : p) {
        auto it = _form2index.find(w);
        q_tok.push_back(it == _form2index.end() ? 1 : (int32_t)it->second);  // VOCAB_UNK
      }
      if (reals && (*reals)[k].size() != p.size()) throw std::invalid_argument("real and normalised pattern differ in length");
      append_real(reals ? (*reals)[k] : Sentence(p), q_real, q_gaps);
      q_off.push_back((int64_t)q_tok.size());
    }
    fm_params prm;
    prm.fuzzy = fuzzy; prm.number_of_matches = (int32_t)number_of_matches; prm.no_perfect = no_perfect;
    prm.min_subseq_length = min_subseq_length; prm.min_subseq_ratio = min_subseq_ratio; prm.vocab_idf_penalty = vocab_idf_penalty;
    prm.insert_cost = edit_costs.insert_cost; prm.delete_cost = edit_costs.delete_cost; prm.replace_cost = edit_costs.replace_cost;
    prm.contrastive_factor = contrastive_factor; prm.contrast_reduce = reduce == ContrastReduce::MAX; prm.contrast_buffer = contrast_buffer;
    const int64_t n_q = (int64_t)patterns.size();
    int64_t cap = number_of_matches > 0 ? number_of_matches : 64;
    std::vector<fm_match> res;
    std::vector<int32_t> cnt((size_t)n_q);
    for (;;) {
      res.assign((size_t)(n_q * cap), fm_match());
      if (!reals && _tm_plain && prior) {  // one pattern, result vector not empty
        const int64_t prior_off[2] = {0, (int64_t)prior->size()};
        check(fm_match_batch_prior(_index, q_tok.data(), q_off.data(), n_q, &prm, prior->data(), prior_off, cap, res.data(), cnt.data()));
      } else if (!reals && _tm_plain) {  // nothing but normalised tokens anywhere: the plain path is equivalent
        check(fm_match_batch(_index, q_tok.data(), q_off.data(), n_q, &prm, cap, res.data(), cnt.data()));
      } else {
        if (!_real_uploaded) {
          check(fm_index_set_real(_index, _tm_real.data(), _tm_gaps.data(), _tm_off.data(), (int64_t)_tm_off.size() - 1));
          _real_uploaded = true;
        }
        check(fm_match_batch_real(_index, q_tok.data(), q_real.data(), q_gaps.data(), q_off.data(), n_q, &prm, itok_table().data(),
                                  (int32_t)_itoks.size(), cap, res.data(), cnt.data()));
      }
      int64_t mx = 0;
      for (auto c : cnt) mx = c > mx ? c : mx;
      if (mx <= cap) break;
      cap = mx;  // number_of_matches == 0 returns everything
    }
    out.resize((size_t)n_q);
    for (int64_t q = 0; q < n_q; q++)
      for (int32_t k = 0; k < cnt[q]; k++) {
        const fm_match& r = res[(size_t)(q * cap + k)];
        const int32_t* toks = nullptr;
        int32_t len = 0;
        check(fm_index_sentence(_index, r.s_id, &toks, &len));
        Match m(reinterpret_cast<const unsigned*>(toks), r.length);
        m.score = r.score; m.penalty = r.penalty; m.max_subseq = r.max_subseq; m.s_id = r.s_id; m.id = _ids[r.s_id];
        out[(size_t)q].push_back(m);
      }
  }

  // real forms and penalty tokens are interned here; equal strings <=> equal ids
  void append_real(const Sentence& s, std::vector<int32_t>& real, std::vector<int32_t>& gaps) const {
    for (size_t i = 0; i < s.size(); i++) {
      const std::string& t = s[i];
      auto it = _real2id.find(t);
      if (it == _real2id.end()) it = _real2id.emplace(t, (int32_t)_real2id.size()).first;
      const bool case_class = t.empty() || std::strchr("LUMC", t[0]) != nullptr;  // src/edit_distance.cc:55
      real.push_back((it->second << 1) | (case_class ? 1 : 0));
    }
    const size_t g0 = gaps.size();
    gaps.resize(g0 + s.size() + 1, 0);
    for (const auto& kv : s.itoks()) {
      if (kv.first > s.size() || kv.second.empty()) continue;
      auto it = _itok2id.find(kv.second);
      if (it == _itok2id.end()) {
        it = _itok2id.emplace(kv.second, (int32_t)_itoks.size()).first;
        _itoks.push_back(kv.second);
        _itok_dist.clear();  // rebuilt on demand
      }
      gaps[g0 + kv.first] = it->second;
    }
  }
  // pairwise _edit_distance_char (include/fuzzy/edit_distance.hxx:7-35); row / column 0 = lengths
  const std::vector<int32_t>& itok_table() const {
    const size_t k = _itoks.size();
    if (_itok_dist.size() == k * k) return _itok_dist;
    _itok_dist.assign(k * k, 0);
    for (size_t a = 0; a < k; a++)
      for (size_t b = 0; b < k; b++) {
        const std::string &s1 = _itoks[a], &s2 = _itoks[b];
        if (s1.empty() || s2.empty()) { _itok_dist[a * k + b] = (int32_t)(s1.size() + s2.size()); continue; }
        std::vector<int> prev(s2.size() + 1), cur(s2.size() + 1);
        for (size_t j = 0; j <= s2.size(); j++) prev[j] = (int)j;
        for (size_t i = 1; i <= s1.size(); i++) {
          cur[0] = (int)i;
          for (size_t j = 1; j <= s2.size(); j++)
            cur[j] = std::min(std::min(prev[j] + 1, cur[j - 1] + 1), prev[j - 1] + (s1[i - 1] == s2[j - 1] ? 0 : 1));
          prev.swap(cur);
        }
        _itok_dist[a * k + b] = prev[s2.size()];
      }
    return _itok_dist;
  }
  unsigned add_word(const std::string& w) {
    auto it = _form2index.find(w);
    if (it != _form2index.end()) return it->second;
    const unsigned id = (unsigned)_forms.size();
    _form2index.emplace(w, id);
    _forms.push_back(w);
    return id;
  }
  static void check(int rc) {
    if (rc != FM_OK) throw std::runtime_error(std::string("fuzzy_match_b200: ") + fm_last_error());
  }

  size_t _max_tokens;
  int _device;
  bool _dirty = true;
  bool _tm_plain = true;               // every TM sentence so far had real == normalised tokens and no itoks
  mutable bool _real_uploaded = false;  // fm_index_set_real done for the current index
  fm_index* _index = nullptr;
  std::vector<std::string> _forms;
  std::unordered_map<std::string, unsigned> _form2index;
  std::vector<std::string> _ids;
  std::vector<int32_t> _tm_tokens, _tm_real, _tm_gaps;
  std::vector<int64_t> _tm_off = std::vector<int64_t>(1, 0);
  // interning tables grow during const match() calls, like a cache (not thread-safe across concurrent
  // match() calls that introduce new real forms; guard externally or pre-intern)
  mutable std::unordered_map<std::string, int32_t> _real2id;
  mutable std::unordered_map<std::string, int32_t> _itok2id;
  mutable std::vector<std::string> _itoks = std::vector<std::string>(1);  // id 0 = no penalty token
  mutable std::vector<int32_t> _itok_dist;
};

}  // namespace FUZZY_MATCH_B200_NAMESPACE
