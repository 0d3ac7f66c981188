// oracle/ref_driver.cc -- TEST INFRASTRUCTURE, not product code.
//
// C-ABI driver around the UNMODIFIED reference implementation (fuzzy::FuzzyMatch, compiled from
// /root/reference/src/*.cc by oracle/Makefile into oracle/_ref/libfm_ref.so). It feeds the
// reference through its own public API only:
//   FuzzyMatch::add_tm(id, Tokens, sort=false)   include/fuzzy/fuzzy_match.hh:52
//   FuzzyMatch::sort()                           include/fuzzy/fuzzy_match.hh:57
//   FuzzyMatch::match(Tokens, ...)               include/fuzzy/fuzzy_match.hh:59-69
// Token ids are turned into decimal strings so that the reference's own VocabIndexer assigns
// its ids; results (s_id, score, max_subseq, penalty, length) do not depend on that assignment.
// Used (a) to pin the C restatement in oracle/fm_oracle.c and (b) as the "reference" CPU baseline
// of bench.py (worker threads pulling queries from an atomic counter on one shared index, the
// equivalent of FuzzyMatch-cli -N <threads>, cli/src/FuzzyMatch-cli.cc:112-193).
#include <fuzzy/fuzzy_match.hh>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <string>
#include <thread>
#include <vector>

namespace {
struct RefHandle {
  fuzzy::FuzzyMatch fm;
  explicit RefHandle(int max_tok) : fm(fuzzy::FuzzyMatch::pt_none, (size_t)max_tok) {}
};
fuzzy::Tokens to_tokens(const int32_t* t, int64_t n) {
  fuzzy::Tokens out;
  out.reserve((size_t)n);
  for (int64_t i = 0; i < n; i++) out.push_back(std::to_string(t[i]));
  return out;
}
}  // namespace

extern "C" {

struct fmref_params {
  float fuzzy;
  int32_t number_of_matches;
  int32_t min_subseq_length;
  float min_subseq_ratio;
  float vocab_idf_penalty;
  float insert_cost, delete_cost, replace_cost;
  float contrastive_factor;
  int32_t contrast_reduce;  // 0 = MEAN, 1 = MAX
  int32_t contrast_buffer;
};

struct fmref_match {
  uint32_t s_id;
  float score;
  float penalty;
  int32_t max_subseq;
  int32_t length;
};

void* fmref_create(int max_tokens_in_pattern) { return new RefHandle(max_tokens_in_pattern); }
void fmref_destroy(void* h) { delete static_cast<RefHandle*>(h); }

void fmref_add_tm(void* h, const int32_t* tokens, const int64_t* off, int64_t n_sent) {
  auto* r = static_cast<RefHandle*>(h);
  for (int64_t s = 0; s < n_sent; s++)
    r->fm.add_tm(std::to_string(s), to_tokens(tokens + off[s], off[s + 1] - off[s]), /*sort=*/false);
}

void fmref_sort(void* h) { static_cast<RefHandle*>(h)->fm.sort(); }

// Runs match(Tokens) for every query. out is [n_q * cap]; out_count[q] = number of matches the
// reference returned (may exceed cap; only the first cap are stored). Returns wall seconds spent
// inside the matching loop (string conversion of the queries happens before the clock starts).
double fmref_match_batch(void* h, const int32_t* q_tokens, const int64_t* q_off, int64_t n_q,
                         const fmref_params* p, int nthreads, int64_t cap, fmref_match* out,
                         int32_t* out_count) {
  auto* r = static_cast<RefHandle*>(h);
  std::vector<fuzzy::Tokens> queries((size_t)n_q);
  for (int64_t q = 0; q < n_q; q++) queries[q] = to_tokens(q_tokens + q_off[q], q_off[q + 1] - q_off[q]);
  const fuzzy::EditCosts costs(p->insert_cost, p->delete_cost, p->replace_cost);
  const auto reduce = p->contrast_reduce ? fuzzy::ContrastReduce::MAX : fuzzy::ContrastReduce::MEAN;
  std::atomic<int64_t> next(0);
  auto work = [&]() {
    std::vector<fuzzy::FuzzyMatch::Match> matches;
    for (;;) {
      const int64_t q = next.fetch_add(1);
      if (q >= n_q) break;
      matches.clear();
      r->fm.match(queries[q], p->fuzzy, (unsigned)p->number_of_matches, matches, p->min_subseq_length,
                  p->min_subseq_ratio, p->vocab_idf_penalty, costs, p->contrastive_factor, reduce,
                  p->contrast_buffer);
      if (out_count) out_count[q] = (int32_t)matches.size();
      if (out)
        for (size_t k = 0; k < matches.size() && (int64_t)k < cap; k++) {
          fmref_match& o = out[q * cap + (int64_t)k];
          o.s_id = matches[k].s_id;
          o.score = matches[k].score;
          // Match::penalty is never initialised by the reference unless contrastive rerank runs
          // (fuzzy_match.hh:34-38); report 0 there.
          o.penalty = p->contrastive_factor > 0 ? matches[k].penalty : 0.f;
          o.max_subseq = matches[k].max_subseq;
          o.length = matches[k].length;
        }
    }
  };
  const auto t0 = std::chrono::steady_clock::now();
  if (nthreads <= 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; t++) pool.emplace_back(work);
    for (auto& t : pool) t.join();
  }
  const auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// ---- Sentence API: real tokens and penalty tokens (itoks) given explicitly
// (FuzzyMatch::add_tm(id, Sentence, Tokens) fuzzy_match.hh:53, match(Sentence, Tokens, ...) :70-82).
// real[k] = (real form id << 1) | case_class: the real token string is "L<id>" when case_class is set
// (a case-feature token, first character in "LUMC", src/edit_distance.cc:55) and "r<id>" otherwise, so
// string equality == (id, class) equality. gaps: for a sentence of n tokens, n+1 itok ids (0 = none)
// at positions off[s] + s ... ; itok strings come from (itok_blob, itok_off).
static fuzzy::Sentence make_sentence(const int32_t* real, const int32_t* gaps, int64_t n, const char* blob,
                                     const int32_t* itok_off) {
  fuzzy::Sentence s;
  for (int64_t i = 0; i < n; i++) s.push_back(std::string(real[i] & 1 ? "L" : "r") + std::to_string(real[i] >> 1));
  for (int64_t i = 0; i <= n; i++)
    if (gaps[i]) s.set_itok((size_t)i, std::string(blob + itok_off[gaps[i]], blob + itok_off[gaps[i] + 1]));
  return s;
}

void fmref_add_tm_real(void* h, const int32_t* tokens, const int32_t* real, const int32_t* gaps, const int64_t* off,
                       int64_t n_sent, const char* itok_blob, const int32_t* itok_off) {
  auto* r = static_cast<RefHandle*>(h);
  for (int64_t s = 0; s < n_sent; s++) {
    const int64_t n = off[s + 1] - off[s];
    r->fm.add_tm(std::to_string(s), make_sentence(real + off[s], gaps + off[s] + s, n, itok_blob, itok_off),
                 to_tokens(tokens + off[s], n), /*sort=*/false);
  }
}

double fmref_match_batch_real(void* h, const int32_t* q_tokens, const int32_t* q_real, const int32_t* q_gaps,
                              const int64_t* q_off, int64_t n_q, const fmref_params* p, int no_perfect, int nthreads,
                              int64_t cap, fmref_match* out, int32_t* out_count, const char* itok_blob,
                              const int32_t* itok_off) {
  auto* r = static_cast<RefHandle*>(h);
  std::vector<fuzzy::Tokens> queries((size_t)n_q);
  std::vector<fuzzy::Sentence> reals((size_t)n_q);
  for (int64_t q = 0; q < n_q; q++) {
    const int64_t n = q_off[q + 1] - q_off[q];
    queries[q] = to_tokens(q_tokens + q_off[q], n);
    reals[q] = make_sentence(q_real + q_off[q], q_gaps + q_off[q] + q, n, itok_blob, itok_off);
  }
  const fuzzy::EditCosts costs(p->insert_cost, p->delete_cost, p->replace_cost);
  const auto reduce = p->contrast_reduce ? fuzzy::ContrastReduce::MAX : fuzzy::ContrastReduce::MEAN;
  std::atomic<int64_t> next(0);
  auto work = [&]() {
    std::vector<fuzzy::FuzzyMatch::Match> matches;
    for (;;) {
      const int64_t q = next.fetch_add(1);
      if (q >= n_q) break;
      matches.clear();
      r->fm.match(reals[q], queries[q], p->fuzzy, (unsigned)p->number_of_matches, no_perfect != 0, matches,
                  p->min_subseq_length, p->min_subseq_ratio, p->vocab_idf_penalty, costs, p->contrastive_factor, reduce,
                  p->contrast_buffer);
      if (out_count) out_count[q] = (int32_t)matches.size();
      if (out)
        for (size_t k = 0; k < matches.size() && (int64_t)k < cap; k++) {
          fmref_match& o = out[q * cap + (int64_t)k];
          o.s_id = matches[k].s_id;
          o.score = matches[k].score;
          o.penalty = p->contrastive_factor > 0 ? matches[k].penalty : 0.f;
          o.max_subseq = matches[k].max_subseq;
          o.length = matches[k].length;
        }
    }
  };
  const auto t0 = std::chrono::steady_clock::now();
  if (nthreads <= 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; t++) pool.emplace_back(work);
    for (auto& t : pool) t.join();
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// FuzzyMatch::subsequence(string, ...) (include/fuzzy/fuzzy_match.hh:96-102) for a batch: the pattern is handed over as
// the decimal token strings joined by blanks, which the stand-in tokenizer splits again and the reference's own
// _tokenize_and_normalize passes through unchanged (pt_none). text receives what the reference appends to Match::id
// behind the tab (the detokenised sub-sequence), NUL-terminated, text_stride bytes per query.
struct fmref_subseq { uint32_t s_id; float score; int32_t max_subseq; int32_t found; };
void fmref_subsequence_batch(void* h, const int32_t* q_tokens, const int64_t* q_off, int64_t n_q, int32_t number_of_matches,
                             int32_t no_perfect, int32_t min_subseq_length, float min_subseq_ratio, int32_t idf_weighting,
                             fmref_subseq* out, char* text, int64_t text_stride) {
  fuzzy::FuzzyMatch& fm = static_cast<RefHandle*>(h)->fm;
  for (int64_t q = 0; q < n_q; q++) {
    std::string sentence;
    for (int64_t i = q_off[q]; i < q_off[q + 1]; i++) { if (i > q_off[q]) sentence += " "; sentence += std::to_string(q_tokens[i]); }
    std::vector<fuzzy::FuzzyMatch::Match> matches;
    const bool ok = fm.subsequence(sentence, (unsigned)number_of_matches, no_perfect != 0, matches, min_subseq_length, min_subseq_ratio,
                                   idf_weighting != 0);
    out[q] = fmref_subseq{0, 0.f, 0, 0};
    text[q * text_stride] = 0;
    if (ok && !matches.empty()) {
      const auto& m = matches.back();
      out[q] = fmref_subseq{m.s_id, m.score, m.max_subseq, 1};
      const size_t tab = m.id.find('\t');
      const std::string sub = tab == std::string::npos ? std::string() : m.id.substr(tab + 1);
      snprintf(text + q * text_stride, (size_t)text_stride, "%s", sub.c_str());
    }
  }
}

// match(Tokens) into a result vector that already holds matches: for every query the vector is first filled by
// match(first pattern, p1) and then handed, as it is, to match(second pattern, p2) -- the reference appends, counts the
// earlier entries against number_of_matches and penalises contrastive candidates against them
// (src/fuzzy_match.cc:626-679). prior_* receive the entries of the first call, out_* what the second call appended.
void fmref_match_batch_twice(void* h, const int32_t* q1_tokens, const int64_t* q1_off, const int32_t* q2_tokens, const int64_t* q2_off,
                             int64_t n_q, const fmref_params* p1, const fmref_params* p2, int64_t cap, fmref_match* prior,
                             int32_t* prior_count, fmref_match* out, int32_t* out_count) {
  auto* r = static_cast<RefHandle*>(h);
  auto call = [&](const fuzzy::Tokens& pat, const fmref_params* p, std::vector<fuzzy::FuzzyMatch::Match>& matches) {
    const fuzzy::EditCosts costs(p->insert_cost, p->delete_cost, p->replace_cost);
    r->fm.match(pat, p->fuzzy, (unsigned)p->number_of_matches, matches, p->min_subseq_length, p->min_subseq_ratio, p->vocab_idf_penalty,
                costs, p->contrastive_factor, p->contrast_reduce ? fuzzy::ContrastReduce::MAX : fuzzy::ContrastReduce::MEAN,
                p->contrast_buffer);
  };
  auto store = [&](const fuzzy::FuzzyMatch::Match& m, fmref_match& o, bool penalty_valid) {
    o.s_id = m.s_id; o.score = m.score; o.penalty = penalty_valid ? m.penalty : 0.f; o.max_subseq = m.max_subseq; o.length = m.length;
  };
  for (int64_t q = 0; q < n_q; q++) {
    std::vector<fuzzy::FuzzyMatch::Match> matches;
    call(to_tokens(q1_tokens + q1_off[q], q1_off[q + 1] - q1_off[q]), p1, matches);
    const size_t n1 = matches.size();
    prior_count[q] = (int32_t)n1;
    for (size_t k = 0; k < n1 && (int64_t)k < cap; k++) store(matches[k], prior[q * cap + (int64_t)k], false);
    call(to_tokens(q2_tokens + q2_off[q], q2_off[q + 1] - q2_off[q]), p2, matches);
    out_count[q] = (int32_t)(matches.size() - n1);
    // Match::penalty of an appended entry is computed iff the rerank ran with a non-empty vector (:634-662)
    for (size_t k = n1; k < matches.size() && (int64_t)(k - n1) < cap; k++)
      store(matches[k], out[q * cap + (int64_t)(k - n1)], p2->contrastive_factor > 0 && k > 0);
  }
}

int64_t fmref_max_tokens_in_pattern(void* h) { return (int64_t) static_cast<RefHandle*>(h)->fm.max_tokens_in_pattern(); }

}  // extern "C"
