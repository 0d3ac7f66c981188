/*
 * oracle/fm_oracle.c -- CPU ORACLE. TEST INFRASTRUCTURE ONLY, never on the product path.
 *
 * Plain-C restatement of the SYSTRAN/fuzzy-match hot path FuzzyMatch::match()
 * (reference src/fuzzy_match.cc:435-681) on pre-tokenised int32 word ids, i.e. what the
 * reference computes through match(const Tokens&, ...) (src/fuzzy_match.cc:415-432) where the
 * "real" sentence equals the normalised one and there are no penalty tokens (itoks).
 *
 * Parity pin: tests/test_oracle.py checks this file against the reference's own sources
 * compiled unmodified (oracle/_ref/libfm_ref.so, see oracle/Makefile) on the reference's
 * tokenizer-free known-answer tests (test/test.cc:223-262, 337-632), on the order-dependence
 * vectors Q1/Q2 of SURVEY.md section 3.1, and on seeded random TMs; the resulting vectors are
 * committed under tests/golden/ so the pin also holds where /root/reference is absent.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library (as the checker or the timed CPU baseline).
 *
 * Conventions: word ids >= 2 are vocabulary words, 0 is the sentence separator and 1 is
 * "unknown" (reference src/vocab_indexer.cc:10-11). Query ids that never occur in the TM behave
 * as unknown. Every float expression below is evaluated as separate IEEE single operations in
 * the reference's source order (this file must be built with -ffp-contract=off and no -march).
 */
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ public structs */

typedef struct fmo_params {
  float fuzzy;
  int32_t number_of_matches;
  int32_t no_perfect;
  int32_t min_subseq_length;
  float min_subseq_ratio;
  float vocab_idf_penalty;
  float insert_cost, delete_cost, replace_cost;
  float contrastive_factor;
  int32_t contrast_reduce; /* 0 = MEAN, 1 = MAX */
  int32_t contrast_buffer;
} fmo_params;

typedef struct fmo_match {
  uint32_t s_id;
  float score;
  float penalty;
  int32_t max_subseq;
  int32_t length;
  float cost;
} fmo_match;

/* implementation-independent work counters */
typedef struct fmo_counters {
  int64_t queries;
  int64_t equal_range_calls; /* calls with length >= 2 */
  int64_t probes;            /* sum over those calls of 2*ceil(log2(R+1)) */
  int64_t elements_walked;   /* SA elements in registered ranges with match_len >= ml */
  int64_t candidates;        /* deduplicated candidates after the length filter */
  int64_t candidate_tokens;  /* sum of their sentence lengths */
  int64_t dp_pairs;          /* pairs that reach the edit distance */
  int64_t dp_tokens;         /* sum of their sentence lengths */
  int64_t dp_cells;          /* sum of s*p over those pairs */
  int64_t pattern_tokens;
  int64_t matches_out;
} fmo_counters;

/* per-candidate record for kernel-level tests, in the reference's processing order */
typedef struct fmo_candidate {
  uint32_t s_id;
  int32_t longest_match;
  int32_t s_length;
  int32_t cover;
  int32_t rejected; /* theoretical_rejection_cover said no */
  float cost_full;  /* DP without bound (C); only when !rejected */
  float rowmin_max; /* K = max over rows of the row minimum */
  int32_t accepted; /* pushed into the result heap */
} fmo_candidate;

typedef struct fmo_index {
  int64_t n_sent;    /* kept sentences */
  int64_t n_tok;     /* = number of suffixes */
  int32_t vocab_size;
  int32_t max_tokens;
  int32_t* buf;      /* kept sentences, each followed by one 0 separator */
  int64_t* sent_pos; /* [n_sent+1] offset of each sentence in buf */
  int64_t* sa_pos;   /* [n_tok] suffix start offsets into buf, sorted */
  uint32_t* sa_sid;  /* [n_tok] sentence id of each suffix */
  uint16_t* sa_len;  /* [n_tok] sentence length of each suffix */
  int64_t* qva;      /* [vocab_size+1] first suffix starting with each word id */
  uint32_t* sfreq;   /* [vocab_size] number of sentences containing the word */
  int64_t n_sent_global;
  int64_t* kept_src; /* [n_sent] index of each kept sentence in the caller's CSR */
  /* optional real tokens / penalty tokens (Sentence API), same layout as buf */
  int32_t* real;     /* (real form id << 1) | case class, per token */
  int32_t* gap;      /* itok id of the gap before each token; the separator slot holds the trailing gap */
  char* itok_blob; int32_t* itok_off; int32_t n_itok; /* itok strings, id 0 = none (empty) */
} fmo_index;

/* ------------------------------------------------------------------ index build */

/* suffix order: token-wise, a shorter suffix that is a prefix sorts first, ties by sentence id
 * (reference src/suffix_array.cc:214-251). The 0 separator is smaller than any word id, so a
 * plain compare that stops after a separator implements "shorter first". */
static const int32_t* g_cmp_buf;
static int g_cmp_depth;
static int cmp_suffix(const void* a, const void* b) {
  const int64_t pa = *(const int64_t*)a, pb = *(const int64_t*)b;
  const int32_t* x = g_cmp_buf + pa + g_cmp_depth;
  const int32_t* y = g_cmp_buf + pb + g_cmp_depth;
  for (;; x++, y++) {
    if (*x != *y) return *x < *y ? -1 : 1;
    if (*x == 0) break;
  }
  return pa < pb ? -1 : (pa > pb ? 1 : 0); /* position order == sentence-id order here */
}

typedef struct { uint64_t key; int64_t pos; } keypos;

static void radix_sort_keypos(keypos* a, keypos* tmp, int64_t n) {
  for (int pass = 0; pass < 4; pass++) {
    const int shift = pass * 16;
    int64_t* cnt = (int64_t*)calloc(65537, sizeof(int64_t));
    for (int64_t i = 0; i < n; i++) cnt[((a[i].key >> shift) & 0xffff) + 1]++;
    for (int i = 0; i < 65536; i++) cnt[i + 1] += cnt[i];
    for (int64_t i = 0; i < n; i++) tmp[cnt[(a[i].key >> shift) & 0xffff]++] = a[i];
    memcpy(a, tmp, (size_t)n * sizeof(keypos));
    free(cnt);
  }
}

/* Builds the index the way add_tm + sort do (reference src/suffix_array_index.cc:10-30,
 * src/suffix_array.cc:9-27,58-102, src/vocab_indexer.cc:73-90): sentences that are empty or longer
 * than max_tokens are dropped and the kept ones are numbered consecutively. sfreq_global/
 * n_sent_global (optional) override the IDF statistics for a sentence-id shard of a larger TM. */
fmo_index* fmo_index_create(const int32_t* tokens, const int64_t* off, int64_t n_in, int32_t vocab_size,
                            int32_t max_tokens, const uint32_t* sfreq_global, int64_t n_sent_global) {
  fmo_index* ix = (fmo_index*)calloc(1, sizeof(fmo_index));
  ix->vocab_size = vocab_size;
  ix->max_tokens = max_tokens;
  int64_t n_keep = 0, n_tok = 0;
  for (int64_t s = 0; s < n_in; s++) {
    const int64_t len = off[s + 1] - off[s];
    if (len > 0 && len <= max_tokens) { n_keep++; n_tok += len; }
  }
  ix->n_sent = n_keep;
  ix->n_tok = n_tok;
  ix->buf = (int32_t*)malloc((size_t)(n_tok + n_keep + 1) * sizeof(int32_t));
  ix->sent_pos = (int64_t*)malloc((size_t)(n_keep + 1) * sizeof(int64_t));
  ix->kept_src = (int64_t*)malloc((size_t)(n_keep + 1) * sizeof(int64_t));
  ix->sfreq = (uint32_t*)calloc((size_t)vocab_size, sizeof(uint32_t));
  int64_t w = 0, k = 0;
  for (int64_t s = 0; s < n_in; s++) {
    const int64_t len = off[s + 1] - off[s];
    if (!(len > 0 && len <= max_tokens)) continue;
    ix->sent_pos[k] = w;
    ix->kept_src[k] = s;
    for (int64_t i = 0; i < len; i++) {
      const int32_t t = tokens[off[s] + i];
      if (t < 2 || t >= vocab_size) { /* invalid TM token: refuse */
        free(ix->buf); free(ix->sent_pos); free(ix->kept_src); free(ix->sfreq); free(ix);
        return NULL;
      }
      ix->buf[w + i] = t;
      int seen = 0; /* count each word once per sentence */
      for (int64_t j = 0; j < i; j++) if (ix->buf[w + j] == t) { seen = 1; break; }
      if (!seen) ix->sfreq[t]++;
    }
    w += len;
    ix->buf[w++] = 0;
    k++;
  }
  ix->sent_pos[n_keep] = w;
  ix->buf[w] = 0;
  if (sfreq_global) memcpy(ix->sfreq, sfreq_global, (size_t)vocab_size * sizeof(uint32_t));
  ix->n_sent_global = n_sent_global > 0 ? n_sent_global : n_keep;

  /* sort suffixes: radix on (tok0, tok1), comparator from depth 2 inside equal-key runs */
  keypos* kp = (keypos*)malloc((size_t)(n_tok + 1) * sizeof(keypos));
  keypos* tmp = (keypos*)malloc((size_t)(n_tok + 1) * sizeof(keypos));
  int64_t m = 0;
  for (int64_t s = 0; s < n_keep; s++)
    for (int64_t p = ix->sent_pos[s]; ix->buf[p] != 0; p++) {
      kp[m].key = ((uint64_t)(uint32_t)ix->buf[p] << 32) | (uint32_t)ix->buf[p + 1];
      kp[m].pos = p;
      m++;
    }
  radix_sort_keypos(kp, tmp, n_tok); /* stable: equal keys stay in position order */
  free(tmp);
  ix->sa_pos = (int64_t*)malloc((size_t)(n_tok + 1) * sizeof(int64_t));
  for (int64_t i = 0; i < n_tok; i++) ix->sa_pos[i] = kp[i].pos;
  g_cmp_buf = ix->buf;
  g_cmp_depth = 2;
  for (int64_t i = 0; i < n_tok;) {
    int64_t j = i + 1;
    while (j < n_tok && kp[j].key == kp[i].key) j++;
    if (j - i > 1 && (kp[i].key & 0xffffffffu) != 0) qsort(ix->sa_pos + i, (size_t)(j - i), sizeof(int64_t), cmp_suffix);
    i = j;
  }
  free(kp);

  /* per-suffix sentence id / length (reference src/suffix_array.cc:253-261) */
  uint32_t* sid_of_pos = (uint32_t*)malloc((size_t)(w + 1) * sizeof(uint32_t));
  for (int64_t s = 0; s < n_keep; s++)
    for (int64_t p = ix->sent_pos[s]; p < ix->sent_pos[s + 1]; p++) sid_of_pos[p] = (uint32_t)s;
  ix->sa_sid = (uint32_t*)malloc((size_t)(n_tok + 1) * sizeof(uint32_t));
  ix->sa_len = (uint16_t*)malloc((size_t)(n_tok + 1) * sizeof(uint16_t));
  ix->qva = (int64_t*)malloc((size_t)(vocab_size + 1) * sizeof(int64_t));
  for (int64_t i = 0; i < n_tok; i++) {
    const uint32_t s = sid_of_pos[ix->sa_pos[i]];
    ix->sa_sid[i] = s;
    ix->sa_len[i] = (uint16_t)(ix->sent_pos[s + 1] - ix->sent_pos[s] - 1);
  }
  free(sid_of_pos);
  /* first-word bucket table (reference _quickVocabAccess, src/suffix_array.cc:82-98) */
  int64_t i = 0;
  for (int32_t wid = 0; wid <= vocab_size; wid++) {
    while (i < n_tok && ix->buf[ix->sa_pos[i]] < wid) i++;
    ix->qva[wid] = i;
  }
  return ix;
}

void fmo_index_destroy(fmo_index* ix) {
  if (!ix) return;
  free(ix->buf); free(ix->sent_pos); free(ix->sa_pos); free(ix->sa_sid); free(ix->sa_len);
  free(ix->qva); free(ix->sfreq); free(ix->kept_src); free(ix->real); free(ix->gap); free(ix->itok_blob); free(ix->itok_off); free(ix);
}

int64_t fmo_index_num_sentences(const fmo_index* ix) { return ix->n_sent; }
int64_t fmo_index_num_suffixes(const fmo_index* ix) { return ix->n_tok; }
const uint32_t* fmo_index_sfreq(const fmo_index* ix) { return ix->sfreq; }
const int64_t* fmo_index_kept(const fmo_index* ix) { return ix->kept_src; }

/* ------------------------------------------------------------------ suffix-array search */

/* start_by(): compare suffix against the n-gram, "equal" if the suffix starts with it
 * (reference src/suffix_array.cc:214-233, 263-273) */
static int start_by(const fmo_index* ix, int64_t suffix, const int32_t* ngram, int64_t length) {
  const int32_t* s = ix->buf + ix->sa_pos[suffix];
  for (int64_t i = 0; i < length; i++) {
    if (s[i] == 0) return -1; /* suffix shorter than the n-gram */
    if (s[i] < ngram[i]) return -1;
    if (s[i] > ngram[i]) return 1;
  }
  return 0;
}

static int64_t ceil_log2(int64_t x) { /* ceil(log2(x)) for x >= 1 */
  int64_t r = 0;
  while (((int64_t)1 << r) < x) r++;
  return r;
}

/* equal_range(): the canonical [lower, upper) of suffixes starting with the n-gram, searched
 * inside [min,max) (or inside the first-word bucket when max == 0). The reference's narrowing
 * loop (src/suffix_array.cc:105-212) returns exactly this range (its post-conditions are asserted
 * at :206-210), or an empty one. */
static void equal_range(const fmo_index* ix, const int32_t* ngram, int64_t length, int64_t min, int64_t max,
                        int64_t* lo_out, int64_t* hi_out, fmo_counters* ct) {
  *lo_out = *hi_out = 0;
  if (length == 0) return;
  if (max == 0) {
    if (ngram[0] < 0 || ngram[0] >= ix->vocab_size) return;
    min = ix->qva[ngram[0]];
    max = ix->qva[ngram[0] + 1];
    if (length == 1) { *lo_out = min; *hi_out = max; return; }
  }
  if (ct && length >= 2) { ct->equal_range_calls++; ct->probes += 2 * ceil_log2(max - min + 1); }
  int64_t lo = min, hi = max;
  while (lo < hi) { /* first suffix that is not < ngram */
    const int64_t mid = lo + (hi - lo) / 2;
    if (start_by(ix, mid, ngram, length) < 0) lo = mid + 1; else hi = mid;
  }
  const int64_t lower = lo;
  hi = max;
  while (lo < hi) { /* first suffix that is > ngram */
    const int64_t mid = lo + (hi - lo) / 2;
    if (start_by(ix, mid, ngram, length) <= 0) lo = mid + 1; else hi = mid;
  }
  *lo_out = lower;
  *hi_out = lo;
}

/* ------------------------------------------------------------------ costs and rejection bounds */

/* Costs::get_normalizer (reference include/fuzzy/costs.hh:33-47) */
static float get_normalizer(int64_t p, int64_t s, const fmo_params* pr) {
  const float ins = pr->insert_cost, del = pr->delete_cost, rep = pr->replace_cost;
  if (ins == 0.f && del == 0.f && rep == 0.f) return 1.f;
  if (ins + del <= rep) return ins * (float)p + del * (float)s;
  if (p <= s) return (rep - del) * (float)p + del * (float)s;
  return (rep - ins) * (float)s + ins * (float)p;
}

/* NGramMatches::theoretical_rejection (reference src/ngram_matches.cc:32-39) */
static int theoretical_rejection(int64_t p, int64_t s, const fmo_params* pr) {
  const float size_difference = fabsf((float)p - (float)s);
  const float remaining_cost = (p >= s) ? pr->insert_cost : pr->delete_cost;
  const float bound = 1.f - remaining_cost * size_difference / get_normalizer(p, s, pr);
  return (double)bound + 0.000005 < (double)pr->fuzzy;
}

/* NGramMatches::theoretical_rejection_cover (reference src/ngram_matches.cc:42-59) */
static int theoretical_rejection_cover(int64_t p, int64_t s, int64_t cover, const fmo_params* pr) {
  const float ins = pr->insert_cost, del = pr->delete_cost, rep = pr->replace_cost;
  float bound;
  if (ins + del < rep) {
    bound = 1.f - (ins * ((float)s - (float)cover) + del * ((float)p - (float)cover)) / get_normalizer(p, s, pr);
  } else {
    const float cost_remaining = (p > s) ? ins : del;
    const float min_length = (p > s) ? (float)s : (float)p;
    const float max_length = (p > s) ? (float)p : (float)s;
    bound = 1.f - (rep * (min_length - (float)cover) + cost_remaining * (max_length - min_length)) /
                      get_normalizer(p, s, pr);
  }
  return (double)bound + 0.000005 < (double)pr->fuzzy;
}

/* PatternCoverage::count_covered_words (reference src/pattern_coverage.cc:15-28): sum over the
 * distinct pattern words present in the sentence of their multiplicity in the pattern == number of
 * pattern positions whose word occurs in the sentence. */
static int64_t count_covered_words(const int32_t* pat, int64_t p, const int32_t* sent, int64_t s) {
  int64_t covered = 0;
  for (int64_t j = 0; j < p; j++)
    for (int64_t i = 0; i < s; i++)
      if (sent[i] == pat[j]) { covered++; break; }
  return covered;
}

/* ------------------------------------------------------------------ edit distances */

/* _edit_distance, full variant without real-token / penalty-token terms
 * (reference src/edit_distance.cc:5-77; rows = TM sentence s1, columns = pattern s2).
 * Returns the value the reference returns (row minimum on early exit). If rowmin_max is non-NULL
 * it receives K = max over the rows actually evaluated of the row minimum. */
static float edit_distance_full(const int32_t* s1, int n1, const int32_t* s2, int n2, const float* idf_penalty,
                                float idf_weight, const fmo_params* pr, float diff_word, float max_fuzzyness,
                                float* rowmin_max, float* scratch) {
  float* prev = scratch;
  float* cur = scratch + (n2 + 1);
  prev[0] = 0.f;
  for (int j = 1; j < n2 + 1; j++) {
    prev[j] = prev[j - 1] + diff_word * pr->insert_cost;
    if (idf_weight) prev[j] += idf_penalty[j - 1] * idf_weight;
  }
  float col0 = 0.f, kmax = -FLT_MAX;
  for (int i = 1; i < n1 + 1; i++) {
    col0 = col0 + diff_word * pr->delete_cost;
    cur[0] = col0;
    float min = FLT_MAX;
    for (int j = 1; j < n2 + 1; j++) {
      float diff = 0.f, penalty_j1 = 0.f;
      if (idf_weight) penalty_j1 = idf_penalty[j - 1] * idf_weight;
      if (s1[i - 1] != s2[j - 1]) diff = pr->replace_cost * diff_word + penalty_j1;
      const float a = prev[j] + pr->delete_cost * diff_word;
      const float b = cur[j - 1] + pr->insert_cost * diff_word + penalty_j1;
      const float c = prev[j - 1] + diff;
      float d = a < b ? a : b;
      d = c < d ? c : d;
      cur[j] = d;
      if (d < min) min = d;
    }
    if (min > kmax) kmax = min;
    if (min > max_fuzzyness) { if (rowmin_max) *rowmin_max = kmax; return min; }
    float* t = prev; prev = cur; cur = t;
  }
  if (rowmin_max) *rowmin_max = kmax;
  return prev[n2];
}

/* _edit_distance, plain variant used by the contrastive rerank (reference src/edit_distance.cc:79-122) */
static float edit_distance_plain(const int32_t* s1, int n1, const int32_t* s2, int n2, float ins, float del, float rep,
                                 float diff_word, float* scratch) {
  float* prev = scratch;
  float* cur = scratch + (n2 + 1);
  prev[0] = 0.f;
  for (int j = 1; j < n2 + 1; j++) prev[j] = prev[j - 1] + diff_word * ins;
  float col0 = 0.f;
  for (int i = 1; i < n1 + 1; i++) {
    col0 = col0 + diff_word * del;
    cur[0] = col0;
    for (int j = 1; j < n2 + 1; j++) {
      float diff = 0.f;
      if (s1[i - 1] != s2[j - 1]) diff = rep * diff_word;
      const float a = prev[j] + del * diff_word;
      const float b = cur[j - 1] + ins * diff_word;
      const float c = prev[j - 1] + diff;
      float d = a < b ? a : b;
      d = c < d ? c : d;
      cur[j] = d;
    }
    float* t = prev; prev = cur; cur = t;
  }
  return prev[n2];
}

/* _edit_distance_char (reference include/fuzzy/edit_distance.hxx:7-35) on two penalty-token strings */
static int edit_distance_char(const char* s1, int n1, const char* s2, int n2) {
  if (n1 == 0) return n2;
  if (n2 == 0) return n1;
  int prev[64], cur[64];
  if (n2 > 62) n2 = 62; /* itoks are a few characters */
  for (int j = 0; j <= n2; j++) prev[j] = j;
  for (int i = 1; i <= n1; i++) {
    cur[0] = i;
    for (int j = 1; j <= n2; j++) {
      int d = prev[j] + 1;
      if (cur[j - 1] + 1 < d) d = cur[j - 1] + 1;
      const int c = prev[j - 1] + (s1[i - 1] == s2[j - 1] ? 0 : 1);
      if (c < d) d = c;
      cur[j] = d;
    }
    memcpy(prev, cur, sizeof(int) * (size_t)(n2 + 1));
  }
  return prev[n2];
}
static int itok_len(const fmo_index* ix, int id) { return ix->itok_off[id + 1] - ix->itok_off[id]; }
static int cost_tag(const fmo_index* ix, int a, int b) {
  return edit_distance_char(ix->itok_blob + ix->itok_off[a], itok_len(ix, a), ix->itok_blob + ix->itok_off[b], itok_len(ix, b));
}

/* _edit_distance, full variant WITH real-token and penalty-token terms (reference
 * src/edit_distance.cc:5-77). g1 / g2: itok ids of the n1+1 / n2+1 gaps; r1 / r2: real tokens. */
static float edit_distance_real(const fmo_index* ix, const int32_t* s1, const int32_t* r1, const int32_t* g1, int n1,
                                const int32_t* s2, const int32_t* r2, const int32_t* g2, int n2, const float* idf_penalty,
                                float idf_weight, const fmo_params* pr, float diff_word, float max_fuzzyness,
                                float* rowmin_max, float* scratch) {
  float* prev = scratch;
  float* cur = scratch + (n2 + 1);
  prev[0] = (float)cost_tag(ix, g1[n1], g2[n2]); /* :25 trailing penalty tokens */
  for (int j = 1; j < n2 + 1; j++) {
    prev[j] = prev[j - 1] + diff_word * pr->insert_cost + itok_len(ix, g2[j]); /* :35 */
    if (idf_weight) prev[j] += idf_penalty[j - 1] * idf_weight;
  }
  float col0 = prev[0], kmax = -FLT_MAX;
  for (int i = 1; i < n1 + 1; i++) {
    col0 = col0 + diff_word * pr->delete_cost + itok_len(ix, g1[i]); /* :30 */
    cur[0] = col0;
    float min = FLT_MAX;
    for (int j = 1; j < n2 + 1; j++) {
      float diff = 0.f, penalty_j1 = 0.f;
      if (idf_weight) penalty_j1 = idf_penalty[j - 1] * idf_weight;
      if (s1[i - 1] != s2[j - 1]) diff = pr->replace_cost * diff_word + penalty_j1;
      else if (r1[i - 1] != r2[j - 1]) diff = (r1[i - 1] & 1) ? pr->replace_cost * 1.0f : pr->replace_cost * 2.0f; /* :53-59 */
      const float a = prev[j] + pr->delete_cost * diff_word + cost_tag(ix, g1[i - 1], g2[j]);
      const float b = cur[j - 1] + pr->insert_cost * diff_word + cost_tag(ix, g1[i], g2[j - 1]) + penalty_j1;
      const float c = prev[j - 1] + diff + cost_tag(ix, g1[i - 1], g2[j - 1]);
      float d = a < b ? a : b;
      d = c < d ? c : d;
      cur[j] = d;
      if (d < min) min = d;
    }
    if (min > kmax) kmax = min;
    if (min > max_fuzzyness) { if (rowmin_max) *rowmin_max = kmax; return min; }
    float* t = prev; prev = cur; cur = t;
  }
  if (rowmin_max) *rowmin_max = kmax;
  return prev[n2];
}

/* Attach real tokens / penalty tokens to the indexed sentences (FuzzyMatch::add_tm(id, Sentence, Tokens),
 * reference include/fuzzy/fuzzy_match.hh:53). real / gaps follow the CSR given to fmo_index_create:
 * real[off[s] + i], gaps[off[s] + s + i] (n+1 gaps per sentence). */
int fmo_index_set_real(fmo_index* ix, const int32_t* real, const int32_t* gaps, const int64_t* off, const char* itok_blob,
                       const int32_t* itok_off, int32_t n_itok) {
  const int64_t total = ix->sent_pos[ix->n_sent];
  ix->real = (int32_t*)calloc((size_t)total + 1, 4);
  ix->gap = (int32_t*)calloc((size_t)total + 1, 4);
  for (int64_t k = 0; k < ix->n_sent; k++) {
    const int64_t s = ix->kept_src[k], n = ix->sent_pos[k + 1] - ix->sent_pos[k] - 1;
    for (int64_t i = 0; i < n; i++) ix->real[ix->sent_pos[k] + i] = real[off[s] + i];
    for (int64_t i = 0; i <= n; i++) {
      const int32_t g = gaps[off[s] + s + i];
      if (g < 0 || g >= n_itok) return 1;
      ix->gap[ix->sent_pos[k] + i] = g;
    }
  }
  ix->n_itok = n_itok;
  ix->itok_off = (int32_t*)malloc((size_t)(n_itok + 1) * 4);
  memcpy(ix->itok_off, itok_off, (size_t)(n_itok + 1) * 4);
  ix->itok_blob = (char*)malloc((size_t)itok_off[n_itok] + 1);
  memcpy(ix->itok_blob, itok_blob, (size_t)itok_off[n_itok]);
  return 0;
}

/* ------------------------------------------------------------------ per-thread scratch */

typedef struct { uint32_t sid; uint32_t lm; } sidlm;

typedef struct scratch_t {
  /* open-addressing map sentence id -> longest n-gram match */
  uint32_t* hkey; uint32_t* hval; int64_t hcap, hcount;
  sidlm* cands; int64_t cands_cap;
  float* heap; int64_t heap_cap, heap_n;
  fmo_match* res; int64_t res_cap;
  float* dp; int64_t dp_cap;
  float* idf; int32_t* pat; int64_t pat_cap;
  float* pen_sum; float* pen_max; int32_t* pen_n;
  /* sentence ids already in the caller's `matches` vector when match() is called (the reference appends, counts them
   * against number_of_matches and penalises the contrastive candidates against them, src/fuzzy_match.cc:626-679) */
  const uint32_t* prior; int64_t n_prior;
} scratch_t;

static void hmap_reset(scratch_t* sc) {
  if (!sc->hkey) {
    sc->hcap = 1024;
    sc->hkey = (uint32_t*)malloc((size_t)sc->hcap * 4);
    sc->hval = (uint32_t*)malloc((size_t)sc->hcap * 4);
  }
  memset(sc->hval, 0, (size_t)sc->hcap * 4); /* value 0 == empty (match lengths are >= 1) */
  sc->hcount = 0;
}
static uint32_t hash32(uint32_t x) {
  x = ((x >> 16) ^ x) * 0x45d9f3bu; x = ((x >> 16) ^ x) * 0x45d9f3bu; return (x >> 16) ^ x;
}
static void hmap_put_max(scratch_t* sc, uint32_t key, uint32_t val);
static void hmap_grow(scratch_t* sc) {
  uint32_t* ok = sc->hkey; uint32_t* ov = sc->hval; const int64_t oc = sc->hcap;
  sc->hcap *= 2;
  sc->hkey = (uint32_t*)malloc((size_t)sc->hcap * 4);
  sc->hval = (uint32_t*)calloc((size_t)sc->hcap, 4);
  sc->hcount = 0;
  for (int64_t i = 0; i < oc; i++) if (ov[i]) hmap_put_max(sc, ok[i], ov[i]);
  free(ok); free(ov);
}
static void hmap_put_max(scratch_t* sc, uint32_t key, uint32_t val) {
  if ((sc->hcount + 1) * 2 > sc->hcap) hmap_grow(sc);
  int64_t h = hash32(key) & (uint32_t)(sc->hcap - 1);
  while (sc->hval[h] && sc->hkey[h] != key) h = (h + 1) & (sc->hcap - 1);
  if (!sc->hval[h]) { sc->hkey[h] = key; sc->hval[h] = val; sc->hcount++; }
  else if (val > sc->hval[h]) sc->hval[h] = val;
}

static int cmp_cand(const void* a, const void* b) { /* longest match desc, sentence id asc */
  const sidlm* x = (const sidlm*)a; const sidlm* y = (const sidlm*)b;
  if (x->lm != y->lm) return x->lm > y->lm ? -1 : 1;
  return x->sid < y->sid ? -1 : (x->sid > y->sid ? 1 : 0);
}
static int cmp_match(const void* a, const void* b) { /* score desc, sentence id asc (CompareMatch) */
  const fmo_match* x = (const fmo_match*)a; const fmo_match* y = (const fmo_match*)b;
  if (x->score != y->score) return x->score > y->score ? -1 : 1;
  return x->s_id < y->s_id ? -1 : (x->s_id > y->s_id ? 1 : 0);
}

/* max-heap of floats == std::priority_queue<float> lowest_costs */
static void heap_push(scratch_t* sc, float v) {
  if (sc->heap_n + 1 > sc->heap_cap) {
    sc->heap_cap = sc->heap_cap ? sc->heap_cap * 2 : 64;
    sc->heap = (float*)realloc(sc->heap, (size_t)sc->heap_cap * sizeof(float));
  }
  int64_t i = sc->heap_n++;
  while (i > 0 && sc->heap[(i - 1) / 2] < v) { sc->heap[i] = sc->heap[(i - 1) / 2]; i = (i - 1) / 2; }
  sc->heap[i] = v;
}
static void heap_pop(scratch_t* sc) {
  const float v = sc->heap[--sc->heap_n];
  int64_t i = 0;
  for (;;) {
    int64_t c = 2 * i + 1;
    if (c >= sc->heap_n) break;
    if (c + 1 < sc->heap_n && sc->heap[c + 1] > sc->heap[c]) c++;
    if (!(sc->heap[c] > v)) break;
    sc->heap[i] = sc->heap[c];
    i = c;
  }
  if (sc->heap_n > 0) sc->heap[i] = v;
}

static void scratch_free(scratch_t* sc) {
  free(sc->hkey); free(sc->hval); free(sc->cands); free(sc->heap); free(sc->res); free(sc->dp);
  free(sc->idf); free(sc->pat); free(sc->pen_sum); free(sc->pen_max); free(sc->pen_n);
}

/* NGramMatches::register_suffix_range_match (reference src/ngram_matches.cc:62-84) */
static void register_range(const fmo_index* ix, scratch_t* sc, int64_t begin, int64_t end, uint32_t match_length,
                           uint32_t min_seq_len, int64_t p, const fmo_params* pr, fmo_counters* ct) {
  if (match_length < min_seq_len) return;
  for (int64_t i = begin; i < end; i++) {
    if (ct) ct->elements_walked++;
    if (theoretical_rejection(p, ix->sa_len[i], pr)) continue;
    hmap_put_max(sc, ix->sa_sid[i], match_length);
  }
}

/* ------------------------------------------------------------------ match() */

/* FuzzyMatch::match core (reference src/fuzzy_match.cc:435-681). Writes up to cap matches, returns
 * the number the reference would append. dbg (optional, dbg_cap entries) receives the candidate
 * records in processing order; *dbg_n their count. */
static int64_t match_one(const fmo_index* ix, const int32_t* pattern_in, const int32_t* p_real, const int32_t* p_gap,
                         int64_t p_length, const fmo_params* pr, scratch_t* sc, fmo_match* out, int64_t cap,
                         fmo_counters* ct, fmo_candidate* dbg, int64_t dbg_cap, int64_t* dbg_n) {
  if (dbg_n) *dbg_n = 0;
  if (ct) ct->queries++;
  int contrast_buffer = pr->contrast_buffer;
  const unsigned number_of_matches = (unsigned)pr->number_of_matches;
  if (contrast_buffer == -1) contrast_buffer = (int)number_of_matches;     /* :451-452 */
  if (p_length > ix->max_tokens) return 0;                                  /* :455-458 */
  if (!p_length) return 0;                                                  /* :460-461 */
  int min_subseq_length = pr->min_subseq_length;
  if ((size_t)min_subseq_length > (size_t)p_length) min_subseq_length = (int)p_length;           /* :463-464 */
  if ((int)(pr->min_subseq_ratio * p_length) > min_subseq_length)
    min_subseq_length = (int)(pr->min_subseq_ratio * p_length);                                   /* :466-467 */
  if (ct) ct->pattern_tokens += p_length;

  if (p_length > sc->pat_cap) {
    sc->pat_cap = p_length + 64;
    sc->pat = (int32_t*)realloc(sc->pat, (size_t)sc->pat_cap * 4);
    sc->idf = (float*)realloc(sc->idf, (size_t)sc->pat_cap * 4);
  }
  /* vocabulary lookup: anything that is not a TM word is VOCAB_UNK (src/vocab_indexer.cc:52-60) */
  int32_t* pattern = sc->pat;
  for (int64_t j = 0; j < p_length; j++) {
    const int32_t t = pattern_in[j];
    pattern[j] = (t >= 2 && t < ix->vocab_size && ix->sfreq[t] > 0) ? t : 1;
  }
  /* IDF (src/fuzzy_match.cc:367-390, 472-477) */
  float idf_max = 0.01f;
  const float vocab_idf_penalty = pr->vocab_idf_penalty;
  if (vocab_idf_penalty) {
    const unsigned num_sentences = (unsigned)ix->n_sent_global;
    for (int64_t j = 0; j < p_length; j++)
      sc->idf[j] = pattern[j] != 1 ? logf((float)num_sentences / (float)ix->sfreq[pattern[j]]) : 0.f;
    idf_max = (float)log((double)num_sentences);
  }

  /* n-gram walk (src/fuzzy_match.cc:482-551) */
  hmap_reset(sc);
  const uint32_t min_seq_len = (uint32_t)min_subseq_length;
  if (p_length == 1) {
    int64_t lo, hi;
    equal_range(ix, pattern, 1, 0, 0, &lo, &hi, ct);
    if (lo != hi) register_range(ix, sc, lo, hi, 1, min_seq_len, p_length, pr, ct);
  }
  for (int64_t it = 0; it < p_length; it++) {
    int64_t prev_lo = 0, prev_hi = 0;
    int64_t subseq_length = 0;
    for (int64_t jt = it; jt < p_length; jt++) {
      ++subseq_length;
      int64_t lo, hi;
      equal_range(ix, pattern + it, subseq_length, prev_lo, prev_hi, &lo, &hi, ct);
      if (lo != hi) {
        if (subseq_length > 2) {
          register_range(ix, sc, prev_lo, lo, (uint32_t)(subseq_length - 1), min_seq_len, p_length, pr, ct);
          register_range(ix, sc, hi, prev_hi, (uint32_t)(subseq_length - 1), min_seq_len, p_length, pr, ct);
        }
        prev_lo = lo; prev_hi = hi;
      } else {
        --subseq_length;
        break;
      }
    }
    if (subseq_length >= 2)
      register_range(ix, sc, prev_lo, prev_hi, (uint32_t)subseq_length, min_seq_len, p_length, pr, ct);
  }

  /* get_longest_matches (src/ngram_matches.cc:20-29) */
  if (sc->hcount > sc->cands_cap) {
    sc->cands_cap = sc->hcount + 1024;
    sc->cands = (sidlm*)realloc(sc->cands, (size_t)sc->cands_cap * sizeof(sidlm));
  }
  int64_t n_cand = 0;
  for (int64_t h = 0; h < sc->hcap; h++)
    if (sc->hval[h]) { sc->cands[n_cand].sid = sc->hkey[h]; sc->cands[n_cand].lm = sc->hval[h]; n_cand++; }
  qsort(sc->cands, (size_t)n_cand, sizeof(sidlm), cmp_cand);

  /* candidate loop (src/fuzzy_match.cc:557-612) */
  const int64_t dp_need = 2 * (p_length + 1) + 2 * (ix->max_tokens + 2);
  if (dp_need > sc->dp_cap) { sc->dp_cap = dp_need; sc->dp = (float*)realloc(sc->dp, (size_t)dp_need * sizeof(float)); }
  sc->heap_n = 0;
  heap_push(sc, FLT_MAX);
  int64_t n_res = 0;
  for (int64_t c = 0; c < n_cand; c++) {
    const uint32_t s_id = sc->cands[c].sid;
    const int64_t longest_match = sc->cands[c].lm;
    const int32_t* sentence = ix->buf + ix->sent_pos[s_id];
    const int64_t s_length = ix->sent_pos[s_id + 1] - ix->sent_pos[s_id] - 1;
    if (ct) { ct->candidates++; ct->candidate_tokens += s_length; }
    const int64_t cover = longest_match < p_length ? count_covered_words(pattern, p_length, sentence, s_length) : p_length;
    fmo_candidate* d = (dbg && c < dbg_cap) ? &dbg[c] : NULL;
    if (d) {
      d->s_id = s_id; d->longest_match = (int32_t)longest_match; d->s_length = (int32_t)s_length;
      d->cover = (int32_t)cover; d->rejected = 1; d->cost_full = 0.f; d->rowmin_max = 0.f; d->accepted = 0;
    }
    if (theoretical_rejection_cover(p_length, s_length, cover, pr)) continue;
    const float diff_word = 100.f / get_normalizer(p_length, s_length, pr); /* costs.hh:54-57 */
    const float cost_upper_bound = sc->heap[0];
    const float idf_weight = diff_word * vocab_idf_penalty / idf_max;
    if (ct) { ct->dp_pairs++; ct->dp_tokens += s_length; ct->dp_cells += s_length * p_length; }
    float cost;
    if (p_real && ix->real) { /* Sentence API: real-token and penalty-token terms */
      const int32_t* r1 = ix->real + ix->sent_pos[s_id];
      const int32_t* g1 = ix->gap + ix->sent_pos[s_id];
      cost = edit_distance_real(ix, sentence, r1, g1, (int)s_length, pattern, p_real, p_gap, (int)p_length, sc->idf,
                                idf_weight, pr, diff_word, cost_upper_bound, NULL, sc->dp);
      if (d) {
        d->rejected = 0;
        d->cost_full = edit_distance_real(ix, sentence, r1, g1, (int)s_length, pattern, p_real, p_gap, (int)p_length,
                                          sc->idf, idf_weight, pr, diff_word, FLT_MAX, &d->rowmin_max, sc->dp);
      }
    } else {
      cost = edit_distance_full(sentence, (int)s_length, pattern, (int)p_length, sc->idf, idf_weight, pr, diff_word,
                                cost_upper_bound, NULL, sc->dp);
      if (d) {
        d->rejected = 0;
        d->cost_full = edit_distance_full(sentence, (int)s_length, pattern, (int)p_length, sc->idf, idf_weight, pr,
                                          diff_word, FLT_MAX, &d->rowmin_max, sc->dp);
      }
    }
    if ((pr->no_perfect && cost == 0 && s_length == p_length) || cost > cost_upper_bound) continue;
    const float score = (float)((int)(10000 - cost * 100) / 10000.0);
    heap_push(sc, cost);
    if (score < pr->fuzzy || (contrast_buffer > 0 && sc->heap_n > (int64_t)contrast_buffer)) heap_pop(sc);
    if (score >= pr->fuzzy) {
      if (n_res + 1 > sc->res_cap) {
        sc->res_cap = sc->res_cap ? sc->res_cap * 2 : 64;
        sc->res = (fmo_match*)realloc(sc->res, (size_t)sc->res_cap * sizeof(fmo_match));
      }
      fmo_match* m = &sc->res[n_res++];
      m->s_id = s_id; m->score = score; m->penalty = 0.f; m->max_subseq = (int32_t)longest_match;
      m->length = (int32_t)s_length; m->cost = cost;
      if (d) d->accepted = 1;
    }
  }
  if (dbg_n) *dbg_n = n_cand < dbg_cap ? n_cand : dbg_cap;

  /* result heap drained best-first: score desc, s_id asc (src/fuzzy_match.cc:25-33) */
  qsort(sc->res, (size_t)n_res, sizeof(fmo_match), cmp_match);
  int64_t n_out = 0;
  if (pr->contrastive_factor > 0) { /* src/fuzzy_match.cc:613-669 */
    sc->pen_sum = (float*)realloc(sc->pen_sum, (size_t)(n_res + 1) * 4);
    sc->pen_max = (float*)realloc(sc->pen_max, (size_t)(n_res + 1) * 4);
    sc->pen_n = (int32_t*)realloc(sc->pen_n, (size_t)(n_res + 1) * 4);
    for (int64_t i = 0; i < n_res; i++) { sc->pen_sum[i] = 0.f; sc->pen_max[i] = 0.f; sc->pen_n[i] = 0; }
    int64_t remaining = n_res;
    fmo_params unit_costs = *pr;
    unit_costs.insert_cost = unit_costs.delete_cost = unit_costs.replace_cost = 1.f; /* :630 */
    const fmo_match* last = NULL;
    fmo_match last_copy;
    /* penalties are accumulated in the order of `matches`: the entries that were there before the call first
     * (j < n_prior), then the matches selected by this call (last) */
    int64_t j_prior = 0;
    while (remaining > 0 && (number_of_matches == 0 || (uint64_t)(n_out + sc->n_prior) < number_of_matches)) {
      while (last || j_prior < sc->n_prior) { /* rescore penalties against the entries not yet accounted for (the older ones are memoised) */
        const uint32_t other = last ? last->s_id : sc->prior[j_prior];
        const int32_t other_len = last ? last->length : (int32_t)(ix->sent_pos[other + 1] - ix->sent_pos[other] - 1);
        for (int64_t i = 0; i < remaining; i++) {
          fmo_match* m = &sc->res[i];
          const int32_t* a = ix->buf + ix->sent_pos[m->s_id];
          const int32_t* b = ix->buf + ix->sent_pos[other];
          const float dw = 100.f / get_normalizer(m->length, other_len, &unit_costs); /* Costs(c.len, m.len, EditCosts()) */
          float pen = edit_distance_plain(a, m->length, b, other_len, 1.f, 1.f, 1.f, dw, sc->dp);
          pen = (float)((int)(10000 - pen * 100) / 10000.0);
          sc->pen_sum[i] = sc->pen_sum[i] + pen;
          sc->pen_max[i] = (sc->pen_n[i] == 0 || pen > sc->pen_max[i]) ? pen : sc->pen_max[i];
          sc->pen_n[i]++;
          m->penalty = pr->contrast_reduce == 1 ? sc->pen_max[i] : sc->pen_sum[i] / (float)sc->pen_n[i];
        }
        if (last) last = NULL; else j_prior++;
      }
      int64_t best = 0; /* std::max_element: first maximum in list order */
      for (int64_t i = 1; i < remaining; i++) {
        const float kb = sc->res[best].score - pr->contrastive_factor * sc->res[best].penalty;
        const float ki = sc->res[i].score - pr->contrastive_factor * sc->res[i].penalty;
        if (kb < ki) best = i;
      }
      last_copy = sc->res[best];
      last = &last_copy;
      if (out && n_out < cap) out[n_out] = last_copy;
      n_out++;
      for (int64_t i = best; i + 1 < remaining; i++) { /* list erase keeps order */
        sc->res[i] = sc->res[i + 1];
        sc->pen_sum[i] = sc->pen_sum[i + 1]; sc->pen_max[i] = sc->pen_max[i + 1]; sc->pen_n[i] = sc->pen_n[i + 1];
      }
      remaining--;
    }
  } else { /* src/fuzzy_match.cc:670-679 */
    for (int64_t i = 0; i < n_res && (number_of_matches == 0 || (uint64_t)(n_out + sc->n_prior) < number_of_matches); i++) {
      if (out && n_out < cap) out[n_out] = sc->res[i];
      n_out++;
    }
  }
  if (ct) ct->matches_out += n_out;
  return n_out;
}

/* ------------------------------------------------------------------ batch API */

typedef struct job_t {
  const fmo_index* ix;
  const int32_t* q_tokens; const int64_t* q_off; int64_t n_q;
  const int32_t* q_real; const int32_t* q_gaps; /* optional (Sentence API) */
  const uint32_t* prior_sid; const int64_t* prior_off; /* optional: entries already in `matches` per query (CSR) */
  const fmo_params* pr;
  int64_t cap; fmo_match* out; int32_t* out_count;
  int64_t next; pthread_mutex_t mu;
  fmo_counters total;
  int want_counters;
} job_t;

static void* worker(void* arg) {
  job_t* jb = (job_t*)arg;
  scratch_t sc; memset(&sc, 0, sizeof sc);
  fmo_counters ct; memset(&ct, 0, sizeof ct);
  for (;;) {
    const int64_t q0 = __atomic_fetch_add(&jb->next, 16, __ATOMIC_RELAXED);
    if (q0 >= jb->n_q) break;
    const int64_t q1 = q0 + 16 < jb->n_q ? q0 + 16 : jb->n_q;
    for (int64_t q = q0; q < q1; q++) {
      sc.prior = jb->prior_sid ? jb->prior_sid + jb->prior_off[q] : NULL;
      sc.n_prior = jb->prior_sid ? jb->prior_off[q + 1] - jb->prior_off[q] : 0;
      const int64_t n = match_one(jb->ix, jb->q_tokens + jb->q_off[q], jb->q_real ? jb->q_real + jb->q_off[q] : NULL,
                                  jb->q_gaps ? jb->q_gaps + jb->q_off[q] + q : NULL, jb->q_off[q + 1] - jb->q_off[q], jb->pr, &sc,
                                  jb->out ? jb->out + q * jb->cap : NULL, jb->cap, jb->want_counters ? &ct : NULL, NULL, 0, NULL);
      if (jb->out_count) jb->out_count[q] = (int32_t)n;
    }
  }
  if (jb->want_counters) {
    pthread_mutex_lock(&jb->mu);
    int64_t* a = (int64_t*)&jb->total; const int64_t* b = (const int64_t*)&ct;
    for (size_t i = 0; i < sizeof(fmo_counters) / sizeof(int64_t); i++) a[i] += b[i];
    pthread_mutex_unlock(&jb->mu);
  }
  scratch_free(&sc);
  return NULL;
}

/* Matches n_q queries (CSR) with nthreads workers sharing the index. out is [n_q*cap];
 * out_count[q] is the number of matches the reference would return (may exceed cap). */
void fmo_match_batch_real(const fmo_index* ix, const int32_t* q_tokens, const int32_t* q_real, const int32_t* q_gaps,
                          const int64_t* q_off, int64_t n_q, const fmo_params* pr, int nthreads, int64_t cap,
                          fmo_match* out, int32_t* out_count, fmo_counters* counters);
void fmo_match_batch(const fmo_index* ix, const int32_t* q_tokens, const int64_t* q_off, int64_t n_q,
                     const fmo_params* pr, int nthreads, int64_t cap, fmo_match* out, int32_t* out_count,
                     fmo_counters* counters) {
  fmo_match_batch_real(ix, q_tokens, NULL, NULL, q_off, n_q, pr, nthreads, cap, out, out_count, counters);
}
/* Same with the pattern's real tokens and penalty tokens (match(const Sentence& real, const Tokens&, ...),
 * reference include/fuzzy/fuzzy_match.hh:70-82); q_gaps holds n+1 itok ids per query at q_off[q] + q. */
void fmo_match_batch_real(const fmo_index* ix, const int32_t* q_tokens, const int32_t* q_real, const int32_t* q_gaps,
                          const int64_t* q_off, int64_t n_q, const fmo_params* pr, int nthreads, int64_t cap,
                          fmo_match* out, int32_t* out_count, fmo_counters* counters) {
  job_t jb; memset(&jb, 0, sizeof jb);
  jb.ix = ix; jb.q_tokens = q_tokens; jb.q_off = q_off; jb.n_q = n_q; jb.pr = pr; jb.q_real = q_real; jb.q_gaps = q_gaps;
  jb.cap = cap; jb.out = out; jb.out_count = out_count; jb.want_counters = counters != NULL;
  pthread_mutex_init(&jb.mu, NULL);
  if (nthreads <= 1) {
    worker(&jb);
  } else {
    pthread_t* th = (pthread_t*)malloc((size_t)nthreads * sizeof(pthread_t));
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, worker, &jb);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
  }
  if (counters) *counters = jb.total;
  pthread_mutex_destroy(&jb.mu);
}

/* match(Tokens) into result vectors that already hold matches of the same TM: prior_sid[prior_off[q] .. prior_off[q+1])
 * are their sentence ids in vector order. Returns per query what the call appends (src/fuzzy_match.cc:626-679). */
void fmo_match_batch_prior(const fmo_index* ix, const int32_t* q_tokens, const int64_t* q_off, int64_t n_q, const fmo_params* pr,
                           const uint32_t* prior_sid, const int64_t* prior_off, int64_t cap, fmo_match* out, int32_t* out_count) {
  job_t jb; memset(&jb, 0, sizeof jb);
  jb.ix = ix; jb.q_tokens = q_tokens; jb.q_off = q_off; jb.n_q = n_q; jb.pr = pr; jb.prior_sid = prior_sid; jb.prior_off = prior_off;
  jb.cap = cap; jb.out = out; jb.out_count = out_count;
  pthread_mutex_init(&jb.mu, NULL);
  worker(&jb);
  pthread_mutex_destroy(&jb.mu);
}

/* One query with the per-candidate trace (kernel-level parity tests). Returns the match count;
 * *dbg_n receives the number of candidate records written (processing order). */
int64_t fmo_match_debug(const fmo_index* ix, const int32_t* pattern, int64_t p_length, const fmo_params* pr,
                        int64_t cap, fmo_match* out, fmo_candidate* dbg, int64_t dbg_cap, int64_t* dbg_n) {
  scratch_t sc; memset(&sc, 0, sizeof sc);
  const int64_t n = match_one(ix, pattern, NULL, NULL, p_length, pr, &sc, out, cap, NULL, dbg, dbg_cap, dbg_n);
  scratch_free(&sc);
  return n;
}

/* The suffix-array range of one n-gram (search-kernel parity tests). */
void fmo_equal_range(const fmo_index* ix, const int32_t* ngram, int64_t length, int64_t* lo, int64_t* hi) {
  int64_t plo = 0, phi = 0;
  for (int64_t k = 1; k <= length; k++) {
    equal_range(ix, ngram, k, plo, phi, lo, hi, NULL);
    if (*lo == *hi) return;
    plo = *lo; phi = *hi;
  }
}

/* ------------------------------------------------------------------ subsequence()
 *
 * FuzzyMatch::subsequence (reference src/fuzzy_match.cc:238-365) behind its tokenizer: the pattern arrives as
 * word ids (real == normalised tokens, no penalty tokens, like match(Tokens)); the text the reference appends to
 * Match::id (the detokenised sub-sequence) is described by (position, length) instead. */
typedef struct fmo_subseq {
  uint32_t s_id;
  float score;
  float cost;
  int32_t position; /* first pattern token of the sub-sequence that located the match */
  int32_t length;   /* its length = Match::max_subseq */
  int32_t found;
} fmo_subseq;

typedef struct { float weight; int32_t position; int32_t length; } subseq_t;
/* priority_queue<Subseq> pops the largest element: weight desc, then position asc (operator< :238-248) */
static int cmp_subseq(const void* a, const void* b) {
  const subseq_t* x = (const subseq_t*)a; const subseq_t* y = (const subseq_t*)b;
  if (x->weight != y->weight) return x->weight > y->weight ? -1 : 1;
  if (x->position != y->position) return x->position < y->position ? -1 : 1;
  return x->length > y->length ? -1 : (x->length < y->length ? 1 : 0); /* (only reachable with zero idf weights) */
}

static void subsequence_one(const fmo_index* ix, const int32_t* pattern_in, int64_t p_length, unsigned number_of_matches,
                            int no_perfect, int min_subseq_length, float min_subseq_ratio, int idf_weighting, fmo_subseq* out) {
  memset(out, 0, sizeof *out);
  if ((int)(min_subseq_ratio * p_length) > min_subseq_length) min_subseq_length = (int)(min_subseq_ratio * p_length); /* :263-264 */
  if ((int)p_length < min_subseq_length) return;                                                                   /* :266-267 */
  int32_t* pidx = (int32_t*)malloc((size_t)(p_length + 1) * 4);
  float* idf = (float*)malloc((size_t)(p_length + 1) * 4);
  const unsigned num_sentences = (unsigned)ix->n_sent_global;
  for (int64_t j = 0; j < p_length; j++) {
    const int32_t t = pattern_in[j];
    pidx[j] = (t >= 2 && t < ix->vocab_size && ix->sfreq[t] > 0) ? t : 1;
    idf[j] = pidx[j] != 1 ? logf((float)num_sentences / (float)ix->sfreq[pidx[j]]) : -1.f; /* :372-390, unknown = -1 */
  }
  /* all sub-sequences without unknown words, by weight (:277-288) */
  subseq_t* sq = (subseq_t*)malloc((size_t)(p_length * (p_length + 1) / 2 + 1) * sizeof(subseq_t));
  int64_t n_sq = 0;
  for (int64_t it = 0; it < p_length; it++) {
    float idf_weight = 0;
    for (int64_t jt = it; jt < p_length; jt++) {
      const float weight = idf[jt];
      if (weight == -1) break;
      idf_weight += idf_weighting ? weight : 1;
      if ((int)(jt - it + 1) >= min_subseq_length) { sq[n_sq].weight = idf_weight; sq[n_sq].position = (int32_t)it; sq[n_sq].length = (int32_t)(jt - it + 1); n_sq++; }
    }
  }
  qsort(sq, (size_t)n_sq, sizeof(subseq_t), cmp_subseq);
  int max_distance = 10000;
  uint32_t* candidates = (uint32_t*)malloc(((size_t)number_of_matches + 1) * 4);
  size_t n_cand = 0, n_perfect = 0, perfect_cap = 16;
  uint32_t* perfect = (uint32_t*)malloc(perfect_cap * 4);
  float* scratch = (float*)malloc((size_t)(2 * (p_length + 2)) * 4);
  fmo_params unit; memset(&unit, 0, sizeof unit);
  unit.insert_cost = unit.delete_cost = unit.replace_cost = 1.f;
  for (int64_t k = 0; k < n_sq && max_distance == 10000; k++) { /* :300-301 */
    int64_t lo = 0, hi = 0, plo = 0, phi = 0;
    for (int64_t len = 1; len <= sq[k].length; len++) { /* canonical range of the whole n-gram (:306) */
      equal_range(ix, pidx + sq[k].position, len, plo, phi, &lo, &hi, NULL);
      if (lo == hi) break;
      plo = lo; phi = hi;
    }
    for (int64_t su = lo; su < hi && n_cand < number_of_matches; su++) { /* :308-309 */
      const uint32_t s_id = ix->sa_sid[su];
      int seen = 0;
      for (size_t i = 0; i < n_cand && !seen; i++) seen = candidates[i] == s_id;
      for (size_t i = 0; i < n_perfect && !seen; i++) seen = perfect[i] == s_id;
      if (seen) continue;
      const int64_t s_length = ix->sent_pos[s_id + 1] - ix->sent_pos[s_id] - 1;
      const float diff_word = 100.f / get_normalizer(p_length, s_length, &unit); /* Costs(p, s, EditCosts()) :317-318 */
      const float cost = edit_distance_full(ix->buf + ix->sent_pos[s_id], (int)s_length, pidx, (int)p_length, idf, 0.f, &unit, diff_word,
                                            (float)max_distance, NULL, scratch); /* :321-326 */
      if (cost == 0 && no_perfect) { /* :327-330 */
        if (n_perfect == perfect_cap) { perfect_cap *= 2; perfect = (uint32_t*)realloc(perfect, perfect_cap * 4); }
        perfect[n_perfect++] = s_id;
        continue;
      }
      if (cost < max_distance) { /* :331-350 */
        out->found = 1;
        out->score = (float)((int)(10000 - cost * 100) / 10000.0);
        out->cost = cost;
        out->length = sq[k].length;
        out->position = sq[k].position;
        out->s_id = s_id;
        max_distance = (int)cost; /* int max_distance = cost: truncated */
        if (cost == 0) break;
      }
      candidates[n_cand++] = s_id;
    }
  }
  free(pidx); free(idf); free(sq); free(candidates); free(perfect); free(scratch);
}

void fmo_subsequence_batch(const fmo_index* ix, const int32_t* q_tokens, const int64_t* q_off, int64_t n_q, int32_t number_of_matches,
                           int32_t no_perfect, int32_t min_subseq_length, float min_subseq_ratio, int32_t idf_weighting, fmo_subseq* out) {
  for (int64_t q = 0; q < n_q; q++)
    subsequence_one(ix, q_tokens + q_off[q], q_off[q + 1] - q_off[q], (unsigned)number_of_matches, no_perfect, min_subseq_length,
                    min_subseq_ratio, idf_weighting, out + q);
}
