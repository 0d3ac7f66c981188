// fm_internal.h -- shared declarations of the B200 fuzzy-match library (host + device).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/fuzzy_match_b200.h"

namespace fm {

// ---------------------------------------------------------------- device-resident index (HBM layout)
//
// tok      int32[n_buf]   kept sentences back to back; every sentence starts on a 16-byte boundary
//                         (multiple of 4 tokens) and is followed by >= 1 zero (the separator); the zero
//                         sorts before every word id, so "shorter suffix first" needs no length check.
// sa_pos   int32[n_suf]   suffix array: absolute offset into tok of each suffix, sorted.
// sa_rec   uint2[n_suf]   what the range walk tests per suffix, 8 bytes so that one 256-bit load brings four:
//                         bits 0-5 = sentence length (<= kWideMin), bits 6-63 = 58-bit word signature of the
//                         sentence (bit sig_bit(w) set for every word w): an upper bound on the coverage
//                         without touching the sentence. Long sentences: bits 0-5 = 63, bits 6-15 = length,
//                         high word = row of the sentence's wide signature (wsig).
// sa_aux   int4[2*n_suf]  what the verify kernel needs of an element that passed the walk's test, in ONE 32-byte read
//                         (a single sector): word 0 = sentence start in tok, word 1 = sentence length, words 2..7 =
//                         second, independent 192-bit word signature of the sentence (sig2_bit; tested before the
//                         sentence is fetched for the exact count). Long sentences: word 1 = length | 1 << 31,
//                         word 2 = row of the wide signature.
// sa_next  int32[n_suf]   token at depth 3 of each suffix (tok[sa_pos[k] + 3], 0 when the suffix is shorter): keys of
//                         the bisection that narrows a trigram range in subsequence() -- one array instead of
//                         sa_pos -> tok (two dependent misses); the search kernel uses qg_tab at that level.
// tg_tab   int4[2*pow2]   trigram directory, 32-byte entries (one sector): (word0, word1, word2, lo | hi, -, -, -), hi =
//                         -position-1 for a trigram that occurs once. Keyed by the three words, not by the bigram's
//                         slot: with min_subseq_length >= 3 (the reference's CLI default) a chain that does not reach
//                         a trigram registers nothing, so the search probes this table FIRST and never reads the
//                         bigram directory -- one random sector and one dependent round less per chain.
// qg_tab   int4[pow2]     4-gram directory: (trigram slot, word3) -> [lo, hi), or (lo, -position-1) for a 4-gram that
//                         occurs once; only for trigrams that occur more than once. The fourth narrowing step --
//                         the one where ranges are still wide (a frequent trigram: thousands of suffixes, a dozen
//                         dependent bisection rounds) -- in one probe.
// qva      int32[V+1]     first-word bucket table (reference _quickVocabAccess).
// bg_tab   int4[pow2]     bigram directory: open-addressing table (word0, word1) -> [lo, hi) of the suffixes
//                         that start with that bigram (whole array -> first word -> bigram in one probe instead of
//                         ~40); read when min_subseq_length < 3 -- bigram ranges are registered then -- and by
//                         subsequence().
// sid_at   int32[n_buf/4] local sentence id, stored at (sentence start / 4); only read for survivors.
// wsig     uint32[n_wide*32] 1024-bit word signatures of the sentences longer than kWideMin tokens (a 64-bit
//                         signature saturates there); their walk records carry the row number instead of the
//                         64-bit signature. One row = one 128-byte line.
// idf      float[V]       logf(N / sfreq[w]) computed on the host with glibc (0 for unseen words).
struct IndexDev {
  const int32_t* tok;
  const int32_t* sa_pos;
  const uint2* sa_rec;
  const int4* sa_aux;
  const int32_t* sa_next;
  const int32_t* qva;
  const int4* bg_tab;
  uint32_t bg_mask;
  const int4* tg_tab;
  uint32_t tg_mask;
  const int4* qg_tab;
  uint32_t qg_mask;
  const int32_t* sid_at;
  const uint32_t* wsig;
  int32_t n_wide;
  const float* idf;
  const int32_t* real;  // optional (Sentence API): (real form id << 1) | case class, parallel to tok
  const int32_t* gap;   // optional: penalty-token id of the gap before each token (separator slot = trailing gap)
  int32_t vocab_size;
  int32_t max_tokens;
  int64_t n_suf;
  int32_t n_buf;  // tokens in tok (the last one is a separator)
  uint32_t sid_base;
  float idf_max;  // (float)log((double)N_sent_global)
  const int32_t* sent_start;  // optional [n_sent]: start of local sentence s in tok (uploaded on first use: contrastive rerank on a sharded TM)
};

// per-query metadata written by the prepare kernel
//   x = pattern length p (0 if the query is skipped), y = effective min_subseq_length,
//   z = offset of the pattern in the token arrays,
//   w = bit0: query takes part; bits 8..17: max(0, (largest number of pattern positions on one signature bit) - 3);
//       bits 18..27: the same for the second signature
typedef int4 QMeta;
static const int kQValid = 1;

// word -> signature bit 6..63 (must be identical on host and device); bits 0-5 of a record hold the length
__host__ __device__ inline unsigned sig_bit(int w) { return 6u + (((((unsigned)w * 0x9E3779B1u) >> 16) * 58u) >> 16); }
// second, independent word -> bit map (kSig2Words * 32 = 192 bits) of the per-sentence signature the verify kernel
// tests before it fetches a sentence (sa_aux): three times the bits of the walk's signature, so that few candidates whose
// exact coverage fails get as far as the sentence fetch (at f=0.5: 2.7 times fewer than with 64 bits)
static const int kSig2Words = 6;
__host__ __device__ inline unsigned sig2_bit(int w) {
  unsigned x = (unsigned)w * 0x7FEB352Du;
  x ^= x >> 15;
  x *= 0x846CA68Bu;
  x ^= x >> 16;
  return (((x & 0xffffu) * (unsigned)kSig2Words) >> 16) * 32u + (x >> 27);
}
// sentences longer than this carry a 1024-bit signature (wsig) instead of the 64-bit one
static const int kWideMin = 48;
static const int kWideWords = 32;  // 32-bit words per wide signature
// Per query, for the wide signatures: three bit-sliced planes of min(pattern positions per wide bit, 7), then up to
// kWideBig entries (bit | excess << 16) for the bits that collect more than 7 positions, then one word: the excess
// of further such bits (counted as present).
static const int kWideBig = 8;
static const int kWideStride = 3 * kWideWords + 16;
__host__ __device__ inline unsigned wsig_bit(int w) {
  unsigned x = (unsigned)w * 0x85EBCA6Bu;
  x ^= x >> 15;
  return (x * 0xC2B2AE35u) >> 22;
}
// bigram -> slot hash (host build and device lookup)
__host__ __device__ inline uint32_t bigram_hash(int w0, int w1) {
  unsigned long long k = ((unsigned long long)(unsigned)w0 << 32) | (unsigned)w1;
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return (uint32_t)k;
}

// trigram -> first slot probed in the trigram directory
__host__ __device__ inline uint32_t trigram_hash(int w0, int w1, int w2) { return bigram_hash((int)bigram_hash(w0, w1), w2); }

// A scored candidate: what the replay of the candidate loop needs. rowmin_max = max over DP rows of the row
// minimum (reproduces the reference's early exit); reserved[0] = sentence start in tok (contrastive rerank),
// reserved[1..2] = scratch of the rerank.
struct fm_record {
  uint32_t s_id;  // global sentence id
  int32_t longest_match;
  int32_t length;
  float cost;
  float rowmin_max;
  int32_t reserved[3];
};

struct SurvRec {  // one distinct (query, sentence) that passed both rejection bounds
  int32_t q;
  int32_t start;  // sentence start in tok (unique per sentence, ascending with s_id)
  int32_t hslot;  // slot in the dedup table (holds the running max of the n-gram match length)
  int32_t j;      // arrival index inside its query
};

struct Counters {
  unsigned long long slice_elem;  // (n_slices << 38) | n_elements, one packed atomic
  unsigned int n_surv;
  unsigned int overflow;  // bit0 slices, bit1 survivors, bit2 span index, bit3 candidate list
  unsigned int n_matches;
  unsigned int n_heavy;  // queries with more than kWarpMax scored candidates (CTA each)
  unsigned int n_mid;    // queries with 2..kWarpMax scored candidates (warp each)
  unsigned int n_prep;  // queries the thread-per-query prepare kernel left to the warp-per-query one (prep_list)
  unsigned int n_small;  // slices of at most kSmallSlice elements (their own list, a lane each)
  unsigned int n_long;  // bit0 / bit1: some survivor's pattern is too long for the first / second scoring kernel
  unsigned int n_cand;   // suffix-array elements that passed stage 1 of the gather (the candidate list)
  unsigned int n_verified;  // exact coverage counts (profiling)
  unsigned int wire_need;   // merge stage: largest number of accepted records of one shard
};
static const int kElemBits = 38;
#ifndef FM_SMALL_SLICE
#define FM_SMALL_SLICE 8
#endif
static const int kSmallSlice = FM_SMALL_SLICE;
static const int kSpan = 1024;  // flattened elements per gather work unit (one warp)

// per-batch workspace pointers (device)
struct BatchDev {
  // inputs
  const int32_t* q_tok_in;
  const int32_t* q_off;  // [n_q+1]
  const int32_t* q_real;     // optional (Sentence API) [n_tok]
  const int32_t* q_gap;      // optional [n_tok + n_q]: p+1 gaps per query at q_off[q] + q
  const int32_t* itok_dist;  // optional [n_itok^2] pairwise _edit_distance_char of the penalty tokens
  int32_t n_itok;
  int32_t n_q;
  int32_t n_tok;
  // prepared
  int32_t* pat;      // [n_tok] sanitised pattern tokens
  int2* chain_rec;   // [n_tok] per chain (= pattern position) of the search: (query, start position | pattern length << 10 | mult << 20); length 0 = dead
  int32_t* prep_list;  // [n_q] queries for the warp-per-query prepare kernel
  QMeta* qmeta;      // [n_q]
  int2* tbl;         // [4*n_tok] per-query open-addressing tables: (word, distinct_idx | count<<16)
  const uint16_t* cmin_tab; // [(max_tokens+1) << 10] per (pattern length << 10 | sentence length): smallest coverage that passes
  const uint16_t* cmin64;   // [(max_tokens+1) << 6] the same for the 6-bit length field of a walk record (stage 1 of the gather)
  int4* qmask2;      // [3*n_q] the same planes over the 192 bits of the second signature (sig2_bit): B0 words 0..5, then B1 words 0..5; its mult sits in qmeta.w bits 18..27
  int4* qmask;       // [n_q] per query, in record layout: planes (B0 lo, B0 hi, B1 lo, B1 hi) of min(pattern positions per signature bit, 3)
  uint32_t* wq;      // [kWideStride*n_q] or NULL (index without wide signatures): planes and excess list over the 1024 wide bits
  unsigned long long* peq64;  // [n_tok] patterns of <= 64 tokens: position mask of each distinct word, at q_off + distinct index
  // search output: slices of more than kSmallSlice elements, flattened, and the small ones
  long long* sl_start;  // [slice_cap+1] first flattened element of each slice (ascending)
  int4* sl_rec;         // [2*slice_cap] per slice (q, sa_begin, match_len | p << 10 | mult << 20, size) and the query's
                        //               signature planes (qmask[q]): everything stage 1 needs in one 32-byte record
  int4* sm_rec;         // [2*slice_cap] same records, slices of <= kSmallSlice elements
  int64_t slice_cap;
  int32_t* span_slice;  // [span_cap] slice holding flattened element k*kSpan
  int64_t span_cap;
  // gather output
  unsigned long long* hkey;  // [hsize] (q<<32 | start), ~0 = empty
  unsigned int* hlm;         // [hsize] max match length
  uint32_t hmask;
  int2* cand;         // [cand_cap] walk kernel -> verify kernel: (q | match length << 20, suffix-array index)
  int64_t cand_cap;
  SurvRec* surv;      // [surv_cap]
  uint16_t* surv_len; // [surv_cap]
  int64_t surv_cap;
  int32_t* q_cnt;   // [n_q+1] distinct survivors per query
  int32_t* q_base;  // [n_q+1] exclusive scan of q_cnt
  // scoring / replay
  fm_record* rec;   // [surv_cap] grouped by query
  float* heapbuf;   // [surv_cap + n_q]
  int32_t* acc_cnt; // [n_q] accepted matches per query (contrastive path)
  Counters* ctr;
};

struct Params {  // fm_params + derived
  float fuzzy;
  int32_t n_matches;
  int32_t no_perfect;
  int32_t ml;
  float mr;
  float idf_penalty;
  float ins, del, rep;
  float contrast;
  int32_t reduce;
  int32_t buffer;  // already resolved (-1 -> n_matches)
};

// ---------------------------------------------------------------- host objects

struct Workspace {
  int device = 0;
  cudaStream_t stream = nullptr;  // owned stream for the host-buffer API
  cudaStream_t stream2 = nullptr;  // side stream: the CTA-per-query replay runs next to the warp-per-query one
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_done = nullptr;  // recorded behind the counter read-back of a submitted batch
  int64_t surv_hint = -1;         // survivors of the previous batch on this workspace: sizes the dedup table
  uint32_t hs_use = 0;            // dedup-table slots cleared and used by the batch in flight (power of two <= hsize)
  // capacities
  int64_t cap_q = 0, cap_tok = 0, cap_slices = 0, cap_surv = 0, cap_out = 0, cap_cand = 0;
  uint32_t hsize = 0;
  // device buffers
  int32_t *d_q_tok = nullptr, *d_q_off = nullptr;  // staging for host inputs
  int32_t *d_q_real = nullptr, *d_q_gap = nullptr, *d_itok_dist = nullptr;  // Sentence API staging
  int64_t cap_real_tok = 0, cap_real_gap = 0, cap_itok = 0;
  int32_t n_itok = 0;
  bool real_active = false;  // the batch in flight carries real tokens / penalty tokens
  // entries already in the callers' result vectors (fm_match_batch_prior): (sentence start in tok, length) per entry
  int2* d_prior = nullptr;
  int32_t* d_prior_off = nullptr;  // [n_q+1]
  int64_t cap_prior = 0, cap_prior_q = 0;
  bool prior_active = false;
  int32_t *pat = nullptr, *prep_list = nullptr;
  int2* chain_rec = nullptr;
  QMeta* qmeta = nullptr;
  int2* tbl = nullptr;
  uint16_t* cmin_tab = nullptr;
  uint16_t* cmin64 = nullptr;
  int4* sm_rec = nullptr;
  int32_t* span_slice = nullptr;
  int64_t cap_spans = 0;
  Params bounds_params{};
  bool bounds_valid = false;
  int4* qmask = nullptr;
  int4* qmask2 = nullptr;
  uint32_t* wq = nullptr;
  unsigned long long* peq64 = nullptr;
  int64_t cap_wq = 0;
  long long* sl_start = nullptr;
  int4* sl_rec = nullptr;
  unsigned long long* hkey = nullptr;
  unsigned int* hlm = nullptr;
  SurvRec* surv = nullptr;
  int2* cand = nullptr;
  uint16_t* surv_len = nullptr;
  int32_t *q_cnt = nullptr, *q_base = nullptr, *acc_cnt = nullptr, *heavy_q = nullptr, *m_heavy = nullptr, *mid_q = nullptr, *m_mid = nullptr;
  fm_record* rec = nullptr;
  float* heapbuf = nullptr;
  unsigned long long *sort_key = nullptr, *m_key = nullptr, *sort_key2 = nullptr, *m_key2 = nullptr;
  int32_t *sort_idx = nullptr, *m_idx = nullptr;
  Counters* ctr = nullptr;
  unsigned long long* scan_chain = nullptr;  // scan mailboxes: (epoch << 32 | tile total) per CTA
  unsigned int scan_epoch = 0;
  fm_match* d_out = nullptr;
  int32_t* d_out_count = nullptr;
  // sharded TM: accepted records of the batch before they are packed (sized like the survivor arrays), this
  // rank's block and the gathered blocks of all ranks
  fm_wire* wire_stage = nullptr;
  bool want_stage = false;
  char* wire_send = nullptr;
  char* wire_recv = nullptr;
  int64_t cap_wire_block = 0;
  // merged-shard buffers
  Counters* mctr = nullptr;    // counters of the merge stage (the shard stage's stay intact for the overflow check)
  Counters* h_mctr = nullptr;  // pinned
  fm_record* mrec = nullptr;
  int32_t *m_cnt = nullptr, *m_base = nullptr, *m_acc = nullptr;
  float* m_heap = nullptr;
  int64_t cap_mrec = 0, cap_mq = 0;
  // contrastive rerank on a sharded TM: tokens of every accepted sentence, filled by the owning shard and summed over the ranks
  int32_t* ctok = nullptr;
  int32_t *c_cnt = nullptr, *c_base = nullptr;
  int64_t cap_ctok = 0, cap_cq = 0;
  int32_t* h_ctotal = nullptr;  // pinned
  // pinned host staging
  Counters* h_ctr = nullptr;
  int32_t* h_q_off32 = nullptr;
  int64_t cap_hq = 0;
  // profiling
  cudaEvent_t ev[10] = {};  // 0..6 stage boundaries, 7 host timing, 8 between the two gather kernels
  bool in_use = false;
};

struct Index {
  int device = 0;
  IndexDev dev{};
  // host copies
  std::vector<int32_t> h_tok;
  std::vector<int32_t> h_sent_start;  // [n_sent+1]
  std::vector<int64_t> kept;
  std::vector<uint32_t> sfreq;
  int64_t n_sent = 0, n_suf = 0, n_buf = 0;
  int64_t device_bytes = 0;
  int32_t vocab_size = 0, max_tokens = 0;
  void* d_blocks[16] = {};
  size_t blk_bytes[16] = {};
  int64_t n_sent_global = 0;
  int sm_count = 148;
  // workspaces
  std::mutex mu;
  std::vector<Workspace*> pool;
  int32_t max_gap_id = 0;  // largest penalty-token id of the TM side (fm_index_set_real): the caller's table must cover it
  int32_t* d_sent_start = nullptr;  // device copy of h_sent_start (lazily, under mu)
  std::mutex shard_mu;  // the sharded calls of one index run one at a time (collectives must not interleave)
  bool profiling = false;
  fm_profile last_profile{};
};

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
#define FM_CUDA(call)                                     \
  do {                                                    \
    cudaError_t e__ = (call);                             \
    if (e__ != cudaSuccess) return fm::cuda_fail(e__, #call); \
  } while (0)

// fm_index.cu
int build_index(const int32_t* tokens, const int64_t* sent_off, int64_t n_sent, int32_t vocab_size, int32_t max_tokens,
                const uint32_t* sfreq_global, int64_t n_sent_global, int64_t s_id_base, int device, Index** out);
void free_index(Index* ix);
int save_index(const Index* ix, const char* path);
int load_index(const char* path, int device, Index** out);
int set_idf_stats(Index* ix, const uint32_t* sfreq, int64_t n_sent_global);
int set_real(Index* ix, const int32_t* real, const int32_t* gaps, const int64_t* sent_off, int64_t n_sent);

// fm_sort.cu
int gpu_suffix_sort(const int32_t* d_tok, int64_t n_buf, const std::vector<int32_t>& sent_start,
                    const std::vector<int32_t>& compact_off, int64_t n_suf, int max_len, int32_t vocab_size, int sm_count,
                    int32_t* d_sa);

// fm_kernels.cu -- launchers (all asynchronous on `st`)
void launch_bounds(const IndexDev& ix, const BatchDev& b, const Params& p, cudaStream_t st);
int launch_prepare(const IndexDev& ix, const BatchDev& b, const Params& p, int sm_count, cudaStream_t st);
void launch_search(const IndexDev& ix, const BatchDev& b, const Params& p, cudaStream_t st);
void launch_gather(const IndexDev& ix, const BatchDev& b, const Params& p, int sm_count, cudaStream_t st, cudaEvent_t between);  // walk + verify
void launch_scan(const int32_t* in, int32_t* out, int32_t n, unsigned long long* chain, unsigned int epoch, int sm_count,
                 cudaStream_t st);
void launch_score(const IndexDev& ix, const BatchDev& b, const Params& p, int sm_count, cudaStream_t st);
void launch_replay(const IndexDev& ix, const fm_record* rec, const int32_t* q_cnt, const int32_t* q_base, float* heapbuf,
                   unsigned long long* sort_key, unsigned long long* sort_key2, int32_t* sort_idx, int32_t* acc_cnt,
                   int32_t* mid_q, int32_t* heavy_q, const int32_t* q_off, int32_t n_q, const Params& p, int64_t cap,
                   fm_match* out, int32_t* out_count, Counters* ctr, int sm_count, cudaStream_t st, cudaStream_t st2, cudaEvent_t ev_fork, cudaEvent_t ev_join,
                   int32_t* wire_cnt = nullptr, fm_wire* wire_stage = nullptr);  // wire_cnt != NULL: shard mode (accepted records out)
void launch_contrast(const IndexDev& ix, fm_record* rec, const int32_t* q_base, const int32_t* sort_idx,
                     const int32_t* acc_cnt, int32_t n_q, const Params& p, int64_t cap, fm_match* out, int32_t* out_count,
                     Counters* ctr, int sm_count, cudaStream_t st, const int2* prior = nullptr, const int32_t* prior_off = nullptr);
void launch_contrast_need(const fm_record* rec, const int32_t* q_base, const int32_t* sort_idx, const int32_t* acc_cnt, int32_t n_q,
                          int32_t* tok_cnt, cudaStream_t st);
void launch_contrast_fill(const IndexDev& ix, int64_t n_sent_local, fm_record* rec, const int32_t* q_base, const int32_t* sort_idx,
                          const int32_t* acc_cnt, const int32_t* tok_base, int32_t n_q, int32_t* slab, cudaStream_t st);
void launch_subseq(const IndexDev& ix, const int32_t* q_tok, const int32_t* q_off, int32_t n_q, int n_matches, int no_perfect, int ml, float mr,
                   int idf_weighting, uint32_t* seen, int seen_cap, fm_subseq* out, cudaStream_t st);
// wire blocks (one shard's accepted records of a batch; layout in fm_kernels.cu / include/fuzzy_match_b200.h)
inline long long wire_off_words_host(long long n_q) { return (n_q + 1 + 3) / 4 * 4; }
inline long long wire_block_bytes(long long n_q, long long capacity) { return 4 * (4 + wire_off_words_host(n_q)) + (long long)sizeof(fm_wire) * capacity; }
void launch_wire_pack(int32_t* block, const Counters* ctr, const fm_wire* stage, const int32_t* q_base, int32_t n_q, int capacity,
                      cudaStream_t st);
void launch_wire_count(int n_shards, const int32_t* const* blocks, int32_t* m_cnt, int32_t n_q, Counters* mctr, cudaStream_t st);
void launch_wire_copy(int n_shards, const int32_t* const* blocks, const int32_t* m_base, fm_record* mrec, int32_t n_q,
                      const Counters* mctr, cudaStream_t st);

}  // namespace fm
