// fm_index.cu -- host-side index build and upload.
//
// Replaces, for pre-tokenised int32 input, what the reference does in FuzzyMatch::add_tm(Tokens) +
// sort() (reference src/suffix_array_index.cc:10-30, src/suffix_array.cc:9-27,58-102,253-261,
// src/vocab_indexer.cc:73-90): drop empty / over-long sentences, count word-in-sentence
// frequencies, sort the sentence-bounded suffixes and build the first-word bucket table.
// The order among suffixes that compare equal is immaterial to match() (ranges are sets), so any
// total order works; here ties break by position, which equals the reference's sentence-id order.
// The build runs on host threads for now (a GPU build is the first "next" row of SURVEY.md 8f).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <numeric>
#include <thread>

#include "fm_internal.h"

namespace fm {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
const std::string& get_error() { return g_error; }
int cuda_fail(cudaError_t e, const char* what) {
  set_error(std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what);
  return FM_ERR_CUDA;
}

template <class T>
static int upload(const std::vector<T>& h, size_t extra, void** slot, const T** out, int64_t* bytes) {
  void* d = nullptr;
  const size_t n = (h.size() + extra) * sizeof(T);
  FM_CUDA(cudaMalloc(&d, n ? n : 16));
  FM_CUDA(cudaMemset(d, 0, n ? n : 16));
  if (!h.empty()) FM_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  *slot = d;
  *out = static_cast<const T*>(d);
  *bytes += (int64_t)n;
  return FM_OK;
}

// IDF table with host libm exactly as src/fuzzy_match.cc:367-390: logf((float)N / (float)sfreq[w]) per
// word, idf_max = (float)log((double)N).
int set_idf_stats(Index* ix, const uint32_t* sf, int64_t n_sent_global) {
  const unsigned num_sentences = (unsigned)n_sent_global;
  std::vector<float> idf((size_t)ix->vocab_size, 0.f);
  for (int32_t w = 2; w < ix->vocab_size; w++)
    if (sf[w] > 0) idf[w] = std::log((float)num_sentences / (float)sf[w]);
  FM_CUDA(cudaMemcpy(const_cast<float*>(ix->dev.idf), idf.data(), idf.size() * sizeof(float), cudaMemcpyHostToDevice));
  if (sf != ix->sfreq.data()) std::copy(sf, sf + ix->vocab_size, ix->sfreq.begin());
  ix->dev.idf_max = (float)std::log((double)num_sentences);
  return FM_OK;
}

void free_index(Index* ix) {
  if (!ix) return;
  cudaSetDevice(ix->device);
  for (void* p : ix->d_blocks)
    if (p) cudaFree(p);
  delete ix;
}

int build_index(const int32_t* tokens, const int64_t* sent_off, int64_t n_in, int32_t vocab_size, int32_t max_tokens,
                const uint32_t* sfreq_global, int64_t n_sent_global, int64_t s_id_base, int device, Index** out) {
  if (!out) { set_error("out is NULL"); return FM_ERR_INVALID; }
  *out = nullptr;
  if (n_in < 0 || (n_in > 0 && (!tokens || !sent_off))) { set_error("bad TM arrays"); return FM_ERR_INVALID; }
  if (vocab_size < 2) { set_error("vocab_size must be >= 2"); return FM_ERR_INVALID; }
  if (max_tokens < 1 || max_tokens > FM_MAX_TOKENS) {
    set_error("max_tokens_in_pattern must be in [1, " + std::to_string(FM_MAX_TOKENS) + "]");
    return FM_ERR_INVALID;
  }
  Index* ix = new Index();
  ix->device = device;
  ix->vocab_size = vocab_size;
  ix->max_tokens = max_tokens;

  // ---- kept sentences and the padded token buffer
  int64_t n_keep = 0, n_suf = 0, n_buf = 0;
  for (int64_t s = 0; s < n_in; s++) {
    const int64_t len = sent_off[s + 1] - sent_off[s];
    if (len < 0) { delete ix; set_error("sent_off is not non-decreasing"); return FM_ERR_INVALID; }
    if (len > 0 && len <= max_tokens) {
      n_keep++;
      n_suf += len;
      n_buf += (len + 1 + 3) & ~int64_t(3);
    }
  }
  n_buf += 8;  // zero tail so 128-bit loads of the last sentence stay in bounds
  if (n_buf >= (int64_t(1) << 31) - 64) { delete ix; set_error("TM shard too large for int32 offsets"); return FM_ERR_INVALID; }
  ix->n_sent = n_keep;
  ix->n_suf = n_suf;
  ix->n_buf = n_buf;
  ix->h_tok.assign((size_t)n_buf, 0);
  ix->h_sent_start.resize((size_t)n_keep + 1);
  ix->kept.resize((size_t)n_keep);
  ix->sfreq.assign((size_t)vocab_size, 0);
  std::vector<uint32_t> meta_of_pos((size_t)n_buf, 0);
  std::vector<int32_t> sid_at((size_t)(n_buf / 4) + 1, -1);
  {
    std::vector<int64_t> stamp((size_t)vocab_size, -1);
    int64_t cur = 0, k = 0;
    for (int64_t s = 0; s < n_in; s++) {
      const int64_t len = sent_off[s + 1] - sent_off[s];
      if (!(len > 0 && len <= max_tokens)) continue;
      ix->h_sent_start[k] = (int32_t)cur;
      ix->kept[k] = s;
      sid_at[cur >> 2] = (int32_t)k;
      for (int64_t i = 0; i < len; i++) {
        const int32_t t = tokens[sent_off[s] + i];
        if (t < 2 || t >= vocab_size) {
          delete ix;
          set_error("TM token id outside [2, vocab_size) in sentence " + std::to_string(s));
          return FM_ERR_INVALID;
        }
        ix->h_tok[cur + i] = t;
        meta_of_pos[cur + i] = ((uint32_t)len << 16) | (uint32_t)i;
        if (stamp[t] != k) { stamp[t] = k; ix->sfreq[t]++; }
      }
      cur += (len + 1 + 3) & ~int64_t(3);
      k++;
    }
    ix->h_sent_start[n_keep] = (int32_t)cur;
  }

  // ---- suffix sort: counting sort on the first token, then each bucket by the rest
  std::vector<int32_t> qva((size_t)vocab_size + 1, 0);
  std::vector<int32_t> sa((size_t)n_suf);
  {
    std::vector<int64_t> cnt((size_t)vocab_size + 1, 0);
    for (int64_t k = 0; k < n_keep; k++)
      for (int32_t pos = ix->h_sent_start[k]; ix->h_tok[pos] != 0; pos++) cnt[ix->h_tok[pos] + 1]++;
    for (int32_t w = 0; w < vocab_size; w++) cnt[w + 1] += cnt[w];
    for (int32_t w = 0; w <= vocab_size; w++) qva[w] = (int32_t)cnt[w];
    for (int64_t k = 0; k < n_keep; k++)
      for (int32_t pos = ix->h_sent_start[k]; ix->h_tok[pos] != 0; pos++) sa[cnt[ix->h_tok[pos]]++] = pos;
  }
  {
    std::vector<int32_t> order;  // non-trivial buckets, largest first
    for (int32_t w = 2; w < vocab_size; w++)
      if (qva[w + 1] - qva[w] > 1) order.push_back(w);
    std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return qva[a + 1] - qva[a] > qva[b + 1] - qva[b]; });
    const int32_t* tok = ix->h_tok.data();
    auto less = [tok](int32_t a, int32_t b) {
      const int32_t* x = tok + a + 1;
      const int32_t* y = tok + b + 1;
      for (;; x++, y++) {
        if (*x != *y) return *x < *y;
        if (*x == 0) return a < b;
      }
    };
    std::atomic<size_t> next(0);
    auto work = [&]() {
      for (;;) {
        const size_t i = next.fetch_add(1);
        if (i >= order.size()) break;
        const int32_t w = order[i];
        std::sort(sa.begin() + qva[w], sa.begin() + qva[w + 1], less);
      }
    };
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 4;
    if (nt > 64) nt = 64;
    if (order.size() < 64) nt = 1;
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nt; t++) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
  }
  // per-suffix walk record: (sentence start, length, signature lo, signature hi)
  std::vector<int4> sa_walk((size_t)n_suf);
  {
    std::vector<unsigned long long> sig_of_sent((size_t)n_keep, 0);
    for (int64_t k = 0; k < n_keep; k++) {
      unsigned long long sg = 0;
      for (int32_t pos = ix->h_sent_start[k]; ix->h_tok[pos] != 0; pos++) sg |= 1ull << sig_bit(ix->h_tok[pos]);
      sig_of_sent[k] = sg;
    }
    for (int64_t i = 0; i < n_suf; i++) {
      const uint32_t m = meta_of_pos[sa[i]];
      const int32_t start = sa[i] - (int32_t)(m & 0xffffu);
      const unsigned long long sg = sig_of_sent[sid_at[start >> 2]];
      sa_walk[i] = make_int4(start, (int32_t)(m >> 16), (int32_t)(uint32_t)sg, (int32_t)(uint32_t)(sg >> 32));
    }
  }
  std::vector<uint32_t>().swap(meta_of_pos);

  // ---- bigram directory: one entry per distinct (word0, word1) with its suffix-array range
  std::vector<int4> bg_tab;
  uint32_t bg_mask = 0;
  {
    const int32_t* tok = ix->h_tok.data();
    int64_t n_bg = 0;
    for (int64_t i2 = 0; i2 < n_suf; i2++) {
      const int32_t t1 = tok[sa[i2] + 1];
      if (t1 != 0 && (i2 == 0 || tok[sa[i2 - 1]] != tok[sa[i2]] || tok[sa[i2 - 1] + 1] != t1)) n_bg++;
    }
    uint64_t cap = 1024;
    while (cap < (uint64_t)n_bg * 2) cap <<= 1;
    bg_mask = (uint32_t)(cap - 1);
    bg_tab.assign((size_t)cap, make_int4(-1, -1, 0, 0));
    for (int64_t i2 = 0; i2 < n_suf;) {
      const int32_t t0 = tok[sa[i2]], t1 = tok[sa[i2] + 1];
      int64_t j2 = i2 + 1;
      while (j2 < n_suf && tok[sa[j2]] == t0 && tok[sa[j2] + 1] == t1) j2++;
      if (t1 != 0) {
        uint32_t hsl = bigram_hash(t0, t1) & bg_mask;
        while (bg_tab[hsl].x != -1) hsl = (hsl + 1) & bg_mask;
        bg_tab[hsl] = make_int4(t0, t1, (int32_t)i2, (int32_t)j2);
      }
      i2 = j2;
    }
  }

  // ---- trigram directory: (slot of the bigram in bg_tab, word2) -> suffix-array range
  std::vector<int4> tg_tab;
  uint32_t tg_mask = 0;
  {
    const int32_t* tok = ix->h_tok.data();
    auto is_tri = [&](int64_t k) { return tok[sa[k] + 1] != 0 && tok[sa[k] + 2] != 0; };
    auto same_tri = [&](int64_t a, int64_t b2) {
      return tok[sa[a]] == tok[sa[b2]] && tok[sa[a] + 1] == tok[sa[b2] + 1] && tok[sa[a] + 2] == tok[sa[b2] + 2];
    };
    int64_t n_tg = 0;
    for (int64_t k = 0; k < n_suf; k++)
      if (is_tri(k) && (k == 0 || !same_tri(k - 1, k))) n_tg++;
    uint64_t cap = 1024;
    while (cap < (uint64_t)n_tg * 2) cap <<= 1;
    tg_mask = (uint32_t)(cap - 1);
    tg_tab.assign((size_t)cap, make_int4(-1, -1, 0, 0));
    for (int64_t k = 0; k < n_suf;) {
      int64_t j2 = k + 1;
      if (!is_tri(k)) { k = j2; continue; }
      while (j2 < n_suf && same_tri(k, j2)) j2++;
      const int32_t t0 = tok[sa[k]], t1 = tok[sa[k] + 1], t2 = tok[sa[k] + 2];
      uint32_t bs = bigram_hash(t0, t1) & bg_mask;
      while (!(bg_tab[bs].x == t0 && bg_tab[bs].y == t1)) bs = (bs + 1) & bg_mask;
      uint32_t hsl = bigram_hash((int32_t)bs, t2) & tg_mask;
      while (tg_tab[hsl].x != -1) hsl = (hsl + 1) & tg_mask;
      tg_tab[hsl] = make_int4((int32_t)bs, t2, (int32_t)k, (int32_t)j2);
      k = j2;
    }
  }

  // ---- upload
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) { delete ix; return cuda_fail(e, "cudaSetDevice"); }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ix->sm_count = prop.multiProcessorCount;
  int rc;
  IndexDev& d = ix->dev;
  if ((rc = upload(ix->h_tok, 0, &ix->d_blocks[0], &d.tok, &ix->device_bytes)) ||
      (rc = upload(sa, 4, &ix->d_blocks[1], &d.sa_pos, &ix->device_bytes)) ||
      (rc = upload(sa_walk, 4, &ix->d_blocks[2], &d.sa_walk, &ix->device_bytes)) ||
      (rc = upload(qva, 0, &ix->d_blocks[3], &d.qva, &ix->device_bytes)) ||
      (rc = upload(bg_tab, 0, &ix->d_blocks[6], &d.bg_tab, &ix->device_bytes)) ||
      (rc = upload(tg_tab, 0, &ix->d_blocks[7], &d.tg_tab, &ix->device_bytes)) ||
      (rc = upload(sid_at, 0, &ix->d_blocks[4], &d.sid_at, &ix->device_bytes)) ||
      (rc = upload(std::vector<float>((size_t)vocab_size, 0.f), 0, &ix->d_blocks[5], &d.idf, &ix->device_bytes)) ||
      (rc = set_idf_stats(ix, sfreq_global ? sfreq_global : ix->sfreq.data(), n_sent_global > 0 ? n_sent_global : n_keep))) {
    free_index(ix);
    return rc;
  }
  d.bg_mask = bg_mask;
  d.tg_mask = tg_mask;
  d.vocab_size = vocab_size;
  d.max_tokens = max_tokens;
  d.n_suf = n_suf;
  d.sid_base = (uint32_t)s_id_base;
  *out = ix;
  return FM_OK;
}

}  // namespace fm
