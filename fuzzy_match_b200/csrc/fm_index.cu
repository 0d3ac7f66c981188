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
#include <cstdio>
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
static int upload(const std::vector<T>& h, size_t extra, Index* ix, int blk, const T** out) {
  void* d = nullptr;
  const size_t n = (h.size() + extra) * sizeof(T);
  FM_CUDA(cudaMalloc(&d, n ? n : 16));
  FM_CUDA(cudaMemset(d, 0, n ? n : 16));
  if (!h.empty()) FM_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  ix->d_blocks[blk] = d;
  ix->blk_bytes[blk] = n;
  *out = static_cast<const T*>(d);
  ix->device_bytes += (int64_t)n;
  return FM_OK;
}

// ---- on-disk format: one flat little-endian file (replaces the reference's
// Boost binary archive, src/fuzzy_matcher_binarization.cc). Header, then the device blocks exactly as
// they sit in HBM, then the host-side tables. Loading is read + upload: no sort, no hashing.
static const char kMagic[8] = {'F', 'M', 'B', '2', '0', '0', 'I', 1};
enum { BLK_TOK = 0, BLK_SA = 1, BLK_WALK = 2, BLK_QVA = 3, BLK_SID = 4, BLK_IDF = 5, BLK_BG = 6, BLK_TG = 7, N_BLK = 8 };

static void bind_blocks(Index* ix) {
  IndexDev& d = ix->dev;
  d.tok = static_cast<const int32_t*>(ix->d_blocks[BLK_TOK]);
  d.sa_pos = static_cast<const int32_t*>(ix->d_blocks[BLK_SA]);
  d.sa_walk = static_cast<const int4*>(ix->d_blocks[BLK_WALK]);
  d.qva = static_cast<const int32_t*>(ix->d_blocks[BLK_QVA]);
  d.sid_at = static_cast<const int32_t*>(ix->d_blocks[BLK_SID]);
  d.idf = static_cast<const float*>(ix->d_blocks[BLK_IDF]);
  d.bg_tab = static_cast<const int4*>(ix->d_blocks[BLK_BG]);
  d.tg_tab = static_cast<const int4*>(ix->d_blocks[BLK_TG]);
}

int save_index(const Index* ix, const char* path) {
  FILE* f = fopen(path, "wb");
  if (!f) { set_error(std::string("cannot open ") + path + " for writing"); return FM_ERR_INVALID; }
  FM_CUDA(cudaSetDevice(ix->device));
  int64_t hdr[16] = {1, ix->vocab_size, ix->max_tokens, ix->n_sent, ix->n_suf, ix->n_buf, (int64_t)ix->dev.bg_mask,
                     (int64_t)ix->dev.tg_mask, (int64_t)ix->dev.sid_base, ix->n_sent_global, 0, 0, 0, 0, 0, 0};
  memcpy(&hdr[10], &ix->dev.idf_max, sizeof(float));
  bool ok = fwrite(kMagic, 1, 8, f) == 8 && fwrite(hdr, sizeof(int64_t), 16, f) == 16;
  std::vector<char> buf;
  for (int k = 0; k < N_BLK && ok; k++) {
    const int64_t n = (int64_t)ix->blk_bytes[k];
    buf.resize((size_t)n);
    if (n && cudaMemcpy(buf.data(), ix->d_blocks[k], (size_t)n, cudaMemcpyDeviceToHost) != cudaSuccess) ok = false;
    ok = ok && fwrite(&n, sizeof n, 1, f) == 1 && (n == 0 || fwrite(buf.data(), 1, (size_t)n, f) == (size_t)n);
  }
  ok = ok && fwrite(ix->h_sent_start.data(), sizeof(int32_t), ix->h_sent_start.size(), f) == ix->h_sent_start.size();
  ok = ok && fwrite(ix->kept.data(), sizeof(int64_t), ix->kept.size(), f) == ix->kept.size();
  ok = ok && fwrite(ix->sfreq.data(), sizeof(uint32_t), ix->sfreq.size(), f) == ix->sfreq.size();
  ok = (fclose(f) == 0) && ok;
  if (!ok) { set_error(std::string("write to ") + path + " failed"); return FM_ERR_INVALID; }
  return FM_OK;
}

int load_index(const char* path, int device, Index** out) {
  *out = nullptr;
  FILE* f = fopen(path, "rb");
  if (!f) { set_error(std::string("cannot open ") + path); return FM_ERR_INVALID; }
  char magic[8];
  int64_t hdr[16];
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, kMagic, 8) != 0 || fread(hdr, sizeof(int64_t), 16, f) != 16 || hdr[0] != 1) {
    fclose(f);
    set_error(std::string(path) + " is not a fuzzy_match_b200 index (version 1)");
    return FM_ERR_INVALID;
  }
  Index* ix = new Index();
  ix->device = device;
  ix->vocab_size = (int32_t)hdr[1]; ix->max_tokens = (int32_t)hdr[2];
  ix->n_sent = hdr[3]; ix->n_suf = hdr[4]; ix->n_buf = hdr[5]; ix->n_sent_global = hdr[9];
  cudaError_t e = cudaSetDevice(device);
  cudaDeviceProp prop;
  if (e == cudaSuccess && cudaGetDeviceProperties(&prop, device) == cudaSuccess) ix->sm_count = prop.multiProcessorCount;
  bool ok = e == cudaSuccess;
  std::vector<char> buf;
  for (int k = 0; k < N_BLK && ok; k++) {
    int64_t n = 0;
    ok = fread(&n, sizeof n, 1, f) == 1 && n >= 0 && n < (int64_t(1) << 40);
    if (!ok) break;
    buf.resize((size_t)n);
    ok = n == 0 || fread(buf.data(), 1, (size_t)n, f) == (size_t)n;
    void* d = nullptr;
    ok = ok && cudaMalloc(&d, n ? (size_t)n : 16) == cudaSuccess;
    ok = ok && (n == 0 || cudaMemcpy(d, buf.data(), (size_t)n, cudaMemcpyHostToDevice) == cudaSuccess);
    ix->d_blocks[k] = d;
    ix->blk_bytes[k] = (size_t)n;
    ix->device_bytes += n;
    if (k == BLK_TOK && ok) ix->h_tok.assign(reinterpret_cast<int32_t*>(buf.data()), reinterpret_cast<int32_t*>(buf.data()) + n / 4);
  }
  ix->h_sent_start.resize((size_t)ix->n_sent + 1);
  ix->kept.resize((size_t)ix->n_sent);
  ix->sfreq.resize((size_t)ix->vocab_size);
  ok = ok && fread(ix->h_sent_start.data(), sizeof(int32_t), ix->h_sent_start.size(), f) == ix->h_sent_start.size();
  ok = ok && fread(ix->kept.data(), sizeof(int64_t), ix->kept.size(), f) == ix->kept.size();
  ok = ok && fread(ix->sfreq.data(), sizeof(uint32_t), ix->sfreq.size(), f) == ix->sfreq.size();
  fclose(f);
  if (!ok) { free_index(ix); set_error(std::string("reading ") + path + " failed (truncated file or CUDA error)"); return FM_ERR_INVALID; }
  bind_blocks(ix);
  IndexDev& d = ix->dev;
  d.vocab_size = ix->vocab_size; d.max_tokens = ix->max_tokens; d.n_suf = ix->n_suf;
  d.bg_mask = (uint32_t)hdr[6]; d.tg_mask = (uint32_t)hdr[7]; d.sid_base = (uint32_t)hdr[8];
  memcpy(&d.idf_max, &hdr[10], sizeof(float));
  *out = ix;
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
  ix->n_sent_global = n_sent_global;
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
  if ((rc = upload(ix->h_tok, 0, ix, BLK_TOK, &d.tok)) ||
      (rc = upload(sa, 4, ix, BLK_SA, &d.sa_pos)) ||
      (rc = upload(sa_walk, 4, ix, BLK_WALK, &d.sa_walk)) ||
      (rc = upload(qva, 0, ix, BLK_QVA, &d.qva)) ||
      (rc = upload(bg_tab, 0, ix, BLK_BG, &d.bg_tab)) ||
      (rc = upload(tg_tab, 0, ix, BLK_TG, &d.tg_tab)) ||
      (rc = upload(sid_at, 0, ix, BLK_SID, &d.sid_at)) ||
      (rc = upload(std::vector<float>((size_t)vocab_size, 0.f), 0, ix, BLK_IDF, &d.idf)) ||
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
