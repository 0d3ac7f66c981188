// fm_index.cu -- index build, upload, save / load.
//
// Replaces, for pre-tokenised int32 input, what the reference does in FuzzyMatch::add_tm(Tokens) +
// sort() (reference src/suffix_array_index.cc:10-30, src/suffix_array.cc:9-27,58-102,253-261,
// src/vocab_indexer.cc:73-90): drop empty / over-long sentences, count word-in-sentence
// frequencies, sort the sentence-bounded suffixes and build the first-word bucket table.
// The host only lays out the token buffer, counts sfreq and the bucket table (one pass over the
// tokens); the suffix sort (fm_sort.cu) and everything derived from the sorted array (walk records,
// signatures, bigram / trigram directories, kernels below) run on the GPU. The order among suffixes
// that compare equal is immaterial to match() (ranges are sets), so any total order works.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <chrono>
#include <thread>

#include "fm_internal.h"

namespace fm {

struct PhaseTimer {
  bool on = getenv("FM_BUILD_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  void lap(const char* what) {
    if (!on) return;
    const auto n = std::chrono::steady_clock::now();
    fprintf(stderr, "[fm build] %-28s %.3f s\n", what, std::chrono::duration<double>(n - t).count());
    t = n;
  }
};

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
static const int64_t kFileVersion = 7;  // 2: wide signatures (walk records of long sentences carry a wsig row); 4: sa_aux (start + second signature + length); 5: 192-bit second signature, 32-byte sa_aux records; 6: 4-gram directory; 7: trigram directory keyed by the three words, 32-byte entries
enum { BLK_TOK = 0, BLK_SA = 1, BLK_WALK = 2, BLK_QVA = 3, BLK_SID = 4, BLK_IDF = 5, BLK_BG = 6, BLK_TG = 7, BLK_REAL = 8, BLK_GAP = 9, BLK_NEXT = 10, BLK_WSIG = 11, BLK_START = 12, BLK_QG = 13, N_BLK = 14 };

static void bind_blocks(Index* ix) {
  IndexDev& d = ix->dev;
  d.tok = static_cast<const int32_t*>(ix->d_blocks[BLK_TOK]);
  d.sa_pos = static_cast<const int32_t*>(ix->d_blocks[BLK_SA]);
  d.sa_rec = static_cast<const uint2*>(ix->d_blocks[BLK_WALK]);
  d.sa_aux = static_cast<const int4*>(ix->d_blocks[BLK_START]);
  d.sa_next = static_cast<const int32_t*>(ix->d_blocks[BLK_NEXT]);
  d.qva = static_cast<const int32_t*>(ix->d_blocks[BLK_QVA]);
  d.sid_at = static_cast<const int32_t*>(ix->d_blocks[BLK_SID]);
  d.idf = static_cast<const float*>(ix->d_blocks[BLK_IDF]);
  d.bg_tab = static_cast<const int4*>(ix->d_blocks[BLK_BG]);
  d.tg_tab = static_cast<const int4*>(ix->d_blocks[BLK_TG]);
  d.qg_tab = static_cast<const int4*>(ix->d_blocks[BLK_QG]);
  d.wsig = static_cast<const uint32_t*>(ix->d_blocks[BLK_WSIG]);
  d.n_wide = (int32_t)(ix->blk_bytes[BLK_WSIG] / (kWideWords * sizeof(uint32_t)));
  d.real = ix->blk_bytes[BLK_REAL] ? static_cast<const int32_t*>(ix->d_blocks[BLK_REAL]) : nullptr;
  d.gap = ix->blk_bytes[BLK_GAP] ? static_cast<const int32_t*>(ix->d_blocks[BLK_GAP]) : nullptr;
}

int save_index(const Index* ix, const char* path) {
  FM_CUDA(cudaSetDevice(ix->device));
  FILE* f = fopen(path, "wb");
  if (!f) { set_error(std::string("cannot open ") + path + " for writing"); return FM_ERR_INVALID; }
  int64_t hdr[16] = {kFileVersion, ix->vocab_size, ix->max_tokens, ix->n_sent, ix->n_suf, ix->n_buf, (int64_t)ix->dev.bg_mask,
                     (int64_t)ix->dev.tg_mask, (int64_t)ix->dev.sid_base, ix->n_sent_global, 0, N_BLK, (int64_t)ix->dev.qg_mask, 0, 0, 0};
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

// The header is untrusted input: every count is range-checked and cross-checked against the block sizes
// before anything is allocated from it, so a truncated or corrupt file yields FM_ERR_INVALID, not a crash
// or an out-of-bounds read in a later kernel.
static bool header_ok(const int64_t* hdr, const int64_t* blk, int n_blk) {
  const int64_t vocab = hdr[1], max_tok = hdr[2], n_sent = hdr[3], n_suf = hdr[4], n_buf = hdr[5], bgm = hdr[6], tgm = hdr[7], qgm = hdr[12];
  auto pow2m1 = [](int64_t m) { return m >= 0 && m < (int64_t(1) << 32) && ((m + 1) & m) == 0; };
  if (vocab < 2 || vocab > (int64_t(1) << 30) || max_tok < 1 || max_tok > FM_MAX_TOKENS) return false;
  if (n_sent < 0 || n_suf < 0 || n_buf < 8 || n_buf >= (int64_t(1) << 31) || n_sent > n_suf || n_suf > n_buf) return false;
  if (!pow2m1(bgm) || !pow2m1(tgm) || !pow2m1(qgm) || n_blk != N_BLK) return false;
  if (blk[BLK_TOK] != n_buf * 4 || blk[BLK_SA] < n_suf * 4 || blk[BLK_NEXT] < n_suf * 4 || blk[BLK_WALK] < (n_suf + 8) * 8 ||
      blk[BLK_START] < n_suf * 32)
    return false;
  if (blk[BLK_QVA] != (vocab + 1) * 4 || blk[BLK_IDF] != vocab * 4 || blk[BLK_SID] != (n_buf / 4 + 1) * 4) return false;
  if (blk[BLK_BG] != (bgm + 1) * 16 || blk[BLK_TG] != (tgm + 1) * 32 || blk[BLK_QG] != (qgm + 1) * 16) return false;
  if (blk[BLK_WSIG] % (kWideWords * 4) != 0 || blk[BLK_WSIG] / (kWideWords * 4) > n_sent) return false;
  if ((blk[BLK_REAL] != 0 && blk[BLK_REAL] != n_buf * 4) || blk[BLK_GAP] != blk[BLK_REAL]) return false;
  return true;
}

int load_index(const char* path, int device, Index** out) {
  *out = nullptr;
  FILE* f = fopen(path, "rb");
  if (!f) { set_error(std::string("cannot open ") + path); return FM_ERR_INVALID; }
  char magic[8];
  int64_t hdr[16];
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, kMagic, 8) != 0 || fread(hdr, sizeof(int64_t), 16, f) != 16 || hdr[0] != kFileVersion) {
    fclose(f);
    set_error(std::string(path) + " is not a fuzzy_match_b200 index (version " + std::to_string(kFileVersion) + ")");
    return FM_ERR_INVALID;
  }
  Index* ix = nullptr;
  try {
    // pass 1: block sizes (seek over the payloads), validated against the header
    int64_t blk[N_BLK] = {};
    const int n_blk = (int)hdr[11];
    bool ok = n_blk == N_BLK;
    const long data_pos = ftell(f);
    for (int k = 0; k < N_BLK && ok; k++) {
      ok = fread(&blk[k], sizeof(int64_t), 1, f) == 1 && blk[k] >= 0 && blk[k] < (int64_t(1) << 40) && fseek(f, (long)blk[k], SEEK_CUR) == 0;
    }
    if (ok) {  // the host tables must be there in full as well
      const long tables_pos = ftell(f);
      ok = fseek(f, 0, SEEK_END) == 0 && ftell(f) - tables_pos == (hdr[3] + 1) * 4 + hdr[3] * 8 + hdr[1] * 4;
    }
    if (!ok || !header_ok(hdr, blk, n_blk) || fseek(f, data_pos, SEEK_SET) != 0) {
      fclose(f);
      set_error(std::string(path) + ": corrupt or truncated index file");
      return FM_ERR_INVALID;
    }
    ix = new Index();
    ix->device = device;
    ix->vocab_size = (int32_t)hdr[1]; ix->max_tokens = (int32_t)hdr[2];
    ix->n_sent = hdr[3]; ix->n_suf = hdr[4]; ix->n_buf = hdr[5]; ix->n_sent_global = hdr[9];
    cudaError_t e = cudaSetDevice(device);
    cudaDeviceProp prop;
    if (e == cudaSuccess && cudaGetDeviceProperties(&prop, device) == cudaSuccess) ix->sm_count = prop.multiProcessorCount;
    ok = e == cudaSuccess;
    std::vector<char> buf;
    for (int k = 0; k < N_BLK && ok; k++) {
      int64_t n = 0;
      ok = fread(&n, sizeof n, 1, f) == 1 && n == blk[k];
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
      if (k == BLK_GAP && ok)  // penalty-token ids index the caller's table on the device: remember (and bound) the largest
        for (int64_t i = 0; i < n / 4 && ok; i++) {
          const int32_t g = reinterpret_cast<int32_t*>(buf.data())[i];
          ok = g >= 0 && g < 2048;
          if (g > ix->max_gap_id) ix->max_gap_id = g;
        }
    }
    ix->h_sent_start.resize((size_t)ix->n_sent + 1);
    ix->kept.resize((size_t)ix->n_sent);
    ix->sfreq.resize((size_t)ix->vocab_size);
    ok = ok && fread(ix->h_sent_start.data(), sizeof(int32_t), ix->h_sent_start.size(), f) == ix->h_sent_start.size();
    ok = ok && fread(ix->kept.data(), sizeof(int64_t), ix->kept.size(), f) == ix->kept.size();
    ok = ok && fread(ix->sfreq.data(), sizeof(uint32_t), ix->sfreq.size(), f) == ix->sfreq.size();
    fclose(f);
    f = nullptr;
    for (int64_t k = 0; ok && k <= ix->n_sent; k++)  // sentence starts index h_tok: keep them inside it
      ok = ix->h_sent_start[k] >= 0 && ix->h_sent_start[k] < ix->n_buf && (ix->h_sent_start[k] & 3) == 0;
    if (!ok) { free_index(ix); set_error(std::string("reading ") + path + " failed (corrupt file or CUDA error)"); return FM_ERR_INVALID; }
  } catch (const std::exception&) {
    if (f) fclose(f);
    if (ix) free_index(ix);
    set_error(std::string("out of memory while loading ") + path);
    return FM_ERR_NOMEM;
  }
  bind_blocks(ix);
  IndexDev& d = ix->dev;
  d.vocab_size = ix->vocab_size; d.max_tokens = ix->max_tokens; d.n_suf = ix->n_suf; d.n_buf = (int32_t)ix->n_buf;
  d.bg_mask = (uint32_t)hdr[6]; d.tg_mask = (uint32_t)hdr[7]; d.qg_mask = (uint32_t)hdr[12]; d.sid_base = (uint32_t)hdr[8];
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


// ---------------------------------------------------------------- device side of the build
// Once the token buffer and the sorted suffix array are in HBM, everything derived from them is
// data-parallel and is built by these kernels instead of host loops: per-sentence signatures, the
// per-suffix walk records, and the bigram / trigram directories (run boundaries of the suffix array
// inserted into open-addressing tables with 64-bit CAS).

// Per sentence: length, 64-bit word signature, sid_at entry. A sentence longer than kWideMin tokens gets
// row wide_row[s] of the 1024-bit signature table instead (the row was zeroed at allocation; one thread
// owns a row), and its walk records carry that row number in place of the 64-bit signature.
__global__ void fm_build_sentence_kernel(const int32_t* __restrict__ tok, const int32_t* __restrict__ sent_start, int n_sent,
                                         const int32_t* __restrict__ wide_row, uint32_t* wsig, unsigned long long* sig,
                                         int32_t* sent_len, int32_t* sid_at, uint32_t* sig2) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_sent) return;
  const int st = sent_start[s];
  const int row = wide_row[s];
  unsigned long long sg = 0;
  uint32_t sg2[kSig2Words];  // (long sentences are tested against their wide signature instead)
#pragma unroll
  for (int k = 0; k < kSig2Words; k++) sg2[k] = 0;
  int n = 0;
  if (row < 0) {
    for (int t; (t = tok[st + n]) != 0; n++) {
      sg |= 1ull << sig_bit(t);
      const unsigned b2 = sig2_bit(t);
#pragma unroll
      for (int k = 0; k < kSig2Words; k++)
        if ((int)(b2 >> 5) == k) sg2[k] |= 1u << (b2 & 31u);
    }
    sg |= (unsigned long long)n;  // bits 0-5: the length (<= kWideMin < 63)
  } else {
    uint32_t* w = wsig + (size_t)row * kWideWords;
    for (int t; (t = tok[st + n]) != 0; n++) {
      const unsigned b = wsig_bit(t);
      w[b >> 5] |= 1u << (b & 31);
    }
    sg = ((unsigned long long)(unsigned)row << 32) | ((unsigned long long)n << 6) | 63ull;
  }
  sig[s] = sg;  // the walk record of every suffix of this sentence
  sent_len[s] = n;
  sid_at[st >> 2] = s;
#pragma unroll
  for (int k = 0; k < kSig2Words; k++) sig2[(size_t)s * kSig2Words + k] = sg2[k];
}

__global__ void fm_build_walk_kernel(const int32_t* __restrict__ sa_pos, long long n_suf, const int32_t* __restrict__ sent_start,
                                     int n_sent, const unsigned long long* __restrict__ sig,
                                     const uint32_t* __restrict__ sig2, uint2* sa_rec, int4* sa_aux) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_suf) return;
  const int pos = sa_pos[i];
  int a = 0, e = n_sent;  // sentence whose start is the largest one <= pos
  while (e - a > 1) {
    const int mid = (a + e) >> 1;
    if (sent_start[mid] <= pos) a = mid; else e = mid;
  }
  const unsigned long long sg = sig[a];
  sa_rec[i] = make_uint2((unsigned)sg, (unsigned)(sg >> 32));
  if (((unsigned)sg & 63u) != 63u) {  // short sentence: length and second signature
    const uint32_t* s2 = sig2 + (size_t)a * kSig2Words;
    sa_aux[2 * i] = make_int4(sent_start[a], (int)((unsigned)sg & 63u), (int)s2[0], (int)s2[1]);
    sa_aux[2 * i + 1] = make_int4((int)s2[2], (int)s2[3], (int)s2[4], (int)s2[5]);
  } else {  // long sentence: length | 1 << 31, row of the wide signature
    sa_aux[2 * i] = make_int4(sent_start[a], (int)((((unsigned)sg >> 6) & 1023u) | 0x80000000u), (int)(unsigned)(sg >> 32), 0);
    sa_aux[2 * i + 1] = make_int4(0, 0, 0, 0);
  }
}

// sa_next[i] = token at depth 3 of suffix i (0 if it has fewer than four tokens)
__global__ void fm_build_next_kernel(const int32_t* __restrict__ tok, const int32_t* __restrict__ sa_pos, long long n_suf,
                                     int32_t* sa_next) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_suf) return;
  const int p = sa_pos[i];
  sa_next[i] = (tok[p + 1] != 0 && tok[p + 2] != 0) ? tok[p + 3] : 0;
}

// counts[0] = distinct bigrams, counts[1] = distinct trigrams (run starts in the suffix array), counts[2] = distinct
// 4-grams whose trigram occurs more than once (the entries of the 4-gram directory)
__global__ void fm_count_runs_kernel(const int32_t* __restrict__ tok, const int32_t* __restrict__ sa_pos, long long n_suf,
                                     unsigned long long* counts) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int bg = 0, tg = 0, qg = 0;
  if (i < n_suf) {
    const int p = sa_pos[i];
    const int t0 = tok[p], t1 = tok[p + 1];
    if (t1 != 0) {
      const int t2 = tok[p + 2];
      int q0 = -1, q1 = -1, q2 = -1, q3 = -1;
      if (i > 0) { const int pp = sa_pos[i - 1]; q0 = tok[pp]; q1 = tok[pp + 1]; q2 = q1 ? tok[pp + 2] : 0; q3 = q2 ? tok[pp + 3] : 0; }
      bg = q0 != t0 || q1 != t1;
      tg = t2 != 0 && (bg || q2 != t2);
      if (t2 != 0 && tok[p + 3] != 0) {
        const bool same_prev = !bg && q2 == t2;  // the previous suffix shares the trigram
        bool same_next = false;
        if (i + 1 < n_suf) { const int pn = sa_pos[i + 1]; same_next = tok[pn] == t0 && tok[pn + 1] == t1 && tok[pn + 2] == t2; }
        qg = (same_prev || same_next) && !(same_prev && q3 == tok[p + 3]);
      }
    }
  }
  const unsigned b1 = __ballot_sync(0xffffffffu, bg), b2 = __ballot_sync(0xffffffffu, tg), b3 = __ballot_sync(0xffffffffu, qg);
  if ((threadIdx.x & 31) == 0) {
    if (b1) atomicAdd(&counts[0], (unsigned long long)__popc(b1));
    if (b2) atomicAdd(&counts[1], (unsigned long long)__popc(b2));
    if (b3) atomicAdd(&counts[2], (unsigned long long)__popc(b3));
  }
}

__device__ __forceinline__ uint32_t dir_insert(int4* tab, uint32_t mask, int k0, int k1) {
  const unsigned long long key = ((unsigned long long)(unsigned)k1 << 32) | (unsigned)k0;  // (x, y) little-endian
  uint32_t h = bigram_hash(k0, k1) & mask;
  for (;;) {
    const unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(tab + h), ~0ull, key);
    if (prev == ~0ull || prev == key) return h;
    h = (h + 1) & mask;
  }
}
__device__ __forceinline__ uint32_t dir_find(const int4* tab, uint32_t mask, int k0, int k1) {
  uint32_t h = bigram_hash(k0, k1) & mask;
  while (!(tab[h].x == k0 && tab[h].y == k1)) h = (h + 1) & mask;
  return h;
}

// pass 0: run starts insert their key and write lo; pass 1: run ends write hi (keys are all present)
__global__ void fm_build_bigram_kernel(const int32_t* __restrict__ tok, const int32_t* __restrict__ sa_pos, long long n_suf,
                                       int4* bg_tab, uint32_t bg_mask, int pass) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_suf) return;
  const int p = sa_pos[i];
  const int t0 = tok[p], t1 = tok[p + 1];
  if (t1 == 0) return;
  if (pass == 0) {
    bool start = i == 0;
    if (!start) { const int pp = sa_pos[i - 1]; start = tok[pp] != t0 || tok[pp + 1] != t1; }
    if (start) bg_tab[dir_insert(bg_tab, bg_mask, t0, t1)].z = (int)i;
  } else {
    bool end = i == n_suf - 1;
    if (!end) { const int pn = sa_pos[i + 1]; end = tok[pn] != t0 || tok[pn + 1] != t1; }
    if (end) bg_tab[dir_find(bg_tab, bg_mask, t0, t1)].w = (int)(i + 1);
  }
}
// trigram directory: 32-byte entries keyed by the three words; the slot is claimed in two steps (words 0-1 with a
// 64-bit CAS, then word 2: contenders for the second step agree on the first)
__device__ __forceinline__ uint32_t tg_insert(int4* tab, uint32_t mask, int t0, int t1, int t2) {
  const unsigned long long key = ((unsigned long long)(unsigned)t1 << 32) | (unsigned)t0;
  uint32_t h = trigram_hash(t0, t1, t2) & mask;
  for (;;) {
    const unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(tab + 2 * (size_t)h), ~0ull, key);
    if (prev == ~0ull || prev == key) {
      const int prev2 = atomicCAS(&tab[2 * (size_t)h].z, -1, t2);
      if (prev2 == -1 || prev2 == t2) return h;
    }
    h = (h + 1) & mask;
  }
}
__device__ __forceinline__ uint32_t tg_find(const int4* tab, uint32_t mask, int t0, int t1, int t2) {
  uint32_t h = trigram_hash(t0, t1, t2) & mask;
  for (;;) {
    const int4 e = tab[2 * (size_t)h];
    if (e.x == t0 && e.y == t1 && e.z == t2) return h;
    h = (h + 1) & mask;
  }
}
__global__ void fm_build_trigram_kernel(const int32_t* __restrict__ tok, const int32_t* __restrict__ sa_pos, long long n_suf,
                                        int4* tg_tab, uint32_t tg_mask, int pass) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_suf) return;
  const int p = sa_pos[i];
  const int t0 = tok[p], t1 = tok[p + 1];
  if (t1 == 0) return;
  const int t2 = tok[p + 2];
  if (t2 == 0) return;
  const long long j = pass == 0 ? i - 1 : i + 1;
  bool edge = j < 0 || j >= n_suf;
  if (!edge) { const int pj = sa_pos[j]; edge = tok[pj] != t0 || tok[pj + 1] != t1 || tok[pj + 2] != t2; }
  if (!edge) return;
  if (pass == 0) {
    tg_tab[2 * (size_t)tg_insert(tg_tab, tg_mask, t0, t1, t2)].w = (int)i;
  } else {
    // hi, or for a trigram that occurs once -(position) - 1: the search then follows that suffix
    // without reading sa_pos
    bool single = i == 0;
    if (!single) { const int pj = sa_pos[i - 1]; single = tok[pj] != t0 || tok[pj + 1] != t1 || tok[pj + 2] != t2; }
    tg_tab[2 * (size_t)tg_find(tg_tab, tg_mask, t0, t1, t2) + 1].x = single ? -p - 1 : (int)(i + 1);
  }
}
// 4-gram directory: (trigram slot, word3) -> [lo, hi), or (lo, -position-1) for a 4-gram that occurs once. Only
// for trigrams that occur more than once (the others carry their position in the trigram directory): it replaces
// the one narrowing step of the search where ranges are still wide by one probe.
__global__ void fm_build_quadgram_kernel(const int32_t* __restrict__ tok, const int32_t* __restrict__ sa_pos, long long n_suf,
                                         const int4* __restrict__ tg_tab, uint32_t tg_mask, int4* qg_tab, uint32_t qg_mask,
                                         int pass) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_suf) return;
  const int p = sa_pos[i];
  const int t0 = tok[p], t1 = tok[p + 1];
  if (t1 == 0) return;
  const int t2 = tok[p + 2];
  if (t2 == 0) return;
  const int t3 = tok[p + 3];
  if (t3 == 0) return;
  bool same_prev = false, same_next = false, eq_prev = false, eq_next = false;  // trigram / 4-gram shared with the neighbours
  if (i > 0) {
    const int pj = sa_pos[i - 1];
    same_prev = tok[pj] == t0 && tok[pj + 1] == t1 && tok[pj + 2] == t2;
    eq_prev = same_prev && tok[pj + 3] == t3;
  }
  if (i + 1 < n_suf) {
    const int pj = sa_pos[i + 1];
    same_next = tok[pj] == t0 && tok[pj + 1] == t1 && tok[pj + 2] == t2;
    eq_next = same_next && tok[pj + 3] == t3;
  }
  if (!same_prev && !same_next) return;  // the trigram occurs once
  if (pass == 0 ? eq_prev : eq_next) return;  // not a run start / run end
  const int ts = (int)tg_find(tg_tab, tg_mask, t0, t1, t2);
  if (pass == 0) qg_tab[dir_insert(qg_tab, qg_mask, ts, t3)].z = (int)i;
  else qg_tab[dir_find(qg_tab, qg_mask, ts, t3)].w = !eq_prev ? -p - 1 : (int)(i + 1);
}

template <class T>
static int dev_alloc(Index* ix, int blk, size_t count, int fill_byte, const T** out) {
  void* d = nullptr;
  const size_t n = count * sizeof(T);
  FM_CUDA(cudaMalloc(&d, n ? n : 16));
  FM_CUDA(cudaMemset(d, fill_byte, n ? n : 16));
  ix->d_blocks[blk] = d;
  ix->blk_bytes[blk] = n;
  ix->device_bytes += (int64_t)n;
  *out = static_cast<const T*>(d);
  return FM_OK;
}

static int derive_next(Index* ix) {
  const int32_t* next = nullptr;
  int rc = dev_alloc(ix, BLK_NEXT, (size_t)ix->n_suf + 4, 0, &next);
  if (rc || ix->n_suf == 0) return rc;
  fm_build_next_kernel<<<(unsigned)((ix->n_suf + 255) / 256), 256>>>(static_cast<const int32_t*>(ix->d_blocks[BLK_TOK]),
                                                                      static_cast<const int32_t*>(ix->d_blocks[BLK_SA]), ix->n_suf,
                                                                      const_cast<int32_t*>(next));
  FM_CUDA(cudaDeviceSynchronize());
  return FM_OK;
}

// Builds sa_rec, sa_aux, sa_next, sid_at and the two directories on the device from tok + sa_pos (already uploaded).
static int build_on_device(Index* ix, const std::vector<int32_t>& sent_start, const std::vector<int32_t>& wide_row, int64_t n_wide) {
  IndexDev& d = ix->dev;
  const long long n_suf = ix->n_suf;
  const int n_sent = (int)ix->n_sent;
  int32_t* d_start = nullptr; int32_t* d_len = nullptr; unsigned long long* d_sig = nullptr; unsigned long long* d_counts = nullptr;
  int32_t* d_wrow = nullptr;
  uint32_t* d_sig2 = nullptr;
  FM_CUDA(cudaMalloc((void**)&d_sig2, (size_t)(n_sent + 1) * kSig2Words * 4));
  FM_CUDA(cudaMalloc((void**)&d_start, (size_t)(n_sent + 1) * 4));
  FM_CUDA(cudaMalloc((void**)&d_wrow, (size_t)(n_sent + 1) * 4));
  if (n_sent > 0) FM_CUDA(cudaMemcpy(d_wrow, wide_row.data(), (size_t)n_sent * 4, cudaMemcpyHostToDevice));
  FM_CUDA(cudaMalloc((void**)&d_len, (size_t)(n_sent + 1) * 4));
  FM_CUDA(cudaMalloc((void**)&d_sig, (size_t)(n_sent + 1) * 8));
  FM_CUDA(cudaMalloc((void**)&d_counts, 32));
  FM_CUDA(cudaMemset(d_counts, 0, 32));
  FM_CUDA(cudaMemcpy(d_start, sent_start.data(), (size_t)(n_sent + 1) * 4, cudaMemcpyHostToDevice));
  int rc;
  if ((rc = dev_alloc(ix, BLK_SID, (size_t)(ix->n_buf / 4) + 1, 0xff, &d.sid_at)) ||
      (rc = dev_alloc(ix, BLK_WALK, (size_t)n_suf + 8, 0, &d.sa_rec)) || (rc = dev_alloc(ix, BLK_START, 2 * ((size_t)n_suf + 4), 0, &d.sa_aux)) ||
      (rc = derive_next(ix)) ||
      (rc = dev_alloc(ix, BLK_WSIG, (size_t)n_wide * kWideWords, 0, &d.wsig)))
    return rc;
  d.n_wide = (int32_t)n_wide;
  d.sa_next = static_cast<const int32_t*>(ix->d_blocks[BLK_NEXT]);
  const int tb = 256;
  const unsigned gs = (unsigned)((n_suf + tb - 1) / tb);
  if (n_sent > 0)
    fm_build_sentence_kernel<<<(n_sent + tb - 1) / tb, tb>>>(d.tok, d_start, n_sent, d_wrow, const_cast<uint32_t*>(d.wsig), d_sig, d_len,
                                                             const_cast<int32_t*>(d.sid_at), d_sig2);
  unsigned long long counts[4] = {0, 0, 0, 0};
  if (n_suf > 0) {
    fm_build_walk_kernel<<<gs, tb>>>(d.sa_pos, n_suf, d_start, n_sent, d_sig, d_sig2, const_cast<uint2*>(d.sa_rec), const_cast<int4*>(d.sa_aux));
    fm_count_runs_kernel<<<gs, tb>>>(d.tok, d.sa_pos, n_suf, d_counts);
  }
  FM_CUDA(cudaMemcpy(counts, d_counts, 32, cudaMemcpyDeviceToHost));
  uint64_t cap_bg = 1024, cap_tg = 1024, cap_qg = 1024;
  while (cap_bg < counts[0] * 2) cap_bg <<= 1;
  while (cap_tg < counts[1] * 2) cap_tg <<= 1;
  while (cap_qg < counts[2] * 2) cap_qg <<= 1;
  d.bg_mask = (uint32_t)(cap_bg - 1);
  d.tg_mask = (uint32_t)(cap_tg - 1);
  d.qg_mask = (uint32_t)(cap_qg - 1);
  if ((rc = dev_alloc(ix, BLK_BG, (size_t)cap_bg, 0xff, &d.bg_tab)) || (rc = dev_alloc(ix, BLK_TG, 2 * (size_t)cap_tg, 0xff, &d.tg_tab)) ||
      (rc = dev_alloc(ix, BLK_QG, (size_t)cap_qg, 0xff, &d.qg_tab)))
    return rc;
  if (n_suf > 0) {
    fm_build_bigram_kernel<<<gs, tb>>>(d.tok, d.sa_pos, n_suf, const_cast<int4*>(d.bg_tab), d.bg_mask, 0);
    fm_build_bigram_kernel<<<gs, tb>>>(d.tok, d.sa_pos, n_suf, const_cast<int4*>(d.bg_tab), d.bg_mask, 1);
    fm_build_trigram_kernel<<<gs, tb>>>(d.tok, d.sa_pos, n_suf, const_cast<int4*>(d.tg_tab), d.tg_mask, 0);
    fm_build_trigram_kernel<<<gs, tb>>>(d.tok, d.sa_pos, n_suf, const_cast<int4*>(d.tg_tab), d.tg_mask, 1);
    fm_build_quadgram_kernel<<<gs, tb>>>(d.tok, d.sa_pos, n_suf, d.tg_tab, d.tg_mask, const_cast<int4*>(d.qg_tab), d.qg_mask, 0);
    fm_build_quadgram_kernel<<<gs, tb>>>(d.tok, d.sa_pos, n_suf, d.tg_tab, d.tg_mask, const_cast<int4*>(d.qg_tab), d.qg_mask, 1);
  }
  FM_CUDA(cudaDeviceSynchronize());
  FM_CUDA(cudaGetLastError());
  cudaFree(d_start); cudaFree(d_len); cudaFree(d_sig); cudaFree(d_counts); cudaFree(d_wrow); cudaFree(d_sig2);
  return FM_OK;
}

// Real tokens and penalty tokens of the TM (FuzzyMatch::add_tm(id, Sentence, Tokens), reference
// include/fuzzy/fuzzy_match.hh:53, include/fuzzy/sentence.hh:24-48) laid out parallel to the token buffer.
int set_real(Index* ix, const int32_t* real, const int32_t* gaps, const int64_t* sent_off, int64_t n_sent) {
  for (int64_t k = 0; k < ix->n_sent; k++)
    if (ix->kept[k] >= n_sent) { set_error("sent_off does not match the CSR the index was built from"); return FM_ERR_INVALID; }
  std::vector<int32_t> h_real((size_t)ix->n_buf, 0), h_gap((size_t)ix->n_buf, 0);
  ix->max_gap_id = 0;
  for (int64_t k = 0; k < ix->n_sent; k++) {
    const int64_t s = ix->kept[k], st = ix->h_sent_start[k];
    const int64_t n = sent_off[s + 1] - sent_off[s];
    int64_t stored = 0;
    while (ix->h_tok[(size_t)(st + stored)] != 0) stored++;
    if (n != stored) { set_error("sent_off does not match the CSR the index was built from (sentence length differs)"); return FM_ERR_INVALID; }
    for (int64_t j = 0; j < n; j++) h_real[st + j] = real[sent_off[s] + j];
    for (int64_t j = 0; j <= n; j++) {
      const int32_t g = gaps[sent_off[s] + s + j];
      if (g < 0 || g >= 2048) { set_error("penalty-token id outside [0, 2048)"); return FM_ERR_INVALID; }
      if (g > ix->max_gap_id) ix->max_gap_id = g;
      h_gap[st + j] = g;
    }
  }
  for (int blk : {BLK_REAL, BLK_GAP})
    if (ix->d_blocks[blk]) { cudaFree(ix->d_blocks[blk]); ix->device_bytes -= (int64_t)ix->blk_bytes[blk]; ix->d_blocks[blk] = nullptr; }
  int rc;
  if ((rc = upload(h_real, 0, ix, BLK_REAL, &ix->dev.real)) || (rc = upload(h_gap, 0, ix, BLK_GAP, &ix->dev.gap))) return rc;
  return FM_OK;
}

void free_index(Index* ix) {
  if (!ix) return;
  cudaSetDevice(ix->device);
  for (void* p : ix->d_blocks)
    if (p) cudaFree(p);
  cudaFree(ix->d_sent_start);
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
  {
    std::vector<int64_t> stamp((size_t)vocab_size, -1);
    int64_t cur = 0, k = 0;
    for (int64_t s = 0; s < n_in; s++) {
      const int64_t len = sent_off[s + 1] - sent_off[s];
      if (!(len > 0 && len <= max_tokens)) continue;
      ix->h_sent_start[k] = (int32_t)cur;
      ix->kept[k] = s;
      for (int64_t i = 0; i < len; i++) {
        const int32_t t = tokens[sent_off[s] + i];
        if (t < 2 || t >= vocab_size) {
          delete ix;
          set_error("TM token id outside [2, vocab_size) in sentence " + std::to_string(s));
          return FM_ERR_INVALID;
        }
        ix->h_tok[cur + i] = t;
        if (stamp[t] != k) { stamp[t] = k; ix->sfreq[t]++; }
      }
      cur += (len + 1 + 3) & ~int64_t(3);
      k++;
    }
    ix->h_sent_start[n_keep] = (int32_t)cur;
  }

  PhaseTimer pt;
  pt.lap("token buffer + sfreq");
  // ---- first-word bucket table (reference _quickVocabAccess, src/suffix_array.cc:82-98)
  std::vector<int32_t> qva((size_t)vocab_size + 1, 0);
  std::vector<int32_t> compact_off((size_t)n_keep + 1, 0);  // suffixes before each sentence
  int max_len = 0;
  {
    std::vector<int64_t> cnt((size_t)vocab_size + 1, 0);
    for (int64_t k = 0; k < n_keep; k++) {
      int32_t len = 0;
      for (int32_t pos = ix->h_sent_start[k]; ix->h_tok[pos] != 0; pos++, len++) cnt[ix->h_tok[pos] + 1]++;
      compact_off[k + 1] = compact_off[k] + len;
      max_len = std::max(max_len, (int)len);
    }
    for (int32_t w = 0; w < vocab_size; w++) cnt[w + 1] += cnt[w];
    for (int32_t w = 0; w <= vocab_size; w++) qva[w] = (int32_t)cnt[w];
  }
  // ---- sentences that get a wide signature (kWideMin) and their row numbers
  std::vector<int32_t> wide_row((size_t)n_keep, -1);
  int64_t n_wide = 0;
  for (int64_t k = 0; k < n_keep; k++)
    if (compact_off[k + 1] - compact_off[k] > kWideMin) wide_row[k] = (int32_t)n_wide++;
  // ---- suffix sort: on the GPU (fm_sort.cu); FM_HOST_SORT=1 keeps the host-thread sort for cross-checks
  const bool host_sort = getenv("FM_HOST_SORT") != nullptr;
  std::vector<int32_t> sa;
  if (host_sort) {
    sa.resize((size_t)n_suf);
    {
      std::vector<int64_t> cnt(qva.begin(), qva.end());
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
  }
  pt.lap("host tables (+ host sort)");
  // ---- upload
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) { delete ix; return cuda_fail(e, "cudaSetDevice"); }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ix->sm_count = prop.multiProcessorCount;
  int rc;
  IndexDev& d = ix->dev;
  rc = upload(ix->h_tok, 0, ix, BLK_TOK, &d.tok);
  pt.lap("upload tokens");
  if (!rc) rc = host_sort ? upload(sa, 4, ix, BLK_SA, &d.sa_pos) : dev_alloc(ix, BLK_SA, (size_t)n_suf + 4, 0, &d.sa_pos);
  if (!rc && !host_sort)
    rc = gpu_suffix_sort(d.tok, n_buf, ix->h_sent_start, compact_off, n_suf, max_len, vocab_size, ix->sm_count,
                         const_cast<int32_t*>(d.sa_pos));
  pt.lap("suffix sort (device)");
  if (rc ||
      (rc = upload(qva, 0, ix, BLK_QVA, &d.qva)) ||
      (rc = build_on_device(ix, ix->h_sent_start, wide_row, n_wide)) ||
      (rc = upload(std::vector<float>((size_t)vocab_size, 0.f), 0, ix, BLK_IDF, &d.idf)) ||
      (rc = set_idf_stats(ix, sfreq_global ? sfreq_global : ix->sfreq.data(), n_sent_global > 0 ? n_sent_global : n_keep))) {
    free_index(ix);
    return rc;
  }
  pt.lap("walk records + directories");
  d.vocab_size = vocab_size;
  d.max_tokens = max_tokens;
  d.n_suf = n_suf;
  d.n_buf = (int32_t)ix->n_buf;
  d.sid_base = (uint32_t)s_id_base;
  *out = ix;
  return FM_OK;
}

}  // namespace fm
