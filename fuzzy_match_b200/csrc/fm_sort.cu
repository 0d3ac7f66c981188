// fm_sort.cu -- GPU construction of the sentence-bounded suffix array.
//
// Replaces SuffixArray::sort (reference src/suffix_array.cc:58-102: bucket by first word, std::sort of
// every bucket with a token-wise comparator, :214-251) by prefix doubling on the device: ranks of the
// first h tokens of every suffix are combined pairwise, rank_2h(i) = order of (rank_h(i), rank_h(i+h)),
// each round one hand-written LSD radix sort of 64-bit keys. The separator behind sentence s has the rank s,
// below every word: a suffix that ends sorts before every longer suffix with the same prefix, and suffixes
// with identical content are ordered by sentence id -- exactly the reference's comparator (shorter first,
// ties by sentence id). match() cannot observe the order inside a range, subsequence() can: it walks ranges
// in this order and stops after number_of_matches candidates (src/fuzzy_match.cc:308-309).
//
// Separators are unique, so every suffix is a unique string up to and including its separator: what
// rank_h(i+h) reads from beyond a separator never decides a comparison.
#include <algorithm>
#include <vector>

#include "fm_internal.h"

namespace fm {

#define FULL 0xffffffffu
static const int kTile = 2048;  // keys per CTA and pass (256 threads, 8 warps x 256 keys)

// ---- radix sort pass: per-CTA digit histogram -> exclusive scan over (digit, CTA) -> stable scatter

__global__ void __launch_bounds__(256) fm_radix_hist_kernel(const unsigned long long* __restrict__ keys, long long n, int shift,
                                                            int32_t* hist, int n_cta) {
  __shared__ int s_hist[256];
  s_hist[threadIdx.x] = 0;
  __syncthreads();
  const long long beg = (long long)blockIdx.x * kTile;
  for (int k = threadIdx.x; k < kTile; k += 256) {
    const long long i = beg + k;
    if (i < n) atomicAdd(&s_hist[(int)((keys[i] >> shift) & 255)], 1);
  }
  __syncthreads();
  hist[(long long)threadIdx.x * n_cta + blockIdx.x] = s_hist[threadIdx.x];
}

__global__ void __launch_bounds__(256) fm_radix_scatter_kernel(const unsigned long long* __restrict__ kin,
                                                               const uint32_t* __restrict__ vin, unsigned long long* kout,
                                                               uint32_t* vout, long long n, int shift,
                                                               const int32_t* __restrict__ scanned, int n_cta) {
  __shared__ int s_off[8][256];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  for (int k = 0; k < 8; k++) s_off[k][t] = 0;
  __syncthreads();
  const long long wbeg = (long long)blockIdx.x * kTile + w * 256;  // every warp owns 256 consecutive keys
  for (int it = 0; it < 8; it++) {
    const long long i = wbeg + it * 32 + lane;
    if (i < n) atomicAdd(&s_off[w][(int)((kin[i] >> shift) & 255)], 1);
  }
  __syncthreads();
  {  // thread t owns digit t: global base of this CTA, then prefix over its warps
    int base = scanned[(long long)t * n_cta + blockIdx.x];
    for (int k = 0; k < 8; k++) {
      const int c = s_off[k][t];
      s_off[k][t] = base;
      base += c;
    }
  }
  __syncthreads();
  for (int it = 0; it < 8; it++) {
    const long long i = wbeg + it * 32 + lane;
    const bool valid = i < n;
    const unsigned long long key = valid ? kin[i] : 0;
    const int d = valid ? (int)((key >> shift) & 255) : 256 + lane;
    const unsigned m = __match_any_sync(FULL, d);
    const int rank = __popc(m & ((1u << lane) - 1));
    if (valid) {
      const int dst = s_off[w][d] + rank;
      kout[dst] = key;
      vout[dst] = vin[i];
    }
    __syncwarp();
    if (valid && rank == 0) s_off[w][d] += __popc(m);
    __syncwarp();
  }
}

// ---- prefix doubling

__global__ void fm_sa_init_kernel(const int32_t* __restrict__ tok, const int32_t* __restrict__ sent_start,
                                  const int32_t* __restrict__ coff, int n_sent, uint32_t* sa, int32_t* rank) {
  // one warp per sentence: list its suffix positions, seed the ranks with the word ids
  const int lane = threadIdx.x & 31;
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (s >= n_sent) return;
  const int st = sent_start[s], len = coff[s + 1] - coff[s], c0 = coff[s];
  for (int j = lane; j < len; j += 32) {
    sa[c0 + j] = (uint32_t)(st + j);
    rank[st + j] = n_sent + tok[st + j];
  }
  if (lane == 0) rank[st + len] = s;  // the separator: below every word, sentences in order
}
__global__ void fm_sa_keys_kernel(const uint32_t* __restrict__ sa, const int32_t* __restrict__ rank, long long n, int h,
                                  unsigned long long* keys) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t p = sa[k];
  keys[k] = ((unsigned long long)(uint32_t)rank[p] << 32) | (uint32_t)rank[p + h];
}
__global__ void fm_sa_heads_kernel(const unsigned long long* __restrict__ keys, long long n, int32_t* head) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  head[k] = (k == 0 || keys[k] != keys[k - 1]) ? 1 : 0;
}
__global__ void fm_sa_rank_kernel(const uint32_t* __restrict__ sa, const int32_t* __restrict__ head,
                                  const int32_t* __restrict__ excl, long long n, int n_sent, int32_t* rank) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  rank[sa[k]] = n_sent + excl[k] + head[k];  // group index above the separators' ranks 0 .. n_sent-1
}

static int bits_for(unsigned long long v) {
  int b = 1;
  while (b < 32 && (v >> b)) b++;
  return b;
}

#define SORT_CUDA(call)                                \
  do {                                                 \
    cudaError_t e__ = (call);                          \
    if (e__ != cudaSuccess) { rc = cuda_fail(e__, #call); goto done; } \
  } while (0)

// d_tok: padded token buffer on the device (n_buf entries); the suffix array (absolute offsets,
// sorted) is written to d_sa (n_suf entries). max_len = longest kept sentence.
int gpu_suffix_sort(const int32_t* d_tok, int64_t n_buf, const std::vector<int32_t>& sent_start,
                    const std::vector<int32_t>& compact_off, int64_t n_suf, int max_len, int32_t vocab_size, int sm_count,
                    int32_t* d_sa) {
  if (n_suf == 0) return FM_OK;
  const int n_sent = (int)sent_start.size() - 1;
  const int n_cta = (int)((n_suf + kTile - 1) / kTile);
  const long long n_hist = 256ll * n_cta;
  int rc = FM_OK;
  int32_t *d_start = nullptr, *d_coff = nullptr, *d_rank = nullptr, *d_head = nullptr, *d_excl = nullptr, *d_hist = nullptr,
          *d_hscan = nullptr;
  uint32_t *d_va = nullptr, *d_vb = nullptr;
  unsigned long long *d_ka = nullptr, *d_kb = nullptr, *d_chain = nullptr;
  unsigned epoch = 0;
  const int tb = 256;
  const unsigned gs = (unsigned)((n_suf + tb - 1) / tb);
  SORT_CUDA(cudaMalloc((void**)&d_start, (size_t)(n_sent + 1) * 4));
  SORT_CUDA(cudaMalloc((void**)&d_coff, (size_t)(n_sent + 1) * 4));
  SORT_CUDA(cudaMalloc((void**)&d_rank, (size_t)(n_buf + 1024) * 4));
  SORT_CUDA(cudaMalloc((void**)&d_head, (size_t)n_suf * 4));
  SORT_CUDA(cudaMalloc((void**)&d_excl, (size_t)(n_suf + 1) * 4));
  SORT_CUDA(cudaMalloc((void**)&d_hist, (size_t)n_hist * 4));
  SORT_CUDA(cudaMalloc((void**)&d_hscan, (size_t)(n_hist + 1) * 4));
  SORT_CUDA(cudaMalloc((void**)&d_va, (size_t)n_suf * 4));
  SORT_CUDA(cudaMalloc((void**)&d_vb, (size_t)n_suf * 4));
  SORT_CUDA(cudaMalloc((void**)&d_ka, (size_t)n_suf * 8));
  SORT_CUDA(cudaMalloc((void**)&d_kb, (size_t)n_suf * 8));
  SORT_CUDA(cudaMalloc((void**)&d_chain, 256 * 8));
  SORT_CUDA(cudaMemset(d_chain, 0, 256 * 8));
  SORT_CUDA(cudaMemset(d_rank, 0, (size_t)(n_buf + 1024) * 4));
  SORT_CUDA(cudaMemcpy(d_start, sent_start.data(), (size_t)(n_sent + 1) * 4, cudaMemcpyHostToDevice));
  SORT_CUDA(cudaMemcpy(d_coff, compact_off.data(), (size_t)(n_sent + 1) * 4, cudaMemcpyHostToDevice));
  fm_sa_init_kernel<<<(n_sent + 7) / 8, 256>>>(d_tok, d_start, d_coff, n_sent, d_va, d_rank);
  {
    unsigned long long max_rank = (unsigned long long)n_sent + (unsigned long long)std::max<int64_t>(vocab_size, 2);
    for (int h = 1; h < std::max(max_len + 1, 2); h <<= 1) {  // until the ranks cover the longest sentence and its separator
      fm_sa_keys_kernel<<<gs, tb>>>(d_va, d_rank, n_suf, h, d_ka);
      // LSD passes over the bits that can differ: low word (rank at i+h) then high word (rank at i)
      const int nb = bits_for(max_rank);
      for (int word = 0; word < 2; word++)
        for (int bit = 0; bit < nb; bit += 8) {
          const int shift = word * 32 + bit;
          fm_radix_hist_kernel<<<n_cta, 256>>>(d_ka, n_suf, shift, d_hist, n_cta);
          launch_scan(d_hist, d_hscan, (int32_t)n_hist, d_chain, ++epoch, sm_count, 0);
          fm_radix_scatter_kernel<<<n_cta, 256>>>(d_ka, d_va, d_kb, d_vb, n_suf, shift, d_hscan, n_cta);
          std::swap(d_ka, d_kb);
          std::swap(d_va, d_vb);
        }
      fm_sa_heads_kernel<<<gs, tb>>>(d_ka, n_suf, d_head);
      launch_scan(d_head, d_excl, (int32_t)n_suf, d_chain, ++epoch, sm_count, 0);
      fm_sa_rank_kernel<<<gs, tb>>>(d_va, d_head, d_excl, n_suf, n_sent, d_rank);
      max_rank = (unsigned long long)n_sent + (unsigned long long)n_suf + 1;
    }
  }
  SORT_CUDA(cudaMemcpy(d_sa, d_va, (size_t)n_suf * 4, cudaMemcpyDeviceToDevice));
  SORT_CUDA(cudaDeviceSynchronize());
  SORT_CUDA(cudaGetLastError());
done:
  cudaFree(d_start); cudaFree(d_coff); cudaFree(d_rank); cudaFree(d_head); cudaFree(d_excl); cudaFree(d_hist); cudaFree(d_hscan);
  cudaFree(d_va); cudaFree(d_vb); cudaFree(d_ka); cudaFree(d_kb); cudaFree(d_chain);
  return rc;
}

}  // namespace fm
