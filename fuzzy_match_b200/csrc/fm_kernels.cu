// fm_kernels.cu -- hand-written sm_100a kernels of the fuzzy-match hot path.
//
// One batch of patterns streams through eleven launches (no host round trip in between):
//   prepare      per query: clamp ml, sanitise ids, build the pattern's word table and signature planes (a thread
//                per query for short patterns, a warp per query for the rest: two launches)
//   search       per (query, start position): trigram directory probe (bigram first when min_subseq_length < 3),
//                4-gram directory, then narrow the suffix-array range token by token; emit range slices
//   gather       walk: per suffix-array element of every slice, length window + signature bound from its 8-byte
//                record (four per 256-bit load); verify: second signature, exact coverage for the few that
//                pass, dedup (query, sentence) with max match length (two launches)
//                                                                      <- the "suffix-range gather"
//   scan         exclusive scan of survivors per query (co-resident CTAs, epoch-tagged tile totals)
//   score        per surviving (query, sentence): edit-distance DP -- registers for p <= 32 (thread per
//                pair), warp-wide wavefront in shared memory above     <- the "DP kernel"
//   replay       per query: the reference's sequential bound heap / top-N over the scored candidates
//                (thread per query for 0-1 candidates, warp per query, CTA per query for long lists)
//   (+ bounds when the parameters change: per pattern length tables of the two rejection bounds;
//    + contrast: per query warp, contrastive rerank)
//
// Integer indexing and scalar fp32 only -- no tensor cores. All float arithmetic that reaches a
// result is written with __fadd_rn/__fmul_rn/__fdiv_rn in the reference's evaluation order so no
// FMA contraction or reassociation can change a bit (the file is also compiled with -fmad=false).
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cstdlib>

#include "fm_internal.h"

namespace fm {

#define FULL 0xffffffffu

// ---------------------------------------------------------------- small device helpers

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x = ((x >> 16) ^ x) * 0x45d9f3bu;
  x = ((x >> 16) ^ x) * 0x45d9f3bu;
  return (x >> 16) ^ x;
}
__device__ __forceinline__ uint32_t hash64(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return (uint32_t)k;
}
__device__ __forceinline__ int next_pow2(int v) {  // smallest power of two >= v (v >= 1)
  return v <= 1 ? 1 : 1 << (32 - __clz(v - 1));
}
__device__ __forceinline__ int4 ldg_nc_v4(const int4* p) {
  int4 r;
  asm volatile("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// Costs::get_normalizer (reference include/fuzzy/costs.hh:33-47)
__device__ __forceinline__ float normalizer(int p, int s, const Params& pr) {
  if (pr.ins == 0.f && pr.del == 0.f && pr.rep == 0.f) return 1.f;
  const float fp = (float)p, fs = (float)s;
  if (__fadd_rn(pr.ins, pr.del) <= pr.rep) return __fadd_rn(__fmul_rn(pr.ins, fp), __fmul_rn(pr.del, fs));
  if (p <= s) return __fadd_rn(__fmul_rn(__fsub_rn(pr.rep, pr.del), fp), __fmul_rn(pr.del, fs));
  return __fadd_rn(__fmul_rn(__fsub_rn(pr.rep, pr.ins), fs), __fmul_rn(pr.ins, fp));
}
// NGramMatches::theoretical_rejection (reference src/ngram_matches.cc:32-39)
__device__ __forceinline__ bool reject_length(int p, int s, const Params& pr) {
  const float diff = fabsf(__fsub_rn((float)p, (float)s));
  const float rc = (p >= s) ? pr.ins : pr.del;
  const float bound = __fsub_rn(1.f, __fdiv_rn(__fmul_rn(rc, diff), normalizer(p, s, pr)));
  return (double)bound + 0.000005 < (double)pr.fuzzy;
}
// NGramMatches::theoretical_rejection_cover (reference src/ngram_matches.cc:42-59)
__device__ __forceinline__ bool reject_cover(int p, int s, int cover, const Params& pr) {
  const float fp = (float)p, fs = (float)s, fc = (float)cover;
  float num;
  if (__fadd_rn(pr.ins, pr.del) < pr.rep) {
    num = __fadd_rn(__fmul_rn(pr.ins, __fsub_rn(fs, fc)), __fmul_rn(pr.del, __fsub_rn(fp, fc)));
  } else {
    const float rc = (p > s) ? pr.ins : pr.del;
    const float mn = (p > s) ? fs : fp;
    const float mx = (p > s) ? fp : fs;
    num = __fadd_rn(__fmul_rn(pr.rep, __fsub_rn(mn, fc)), __fmul_rn(rc, __fsub_rn(mx, mn)));
  }
  const float bound = __fsub_rn(1.f, __fdiv_rn(num, normalizer(p, s, pr)));
  return (double)bound + 0.000005 < (double)pr.fuzzy;
}
// score = int(10000 - cost*100) / 10000.0 narrowed to float (reference src/fuzzy_match.cc:598)
__device__ __forceinline__ float score_of(float cost) {
  const int v = (int)__fsub_rn(10000.f, __fmul_rn(cost, 100.f));
  return (float)((double)v / 10000.0);
}

// ---------------------------------------------------------------- bound tables (one warp per pattern length)

// The length bound (ngram_matches.cc:32-39) accepts a window [smin, smax] of sentence lengths around
// p; for each of them cmin = the smallest coverage the coverage bound (ngram_matches.cc:42-59) lets
// through. Both depend only on (p, s, fuzzy, costs), so they are evaluated with the exact float/double
// expressions once per pattern length when the parameters change, instead of once per suffix-array
// element. cmin_tab[(p << 10) | s] = smallest passing coverage, kNeedReject when (p, s) can never pass,
// kNeedNoTable when the bounds must be evaluated per element (negative costs: not monotone).
static const int kNeedReject = 0xffff, kNeedNoTable = 0xfffe;
// cmin64[(p << 6) | l] is the same table over the 6-bit length field of a walk record, in the form stage 1
// of the gather compares against: the smallest passing coverage, kNeedReject, or 0 = "goes to stage 2"
// for l = 63 (a long sentence: its real length and wide signature are looked at there) and when there is
// no table.
__global__ void __launch_bounds__(256) fm_bounds_kernel(int max_tokens, Params pr, uint16_t* cmin_tab, uint16_t* cmin64) {
  const int lane = threadIdx.x & 31;
  const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // row 0 exists for idle lanes of the walk
  if (p > max_tokens) return;
  const bool fast = pr.ins >= 0.f && pr.del >= 0.f && pr.rep >= 0.f;  // reject_cover is monotone in the coverage for costs >= 0
  uint16_t* row = cmin_tab + (p << 10);
  for (int sl = lane; sl < 1024; sl += 32) {
    int need = fast ? kNeedReject : kNeedNoTable;
    if (fast && sl >= 1 && sl <= max_tokens && !reject_length(p, sl, pr) && !reject_cover(p, sl, p, pr)) {
      int lo = 0, hi = p;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (!reject_cover(p, sl, mid, pr)) hi = mid; else lo = mid + 1;
      }
      need = lo;
    }
    row[sl] = (uint16_t)need;
    if (sl < 64) cmin64[(p << 6) | sl] = (uint16_t)(sl == 63 || need == kNeedNoTable ? 0 : (sl > kWideMin ? kNeedReject : need));
  }
}

// ---------------------------------------------------------------- prepare

// One warp per query. Guards and ml clamp of src/fuzzy_match.cc:450-467; ids outside the vocabulary
// become VOCAB_UNK (src/vocab_indexer.cc:52-60); the table is PatternCoverage's multiset
// (src/pattern_coverage.cc:8-13) as an open-addressing table: word -> (distinct index, multiplicity).
__device__ __forceinline__ void prepare_query(const IndexDev& ix, const BatchDev& b, const Params& pr, int q, int lane) {
  const int off = b.q_off[q];
  const int p = b.q_off[q + 1] - off;
  const bool valid = p > 0 && p <= ix.max_tokens;
  int ml = pr.ml;
  if (ml < 0 || ml > p) ml = p;  // (size_t)ml > pattern.size()
  const int by_ratio = (int)__fmul_rn(pr.mr, (float)p);
  if (by_ratio > ml) ml = by_ratio;
  if (lane == 0) {
    if (!valid) b.qmeta[q] = make_int4(0, ml, off, 0);
    b.q_cnt[q] = 0;
    if (q == 0) b.q_cnt[b.n_q] = 0;
  }
  if (p <= 0) return;
  for (int j = lane; j < p; j += 32) {
    const int t = b.q_tok_in[off + j];
    b.pat[off + j] = (t >= 2 && t < ix.vocab_size) ? t : 1;
    if (!valid) b.chain_rec[off + j] = make_int2(q, 0);  // (pattern length 0: the chain is dead)
  }
  if (!valid) return;
  const int ts = next_pow2(2 * p);
  int2* tbl = b.tbl + 4ll * off;
  __shared__ int2 s_tbl[8][128];
  __shared__ int s_cnt[8][64];
  __shared__ int s_cnt2[8][32 * kSig2Words];  // pattern positions per bit of the second signature
  __shared__ unsigned long long s_peq[8][64];
  __shared__ uint32_t s_wcnt[8][512];  // 1024 16-bit counters per warp (wide signature bits)
  // Lanes insert their words concurrently (CAS on the key, atomic add on the multiplicity) -- into a
  // shared-memory table for patterns of up to 64 words, straight into the query's table in global memory for
  // longer ones; distinct indices are then handed out in slot order.
  const bool small = ts <= 128;
  int2* wtbl = small ? s_tbl[threadIdx.x >> 5] : tbl;
  int* cnt = s_cnt[threadIdx.x >> 5];
  int* cnt2 = s_cnt2[threadIdx.x >> 5];
  for (int j = lane; j < ts; j += 32) wtbl[j] = make_int2(-1, 0);
  cnt[lane] = 0;
  cnt[lane + 32] = 0;
#pragma unroll
  for (int k = 0; k < kSig2Words; k++) cnt2[32 * k + lane] = 0;
  __syncwarp();
  for (int j = lane; j < p; j += 32) {
    const int w = b.pat[off + j];
    int h = hash32((uint32_t)w) & (ts - 1);
    for (;;) {
      const int prev = atomicCAS(&wtbl[h].x, -1, w);
      if (prev == -1 || prev == w) break;
      h = (h + 1) & (ts - 1);
    }
    atomicAdd(&wtbl[h].y, 1 << 16);
    if (w >= 2) {  // pattern positions per signature bit; unknown words excluded
      atomicAdd(&cnt[sig_bit(w)], 1);
      atomicAdd(&cnt2[sig2_bit(w)], 1);
    }
  }
  __syncwarp();
  int distinct = 0;
  for (int j0 = 0; j0 < ts; j0 += 32) {
    const int2 e = j0 + lane < ts ? wtbl[j0 + lane] : make_int2(-1, 0);
    const unsigned used = __ballot_sync(FULL, e.x != -1);
    if (e.x != -1) {
      const int d = distinct + __popc(used & ((1u << lane) - 1));
      wtbl[j0 + lane].y = e.y | d;
      if (small) tbl[j0 + lane] = make_int2(e.x, e.y | d);
    } else if (small && j0 + lane < ts) {
      tbl[j0 + lane] = e;
    }
    distinct += __popc(used);
  }
  __syncwarp();
  const int c_lo = cnt[lane], c_hi = cnt[lane + 32];  // pattern positions per signature bit (lane, lane + 32)
  // Position masks for the bit-parallel edit distance of patterns of 33..64 tokens (shorter ones compare
  // against the pattern in registers): peq64[off + d] has bit j set iff pattern[j] is the word with
  // distinct index d.
  if (small && p > 32) {
    unsigned long long* peq = s_peq[threadIdx.x >> 5];
    peq[lane] = 0;
    peq[lane + 32] = 0;
    __syncwarp();
    for (int j = lane; j < p; j += 32) {
      const int w = b.pat[off + j];
      int h = hash32((uint32_t)w) & (ts - 1);
      while (wtbl[h].x != w) h = (h + 1) & (ts - 1);
      atomicOr(&peq[wtbl[h].y & 0xffff], 1ull << j);
    }
    __syncwarp();
    for (int d = lane; d < distinct; d += 32) b.peq64[off + d] = peq[d];
  }
  // Signature masks in the layout of a walk record (bits 6..63), bit-sliced: plane k holds bit k of
  // min(count, 3), count = pattern positions on that signature bit, and mult = (largest count) - 3 (>= 0):
  //   coverage <= popc(sig & B0) + 2 popc(sig & B1) + mult * popc(sig & B0 & B1)
  // (exact while no bit collects more than three positions).
  {
    const int k_lo = min(c_lo, 3), k_hi = min(c_hi, 3);
    const unsigned a_lo = __ballot_sync(FULL, k_lo & 1), a_hi = __ballot_sync(FULL, k_hi & 1);
    const unsigned a2_lo = __ballot_sync(FULL, k_lo & 2), a2_hi = __ballot_sync(FULL, k_hi & 2);
    int mult = max(max(c_lo, c_hi) - 3, 0);
    mult = __reduce_max_sync(FULL, mult);
    // the same planes over the second signature's 192 bits (tested by the verify kernel before the exact count):
    // lane k keeps word k of plane B0 (k < 6) or word k - 6 of plane B1 and writes it
    int dmax = 0;
    unsigned mine = 0;
#pragma unroll
    for (int k = 0; k < kSig2Words; k++) {
      const int c2 = cnt2[32 * k + lane];
      const int j2 = min(c2, 3);
      dmax = max(dmax, c2);
      const unsigned p0 = __ballot_sync(FULL, j2 & 1), p1 = __ballot_sync(FULL, j2 & 2);
      if (lane == k) mine = p0;
      if (lane == k + kSig2Words) mine = p1;
    }
    int mult2 = max(dmax - 3, 0);
    mult2 = __reduce_max_sync(FULL, mult2);
    if (lane < 2 * kSig2Words) reinterpret_cast<unsigned*>(b.qmask2)[(size_t)q * (2 * kSig2Words) + lane] = mine;
    // what a chain of the search needs to start, in one coalesced 8-byte read: query, start position, pattern length
    // and the mult of the walk's signature (the tag of its slice records)
    for (int j = lane; j < p; j += 32) b.chain_rec[off + j] = make_int2(q, j | (p << 10) | (mult << 20));
    mult |= mult2 << 10;  // both travel in qmeta.w (10 bits each: a pattern has at most 1023 positions)
    // The same over the 1024 bits of the wide signatures (sentences longer than kWideMin), exactly: three planes
    // of min(count, 7) and a short list of the bits that collect more (frequent words of a long pattern), so that
    //   coverage <= sum over the signature's bits of count(bit)
    // is evaluated without slack. Lane l owns signature bits [32 l, 32 l + 32), i.e. words 16 l .. 16 l + 15 of
    // the packed counters.
    if (b.wq) {
      uint32_t* wc = s_wcnt[threadIdx.x >> 5];
      for (int k = lane; k < 512; k += 32) wc[k] = 0;
      __syncwarp();
      for (int j = lane; j < p; j += 32) {
        const int w = b.pat[off + j];
        if (w >= 2) {
          const unsigned bit = wsig_bit(w);
          atomicAdd(&wc[bit >> 1], 1u << (16 * (bit & 1)));
        }
      }
      __syncwarp();
      uint32_t* dst = b.wq + (size_t)q * kWideStride;
      if (lane < 16) dst[3 * kWideWords + lane] = 0;
      __syncwarp();
      unsigned w0 = 0, w1 = 0, w2 = 0;
      int n_big = 0, rest = 0;  // entries written so far (warp-uniform) / excess that found no entry (this lane)
#pragma unroll
      for (int k = 0; k < 16; k++) {
        const int c = (k + lane) & 15;  // rotated: lanes hit different banks
        const uint32_t v = wc[16 * lane + c];
        const int c0 = (int)(v & 0xffffu), c1 = (int)(v >> 16);
        const int k0 = min(c0, 7), k1 = min(c1, 7);
        w0 |= (unsigned)((k0 & 1) | ((k1 & 1) << 1)) << (2 * c);
        w1 |= (unsigned)(((k0 >> 1) & 1) | (((k1 >> 1) & 1) << 1)) << (2 * c);
        w2 |= (unsigned)(((k0 >> 2) & 1) | (((k1 >> 2) & 1) << 1)) << (2 * c);
        if (__any_sync(FULL, max(c0, c1) > 7)) {
#pragma unroll
          for (int half = 0; half < 2; half++) {
            const int cc = half ? c1 : c0;
            const unsigned bal = __ballot_sync(FULL, cc > 7);
            if (cc > 7) {
              const int slot = n_big + __popc(bal & ((1u << lane) - 1));
              if (slot < kWideBig) dst[3 * kWideWords + slot] = (uint32_t)(32 * lane + 2 * c + half) | ((uint32_t)(cc - 7) << 16);
              else rest += cc - 7;
            }
            n_big += __popc(bal);
          }
        }
      }
      dst[lane] = w0;
      dst[kWideWords + lane] = w1;
      dst[2 * kWideWords + lane] = w2;
      rest = __reduce_add_sync(FULL, rest);
      if (lane == 0) dst[3 * kWideWords + kWideBig] = (uint32_t)rest;
    }
    if (lane == 0) {
      b.qmask[q] = make_int4((int)a_lo, (int)a_hi, (int)a2_lo, (int)a2_hi);
      b.qmeta[q] = make_int4(p, ml, off, kQValid | (mult << 8));
    }
  }
}
// warp w prepares query w
__global__ void __launch_bounds__(256) fm_prepare_kernel(IndexDev ix, BatchDev b, Params pr) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid < b.n_q) prepare_query(ix, b, pr, wid, threadIdx.x & 31);
}
// the queries fm_prepare_short_kernel left over (patterns of more than kPrepShort words, invalid ones, a signature
// bit with more than three positions): b.prep_list, persistent warps
__global__ void __launch_bounds__(256, 4) fm_prepare_list_kernel(IndexDev ix, BatchDev b, Params pr) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n = (int)b.ctr->n_prep, n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = wid; i < n; i += n_warps) {
    prepare_query(ix, b, pr, b.prep_list[i], threadIdx.x & 31);
    __syncwarp();
  }
}

// The same for the common case, one THREAD per query: patterns of at most kPrepShort words on an index without wide
// signatures. A warp-per-query pass spends ~500 warp instructions on a 15-word pattern, most of them on
// warp-wide bookkeeping of a handful of active lanes; here a lane walks its own pattern: every word goes into the
// lane's table in shared memory (slot-major: a lane only ever touches its own bank) and into bit-sliced 3-bit
// counters of the two signatures (the walk's 58 bits in registers, the 192 bits of the second one in shared
// memory) -- a counter that would pass seven sends the query to the warp kernel, which keeps the exact
// multiplicities. The finished table is copied to the query's place in global memory. Distinct indices are handed
// out in insertion order (any numbering will do: they only name bits of the verify kernel's seen-mask).
static const int kPrepShort = 32;
static const int kPrepThreads = 64;
__global__ void __launch_bounds__(kPrepThreads) fm_prepare_short_kernel(IndexDev ix, BatchDev b, Params pr) {
  __shared__ int s_key[2 * kPrepShort][kPrepThreads];
  __shared__ int s_val[2 * kPrepShort][kPrepThreads];
  __shared__ unsigned s_p2[3 * kSig2Words][kPrepThreads];
  const int lane = threadIdx.x & 31, tid = threadIdx.x;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const bool inb = q < b.n_q;
  int off = 0, p = 0;
  if (inb) {
    off = b.q_off[q];
    p = b.q_off[q + 1] - off;
    b.q_cnt[q] = 0;
    if (q == 0) b.q_cnt[b.n_q] = 0;
  }
  const bool mine = inb && p > 0 && p <= kPrepShort && p <= ix.max_tokens;
  bool redo = inb && !mine;
  if (mine) {
    const int ts = next_pow2(2 * p);
    for (int k = 0; k < ts; k++) s_key[k][tid] = -1;
#pragma unroll
    for (int k = 0; k < 3 * kSig2Words; k++) s_p2[k][tid] = 0;
    unsigned long long a0 = 0, a1 = 0, a2 = 0;  // the walk signature's counters, bit-sliced: planes 0, 1, 2
    unsigned sat = 0;
    int distinct = 0;
    int t = b.q_tok_in[off];
    for (int j = 0; j < p; j++) {
      const int w = (t >= 2 && t < ix.vocab_size) ? t : 1;
      if (j + 1 < p) t = b.q_tok_in[off + j + 1];
      b.pat[off + j] = w;
      int h = hash32((uint32_t)w) & (ts - 1);
      for (;;) {
        const int k = s_key[h][tid];
        if (k == w) { s_val[h][tid] += 1 << 16; break; }
        if (k == -1) { s_key[h][tid] = w; s_val[h][tid] = (1 << 16) | distinct; distinct++; break; }
        h = (h + 1) & (ts - 1);
      }
      if (w >= 2) {  // unknown words take no part in the signatures
        const unsigned long long m = 1ull << sig_bit(w);
        const unsigned long long c0 = a0 & m, c1 = a1 & c0;  // carries into planes 1 and 2
        sat |= (unsigned)((a2 & c1) != 0);
        a0 ^= m;
        a1 ^= c0;
        a2 |= c1;
        const unsigned b2 = sig2_bit(w), k2 = b2 >> 5, m2 = 1u << (b2 & 31u);
        const unsigned x0 = s_p2[k2][tid], x1 = s_p2[kSig2Words + k2][tid], x2 = s_p2[2 * kSig2Words + k2][tid];
        const unsigned d0 = x0 & m2, d1 = x1 & d0;
        sat |= (unsigned)((x2 & d1) != 0);
        s_p2[k2][tid] = x0 ^ m2;
        s_p2[kSig2Words + k2][tid] = x1 ^ d0;
        if (d1) s_p2[2 * kSig2Words + k2][tid] = x2 | d1;
      }
    }
    if (sat) {
      redo = true;  // (the warp kernel builds the table as well)
    } else {
      // four entries -- one 32-byte sector -- per store (the table starts on a 32-byte boundary; ts >= 4 here or the
      // two entries of a one-word pattern go out on their own)
      int2* tbl = b.tbl + 4ll * off;
      if (ts < 4) {
        for (int k = 0; k < ts; k++) {
          const int key = s_key[k][tid];
          tbl[k] = make_int2(key, key == -1 ? 0 : s_val[k][tid]);
        }
      } else {
        for (int k = 0; k < ts; k += 4) {
          const int k0 = s_key[k][tid], k1 = s_key[k + 1][tid], k2 = s_key[k + 2][tid], k3 = s_key[k + 3][tid];
          const int v0 = k0 == -1 ? 0 : s_val[k][tid], v1 = k1 == -1 ? 0 : s_val[k + 1][tid];
          const int v2 = k2 == -1 ? 0 : s_val[k + 2][tid], v3 = k3 == -1 ? 0 : s_val[k + 3][tid];
          asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(tbl + k), "r"(k0), "r"(v0), "r"(k1), "r"(v1),
                       "r"(k2), "r"(v2), "r"(k3), "r"(v3)
                       : "memory");
        }
      }
      int ml = pr.ml;
      if (ml < 0 || ml > p) ml = p;  // (size_t)ml > pattern.size()
      const int by_ratio = (int)__fmul_rn(pr.mr, (float)p);
      if (by_ratio > ml) ml = by_ratio;
      // planes of min(count, 3) and mult = (largest count) - 3, like the warp kernel: a counter of 4..7 has plane 2 set
      int mult = 0;
      if (a2) mult = (a2 & a1 & a0) ? 4 : (a2 & a1) ? 3 : (a2 & a0) ? 2 : 1;
      const unsigned long long s0 = a0 | a2, s1 = a1 | a2;
      b.qmask[q] = make_int4((int)(unsigned)s0, (int)(unsigned)(s0 >> 32), (int)(unsigned)s1, (int)(unsigned)(s1 >> 32));
      unsigned y[2 * kSig2Words];
      int mult2 = 0;
#pragma unroll
      for (int k = 0; k < kSig2Words; k++) {
        const unsigned x0 = s_p2[k][tid], x1 = s_p2[kSig2Words + k][tid], x2 = s_p2[2 * kSig2Words + k][tid];
        y[k] = x0 | x2;
        y[kSig2Words + k] = x1 | x2;
        if (x2) mult2 = max(mult2, (x2 & x1 & x0) ? 4 : (x2 & x1) ? 3 : (x2 & x0) ? 2 : 1);
      }
      int4* mq = b.qmask2 + 3ll * q;
      mq[0] = make_int4((int)y[0], (int)y[1], (int)y[2], (int)y[3]);
      mq[1] = make_int4((int)y[4], (int)y[5], (int)y[6], (int)y[7]);
      mq[2] = make_int4((int)y[8], (int)y[9], (int)y[10], (int)y[11]);
      b.qmeta[q] = make_int4(p, ml, off, kQValid | (mult << 8) | (mult2 << 18));
      for (int j = 0; j < p; j++) b.chain_rec[off + j] = make_int2(q, j | (p << 10) | (mult << 20));
    }
  }
  const unsigned rb = __ballot_sync(FULL, redo);
  if (rb) {
    int base = 0;
    if (lane == 0) base = (int)atomicAdd(&b.ctr->n_prep, (unsigned)__popc(rb));
    base = __shfl_sync(FULL, base, 0);
    if (redo) b.prep_list[base + __popc(rb & ((1u << lane) - 1))] = q;
  }
}

__device__ __forceinline__ void ldg_nc_v8(const void* p, unsigned (&r)[8]) {  // one 256-bit load (32-byte aligned)
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Trigram directory lookup: one 32-byte entry (a sector) per probe. Returns false when no suffix starts with the
// trigram; pos1 = position of the only suffix when it occurs once (else -1), slot = the entry's slot (key of the
// 4-gram directory).
__device__ __forceinline__ bool tg_lookup(const IndexDev& ix, int t0, int t1, int t2, int& lo, int& hi, int& pos1, uint32_t& slot) {
  uint32_t h = trigram_hash(t0, t1, t2) & ix.tg_mask;
  for (;;) {
    unsigned e[8];
    ldg_nc_v8(ix.tg_tab + 2 * (size_t)h, e);
    if ((int)e[0] == t0 && (int)e[1] == t1 && (int)e[2] == t2) {
      lo = (int)e[3];
      hi = (int)e[4];
      pos1 = -1;
      if (hi < 0) { pos1 = -hi - 1; hi = lo + 1; }
      slot = h;
      return true;
    }
    if ((int)e[0] == -1) return false;
    h = (h + 1) & ix.tg_mask;
  }
}

// ---------------------------------------------------------------- search

// Range slices are first buffered per thread in shared memory (kSliceBuf slots) and appended to the
// global slice list by warp-collective flushes: one packed 64-bit atomic per warp and flush reserves
// slice slots and flattened element offsets together, so slice order == element order. Buffering keeps
// the number of same-address atomics near one per warp instead of one per warp and n-gram level.
static const int kSliceBuf = 6;
#ifndef FM_SEARCH_THREADS
#define FM_SEARCH_THREADS 128  // CTA size of the search kernel: the CTA-wide flush waits for its slowest warp (measured:
                               // 0.134 ms with 128 threads x 12 CTAs/SM, 0.143 ms with 256 x 6, 0.144 ms with 64 x 20)
#endif
struct SliceBuf {
  int beg[kSliceBuf][FM_SEARCH_THREADS];
  int sz[kSliceBuf][FM_SEARCH_THREADS];
  int lm[kSliceBuf][FM_SEARCH_THREADS];
};
__device__ __forceinline__ void note_spans(const BatchDev& b, long long slot, long long start, int size) {
  // span_slice[k] = slice that holds flattened element k*kSpan
  for (long long k = (start + kSpan - 1) / kSpan; k * kSpan < start + size; k++) {
    if (k < b.span_cap) b.span_slice[k] = (int32_t)slot;
    else { atomicOr(&b.ctr->overflow, 4u); break; }
  }
}
__device__ __forceinline__ void push_slice(SliceBuf& sb, int& nbuf, int beg, int sz, int lm) {
  sb.beg[nbuf][threadIdx.x] = beg;
  sb.sz[nbuf][threadIdx.x] = sz;
  sb.lm[nbuf][threadIdx.x] = lm;
  nbuf++;
}
// Slices of more than kSmallSlice elements go to the flattened list (the gather walks them with whole
// warps); the many tiny ones -- three quarters of all slices hold a single suffix -- go to their own
// list, where a lane takes a slice. `tag` = p << 10 | mult << 20 of the query.
__device__ __forceinline__ void flush_slices(const BatchDev& b, SliceBuf& sb, int& nbuf, int lane, int q, int tag) {
  int elems = 0, n_big = 0;
  for (int k = 0; k < nbuf; k++) {
    const int sz = sb.sz[k][threadIdx.x];
    if (sz > kSmallSlice) { elems += sz; n_big++; }
  }
  // one scan for three counts: small slices << 54 | big slices << kElemBits | elements of big slices
  const unsigned long long mine = ((unsigned long long)(nbuf - n_big) << 54) | ((unsigned long long)n_big << kElemBits) |
                                  (unsigned long long)(unsigned)elems;
  unsigned long long incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long o = __shfl_up_sync(FULL, incl, d);
    if (lane >= d) incl += o;
  }
  const unsigned long long total = __shfl_sync(FULL, incl, 31);
  if (total == 0) return;
  const unsigned long long low54 = (1ull << 54) - 1;
  unsigned long long base = 0;
  unsigned int sbase = 0;
  if (lane == 31 && (total & low54)) base = atomicAdd(&b.ctr->slice_elem, total & low54);
  if (lane == 30 && (total >> 54)) sbase = atomicAdd(&b.ctr->n_small, (unsigned)(total >> 54));
  base = __shfl_sync(FULL, base, 31);
  sbase = __shfl_sync(FULL, sbase, 30);
  const unsigned long long excl = incl - mine;
  const unsigned long long big_excl = base + (excl & low54);
  long long slot = (long long)(big_excl >> kElemBits);
  long long start = (long long)(big_excl & ((1ull << kElemBits) - 1));
  long long sslot = (long long)sbase + (long long)(excl >> 54);
  if (nbuf > 0) {
    if (slot + n_big > b.slice_cap || sslot + (nbuf - n_big) > b.slice_cap) {
      atomicOr(&b.ctr->overflow, 1u);
    } else {
      const int4 planes = __ldg(b.qmask + q);  // travels with every slice: the walk needs no per-query load
      for (int k = 0; k < nbuf; k++) {
        const int sz = sb.sz[k][threadIdx.x];
        const int4 rec = make_int4(q, sb.beg[k][threadIdx.x], sb.lm[k][threadIdx.x] | tag, sz);
        if (sz > kSmallSlice) {
          b.sl_start[slot] = start;
          b.sl_rec[2 * slot] = rec;
          b.sl_rec[2 * slot + 1] = planes;
          note_spans(b, slot, start, sz);
          slot++;
          start += sz;
        } else {
          b.sm_rec[2 * sslot] = rec;
          b.sm_rec[2 * sslot + 1] = planes;
          sslot++;
        }
      }
    }
  }
  nbuf = 0;
}

// One thread per (query, start position) chain: the n-gram walk of src/fuzzy_match.cc:484-551 with
// SuffixArray::equal_range (src/suffix_array.cc:105-212) restated as the equal range of the ONE new
// token at depth k inside the previous range (every suffix there already shares k tokens).
#ifndef FM_SEARCH_CTAS
#define FM_SEARCH_CTAS 12
#endif
__global__ void __launch_bounds__(FM_SEARCH_THREADS, FM_SEARCH_CTAS) fm_search_kernel(IndexDev ix, BatchDev b, Params pr) {
  __shared__ SliceBuf sb;
  int nbuf = 0;
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  int q = 0, it = 0, p = 0, ml = 0, tag = 0;
  const int32_t* pat = b.pat;
  bool live = c < b.n_tok;
  int t0 = 0, t1 = 0, t2 = 0;
  if (live) {
    // the chain's record and its first three words, independent reads (pat is padded: the reads past the last
    // pattern stay inside)
    const int2 cr = __ldg(b.chain_rec + c);
    t0 = b.pat[c];
    t1 = b.pat[c + 1];
    t2 = b.pat[c + 2];
    q = cr.x;
    it = cr.y & 1023;
    p = (cr.y >> 10) & 1023;
    tag = cr.y & ~1023;  // p << 10 | mult << 20
    pat = b.pat + (c - it);
    live = p > 0;
    ml = pr.ml;  // the clamp of the prepare kernels (src/fuzzy_match.cc:450-467)
    if (ml < 0 || ml > p) ml = p;
    const int by_ratio = (int)__fmul_rn(pr.mr, (float)p);
    if (by_ratio > ml) ml = by_ratio;
  }
  int lo = 0, hi = 0, len = 0;
  uint32_t bslot = 0;  // slot of the chain's trigram in the trigram directory (key of the 4-gram directory)
  int pos1 = -1;  // sa_pos[lo] once the range has shrunk to one suffix
  if (live) {
    if (ml >= 3) {
      // Nothing shorter than three words is registered (src/fuzzy_match.cc:546-550: a chain registers matches of
      // at least min_subseq_length words): whole array -> trigram in ONE probe of the trigram directory, the
      // bigram directory is never read. A chain without a third word, or with an unknown word, is dead at once.
      live = it + 2 < p && t0 >= 2 && t1 >= 2 && t2 >= 2 && tg_lookup(ix, t0, t1, t2, lo, hi, pos1, bslot);
      len = live ? 3 : 0;
    } else if (it + 1 < p) {
      // whole array -> first word -> bigram in one probe of the bigram directory. A chain that does
      // not reach length 2 registers nothing, so length 1 is skipped.
      if (t0 >= 2 && t1 >= 2) {
        uint32_t h = bigram_hash(t0, t1) & ix.bg_mask;
        for (;;) {
          const int4 e = __ldg(ix.bg_tab + h);
          if (e.x == t0 && e.y == t1) { lo = e.z; hi = e.w; len = 2; break; }
          if (e.x == -1) break;
          h = (h + 1) & ix.bg_mask;
        }
      }
      live = len == 2;
    } else {
      if (t0 >= 2) { lo = ix.qva[t0]; hi = ix.qva[t0 + 1]; }
      len = hi > lo ? 1 : 0;
      live = len == 1;
    }
  }
  // p == 1: the unigram range itself is registered (src/fuzzy_match.cc:484-493)
  if (live && p == 1 && 1 >= ml) push_slice(sb, nbuf, lo, hi - lo, 1);
  bool extending = live && it + len < p;
  while (__any_sync(FULL, extending)) {
    if (__any_sync(FULL, nbuf > kSliceBuf - 2)) flush_slices(b, sb, nbuf, lane, q, tag);  // room for two more
    if (extending) {
      const int t = pat[it + len];  // token at depth len
      int nlo = lo, nhi = lo, npos1 = -1;
      bool last = false;
      if (t >= 2 && len == 2) {
        // bigram -> trigram through the trigram directory (min_subseq_length < 3: the bigram range came first)
        if (!tg_lookup(ix, pat[it], pat[it + 1], t, nlo, nhi, npos1, bslot)) { nlo = nhi = lo; npos1 = -1; }
      } else if (t >= 2 && len == 3 && hi - lo > 1) {
        // trigram -> 4-gram through the 4-gram directory (it holds every 4-gram of a trigram that occurs more than
        // once): the one level where ranges are still wide costs one probe instead of a bisection
        uint32_t h = bigram_hash((int)bslot, t) & ix.qg_mask;
        for (;;) {
          const int4 e = __ldg(ix.qg_tab + h);
          if (e.x == (int)bslot && e.y == t) {
            nlo = e.z;
            nhi = e.w;
            if (e.w < 0) { nhi = e.z + 1; npos1 = -e.w - 1; }
            break;
          }
          if (e.x == -1) break;
          h = (h + 1) & ix.qg_mask;
        }
      } else if (t >= 2 && hi - lo == 1) {
        // a single suffix left: the rest of the chain is the common prefix of the pattern and that
        // suffix, and nothing is shaved off on the way (the separator 0 never equals a pattern token)
        if (pos1 < 0) pos1 = __ldg(ix.sa_pos + lo);
        int l2 = len;
        while (it + l2 < p && __ldg(ix.tok + (pos1 + l2)) == pat[it + l2]) l2++;
        if (l2 > len) { nlo = lo; nhi = hi; len = l2 - 1; }
        last = true;
      } else if (t >= 2) {
        // equal range of t at depth len >= 4 inside [lo, hi) (ranges are short there: the typical 4-gram range is
        // 1-2 suffixes). The search is latency bound, so it trades probes for dependent steps: (1) quaternary
        // narrowing, three independent pivots per step, until a pivot hits t; (2) from the hit the two ends of the
        // run of t gallop outwards together and finish by bisection.
        auto key = [&](int k) { return __ldg(ix.tok + (__ldg(ix.sa_pos + k) + len)); };
        int a = lo, e = hi, m = -1;
        while (m < 0 && a < e) {
          const int n = e - a;
          const int q1 = a + (n >> 2), q2 = a + (n >> 1), q3 = a + ((3 * n) >> 2);
          const int v1 = key(q1), v2 = key(q2), v3 = key(q3);
          if (v1 < t) a = q1 + 1;
          if (v2 < t) a = q2 + 1;
          if (v3 < t) a = q3 + 1;
          if (v3 > t) e = q3;
          if (v2 > t) e = q2;
          if (v1 > t) e = q1;
          m = v1 == t ? q1 : v2 == t ? q2 : v3 == t ? q3 : -1;
        }
        if (m >= 0) {
          // keys in [.., a) are < t, keys in [e, ..) are > t, key[m] == t
          int la = a, le = m, ls = 1;      // run start in [la, le]; keys in [le, m] == t
          int ua = m + 1, ue = e, us = 1;  // run end in [ua, ue]; keys in [m, ua) == t
          while (la < le || ua < ue) {
            const bool dl = la < le, du = ua < ue;
            const int pl = ls ? max(le - ls, la) : (int)(((unsigned)la + (unsigned)le) >> 1);
            const int pu = us ? min(ua + us - 1, ue - 1) : (int)(((unsigned)ua + (unsigned)ue) >> 1);
            int vl = 0, vu = 0;
            const int sl = dl ? __ldg(ix.sa_pos + pl) : 0, su = du ? __ldg(ix.sa_pos + pu) : 0;
            if (dl) vl = __ldg(ix.tok + (sl + len));
            if (du) vu = __ldg(ix.tok + (su + len));
            if (dl) {
              if (vl < t) { la = pl + 1; ls = 0; } else { le = pl; ls <<= 1; }
            }
            if (du) {
              if (vu > t) { ue = pu; us = 0; } else { ua = pu + 1; us <<= 1; }
            }
          }
          nlo = la;
          nhi = ua;
        }
      }
      if (nhi > nlo) {
        // range for length len+1 is non-empty; the shaved-off parts matched exactly len tokens
        if (len + 1 > 2 && len >= ml) {
          if (nlo > lo) push_slice(sb, nbuf, lo, nlo - lo, len);
          if (hi > nhi) push_slice(sb, nbuf, nhi, hi - nhi, len);
        }
        if (nlo != lo || nhi != hi || npos1 >= 0) pos1 = npos1;
        lo = nlo; hi = nhi; len++;
        if (it + len >= p || last) extending = false;
      } else {
        extending = false;
      }
    }
  }
  if (__any_sync(FULL, nbuf > kSliceBuf - 1)) flush_slices(b, sb, nbuf, lane, q, tag);
  if (live && len >= 2 && len >= ml) push_slice(sb, nbuf, lo, hi - lo, len);
  // The last flush -- for most chains the only one -- is made by the whole CTA: one pair of atomics on the
  // global counters per CTA instead of one per warp (same-address atomics serialise in L2).
  {
    __shared__ unsigned long long s_wtot[8], s_base;
    __shared__ unsigned int s_sbase;
    int elems = 0, n_big = 0;
    for (int k = 0; k < nbuf; k++) {
      const int sz = sb.sz[k][threadIdx.x];
      if (sz > kSmallSlice) { elems += sz; n_big++; }
    }
    const unsigned long long mine = ((unsigned long long)(nbuf - n_big) << 54) | ((unsigned long long)n_big << kElemBits) |
                                    (unsigned long long)(unsigned)elems;
    unsigned long long incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long o = __shfl_up_sync(FULL, incl, d);
      if (lane >= d) incl += o;
    }
    const int wib = threadIdx.x >> 5;
    if (lane == 31) s_wtot[wib] = incl;
    __syncthreads();
    const unsigned long long low54 = (1ull << 54) - 1;
    if (threadIdx.x == 0) {
      unsigned long long tot = 0;
      for (int k = 0; k < FM_SEARCH_THREADS / 32; k++) { const unsigned long long v = s_wtot[k]; s_wtot[k] = tot; tot += v; }
      s_base = (tot & low54) ? atomicAdd(&b.ctr->slice_elem, tot & low54) : 0ull;
      s_sbase = (tot >> 54) ? atomicAdd(&b.ctr->n_small, (unsigned)(tot >> 54)) : 0u;
    }
    __syncthreads();
    if (nbuf > 0) {
      const unsigned long long excl = s_wtot[wib] + incl - mine;
      const unsigned long long big_excl = s_base + (excl & low54);
      long long slot = (long long)(big_excl >> kElemBits);
      long long start = (long long)(big_excl & ((1ull << kElemBits) - 1));
      long long sslot = (long long)s_sbase + (long long)(excl >> 54);
      if (slot + n_big > b.slice_cap || sslot + (nbuf - n_big) > b.slice_cap) {
        atomicOr(&b.ctr->overflow, 1u);
      } else {
        const int4 planes = __ldg(b.qmask + q);
        for (int k = 0; k < nbuf; k++) {
          const int sz = sb.sz[k][threadIdx.x];
          const int4 rec = make_int4(q, sb.beg[k][threadIdx.x], sb.lm[k][threadIdx.x] | tag, sz);
          if (sz > kSmallSlice) {
            b.sl_start[slot] = start;
            b.sl_rec[2 * slot] = rec;
            b.sl_rec[2 * slot + 1] = planes;
            note_spans(b, slot, start, sz);
            slot++;
            start += sz;
          } else {
            b.sm_rec[2 * sslot] = rec;
            b.sm_rec[2 * sslot + 1] = planes;
            sslot++;
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------- gather

// PatternCoverage::count_covered_words (src/pattern_coverage.cc:15-28): every sentence token is looked up in
// the query's table; the first time a distinct pattern word is seen, its multiplicity is added. `seen` is
// a bitmask over distinct-word indices.
// Exact coverage of one sentence, stopping as soon as `need` is reached (the caller only compares
// the result with `need`; pass need > p to get the exact count). The first probes of the four words of a
// 128-bit load are issued together: the count is a chain of dependent table reads otherwise.
template <int MW>
__device__ __forceinline__ void cover_resolve(const int2* __restrict__ tbl, int tmask, int w, int h, int2 e, unsigned* seen, int& cover) {
  for (;;) {
    if (e.x == w) {
      const int d = e.y & 0xffff;
      const unsigned bit = 1u << (d & 31);
      unsigned& word = seen[MW == 1 ? 0 : (d >> 5)];
      if (!(word & bit)) { word |= bit; cover += e.y >> 16; }
      return;
    }
    if (e.x == -1) return;
    h = (h + 1) & tmask;
    e = __ldg(tbl + h);
  }
}
template <int MW>
__device__ __forceinline__ int cover_sentence(const int32_t* __restrict__ sent, int slen, const int2* tbl, int tmask, int need) {
  unsigned seen[MW];
#pragma unroll
  for (int i = 0; i < MW; i++) seen[i] = 0;
  int cover = 0;
  const int4* s4 = reinterpret_cast<const int4*>(sent);
  int4 t = ldg_nc_v4(s4);  // sentences start on 16-byte boundaries and are zero padded
  for (int k = 0; k < slen && cover < need; k += 4) {
    const int4 cur = t;
    if (k + 4 < slen) t = ldg_nc_v4(s4 + (k >> 2) + 1);
    const int h0 = hash32((uint32_t)cur.x) & tmask, h1 = hash32((uint32_t)cur.y) & tmask;
    const int h2 = hash32((uint32_t)cur.z) & tmask, h3 = hash32((uint32_t)cur.w) & tmask;
    const int2 e0 = __ldg(tbl + h0), e1 = __ldg(tbl + h1), e2 = __ldg(tbl + h2), e3 = __ldg(tbl + h3);
    cover_resolve<MW>(tbl, tmask, cur.x, h0, e0, seen, cover);
    if (k + 1 < slen) cover_resolve<MW>(tbl, tmask, cur.y, h1, e1, seen, cover);
    if (k + 2 < slen) cover_resolve<MW>(tbl, tmask, cur.z, h2, e2, seen, cover);
    if (k + 3 < slen) cover_resolve<MW>(tbl, tmask, cur.w, h3, e3, seen, cover);
  }
  return cover;
}


// Stage 1 for one walk record: upper bound on the coverage from the signature against the smallest
// coverage that passes for this (pattern length, sentence length). row = cmin64 + (p << 6); m = the
// query's planes (B0 lo, B0 hi, B1 lo, B1 hi). Branch-free; the rare correction for signature bits that
// collect more than three pattern positions (mult != 0, uniform over a slice) is added by the caller.
__device__ __forceinline__ int stage1_margin(unsigned lo, unsigned hi, const int4& m, const uint16_t* __restrict__ row) {
  const int need = __ldg(row + (lo & 63u));
  return __popc(lo & (unsigned)m.x) + __popc(hi & (unsigned)m.y) + 2 * (__popc(lo & (unsigned)m.z) + __popc(hi & (unsigned)m.w)) - need;
}
__device__ __forceinline__ int stage1_extra(unsigned lo, unsigned hi, const int4& m, int mult) {
  return mult * (__popc(lo & (unsigned)m.x & (unsigned)m.z) + __popc(hi & (unsigned)m.y & (unsigned)m.w));
}

// "Suffix-range gather", first kernel: register_suffix_range_match's walk (src/ngram_matches.cc:62-84) with
// the signature form of the coverage filter (src/fuzzy_match.cc:576-581) evaluated per element: one 8-byte
// walk record -- length and 58-bit signature -- against the query's planes and the per-length bound table;
// no sentence is touched. A pure streaming filter: the elements that pass go to the candidate list
// (q | match length << 20, suffix-array index) and are looked at by fm_verify_kernel.
// A CTA takes blocks of work in turn (the grid is larger than what is resident, so the hardware balances
// the CTAs):
//  * a block of flattened elements, kSpan per warp. Slices of more than kSmallSlice elements (94 % of the
//    elements sit in slices of 33 or more) are flattened; a warp walks its span slice by slice, four records
//    per lane from one 256-bit load. The 32-byte slice records of the span (query, range, planes) are
//    fetched by the lanes in one go into shared memory and every line of the span is prefetched into L2
//    before the walk starts;
//  * or a block of small slices (three quarters of all slices hold one suffix), one slice per thread.
// Candidates are collected per CTA in shared memory; one atomic per block reserves their place in the list.
#ifndef FM_GATHER_CTAS
#define FM_GATHER_CTAS 5
#endif
static const int kPackMax = 28;  // segments of up to 28 records (32 with the alignment of the 256-bit loads) are walked four to an iteration
static const int kWalkQueue = 3072;  // candidates of a block staged in shared memory (8 * kSpan elements; beyond: straight to the list)
struct SliceWin {  // the slice records of the span a warp is walking
  int4 rec[32];
  int4 planes[32];
  int st[32];  // first flattened element of the slice, relative to the span start
};
// nfill[0] = slots reserved so far, nfill[1] = end of the last reservation that fitted the shared-memory queue
__device__ __forceinline__ void block_push(const BatchDev& b, int2* queue, int* nfill, bool pass, int2 item, int lane) {
  const unsigned bal = __ballot_sync(FULL, pass);
  if (!bal) return;
  const int cnt = __popc(bal);
  int base = 0;
  if (lane == 0) base = atomicAdd(&nfill[0], cnt);
  base = __shfl_sync(FULL, base, 0);
  const int rank = __popc(bal & ((1u << lane) - 1));
  if (base + cnt <= kWalkQueue) {
    if (pass) queue[base + rank] = item;
    if (lane == 0) atomicMax(&nfill[1], base + cnt);
  } else {  // the block's queue is full (dense candidates): this push goes straight to the list
    unsigned g = 0;
    if (lane == 0) g = atomicAdd(&b.ctr->n_cand, (unsigned)cnt);
    g = __shfl_sync(FULL, g, 0);
    if ((long long)g + cnt <= b.cand_cap) {
      if (pass) b.cand[(long long)g + rank] = item;
    } else if (lane == 0) {
      atomicOr(&b.ctr->overflow, 8u);
    }
  }
}
// The same for the four elements a lane tested in one group: one reservation for all of them.
__device__ __forceinline__ void block_push4(const BatchDev& b, int2* queue, int* nfill, bool p0, bool p1, bool p2, bool p3, int qlm,
                                            int sa0, int lane) {
  const unsigned b0 = __ballot_sync(FULL, p0), b1 = __ballot_sync(FULL, p1), b2 = __ballot_sync(FULL, p2), b3 = __ballot_sync(FULL, p3);
  const int c0 = __popc(b0), c1 = __popc(b1), c2 = __popc(b2);
  const int cnt = c0 + c1 + c2 + __popc(b3);
  int base = 0;
  if (lane == 0) base = atomicAdd(&nfill[0], cnt);
  base = __shfl_sync(FULL, base, 0);
  const unsigned lt = (1u << lane) - 1;
  const int o0 = __popc(b0 & lt), o1 = c0 + __popc(b1 & lt), o2 = c0 + c1 + __popc(b2 & lt), o3 = c0 + c1 + c2 + __popc(b3 & lt);
  if (base + cnt <= kWalkQueue) {
    if (p0) queue[base + o0] = make_int2(qlm, sa0);
    if (p1) queue[base + o1] = make_int2(qlm, sa0 + 1);
    if (p2) queue[base + o2] = make_int2(qlm, sa0 + 2);
    if (p3) queue[base + o3] = make_int2(qlm, sa0 + 3);
    if (lane == 0) atomicMax(&nfill[1], base + cnt);
  } else {  // the block's queue is full (dense candidates): this push goes straight to the list
    unsigned g = 0;
    if (lane == 0) g = atomicAdd(&b.ctr->n_cand, (unsigned)cnt);
    g = __shfl_sync(FULL, g, 0);
    if ((long long)g + cnt <= b.cand_cap) {
      if (p0) b.cand[(long long)g + o0] = make_int2(qlm, sa0);
      if (p1) b.cand[(long long)g + o1] = make_int2(qlm, sa0 + 1);
      if (p2) b.cand[(long long)g + o2] = make_int2(qlm, sa0 + 2);
      if (p3) b.cand[(long long)g + o3] = make_int2(qlm, sa0 + 3);
    } else if (lane == 0) {
      atomicOr(&b.ctr->overflow, 8u);
    }
  }
}
__global__ void __launch_bounds__(256, FM_GATHER_CTAS) fm_gather_kernel(IndexDev ix, BatchDev b) {
  __shared__ int2 s_queue[kWalkQueue];
  __shared__ SliceWin s_win[8];
  __shared__ int s_n[2], s_base;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  SliceWin& win = s_win[wib];
  // A worklist of the search overflowed: the host regrows and reruns. One decision per CTA (the CTA's warps
  // work together below). An overflow of the candidate list itself (bit 3, set by other CTAs of this kernel)
  // does not stop the walk: the count goes on, so that the host learns the size to regrow to.
  if (threadIdx.x == 0) s_n[0] = (int)(b.ctr->overflow & 5u);
  __syncthreads();
  const bool stop = s_n[0] != 0;
  __syncthreads();
  if (stop) return;
  const unsigned long long packed = b.ctr->slice_elem;
  const long long total = (long long)(packed & ((1ull << kElemBits) - 1));
  const int n_big = (int)(packed >> kElemBits);
  const int n_small = (int)b.ctr->n_small;
  const int n_spans = (int)((total + kSpan - 1) / kSpan);
  const int span_blocks = (n_spans + 7) >> 3, small_blocks = (n_small + 255) >> 8;
  for (int blk = blockIdx.x; blk < span_blocks + small_blocks; blk += gridDim.x) {
    if (threadIdx.x == 0) s_n[0] = s_n[1] = 0;
    __syncthreads();
    if (blk < span_blocks) {
      const int sp = blk * 8 + wib;
      if (sp < n_spans) {
        const long long span_base = (long long)sp * kSpan;
        const int span_len = (int)min((long long)kSpan, total - span_base);
        int k = __ldg(b.span_slice + sp);  // slice that holds the first element of the span
        int kwin = k - 32;
        int pos = 0;
        while (pos < span_len) {
          if (k >= kwin + 32) {  // fetch the records of slices k .. k+31, one per lane, and prefetch their elements
            kwin = k;
            __syncwarp();
            int4 r0 = make_int4(0, 0, 0, 0);
            int st = 1 << 30;
            if (k + lane < n_big) {
              r0 = __ldg(b.sl_rec + 2 * (k + lane));
              win.rec[lane] = r0;
              win.planes[lane] = __ldg(b.sl_rec + 2 * (k + lane) + 1);
              st = (int)max(-(1ll << 31), min(1ll << 30, __ldg(b.sl_start + k + lane) - span_base));
              win.st[lane] = st;
            }
            // lane l: the first line of walk records of slice k + l inside the rest of the span (the walk itself
            // prefetches two groups ahead inside a slice)
            const int e0 = max(st, pos);
            if (e0 < span_len && e0 < st + r0.w) prefetch_l2(ix.sa_rec + (r0.y + (e0 - st)));
            __syncwarp();
          }
          const int4 sr = win.rec[k - kwin];  // (q, sa_begin, lm | p << 10 | mult << 20, size), same for all lanes
          const int4 m = win.planes[k - kwin];
          const int st = win.st[k - kwin];
          const int seg_end = (int)min((long long)st + sr.w, (long long)span_len);
          if (seg_end - pos <= kPackMax) {
            // Short segments (a third of the loop iterations belong to slices that fill a fraction of the 128
            // records of an iteration): up to four consecutive slices share one iteration, eight lanes -- 32 aligned
            // records -- each, every group of lanes with its own slice record from the window.
            int npk = 1, last_end = seg_end;
#pragma unroll
            for (int j = 1; j < 4; j++) {
              const int wj = k - kwin + j;
              if (npk == j && wj < 32 && k + j < n_big) {
                const int stj = win.st[wj];
                const int endj = (int)min((long long)stj + win.rec[wj].w, (long long)span_len);
                if (stj < span_len && endj - stj <= kPackMax) { npk = j + 1; last_end = endj; }
              }
            }
            const int grp = lane >> 3, sub = lane & 7;
            const bool act = grp < npk;
            const int wi = k - kwin + (act ? grp : 0);
            const int4 gsr = win.rec[wi];
            const int4 gm = win.planes[wi];
            const int gst = win.st[wi];
            const int gbeg = max(gst, pos), gend = (int)min((long long)gst + gsr.w, (long long)span_len);
            const int ga0 = gsr.y + (gbeg - gst), ga1 = gsr.y + (gend - gst);
            const unsigned glen = (unsigned)(ga1 - ga0);
            const int gmult = gsr.z >> 20;
            const uint16_t* grow = b.cmin64 + (((gsr.z >> 10) & 1023) << 6);
            const int gqlm = gsr.x | ((gsr.z & 1023) << 20);
            const int base = (ga0 & ~3) + 4 * sub;
            unsigned r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (act && base < ga1) ldg_nc_v8(ix.sa_rec + base, r);
            int d0 = stage1_margin(r[0], r[1], gm, grow), d1 = stage1_margin(r[2], r[3], gm, grow);
            int d2 = stage1_margin(r[4], r[5], gm, grow), d3 = stage1_margin(r[6], r[7], gm, grow);
            if (__any_sync(FULL, gmult != 0)) {
              d0 += stage1_extra(r[0], r[1], gm, gmult); d1 += stage1_extra(r[2], r[3], gm, gmult);
              d2 += stage1_extra(r[4], r[5], gm, gmult); d3 += stage1_extra(r[6], r[7], gm, gmult);
            }
            const unsigned rel = (unsigned)(base - ga0);
            const bool p0 = act & (d0 >= 0) & (rel < glen), p1 = act & (d1 >= 0) & (rel + 1u < glen);
            const bool p2 = act & (d2 >= 0) & (rel + 2u < glen), p3 = act & (d3 >= 0) & (rel + 3u < glen);
            if (__any_sync(FULL, p0 | p1 | p2 | p3)) block_push4(b, s_queue, s_n, p0, p1, p2, p3, gqlm, base, lane);
            pos = last_end;
            k += npk;
            continue;
          }
          const int mult = sr.z >> 20;
          const uint16_t* row = b.cmin64 + (((sr.z >> 10) & 1023) << 6);
          const int qlm = sr.x | ((sr.z & 1023) << 20);
          const int a0 = sr.y + (pos - st), a1 = sr.y + (seg_end - st);  // suffix-array indices [a0, a1)
          const unsigned len = (unsigned)(a1 - a0);
          pos = seg_end;
          k++;
          for (int g = a0 & ~3; g < a1; g += 128) {  // 256-bit loads cover aligned groups of four records
            const int base = g + 4 * lane;
            unsigned r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (base < a1) ldg_nc_v8(ix.sa_rec + base, r);
            if (base + 256 < a1) prefetch_l2(ix.sa_rec + base + 256);
            int d0 = stage1_margin(r[0], r[1], m, row), d1 = stage1_margin(r[2], r[3], m, row);
            int d2 = stage1_margin(r[4], r[5], m, row), d3 = stage1_margin(r[6], r[7], m, row);
            if (mult) {  // (uniform)
              d0 += stage1_extra(r[0], r[1], m, mult); d1 += stage1_extra(r[2], r[3], m, mult);
              d2 += stage1_extra(r[4], r[5], m, mult); d3 += stage1_extra(r[6], r[7], m, mult);
            }
            const unsigned rel = (unsigned)(base - a0);  // element base + i is inside iff rel + i < len (unsigned)
            const bool p0 = (d0 >= 0) & (rel < len), p1 = (d1 >= 0) & (rel + 1u < len);
            const bool p2 = (d2 >= 0) & (rel + 2u < len), p3 = (d3 >= 0) & (rel + 3u < len);
            if (__any_sync(FULL, p0 | p1 | p2 | p3)) block_push4(b, s_queue, s_n, p0, p1, p2, p3, qlm, base, lane);
          }
        }
      }
    } else {
      const int si = (blk - span_blocks) * 256 + threadIdx.x;
      int4 sr = make_int4(0, 0, 0, 0), sm = make_int4(0, 0, 0, 0);
      if (si < n_small) {
        sr = __ldg(b.sm_rec + 2 * si);
        sm = __ldg(b.sm_rec + 2 * si + 1);
      }
      const uint16_t* srow = b.cmin64 + (((sr.z >> 10) & 1023) << 6);
      const int smult = sr.z >> 20, sqlm = sr.x | ((sr.z & 1023) << 20);
      uint2 rec[kSmallSlice];
#pragma unroll
      for (int e = 0; e < kSmallSlice; e++) rec[e] = e < sr.w ? __ldg(ix.sa_rec + sr.y + e) : make_uint2(0u, 0u);
#pragma unroll
      for (int e = 0; e < kSmallSlice; e++) {
        const bool pass = (e < sr.w) & (stage1_margin(rec[e].x, rec[e].y, sm, srow) + stage1_extra(rec[e].x, rec[e].y, sm, smult) >= 0);
        block_push(b, s_queue, s_n, pass, make_int2(sqlm, sr.y + e), lane);
      }
    }
    __syncthreads();
    const int n = s_n[1];
    if (n) {  // (uniform)
      if (threadIdx.x == 0) s_base = (int)atomicAdd(&b.ctr->n_cand, (unsigned)n);
      __syncthreads();
      const long long base = (unsigned)s_base;
      if (base + n <= b.cand_cap) {
        for (int i = threadIdx.x; i < n; i += 256) b.cand[base + i] = s_queue[i];
      } else if (threadIdx.x == 0) {
        atomicOr(&b.ctr->overflow, 8u);  // the count goes on, so that the host knows the size to regrow to
      }
      __syncthreads();
    }
  }
}

// Insert (q, start) into the dedup table keeping the max match length: NGramMatches::_longest_matches
// (src/ngram_matches.cc:79-81). The survivor's slot in the compact list comes from the caller (reserved per
// CTA); the first inserter also claims the candidate's slot inside its query.
static const int kSurvStage = 512;
struct SurvStage {  // survivors of one CTA block, staged in shared memory
  SurvRec rec[kSurvStage];
  uint16_t len[kSurvStage];
};
__device__ __forceinline__ int add_survivor(const BatchDev& b, SurvStage& stage, int* n_stage, int q, int start, int slen, int lm) {
  const unsigned long long key = ((unsigned long long)(unsigned)q << 32) | (unsigned)start;
  uint32_t h = hash64(key) & b.hmask;
  for (int probes = 0; probes < 4096; probes++) {
    const unsigned long long prev = atomicCAS(&b.hkey[h], ~0ull, key);
    if (prev == ~0ull) {
      const int j = atomicAdd(&b.q_cnt[q], 1);
      atomicMax(&b.hlm[h], (unsigned)lm);
      const int i = atomicAdd(n_stage, 1);
      if (i < kSurvStage) {
        stage.rec[i] = SurvRec{q, start, (int32_t)h, j};
        stage.len[i] = (uint16_t)slen;
      } else {  // the block's stage is full (a block of several rounds with dense survivors): straight to the list
        const unsigned g = atomicAdd(&b.ctr->n_surv, 1u);
        if ((long long)g < b.surv_cap) {
          b.surv[g] = SurvRec{q, start, (int32_t)h, j};
          b.surv_len[g] = (uint16_t)slen;
        } else {
          atomicOr(&b.ctr->overflow, 2u);
        }
      }
      return (int)h;
    }
    if (prev == key) {
      atomicMax(&b.hlm[h], (unsigned)lm);
      return (int)h;
    }
    h = (h + 1) & b.hmask;
  }
  atomicOr(&b.ctr->overflow, 2u);  // (a probe sequence this long: the table is all but full)
  return -1;
}
// Slot of (q, start) if it already is a survivor, else -1 (a concurrent insert may be missed: the caller
// then verifies the pair again, which is harmless).
__device__ __forceinline__ int find_survivor(const BatchDev& b, int q, int start) {
  const unsigned long long key = ((unsigned long long)(unsigned)q << 32) | (unsigned)start;
  uint32_t h = hash64(key) & b.hmask;
  for (int probes = 0; probes < 4096; probes++) {
    const unsigned long long k = *(volatile const unsigned long long*)&b.hkey[h];
    if (k == key) return (int)h;
    if (k == ~0ull) return -1;
    h = (h + 1) & b.hmask;
  }
  return -1;
}

// "Suffix-range gather", second kernel: the candidates of the walk, 32 per warp and round. Each lane
// resolves one item -- walk record, sentence start, the query's length / table offset, the smallest passing
// coverage for this length pair -- and then:
//  * sentences with a wide signature (longer than kWideMin tokens): first the upper bound on the coverage
//    from the 1024-bit signature and the query's bit-sliced planes -- 8 lanes per candidate, 128-bit loads,
//    4 candidates per round -- so that only plausible pairs reach the exact count;
//  * patterns of up to 32 words against short sentences: exact coverage, one candidate per lane, the set
//    of distinct words seen is one register, the count stops as soon as the bound is met;
//  * everything else: one candidate at a time, the 32 lanes probe the sentence's tokens in parallel and
//    mark the distinct words in a shared-memory bit set (PatternCoverage::count_covered_words,
//    src/pattern_coverage.cc:15-28).
// A pair that passes the coverage bound (theoretical_rejection_cover, src/ngram_matches.cc:42-59) enters the
// dedup table with its match length; survivors are staged per CTA and appended to the compact list with
// one atomic per block of 512 candidates.
struct Cand {  // a resolved candidate, exchanged between lanes through shared memory
  int q_lm;    // q | match length << 20
  int start;   // sentence start in tok
  int len_need;  // sentence length | need << 16 (need = 0xffff: no bound table, evaluate the bounds)
  int wrow;    // wide signature row or -1
};
__device__ __forceinline__ int verify_candidates(const IndexDev& ix, const BatchDev& b, const Params& pr, const int2* items, int n,
                                                 Cand* cand, unsigned* seen, SurvStage& stage, int* n_stage, int lane) {
  bool have = lane < n;
  int q = 0, lm = 0, start = 0, slen = 0, need = 0, wrow = -1, p = 0, off = 0;
  if (have) {
    const int2 item = items[lane];
    q = item.x & 0xfffff;
    lm = item.x >> 20;
    unsigned ax[8];
    ldg_nc_v8(ix.sa_aux + 2ll * item.y, ax);  // start, length, second signature (or wide row): one 32-byte read
    start = (int)ax[0];
    const QMeta qm = __ldg(b.qmeta + q);
    p = qm.x;
    off = qm.z;
    slen = (int)(ax[1] & 0x7fffffffu);
    if ((int)ax[1] < 0) wrow = (int)ax[2];
    need = __ldg(b.cmin_tab + ((p << 10) | slen));
    if (need == kNeedReject) have = false;  // (a long sentence outside the length window)
    else if (need == kNeedNoTable) {
      need = 0xffff;
      if (reject_length(p, slen, pr)) have = false;
    } else if (wrow < 0) {
      // Second signature: an independent 192-bit word -> bit map of the sentence (in sa_aux) against the query's planes
      // over the same map (qmask2). Like stage 1 of the walk it bounds the coverage from above, so a candidate whose
      // bound stays below the smallest passing coverage cannot pass the exact count: no sentence fetch and no table
      // probes for it. Three times as wide as the walk's signature, so word collisions add little to the bound:
      // at f=0.5, 83 % of the candidates whose exact count would fail stop here (56 % with a 64-bit map).
      const int4* mq = b.qmask2 + 3ll * q;
      const int4 ma = __ldg(mq), mb = __ldg(mq + 1), mc = __ldg(mq + 2);
      const unsigned b0[kSig2Words] = {(unsigned)ma.x, (unsigned)ma.y, (unsigned)ma.z, (unsigned)ma.w, (unsigned)mb.x, (unsigned)mb.y};
      const unsigned b1[kSig2Words] = {(unsigned)mb.z, (unsigned)mb.w, (unsigned)mc.x, (unsigned)mc.y, (unsigned)mc.z, (unsigned)mc.w};
      const int mult2 = (qm.w >> 18) & 0x3ff;
      int u0 = 0, u1 = 0, u2 = 0;
#pragma unroll
      for (int k = 0; k < kSig2Words; k++) {
        u0 += __popc(ax[2 + k] & b0[k]);
        u1 += __popc(ax[2 + k] & b1[k]);
        u2 += __popc(ax[2 + k] & b0[k] & b1[k]);
      }
      if (u0 + 2 * u1 + mult2 * u2 < need) have = false;
    }
  }
  // The sentence a query really matches is reached through many of its n-grams, so many candidates are
  // repeats of a pair that is a survivor already: those only raise the recorded match length
  // (NGramMatches::_longest_matches keeps the maximum, src/ngram_matches.cc:79-81). Repeats inside this
  // round are verified once, by their lowest lane.
  int slot = -1;
  if (have) {
    slot = find_survivor(b, q, start);
    if (slot >= 0) {
      atomicMax(&b.hlm[slot], (unsigned)lm);
      have = false;
    }
  }
  const unsigned long long pair_key = have ? (((unsigned long long)(unsigned)q << 32) | (unsigned)start) : ~(unsigned long long)lane;
  const int leader = __ffs(__match_any_sync(FULL, pair_key)) - 1;
  const bool repeat = have && leader != lane;
  if (repeat) have = false;
  slot = -1;
  const int n_verified = __popc(__ballot_sync(FULL, have));
  const bool wide = have && wrow >= 0;
  unsigned wide_pass = 0;
  if (__any_sync(FULL, (have && p > 32) || wide)) {
    __syncwarp();
    cand[lane] = Cand{q | (lm << 20), start, slen | (need << 16), wrow};
    __syncwarp();
  }
  if (__any_sync(FULL, wide)) {
    wide_pass = __ballot_sync(FULL, wide && need == 0xffff);  // no bound table: straight to the exact count
    unsigned todo = __ballot_sync(FULL, wide && need != 0xffff);
    const int grp = lane >> 3, sub = lane & 7;
    const unsigned gmask = 0xffu << (8 * grp);
    while (todo) {
      // the next (up to) four pending candidates, one per group of eight lanes
      unsigned t = todo;
      int src = -1;
#pragma unroll
      for (int g = 0; g < 4; g++) {
        const int s0 = t ? __ffs(t) - 1 : -1;
        if (g == grp) src = s0;
        t &= t - 1;
      }
      todo = t;
      const bool valid = src >= 0;
      int ub = 0, cneed = 0;
      if (valid) {
        const Cand c = cand[src];
        const int cq = c.q_lm & 0xfffff;
        cneed = (int)((unsigned)c.len_need >> 16);
        const uint32_t* sig = ix.wsig + (size_t)c.wrow * kWideWords;
        const uint32_t* wq = b.wq + (size_t)cq * kWideStride;
        const uint4 sg = __ldg(reinterpret_cast<const uint4*>(sig) + sub);
        const uint4* pl = reinterpret_cast<const uint4*>(wq) + sub;
        const uint4 b0 = __ldg(pl), b1 = __ldg(pl + kWideWords / 4), b2 = __ldg(pl + 2 * (kWideWords / 4));
        const uint32_t big = __ldg(wq + 3 * kWideWords + sub);  // (bit | excess << 16) of a bit with more than 7 positions
        ub = __popc(sg.x & b0.x) + __popc(sg.y & b0.y) + __popc(sg.z & b0.z) + __popc(sg.w & b0.w) +
             2 * (__popc(sg.x & b1.x) + __popc(sg.y & b1.y) + __popc(sg.z & b1.z) + __popc(sg.w & b1.w)) +
             4 * (__popc(sg.x & b2.x) + __popc(sg.y & b2.y) + __popc(sg.z & b2.z) + __popc(sg.w & b2.w));
        if (big >> 16) {
          const unsigned bit = big & 0x3ffu;
          if ((__ldg(sig + (bit >> 5)) >> (bit & 31u)) & 1u) ub += (int)(big >> 16);
        }
        if (sub == 0) ub += (int)__ldg(wq + 3 * kWideWords + kWideBig);
      }
      ub = __reduce_add_sync(gmask, ub);
      wide_pass |= __reduce_or_sync(FULL, (valid && sub == 0 && ub >= cneed) ? (1u << src) : 0u);
    }
  }
  if (have && !wide && p <= 32) {
    const int2* tbl = b.tbl + 4ll * off;
    const int tmask = next_pow2(2 * p) - 1;
    bool ok;
    if (need == 0xffff) {  // no bound table for this query: evaluate the bound itself
      ok = !reject_cover(p, slen, cover_sentence<1>(ix.tok + start, slen, tbl, tmask, p + 1), pr);
    } else {
      ok = cover_sentence<1>(ix.tok + start, slen, tbl, tmask, need) >= need;
    }
    if (ok) slot = add_survivor(b, stage, n_stage, q, start, slen, lm);
  }
  unsigned todo = __ballot_sync(FULL, have && !wide && p > 32) | wide_pass;
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const Cand c = cand[src];
    const int cq = c.q_lm & 0xfffff, clm = c.q_lm >> 20, cstart = c.start;
    const int cslen = c.len_need & 0xffff, cneed = (int)((unsigned)c.len_need >> 16);
    const int cp = __shfl_sync(FULL, p, src), coff = __shfl_sync(FULL, off, src);
    const int2* tbl = b.tbl + 4ll * coff;
    const int tmask = next_pow2(2 * cp) - 1;
    seen[lane] = 0;
    __syncwarp();
    int cover = 0;
    for (int k = lane; k < cslen; k += 32) {
      const int w = __ldg(ix.tok + cstart + k);
      int h = hash32((uint32_t)w) & tmask;
      for (;;) {
        const int2 e = __ldg(tbl + h);
        if (e.x == w) {
          const int d = e.y & 0xffff;
          const unsigned bit = 1u << (d & 31);
          if (!(atomicOr(&seen[d >> 5], bit) & bit)) cover += e.y >> 16;
          break;
        }
        if (e.x == -1) break;
        h = (h + 1) & tmask;
      }
    }
    cover = __reduce_add_sync(FULL, cover);
    __syncwarp();
    const bool ok = cneed == 0xffff ? !reject_cover(cp, cslen, cover, pr) : cover >= cneed;
    int hs = -1;
    if (ok && lane == 0) hs = add_survivor(b, stage, n_stage, cq, cstart, cslen, clm);
    hs = __shfl_sync(FULL, hs, 0);
    if (lane == src) slot = hs;
  }
  const int lslot = __shfl_sync(FULL, slot, leader);
  if (repeat && lslot >= 0) atomicMax(&b.hlm[lslot], (unsigned)lm);
  return n_verified;
}

#ifndef FM_VERIFY_CTAS
#define FM_VERIFY_CTAS 5  // 48 registers, 40 warps per SM: verify 0.092 -> 0.083 ms at f=0.7, 0.49 -> 0.41 ms at f=0.5 (6 CTAs: spills, no better)
#endif
__global__ void __launch_bounds__(256, FM_VERIFY_CTAS) fm_verify_kernel(IndexDev ix, BatchDev b, Params pr) {
  __shared__ SurvStage s_stage;
  __shared__ Cand s_cand[8][32];
  __shared__ unsigned s_seen[8][32];
  __shared__ int s_n, s_base;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  // one decision per CTA: other CTAs of this kernel may raise the flag while this one starts
  if (threadIdx.x == 0) s_n = (int)b.ctr->overflow;
  __syncthreads();
  const bool stop = s_n != 0;
  __syncthreads();
  if (stop) return;
  const long long n_cand = (long long)b.ctr->n_cand;
  // A block ends with a CTA barrier (its survivors leave the shared-memory stage together), and the warps of a CTA
  // finish their rounds at different times (exact counts are unevenly spread): 19 % of the stall samples at f=0.5
  // with two rounds per block. Long candidate lists take eight rounds per block -- a quarter of the barriers, and
  // the differences average out over the rounds; short lists keep small blocks so that every CTA gets some.
  const int rounds = n_cand > (long long)gridDim.x * 2048 ? 8 : n_cand > (long long)gridDim.x * 1024 ? 4 : 2;
  const int per_block = rounds * 256;
  const int n_blocks = (int)((n_cand + per_block - 1) / per_block);
  int verified = 0;
  for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int r = 0; r < rounds; r++) {
      const long long first = (long long)blk * per_block + (r * 8 + wib) * 32;
      const int n = (int)min(32ll, n_cand - first);
      if (n > 0) verified += verify_candidates(ix, b, pr, b.cand + first, n, s_cand[wib], s_seen[wib], s_stage, &s_n, lane);
    }
    __syncthreads();
    const int n = min(s_n, kSurvStage);
    if (n) {  // (uniform)
      if (threadIdx.x == 0) s_base = (int)atomicAdd(&b.ctr->n_surv, (unsigned)n);
      __syncthreads();
      const long long base = (unsigned)s_base;
      if (base + n <= b.surv_cap) {
        for (int i = threadIdx.x; i < n; i += 256) {
          b.surv[base + i] = s_stage.rec[i];
          b.surv_len[base + i] = s_stage.len[i];
        }
      } else if (threadIdx.x == 0) {
        atomicOr(&b.ctr->overflow, 2u);
      }
      __syncthreads();
    }
  }
  if (lane == 0 && verified) atomicAdd(&b.ctr->n_verified, (unsigned)verified);
}

// ---------------------------------------------------------------- scan (<= 128 co-resident CTAs)

// Exclusive scan of the per-query survivor counts over co-resident CTAs: CTA i reduces its tile,
// publishes the total (a flag/value pair in `chain`) and sums the totals of CTAs 0..i-1. The grid
// (<= 128 CTAs of 1024 threads) is always fully resident, so the spin cannot deadlock.
static const int kScanCtas = 128;
__global__ void __launch_bounds__(1024) fm_scan_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out, int n,
                                                       unsigned long long* chain, unsigned int epoch) {
  __shared__ int warp_sum[32];
  __shared__ int carry_s;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int per_cta = (((n + gridDim.x - 1) / gridDim.x) + 4095) / 4096 * 4096;
  const int beg = blockIdx.x * per_cta, end = min(n, beg + per_cta);
  // pass 1: tile total
  int total = 0;
  for (int i = beg + t; i < end; i += 1024) total += in[i];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) total += __shfl_xor_sync(FULL, total, d);
  if (lane == 0) warp_sum[wid] = total;
  __syncthreads();
  if (wid == 0) {
    // publish this tile's total, then add up the totals of all predecessors (each one is read as soon
    // as its CTA has published it: no serial hand-over from CTA to CTA)
    int tile = warp_sum[lane];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) tile += __shfl_xor_sync(FULL, tile, d);
    if (lane == 0) *(volatile unsigned long long*)(chain + blockIdx.x) = ((unsigned long long)epoch << 32) | (unsigned)tile;
    int carry = 0;
    for (int j = lane; j < (int)blockIdx.x; j += 32) {
      volatile unsigned long long* pj = chain + j;
      unsigned long long v;
      do { v = *pj; } while ((unsigned)(v >> 32) != epoch);
      carry += (int)(unsigned)v;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) carry += __shfl_xor_sync(FULL, carry, d);
    if (lane == 0) {
      carry_s = carry;
      if (blockIdx.x == gridDim.x - 1) out[n] = carry + tile;
    }
  }
  __syncthreads();
  // pass 2: exclusive scan of the tile, 4096 items per round
  int carry = carry_s;
  for (int base = beg; base < end; base += 4096) {
    const int i0 = base + 4 * t;
    int v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = i0 + k < end ? in[i0 + k] : 0;
    const int mine = v[0] + v[1] + v[2] + v[3];
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(FULL, incl, d);
      if (lane >= d) incl += o;
    }
    __syncthreads();
    if (lane == 31) warp_sum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int ws = warp_sum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(FULL, ws, d);
        if (lane >= d) ws += o;
      }
      warp_sum[lane] = ws;
    }
    __syncthreads();
    int run = carry + (wid ? warp_sum[wid - 1] : 0) + incl - mine;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (i0 + k < end) out[i0 + k] = run;
      run += v[k];
    }
    carry += warp_sum[31];
  }
}

// ---------------------------------------------------------------- edit distance (warp wavefront)

// Token-level Levenshtein of src/edit_distance.cc:5-77 (and :79-122 with pen == 0), rows = TM
// sentence (n1 = s), columns = pattern (n2 = p). Lane l owns columns l*c+1 .. (l+1)*c; at step t it
// works on row t-l, so the anti-diagonal moves one lane per step and the only exchange is one
// __shfl_up of (left value, running row minimum). Column state lives in shared memory.
// Returns C = arr[s][p] and K = max over rows of min_{j>=1} arr[i][j] in every lane.
// With REAL (Sentence API: src/edit_distance.cc:19-26,31,38,53-62): rs holds the real tokens and gap
// itok ids of both sides in shared memory; ct(a, b) = _edit_distance_char of two penalty tokens comes
// from the caller's table dist[a * n_itok + b] (dist[0][b] = length of b).
struct RealSide {
  const int32_t* real1;  // [s]   (form id << 1) | case class of the TM sentence
  const int32_t* gap1;   // [s+1] itok id of each gap of the TM sentence
  const int32_t* real2;  // [p]   pattern
  const int32_t* gap2;   // [p+1]
  const int32_t* dist;   // [n_itok * n_itok]
  int n_itok;
  float rep;             // replace cost (for the case / real-form differences)
};
template <bool REAL>
__device__ void warp_edit_distance(const int32_t* s_sent, int s, const int32_t* s_pat, int p, const float* s_pen,
                                   float* s_up, float delw, float insw, float repw, float& C_out, float& K_out,
                                   const RealSide& rs) {
  const int lane = threadIdx.x & 31;
  const int c = (p + 31) >> 5;
  const int nl = (p + c - 1) / c;  // lanes that own columns
  float origin = 0.f;               // arr[0][0]: edit distance of the trailing penalty tokens, :25
  if (REAL) origin = (float)__ldg(rs.dist + rs.gap1[s] * rs.n_itok + rs.gap2[p]);
  if (lane == 0) {  // row 0: arr[0][j] = arr[0][j-1] + w*ins + sn2[j] (+ idf penalty), :33-39
    float v = origin;
    for (int j = 1; j <= p; j++) {
      v = __fadd_rn(v, insw);
      if (REAL) v = __fadd_rn(v, (float)__ldg(rs.dist + rs.gap2[j]));
      v = __fadd_rn(v, s_pen[j - 1]);
      s_up[((j - 1) % c) * 32 + (j - 1) / c] = v;
    }
  }
  __syncwarp();
  float diag = lane == 0 ? origin : s_up[(c - 1) * 32 + (lane - 1)];  // arr[0][lane*c]
  float col0 = origin;
  float recv_left = 0.f, recv_min = FLT_MAX;
  float K = -FLT_MAX, C = 0.f;
  const int steps = s + nl - 1;
  const int jbase = lane * c;
  for (int t = 1; t <= steps; t++) {
    const int i = t - lane;
    const bool active = lane < nl && i >= 1 && i <= s;
    float left = recv_left, rmin = recv_min;
    if (active) {
      if (lane == 0) {  // arr[i][0] = arr[i-1][0] + w*del + sn1[i], :28-32
        diag = col0;
        col0 = __fadd_rn(col0, delw);
        if (REAL) col0 = __fadd_rn(col0, (float)__ldg(rs.dist + rs.gap1[i] * rs.n_itok));
        left = col0;
        rmin = FLT_MAX;
      }
      const float next_diag = left;
      const int tok = s_sent[i - 1];
      int r1 = 0, ga_up = 0, ga = 0;
      if (REAL) {
        r1 = rs.real1[i - 1];
        ga_up = rs.gap1[i - 1] * rs.n_itok;  // row of the table for the gap above / on this row
        ga = rs.gap1[i] * rs.n_itok;
      }
      float d = left;
      for (int r = 0; r < c; r++) {
        const int j = jbase + r;  // 0-based column; the cell is (i, j+1)
        if (j >= p) break;
        const float up = s_up[r * 32 + lane];
        const float pen = s_pen[j];
        float a = __fadd_rn(up, delw);
        float bb = __fadd_rn(left, insw);
        float diff = (tok != s_pat[j]) ? __fadd_rn(repw, pen) : 0.f;
        if (REAL && tok == s_pat[j] && r1 != rs.real2[j]) diff = (r1 & 1) ? __fmul_rn(rs.rep, 1.f) : __fmul_rn(rs.rep, 2.f);
        float cc = __fadd_rn(diag, diff);
        if (REAL) {
          a = __fadd_rn(a, (float)__ldg(rs.dist + ga_up + rs.gap2[j + 1]));   // cost_tag[i-1][j]
          bb = __fadd_rn(bb, (float)__ldg(rs.dist + ga + rs.gap2[j]));        // cost_tag[i][j-1]
          cc = __fadd_rn(cc, (float)__ldg(rs.dist + ga_up + rs.gap2[j]));     // cost_tag[i-1][j-1]
        }
        bb = __fadd_rn(bb, pen);
        d = fminf(fminf(a, bb), cc);
        s_up[r * 32 + lane] = d;
        diag = up;
        left = d;
        rmin = fminf(rmin, d);
      }
      if (lane == nl - 1) {
        K = fmaxf(K, rmin);
        if (i == s) C = d;
      }
      diag = next_diag;  // arr[i][lane*c] is the diagonal of the next row
    }
    recv_left = __shfl_up_sync(FULL, left, 1);
    recv_min = __shfl_up_sync(FULL, rmin, 1);
  }
  C_out = __shfl_sync(FULL, C, nl - 1);
  K_out = __shfl_sync(FULL, K, nl - 1);
  __syncwarp();
}

// Same recurrence for short patterns (p <= 32), one THREAD per pair: the whole DP row, the pattern
// and the IDF penalties live in registers (fully unrolled over the 32 possible columns), no
// shuffles and no idle lanes -- ~30x fewer issue slots per pair than the warp wavefront, which
// remains the path for p > 32. Identical float operations in identical order.
template <bool IDF>
__device__ __forceinline__ void thread_edit_distance(const int32_t* __restrict__ sent, int s, const int32_t* __restrict__ pat,
                                                     int p, const float* __restrict__ idf, float idf_weight, float delw,
                                                     float insw, float repw, float& C_out, float& K_out) {
  int pt[32];
  float pn[32], row[32];
#pragma unroll
  for (int j = 0; j < 32; j++) {
    pt[j] = j < p ? __ldg(pat + j) : -1;
    pn[j] = (IDF && j < p) ? __fmul_rn(__ldg(idf + pt[j]), idf_weight) : 0.f;
  }
  float v = 0.f;
#pragma unroll
  for (int j = 0; j < 32; j++) {
    if (j < p) v = IDF ? __fadd_rn(__fadd_rn(v, insw), pn[j]) : __fadd_rn(v, insw);
    row[j] = v;
  }
  float col0 = 0.f, K = -FLT_MAX;
  for (int i = 0; i < s; i++) {
    const int tok = __ldg(sent + i);
    float diag = col0;
    col0 = __fadd_rn(col0, delw);
    float left = col0, rmin = FLT_MAX;
#pragma unroll
    for (int j = 0; j < 32; j++) {
      if (j < p) {
        const float up = row[j];
        const float a = __fadd_rn(up, delw);
        const float bb = IDF ? __fadd_rn(__fadd_rn(left, insw), pn[j]) : __fadd_rn(left, insw);
        const float diff = (tok != pt[j]) ? (IDF ? __fadd_rn(repw, pn[j]) : repw) : 0.f;
        const float cc = __fadd_rn(diag, diff);
        const float d = fminf(fminf(a, bb), cc);
        row[j] = d;
        diag = up;
        left = d;
        rmin = fminf(rmin, d);
      }
    }
    K = fmaxf(K, rmin);
  }
  float C = 0.f;  // row[p - 1] as a chain of selects: a dynamic index would put the row into local memory
#pragma unroll
  for (int j = 0; j < 32; j++) C = (j < p) ? row[j] : C;
  C_out = C;
  K_out = K;
}

__device__ __forceinline__ void write_record(const IndexDev& ix, const BatchDev& b, const SurvRec& sr, int slen, float C, float K) {
  fm_record r;
  r.s_id = (uint32_t)ix.sid_at[sr.start >> 2] + ix.sid_base;
  r.longest_match = (int32_t)b.hlm[sr.hslot];
  r.length = slen;
  r.cost = C;
  r.rowmin_max = K;
  r.reserved[0] = sr.start;
  r.reserved[1] = 0;
  r.reserved[2] = 0;
  b.rec[b.q_base[sr.q] + sr.j] = r;
}

// One thread per surviving (query, sentence) with p <= 32.
template <bool IDF>
__global__ void __launch_bounds__(128) fm_score_short_kernel(IndexDev ix, BatchDev b, Params pr) {
  const long long n = min((long long)b.ctr->n_surv, (long long)b.surv_cap);
  if (b.ctr->overflow) return;
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n) return;
  const SurvRec sr = b.surv[w];
  const int slen = b.surv_len[w];
  const QMeta qm = b.qmeta[sr.q];
  const int p = qm.x;
  if (p > 32) {  // left to the wavefront kernel, which only scans the survivors when this flag is set
    b.ctr->n_long = 1;
    return;
  }
  const float wdiff = __fdiv_rn(100.f, normalizer(p, slen, pr));
  const float idf_weight = __fdiv_rn(__fmul_rn(wdiff, pr.idf_penalty), pr.idf_penalty != 0.f ? ix.idf_max : 0.01f);
  float C, K;
  thread_edit_distance<IDF>(ix.tok + sr.start, slen, b.pat + qm.z, p, ix.idf, idf_weight, __fmul_rn(pr.del, wdiff),
                            __fmul_rn(pr.ins, wdiff), __fmul_rn(pr.rep, wdiff), C, K);
  write_record(ix, b, sr, slen, C, K);
}

// One warp per surviving (query, sentence) with p > 32: Costs (include/fuzzy/costs.hh:54-57), idf
// weight (src/fuzzy_match.cc:591) and the full edit distance without upper bound; writes the record
// at the candidate's slot inside its query group.
template <bool REAL>
__global__ void __launch_bounds__(256) fm_score_kernel(IndexDev ix, BatchDev b, Params pr, int stride, int min_p, unsigned long_mask) {
  extern __shared__ int smem[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int n_arr = REAL ? 8 : 4;
  int32_t* s_sent = smem + wib * n_arr * stride;
  int32_t* s_pat = s_sent + stride;
  float* s_pen = reinterpret_cast<float*>(s_pat + stride);
  float* s_up = s_pen + stride;
  int32_t* s_real1 = reinterpret_cast<int32_t*>(s_up + stride);
  int32_t* s_gap1 = s_real1 + stride;
  int32_t* s_real2 = s_gap1 + stride;
  int32_t* s_gap2 = s_real2 + stride;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long n = min((long long)b.ctr->n_surv, (long long)b.surv_cap);
  if (b.ctr->overflow) return;
  if (min_p > 0 && (b.ctr->n_long & long_mask) == 0) return;  // nothing for the wavefront: the earlier kernels took every pair
  for (long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n; w += n_warps) {
    const SurvRec sr = b.surv[w];
    const int slen = b.surv_len[w];
    const QMeta qm = b.qmeta[sr.q];
    const int p = qm.x;
    if (p < min_p) continue;  // short patterns are scored by fm_score_short_kernel
    const float norm = normalizer(p, slen, pr);
    const float wdiff = __fdiv_rn(100.f, norm);
    const float idf_weight = __fdiv_rn(__fmul_rn(wdiff, pr.idf_penalty), pr.idf_penalty != 0.f ? ix.idf_max : 0.01f);
    for (int k = lane; k < slen; k += 32) s_sent[k] = ix.tok[sr.start + k];
    for (int k = lane; k < p; k += 32) {
      const int t = b.pat[qm.z + k];
      s_pat[k] = t;
      s_pen[k] = idf_weight != 0.f ? __fmul_rn(ix.idf[t], idf_weight) : 0.f;
    }
    RealSide rs{};
    if (REAL) {
      for (int k = lane; k < slen; k += 32) s_real1[k] = ix.real[sr.start + k];
      for (int k = lane; k <= slen; k += 32) s_gap1[k] = ix.gap[sr.start + k];
      for (int k = lane; k < p; k += 32) s_real2[k] = b.q_real[qm.z + k];
      for (int k = lane; k <= p; k += 32) s_gap2[k] = b.q_gap[qm.z + sr.q + k];
      rs.real1 = s_real1; rs.gap1 = s_gap1; rs.real2 = s_real2; rs.gap2 = s_gap2;
      rs.dist = b.itok_dist; rs.n_itok = b.n_itok; rs.rep = pr.rep;
    }
    __syncwarp();
    float C, K;
    warp_edit_distance<REAL>(s_sent, slen, s_pat, p, s_pen, s_up, __fmul_rn(pr.del, wdiff), __fmul_rn(pr.ins, wdiff),
                             __fmul_rn(pr.rep, wdiff), C, K, rs);
    if (lane == 0) write_record(ix, b, sr, slen, C, K);
  }
}

// ---------------------------------------------------------------- edit distance, bit-parallel
//
// When insert, delete and replace cost the same positive amount c and there are no IDF / real-token /
// penalty-token terms, every cell of the reference's float DP (src/edit_distance.cc:41-75) is the float
// sum of k copies of m = c * w added one at a time from 0 (each transition adds m or 0, min is exact), so
// arr[i][j] = acc(D[i][j]) with D the integer Levenshtein distance and acc(k) = ((m + m) + m) ... k times
// (SURVEY.md 3.1 Q3). D comes from Myers' bit-vector recurrence (Myers 1999, Hyyro 2003; global variant:
// the top boundary grows by one per row) with the pattern along the bits: 32 (or 64) cells per ~17 integer
// instructions. The early-exit value K = max_i min_{j>=1} arr[i][j] never exceeds the final cost for such
// costs (row i's minimum is at most arr[i][1] <= arr[i][0] + m = arr[i+1][0] <= arr[s][p]), and the replay
// only uses max(K, C), so K = C is recorded.

__device__ __forceinline__ float chain_cost(int k, float m) {  // acc(k)
  float v = 0.f;
  for (int i = 0; i < k; i++) v = __fadd_rn(v, m);
  return v;
}

// One row of the recurrence for a whole bit vector (horizontal input +1: the global top boundary);
// returns the horizontal delta leaving at bit `high`.
template <typename W>
__device__ __forceinline__ int myers_row(W eq, W& pv, W& mv, W high) {
  const W xv = eq | mv;
  const W xh = (((eq & pv) + pv) ^ pv) | eq;
  W ph = mv | ~(xh | pv);
  W mh = pv & xh;
  const int hout = (ph & high) ? 1 : ((mh & high) ? -1 : 0);
  ph = (ph << 1) | 1;
  mh <<= 1;
  pv = mh | ~(xv | ph);
  mv = ph & xv;
  return hout;
}

// position mask of word w in the query's pattern (0 when the pattern does not contain it)
__device__ __forceinline__ unsigned long long peq_lookup(const int2* __restrict__ tbl, int tmask,
                                                         const unsigned long long* __restrict__ peq, int w, int2 e, int h) {
  for (;;) {
    if (e.x == w) return __ldg(peq + (e.y & 0xffff));
    if (e.x == -1) return 0ull;
    h = (h + 1) & tmask;
    e = __ldg(tbl + h);
  }
}

// p <= 32: the pattern sits in registers and the match vector of a sentence token is built by comparing
// against all of it (no table lookups, no dependent loads: the only memory traffic is the sentence).
__device__ __forceinline__ int myers_thread32(const int32_t* __restrict__ sent, int s, const int32_t* __restrict__ pat, int p) {
  int pt[32];
#pragma unroll
  for (int j = 0; j < 32; j++) pt[j] = j < p ? __ldg(pat + j) : -1;
  uint32_t pv = ~0u, mv = 0;
  const uint32_t high = 1u << (p - 1);
  int score = p;
  const int4* s4 = reinterpret_cast<const int4*>(sent);
  int4 t = ldg_nc_v4(s4);  // sentences start on 16-byte boundaries and are zero padded
  for (int i0 = 0; i0 < s; i0 += 4) {
    const int4 cur = t;
    if (i0 + 4 < s) t = ldg_nc_v4(s4 + (i0 >> 2) + 1);
    uint32_t e0 = 0, e1 = 0, e2 = 0, e3 = 0;
#pragma unroll
    for (int j = 0; j < 32; j++) {
      const uint32_t bit = 1u << j;
      if (cur.x == pt[j]) e0 |= bit;
      if (cur.y == pt[j]) e1 |= bit;
      if (cur.z == pt[j]) e2 |= bit;
      if (cur.w == pt[j]) e3 |= bit;
    }
    score += myers_row<uint32_t>(e0, pv, mv, high);
    if (i0 + 1 < s) score += myers_row<uint32_t>(e1, pv, mv, high);
    if (i0 + 2 < s) score += myers_row<uint32_t>(e2, pv, mv, high);
    if (i0 + 3 < s) score += myers_row<uint32_t>(e3, pv, mv, high);
  }
  return score;
}

// 32 < p <= 64: one 64-bit vector; the match vectors come from the query's position masks (peq64, built
// by the prepare kernel) through its word table.
__device__ __forceinline__ int myers_thread64(const int32_t* __restrict__ sent, int s, int p, const int2* __restrict__ tbl, int tmask,
                                              const unsigned long long* __restrict__ peq) {
  typedef unsigned long long W;
  W pv = ~(W)0, mv = 0;
  const W high = (W)1 << (p - 1);
  int score = p;
  const int4* s4 = reinterpret_cast<const int4*>(sent);
  for (int i0 = 0; i0 < s; i0 += 4) {
    const int4 t = ldg_nc_v4(s4 + (i0 >> 2));
    // first probes of the four lookups in flight together
    const int h0 = hash32((uint32_t)t.x) & tmask, h1 = hash32((uint32_t)t.y) & tmask;
    const int h2 = hash32((uint32_t)t.z) & tmask, h3 = hash32((uint32_t)t.w) & tmask;
    const int2 e0 = __ldg(tbl + h0), e1 = __ldg(tbl + h1), e2 = __ldg(tbl + h2), e3 = __ldg(tbl + h3);
    const W q0 = peq_lookup(tbl, tmask, peq, t.x, e0, h0);
    const W q1 = peq_lookup(tbl, tmask, peq, t.y, e1, h1);
    const W q2 = peq_lookup(tbl, tmask, peq, t.z, e2, h2);
    const W q3 = peq_lookup(tbl, tmask, peq, t.w, e3, h3);
    score += myers_row<W>(q0, pv, mv, high);
    if (i0 + 1 < s) score += myers_row<W>(q1, pv, mv, high);
    if (i0 + 2 < s) score += myers_row<W>(q2, pv, mv, high);
    if (i0 + 3 < s) score += myers_row<W>(q3, pv, mv, high);
  }
  return score;
}

static const int kBpThreadMax = 64;  // thread per pair up to here (one 64-bit vector)
static const int kBpWarpMax = 320;   // warp per pair up to here (Peq table of the pair in shared memory)

// One thread per surviving (query, sentence) with p <= 64.
__global__ void __launch_bounds__(128) fm_score_bp_kernel(IndexDev ix, BatchDev b, Params pr) {
  const long long n = min((long long)b.ctr->n_surv, (long long)b.surv_cap);
  if (b.ctr->overflow) return;
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n) return;
  const SurvRec sr = b.surv[w];
  const int slen = b.surv_len[w];
  const QMeta qm = b.qmeta[sr.q];
  const int p = qm.x;
  if (p > kBpThreadMax) {  // left to the warp kernels, which only scan the survivors when their flag is set
    atomicOr(&b.ctr->n_long, p > kBpWarpMax ? 2u : 1u);
    return;
  }
  const int32_t* sent = ix.tok + sr.start;
  const int d = p <= 32 ? myers_thread32(sent, slen, b.pat + qm.z, p)
                        : myers_thread64(sent, slen, p, b.tbl + 4ll * qm.z, next_pow2(2 * p) - 1, b.peq64 + qm.z);
  const float wdiff = __fdiv_rn(100.f, normalizer(p, slen, pr));
  const float C = chain_cost(d, __fmul_rn(pr.del, wdiff));
  write_record(ix, b, sr, slen, C, C);
}

// One warp per surviving pair with 64 < p <= kBpWarpMax: lane l owns pattern positions [32 l, 32 l + 32)
// and works on row t - l at step t (a wavefront over the blocks: the horizontal delta leaving block l
// for row i enters block l + 1 one step later, one __shfl_up per step). The pair's position masks
// (distinct word x block) and the distinct index of every sentence token are staged in shared memory.
__global__ void __launch_bounds__(128) fm_score_bpw_kernel(IndexDev ix, BatchDev b, Params pr, int peq_words, int didx_words) {
  extern __shared__ uint32_t bsm[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint32_t* s_peq = bsm + (size_t)wib * (peq_words + didx_words);
  uint16_t* s_didx = reinterpret_cast<uint16_t*>(s_peq + peq_words);
  if (b.ctr->overflow || !(b.ctr->n_long & 1u)) return;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long n = min((long long)b.ctr->n_surv, (long long)b.surv_cap);
  for (long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n; w += n_warps) {
    const SurvRec sr = b.surv[w];
    const QMeta qm = b.qmeta[sr.q];
    const int p = qm.x;
    if (p <= kBpThreadMax || p > kBpWarpMax) continue;
    const int slen = b.surv_len[w];
    const int nw = (p + 31) >> 5;
    const int2* tbl = b.tbl + 4ll * qm.z;
    const int tmask = next_pow2(2 * p) - 1;
    auto didx_of = [&](int word) -> int {
      int h = hash32((uint32_t)word) & tmask;
      for (;;) {
        const int2 e = __ldg(tbl + h);
        if (e.x == word) return e.y & 0xffff;
        if (e.x == -1) return 0xffff;
        h = (h + 1) & tmask;
      }
    };
    for (int k = lane; k < p * nw; k += 32) s_peq[k] = 0;
    __syncwarp();
    for (int j = lane; j < p; j += 32) atomicOr(&s_peq[didx_of(b.pat[qm.z + j]) * nw + (j >> 5)], 1u << (j & 31));
    for (int i = lane; i < slen; i += 32) s_didx[i] = (uint16_t)didx_of(__ldg(ix.tok + sr.start + i));
    __syncwarp();
    uint32_t pv = ~0u, mv = 0;
    const uint32_t high = lane == nw - 1 ? 1u << ((p - 1) & 31) : 0x80000000u;
    int hout = 0, score = p;
    const int steps = slen + nw - 1;
    for (int t = 1; t <= steps; t++) {
      int hin = __shfl_up_sync(FULL, hout, 1);
      if (lane == 0) hin = 1;
      const int i = t - lane;
      if (lane < nw && i >= 1 && i <= slen) {
        const int d = s_didx[i - 1];
        uint32_t eq = d != 0xffff ? s_peq[d * nw + lane] : 0u;
        const uint32_t xv = eq | mv;
        if (hin < 0) eq |= 1u;
        const uint32_t xh = (((eq & pv) + pv) ^ pv) | eq;
        uint32_t ph = mv | ~(xh | pv);
        uint32_t mh = pv & xh;
        hout = (ph & high) ? 1 : ((mh & high) ? -1 : 0);
        ph <<= 1;
        mh <<= 1;
        if (hin < 0) mh |= 1u;
        else if (hin > 0) ph |= 1u;
        pv = mh | ~(xv | ph);
        mv = ph & xv;
        if (lane == nw - 1) score += hout;
      }
    }
    score = __shfl_sync(FULL, score, nw - 1);
    if (lane == 0) {
      const float wdiff = __fdiv_rn(100.f, normalizer(p, slen, pr));
      const float C = chain_cost(score, __fmul_rn(pr.del, wdiff));
      write_record(ix, b, sr, slen, C, C);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------- replay

__device__ __forceinline__ unsigned long long order_key(const fm_record& r) {  // lm desc, s_id asc
  return ((unsigned long long)(unsigned)(0x7fffffff - r.longest_match) << 32) | r.s_id;
}
// Order n <= 32 records held one per lane: rank = number of smaller keys (keys are distinct).
__device__ __forceinline__ int warp_rank(unsigned long long key, int n) {
  int rank = 0;
  for (int j = 0; j < n; j++) rank += __shfl_sync(FULL, key, j) < key;
  return rank;
}

// std::priority_queue<float> lowest_costs (src/fuzzy_match.cc:567-568)
__device__ __forceinline__ void heap_push(float* h, int& n, float v) {
  int i = n++;
  while (i > 0 && h[(i - 1) >> 1] < v) { h[i] = h[(i - 1) >> 1]; i = (i - 1) >> 1; }
  h[i] = v;
}
__device__ __forceinline__ void heap_pop(float* h, int& n) {
  const float v = h[--n];
  int i = 0;
  for (;;) {
    int c = 2 * i + 1;
    if (c >= n) break;
    if (c + 1 < n && h[c + 1] > h[c]) c++;
    if (!(h[c] > v)) break;
    h[i] = h[c];
    i = c;
  }
  if (n > 0) h[i] = v;
}

__device__ __forceinline__ fm_match to_match(const fm_record& r, float penalty) {
  fm_match m;
  m.s_id = r.s_id; m.score = r.rowmin_max; m.penalty = penalty; m.max_subseq = r.longest_match;
  m.length = r.length; m.cost = r.cost;
  return m;
}

// CTA-cooperative version of warp_sort_pairs (keys/idx in shared or global memory).
__device__ void block_sort_pairs(unsigned long long* keys, int32_t* idx, int n) {
  int np = 1;
  while (np < n) np <<= 1;
  for (int k = 2; k <= np; k <<= 1) {
    for (int j = k >> 1, first = 1; j > 0; j >>= 1, first = 0) {
      for (int i = threadIdx.x; i < np; i += blockDim.x) {
        const int l = first ? (i ^ (k - 1)) : (i ^ j);
        if (l > i && l < n) {
          const unsigned long long a = keys[i], c = keys[l];
          if (a > c) {
            keys[i] = c; keys[l] = a;
            const int t = idx[i]; idx[i] = idx[l]; idx[l] = t;
          }
        }
      }
      __syncthreads();
    }
  }
}

// The sequential heart of the candidate loop (src/fuzzy_match.cc:567-611), run by one warp over the
// records seg[idx[0..n)] given in the reference's order. Lanes prefetch 32 records at a time and
// hand them to lane 0, which owns the bound heap. A candidate is dropped where the reference's
// bounded edit distance would have exited early or exceeded the bound: max(K, C) > bound.
// On return idx[0..nacc) / keys[0..nacc) hold the accepted records and their result keys
// (score desc, s_id asc; CompareMatch :25-33); rowmin_max of an accepted record now holds its score.
struct WireOut {   // shard mode of the replay kernels: where the locally accepted records go (cnt == NULL: off)
  int32_t* cnt;    // [n_q] accepted records of each query
  fm_wire* stage;  // accepted records of query q at stage[q_base[q] ...] (never more than its scored candidates)
};
__device__ __forceinline__ fm_wire to_wire(const fm_record& r) {
  fm_wire w;
  w.s_id = r.s_id;
  w.lm_len = (uint32_t)r.longest_match | ((uint32_t)r.length << 16);
  w.cost = r.cost;
  w.rowmin_max = r.rowmin_max;
  return w;
}
__device__ int replay_sequence(fm_record* seg, int n, int p, int32_t* idx, unsigned long long* keys, float* heap,
                               const Params& pr, fm_wire* wire = nullptr, int wire_cap = 0, int* wire_n = nullptr) {
  const int lane = threadIdx.x & 31;
  int hn = 0, nacc = 0, nwire = 0;
  if (lane == 0) heap_push(heap, hn, FLT_MAX);
  float bound = FLT_MAX;
  for (int c0 = 0; c0 < n; c0 += 32) {
    const bool have = c0 + lane < n;
    const int my = have ? idx[c0 + lane] : 0;
    const fm_record r = seg[my];
    __syncwarp();
    int pos = 0;  // candidates before pos are done
    for (;;) {
      // every lane tests its candidate against the current bound; skipped candidates have no side
      // effect (src/fuzzy_match.cc:595), so only the first one that passes needs the sequential step
      const bool pass = have && lane >= pos && !(r.rowmin_max > bound || r.cost > bound) &&
                        !(pr.no_perfect && r.cost == 0.f && r.length == p);
      const unsigned bal = __ballot_sync(FULL, pass);
      if (!bal) break;
      const int t = __ffs(bal) - 1;
      const float cost = __shfl_sync(FULL, r.cost, t);
      const int id = __shfl_sync(FULL, my, t);
      const unsigned sid = __shfl_sync(FULL, r.s_id, t);
      if (lane == 0) {
        const float score = score_of(cost);
        heap_push(heap, hn, cost);
        if (score < pr.fuzzy || (pr.buffer > 0 && hn > pr.buffer)) heap_pop(heap, hn);
        if (wire) {  // shard mode: every candidate the loop accepts, in the order it accepts them; nothing else is kept
          if (nwire < wire_cap) wire[nwire] = to_wire(seg[id]);
          nwire++;
        } else if (score >= pr.fuzzy) {
          seg[id].rowmin_max = score;  // slot reused: score
          seg[id].reserved[1] = 0;     // contrastive accumulator
          seg[id].reserved[2] = 0;     // contrastive "selected" flag
          const unsigned u = __float_as_uint(score);
          const unsigned ord = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
          keys[nacc] = ((unsigned long long)(~ord) << 32) | sid;
          idx[nacc] = id;  // nacc <= c0 + t: this slot has already been consumed
          nacc++;
        }
        bound = heap[0];
      }
      bound = __shfl_sync(FULL, bound, 0);
      pos = t + 1;
    }
    __syncwarp();
  }
  if (wire_n && lane == 0) *wire_n = nwire;
  return __shfl_sync(FULL, nacc, 0);
}

// The same loop over a list prepared for it: packed[i] = (bits of max(K, C), or +inf for a candidate that
// no_perfect skips) << 32 | record index, in the reference's candidate order, in shared memory or as one
// coalesced global array. The bound test of 32 candidates is then one read and one compare per lane -- no
// dependent record fetch per block (what made long lists slow) -- and only accepted candidates touch their record.
// packed may alias keys_out / idx_out is written at positions already consumed, as above.
__device__ __forceinline__ unsigned long long pack_candidate(const fm_record& r, int id, int p, const Params& pr) {
  float m = fmaxf(r.rowmin_max, r.cost);
  if (pr.no_perfect && r.cost == 0.f && r.length == p) m = __int_as_float(0x7f800000);  // never passes
  return ((unsigned long long)__float_as_uint(m) << 32) | (unsigned)id;
}
__device__ int replay_sequence_packed(fm_record* seg, int n, const unsigned long long* packed, int32_t* idx_out,
                                      unsigned long long* keys_out, float* heap, const Params& pr, fm_wire* wire = nullptr,
                                      int wire_cap = 0, int* wire_n = nullptr) {
  const int lane = threadIdx.x & 31;
  int hn = 0, nacc = 0, nwire = 0;
  if (lane == 0) heap_push(heap, hn, FLT_MAX);
  float bound = FLT_MAX;
  unsigned long long nxt = lane < n ? packed[lane] : 0ull;
  for (int c0 = 0; c0 < n; c0 += 32) {
    const bool have = c0 + lane < n;
    const unsigned long long pk = nxt;
    if (c0 + 32 + lane < n) nxt = packed[c0 + 32 + lane];  // next block in flight while this one is replayed
    const float m = __uint_as_float((unsigned)(pk >> 32));
    __syncwarp();
    int pos = 0;  // candidates before pos are done
    for (;;) {
      const bool pass = have && lane >= pos && !(m > bound);
      const unsigned bal = __ballot_sync(FULL, pass);
      if (!bal) break;
      const int t = __ffs(bal) - 1;
      const int id = (int)__shfl_sync(FULL, (unsigned)pk, t);
      if (lane == 0) {
        const fm_record r = seg[id];
        const float score = score_of(r.cost);
        heap_push(heap, hn, r.cost);
        if (score < pr.fuzzy || (pr.buffer > 0 && hn > pr.buffer)) heap_pop(heap, hn);
        if (wire) {
          if (nwire < wire_cap) wire[nwire] = to_wire(r);
          nwire++;
        } else if (score >= pr.fuzzy) {
          seg[id].rowmin_max = score;  // slot reused: score
          seg[id].reserved[1] = 0;     // contrastive accumulator
          seg[id].reserved[2] = 0;     // contrastive "selected" flag
          const unsigned u = __float_as_uint(score);
          const unsigned ord = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
          keys_out[nacc] = ((unsigned long long)(~ord) << 32) | r.s_id;
          idx_out[nacc] = id;
          nacc++;
        }
        bound = heap[0];
      }
      bound = __shfl_sync(FULL, bound, 0);
      pos = t + 1;
    }
    __syncwarp();
  }
  if (wire_n && lane == 0) *wire_n = nwire;
  return __shfl_sync(FULL, nacc, 0);
}

// Warp-cooperative ascending bitonic sort of 64-bit keys in shared memory (ascending comparators only,
// so the virtual +inf padding up to the next power of two never moves).
__device__ void warp_sort_keys(unsigned long long* keys, int n) {
  const int lane = threadIdx.x & 31;
  int np = 32;
  while (np < n) np <<= 1;
  for (int k = 2; k <= np; k <<= 1) {
    for (int j = k >> 1, first = 1; j > 0; j >>= 1, first = 0) {
      for (int i = lane; i < np; i += 32) {
        const int l = first ? (i ^ (k - 1)) : (i ^ j);
        if (l > i && l < n) {
          const unsigned long long a = keys[i], c = keys[l];
          if (a > c) { keys[i] = c; keys[l] = a; }
        }
      }
      __syncwarp();
    }
  }
}

// One THREAD per query: queries with no scored candidate or exactly one (the great majority at the
// headline configuration) are finished here -- with a single candidate the bound is still FLT_MAX, so
// the loop of src/fuzzy_match.cc:570-611 reduces to the no_perfect / score tests. Queries with 2..kWarpMax
// candidates are queued for fm_replay_kernel (a warp each), longer ones for fm_replay_heavy_kernel.
__global__ void __launch_bounds__(256) fm_replay_small_kernel(fm_record* rec, const int32_t* __restrict__ q_cnt,
                                                              const int32_t* __restrict__ q_base, int32_t* sort_idx,
                                                              int32_t* acc_cnt, int32_t* mid_q, int32_t* heavy_q,
                                                              const int32_t* __restrict__ q_off, int n_q, Params pr, long long cap,
                                                              fm_match* out, int32_t* out_count, Counters* ctr, int warp_max, WireOut wo) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_q) return;
  if (ctr->overflow) return;
  const int n = q_cnt[q];
  if (n > 1) {
    if (n <= warp_max) mid_q[atomicAdd(&ctr->n_mid, 1u)] = q;
    else heavy_q[atomicAdd(&ctr->n_heavy, 1u)] = q;
    return;
  }
  int nacc = 0;
  if (wo.cnt) {  // shard mode: the single candidate is accepted unless no_perfect skips it (the bound is still FLT_MAX)
    int c = 0;
    if (n == 1) {
      const fm_record r = rec[q_base[q]];
      const int p = q_off[q + 1] - q_off[q];
      if (!(r.rowmin_max > FLT_MAX || r.cost > FLT_MAX) && !(pr.no_perfect && r.cost == 0.f && r.length == p)) {
        wo.stage[q_base[q]] = to_wire(r);
        c = 1;
      }
    }
    wo.cnt[q] = c;
    return;
  }
  if (n == 1) {
    const int base = q_base[q];
    fm_record r = rec[base];
    const int p = q_off[q + 1] - q_off[q];
    const float bound = FLT_MAX;
    if (!(r.rowmin_max > bound || r.cost > bound) && !(pr.no_perfect && r.cost == 0.f && r.length == p)) {
      const float score = score_of(r.cost);
      if (score >= pr.fuzzy) {
        r.rowmin_max = score;  // slot reused: score
        r.reserved[1] = 0;
        r.reserved[2] = 0;
        rec[base] = r;
        sort_idx[base] = 0;
        nacc = 1;
      }
    }
    if (pr.contrast <= 0.f && nacc && cap > 0) out[(long long)q * cap] = to_match(r, 0.f);
  }
  if (pr.contrast > 0.f) {
    acc_cnt[q] = nacc;
    if (n == 0) out_count[q] = 0;
    return;
  }
  const int want = pr.n_matches == 0 ? nacc : min(nacc, pr.n_matches);
  out_count[q] = want;
  if (want) atomicAdd(&ctr->n_matches, 1u);
}

// One warp per query with <= kWarpMax scored candidates: candidate order (ngram_matches.cc:20-29:
// longest match desc, s_id asc) by register rank sort (<= 32) or a shared-memory bitonic sort of packed
// keys, the replay, result order, top-N output (:670-679). Larger queries are queued for
// fm_replay_heavy_kernel.
#ifndef FM_KWARPMAX
#define FM_KWARPMAX 128
#endif
static const int kWarpMax = FM_KWARPMAX;
__global__ void __launch_bounds__(256) fm_replay_kernel(fm_record* rec, const int32_t* __restrict__ q_cnt,
                                                        const int32_t* __restrict__ q_base, float* heapbuf,
                                                        unsigned long long* sort_key, int32_t* sort_idx, int32_t* acc_cnt,
                                                        const int32_t* __restrict__ mid_q, const int32_t* __restrict__ q_off,
                                                        Params pr, long long cap, fm_match* out, int32_t* out_count,
                                                        Counters* ctr, WireOut wo) {
  __shared__ float s_heap[8][64];
  __shared__ unsigned long long s_keys[8][kWarpMax];
  const int lane = threadIdx.x & 31;
  if (ctr->overflow) return;
  const int n_mid = (int)ctr->n_mid;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < n_mid; m += n_warps) {
  const int q = mid_q[m];
  const int n = q_cnt[q];
  const int p = q_off[q + 1] - q_off[q];
  const int base = q_base[q];
  fm_record* seg = rec + base;
  unsigned long long* keys = sort_key + base;
  int32_t* idx = sort_idx + base;
  if (n <= 32) {
    const unsigned long long k = lane < n ? order_key(seg[lane]) : ~0ull;
    const int r = warp_rank(k, n);
    if (lane < n) idx[r] = lane;
  } else {
    unsigned long long* sk = s_keys[threadIdx.x >> 5];
    for (int i = lane; i < n; i += 32) {
      const fm_record r = seg[i];
      sk[i] = ((unsigned long long)(unsigned)(1023 - r.longest_match) << 52) | ((unsigned long long)r.s_id << 20) | (unsigned)i;
    }
    __syncwarp();
    warp_sort_keys(sk, n);
    for (int i = lane; i < n; i += 32) {  // in place: sorted key -> (bound-test value, record index)
      const int id = (int)(sk[i] & 0xfffffu);
      sk[i] = pack_candidate(seg[id], id, p, pr);
    }
  }
  __syncwarp();
  const unsigned long long* packed = n > 32 ? s_keys[threadIdx.x >> 5] : nullptr;
  float* heap = (pr.buffer > 0 && pr.buffer <= 62) ? s_heap[threadIdx.x >> 5] : heapbuf + base + q;
  if (wo.cnt) {  // shard mode: only the accepted records travel; the result order is made after the merge
    if (packed) replay_sequence_packed(seg, n, packed, idx, keys, heap, pr, wo.stage + base, n, wo.cnt + q);
    else replay_sequence(seg, n, p, idx, keys, heap, pr, wo.stage + base, n, wo.cnt + q);
    __syncwarp();
    continue;
  }
  const int nacc = packed ? replay_sequence_packed(seg, n, packed, idx, keys, heap, pr) : replay_sequence(seg, n, p, idx, keys, heap, pr);
  if (nacc > 1) {
    if (nacc <= 32) {
      const unsigned long long k = lane < nacc ? keys[lane] : ~0ull;
      const int id = lane < nacc ? idx[lane] : 0;
      const int r = warp_rank(k, nacc);
      __syncwarp();
      if (lane < nacc) idx[r] = id;
    } else {
      // (score, s_id) keys do not fit one word with the index: rank every accepted record against all
      unsigned long long* sk = s_keys[threadIdx.x >> 5];
      for (int i = lane; i < nacc; i += 32) sk[i] = keys[i];
      __syncwarp();
      int my_rank[kWarpMax / 32], my_id[kWarpMax / 32];
#pragma unroll
      for (int c = 0; c < kWarpMax / 32; c++) {
        const int i = c * 32 + lane;
        my_rank[c] = 0;
        my_id[c] = i < nacc ? idx[i] : 0;
        if (i < nacc) {
          const unsigned long long k = sk[i];
          for (int j = 0; j < nacc; j++) my_rank[c] += sk[j] < k;
        }
      }
      __syncwarp();
#pragma unroll
      for (int c = 0; c < kWarpMax / 32; c++)
        if (c * 32 + lane < nacc) idx[my_rank[c]] = my_id[c];
    }
    __syncwarp();
  }
  if (pr.contrast > 0.f) {
    if (lane == 0) acc_cnt[q] = nacc;
    __syncwarp();
    continue;
  }
  const int want = pr.n_matches == 0 ? nacc : min(nacc, pr.n_matches);
  for (int k = lane; k < want && k < cap; k += 32) out[(long long)q * cap + k] = to_match(seg[idx[k]], 0.f);
  if (lane == 0) {
    out_count[q] = want;
    if (want) atomicAdd(&ctr->n_matches, (unsigned)want);
  }
  __syncwarp();
  }
}

// CTA-cooperative ascending bitonic sort of 64-bit keys (ascending comparators only, so the virtual
// +inf padding up to the next power of two never moves).
__device__ void block_sort_keys(unsigned long long* keys, int n) {
  int np = 1;
  while (np < n) np <<= 1;
  for (int k = 2; k <= np; k <<= 1) {
    for (int j = k >> 1, first = 1; j > 0; j >>= 1, first = 0) {
      for (int i = threadIdx.x; i < np; i += blockDim.x) {
        const int l = first ? (i ^ (k - 1)) : (i ^ j);
        if (l > i && l < n) {
          const unsigned long long a = keys[i], c = keys[l];
          if (a > c) { keys[i] = c; keys[l] = a; }
        }
      }
      __syncthreads();
    }
  }
}

// CTA-cooperative stable LSD radix sort (8-bit digits over key bits [lo_bit, hi_bit)) of 64-bit keys
// in global memory, for candidate lists too long for shared memory. Every warp owns a contiguous
// chunk: per-warp digit histograms, one scan over (digit, warp), then each warp scatters its chunk in
// order with intra-warp ranks from __match_any_sync. Passes whose digit is constant are skipped.
// Returns the buffer that holds the sorted keys.
__device__ unsigned long long* block_radix_sort_keys(unsigned long long* a, unsigned long long* b, int n, int lo_bit,
                                                     int hi_bit, int (*s_off)[256], int* s_tot, int* s_flag) {
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int per_warp = (((n + 7) / 8) + 31) / 32 * 32;
  const int beg = min(n, w * per_warp), end = min(n, beg + per_warp);
  for (int shift = lo_bit; shift < hi_bit; shift += 8) {
    for (int k = 0; k < 8; k++) s_off[k][t] = 0;
    if (t == 0) *s_flag = 0;
    __syncthreads();
    for (int i = beg + lane; i < end; i += 32) atomicAdd(&s_off[w][(int)((a[i] >> shift) & 255)], 1);
    __syncthreads();
    {  // thread t owns digit t
      int tot = 0;
      for (int k = 0; k < 8; k++) tot += s_off[k][t];
      if (tot == n) *s_flag = 1;
      int incl = tot;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += o;
      }
      if (lane == 31) s_tot[w] = incl;
      __syncthreads();
      int base = incl - tot;
      for (int k = 0; k < w; k++) base += s_tot[k];
      for (int k = 0; k < 8; k++) {
        const int c = s_off[k][t];
        s_off[k][t] = base;
        base += c;
      }
    }
    __syncthreads();
    if (*s_flag) continue;  // all keys share this digit
    for (int i0 = beg; i0 < end; i0 += 32) {
      const int i = i0 + lane;
      const bool valid = i < end;
      const unsigned long long key = valid ? a[i] : 0;
      const int d = valid ? (int)((key >> shift) & 255) : 256 + lane;
      const unsigned m = __match_any_sync(FULL, d);
      const int rank = __popc(m & ((1u << lane) - 1));
      if (valid) b[s_off[w][d] + rank] = key;
      __syncwarp();
      if (valid && rank == 0) s_off[w][d] += __popc(m);
      __syncwarp();
    }
    __syncthreads();
    unsigned long long* tmp = a; a = b; b = tmp;
  }
  return a;
}

// One CTA per query with more than kWarpMax scored candidates. The candidate order is sorted as ONE packed
// 64-bit key per record -- (1023 - match length) << 52 | s_id << 20 | record index -- in shared
// memory (up to kHeavySmem records, else in global scratch); warp 0 then runs the replay.
static const int kHeavySmem = 4096;  // 32 KB of sort keys per CTA: several CTAs per SM; longer lists are sorted in global memory
__global__ void __launch_bounds__(256) fm_replay_heavy_kernel(fm_record* rec, const int32_t* __restrict__ q_cnt,
                                                              const int32_t* __restrict__ q_base, float* heapbuf,
                                                              unsigned long long* sort_key, unsigned long long* sort_key2,
                                                              int32_t* sort_idx, int32_t* acc_cnt,
                                                              const int32_t* __restrict__ heavy_q,
                                                              const int32_t* __restrict__ q_off, Params pr, long long cap,
                                                              fm_match* out, int32_t* out_count, Counters* ctr, int smem_cap,
                                                              WireOut wo) {
  extern __shared__ unsigned long long s_keys[];
  __shared__ float s_heap[64];
  __shared__ int s_nacc;
  __shared__ int s_tot[8];
  __shared__ int s_flag;
  if (ctr->overflow) return;
  const int n_heavy = (int)ctr->n_heavy;
  for (int h = blockIdx.x; h < n_heavy; h += gridDim.x) {
    const int q = heavy_q[h];
    const int n = q_cnt[q];
    const int p = q_off[q + 1] - q_off[q];
    const int base = q_base[q];
    fm_record* seg = rec + base;
    unsigned long long* gkeys = sort_key + base;
    int32_t* idx = sort_idx + base;
    const unsigned long long* packed = nullptr;  // the list in replay order when it was sorted as packed keys
    if (n < (1 << 20)) {
      unsigned long long* keys = n <= smem_cap ? s_keys : gkeys;
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const fm_record r = seg[i];
        keys[i] = ((unsigned long long)(unsigned)(1023 - r.longest_match) << 52) | ((unsigned long long)r.s_id << 20) | (unsigned)i;
      }
      __syncthreads();
      if (n <= smem_cap) block_sort_keys(keys, n);
      else keys = block_radix_sort_keys(gkeys, sort_key2 + base, n, 20, 62, reinterpret_cast<int(*)[256]>(s_keys), s_tot, &s_flag);
      // in place: sorted key -> (bound-test value, record index); every thread fetches the records of its slots
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int id = (int)(keys[i] & 0xfffffu);
        keys[i] = pack_candidate(seg[id], id, p, pr);
      }
      packed = keys;
    } else {
      for (int i = threadIdx.x; i < n; i += blockDim.x) { gkeys[i] = order_key(seg[i]); idx[i] = i; }
      __syncthreads();
      block_sort_pairs(gkeys, idx, n);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      float* heap = (pr.buffer > 0 && pr.buffer <= 62) ? s_heap : heapbuf + base + q;
      int nacc;
      if (packed) nacc = wo.cnt ? replay_sequence_packed(seg, n, packed, idx, gkeys, heap, pr, wo.stage + base, n, wo.cnt + q)
                                : replay_sequence_packed(seg, n, packed, idx, gkeys, heap, pr);
      else nacc = wo.cnt ? replay_sequence(seg, n, p, idx, gkeys, heap, pr, wo.stage + base, n, wo.cnt + q)
                         : replay_sequence(seg, n, p, idx, gkeys, heap, pr);
      if (threadIdx.x == 0) s_nacc = nacc;
    }
    __syncthreads();
    if (wo.cnt) continue;  // shard mode (uniform)
    const int nacc = s_nacc;
    block_sort_pairs(gkeys, idx, nacc);
    if (pr.contrast > 0.f) {
      if (threadIdx.x == 0) acc_cnt[q] = nacc;
    } else {
      const int want = pr.n_matches == 0 ? nacc : min(nacc, pr.n_matches);
      for (int k = threadIdx.x; k < want && k < cap; k += blockDim.x) out[(long long)q * cap + k] = to_match(seg[idx[k]], 0.f);
      if (threadIdx.x == 0) {
        out_count[q] = want;
        if (want) atomicAdd(&ctr->n_matches, (unsigned)want);
      }
    }
    __syncthreads();
  }
}

// One warp per query: contrastive rerank of src/fuzzy_match.cc:613-669. Penalties are edit distances between TM
// sentences (plain variant, unit costs) against every entry of `matches`, accumulated in vector order (running float
// sum for MEAN, running max for MAX): first the entries that were in the vector before the call (prior, optional:
// (sentence start, length) per entry, CSR by prior_off), then the matches this call selects.
__global__ void __launch_bounds__(256) fm_contrast_kernel(IndexDev ix, fm_record* rec, const int32_t* __restrict__ q_base,
                                                          const int32_t* __restrict__ sort_idx,
                                                          const int32_t* __restrict__ acc_cnt, int n_q, Params pr,
                                                          long long cap, fm_match* out, int32_t* out_count, Counters* ctr,
                                                          int stride, const int2* __restrict__ prior,
                                                          const int32_t* __restrict__ prior_off) {
  extern __shared__ int smem[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  int32_t* s_sent = smem + wib * 4 * stride;
  int32_t* s_pat = s_sent + stride;
  float* s_pen = reinterpret_cast<float*>(s_pat + stride);
  float* s_up = s_pen + stride;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  if (ctr->overflow) return;
  Params unit = pr;
  unit.ins = unit.del = unit.rep = 1.f;
  for (int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < n_q; q += n_warps) {
    const int n = acc_cnt[q];
    fm_record* seg = rec + q_base[q];
    const int32_t* idx = sort_idx + q_base[q];  // accepted records, best first
    const int p0 = prior ? prior_off[q] : 0, n_prior = prior ? prior_off[q + 1] - p0 : 0;
    int n_out = 0, remaining = n;
    int n_pen = 0;  // penalties accumulated so far for every remaining candidate = entries of `matches` accounted for
    int last = -1, j_prior = 0;
    while (remaining > 0 && (pr.n_matches == 0 || n_out + n_prior < pr.n_matches)) {
      while (last >= 0 || j_prior < n_prior) {
        int o_start, o_len;
        if (last >= 0) {
          const fm_record lr = seg[idx[last]];
          o_start = lr.reserved[0]; o_len = lr.length;
          last = -1;
        } else {
          const int2 e = prior[p0 + j_prior++];
          o_start = e.x; o_len = e.y;
        }
        for (int k = lane; k < o_len; k += 32) s_pat[k] = ix.tok[o_start + k];
        for (int k = lane; k < o_len; k += 32) s_pen[k] = 0.f;
        __syncwarp();
        for (int i = 0; i < n; i++) {
          const fm_record cr = seg[idx[i]];
          if (cr.reserved[2]) continue;  // already selected
          for (int k = lane; k < cr.length; k += 32) s_sent[k] = ix.tok[cr.reserved[0] + k];
          __syncwarp();
          const float wdiff = __fdiv_rn(100.f, normalizer(cr.length, o_len, unit));
          float C, K;
          warp_edit_distance<false>(s_sent, cr.length, s_pat, o_len, s_pen, s_up, wdiff, wdiff, wdiff, C, K, RealSide{});
          if (lane == 0) {
            const float pen = score_of(C);
            float acc = __int_as_float(cr.reserved[1]);
            if (pr.reduce == 1) acc = (n_pen == 0 || pen > acc) ? pen : acc;
            else acc = __fadd_rn(acc, pen);
            seg[idx[i]].reserved[1] = __float_as_int(acc);
          }
        }
        n_pen++;
        __syncwarp();
      }
      int best = -1;
      if (lane == 0) {  // std::max_element: first maximum of score - factor*penalty in list order
        float best_key = 0.f;
        for (int i = 0; i < n; i++) {
          const fm_record cr = seg[idx[i]];
          if (cr.reserved[2]) continue;
          const float acc = __int_as_float(cr.reserved[1]);
          const float pen = n_pen == 0 ? 0.f : (pr.reduce == 1 ? acc : __fdiv_rn(acc, (float)n_pen));
          const float key = __fsub_rn(cr.rowmin_max, __fmul_rn(pr.contrast, pen));
          if (best < 0 || best_key < key) { best = i; best_key = key; }
        }
        const fm_record br = seg[idx[best]];
        const float acc = __int_as_float(br.reserved[1]);
        const float pen = n_pen == 0 ? 0.f : (pr.reduce == 1 ? acc : __fdiv_rn(acc, (float)n_pen));
        if (n_out < cap) out[(long long)q * cap + n_out] = to_match(br, pen);
        seg[idx[best]].reserved[2] = 1;
      }
      best = __shfl_sync(FULL, best, 0);
      last = best;
      n_out++; remaining--;
      __syncwarp();
    }
    if (lane == 0) {
      out_count[q] = n_out;
      if (n_out) atomicAdd(&ctr->n_matches, (unsigned)n_out);
    }
  }
}

// Contrastive rerank on a sharded TM (the accepted records of all shards are on every rank, the sentences are not):
// every rank lays out one token slab in the order of the merged accepted lists -- tok_cnt[q] = tokens of the accepted
// sentences of query q, scanned into tok_base -- fills the sentences it owns and leaves zeros elsewhere; the sum over
// the ranks (one all-reduce) is the slab the rerank kernel reads instead of the index's token array.
__global__ void fm_contrast_need_kernel(const fm_record* __restrict__ rec, const int32_t* __restrict__ q_base,
                                        const int32_t* __restrict__ sort_idx, const int32_t* __restrict__ acc_cnt, int n_q,
                                        int32_t* tok_cnt) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_q) return;
  const fm_record* seg = rec + q_base[q];
  const int32_t* idx = sort_idx + q_base[q];
  int c = 0;
  for (int i = 0; i < acc_cnt[q]; i++) c += seg[idx[i]].length;
  tok_cnt[q] = c;
}
__global__ void __launch_bounds__(256) fm_contrast_fill_kernel(IndexDev ix, long long n_sent_local, fm_record* rec,
                                                               const int32_t* __restrict__ q_base, const int32_t* __restrict__ sort_idx,
                                                               const int32_t* __restrict__ acc_cnt, const int32_t* __restrict__ tok_base,
                                                               int n_q, int32_t* slab) {
  const int lane = threadIdx.x & 31;
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= n_q) return;
  fm_record* seg = rec + q_base[q];
  const int32_t* idx = sort_idx + q_base[q];
  int off = tok_base[q];
  for (int i = 0; i < acc_cnt[q]; i++) {
    const fm_record r = seg[idx[i]];
    const long long local = (long long)r.s_id - (long long)ix.sid_base;
    if (local >= 0 && local < n_sent_local) {
      const int start = ix.sent_start[local];
      for (int k = lane; k < r.length; k += 32) slab[off + k] = ix.tok[start + k];
    }
    if (lane == 0) seg[idx[i]].reserved[0] = off;  // where the rerank finds the sentence
    off += r.length;
  }
}

// ---------------------------------------------------------------- subsequence()
//
// FuzzyMatch::subsequence (reference src/fuzzy_match.cc:238-365) behind its tokenizer, one warp per pattern:
// the sub-sequences of the pattern are tried by weight (length, or summed IDF) until one occurs in a sentence
// that is not skipped as perfect; the suffix-array range of that sub-sequence is walked in order for at most
// number_of_matches distinct sentences, each scored with the full edit distance (unit costs), and the best
// under the reference's integer-truncated running bound is returned. Nothing here is throughput critical
// (a handful of dependent lookups and one or two edit distances per pattern); the batch gives the parallelism.

// suffixes that start with the bigram / the trigram (the directories of the search kernel)
__device__ __forceinline__ bool dir_bigram(const IndexDev& ix, int t0, int t1, int& lo, int& hi, uint32_t& slot) {
  if (t0 < 2 || t1 < 2) return false;
  uint32_t h = bigram_hash(t0, t1) & ix.bg_mask;
  for (;;) {
    const int4 e = __ldg(ix.bg_tab + h);
    if (e.x == t0 && e.y == t1) { lo = e.z; hi = e.w; slot = h; return true; }
    if (e.x == -1) return false;
    h = (h + 1) & ix.bg_mask;
  }
}
__device__ __forceinline__ bool dir_trigram(const IndexDev& ix, int t0, int t1, int t2, int& lo, int& hi) {
  if (t2 < 2) return false;
  int pos1;
  uint32_t slot;
  return tg_lookup(ix, t0, t1, t2, lo, hi, pos1, slot);
}
// equal range of word t at depth `depth` inside [lo, hi) (all suffixes there share `depth` words)
__device__ __forceinline__ void narrow_range(const IndexDev& ix, int depth, int t, int& lo, int& hi) {
  if (t < 2) { hi = lo; return; }
  auto key = [&](int k) { return depth == 3 ? __ldg(ix.sa_next + k) : __ldg(ix.tok + (__ldg(ix.sa_pos + k) + depth)); };
  int a = lo, e = hi;
  while (a < e) {  // first suffix whose word is not < t
    const int m = (int)(((unsigned)a + (unsigned)e) >> 1);
    if (key(m) < t) a = m + 1; else e = m;
  }
  const int first = a;
  e = hi;
  while (a < e) {  // first suffix whose word is > t
    const int m = (int)(((unsigned)a + (unsigned)e) >> 1);
    if (key(m) <= t) a = m + 1; else e = m;
  }
  lo = first;
  hi = a;
}
// SuffixArray::equal_range of pat[0..len) (src/suffix_array.cc:105-212); returns false when no suffix starts with it
__device__ bool ngram_range(const IndexDev& ix, const int32_t* pat, int len, int& lo, int& hi) {
  const int t0 = pat[0];
  if (t0 < 2 || t0 >= ix.vocab_size) return false;
  if (len == 1) { lo = __ldg(ix.qva + t0); hi = __ldg(ix.qva + t0 + 1); return hi > lo; }
  uint32_t slot = 0;
  if (!dir_bigram(ix, t0, pat[1], lo, hi, slot)) return false;
  if (len >= 3 && !dir_trigram(ix, t0, pat[1], pat[2], lo, hi)) return false;
  for (int d = 3; d < len; d++) {
    narrow_range(ix, d, pat[d], lo, hi);
    if (hi <= lo) return false;
  }
  return true;
}
// longest prefix of pat[0..n) that occurs in the TM
__device__ int longest_prefix(const IndexDev& ix, const int32_t* pat, int n) {
  int lo = 0, hi = 0;
  if (n < 1 || !ngram_range(ix, pat, 1, lo, hi)) return 0;
  if (n < 2) return 1;
  uint32_t slot = 0;
  if (!dir_bigram(ix, pat[0], pat[1], lo, hi, slot)) return 1;
  if (n < 3 || !dir_trigram(ix, pat[0], pat[1], pat[2], lo, hi)) return 2;
  int len = 3;
  while (len < n) {
    narrow_range(ix, len, pat[len], lo, hi);
    if (hi <= lo) break;
    len++;
  }
  return len;
}

struct SubseqKey {  // priority of a sub-sequence: weight desc, position asc (Subseq::operator<, :238-248), then length desc
  float w;
  int pos, len;
};
__device__ __forceinline__ bool subseq_before(const SubseqKey& a, const SubseqKey& b) {  // a is tried before b
  if (a.w != b.w) return a.w > b.w;
  if (a.pos != b.pos) return a.pos < b.pos;
  return a.len > b.len;
}

__global__ void __launch_bounds__(128) fm_subseq_kernel(IndexDev ix, const int32_t* __restrict__ q_tok, const int32_t* __restrict__ q_off,
                                                        int n_q, int n_matches, int no_perfect, int ml_in, float mr, int idf_weighting,
                                                        int stride, uint32_t* seen_buf, int seen_cap, fm_subseq* out) {
  extern __shared__ int smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int32_t* s_sent = smem + wib * 5 * stride;
  int32_t* s_pat = s_sent + stride;
  float* s_pen = reinterpret_cast<float*>(s_pat + stride);
  float* s_up = s_pen + stride;
  int32_t* s_lmax = reinterpret_cast<int32_t*>(s_up + stride);
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= n_q) return;
  fm_subseq res;
  res.s_id = 0; res.score = 0.f; res.cost = 0.f; res.position = 0; res.length = 0; res.found = 0;
  const int off = q_off[q], p = q_off[q + 1] - off;
  int ml = ml_in;
  if ((int)__fmul_rn(mr, (float)p) > ml) ml = (int)__fmul_rn(mr, (float)p);  // :263-264
  if (p < ml || p < 1 || p > ix.max_tokens) {  // :266-267 (patterns beyond max_tokens_in_pattern are not searched)
    if (lane == 0) out[q] = res;
    return;
  }
  // words the TM does not know are VOCAB_UNK: no sub-sequence runs over them (:281-283)
  for (int j = lane; j < p; j += 32) {
    const int t = q_tok[off + j];
    const bool known = t >= 2 && t < ix.vocab_size && __ldg(ix.qva + t + 1) > __ldg(ix.qva + t);
    s_pat[j] = known ? t : 1;
    s_pen[j] = 0.f;  // the edit distance runs with idf_weight 0 (:321-326)
  }
  __syncwarp();
  for (int it = lane; it < p; it += 32) s_lmax[it] = longest_prefix(ix, s_pat + it, p - it);
  __syncwarp();
  uint32_t* seen = seen_buf + (size_t)q * seen_cap;  // candidates first, then the perfect sentences
  int n_cand = 0, n_perfect = 0;
  int max_distance = 10000;
  SubseqKey prev;
  prev.w = __int_as_float(0x7f800000); prev.pos = -1; prev.len = 0;  // before everything
  while (max_distance == 10000) {
    // the next sub-sequence in priority order among those that occur in the TM (the others have empty ranges)
    SubseqKey best;
    best.w = -1.f; best.pos = 0; best.len = 0;
    for (int it = lane; it < p; it += 32) {
      float w = 0.f;
      const int lm = s_lmax[it];
      for (int len = 1; len <= lm; len++) {
        w = idf_weighting ? __fadd_rn(w, __ldg(ix.idf + s_pat[it + len - 1])) : __fadd_rn(w, 1.f);
        if (len < ml) continue;
        SubseqKey k;
        k.w = w; k.pos = it; k.len = len;
        if (subseq_before(prev, k) && (best.len == 0 || subseq_before(k, best))) best = k;
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      SubseqKey o;
      o.w = __shfl_xor_sync(FULL, best.w, d); o.pos = __shfl_xor_sync(FULL, best.pos, d); o.len = __shfl_xor_sync(FULL, best.len, d);
      if (o.len != 0 && (best.len == 0 || subseq_before(o, best))) best = o;
    }
    if (best.len == 0) break;  // nothing left
    prev = best;
    int lo = 0, hi = 0;
    if (lane == 0) ngram_range(ix, s_pat + best.pos, best.len, lo, hi);
    lo = __shfl_sync(FULL, lo, 0);
    hi = __shfl_sync(FULL, hi, 0);
    for (int su = lo; su < hi && n_cand < n_matches; su++) {  // :308-309
      const int start = __ldg(reinterpret_cast<const int*>(ix.sa_aux + 2ll * su));
      const uint32_t sid = (uint32_t)__ldg(ix.sid_at + (start >> 2));
      bool dup = false;
      for (int i = lane; i < n_cand + n_perfect; i += 32) dup |= seen[i < n_cand ? i : n_matches + (i - n_cand)] == sid;
      if (__any_sync(FULL, dup)) continue;
      int slen = 0;
      for (int k0 = 0;; k0 += 32) {  // stage the sentence (it ends at the separator)
        const int t = __ldg(ix.tok + min(start + k0 + lane, ix.n_buf - 1));  // (the buffer ends with a separator)
        const unsigned z = __ballot_sync(FULL, t == 0);
        s_sent[k0 + lane] = t;
        if (z) { slen = k0 + __ffs(z) - 1; break; }
      }
      __syncwarp();
      Params unit;
      unit.ins = unit.del = unit.rep = 1.f;
      const float wdiff = __fdiv_rn(100.f, normalizer(p, slen, unit));  // Costs(p, s, EditCosts()) :317-318
      float C, K;
      warp_edit_distance<false>(s_sent, slen, s_pat, p, s_pen, s_up, __fmul_rn(1.f, wdiff), __fmul_rn(1.f, wdiff), __fmul_rn(1.f, wdiff), C,
                                K, RealSide{});
      if (C == 0.f && no_perfect) {  // :327-330
        if (n_matches + n_perfect >= seen_cap) { res.found = -1; max_distance = -1; break; }
        if (lane == 0) seen[n_matches + n_perfect] = sid;
        n_perfect++;
        __syncwarp();
        continue;
      }
      if (C < (float)max_distance) {  // :331-350
        res.found = 1;
        res.score = score_of(C);
        res.cost = C;
        res.length = best.len;
        res.position = best.pos;
        res.s_id = sid + ix.sid_base;
        max_distance = (int)C;  // int max_distance = cost
        if (C == 0.f) break;
      }
      if (lane == 0) seen[n_cand] = sid;
      n_cand++;
      __syncwarp();
    }
  }
  if (lane == 0) out[q] = res;
}

// ---------------------------------------------------------------- cross-shard merge helpers
//
// One shard's block for n_q queries with room for `capacity` records (what travels in the all-gather):
//   int32 header[4]     = (overflow flags of the shard's own pipeline, n_q, capacity, accepted records in total)
//   int32 off[n_q + 1]  exclusive offsets of the queries' records (padded to a multiple of 4 words)
//   fm_wire rec[capacity]   the accepted records, query after query, each in the order its loop accepted them
struct WireBlocks {
  const int32_t* blk[16];
};
__host__ __device__ inline long long wire_off_words(long long n_q) { return (n_q + 1 + 3) / 4 * 4; }
// Compacts the staged records into the block (one thread per query) and writes the header.
__global__ void fm_wire_pack_kernel(int32_t* block, const Counters* ctr, const fm_wire* __restrict__ stage,
                                    const int32_t* __restrict__ q_base, int n_q, int capacity) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int32_t* off = block + 4;
  if (q == 0) {
    block[0] = (int32_t)ctr->overflow;
    block[1] = n_q;
    block[2] = capacity;
    block[3] = off[n_q];
  }
  if (q >= n_q || ctr->overflow) return;
  fm_wire* rec = reinterpret_cast<fm_wire*>(block + 4 + wire_off_words(n_q));
  const int o0 = off[q], o1 = off[q + 1];
  const fm_wire* src = stage + q_base[q];
  for (int i = o0; i < o1 && i < capacity; i++) rec[i] = src[i - o0];
}
// m_cnt[q] = records of query q over all shards. mctr->wire_need = the largest total of one shard (the callers
// size the next blocks from it); mctr->overflow: 0x100 a shard's own pipeline overflowed, 0x200 blocks of another
// batch shape, 0x400 a shard accepted more records than its block holds.
__global__ void fm_wire_count_kernel(WireBlocks wb, int n_shards, int32_t* m_cnt, int n_q, Counters* mctr) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q == 0) {
    unsigned bad = 0, mx = 0;
    for (int k = 0; k < n_shards; k++) {
      const int32_t* h = wb.blk[k];
      if (h[0]) bad |= 0x100u;
      if (h[1] != n_q) bad |= 0x200u;
      if (h[3] > h[2]) bad |= 0x400u;
      mx = max(mx, (unsigned)h[3]);
    }
    if (bad) atomicOr(&mctr->overflow, bad);
    mctr->wire_need = mx;
  }
  if (q > n_q) return;
  int c = 0;
  if (q < n_q)
    for (int k = 0; k < n_shards; k++) c += wb.blk[k][4 + q + 1] - wb.blk[k][4 + q];
  m_cnt[q] = c;
}
__global__ void fm_wire_copy_kernel(WireBlocks wb, int n_shards, const int32_t* __restrict__ m_base, fm_record* mrec, int n_q,
                                    const Counters* mctr) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_q || mctr->overflow) return;
  int dst = m_base[q];
  const long long ow = wire_off_words(n_q);
  for (int k = 0; k < n_shards; k++) {
    const int o0 = wb.blk[k][4 + q], o1 = wb.blk[k][4 + q + 1];
    const fm_wire* src = reinterpret_cast<const fm_wire*>(wb.blk[k] + 4 + ow);
    for (int i = o0; i < o1; i++) {
      const fm_wire w = src[i];
      fm_record r;
      r.s_id = w.s_id;
      r.longest_match = (int32_t)(w.lm_len & 0xffffu);
      r.length = (int32_t)(w.lm_len >> 16);
      r.cost = w.cost;
      r.rowmin_max = w.rowmin_max;
      r.reserved[0] = -1;
      r.reserved[1] = 0;
      r.reserved[2] = 0;
      mrec[dst++] = r;
    }
  }
}

// ---------------------------------------------------------------- launchers

static int dp_stride(const IndexDev& ix) { return ((ix.max_tokens + 31) / 32) * 32 + 32; }

void launch_bounds(const IndexDev& ix, const BatchDev& b, const Params& p, cudaStream_t st) {
  fm_bounds_kernel<<<(ix.max_tokens + 8) / 8, 256, 0, st>>>(ix.max_tokens, p, const_cast<uint16_t*>(b.cmin_tab),
                                                            const_cast<uint16_t*>(b.cmin64));
}
int launch_prepare(const IndexDev& ix, const BatchDev& b, const Params& p, int sm_count, cudaStream_t st) {  // returns the launches
  const int warps_per_block = 8;
  const int grid = (b.n_q + warps_per_block - 1) / warps_per_block;
  static const bool warp_only = getenv("FM_PREPARE_WARP_ONLY") != nullptr;  // (tests: the warp kernel for every query)
  if (b.wq || warp_only) {  // wide signatures need the warp kernel's planes for every query
    fm_prepare_kernel<<<grid, 256, 0, st>>>(ix, b, p);
    return 1;
  }
  fm_prepare_short_kernel<<<(b.n_q + kPrepThreads - 1) / kPrepThreads, kPrepThreads, 0, st>>>(ix, b, p);
  fm_prepare_list_kernel<<<std::min(grid, sm_count), 256, 0, st>>>(ix, b, p);
  return 2;
}
void launch_search(const IndexDev& ix, const BatchDev& b, const Params& p, cudaStream_t st) {
  const int grid = (b.n_tok + FM_SEARCH_THREADS - 1) / FM_SEARCH_THREADS;
  if (grid > 0) fm_search_kernel<<<grid, FM_SEARCH_THREADS, 0, st>>>(ix, b, p);
}
void launch_gather(const IndexDev& ix, const BatchDev& b, const Params& p, int sm_count, cudaStream_t st, cudaEvent_t between) {
  // grids several times what is resident: CTAs that finish early make room for the next ones
  fm_gather_kernel<<<sm_count * FM_GATHER_CTAS * 8, 256, 0, st>>>(ix, b);
  if (between) cudaEventRecord(between, st);
  fm_verify_kernel<<<sm_count * FM_VERIFY_CTAS * 4, 256, 0, st>>>(ix, b, p);
}
void launch_scan(const int32_t* in, int32_t* out, int32_t n, unsigned long long* chain, unsigned int epoch, int sm_count,
                 cudaStream_t st) {
  int ctas = (n + 4095) / 4096;
  ctas = ctas < 1 ? 1 : (ctas > kScanCtas ? kScanCtas : ctas);
  if (ctas > sm_count) ctas = sm_count;  // all CTAs must be co-resident
  fm_scan_kernel<<<ctas, 1024, 0, st>>>(in, out, n, chain, epoch);
}
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: opt in once per (kernel, device).
// One process may hold indexes on several GPUs and call from several host threads.
struct SmemOptIn {
  std::atomic<unsigned long long> done{0};
  template <class F>
  void ensure(F fn, int bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return;
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    done.fetch_or(bit, std::memory_order_release);
  }
};

void launch_score(const IndexDev& ix, const BatchDev& b, const Params& p, int sm_count, cudaStream_t st) {
  const int stride = dp_stride(ix);
  static SmemOptIn opt_f, opt_t, opt_bpw;
  opt_f.ensure(fm_score_kernel<false>, 200 * 1024);
  opt_t.ensure(fm_score_kernel<true>, 200 * 1024);
  if (b.q_real) {  // Sentence API: every pair through the wavefront kernel with the real-token terms
    const int warps = 4;  // 8 staging arrays per warp
    fm_score_kernel<true><<<sm_count * 4, warps * 32, (size_t)warps * 8 * stride * sizeof(int), st>>>(ix, b, p, stride, 0, 0u);
    return;
  }
  const size_t smem = (size_t)8 * 4 * stride * sizeof(int);
  // Environment switches (read once per process; tests/test_gpu_scale.py runs each in a fresh process):
  // FM_SCORE_WARP_ONLY=1 sends every pair through the float warp wavefront, FM_SCORE_FLOAT_ONLY=1 disables
  // the bit-parallel kernels so that equal costs take the float kernels too.
  static const bool warp_only = getenv("FM_SCORE_WARP_ONLY") != nullptr;
  static const bool float_only = getenv("FM_SCORE_FLOAT_ONLY") != nullptr;
  const int grid_t = (int)((b.surv_cap + 127) / 128);
  if (warp_only) {
    fm_score_kernel<false><<<sm_count * 4, 256, smem, st>>>(ix, b, p, stride, 0, 0u);
    return;
  }
  const bool equal_costs = p.ins == p.del && p.del == p.rep && p.ins > 0.f && p.idf_penalty == 0.f;
  if (equal_costs && !float_only) {
    // thread per pair (p <= 64) -> warp per pair (p <= 320) -> float wavefront (longer patterns)
    const int mp = std::min(ix.max_tokens, kBpWarpMax);
    const int peq_words = mp * ((mp + 31) / 32), didx_words = (ix.max_tokens + 2) / 2 + 1;
    const size_t bsm = (size_t)4 * (peq_words + didx_words) * sizeof(uint32_t);
    opt_bpw.ensure(fm_score_bpw_kernel, 200 * 1024);
    fm_score_bp_kernel<<<grid_t, 128, 0, st>>>(ix, b, p);
    if (ix.max_tokens > kBpThreadMax) fm_score_bpw_kernel<<<sm_count * 3, 128, bsm, st>>>(ix, b, p, peq_words, didx_words);
    if (ix.max_tokens > kBpWarpMax) fm_score_kernel<false><<<sm_count * 4, 256, smem, st>>>(ix, b, p, stride, kBpWarpMax + 1, 2u);
    return;
  }
  if (p.idf_penalty != 0.f) fm_score_short_kernel<true><<<grid_t, 128, 0, st>>>(ix, b, p);
  else fm_score_short_kernel<false><<<grid_t, 128, 0, st>>>(ix, b, p);
  if (ix.max_tokens > 32) fm_score_kernel<false><<<sm_count * 4, 256, smem, st>>>(ix, b, p, stride, 33, 1u);
}
void launch_replay(const IndexDev& ix, const fm_record* rec, const int32_t* q_cnt, const int32_t* q_base, float* heapbuf,
                   unsigned long long* sort_key, unsigned long long* sort_key2, int32_t* sort_idx, int32_t* acc_cnt,
                   int32_t* mid_q, int32_t* heavy_q, const int32_t* q_off, int32_t n_q, const Params& p, int64_t cap,
                   fm_match* out, int32_t* out_count, Counters* ctr, int sm_count, cudaStream_t st, cudaStream_t st2,
                   cudaEvent_t ev_fork, cudaEvent_t ev_join, int32_t* wire_cnt, fm_wire* wire_stage) {
  const WireOut wo{wire_cnt, wire_stage};
  // FM_WARP_MAX=<n> (32..kWarpMax) moves the warp / CTA boundary (tests/test_gpu_scale.py: fresh process per setting)
  static const int warp_max = getenv("FM_WARP_MAX") ? std::max(32, std::min(kWarpMax, atoi(getenv("FM_WARP_MAX")))) : kWarpMax;
  fm_replay_small_kernel<<<(n_q + 255) / 256, 256, 0, st>>>(const_cast<fm_record*>(rec), q_cnt, q_base, sort_idx, acc_cnt, mid_q,
                                                            heavy_q, q_off, n_q, p, (long long)cap, out, out_count, ctr, warp_max, wo);
  const size_t smem = (size_t)kHeavySmem * sizeof(unsigned long long);
  // FM_HEAVY_SMEM=<n> lowers the shared-memory sort limit: the CTA radix sort then takes shorter lists too
  // (tests/test_gpu_scale.py; without it lists of more than kHeavySmem candidates take that path)
  static const int smem_cap = getenv("FM_HEAVY_SMEM") ? std::max(64, std::min(kHeavySmem, atoi(getenv("FM_HEAVY_SMEM")))) : kHeavySmem;
  static SmemOptIn opt_heavy;
  opt_heavy.ensure(fm_replay_heavy_kernel, (int)smem);
  // The two remaining kernels work on disjoint queries (the lists the small kernel wrote). The heavy
  // one is a few long sequential replays, so it runs on a side stream next to the warp-per-query one.
  cudaStream_t sh = st2 ? st2 : st;
  if (st2) {
    cudaEventRecord(ev_fork, st);
    cudaStreamWaitEvent(st2, ev_fork, 0);
  }
  fm_replay_heavy_kernel<<<sm_count * 6, 256, smem, sh>>>(const_cast<fm_record*>(rec), q_cnt, q_base, heapbuf, sort_key, sort_key2,
                                                          sort_idx, acc_cnt, heavy_q, q_off, p, (long long)cap, out, out_count, ctr, smem_cap, wo);
  if (st2) cudaEventRecord(ev_join, st2);
  int grid = (n_q + 7) / 8;
  if (grid > sm_count * 8) grid = sm_count * 8;
  fm_replay_kernel<<<grid, 256, 0, st>>>(const_cast<fm_record*>(rec), q_cnt, q_base, heapbuf, sort_key, sort_idx, acc_cnt, mid_q,
                                         q_off, p, (long long)cap, out, out_count, ctr, wo);
  if (st2) cudaStreamWaitEvent(st, ev_join, 0);
}
void launch_contrast(const IndexDev& ix, fm_record* rec, const int32_t* q_base, const int32_t* sort_idx,
                     const int32_t* acc_cnt, int32_t n_q, const Params& p, int64_t cap, fm_match* out, int32_t* out_count, Counters* ctr, int sm_count,
                     cudaStream_t st, const int2* prior, const int32_t* prior_off) {
  const int stride = dp_stride(ix);
  const size_t smem = (size_t)8 * 4 * stride * sizeof(int);
  static SmemOptIn opt_contrast;
  opt_contrast.ensure(fm_contrast_kernel, 200 * 1024);
  int grid = (n_q + 7) / 8;
  if (grid > sm_count * 4) grid = sm_count * 4;
  fm_contrast_kernel<<<grid, 256, smem, st>>>(ix, rec, q_base, sort_idx, acc_cnt, n_q, p, (long long)cap, out, out_count, ctr, stride, prior, prior_off);
}

void launch_contrast_need(const fm_record* rec, const int32_t* q_base, const int32_t* sort_idx, const int32_t* acc_cnt, int32_t n_q,
                          int32_t* tok_cnt, cudaStream_t st) {
  fm_contrast_need_kernel<<<(n_q + 255) / 256, 256, 0, st>>>(rec, q_base, sort_idx, acc_cnt, n_q, tok_cnt);
}
void launch_contrast_fill(const IndexDev& ix, int64_t n_sent_local, fm_record* rec, const int32_t* q_base, const int32_t* sort_idx,
                          const int32_t* acc_cnt, const int32_t* tok_base, int32_t n_q, int32_t* slab, cudaStream_t st) {
  fm_contrast_fill_kernel<<<(n_q + 7) / 8, 256, 0, st>>>(ix, (long long)n_sent_local, rec, q_base, sort_idx, acc_cnt, tok_base, n_q, slab);
}
void launch_subseq(const IndexDev& ix, const int32_t* q_tok, const int32_t* q_off, int32_t n_q, int n_matches, int no_perfect, int ml, float mr,
                   int idf_weighting, uint32_t* seen, int seen_cap, fm_subseq* out, cudaStream_t st) {
  const int stride = dp_stride(ix);
  const size_t smem = (size_t)4 * 5 * stride * sizeof(int);
  static SmemOptIn opt;
  opt.ensure(fm_subseq_kernel, 200 * 1024);
  fm_subseq_kernel<<<(n_q + 3) / 4, 128, smem, st>>>(ix, q_tok, q_off, n_q, n_matches, no_perfect, ml, mr, idf_weighting, stride, seen, seen_cap, out);
}
void launch_wire_pack(int32_t* block, const Counters* ctr, const fm_wire* stage, const int32_t* q_base, int32_t n_q, int capacity,
                      cudaStream_t st) {
  fm_wire_pack_kernel<<<(n_q + 255) / 256, 256, 0, st>>>(block, ctr, stage, q_base, n_q, capacity);
}
void launch_wire_count(int n_shards, const int32_t* const* blocks, int32_t* m_cnt, int32_t n_q, Counters* mctr, cudaStream_t st) {
  WireBlocks wb{};
  for (int k = 0; k < n_shards; k++) wb.blk[k] = blocks[k];
  fm_wire_count_kernel<<<(n_q + 256) / 256, 256, 0, st>>>(wb, n_shards, m_cnt, n_q, mctr);
}
void launch_wire_copy(int n_shards, const int32_t* const* blocks, const int32_t* m_base, fm_record* mrec, int32_t n_q,
                      const Counters* mctr, cudaStream_t st) {
  WireBlocks wb{};
  for (int k = 0; k < n_shards; k++) wb.blk[k] = blocks[k];
  fm_wire_copy_kernel<<<(n_q + 255) / 256, 256, 0, st>>>(wb, n_shards, m_base, mrec, n_q, mctr);
}

}  // namespace fm
