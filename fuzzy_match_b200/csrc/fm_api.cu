// fm_api.cu -- the C ABI (include/fuzzy_match_b200.h): workspace management and batch orchestration.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "fm_internal.h"

namespace fm {
const std::string& get_error();

template <class T>
static int dev_realloc(T** p, size_t n) {
  if (*p) cudaFree(*p);
  *p = nullptr;
  void* d = nullptr;
  FM_CUDA(cudaMalloc(&d, (n ? n : 1) * sizeof(T)));
  *p = static_cast<T*>(d);
  return FM_OK;
}

static const int64_t kSpanHost = kSpan;
static int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

static Workspace* acquire(Index* ix) {
  std::lock_guard<std::mutex> g(ix->mu);
  for (Workspace* w : ix->pool)
    if (!w->in_use) { w->in_use = true; return w; }
  Workspace* w = new Workspace();
  w->device = ix->device;
  w->in_use = true;
  ix->pool.push_back(w);
  return w;
}
static void release(Index* ix, Workspace* w) {
  std::lock_guard<std::mutex> g(ix->mu);
  w->in_use = false;
}

static int ensure_base(Workspace* w) {
  if (!w->stream) {
    FM_CUDA(cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking));
    for (auto& e : w->ev) FM_CUDA(cudaEventCreate(&e));
    FM_CUDA(cudaStreamCreateWithFlags(&w->stream2, cudaStreamNonBlocking));
    FM_CUDA(cudaEventCreateWithFlags(&w->ev_fork, cudaEventDisableTiming));
    FM_CUDA(cudaEventCreateWithFlags(&w->ev_join, cudaEventDisableTiming));
    FM_CUDA(cudaEventCreateWithFlags(&w->ev_done, cudaEventDisableTiming));
    FM_CUDA(cudaMalloc((void**)&w->ctr, sizeof(Counters)));
    FM_CUDA(cudaMalloc((void**)&w->scan_chain, 256 * sizeof(unsigned long long)));
    FM_CUDA(cudaMemset(w->scan_chain, 0, 256 * sizeof(unsigned long long)));
    FM_CUDA(cudaMallocHost((void**)&w->h_ctr, sizeof(Counters)));
    FM_CUDA(cudaMalloc((void**)&w->mctr, sizeof(Counters)));
    FM_CUDA(cudaMallocHost((void**)&w->h_mctr, sizeof(Counters)));
  }
  return FM_OK;
}

static int ensure_queries(Workspace* w, int64_t n_q, int64_t n_tok, bool staging) {
  int rc;
  if (n_q > w->cap_q) {
    const int64_t c = round_up(n_q + n_q / 4 + 1024, 1024);
    if ((rc = dev_realloc(&w->qmeta, c)) || (rc = dev_realloc(&w->q_cnt, c + 1)) || (rc = dev_realloc(&w->q_base, c + 1)) ||
        (rc = dev_realloc(&w->acc_cnt, c)) || (rc = dev_realloc(&w->heavy_q, c)) || (rc = dev_realloc(&w->mid_q, c)) || (rc = dev_realloc(&w->qmask, c)) || (rc = dev_realloc(&w->qmask2, 3 * c)) || (rc = dev_realloc(&w->prep_list, c)) || (rc = dev_realloc(&w->d_q_off, c + 1)) || (rc = dev_realloc(&w->d_out_count, c + 1)))
      return rc;
    w->cap_q = c;
    w->cap_surv = 0;  // heapbuf depends on cap_q
  }
  if (n_tok > w->cap_tok) {
    const int64_t c = round_up(n_tok + n_tok / 4 + 4096, 4096);
    if ((rc = dev_realloc(&w->pat, c)) || (rc = dev_realloc(&w->chain_rec, c)) || (rc = dev_realloc(&w->tbl, 4 * c)) ||
        (rc = dev_realloc(&w->d_q_tok, c)) || (rc = dev_realloc(&w->peq64, c)))
      return rc;
    w->cap_tok = c;
  }
  if (staging && n_q + 1 > w->cap_hq) {
    if (w->h_q_off32) cudaFreeHost(w->h_q_off32);
    w->h_q_off32 = nullptr;
    const int64_t c = round_up(n_q + n_q / 4 + 1025, 1024);
    FM_CUDA(cudaMallocHost((void**)&w->h_q_off32, c * sizeof(int32_t)));
    w->cap_hq = c;
  }
  return FM_OK;
}
// per-query planes over the wide signature bits: only for an index that has wide signatures
static int ensure_wide(Index* ix, Workspace* w) {
  if (ix->dev.n_wide == 0 || w->cap_wq >= w->cap_q) return FM_OK;
  int rc;
  if ((rc = dev_realloc(&w->wq, (size_t)w->cap_q * kWideStride))) return rc;
  w->cap_wq = w->cap_q;
  return FM_OK;
}
// (Worklists grow with a quarter of slack: batches of one stream differ a little in size, and a
// reallocation -- cudaFree + cudaMalloc synchronise the device -- must not recur batch after batch.)
static int ensure_slices(Workspace* w, int64_t n) {
  if (n <= w->cap_slices) return FM_OK;
  n = std::min<int64_t>(n + n / 4, (int64_t(1) << 26) - 65537);
  // the packed (slices << 38 | elements) counter leaves 26 bits for the slice count
  if (n >= (int64_t(1) << 26) - 65536) { set_error("too many suffix-array range slices in one batch: split the batch"); return FM_ERR_NOMEM; }
  int rc;
  if ((rc = dev_realloc(&w->sl_start, n + 1)) || (rc = dev_realloc(&w->sl_rec, 2 * n)) || (rc = dev_realloc(&w->sm_rec, 2 * n))) return rc;
  w->cap_slices = n;
  return FM_OK;
}
static int ensure_spans(Workspace* w, int64_t n) {
  if (n <= w->cap_spans) return FM_OK;
  n += n / 4;
  int rc;
  if ((rc = dev_realloc(&w->span_slice, n))) return rc;
  w->cap_spans = n;
  return FM_OK;
}
static int ensure_bounds(Index* ix, Workspace* w) {
  if (w->cmin_tab) return FM_OK;
  int rc;
  const int64_t t = ix->max_tokens;
  if ((rc = dev_realloc(&w->cmin_tab, (t + 1) << 10)) || (rc = dev_realloc(&w->cmin64, (t + 1) << 6))) return rc;
  w->bounds_valid = false;
  return FM_OK;
}
// stage-1 survivors handed from the walk kernel to the verify kernel
static int ensure_cand(Workspace* w, int64_t n) {
  if (n <= w->cap_cand) return FM_OK;
  int rc;
  if (n > (int64_t(1) << 31)) { set_error("too many candidates in one batch: split the batch"); return FM_ERR_NOMEM; }
  n = std::min<int64_t>(n + n / 4, int64_t(1) << 31);
  if ((rc = dev_realloc(&w->cand, n))) return rc;
  w->cap_cand = n;
  return FM_OK;
}
static int ensure_survivors(Workspace* w, int64_t n) {
  if (n <= w->cap_surv) return FM_OK;
  int rc;
  if (n > (int64_t(1) << 28)) { set_error("too many surviving candidates in one batch: split the batch"); return FM_ERR_NOMEM; }
  n = std::min<int64_t>(n + n / 4, int64_t(1) << 28);
  uint32_t hs = 1u << 20;
  while ((int64_t)hs < 4 * n) hs <<= 1;
  if ((rc = dev_realloc(&w->surv, n)) || (rc = dev_realloc(&w->surv_len, n)) || (rc = dev_realloc(&w->rec, n)) ||
      (rc = dev_realloc(&w->heapbuf, n + w->cap_q + 1)) || (rc = dev_realloc(&w->sort_key, n)) || (rc = dev_realloc(&w->sort_key2, n)) || (rc = dev_realloc(&w->sort_idx, n)) || (rc = dev_realloc(&w->hkey, (size_t)hs)) ||
      (rc = dev_realloc(&w->hlm, (size_t)hs)))
    return rc;
  if (w->want_stage && (rc = dev_realloc(&w->wire_stage, n))) return rc;
  w->hsize = hs;
  w->cap_surv = n;
  return FM_OK;
}
// sharded TM: the staging buffer for accepted records follows the survivor capacity
static int ensure_stage(Workspace* w) {
  if (w->want_stage && w->wire_stage) return FM_OK;
  w->want_stage = true;
  if (w->cap_surv == 0) return FM_OK;  // allocated with the survivor arrays
  return dev_realloc(&w->wire_stage, (size_t)w->cap_surv);
}
static int ensure_out(Workspace* w, int64_t n_q, int64_t cap) {
  if (n_q * cap <= w->cap_out) return FM_OK;
  int rc;
  const int64_t c = n_q * cap + n_q * cap / 4 + 1024;
  if ((rc = dev_realloc(&w->d_out, c))) return rc;
  w->cap_out = c;
  return FM_OK;
}

static void free_workspace(Workspace* w) {
  cudaFree(w->d_prior); cudaFree(w->d_prior_off);
  cudaFree(w->ctok); cudaFree(w->c_cnt); cudaFree(w->c_base);
  if (w->h_ctotal) cudaFreeHost(w->h_ctotal);
  cudaFree(w->d_q_tok); cudaFree(w->d_q_off); cudaFree(w->d_q_real); cudaFree(w->d_q_gap); cudaFree(w->d_itok_dist); cudaFree(w->pat); cudaFree(w->chain_rec); cudaFree(w->prep_list); cudaFree(w->qmeta); cudaFree(w->tbl); cudaFree(w->cmin_tab); cudaFree(w->span_slice); cudaFree(w->qmask); cudaFree(w->qmask2); cudaFree(w->wq); cudaFree(w->peq64); cudaFree(w->cmin64); cudaFree(w->sm_rec);
  cudaFree(w->sl_start); cudaFree(w->sl_rec); cudaFree(w->hkey); cudaFree(w->hlm); cudaFree(w->surv); cudaFree(w->cand); cudaFree(w->surv_len);
  cudaFree(w->q_cnt); cudaFree(w->q_base); cudaFree(w->acc_cnt); cudaFree(w->rec); cudaFree(w->heapbuf); cudaFree(w->ctr); cudaFree(w->mctr); cudaFree(w->wire_stage); cudaFree(w->wire_send); cudaFree(w->wire_recv); cudaFree(w->scan_chain);
  cudaFree(w->d_out); cudaFree(w->d_out_count); cudaFree(w->mrec); cudaFree(w->m_cnt); cudaFree(w->m_base); cudaFree(w->m_acc); cudaFree(w->m_heap); cudaFree(w->heavy_q); cudaFree(w->m_heavy); cudaFree(w->mid_q); cudaFree(w->m_mid); cudaFree(w->sort_key); cudaFree(w->sort_key2); cudaFree(w->m_key2); cudaFree(w->sort_idx); cudaFree(w->m_key); cudaFree(w->m_idx);
  if (w->h_ctr) cudaFreeHost(w->h_ctr);
  if (w->h_mctr) cudaFreeHost(w->h_mctr);
  if (w->h_q_off32) cudaFreeHost(w->h_q_off32);
  if (w->stream) {
    cudaStreamDestroy(w->stream);
    if (w->stream2) cudaStreamDestroy(w->stream2);
    if (w->ev_fork) cudaEventDestroy(w->ev_fork);
    if (w->ev_join) cudaEventDestroy(w->ev_join);
    if (w->ev_done) cudaEventDestroy(w->ev_done);
    for (auto& e : w->ev) cudaEventDestroy(e);
  }
  delete w;
}

static int check_params(const fm_params* p, Params* out) {
  if (!p) { set_error("params is NULL"); return FM_ERR_INVALID; }
  out->fuzzy = p->fuzzy;
  out->n_matches = p->number_of_matches;
  out->no_perfect = p->no_perfect;
  out->ml = p->min_subseq_length;
  out->mr = p->min_subseq_ratio;
  out->idf_penalty = p->vocab_idf_penalty;
  out->ins = p->insert_cost; out->del = p->delete_cost; out->rep = p->replace_cost;
  out->contrast = p->contrastive_factor;
  out->reduce = p->contrast_reduce;
  out->buffer = p->contrast_buffer == -1 ? p->number_of_matches : p->contrast_buffer;  // src/fuzzy_match.cc:451-452
  if (p->number_of_matches < 0) { set_error("number_of_matches < 0"); return FM_ERR_INVALID; }
  return FM_OK;
}

// FM_DEBUG_SYNC=1: wait after every stage and name the one that failed (diagnostics only)
static const bool g_debug_sync = getenv("FM_DEBUG_SYNC") != nullptr;
static int stage_check(cudaStream_t st, const char* what) {
  if (!g_debug_sync) return FM_OK;
  cudaError_t e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, what);
  return FM_OK;
}

static BatchDev make_batch(Workspace* w, const int32_t* d_q_tok, const int32_t* d_q_off, int64_t n_q, int64_t n_tok) {
  BatchDev b{};
  b.q_tok_in = d_q_tok; b.q_off = d_q_off; b.n_q = (int32_t)n_q; b.n_tok = (int32_t)n_tok;
  b.pat = w->pat; b.chain_rec = w->chain_rec; b.prep_list = w->prep_list; b.qmeta = w->qmeta; b.tbl = w->tbl; b.cmin_tab = w->cmin_tab; b.cmin64 = w->cmin64; b.qmask = w->qmask; b.qmask2 = w->qmask2; b.wq = w->cap_wq ? w->wq : nullptr; b.peq64 = w->peq64;
  b.span_slice = w->span_slice; b.span_cap = w->cap_spans;
  b.sl_start = w->sl_start; b.sl_rec = w->sl_rec; b.sm_rec = w->sm_rec; b.slice_cap = w->cap_slices;
  b.hkey = w->hkey; b.hlm = w->hlm; b.hmask = w->hs_use - 1;
  b.cand = w->cand; b.cand_cap = w->cap_cand;
  b.surv = w->surv; b.surv_len = w->surv_len; b.surv_cap = std::min<int64_t>(w->cap_surv, w->hs_use / 4);  // load factor <= 1/4
  b.q_cnt = w->q_cnt; b.q_base = w->q_base; b.rec = w->rec; b.heapbuf = w->heapbuf; b.acc_cnt = w->acc_cnt;
  b.ctr = w->ctr;
  if (w->real_active) { b.q_real = w->d_q_real; b.q_gap = w->d_q_gap; b.itok_dist = w->d_itok_dist; b.n_itok = w->n_itok; }
  return b;
}

// Shard half of the pipeline: prepare -> search -> gather -> scan -> score, all asynchronous on st.
// Afterwards w->q_base / w->rec hold the scored candidates grouped by query (unless a worklist
// overflowed, which later kernels detect through the counters and skip).
static int launch_shard(Index* ix, Workspace* w, const int32_t* d_q_tok, const int32_t* d_q_off, int64_t n_q, int64_t n_tok,
                        const Params& pr, cudaStream_t st, int* launches) {
  // The dedup table is cleared per batch, so only as much of it is used as the previous batch on this
  // workspace suggests (twice its survivors at load factor 1/4; a batch that outgrows it is rerun on the
  // whole table by the overflow path).
  {
    const int64_t hint = w->surv_hint >= 0 ? w->surv_hint : n_q / 2;
    uint32_t hs = 1u << 18;
    while (hs < w->hsize && (int64_t)hs < 4 * (2 * hint + 32768)) hs <<= 1;
    w->hs_use = hs;
  }
  BatchDev b = make_batch(w, d_q_tok, d_q_off, n_q, n_tok);
  FM_CUDA(cudaMemsetAsync(w->ctr, 0, sizeof(Counters), st));
  FM_CUDA(cudaMemsetAsync(w->hkey, 0xff, (size_t)w->hs_use * sizeof(unsigned long long), st));
  FM_CUDA(cudaMemsetAsync(w->hlm, 0, (size_t)w->hs_use * sizeof(unsigned int), st));
  if (ix->profiling) cudaEventRecord(w->ev[0], st);
  if (!w->bounds_valid || memcmp(&w->bounds_params, &pr, sizeof(Params)) != 0) {  // per-length bound tables follow the parameters
    launch_bounds(ix->dev, b, pr, st);
    w->bounds_params = pr;
    w->bounds_valid = true;
    (*launches)++;
  }
  int rc;
  if ((rc = stage_check(st, "memsets / bounds"))) return rc;
  *launches += launch_prepare(ix->dev, b, pr, ix->sm_count, st) - 1;
  if (ix->profiling) cudaEventRecord(w->ev[1], st);
  if ((rc = stage_check(st, "prepare kernel"))) return rc;
  launch_search(ix->dev, b, pr, st);
  if (ix->profiling) cudaEventRecord(w->ev[2], st);
  if ((rc = stage_check(st, "search kernel"))) return rc;
  launch_gather(ix->dev, b, pr, ix->sm_count, st, ix->profiling ? w->ev[8] : nullptr);
  if (ix->profiling) cudaEventRecord(w->ev[3], st);
  if ((rc = stage_check(st, "gather kernels"))) return rc;
  launch_scan(w->q_cnt, w->q_base, (int32_t)n_q, w->scan_chain, ++w->scan_epoch, ix->sm_count, st);
  if (ix->profiling) cudaEventRecord(w->ev[4], st);
  if ((rc = stage_check(st, "scan kernel"))) return rc;
  launch_score(ix->dev, b, pr, ix->sm_count, st);
  if (ix->profiling) cudaEventRecord(w->ev[5], st);
  if ((rc = stage_check(st, "score kernels"))) return rc;
  *launches += w->real_active ? 6 : 7;  // prepare, search, gather (walk + verify), scan, score (short + wavefront kernels)
  return FM_OK;
}

static int initial_worklists(Index* ix, Workspace* w, int64_t n_q, int64_t n_tok) {
  int rc;
  if ((rc = ensure_bounds(ix, w)) || (rc = ensure_wide(ix, w))) return rc;
  if ((rc = ensure_slices(w, std::min<int64_t>((int64_t(1) << 26) - (1 << 17), std::max<int64_t>(1 << 16, 8 * n_tok + 65536))))) return rc;
  if ((rc = ensure_spans(w, std::max<int64_t>(1 << 16, 2 * n_tok + 65536)))) return rc;
  if ((rc = ensure_cand(w, std::max<int64_t>(1 << 20, 16 * n_q)))) return rc;
  return ensure_survivors(w, std::max<int64_t>(1 << 18, 8 * n_q));
}

// Behind a batch: read the counters back and mark the point with the workspace's event.
static int enqueue_done(Workspace* w, cudaStream_t st) {
  FM_CUDA(cudaMemcpyAsync(w->h_ctr, w->ctr, sizeof(Counters), cudaMemcpyDeviceToHost, st));
  FM_CUDA(cudaEventRecord(w->ev_done, st));
  return FM_OK;
}

// Waits for that point (not for anything enqueued on the stream afterwards). Returns 1 if a worklist
// overflowed and was regrown (the caller reruns the batch), 0 if the batch is complete, <0 on error (-rc).
static int wait_and_check(Workspace* w, int attempt, int* retries) {
  cudaError_t e = cudaEventSynchronize(w->ev_done);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return -cuda_fail(e, "batch execution");
  if (!w->h_ctr->overflow) {
    w->surv_hint = (int64_t)w->h_ctr->n_surv;
    return 0;
  }
  if (attempt >= 8) { set_error("workspace overflow persists"); return -FM_ERR_NOMEM; }
  (*retries)++;
  int rc;
  const int64_t need_slices = std::max<int64_t>((int64_t)(w->h_ctr->slice_elem >> kElemBits), (int64_t)w->h_ctr->n_small);
  if (need_slices > w->cap_slices && (rc = ensure_slices(w, need_slices + need_slices / 8 + 1024))) return -rc;
  if (w->h_ctr->overflow & 2u) {
    if (w->hs_use < w->hsize) w->surv_hint = w->cap_surv;  // first the whole table ...
    else if ((rc = ensure_survivors(w, w->cap_surv * 4))) return -rc;  // ... then a larger one
    else w->surv_hint = w->cap_surv;
  }
  // the walk kernel counts every candidate, also those that did not fit: the list is regrown to the exact need
  if ((w->h_ctr->overflow & 8u) && (rc = ensure_cand(w, (int64_t)w->h_ctr->n_cand + 1024))) return -rc;
  const int64_t need_spans = (int64_t)((w->h_ctr->slice_elem & ((1ull << kElemBits) - 1)) / kSpanHost) + 2;
  if (need_spans > w->cap_spans && (rc = ensure_spans(w, need_spans + need_spans / 8))) return -rc;
  return 1;
}

static int run_replay(Index* ix, Workspace* w, fm_record* rec, const int32_t* q_cnt, const int32_t* q_base, float* heapbuf,
                      unsigned long long* sort_key, unsigned long long* sort_key2, int32_t* sort_idx, int32_t* acc_cnt,
                      int32_t* mid_q, int32_t* heavy_q, const int32_t* d_q_off, int64_t n_q, const Params& pr, int64_t cap, fm_match* d_out, int32_t* d_out_count,
                      cudaStream_t st, int* launches, Counters* ctr = nullptr, int32_t* wire_cnt = nullptr, fm_wire* wire_stage = nullptr,
                      bool defer_contrast = false) {
  if (!ctr) ctr = w->ctr;
  launch_replay(ix->dev, rec, q_cnt, q_base, heapbuf, sort_key, sort_key2, sort_idx, acc_cnt, mid_q, heavy_q, d_q_off, (int32_t)n_q, pr, cap, d_out,
                d_out_count, ctr, ix->sm_count, st, w->stream2, w->ev_fork, w->ev_join, wire_cnt, wire_stage);
  (*launches) += 3;
  {
    int rc;
    if ((rc = stage_check(st, "replay kernels"))) return rc;
  }
  if (pr.contrast > 0.f && !wire_cnt && !defer_contrast) {
    launch_contrast(ix->dev, rec, q_base, sort_idx, acc_cnt, (int32_t)n_q, pr, cap, d_out, d_out_count, ctr, ix->sm_count, st,
                    w->prior_active ? w->d_prior : nullptr, w->prior_active ? w->d_prior_off : nullptr);
    (*launches)++;
  }
  if (ix->profiling) cudaEventRecord(w->ev[6], st);
  return FM_OK;
}

static void finish_profile(Index* ix, Workspace* w, int64_t n_q, int64_t n_tok, int launches, int retries) {
  if (!ix->profiling) return;
  fm_profile p{};
  float ms = 0;
  cudaEventElapsedTime(&ms, w->ev[0], w->ev[1]); p.ms_prepare = ms;
  cudaEventElapsedTime(&ms, w->ev[1], w->ev[2]); p.ms_search = ms;
  cudaEventElapsedTime(&ms, w->ev[2], w->ev[3]); p.ms_gather = ms;
  cudaEventElapsedTime(&ms, w->ev[2], w->ev[8]); p.ms_walk = ms;
  cudaEventElapsedTime(&ms, w->ev[8], w->ev[3]); p.ms_verify = ms;
  cudaEventElapsedTime(&ms, w->ev[3], w->ev[4]); p.ms_scan = ms;
  cudaEventElapsedTime(&ms, w->ev[4], w->ev[5]); p.ms_score = ms;
  cudaEventElapsedTime(&ms, w->ev[5], w->ev[6]); p.ms_replay = ms;
  cudaEventElapsedTime(&ms, w->ev[0], w->ev[6]); p.ms_total = ms;
  p.n_queries = n_q; p.n_query_tokens = n_tok;
  p.n_slices = (int64_t)(w->h_ctr->slice_elem >> kElemBits) + (int64_t)w->h_ctr->n_small;
  p.n_elements = (int64_t)(w->h_ctr->slice_elem & ((1ull << kElemBits) - 1));
  p.n_survivors = w->h_ctr->n_surv;
  p.n_stage2 = w->h_ctr->n_cand;
  p.n_verified = w->h_ctr->n_verified;
  p.n_matches = w->h_ctr->n_matches;
  p.launches = launches; p.retries = retries;
  std::lock_guard<std::mutex> g(ix->mu);
  ix->last_profile = p;
}

// Device-resident batch: enqueue the whole pipeline on the caller's stream; the ticket is completed by
// finish_device (waits for the batch's own event, reruns after regrowing a worklist that overflowed).
struct DeviceJob {
  Workspace* w = nullptr;
  const int32_t* d_q_tok = nullptr;
  const int32_t* d_q_off = nullptr;
  int64_t n_q = 0, n_tok = 0, cap = 0;
  fm_match* d_out = nullptr;
  int32_t* d_out_count = nullptr;
  cudaStream_t st = nullptr;
  int launches = 0;
};

static int enqueue_device(Index* ix, DeviceJob& j, const Params& pr) {
  int rc;
  if ((rc = launch_shard(ix, j.w, j.d_q_tok, j.d_q_off, j.n_q, j.n_tok, pr, j.st, &j.launches))) return rc;
  if ((rc = run_replay(ix, j.w, j.w->rec, j.w->q_cnt, j.w->q_base, j.w->heapbuf, j.w->sort_key, j.w->sort_key2, j.w->sort_idx, j.w->acc_cnt,
                       j.w->mid_q, j.w->heavy_q, j.d_q_off, j.n_q, pr, j.cap, j.d_out, j.d_out_count, j.st, &j.launches)))
    return rc;
  return enqueue_done(j.w, j.st);
}

static int submit_device(Index* ix, DeviceJob& j, const Params& pr) {
  int rc;
  j.w->real_active = false;
  j.w->prior_active = false;
  if ((rc = ensure_queries(j.w, j.n_q, j.n_tok, false)) || (rc = initial_worklists(ix, j.w, j.n_q, j.n_tok))) return rc;
  return enqueue_device(ix, j, pr);
}

static int finish_device(Index* ix, DeviceJob& j, const Params& pr) {
  int retries = 0;
  for (int attempt = 0;; attempt++) {
    const int again = wait_and_check(j.w, attempt, &retries);
    if (again < 0) return -again;
    if (!again) break;
    int rc;
    if ((rc = enqueue_device(ix, j, pr))) return rc;
  }
  finish_profile(ix, j.w, j.n_q, j.n_tok, j.launches, retries);
  return FM_OK;
}

// One chunk of a host batch in flight on its own workspace / stream.
struct HostChunk {
  Workspace* w = nullptr;
  int64_t q0 = 0, nq = 0, ntok = 0;
  bool in_flight = false;
  int launches = 0;
};

struct RealInputs {  // extras of a host batch (all NULL / 0 when absent): Sentence API; entries already in `matches`
  const int32_t* q_real = nullptr;
  const int32_t* q_gaps = nullptr;
  const int32_t* itok_dist = nullptr;
  int32_t n_itok = 0;
  const uint32_t* prior_sid = nullptr;  // fm_match_batch_prior: sentence ids per query, CSR by prior_off
  const int64_t* prior_off = nullptr;
};

// Entries already in the callers' result vectors -> (sentence start, length) per entry on the device.
static int stage_prior(Index* ix, Workspace* w, const HostChunk& c, const RealInputs& ri, cudaStream_t st) {
  w->prior_active = ri.prior_sid != nullptr;
  if (!w->prior_active) return FM_OK;
  int rc;
  const int64_t p0 = ri.prior_off[c.q0], np = ri.prior_off[c.q0 + c.nq] - p0;
  if (np > w->cap_prior) {
    if ((rc = dev_realloc(&w->d_prior, np + np / 4 + 256))) return rc;
    w->cap_prior = np + np / 4 + 256;
  }
  if (c.nq + 1 > w->cap_prior_q) {
    if ((rc = dev_realloc(&w->d_prior_off, c.nq + 1 + 256))) return rc;
    w->cap_prior_q = c.nq + 1 + 256;
  }
  std::vector<int2> ent((size_t)np);
  std::vector<int32_t> off((size_t)c.nq + 1);
  for (int64_t i = 0; i <= c.nq; i++) off[(size_t)i] = (int32_t)(ri.prior_off[c.q0 + i] - p0);
  for (int64_t k = 0; k < np; k++) {
    const int64_t s = (int64_t)ri.prior_sid[p0 + k] - (int64_t)ix->dev.sid_base;
    if (s < 0 || s >= ix->n_sent) { set_error("prior match: sentence id not in this index"); return FM_ERR_INVALID; }
    const int32_t start = ix->h_sent_start[(size_t)s];
    int32_t n = 0;
    while (ix->h_tok[(size_t)start + n] != 0) n++;
    ent[(size_t)k] = make_int2(start, n);
  }
  // pageable sources: the copies are staged before cudaMemcpyAsync returns, so the vectors may go out of scope
  if (np) FM_CUDA(cudaMemcpyAsync(w->d_prior, ent.data(), (size_t)np * sizeof(int2), cudaMemcpyHostToDevice, st));
  FM_CUDA(cudaMemcpyAsync(w->d_prior_off, off.data(), (size_t)(c.nq + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  FM_CUDA(cudaStreamSynchronize(st));
  return FM_OK;
}

static int stage_real(Workspace* w, const HostChunk& c, const int64_t* q_off, const RealInputs& ri, cudaStream_t st) {
  w->real_active = ri.q_real != nullptr;
  if (!w->real_active) return FM_OK;
  int rc;
  if (c.ntok > w->cap_real_tok) {
    if ((rc = dev_realloc(&w->d_q_real, c.ntok + c.ntok / 4 + 1024))) return rc;
    w->cap_real_tok = c.ntok + c.ntok / 4 + 1024;
  }
  if (c.ntok + c.nq > w->cap_real_gap) {
    const int64_t n = c.ntok + c.nq + (c.ntok + c.nq) / 4 + 1024;
    if ((rc = dev_realloc(&w->d_q_gap, n))) return rc;
    w->cap_real_gap = n;
  }
  const int64_t nd = (int64_t)ri.n_itok * ri.n_itok;
  if (nd > w->cap_itok) {
    if ((rc = dev_realloc(&w->d_itok_dist, nd))) return rc;
    w->cap_itok = nd;
  }
  w->n_itok = ri.n_itok;
  if (c.ntok) FM_CUDA(cudaMemcpyAsync(w->d_q_real, ri.q_real + q_off[c.q0], c.ntok * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  FM_CUDA(cudaMemcpyAsync(w->d_q_gap, ri.q_gaps + q_off[c.q0] + c.q0, (c.ntok + c.nq) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  FM_CUDA(cudaMemcpyAsync(w->d_itok_dist, ri.itok_dist, nd * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  return FM_OK;
}

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static const bool g_host_timing = getenv("FM_HOST_TIMING") != nullptr;

// Enqueue H2D + the whole pipeline + D2H of one chunk; returns without waiting.
static int launch_host_chunk(Index* ix, HostChunk& c, const int32_t* q_tokens, const int64_t* q_off, const Params& pr, int64_t cap,
                             fm_match* out, int32_t* out_count, const RealInputs& ri) {
  Workspace* w = c.w;
  int rc;
  const double t0 = g_host_timing ? now_ms() : 0;
  if ((rc = ensure_queries(w, c.nq, c.ntok, true)) || (rc = ensure_out(w, c.nq, cap)) || (rc = initial_worklists(ix, w, c.nq, c.ntok)))
    return rc;
  cudaStream_t st = w->stream;
  c.in_flight = true;  // from here on the stream may hold work of this chunk: an error return must wait for it
  if (g_host_timing) cudaEventRecord(w->ev[7], st);
  // the token copy does not need the converted offsets: start it first and narrow the offsets to int32
  // on the host while it is in flight
  if (c.ntok) FM_CUDA(cudaMemcpyAsync(w->d_q_tok, q_tokens + q_off[c.q0], c.ntok * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  const double t1 = g_host_timing ? now_ms() : 0;
  for (int64_t i = 0; i <= c.nq; i++) w->h_q_off32[i] = (int32_t)(q_off[c.q0 + i] - q_off[c.q0]);
  FM_CUDA(cudaMemcpyAsync(w->d_q_off, w->h_q_off32, (c.nq + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  if ((rc = stage_real(w, c, q_off, ri, st)) || (rc = stage_prior(ix, w, c, ri, st))) return rc;
  FM_CUDA(cudaMemsetAsync(w->d_out, 0, c.nq * cap * sizeof(fm_match), st));  // slots past the count read as zero
  c.launches = 0;
  if ((rc = launch_shard(ix, w, w->d_q_tok, w->d_q_off, c.nq, c.ntok, pr, st, &c.launches))) return rc;
  if ((rc = run_replay(ix, w, w->rec, w->q_cnt, w->q_base, w->heapbuf, w->sort_key, w->sort_key2, w->sort_idx, w->acc_cnt, w->mid_q, w->heavy_q, w->d_q_off,
                       c.nq, pr, cap, w->d_out, w->d_out_count, st, &c.launches)))
    return rc;
  FM_CUDA(cudaMemcpyAsync(out + c.q0 * cap, w->d_out, c.nq * cap * sizeof(fm_match), cudaMemcpyDeviceToHost, st));
  FM_CUDA(cudaMemcpyAsync(out_count + c.q0, w->d_out_count, c.nq * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if ((rc = enqueue_done(w, st))) return rc;
  if (g_host_timing) {
    const double t2 = now_ms();
    cudaStreamSynchronize(st);
    const double t3 = now_ms();
    fprintf(stderr, "[fm host] ensure+convert %.3f ms, enqueue %.3f ms, wait %.3f ms, total %.3f ms\n", t1 - t0, t2 - t1, t3 - t2, t3 - t0);
  }
  return FM_OK;
}

// Wait for a chunk; if one of its worklists overflowed, regrow and rerun it synchronously.
static int finish_host_chunk(Index* ix, HostChunk& c, const Params& pr, int64_t cap, fm_match* out, int32_t* out_count) {
  if (!c.in_flight) return FM_OK;
  c.in_flight = false;
  Workspace* w = c.w;
  cudaStream_t st = w->stream;
  int retries = 0;
  for (int attempt = 0;; attempt++) {
    const int again = wait_and_check(w, attempt, &retries);
    if (again < 0) return -again;
    if (!again) break;
    int rc;
    FM_CUDA(cudaMemsetAsync(w->d_out, 0, c.nq * cap * sizeof(fm_match), st));
    if ((rc = launch_shard(ix, w, w->d_q_tok, w->d_q_off, c.nq, c.ntok, pr, st, &c.launches))) return rc;
    if ((rc = run_replay(ix, w, w->rec, w->q_cnt, w->q_base, w->heapbuf, w->sort_key, w->sort_key2, w->sort_idx, w->acc_cnt, w->mid_q, w->heavy_q,
                         w->d_q_off, c.nq, pr, cap, w->d_out, w->d_out_count, st, &c.launches)))
      return rc;
    FM_CUDA(cudaMemcpyAsync(out + c.q0 * cap, w->d_out, c.nq * cap * sizeof(fm_match), cudaMemcpyDeviceToHost, st));
    FM_CUDA(cudaMemcpyAsync(out_count + c.q0, w->d_out_count, c.nq * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if ((rc = enqueue_done(w, st))) return rc;
  }
  finish_profile(ix, w, c.nq, c.ntok, c.launches, retries);
  return FM_OK;
}

// fm_match_batch / fm_match_batch_real: chunk the host batch, keep up to kSlots chunks in flight.
static int match_batch_host(fm_index* index, const int32_t* q_tokens, const int64_t* q_off, int64_t n_q, const fm_params* params,
                            int64_t cap, fm_match* out, int32_t* out_count, const RealInputs& ri) {
  Index* ix = reinterpret_cast<Index*>(index);
  Params pr;
  int rc;
  if (!ix || n_q < 0 || cap < 1 || (n_q > 0 && (!q_off || !out || !out_count))) { set_error("bad argument"); return FM_ERR_INVALID; }
  if (ri.q_real && (!ix->dev.real || !ri.q_gaps || !ri.itok_dist || ri.n_itok < 1 || ri.n_itok > 2048)) {
    set_error("Sentence API needs fm_index_set_real on the index, gaps and a penalty-token table (1..2048 entries)");
    return FM_ERR_INVALID;
  }
  if ((rc = check_params(params, &pr))) return rc;
  if (n_q == 0) return FM_OK;
  if (ri.q_real) {  // every penalty-token id indexes the caller's n_itok x n_itok table on the device
    if (ix->max_gap_id >= ri.n_itok) { set_error("penalty-token table smaller than the ids stored with the TM"); return FM_ERR_INVALID; }
    const int64_t n_gap = q_off[n_q] - q_off[0] + n_q;
    const int32_t* g = ri.q_gaps + q_off[0];
    for (int64_t i = 0; i < n_gap; i++)
      if (g[i] < 0 || g[i] >= ri.n_itok) { set_error("query penalty-token id outside the table"); return FM_ERR_INVALID; }
  }
  FM_CUDA(cudaSetDevice(ix->device));
  // The batch is cut into chunks that run on up to kSlots workspaces / streams, so the H2D copy of
  // one chunk and the D2H copy of another overlap the kernels of a third. Chunks also keep the
  // offsets int32 and the workspaces bounded.
  const int kSlots = 3;
  const int64_t kMaxTok = 1 << 22;
  // (measured on B200: for 100k-query batches one chunk is fastest -- per-chunk launch and tail costs
  // outweigh the copy overlap -- so chunking only bounds very large batches unless FM_HOST_CHUNKS asks)
  static const int n_chunks = getenv("FM_HOST_CHUNKS") ? std::max(1, atoi(getenv("FM_HOST_CHUNKS"))) : 1;
  const int64_t chunk_q = std::min<int64_t>(1 << 18, std::max<int64_t>(16384, (n_q + n_chunks - 1) / n_chunks));
  HostChunk slots[kSlots];
  struct Releaser {
    Index* ix; HostChunk* s; int n;
    ~Releaser() { for (int i = 0; i < n; i++) if (s[i].w) { if (s[i].in_flight) cudaStreamSynchronize(s[i].w->stream); release(ix, s[i].w); } }
  } rel{ix, slots, kSlots};
  int64_t q0 = 0;
  for (int k = 0; q0 < n_q; k++) {
    int64_t q1 = std::min(n_q, q0 + chunk_q);
    while (q1 > q0 + 1 && q_off[q1] - q_off[q0] > kMaxTok) q1 = q0 + (q1 - q0) / 2;
    const int64_t ntok = q_off[q1] - q_off[q0];
    if (ntok < 0 || ntok > (int64_t(1) << 29)) { set_error("bad q_off"); return FM_ERR_INVALID; }
    HostChunk& c = slots[k % kSlots];
    if ((rc = finish_host_chunk(ix, c, pr, cap, out, out_count))) return rc;
    if (!c.w) {
      c.w = acquire(ix);
      if ((rc = ensure_base(c.w))) return rc;
    }
    c.q0 = q0; c.nq = q1 - q0; c.ntok = ntok;
    if ((rc = launch_host_chunk(ix, c, q_tokens, q_off, pr, cap, out, out_count, ri))) return rc;
    q0 = q1;
  }
  for (int i = 0; i < kSlots; i++)
    if ((rc = finish_host_chunk(ix, slots[i], pr, cap, out, out_count))) return rc;
  return FM_OK;
}

}  // namespace fm

using namespace fm;

extern "C" {

const char* fm_last_error(void) { return fm::get_error().c_str(); }
const char* fm_version(void) { return "fuzzy_match_b200 0.1 (sm_100a)"; }

int fm_index_create(const int32_t* tokens, const int64_t* sent_off, int64_t n_sent, int32_t vocab_size,
                    int32_t max_tokens_in_pattern, const uint32_t* sfreq_global, int64_t n_sent_global, int64_t s_id_base,
                    int device, fm_index** out) {
  Index* ix = nullptr;
  const int rc = build_index(tokens, sent_off, n_sent, vocab_size, max_tokens_in_pattern, sfreq_global, n_sent_global,
                             s_id_base, device, &ix);
  if (out) *out = reinterpret_cast<fm_index*>(ix);
  return rc;
}

int fm_index_save(const fm_index* index, const char* path) {
  const Index* ix = reinterpret_cast<const Index*>(index);
  if (!ix || !path) { set_error("NULL argument"); return FM_ERR_INVALID; }
  return save_index(ix, path);
}
int fm_index_load(const char* path, int device, fm_index** out) {
  if (!path || !out) { set_error("NULL argument"); return FM_ERR_INVALID; }
  Index* ix = nullptr;
  const int rc = load_index(path, device, &ix);
  *out = reinterpret_cast<fm_index*>(ix);
  return rc;
}

void fm_index_destroy(fm_index* index) {
  Index* ix = reinterpret_cast<Index*>(index);
  if (!ix) return;
  cudaSetDevice(ix->device);
  for (Workspace* w : ix->pool) free_workspace(w);
  free_index(ix);
}

int64_t fm_index_num_sentences(const fm_index* index) { return reinterpret_cast<const Index*>(index)->n_sent; }
int64_t fm_index_num_suffixes(const fm_index* index) { return reinterpret_cast<const Index*>(index)->n_suf; }
int32_t fm_index_max_tokens_in_pattern(const fm_index* index) { return reinterpret_cast<const Index*>(index)->max_tokens; }
int64_t fm_index_device_bytes(const fm_index* index) { return reinterpret_cast<const Index*>(index)->device_bytes; }

int fm_index_kept_sources(const fm_index* index, int64_t* kept) {
  const Index* ix = reinterpret_cast<const Index*>(index);
  if (!ix || !kept) { set_error("NULL argument"); return FM_ERR_INVALID; }
  std::copy(ix->kept.begin(), ix->kept.end(), kept);
  return FM_OK;
}
int fm_index_sfreq(const fm_index* index, uint32_t* sfreq) {
  const Index* ix = reinterpret_cast<const Index*>(index);
  if (!ix || !sfreq) { set_error("NULL argument"); return FM_ERR_INVALID; }
  std::copy(ix->sfreq.begin(), ix->sfreq.end(), sfreq);
  return FM_OK;
}
int fm_index_set_idf_stats(fm_index* index, const uint32_t* sfreq_global, int64_t n_sent_global) {
  Index* ix = reinterpret_cast<Index*>(index);
  if (!ix || !sfreq_global || n_sent_global < 1) { set_error("bad argument"); return FM_ERR_INVALID; }
  FM_CUDA(cudaSetDevice(ix->device));
  return set_idf_stats(ix, sfreq_global, n_sent_global);
}
int fm_index_sentence(const fm_index* index, uint32_t s, const int32_t** tokens, int32_t* length) {
  const Index* ix = reinterpret_cast<const Index*>(index);
  if (!ix || (int64_t)s >= ix->n_sent) { set_error("sentence id out of range"); return FM_ERR_INVALID; }
  const int32_t st = ix->h_sent_start[s];
  if (tokens) *tokens = ix->h_tok.data() + st;
  if (length) {
    int32_t n = 0;
    while (ix->h_tok[st + n] != 0) n++;
    *length = n;
  }
  return FM_OK;
}

int fm_match_batch(fm_index* index, const int32_t* q_tokens, const int64_t* q_off, int64_t n_q, const fm_params* params,
                   int64_t cap, fm_match* out, int32_t* out_count) {
  return match_batch_host(index, q_tokens, q_off, n_q, params, cap, out, out_count, RealInputs());
}

int fm_match_batch_real(fm_index* index, const int32_t* q_tokens, const int32_t* q_real, const int32_t* q_gaps,
                        const int64_t* q_off, int64_t n_q, const fm_params* params, const int32_t* itok_dist, int32_t n_itok,
                        int64_t cap, fm_match* out, int32_t* out_count) {
  if (!q_real) { set_error("q_real is NULL"); return FM_ERR_INVALID; }
  RealInputs ri;
  ri.q_real = q_real; ri.q_gaps = q_gaps; ri.itok_dist = itok_dist; ri.n_itok = n_itok;
  return match_batch_host(index, q_tokens, q_off, n_q, params, cap, out, out_count, ri);
}

int fm_match_batch_prior(fm_index* index, const int32_t* q_tokens, const int64_t* q_off, int64_t n_q, const fm_params* params,
                         const uint32_t* prior_sid, const int64_t* prior_off, int64_t cap, fm_match* out, int32_t* out_count) {
  if (!prior_off || (!prior_sid && n_q > 0 && prior_off[n_q] != prior_off[0])) { set_error("bad argument (prior CSR)"); return FM_ERR_INVALID; }
  for (int64_t q = 0; q < n_q; q++)
    if (prior_off[q + 1] < prior_off[q]) { set_error("prior_off not ascending"); return FM_ERR_INVALID; }
  RealInputs ri;
  static const uint32_t none = 0;
  ri.prior_sid = prior_sid ? prior_sid : &none;
  ri.prior_off = prior_off;
  const int rc = match_batch_host(index, q_tokens, q_off, n_q, params, cap, out, out_count, ri);
  if (rc || !params || params->contrastive_factor > 0.f || params->number_of_matches == 0) return rc;
  // without the rerank the earlier entries only count against number_of_matches (src/fuzzy_match.cc:672)
  for (int64_t q = 0; q < n_q; q++) {
    const int64_t room = std::max<int64_t>(0, (int64_t)params->number_of_matches - (prior_off[q + 1] - prior_off[q]));
    if (out_count[q] > room) {
      for (int64_t k = room; k < std::min<int64_t>(out_count[q], cap); k++) out[q * cap + k] = fm_match{};
      out_count[q] = (int32_t)room;
    }
  }
  return rc;
}

int fm_index_set_real(fm_index* index, const int32_t* real, const int32_t* gaps, const int64_t* sent_off, int64_t n_sent) {
  Index* ix = reinterpret_cast<Index*>(index);
  if (!ix || !real || !gaps || !sent_off) { set_error("NULL argument"); return FM_ERR_INVALID; }
  FM_CUDA(cudaSetDevice(ix->device));
  return set_real(ix, real, gaps, sent_off, n_sent);
}

int fm_subsequence_batch(fm_index* index, const int32_t* q_tokens, const int64_t* q_off, int64_t n_q, int32_t number_of_matches,
                         int32_t no_perfect, int32_t min_subseq_length, float min_subseq_ratio, int32_t idf_weighting, fm_subseq* out) {
  Index* ix = reinterpret_cast<Index*>(index);
  if (!ix || n_q < 0 || number_of_matches < 0 || number_of_matches > 4096 || (n_q > 0 && (!q_off || !out))) {
    set_error("bad argument (number_of_matches 0..4096)");
    return FM_ERR_INVALID;
  }
  if (n_q == 0) return FM_OK;
  const int64_t ntok = q_off[n_q] - q_off[0];
  if (n_q > (1 << 20) || ntok < 0 || ntok > (int64_t(1) << 28)) { set_error("batch too large: split it"); return FM_ERR_INVALID; }
  for (int64_t i = 0; i < n_q; i++)
    if (q_off[i + 1] < q_off[i] || q_off[i + 1] - q_off[i] > ix->max_tokens) {
      // the reference has no cap here (only match() ignores such patterns); the kernel's staging buffers do
      set_error("subsequence: pattern longer than max_tokens_in_pattern (or q_off not ascending)");
      return FM_ERR_INVALID;
    }
  FM_CUDA(cudaSetDevice(ix->device));
  Workspace* w = acquire(ix);
  struct Releaser { Index* ix; Workspace* w; ~Releaser() { if (w->stream) cudaStreamSynchronize(w->stream); release(ix, w); } } rel{ix, w};
  int rc;
  if ((rc = ensure_base(w)) || (rc = ensure_queries(w, n_q, ntok, true))) return rc;
  cudaStream_t st = w->stream;
  const int seen_cap = number_of_matches + 64;  // candidates + sentences skipped as perfect
  uint32_t* d_seen = nullptr;
  fm_subseq* d_out = nullptr;
  FM_CUDA(cudaMalloc((void**)&d_seen, (size_t)n_q * seen_cap * sizeof(uint32_t)));
  if (cudaMalloc((void**)&d_out, (size_t)n_q * sizeof(fm_subseq)) != cudaSuccess) { cudaFree(d_seen); set_error("out of device memory"); return FM_ERR_NOMEM; }
  struct Freer { void* a; void* b; ~Freer() { cudaFree(a); cudaFree(b); } } fr{d_seen, d_out};
  if (ntok) FM_CUDA(cudaMemcpyAsync(w->d_q_tok, q_tokens + q_off[0], ntok * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  for (int64_t i = 0; i <= n_q; i++) w->h_q_off32[i] = (int32_t)(q_off[i] - q_off[0]);
  FM_CUDA(cudaMemcpyAsync(w->d_q_off, w->h_q_off32, (n_q + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  launch_subseq(ix->dev, w->d_q_tok, w->d_q_off, (int32_t)n_q, number_of_matches, no_perfect != 0, min_subseq_length, min_subseq_ratio,
                idf_weighting != 0, d_seen, seen_cap, d_out, st);
  FM_CUDA(cudaMemcpyAsync(out, d_out, (size_t)n_q * sizeof(fm_subseq), cudaMemcpyDeviceToHost, st));
  FM_CUDA(cudaStreamSynchronize(st));
  FM_CUDA(cudaGetLastError());
  for (int64_t i = 0; i < n_q; i++)
    if (out[i].found < 0) { set_error("more than 64 perfect matches skipped for one pattern (no_perfect)"); return FM_ERR_NOMEM; }
  return FM_OK;
}

// ---- submit / wait: the asynchronous form of the two batch calls. A ticket owns one workspace; several
// tickets may be in flight on one index, so the host->device copy of batch i+1 and the device->host copy
// of batch i-1 overlap the kernels of batch i (the reference overlaps I/O and matching the same way with
// its queue of futures, cli/src/FuzzyMatch-cli.cc:112-193).
struct fm_ticket {
  Index* ix = nullptr;
  Params pr{};
  int64_t cap = 0;
  bool host = false;
  DeviceJob dev;
  HostChunk chunk;
  fm_match* out = nullptr;
  int32_t* out_count = nullptr;
  fm_comm* comm = nullptr;  // sharded batch (dev holds the job)
  int64_t capacity = 0;     // record capacity of the blocks of the attempt in flight
  int attempts = 0;
};
static int finish_sharded(fm_ticket* t);

static void drop_ticket(fm_ticket* t) {  // after an error: nothing of the batch may still run on the workspace
  Workspace* w = t->host ? t->chunk.w : t->dev.w;
  if (w) {
    cudaStreamSynchronize(t->host ? w->stream : t->dev.st);
    release(t->ix, w);
  }
  delete t;
}

int fm_match_batch_device_submit(fm_index* index, const int32_t* d_q_tokens, const int32_t* d_q_off, int64_t n_q,
                                 int64_t n_query_tokens, const fm_params* params, int64_t cap, fm_match* d_out,
                                 int32_t* d_out_count, void* stream, fm_ticket** ticket) {
  Index* ix = reinterpret_cast<Index*>(index);
  Params pr;
  int rc;
  if (ticket) *ticket = nullptr;
  if (!ix || !ticket || n_q < 1 || cap < 1 || n_q > (1 << 20) || n_query_tokens > (int64_t(1) << 25)) {
    set_error("bad argument (device batches hold 1 .. 2^20 queries and at most 2^25 tokens)");
    return FM_ERR_INVALID;
  }
  if ((rc = check_params(params, &pr))) return rc;
  FM_CUDA(cudaSetDevice(ix->device));
  fm_ticket* t = new fm_ticket();
  t->ix = ix; t->pr = pr; t->cap = cap; t->host = false;
  DeviceJob& j = t->dev;
  j.w = acquire(ix);
  j.d_q_tok = d_q_tokens; j.d_q_off = d_q_off; j.n_q = n_q; j.n_tok = n_query_tokens; j.cap = cap;
  j.d_out = d_out; j.d_out_count = d_out_count; j.st = static_cast<cudaStream_t>(stream);
  if ((rc = ensure_base(j.w)) || (rc = submit_device(ix, j, pr))) { drop_ticket(t); return rc; }
  *ticket = t;
  return FM_OK;
}

int fm_match_batch_submit(fm_index* index, const int32_t* q_tokens, const int64_t* q_off, int64_t n_q, const fm_params* params,
                          int64_t cap, fm_match* out, int32_t* out_count, fm_ticket** ticket) {
  Index* ix = reinterpret_cast<Index*>(index);
  Params pr;
  int rc;
  if (ticket) *ticket = nullptr;
  if (!ix || !ticket || n_q < 1 || cap < 1 || !q_off || !out || !out_count) { set_error("bad argument"); return FM_ERR_INVALID; }
  const int64_t ntok = q_off[n_q] - q_off[0];
  if (n_q > (1 << 18) || ntok < 0 || ntok > (int64_t(1) << 22)) {
    set_error("a submitted batch holds at most 2^18 queries / 2^22 tokens: split it or use fm_match_batch");
    return FM_ERR_INVALID;
  }
  if ((rc = check_params(params, &pr))) return rc;
  FM_CUDA(cudaSetDevice(ix->device));
  fm_ticket* t = new fm_ticket();
  t->ix = ix; t->pr = pr; t->cap = cap; t->host = true; t->out = out; t->out_count = out_count;
  HostChunk& c = t->chunk;
  c.w = acquire(ix);
  c.q0 = 0; c.nq = n_q; c.ntok = ntok;
  if ((rc = ensure_base(c.w)) || (rc = launch_host_chunk(ix, c, q_tokens, q_off, pr, cap, out, out_count, RealInputs()))) {
    drop_ticket(t);
    return rc;
  }
  *ticket = t;
  return FM_OK;
}

int fm_ticket_wait(fm_ticket* t) {
  if (!t) { set_error("NULL ticket"); return FM_ERR_INVALID; }
  cudaSetDevice(t->ix->device);
  const int rc = t->comm ? finish_sharded(t)
                 : t->host ? finish_host_chunk(t->ix, t->chunk, t->pr, t->cap, t->out, t->out_count) : finish_device(t->ix, t->dev, t->pr);
  if (rc) { drop_ticket(t); return rc; }
  release(t->ix, t->host ? t->chunk.w : t->dev.w);
  delete t;
  return FM_OK;
}

int fm_match_batch_device(fm_index* index, const int32_t* d_q_tokens, const int32_t* d_q_off, int64_t n_q,
                          int64_t n_query_tokens, const fm_params* params, int64_t cap, fm_match* d_out, int32_t* d_out_count,
                          void* stream) {
  if (n_q == 0 && index && cap >= 1 && params) return FM_OK;
  fm_ticket* t = nullptr;
  const int rc = fm_match_batch_device_submit(index, d_q_tokens, d_q_off, n_q, n_query_tokens, params, cap, d_out, d_out_count, stream, &t);
  return rc ? rc : fm_ticket_wait(t);
}

// ---------------------------------------------------------------- sharded TM

int64_t fm_wire_block_bytes(int64_t n_q, int64_t capacity) { return n_q < 0 || capacity < 0 ? -1 : wire_block_bytes(n_q, capacity); }

// Enqueue the shard half of a batch on st: the whole pipeline with the replay in shard mode (accepted records
// staged per query), their offsets scanned straight into the block, then packed behind them with the header.
static int enqueue_accept(Index* ix, Workspace* w, const int32_t* d_q_tok, const int32_t* d_q_off, int64_t n_q, int64_t n_tok,
                          const Params& pr, int64_t capacity, void* d_block, cudaStream_t st, int* launches) {
  int rc;
  int32_t* blk = static_cast<int32_t*>(d_block);
  FM_CUDA(cudaMemsetAsync(w->acc_cnt, 0, n_q * sizeof(int32_t), st));  // queries the replay skips count 0
  if ((rc = launch_shard(ix, w, d_q_tok, d_q_off, n_q, n_tok, pr, st, launches))) return rc;
  if ((rc = run_replay(ix, w, w->rec, w->q_cnt, w->q_base, w->heapbuf, w->sort_key, w->sort_key2, w->sort_idx, nullptr, w->mid_q,
                       w->heavy_q, d_q_off, n_q, pr, 1, nullptr, nullptr, st, launches, nullptr, w->acc_cnt, w->wire_stage)))
    return rc;
  launch_scan(w->acc_cnt, blk + 4, (int32_t)n_q, w->scan_chain, ++w->scan_epoch, ix->sm_count, st);
  launch_wire_pack(blk, w->ctr, w->wire_stage, w->q_base, (int32_t)n_q, (int)capacity, st);
  *launches += 2;
  return stage_check(st, "wire pack");
}

// Enqueue the cross-shard half: counts -> scan -> unpack -> the ordinary replay of the union.
static int ensure_merge(Workspace* w, int64_t n_q, int64_t total) {
  int rc;
  if (n_q > w->cap_mq) {
    if ((rc = dev_realloc(&w->m_cnt, n_q + 1)) || (rc = dev_realloc(&w->m_base, n_q + 1)) || (rc = dev_realloc(&w->m_acc, n_q + 1)) ||
        (rc = dev_realloc(&w->m_heavy, n_q + 1)) || (rc = dev_realloc(&w->m_mid, n_q + 1)))
      return rc;
    w->cap_mq = n_q;
    w->cap_mrec = 0;  // m_heap depends on cap_mq
  }
  if (total + n_q + 1 > w->cap_mrec) {
    const int64_t c = total + total / 4 + n_q + 1024;
    if ((rc = dev_realloc(&w->mrec, c)) || (rc = dev_realloc(&w->m_heap, c + w->cap_mq + 1)) || (rc = dev_realloc(&w->m_key, c)) ||
        (rc = dev_realloc(&w->m_key2, c)) || (rc = dev_realloc(&w->m_idx, c)))
      return rc;
    w->cap_mrec = c;
  }
  return FM_OK;
}
// total_capacity = sum of the blocks' capacities (an upper bound on the records of the union)
static int enqueue_merge(Index* ix, Workspace* w, int n_shards, const int32_t* const* blocks, int64_t total_capacity, const int32_t* d_q_off,
                         int64_t n_q, const Params& pr, int64_t cap, fm_match* d_out, int32_t* d_out_count, cudaStream_t st, int* launches,
                         bool defer_contrast = false) {
  int rc;
  if ((rc = ensure_merge(w, n_q, total_capacity))) return rc;
  FM_CUDA(cudaMemsetAsync(w->mctr, 0, sizeof(Counters), st));
  launch_wire_count(n_shards, blocks, w->m_cnt, (int32_t)n_q, w->mctr, st);
  launch_scan(w->m_cnt, w->m_base, (int32_t)n_q, w->scan_chain, ++w->scan_epoch, ix->sm_count, st);
  launch_wire_copy(n_shards, blocks, w->m_base, w->mrec, (int32_t)n_q, w->mctr, st);
  *launches += 3;
  if ((rc = stage_check(st, "wire merge"))) return rc;
  if ((rc = run_replay(ix, w, w->mrec, w->m_cnt, w->m_base, w->m_heap, w->m_key, w->m_key2, w->m_idx, w->m_acc, w->m_mid, w->m_heavy, d_q_off,
                       n_q, pr, cap, d_out, d_out_count, st, launches, w->mctr, nullptr, nullptr, defer_contrast)))
    return rc;
  FM_CUDA(cudaMemcpyAsync(w->h_mctr, w->mctr, sizeof(Counters), cudaMemcpyDeviceToHost, st));
  return FM_OK;
}

int fm_shard_accept_device(fm_index* index, const int32_t* d_q_tokens, const int32_t* d_q_off, int64_t n_q, int64_t n_query_tokens,
                           const fm_params* params, int64_t capacity, void* d_block, void* stream) {
  Index* ix = reinterpret_cast<Index*>(index);
  Params pr;
  int rc;
  if (!ix || n_q < 1 || n_q > (1 << 20) || capacity < 0 || capacity > (int64_t(1) << 30) || !d_block) { set_error("bad argument"); return FM_ERR_INVALID; }
  if ((rc = check_params(params, &pr))) return rc;
  FM_CUDA(cudaSetDevice(ix->device));
  Workspace* w = acquire(ix);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  struct Releaser { Index* ix; Workspace* w; cudaStream_t st; ~Releaser() { cudaStreamSynchronize(st); release(ix, w); } } rel{ix, w, st};
  if ((rc = ensure_base(w)) || (rc = ensure_queries(w, n_q, n_query_tokens, false))) return rc;
  int launches = 0, retries = 0;
  w->real_active = false;
  w->prior_active = false;
  if ((rc = ensure_stage(w)) || (rc = initial_worklists(ix, w, n_q, n_query_tokens))) return rc;
  for (int attempt = 0;; attempt++) {
    if ((rc = enqueue_accept(ix, w, d_q_tokens, d_q_off, n_q, n_query_tokens, pr, capacity, d_block, st, &launches))) return rc;
    if (ix->profiling) cudaEventRecord(w->ev[6], st);
    if ((rc = enqueue_done(w, st))) return rc;
    const int again = wait_and_check(w, attempt, &retries);
    if (again < 0) return -again;
    if (!again) break;
  }
  finish_profile(ix, w, n_q, n_query_tokens, launches, retries);
  return FM_OK;
}

int fm_merge_accepted_device(fm_index* index, int n_shards, const void* const* d_blocks, int64_t total_capacity, const int32_t* d_q_off,
                             int64_t n_q, const fm_params* params, int64_t cap, fm_match* d_out, int32_t* d_out_count,
                             int64_t* need_capacity, void* stream) {
  Index* ix = reinterpret_cast<Index*>(index);
  Params pr;
  int rc;
  if (need_capacity) *need_capacity = 0;
  if (!ix || n_shards < 1 || n_shards > 16 || n_q < 1 || total_capacity < 0 || !d_blocks || cap < 1) { set_error("bad argument"); return FM_ERR_INVALID; }
  if ((rc = check_params(params, &pr))) return rc;
  if (pr.contrast > 0.f) {
    set_error("contrastive rerank needs the sentences behind the records; not supported on accepted-record blocks");
    return FM_ERR_INVALID;
  }
  FM_CUDA(cudaSetDevice(ix->device));
  Workspace* w = acquire(ix);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  struct Releaser { Index* ix; Workspace* w; cudaStream_t st; ~Releaser() { cudaStreamSynchronize(st); release(ix, w); } } rel{ix, w, st};
  if ((rc = ensure_base(w))) return rc;
  int launches = 0;
  const int32_t* blocks[16];
  for (int k = 0; k < n_shards; k++) blocks[k] = static_cast<const int32_t*>(d_blocks[k]);
  if ((rc = enqueue_merge(ix, w, n_shards, blocks, total_capacity, d_q_off, n_q, pr, cap, d_out, d_out_count, st, &launches))) return rc;
  FM_CUDA(cudaStreamSynchronize(st));
  FM_CUDA(cudaGetLastError());
  if (w->h_mctr->overflow & 0x200u) { set_error("the blocks were made for another batch size"); return FM_ERR_INVALID; }
  if (w->h_mctr->overflow & 0x100u) { set_error("a shard block is marked incomplete (workspace overflow in fm_shard_accept_device)"); return FM_ERR_INVALID; }
  if ((w->h_mctr->overflow & 0x400u) && need_capacity) *need_capacity = (int64_t)w->h_mctr->wire_need;
  if ((w->h_mctr->overflow & 0x400u) && !need_capacity) { set_error("a shard accepted more records than its block holds"); return FM_ERR_NOMEM; }
  return FM_OK;
}

// ---- NCCL, opened at run time (the library has no link-time dependency on it)
}  // extern "C"
#include <dlfcn.h>
namespace fm {
struct NcclId { char internal[128]; };
struct Nccl {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void** comm, int nranks, NcclId id, int rank) = nullptr;
  int (*CommDestroy)(void* comm) = nullptr;
  int (*AllGather)(const void* send, void* recv, size_t count, int dtype, void* comm, cudaStream_t st) = nullptr;
  int (*AllReduce)(const void* send, void* recv, size_t count, int dtype, int op, void* comm, cudaStream_t st) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
static Nccl& nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {getenv("FM_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      if (!nm) continue;
      n.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (n.lib) break;
    }
    if (!n.lib) return;
    n.GetUniqueId = reinterpret_cast<int (*)(NcclId*)>(dlsym(n.lib, "ncclGetUniqueId"));
    n.CommInitRank = reinterpret_cast<int (*)(void**, int, NcclId, int)>(dlsym(n.lib, "ncclCommInitRank"));
    n.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(n.lib, "ncclCommDestroy"));
    n.AllGather = reinterpret_cast<int (*)(const void*, void*, size_t, int, void*, cudaStream_t)>(dlsym(n.lib, "ncclAllGather"));
    n.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)>(dlsym(n.lib, "ncclAllReduce"));
    n.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(n.lib, "ncclGetErrorString"));
    n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllGather && n.AllReduce;
  });
  return n;
}
static int nccl_fail(int r, const char* what) {
  set_error(std::string("NCCL error in ") + what + ": " + (nccl().GetErrorString ? nccl().GetErrorString(r) : "?"));
  return FM_ERR_CUDA;
}
}  // namespace fm

struct fm_comm {
  void* comm = nullptr;
  int rank = 0, world = 1, device = 0;
  double rate = 0.5;  // accepted records per query of the fullest shard (recent batches): sizes the blocks
  int64_t last_capacity = 0;
  int64_t last_gather_bytes = 0;
};

extern "C" {

int fm_comm_unique_id(void* id_out) {
  if (!id_out) { set_error("NULL argument"); return FM_ERR_INVALID; }
  if (!nccl().ok) { set_error("libnccl.so.2 not found (set FM_NCCL_LIB)"); return FM_ERR_INVALID; }
  NcclId id;
  const int r = nccl().GetUniqueId(&id);
  if (r) return nccl_fail(r, "ncclGetUniqueId");
  memcpy(id_out, &id, sizeof(id));
  return FM_OK;
}
int fm_comm_create(const void* id, int rank, int world, int device, fm_comm** out) {
  if (out) *out = nullptr;
  if (!id || !out || world < 1 || world > 16 || rank < 0 || rank >= world) { set_error("bad argument (1..16 ranks)"); return FM_ERR_INVALID; }
  if (!nccl().ok) { set_error("libnccl.so.2 not found (set FM_NCCL_LIB)"); return FM_ERR_INVALID; }
  FM_CUDA(cudaSetDevice(device));
  NcclId nid;
  memcpy(&nid, id, sizeof(nid));
  fm_comm* c = new fm_comm();
  c->rank = rank; c->world = world; c->device = device;
  const int r = nccl().CommInitRank(&c->comm, world, nid, rank);
  if (r) { delete c; return nccl_fail(r, "ncclCommInitRank"); }
  *out = c;
  return FM_OK;
}
void fm_comm_destroy(fm_comm* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->comm) nccl().CommDestroy(c->comm);
  delete c;
}
int64_t fm_comm_last_gather_bytes(const fm_comm* c) { return c ? c->last_gather_bytes : 0; }
int64_t fm_comm_block_capacity(const fm_comm* c) { return c ? c->last_capacity : 0; }

}  // extern "C"

// One attempt of a sharded batch on the ticket's stream: shard half -> all-gather -> cross-shard half.
static int enqueue_sharded(fm_ticket* t) {
  Index* ix = t->ix;
  fm_comm* c = t->comm;
  DeviceJob& j = t->dev;
  Workspace* w = j.w;
  int rc;
  const int64_t capacity = std::max<int64_t>(1024, ((int64_t)(c->rate * 1.25 * (double)j.n_q) + 1023) / 1024 * 1024);
  const int64_t bytes = wire_block_bytes(j.n_q, capacity);
  if (bytes > w->cap_wire_block) {
    cudaFree(w->wire_send); cudaFree(w->wire_recv);
    w->wire_send = w->wire_recv = nullptr;
    w->cap_wire_block = 0;
    const int64_t cb = bytes + bytes / 4;
    FM_CUDA(cudaMalloc((void**)&w->wire_send, cb));
    FM_CUDA(cudaMalloc((void**)&w->wire_recv, cb * c->world));
    w->cap_wire_block = cb;
  }
  t->capacity = c->last_capacity = capacity;
  if ((rc = enqueue_accept(ix, w, j.d_q_tok, j.d_q_off, j.n_q, j.n_tok, t->pr, capacity, w->wire_send, j.st, &j.launches))) return rc;
  if (ix->profiling) cudaEventRecord(w->ev[6], j.st);
  const int nr = nccl().AllGather(w->wire_send, w->wire_recv, (size_t)bytes, /*ncclInt8*/ 0, c->comm, j.st);
  if (nr) return nccl_fail(nr, "ncclAllGather");
  c->last_gather_bytes = bytes * c->world;
  const int32_t* blocks[16];
  for (int k = 0; k < c->world; k++) blocks[k] = reinterpret_cast<const int32_t*>(w->wire_recv + (size_t)k * bytes);
  if ((rc = enqueue_merge(ix, w, c->world, blocks, capacity * c->world, j.d_q_off, j.n_q, t->pr, j.cap, j.d_out, j.d_out_count, j.st, &j.launches,
                          /*defer_contrast=*/true)))
    return rc;
  return enqueue_done(w, j.st);
}

// Contrastive rerank of a settled sharded batch (src/fuzzy_match.cc:613-669): the merged, accepted records are on every
// rank; the sentences behind them are gathered into one token slab (each rank fills what it owns, one all-reduce sums the
// slabs) and every rank runs the rerank on it. Collective: all ranks see the same accepted lists, hence the same layout.
static int contrast_sharded(fm_ticket* t) {
  Index* ix = t->ix;
  fm_comm* c = t->comm;
  DeviceJob& j = t->dev;
  Workspace* w = j.w;
  cudaStream_t st = j.st;
  int rc;
  {
    std::lock_guard<std::mutex> g(ix->mu);
    if (!ix->d_sent_start) {
      FM_CUDA(cudaMalloc((void**)&ix->d_sent_start, (size_t)(ix->n_sent + 1) * sizeof(int32_t)));
      FM_CUDA(cudaMemcpy(ix->d_sent_start, ix->h_sent_start.data(), (size_t)(ix->n_sent + 1) * sizeof(int32_t), cudaMemcpyHostToDevice));
      ix->dev.sent_start = ix->d_sent_start;
    }
  }
  if (j.n_q + 1 > w->cap_cq) {
    if ((rc = dev_realloc(&w->c_cnt, j.n_q + 1 + 256)) || (rc = dev_realloc(&w->c_base, j.n_q + 1 + 256))) return rc;
    w->cap_cq = j.n_q + 1 + 256;
  }
  if (!w->h_ctotal) FM_CUDA(cudaMallocHost((void**)&w->h_ctotal, sizeof(int32_t)));
  launch_contrast_need(w->mrec, w->m_base, w->m_idx, w->m_acc, (int32_t)j.n_q, w->c_cnt, st);
  launch_scan(w->c_cnt, w->c_base, (int32_t)j.n_q, w->scan_chain, ++w->scan_epoch, ix->sm_count, st);
  FM_CUDA(cudaMemcpyAsync(w->h_ctotal, w->c_base + j.n_q, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  FM_CUDA(cudaStreamSynchronize(st));
  const int64_t total = *w->h_ctotal;
  if (total < 0) { set_error("contrastive rerank: more than 2^31 tokens behind the accepted records of one batch (split it)"); return FM_ERR_NOMEM; }
  if (total + 1 > w->cap_ctok) {
    if ((rc = dev_realloc(&w->ctok, total + total / 4 + 1024))) return rc;
    w->cap_ctok = total + total / 4 + 1024;
  }
  FM_CUDA(cudaMemsetAsync(w->ctok, 0, (size_t)(total + 1) * sizeof(int32_t), st));
  launch_contrast_fill(ix->dev, ix->n_sent, w->mrec, w->m_base, w->m_idx, w->m_acc, w->c_base, (int32_t)j.n_q, w->ctok, st);
  if (total) {
    const int nr = nccl().AllReduce(w->ctok, w->ctok, (size_t)total, /*ncclInt32*/ 2, /*ncclSum*/ 0, c->comm, st);
    if (nr) return nccl_fail(nr, "ncclAllReduce");
  }
  IndexDev slab = ix->dev;
  slab.tok = w->ctok;
  launch_contrast(slab, w->mrec, w->m_base, w->m_idx, w->m_acc, (int32_t)j.n_q, t->pr, j.cap, j.d_out, j.d_out_count, w->mctr, ix->sm_count, st);
  j.launches += 4;
  FM_CUDA(cudaStreamSynchronize(st));
  FM_CUDA(cudaGetLastError());
  return FM_OK;
}

// Every rank takes the same decisions from the same gathered data, so the collectives stay matched: a batch
// is rerun by all ranks when any shard overflowed its workspace (flag in its block header) or accepted more
// records than a block holds (its total travels in the header; the next size follows the largest).
static int finish_sharded(fm_ticket* t) {
  Index* ix = t->ix;
  fm_comm* c = t->comm;
  DeviceJob& j = t->dev;
  Workspace* w = j.w;
  int retries = 0, rc;
  std::lock_guard<std::mutex> coll(ix->shard_mu);
  for (;;) {
    const int mine = wait_and_check(w, t->attempts, &retries);  // own overflow: regrown here
    if (mine < 0) return -mine;
    const unsigned flags = w->h_mctr->overflow;
    if (flags & 0x200u) { set_error("ranks disagree on the batch (number of queries)"); return FM_ERR_INVALID; }
    const double seen = (double)w->h_mctr->wire_need / (double)j.n_q;
    if (!(flags & 0x100u)) c->rate = std::max(c->rate * 0.98, seen);  // (totals of an overflowed pipeline mean nothing)
    if (!(flags & (0x100u | 0x400u))) break;
    if (++t->attempts >= 10) { set_error("sharded batch does not settle (workspace / block size keep growing)"); return FM_ERR_NOMEM; }
    if ((rc = enqueue_sharded(t))) return rc;
  }
  if (t->pr.contrast > 0.f && (rc = contrast_sharded(t))) return rc;
  finish_profile(ix, w, j.n_q, j.n_tok, j.launches, retries);
  return FM_OK;
}

extern "C" {

int fm_match_batch_sharded_submit(fm_index* index, fm_comm* c, const int32_t* d_q_tokens, const int32_t* d_q_off, int64_t n_q,
                                  int64_t n_query_tokens, const fm_params* params, int64_t cap, fm_match* d_out, int32_t* d_out_count,
                                  void* stream, fm_ticket** ticket) {
  Index* ix = reinterpret_cast<Index*>(index);
  Params pr;
  int rc;
  if (ticket) *ticket = nullptr;
  if (!ix || !c || !ticket || n_q < 1 || n_q > (1 << 20) || cap < 1) { set_error("bad argument"); return FM_ERR_INVALID; }
  if (ix->device != c->device) { set_error("index and communicator live on different devices"); return FM_ERR_INVALID; }
  if ((rc = check_params(params, &pr))) return rc;
  if (c->world == 1)  // one shard is the whole TM
    return fm_match_batch_device_submit(index, d_q_tokens, d_q_off, n_q, n_query_tokens, params, cap, d_out, d_out_count, stream, ticket);
  FM_CUDA(cudaSetDevice(ix->device));
  std::lock_guard<std::mutex> coll(ix->shard_mu);  // collectives of one communicator are issued one call at a time
  fm_ticket* t = new fm_ticket();
  t->ix = ix; t->pr = pr; t->cap = cap; t->host = false; t->comm = c;
  DeviceJob& j = t->dev;
  j.w = acquire(ix);
  j.d_q_tok = d_q_tokens; j.d_q_off = d_q_off; j.n_q = n_q; j.n_tok = n_query_tokens; j.cap = cap;
  j.d_out = d_out; j.d_out_count = d_out_count; j.st = static_cast<cudaStream_t>(stream);
  j.w->real_active = false;
  j.w->prior_active = false;
  if ((rc = ensure_base(j.w)) || (rc = ensure_queries(j.w, n_q, n_query_tokens, false)) || (rc = ensure_stage(j.w)) ||
      (rc = initial_worklists(ix, j.w, n_q, n_query_tokens)) || (rc = enqueue_sharded(t))) {
    drop_ticket(t);
    return rc;
  }
  *ticket = t;
  return FM_OK;
}

int fm_match_batch_sharded_device(fm_index* index, fm_comm* c, const int32_t* d_q_tokens, const int32_t* d_q_off, int64_t n_q,
                                  int64_t n_query_tokens, const fm_params* params, int64_t cap, fm_match* d_out, int32_t* d_out_count,
                                  void* stream) {
  fm_ticket* t = nullptr;
  const int rc = fm_match_batch_sharded_submit(index, c, d_q_tokens, d_q_off, n_q, n_query_tokens, params, cap, d_out, d_out_count, stream, &t);
  return rc ? rc : fm_ticket_wait(t);
}

int fm_set_profiling(fm_index* index, int enabled) {
  Index* ix = reinterpret_cast<Index*>(index);
  if (!ix) { set_error("NULL index"); return FM_ERR_INVALID; }
  ix->profiling = enabled != 0;
  return FM_OK;
}
int fm_get_profile(const fm_index* index, fm_profile* out) {
  Index* ix = const_cast<Index*>(reinterpret_cast<const Index*>(index));
  if (!ix || !out) { set_error("NULL argument"); return FM_ERR_INVALID; }
  std::lock_guard<std::mutex> g(ix->mu);
  *out = ix->last_profile;
  return FM_OK;
}

}  // extern "C"

// Debug hooks (not part of the public header, used while bringing the kernels up and kept for
// stage-level inspection): fm_debug_last_slices copies the range slices of the last batch run on the
// first workspace, rec = int4 (query, sa_begin, match_len | p << 10 | mult << 20, size) per slice (the flattened
// ones first, then the small ones with start = -1);
// fm_debug_last_survivors copies its (query, sentence start, table slot, arrival index) records and the
// max match length recorded for each.
extern "C" int64_t fm_debug_last_slices(fm_index* index, int32_t* rec, int64_t* start, int64_t cap) {
  Index* ix = reinterpret_cast<Index*>(index);
  if (!ix || ix->pool.empty()) return -1;
  Workspace* w = ix->pool[0];
  cudaSetDevice(ix->device);
  cudaDeviceSynchronize();
  const int64_t n_big = std::min<int64_t>(cap, (int64_t)(w->h_ctr->slice_elem >> kElemBits));
  const int64_t n_small = std::min<int64_t>(cap - n_big, (int64_t)w->h_ctr->n_small);
  cudaMemcpy2D(rec, sizeof(int4), w->sl_rec, 2 * sizeof(int4), sizeof(int4), n_big, cudaMemcpyDeviceToHost);  // first half of each record
  cudaMemcpy(start, w->sl_start, n_big * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaMemcpy2D(rec + 4 * n_big, sizeof(int4), w->sm_rec, 2 * sizeof(int4), sizeof(int4), n_small, cudaMemcpyDeviceToHost);  // (no flattened start)
  for (int64_t i = 0; i < n_small; i++) start[n_big + i] = -1;
  return n_big + n_small;
}
extern "C" int64_t fm_debug_last_survivors(fm_index* index, int32_t* surv, uint32_t* lm, int64_t cap) {
  Index* ix = reinterpret_cast<Index*>(index);
  if (!ix || ix->pool.empty()) return -1;
  Workspace* w = ix->pool[0];
  cudaSetDevice(ix->device);
  cudaDeviceSynchronize();
  const int64_t n = std::min<int64_t>(cap, (int64_t)w->h_ctr->n_surv);
  cudaMemcpy(surv, w->surv, n * sizeof(SurvRec), cudaMemcpyDeviceToHost);
  std::vector<uint32_t> hl(w->hsize);
  cudaMemcpy(hl.data(), w->hlm, (size_t)w->hsize * 4, cudaMemcpyDeviceToHost);
  for (int64_t i = 0; i < n; i++) lm[i] = hl[surv[4 * i + 2]];
  return n;
}
