#!/usr/bin/env python
"""bench.py -- fuzzy-match queries/sec on the BASELINE.json headline workload.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--no-configs] [--no-config3] [--no-cpu-baseline] [--no-sustained]

A step = one pass of the hot path (FuzzyMatch::match for every pattern of one batch) over one batch
of --queries synthetic patterns against the synthetic 1M-sentence TM (BASELINE.json configs[1]:
1M sentences avg 15 tokens, f=0.7, n=1, ml=3, mr=0, unit edit costs).

  value      queries/s with the batches already resident in HBM (fm_match_batch_device_submit /
             fm_ticket_wait, three batches in flight, each on its own stream so that their kernels overlap),
             CUDA events around the whole region (the first stream's start event gates the other stream, its
             end event waits for it), max over ranks.
  e2e        the same through the host-buffer calls fm_match_batch_submit / fm_ticket_wait: pinned host
             buffers in, host results out, every step's H2D and D2H copies inside the timed region.
  sustained  `value` again over a region of >= 2 s with its own clock record.
  roofline   the dominant kernel's algorithmic bytes / its CUDA-event time vs the measured HBM peak, the
             same with its ncu DRAM traffic (dram_frac), and the DP kernel's cell updates/s vs the ALU ceiling.
  configs    (N=1) the other BASELINE.json configurations on one GPU: CLI defaults, configs[2] parameters,
             configs[3] long patterns, configs[4] contrastive + idf -- q/s, e2e, stage times, and whether a
             sample of the results equals the CPU oracle's.
  config3    BASELINE.json configs[2] (10M-sentence TM, f=0.5, ml=3): the SAME query batches at every N --
             N=1 on the unsharded index; N>1 on N sentence-id shards with one NCCL all-gather per batch
             (north_star layout, strong scaling) and, beside it, on N replicas of the unsharded index that
             split each batch by query.
  cpu_baseline / --impl reference   the reference's own CPU implementation (oracle/_ref, built from the
             reference sources; falls back to the C restatement) on this box's host cores.

N > 1 (torchrun, one rank per GPU): the headline replicates the index and gives each rank its own
batches -- no data-path collective, weak scaling; the north-star partition is the `config3` leg.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PARAMS = dict(fuzzy=0.7, n=1, ml=3, mr=0.0)
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md
# 148 SMs x 128 lanes x ~1.9 GHz = 3.6e13 lane-ops/s, ~10 ops per cell update
DP_ALU_CEILING_GCUPS = 3600.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sentences", type=int, default=1000000)
    ap.add_argument("--queries", type=int, default=100000)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the baseline sample")
    ap.add_argument("--config3-sentences", type=int, default=10000000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--no-config3", action="store_true")
    ap.add_argument("--no-sustained", action="store_true")
    return ap.parse_args()


def measured_peak():
    """HBM peak for the roofline: hbm_gbs of the driver-written MEASURED_PEAKS.json, else the recipe's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during a timed region."""

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "power_w_max": None, "samples": 0, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 8:
                    continue
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                try:
                    pw.append(float(f[2]))
                except ValueError:
                    pass
                for name, v in zip(names, f[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
        if pw:
            out["power_w_max"] = float(max(pw))
        out["reasons"] = sorted(reasons)
        return out


def workload(args, rank=0, n_batches=4, shard_queries=False):
    from fuzzy_match_b200 import synth
    tm, off, V = synth.make_tm(args.sentences, seed=1234)
    batches = []
    for b in range(n_batches):
        seed = 5678 + b + (1000 * rank if shard_queries else 0)
        batches.append(synth.make_queries(tm, off, args.queries, seed=seed))
    return tm, off, V, batches


def workload_config(args):
    """Identical in both arms (the driver compares the two dicts)."""
    return {"workload": "BASELINE.json configs[1]: %d-sentence synthetic TM (Zipf(1) over 50k words, len U[5,25]), "
                        "%d queries per step (80%% perturbed TM sentences, 20%% random), f=0.7, n=1, ml=3, mr=0, unit costs"
                        % (args.sentences, args.queries),
            "sentences": args.sentences, "queries_per_step": args.queries, "fuzzy": 0.7, "number_of_matches": 1,
            "min_subseq_length": 3, "min_subseq_ratio": 0.0,
            "cache": "inputs larger than L2: the index arrays (> 1 GB at 1M sentences) exceed the 126 MB L2 and successive "
                     "steps use different query batches"}


# ------------------------------------------------------------------------------------------ CPU arm


def cpu_reference_run(tm, off, V, q, qo, seconds, threads, params=PARAMS, cap=1):
    """Times the reference's own CPU implementation on a bounded sample of the batch."""
    from oracle import binding as ob
    if not (os.path.exists(ob.ORACLE_SO) and (os.path.exists(ob.REF_SO) or not os.path.isdir("/root/reference"))):
        ob.build()
    kind = "reference" if ob.ref_available() else "port"
    t0 = time.time()
    idx = ob.RefIndex(tm, off) if kind == "reference" else ob.OracleIndex(tm, off, V)
    build_s = time.time() - t0
    n_q = len(qo) - 1
    probe = min(n_q, 2000)

    def run(n):
        t = time.time()
        idx.match_batch(q[:qo[n]], qo[:n + 1], cap=cap, nthreads=threads, **params)
        return idx.last_seconds if kind == "reference" else time.time() - t

    dt = run(probe)  # warm-up + rate estimate
    n = int(min(n_q, max(probe, probe / max(dt, 1e-6) * seconds)))
    dt = run(n)
    return dict(kind=kind, n=n, seconds=dt, qps=n / dt, build_s=build_s, index=idx)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tm, off, V, batches = workload(args, n_batches=1)
    q, qo = batches[0]
    threads = os.cpu_count() or 1
    first = cpu_reference_run(tm, off, V, q, qo, args.cpu_seconds / max(1, args.steps), threads)
    idx, kind, n = first["index"], first["kind"], first["n"]
    times = []
    for _ in range(args.warmup + args.steps):
        t = time.time()
        idx.match_batch(q[:qo[n]], qo[:n + 1], cap=1, nthreads=threads, **PARAMS)
        times.append(idx.last_seconds if kind == "reference" else time.time() - t)
    times = times[args.warmup:]
    total = sum(times)
    value = n * len(times) / total
    line = {
        "impl": "reference", "metric": "fuzzy-match queries/sec @ 1M-sent TM f=0.7", "value": value, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "run": {"sample_queries_per_step": n, "index_build_s": round(first["build_s"], 2)},
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": threads, "kind": kind,
                         "sample": "%d of the %d queries of one batch per step, %d host threads sharing one index" % (n, len(qo) - 1, threads)},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------ GPU arm


_REAL_STDOUT = None


def _capture_stdout():
    """Everything libraries print on fd 1 (e.g. the NCCL version banner) goes to stderr; only the JSON
    line is written to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


class Runner:
    """One index + a set of query batches: device-resident and host-buffer loops with `depth` batches in flight."""

    def __init__(self, index, batches, params, cap, dev, torch, capi, depth=int(os.environ.get("FM_BENCH_DEPTH", "3"))):
        self.index, self.params, self.cap, self.dev, self.torch, self.capi, self.depth = index, params, cap, dev, torch, capi, depth
        self.streams = [torch.cuda.Stream(dev) for _ in range(depth)]  # one per batch in flight: their kernels may overlap
        self.stream = self.streams[0]
        self.batches = batches
        self.dbatches = [(torch.as_tensor(q, device=dev), torch.as_tensor(qo.astype(np.int32), device=dev), len(qo) - 1, int(qo[-1]))
                         for q, qo in batches]
        nq_max = max(len(qo) - 1 for _, qo in batches)
        self.d_out = [torch.zeros(nq_max * cap * 24, dtype=torch.uint8, device=dev) for _ in range(depth)]
        self.d_cnt = [torch.zeros(nq_max, dtype=torch.int32, device=dev) for _ in range(depth)]
        self.pinned = None

    def device_loop(self, first, steps):
        """Enqueue `steps` batches on the stream, at most `depth` unfinished tickets; returns after the last wait."""
        tickets = []
        for i in range(steps):
            dq, dqo, nq, ntok = self.dbatches[(first + i) % len(self.dbatches)]
            s = i % self.depth
            if len(tickets) >= self.depth:
                self.index.wait(tickets.pop(0))
            tickets.append(self.index.submit_device(dq.data_ptr(), dqo.data_ptr(), nq, ntok, self.d_out[s].data_ptr(),
                                                    self.d_cnt[s].data_ptr(), self.cap, self.streams[s].cuda_stream, self.params))
        for t in tickets:
            self.index.wait(t)

    def time_device(self, first, steps):
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(self.dev)
        e0.record(self.streams[0])
        for st in self.streams[1:]:  # the other streams start behind the first event ...
            st.wait_event(e0)
        self.device_loop(first, steps)
        for st in self.streams[1:]:  # ... and the last event waits for all of them
            self.streams[0].wait_stream(st)
        e1.record(self.streams[0])
        torch.cuda.synchronize(self.dev)
        return e0.elapsed_time(e1)

    def pin(self):
        torch, capi = self.torch, self.capi
        self.pinned = []
        for q, qo in self.batches:
            pq = torch.empty(len(q), dtype=torch.int32).pin_memory().numpy()
            pq[:] = q
            po = torch.empty(len(qo), dtype=torch.int64).pin_memory().numpy()
            po[:] = qo
            self.pinned.append((pq, po))
        nq_max = max(len(qo) - 1 for _, qo in self.batches)
        self.pout = [torch.empty(nq_max * self.cap * 24, dtype=torch.uint8).pin_memory().numpy().view(capi.MATCH_DTYPE).reshape(nq_max, self.cap)
                     for _ in range(self.depth + 1)]
        self.pcnt = [torch.empty(nq_max, dtype=torch.int32).pin_memory().numpy() for _ in range(self.depth + 1)]

    def host_loop(self, first, steps):
        """The reference-facing path: host CSR in, host fm_match[] out, `depth` batches in flight."""
        if self.pinned is None:
            self.pin()
        tickets = []
        for i in range(steps):
            pq, po = self.pinned[(first + i) % len(self.pinned)]
            s = i % (self.depth + 1)
            if len(tickets) >= self.depth:
                self.index.wait(tickets.pop(0))
            nq = len(po) - 1
            tickets.append(self.index.submit(pq, po, self.pout[s][:nq], self.pcnt[s][:nq], self.cap, self.params))
        for t in tickets:
            self.index.wait(t)

    def time_host(self, first, steps):
        self.torch.cuda.synchronize(self.dev)
        t = time.perf_counter()
        self.host_loop(first, steps)
        self.torch.cuda.synchronize(self.dev)
        return 1e3 * (time.perf_counter() - t)

    def bytes_per_step(self):
        h2d = int(np.mean([4 * len(q) + 4 * len(qo) for q, qo in self.batches]))
        nq = int(np.mean([len(qo) - 1 for _, qo in self.batches]))
        return h2d, nq * self.cap * 24 + nq * 4 + 32

    def profile(self, reps):
        """Per-stage CUDA-event times inside the library (synchronous calls, one batch at a time)."""
        self.index.set_profiling(True)
        acc = {}
        for i in range(reps):
            dq, dqo, nq, ntok = self.dbatches[i % len(self.dbatches)]
            self.index.match_batch_device(dq.data_ptr(), dqo.data_ptr(), nq, ntok, self.d_out[0].data_ptr(), self.d_cnt[0].data_ptr(),
                                          self.cap, stream=self.stream.cuda_stream, params=self.params)
            self.torch.cuda.synchronize(self.dev)
            for k, v in self.index.profile().items():
                acc[k] = acc.get(k, 0) + v
        self.index.set_profiling(False)
        return {k: v / reps for k, v in acc.items()}

    def results(self, b=0):
        """Synchronous host call on batch b (for the parity samples)."""
        q, qo = self.batches[b]
        return self.index.match_batch(q, qo, cap=self.cap, params=self.params)


def parity_sample(tm, off, V, q, qo, out, cnt, n_sample, cap, oracle_cache, **params):
    """First n queries of a batch against the CPU oracle (checker only; outside every timed region)."""
    from oracle import binding as ob
    key = id(tm)
    if key not in oracle_cache:
        if not os.path.exists(ob.ORACLE_SO):
            ob.build()
        oracle_cache[key] = ob.OracleIndex(tm, off, V)
    n = min(n_sample, len(qo) - 1)
    ro, oc = oracle_cache[key].match_batch(q[:qo[n]], qo[:n + 1], cap=cap, nthreads=os.cpu_count() or 1, **params)
    return bool((cnt[:n] == oc).all() and all(out[i, :min(cnt[i], cap)].tobytes() == ro[i].tobytes() for i in range(n)))


def oracle_counters(tm, off, V, q, qo, n_sample, oracle_cache, **params):
    """Implementation-independent work counters from the C restatement on a sample."""
    from oracle import binding as ob
    key = id(tm)
    if key not in oracle_cache:
        if not os.path.exists(ob.ORACLE_SO):
            ob.build()
        oracle_cache[key] = ob.OracleIndex(tm, off, V)
    n = min(n_sample, len(qo) - 1)
    _, _, ct = oracle_cache[key].match_batch(q[:qo[n]], qo[:n + 1], cap=1, nthreads=os.cpu_count() or 1, counters=True, **params)
    return {k: v / n for k, v in ct.items()}


def config_entry(name, desc, index, tm, off, V, batches, kw, cap, dev, torch, capi, oracle_cache, steps, sample):
    """One secondary configuration on one GPU: device-resident q/s, e2e q/s, stage times, parity sample."""
    params = capi.Params.make(**kw)
    r = Runner(index, batches, params, cap, dev, torch, capi)
    r.device_loop(0, max(3, len(batches)))
    ms = r.time_device(0, steps)
    r.host_loop(0, 3)
    ems = r.time_host(0, steps)
    prof = r.profile(3)
    out, cnt = r.results(0)
    q, qo = batches[0]
    ok = parity_sample(tm, off, V, q, qo, out, cnt, sample, cap, oracle_cache, **kw)
    nq = len(qo) - 1
    ct = oracle_counters(tm, off, V, q, qo, min(sample, 500), oracle_cache, **kw)
    h2d, d2h = r.bytes_per_step()
    e = {"workload": desc, "queries_per_step": nq, "value": nq * steps / (ms / 1e3), "unit": "queries/s", "ms_per_step": ms / steps,
         "e2e": {"value": nq * steps / (ems / 1e3), "unit": "queries/s", "ms_per_step": ems / steps, "h2d_bytes_per_step": h2d,
                 "d2h_bytes_per_step": d2h},
         "stage_ms": {k[3:]: round(prof[k], 4) for k in ("ms_prepare", "ms_search", "ms_walk", "ms_verify", "ms_scan", "ms_score", "ms_replay")},
         "found_fraction": float((cnt > 0).mean()), "parity_sample": {"queries": min(sample, nq), "identical_to_oracle": ok},
         "dp": dp_entry(ct, nq, prof["ms_score"])}
    log(name, json.dumps(e))
    return e


def dp_entry(counters, n_q, ms_score):
    """Cell updates of the reference's DP (edit_distance.cc:41-75, counted by the oracle) per second of the score stage."""
    cells = counters["dp_cells"] * n_q
    g = cells / (ms_score * 1e-3) / 1e9 if ms_score > 0 else None
    return {"cells_per_step": int(cells), "score_ms": round(ms_score, 4), "gcups": g, "alu_ceiling_gcups": DP_ALU_CEILING_GCUPS,
            "frac": (g / DP_ALU_CEILING_GCUPS) if g else None,
            "note": "equal-cost pairs run the bit-parallel kernel (32-64 cells per ~17 integer instructions), so cell "
                    "updates/s can exceed the scalar-ALU ceiling of ~10 ops per cell"}


def config3_leg(args, world, rank, dev, torch, dist, capi, fmb):
    """BASELINE.json configs[2]: the same query batches at every N, TM-sharded vs query-split replicas."""
    from fuzzy_match_b200 import synth
    t0 = time.time()
    tm, off, V = synth.make_tm(args.config3_sentences, seed=1234)
    n_batches = 3
    batches = [synth.make_queries(tm, off, args.queries, seed=5678 + b) for b in range(n_batches)]
    gen_s = time.time() - t0
    kw = dict(fuzzy=0.5, n=1, ml=3)
    params = capi.Params.make(**kw)
    steps = max(3, min(args.steps, 12))
    out = {"workload": "BASELINE.json configs[2]: %d-sentence synthetic TM, the same %d batches of %d queries at every N, f=0.5, ml=3, n=1"
                       % (args.config3_sentences, n_batches, args.queries),
           "steps": steps, "tm_generate_s": round(gen_s, 1)}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_ms(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # (a) replicas of the unsharded index; every rank takes its 1/N of each batch
    t0 = time.time()
    full = fmb.Index(tm, off, V, device=dev.index)
    build_full = time.time() - t0
    mine = []
    for q, qo in batches:
        nq = len(qo) - 1
        a, b = rank * nq // world, (rank + 1) * nq // world
        mine.append((np.ascontiguousarray(q[qo[a]:qo[b]]), np.ascontiguousarray(qo[a:b + 1] - qo[a])))
    r = Runner(full, mine, params, 1, dev, torch, capi)
    r.device_loop(0, 3)
    barrier()
    ms = max_ms(r.time_device(0, steps))
    barrier()
    rep = {"value": args.queries * steps / (ms / 1e3), "unit": "queries/s", "ms_per_step": ms / steps,
           "index_device_bytes": int(full.device_bytes), "index_build_s": round(build_full, 2),
           "layout": "1 GPU, unsharded index" if world == 1 else
                     "%d replicas of the unsharded index, each batch split by query (no collective)" % world}
    out["replicas" if world > 1 else "unsharded"] = rep
    found = None
    if world == 1:
        o, c = r.results(0)
        found = float((c > 0).mean())
        rep["found_fraction"] = found
    del r
    full.close()
    # (b) north-star layout: N sentence-id shards, one NCCL all-gather of accepted records per batch
    if world > 1:
        from fuzzy_match_b200.sharded import ShardedIndex
        t0 = time.time()
        sidx = ShardedIndex(tm, off, V, device=dev)
        barrier()
        build_sh = time.time() - t0
        depth = int(os.environ.get("FM_BENCH_DEPTH", "3"))
        streams = [torch.cuda.Stream(dev) for _ in range(depth)]
        dbat = [(torch.as_tensor(q, device=dev), torch.as_tensor(qo.astype(np.int32), device=dev), len(qo) - 1, int(qo[-1])) for q, qo in batches]
        d_out = [torch.zeros(args.queries * 24, dtype=torch.uint8, device=dev) for _ in range(depth)]
        d_cnt = [torch.zeros(args.queries, dtype=torch.int32, device=dev) for _ in range(depth)]

        def sharded_loop(n):  # `depth` batches in flight, each on its own stream; every rank in the same order
            tickets = []
            for i in range(n):
                dq, dqo, nq, ntok = dbat[i % n_batches]
                s = i % depth
                if len(tickets) >= depth:
                    sidx.wait(tickets.pop(0))
                tickets.append(sidx.submit_device(dq, dqo, nq, ntok, d_out[s], d_cnt[s], 1, params, streams[s]))
            for t in tickets:
                sidx.wait(t)

        sharded_loop(max(3, depth))
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(streams[0])
        for st in streams[1:]:
            st.wait_event(e0)
        sharded_loop(steps)
        for st in streams[1:]:
            streams[0].wait_stream(st)
        e1.record(streams[0])
        barrier()
        ms = max_ms(e0.elapsed_time(e1))
        last = (steps - 1) % depth
        out["tm_sharded"] = {"value": args.queries * steps / (ms / 1e3), "unit": "queries/s", "ms_per_step": ms / steps,
                             "scaling": "strong", "shard_device_bytes": int(sidx.index.device_bytes), "shard_build_s": round(build_sh, 2),
                             "allgather_bytes_per_step": int(sidx.last_gather_bytes),
                             "found_fraction": float((d_cnt[last][:dbat[(steps - 1) % n_batches][2]] > 0).float().mean().item()),
                             "batches_in_flight": depth, "block_capacity_records": int(sidx.block_capacity),
                             "layout": "%d sentence-id shards, queries replicated, local replay per shard, one NCCL all-gather of the "
                                       "locally accepted records per batch, merged replay" % world}
    return out


def main():
    args = parse()
    _capture_stdout()
    if args.impl == "reference":
        reference_arm(args)
        return
    import torch
    import torch.distributed as dist

    import fuzzy_match_b200 as fmb
    from fuzzy_match_b200 import capi
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if not os.path.exists(fmb.library_path()):
        raise SystemExit("libfm_b200.so missing: run __graft_entry__.build() (there is no CPU fallback)")

    n_batches = min(4, args.steps + args.warmup)
    tm, off, V, batches = workload(args, rank=rank, n_batches=n_batches, shard_queries=world > 1)
    params = capi.Params.make(**PARAMS)
    cap = 1
    t0 = time.time()
    index = fmb.Index(tm, off, V, device=local_rank)
    build_s = time.time() - t0
    run = Runner(index, batches, params, cap, dev, torch, capi)
    n_q = args.queries

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- value: device-resident, CUDA events on the launching stream
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(1.5)  # let nvidia-smi finish attaching before anything is timed
    run.device_loop(0, max(args.warmup, n_batches))  # every distinct batch once, so no workspace growth is timed
    barrier()
    dev_ms = run.time_device(args.warmup, args.steps)
    barrier()
    found = int((run.d_cnt[(args.steps - 1) % run.depth][:n_q] > 0).sum().item())

    # ---- e2e: host buffers through fm_match_batch_submit / fm_ticket_wait (pinned), wall clock
    run.host_loop(0, max(args.warmup, n_batches))
    barrier()
    e2e_ms = run.time_host(args.warmup, args.steps)
    h2d, d2h = run.bytes_per_step()
    clocks = sampler.stop() if rank == 0 else None

    # ---- sustained: the device-resident loop again for >= 2 s, own clock record
    sustained = None
    if not args.no_sustained:
        s2 = ClockSampler(local_rank)
        if rank == 0:
            s2.start()
            time.sleep(0.5)
        n_sus = int(max(args.steps, 2500.0 / max(dev_ms / args.steps, 1e-3)))
        barrier()
        sus_ms = run.time_device(0, n_sus)
        barrier()
        t = torch.tensor([sus_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sustained = {"value": n_q * n_sus * world / (float(t[0]) / 1e3), "unit": "queries/s", "steps": n_sus, "seconds": float(t[0]) / 1e3,
                     "ms_per_step": float(t[0]) / n_sus, "clocks": s2.stop() if rank == 0 else None}

    # ---- per-kernel times (CUDA events inside the library, same stream) for the roofline
    prof = run.profile(max(3, min(args.steps, 10)))

    # ---- max over ranks
    times = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms_max = float(times[0]), float(times[1])
    units = args.queries * args.steps * world
    value = units / (dev_ms / 1e3)

    config3 = None
    if not args.no_config3:
        log("config 3 leg: generating the %d-sentence TM" % args.config3_sentences)
        config3 = config3_leg(args, world, rank, dev, torch, dist, capi, fmb)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    line = {
        "metric": "fuzzy-match queries/sec @ 1M-sent TM f=0.7", "value": value, "unit": "queries/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "run": {"parallelism": "query-sharded replicas x%d" % world if world > 1 else "1 GPU", "batches_in_flight": run.depth,
                "index_build_s": round(build_s, 2), "index_device_bytes": int(index.device_bytes), "found_fraction": found / n_q},
        "gpu_launches": int(prof["launches"] * args.steps),
        "clocks": clocks,
        "e2e": {"value": units / (e2e_ms_max / 1e3), "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms_max / args.steps,
                "api": "fm_match_batch_submit / fm_ticket_wait (host CSR in pinned memory -> host fm_match[]), %d batches in flight" % run.depth},
    }
    if sustained:
        line["sustained"] = sustained
    if config3:
        line["config3"] = config3

    oracle_cache = {}
    counters = None
    if world == 1 and not args.no_cpu_baseline:
        q, qo = batches[0]
        threads = os.cpu_count() or 1
        cpu = cpu_reference_run(tm, off, V, q, qo, args.cpu_seconds, threads)
        line["cpu_baseline"] = {"value": cpu["qps"], "unit": "queries/s", "cores": threads, "kind": cpu["kind"],
                                "sample": "first %d queries of batch 0 in %.1f s, %d host threads on one shared index "
                                          "(index build %.1f s excluded)" % (cpu["n"], cpu["seconds"], threads, cpu["build_s"])}
        del cpu
    if world == 1:
        q, qo = batches[0]
        counters = oracle_counters(tm, off, V, q, qo, 4000, oracle_cache, **PARAMS)
    peak, peak_src = measured_peak()
    # the two kernels of the gather stage are timed separately
    stages = {k[3:]: prof[k] for k in ("ms_prepare", "ms_search", "ms_walk", "ms_verify", "ms_scan", "ms_score", "ms_replay")}
    kernel_of = {"walk": "fm_gather_kernel", "verify": "fm_verify_kernel", "search": "fm_search_kernel", "prepare": "fm_prepare_short_kernel",
                 "scan": "fm_scan_kernel", "score": "fm_score_bp_kernel", "replay": "fm_replay_small_kernel"}
    dom = max(stages, key=stages.get)
    ncu = {}
    try:  # per-launch DRAM bytes / issue-slot utilisation from the committed ncu --set full capture of this workload
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            ncu = json.load(f)
    except Exception:
        pass
    kname = kernel_of[dom]
    traffic = ncu.get(kname, {}).get("traffic_bytes_per_launch")
    roof = {"bound": "hbm", "kernel": kname, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "traffic": traffic,
            "stage_ms": {k: round(v, 4) for k, v in stages.items()}, "flattened_elements_per_step": prof["n_elements"],
            "slices_per_step": prof["n_slices"], "candidates_per_step": prof["n_stage2"], "survivors_per_step": prof["n_survivors"]}
    if counters:
        # algorithmic bytes per query, SURVEY.md 8d: search 16 B/probe; range walk ("suffix-range gather") 8 B per
        # element walked; verify (8 + 4*s) per deduplicated candidate (the coverage fetch); score 4*s per DP pair +
        # 4*p pattern; 16 B per returned match.
        per_q = {"search": 16 * counters["probes"], "walk": 8 * counters["elements_walked"],
                 "verify": 8 * counters["candidates"] + 4 * counters["candidate_tokens"],
                 "score": 4 * counters["dp_tokens"] + 4 * counters["pattern_tokens"],
                 "replay": 16 * counters["matches_out"], "prepare": 8 * counters["pattern_tokens"], "scan": 8.0}
        roof["algorithmic_bytes_per_query"] = {k: round(v, 1) for k, v in per_q.items()}
        kern = {}
        for k in stages:
            e = {"ms": round(stages[k], 4), "algorithmic_gbs": round(per_q[k] * n_q / (stages[k] * 1e-3) / 1e9, 1) if stages[k] > 0 else None}
            if e["algorithmic_gbs"] is not None:
                e["frac"] = round(e["algorithmic_gbs"] / peak, 4)
            m = ncu.get(kernel_of[k])
            if m and stages[k] > 0:  # what the kernel really moves through HBM, and how busy its issue slots are (ncu)
                e["dram_bytes"] = m.get("traffic_bytes_per_launch")
                e["dram_gbs"] = round(m["traffic_bytes_per_launch"] / (stages[k] * 1e-3) / 1e9, 1)
                e["dram_frac"] = round(e["dram_gbs"] / peak, 4)
                e["issue_active_pct"] = m.get("issue_active_pct")
                e["warp_instructions"] = m.get("inst_executed")
            kern[kernel_of[k]] = e
        roof["kernels"] = kern
        ach = per_q[dom] * n_q / (stages[dom] * 1e-3) / 1e9
        roof.update({"achieved": ach, "frac": ach / peak})
        if traffic and stages[dom] > 0:
            roof["dram_achieved"] = traffic / (stages[dom] * 1e-3) / 1e9
            roof["dram_frac"] = roof["dram_achieved"] / peak
            roof["issue_active"] = ncu.get(kname, {}).get("issue_active_pct")
            roof["warp_instructions"] = ncu.get(kname, {}).get("inst_executed")
        total_bytes = sum(per_q[k] for k in ("search", "walk", "verify", "score", "replay"))
        roof["whole_step"] = {"bytes_per_query": round(total_bytes, 1), "achieved": total_bytes * n_q / (dev_ms / args.steps * 1e-3) / 1e9,
                              "frac": total_bytes * n_q / (dev_ms / args.steps * 1e-3) / 1e9 / peak}
        roof["dp"] = dp_entry(counters, n_q, stages["score"])
        roof["note"] = ("achieved = algorithmic bytes (SURVEY.md 8d, counted by the oracle) / CUDA-event time: the signature test "
                        "rejects ~97% of the walked elements without touching their sentences and hot ranges stay in L2, so the "
                        "bytes that really cross HBM (dram_*) are far fewer; frac can exceed 1 for that reason")
    line["roofline"] = roof

    if world == 1 and not args.no_configs:
        from fuzzy_match_b200 import synth
        steps = max(3, min(args.steps, 10))
        cfgs = {}
        cfgs["cli_defaults"] = config_entry("cli_defaults", "configs[1] TM, FuzzyMatch-cli defaults f=0.8 n=5 ml=3 mr=0.3", index, tm, off, V,
                                            batches[:2], dict(fuzzy=0.8, n=5, ml=3, mr=0.3), 5, dev, torch, capi, oracle_cache, steps, 1000)
        cfgs["c3_shape"] = config_entry("c3_shape", "configs[2] parameters (f=0.5, ml=3, n=1) on the 1M-sentence TM", index, tm, off, V,
                                        batches[:2], dict(fuzzy=0.5, n=1, ml=3), 1, dev, torch, capi, oracle_cache, steps, 1000)
        cfgs["c5"] = config_entry("c5", "configs[4]: contrastive n=10 contrast=0.5 (mean) + idf-penalty 1.0, f=0.7, 1M-sentence TM", index, tm, off,
                                  V, batches[:2], dict(fuzzy=0.7, n=10, ml=3, idf=1.0, contrast=0.5), 10, dev, torch, capi, oracle_cache, steps, 1000)
        oracle_cache.clear()
        n_long = max(200, args.sentences // 50)
        tm4, off4, V4 = synth.make_tm(args.sentences, seed=1234, n_long=n_long)
        src = np.arange(args.sentences - n_long, args.sentences)
        nq4 = max(100, args.queries // 50)
        b4 = [synth.make_queries(tm4, off4, nq4, seed=5678 + b, source_ids=src, frac_random=0.2, len_lo=200, len_hi=300) for b in range(2)]
        index4 = fmb.Index(tm4, off4, V4, device=local_rank)
        cfgs["c4"] = config_entry("c4", "configs[3]: %d short + %d sentences of 200-300 tokens, %d queries per step = perturbed long sentences, "
                                  "max_tokens_in_pattern=300, f=0.7 n=1 ml=3" % (args.sentences - n_long, n_long, nq4), index4, tm4, off4, V4, b4,
                                  dict(fuzzy=0.7, n=1, ml=3), 1, dev, torch, capi, oracle_cache, steps, 300)
        line["configs"] = cfgs
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
