#!/usr/bin/env python
"""bench.py -- fuzzy-match queries/sec on the BASELINE.json headline workload.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--shard queries|tm]

A step = one pass of the hot path (FuzzyMatch::match for every pattern of one batch) over one batch
of --queries synthetic patterns against the synthetic 1M-sentence TM (BASELINE.json configs[1]:
1M sentences avg 15 tokens, f=0.7, n=1, ml=3, mr=0, unit edit costs).

  value      queries/s with the batch already resident in HBM (fm_match_batch_device), CUDA events on
             the launching stream, max over ranks.
  e2e        the same through the reference-facing host-buffer call fm_match_batch: pinned host
             buffers in, host results out, H2D/D2H inside the timed region.
  roofline   the dominant kernel's algorithmic bytes / its CUDA-event time vs the measured HBM peak.
  cpu_baseline / --impl reference   the reference's own CPU implementation (oracle/_ref, built from the
             reference sources; falls back to the C restatement) on this box's host cores.

N > 1 (torchrun, one rank per GPU): --shard queries (default) replicates the 214 MB index and gives
each rank its own batch -- no data-path collective, weak scaling; --shard tm is the north-star layout
(sentence-id shards + one NCCL all-gather of scored candidates + merged replay) and is also measured
as the secondary "tm_sharded" entry of the default run.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shard", default="queries", choices=["queries", "tm"])
    ap.add_argument("--sentences", type=int, default=1000000)
    ap.add_argument("--queries", type=int, default=100000)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_peak():
    """HBM peak for the roofline: the driver-written MEASURED_PEAKS.json when present (the sustained
    figure, since the kernel is timed inside a step), else the profiling recipe's fallback."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        found = []

        def walk(x, trail):
            if isinstance(x, dict):
                for k, v in x.items():
                    walk(v, trail + [str(k)])
            elif isinstance(x, (int, float)) and not isinstance(x, bool):
                name = ".".join(trail).lower()
                if any(t in name for t in ("hbm", "dram", "copy", "mem_bw", "membw")) and not any(t in name for t in ("tflop", "tf_s", "bf16", "fp8")):
                    found.append((name, float(x)))

        walk(d, [])
        if found:
            found.sort(key=lambda kv: (0 if "sustain" in kv[0] else 1 if "burst" not in kv[0] else 2))
            name, v = found[0]
            if v < 100:  # given in TB/s
                v *= 1000.0
            if 1000.0 < v < 20000.0:
                return v, "measured (MEASURED_PEAKS.json %s)" % name
    except Exception:
        pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

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
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 8:
                    continue
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                for name, v in zip(names, f[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
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


# ------------------------------------------------------------------------------------------ CPU arm


def cpu_reference_run(tm, off, V, q, qo, seconds, threads):
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
        if kind == "reference":
            idx.match_batch(q[:qo[n]], qo[:n + 1], cap=1, nthreads=threads, **PARAMS)
            dt = idx.last_seconds
        else:
            idx.match_batch(q[:qo[n]], qo[:n + 1], cap=1, nthreads=threads, **PARAMS)
            dt = time.time() - t
        return dt

    dt = run(probe)  # warm-up + rate estimate
    n = int(min(n_q, max(probe, probe / max(dt, 1e-6) * seconds)))
    dt = run(n)
    return dict(kind=kind, n=n, seconds=dt, qps=n / dt, build_s=build_s, index=idx)


def oracle_counters(tm, off, V, q, qo, n):
    """Implementation-independent work counters from the C restatement on a sample."""
    from oracle import binding as ob
    O = ob.OracleIndex(tm, off, V)
    _, _, ct = O.match_batch(q[:qo[n]], qo[:n + 1], cap=1, nthreads=os.cpu_count() or 1, counters=True, **PARAMS)
    return {k: v / n for k, v in ct.items()}


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
        "config": workload_config(args, extra={"sample_queries_per_step": n}),
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": threads, "kind": kind,
                         "sample": "%d of the %d queries of one batch per step, %d host threads sharing one index" % (n, len(qo) - 1, threads)},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, extra=None):
    c = {"workload": "BASELINE.json configs[1]: %d-sentence synthetic TM (Zipf(1) over 50k words, len U[5,25]), "
                     "%d queries per step (80%% perturbed TM sentences, 20%% random), f=0.7, n=1, ml=3, mr=0, unit costs"
                     % (args.sentences, args.queries),
         "sentences": args.sentences, "queries_per_step": args.queries, "fuzzy": 0.7, "number_of_matches": 1,
         "min_subseq_length": 3, "min_subseq_ratio": 0.0,
         "cache": "inputs larger than L2: index arrays (~214 MB at 1M sentences) exceed the 126 MB L2 and successive "
                  "steps use different query batches"}
    if extra:
        c.update(extra)
    return c


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

    shard_tm = world > 1 and args.shard == "tm"
    n_batches = min(4, args.steps + args.warmup)
    tm, off, V, batches = workload(args, rank=rank, n_batches=n_batches, shard_queries=(world > 1 and not shard_tm))
    params = capi.Params.make(**PARAMS)
    cap = 1
    t0 = time.time()
    if shard_tm:
        from fuzzy_match_b200.sharded import ShardedIndex
        sharded_index = ShardedIndex(tm, off, V, device=dev)
        index = sharded_index.index
    else:
        index = fmb.Index(tm, off, V, device=local_rank)
    build_s = time.time() - t0

    # device-resident copies of the batches + output buffers
    stream = torch.cuda.Stream(dev)
    dbatches = []
    for q, qo in batches:
        dbatches.append((torch.as_tensor(q, device=dev), torch.as_tensor(qo.astype(np.int32), device=dev), len(qo) - 1, int(qo[-1])))
    n_q = args.queries
    d_out = torch.zeros(n_q * cap * 24, dtype=torch.uint8, device=dev)
    d_cnt = torch.zeros(n_q, dtype=torch.int32, device=dev)

    def step_device(i):
        dq, dqo, nq, ntok = dbatches[i % n_batches]
        if shard_tm:
            sharded_index.match_batch_device(dq, dqo, nq, ntok, d_out, d_cnt, cap, params, stream=stream)
        else:
            index.match_batch_device(dq.data_ptr(), dqo.data_ptr(), nq, ntok, d_out.data_ptr(), d_cnt.data_ptr(), cap,
                                     stream=stream.cuda_stream, params=params)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- value: device-resident, CUDA events on the launching stream
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(1.5)  # let nvidia-smi finish attaching before anything is timed
    for i in range(max(args.warmup, n_batches)):  # every distinct batch once, so no workspace growth is timed
        step_device(i)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for i in range(args.steps):
            step_device(args.warmup + i)
        ev1.record(stream)
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    found = int((d_cnt > 0).sum().item())

    # ---- e2e: host buffers through fm_match_batch (pinned), wall clock around synchronous calls
    e2e_ms = None
    h2d = d2h = 0
    if not shard_tm:
        pinned = []
        for q, qo in batches:
            pq = torch.empty(len(q), dtype=torch.int32).pin_memory()
            pq.numpy()[:] = q
            pinned.append((pq.numpy(), qo))
        pout = torch.empty(n_q * cap * 24, dtype=torch.uint8).pin_memory().numpy().view(capi.MATCH_DTYPE).reshape(n_q, cap)
        pcnt = torch.empty(n_q, dtype=torch.int32).pin_memory().numpy()
        for i in range(args.warmup):
            index.match_batch(pinned[i % n_batches][0], pinned[i % n_batches][1], cap=cap, params=params, out=pout, cnt=pcnt)
        barrier()
        t = time.perf_counter()
        for i in range(args.steps):
            pq, qo = pinned[(args.warmup + i) % n_batches]
            index.match_batch(pq, qo, cap=cap, params=params, out=pout, cnt=pcnt)
        torch.cuda.synchronize(dev)
        e2e_ms = 1e3 * (time.perf_counter() - t)
        h2d = int(np.mean([4 * len(b[0]) + 4 * len(b[1]) for b in batches]))
        d2h = n_q * cap * 24 + n_q * 4 + 32
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel times (CUDA events inside the library, same stream) for the roofline
    prof = None
    if not shard_tm:
        index.set_profiling(True)
        acc = {}
        reps = max(3, min(args.steps, 10))
        for i in range(reps):
            step_device(i)
            torch.cuda.synchronize(dev)
            p = index.profile()
            for k, v in p.items():
                acc[k] = acc.get(k, 0) + v
        prof = {k: v / reps for k, v in acc.items()}
        index.set_profiling(False)

    # ---- max over ranks
    times = torch.tensor([dev_ms, e2e_ms if e2e_ms is not None else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms_max = float(times[0]), float(times[1])
    units = args.queries * args.steps * (world if (world > 1 and not shard_tm) else 1)
    value = units / (dev_ms / 1e3)

    # secondary measurement on N > 1: the north-star TM-sharded layout on the same TM
    tm_sharded = None
    if world > 1 and not shard_tm:
        from fuzzy_match_b200.sharded import ShardedIndex
        sidx = ShardedIndex(tm, off, V, device=dev)
        from fuzzy_match_b200 import synth
        q0, qo0 = synth.make_queries(tm, off, args.queries, seed=5678)
        dq, dqo = torch.as_tensor(q0, device=dev), torch.as_tensor(qo0.astype(np.int32), device=dev)
        for i in range(args.warmup):
            sidx.match_batch_device(dq, dqo, len(qo0) - 1, int(qo0[-1]), d_out, d_cnt, cap, params, stream=stream)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for i in range(args.steps):
                sidx.match_batch_device(dq, dqo, len(qo0) - 1, int(qo0[-1]), d_out, d_cnt, cap, params, stream=stream)
            e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tm_sharded = {"value": args.queries * args.steps / (float(t[0]) / 1e3), "unit": "queries/s", "scaling": "strong",
                      "layout": "TM split into %d sentence-id shards, queries replicated, one NCCL all-gather of scored "
                                "candidates per step + merged replay" % world,
                      "ms_per_step": float(t[0]) / args.steps, "allgather_bytes_per_step": sidx.last_gather_bytes}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    line = {
        "metric": "fuzzy-match queries/sec @ 1M-sent TM f=0.7", "value": value, "unit": "queries/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if shard_tm else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, extra={
            "parallelism": ("tm-sharded x%d + NCCL all-gather" % world) if shard_tm else ("query-sharded replicas x%d" % world if world > 1 else "1 GPU"),
            "index_build_s": round(build_s, 2), "index_device_bytes": int(index.device_bytes), "found_fraction": found / n_q}),
        "gpu_launches": int((prof["launches"] if prof else 8) * args.steps),
        "clocks": clocks,
    }
    if e2e_ms is not None:
        line["e2e"] = {"value": units / (e2e_ms_max / 1e3), "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "ms_per_step": e2e_ms_max / args.steps, "api": "fm_match_batch (host CSR in pinned memory -> host fm_match[])"}
    if tm_sharded:
        line["tm_sharded"] = tm_sharded

    cpu = None
    counters = None
    if world == 1 and not args.no_cpu_baseline:
        q, qo = batches[0]
        threads = os.cpu_count() or 1
        cpu = cpu_reference_run(tm, off, V, q, qo, args.cpu_seconds, threads)
        counters = oracle_counters(tm, off, V, q, qo, min(len(qo) - 1, 4000))
        line["cpu_baseline"] = {"value": cpu["qps"], "unit": "queries/s", "cores": threads, "kind": cpu["kind"],
                                "sample": "first %d queries of batch 0 in %.1f s, %d host threads on one shared index "
                                          "(index build %.1f s excluded)" % (cpu["n"], cpu["seconds"], threads, cpu["build_s"])}
    if prof:
        peak, peak_src = measured_peak()
        stages = {k[3:]: prof[k] for k in ("ms_prepare", "ms_search", "ms_gather", "ms_scan", "ms_score", "ms_replay")}
        dom = max(stages, key=stages.get)
        traffic = None
        try:  # DRAM bytes per launch from the committed ncu --set full capture of this workload
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                traffic = json.load(f)["fm_%s_kernel" % dom]["traffic_bytes_per_launch"]
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": "fm_%s_kernel" % dom, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "traffic": traffic,
                "stage_ms": {k: round(v, 4) for k, v in stages.items()}, "elements_per_step": prof["n_elements"],
                "slices_per_step": prof["n_slices"], "survivors_per_step": prof["n_survivors"]}
        if counters:
            # algorithmic bytes per query, SURVEY.md 8d: search 16 B/probe; gather 8 B/element walked +
            # (8 + 4*s) per deduplicated candidate (the coverage fetch is fused into the gather kernel);
            # score 4*s per DP pair + 4*p pattern; 16 B per returned match.
            per_q = {"search": 16 * counters["probes"],
                     "gather": 8 * counters["elements_walked"] + 8 * counters["candidates"] + 4 * counters["candidate_tokens"],
                     "score": 4 * counters["dp_tokens"] + 4 * counters["pattern_tokens"],
                     "replay": 16 * counters["matches_out"], "prepare": 8 * counters["pattern_tokens"], "scan": 8.0}
            roof["algorithmic_bytes_per_query"] = {k: round(v, 1) for k, v in per_q.items()}
            for k in stages:
                roof.setdefault("achieved_by_stage", {})[k] = round(per_q[k] * n_q / (stages[k] * 1e-3) / 1e9, 2) if stages[k] > 0 else None
            ach = per_q[dom] * n_q / (stages[dom] * 1e-3) / 1e9
            roof.update({"achieved": ach, "frac": ach / peak})
            total_bytes = sum(per_q[k] for k in ("search", "gather", "score", "replay"))
            roof["whole_step"] = {"bytes_per_query": round(total_bytes, 1), "achieved": total_bytes * n_q / (prof["ms_total"] * 1e-3) / 1e9,
                                  "frac": total_bytes * n_q / (prof["ms_total"] * 1e-3) / 1e9 / peak}
        line["roofline"] = roof
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
