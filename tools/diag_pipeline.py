#!/usr/bin/env python
"""Diagnostic: device-resident loop at 1 and 2 batches in flight, with retries and host-side call times."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import fuzzy_match_b200 as fmb
from fuzzy_match_b200 import capi, synth
import bench

class A: sentences = 1000000; queries = 100000
tm, off, V, batches = bench.workload(A, n_batches=4)
dev = torch.device("cuda", 0)
index = fmb.Index(tm, off, V)
params = capi.Params.make(**bench.PARAMS)
for depth in (1, 2, 3):
    r = bench.Runner(index, batches, params, 1, dev, torch, capi, depth=depth)
    r.device_loop(0, 8)
    torch.cuda.synchronize()
    ms = r.time_device(3, 20)
    index.set_profiling(True)
    t0 = time.perf_counter(); r.device_loop(0, 4); torch.cuda.synchronize(); t1 = time.perf_counter()
    p = index.profile()
    index.set_profiling(False)
    print("depth", depth, "ms/step", ms / 20, "profiled 4 steps wall ms", 1e3 * (t1 - t0), "retries", p["retries"], "launches", p["launches"],
          {k: round(v, 3) for k, v in p.items() if k.startswith("ms_")}, flush=True)
    # host-side cost of submit and wait
    ts, tw = [], []
    tickets = []
    for i in range(12):
        dq, dqo, nq, ntok = r.dbatches[i % 4]
        if len(tickets) >= depth:
            t = time.perf_counter(); index.wait(tickets.pop(0)); tw.append(time.perf_counter() - t)
        t = time.perf_counter()
        tickets.append(index.submit_device(dq.data_ptr(), dqo.data_ptr(), nq, ntok, r.d_out[i % depth].data_ptr(), r.d_cnt[i % depth].data_ptr(), 1, r.stream.cuda_stream, params))
        ts.append(time.perf_counter() - t)
    for t in tickets: index.wait(t)
    print("   submit ms", [round(1e3 * x, 3) for x in ts], "wait ms", [round(1e3 * x, 3) for x in tw], flush=True)
