#!/usr/bin/env python
"""Runs a few device-resident batches of one parameter set on the synthetic 1M-sentence TM (for ncu captures of
configurations other than the headline). Usage: python tools/run_params.py --fuzzy 0.5 [--n 1] [--ml 3] [--steps 4]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import fuzzy_match_b200 as fmb  # noqa: E402
from fuzzy_match_b200 import capi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sentences", type=int, default=1000000)
    ap.add_argument("--queries", type=int, default=100000)
    ap.add_argument("--fuzzy", type=float, default=0.5)
    ap.add_argument("--n", type=int, default=1)
    ap.add_argument("--ml", type=int, default=3)
    ap.add_argument("--steps", type=int, default=4)
    a = ap.parse_args()
    tm, off, V = synth.make_tm(a.sentences, seed=1234)
    q, qo = synth.make_queries(tm, off, a.queries, seed=5678)
    index = fmb.Index(tm, off, V)
    dev = torch.device("cuda", 0)
    dq, dqo = torch.as_tensor(q, device=dev), torch.as_tensor(qo.astype(np.int32), device=dev)
    cap = max(1, a.n)
    d_out = torch.zeros(a.queries * cap * 24, dtype=torch.uint8, device=dev)
    d_cnt = torch.zeros(a.queries, dtype=torch.int32, device=dev)
    params = capi.Params.make(fuzzy=a.fuzzy, n=a.n, ml=a.ml)
    index.set_profiling(True)
    for _ in range(a.steps):
        index.match_batch_device(dq.data_ptr(), dqo.data_ptr(), a.queries, int(qo[-1]), d_out.data_ptr(), d_cnt.data_ptr(), cap, params=params)
    torch.cuda.synchronize()
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in index.profile().items()})


if __name__ == "__main__":
    main()
