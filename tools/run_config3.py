#!/usr/bin/env python
"""BASELINE.json configs[2]: 10M-sentence TM sharded by sentence-id over the GPUs of one box,
1M queries, f=0.5, ml=3, n=1, one NCCL all-gather of scored candidates per batch + merged replay.

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_config3.py

Prints one JSON line (rank 0). Parity of the sharded path is established at smaller scale
(tools/check_sharded_nccl.py, tests/test_gpu_parity.py); here a sample of queries is additionally
checked against a single-GPU unsharded index when --check is given (needs ~7 GB on rank 0)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fuzzy_match_b200 import capi, synth  # noqa: E402
from fuzzy_match_b200.sharded import ShardedIndex  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sentences", type=int, default=10000000)
    ap.add_argument("--queries", type=int, default=1000000)
    ap.add_argument("--batch", type=int, default=100000)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    os.dup2(2, 1) if False else None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    t0 = time.time()
    tm, off, V = synth.make_tm(args.sentences, seed=1234)
    t_gen = time.time() - t0
    t0 = time.time()
    index = ShardedIndex(tm, off, V, device=dev)
    dist.barrier()
    t_build = time.time() - t0
    params = capi.Params.make(fuzzy=0.5, n=1, ml=3)
    n_batches = (args.queries + args.batch - 1) // args.batch
    batches = []
    for b in range(min(n_batches, 4)):
        q, qo = synth.make_queries(tm, off, args.batch, seed=5678 + b)
        batches.append((torch.as_tensor(q, device=dev), torch.as_tensor(qo.astype(np.int32), device=dev), len(qo) - 1, int(qo[-1]), q, qo))
    d_out = torch.zeros(args.batch * 24, dtype=torch.uint8, device=dev)
    d_cnt = torch.zeros(args.batch, dtype=torch.int32, device=dev)
    stream = torch.cuda.Stream(dev)
    for b in range(min(3, len(batches))):  # warm-up
        dq, dqo, nq, ntok, _, _ = batches[b]
        index.match_batch_device(dq, dqo, nq, ntok, d_out, d_cnt, 1, params, stream=stream)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    found = 0
    with torch.cuda.stream(stream):
        e0.record(stream)
        for b in range(n_batches):
            dq, dqo, nq, ntok, _, _ = batches[b % len(batches)]
            index.match_batch_device(dq, dqo, nq, ntok, d_out, d_cnt, 1, params, stream=stream)
        e1.record(stream)
    dist.barrier()
    torch.cuda.synchronize()
    found = int((d_cnt > 0).sum().item())
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    check = None
    if args.check:
        # rank 0: same last batch against an unsharded index of the whole TM on its own GPU
        out_sh = d_out.cpu().numpy().view(capi.MATCH_DTYPE).reshape(args.batch, 1)
        cnt_sh = d_cnt.cpu().numpy()
        if rank == 0:
            full = capi.Index(tm, off, V, device=local)
            _, _, _, _, q, qo = batches[(n_batches - 1) % len(batches)]
            out1, cnt1 = full.match_batch(q, qo, cap=1, params=params)
            check = bool((cnt1 == cnt_sh).all() and all(out1[i, :cnt1[i]].tobytes() == out_sh[i, :cnt1[i]].tobytes() for i in range(len(cnt1))))
            full.close()
        dist.barrier()
    if rank == 0:
        print(json.dumps({"config": "BASELINE.json configs[2]: %d-sentence TM in %d sentence-id shards, %d queries in batches of %d, f=0.5 ml=3 n=1"
                                    % (args.sentences, world, n_batches * args.batch, args.batch),
                          "n_gpus": world, "value": n_batches * args.batch / (ms / 1e3), "unit": "queries/s", "ms_per_batch": ms / n_batches,
                          "found_fraction_last_batch": found / args.batch, "tm_generate_s": round(t_gen, 1),
                          "shard_build_s": round(t_build, 1), "shard_device_bytes": int(index.index.device_bytes),
                          "allgather_bytes_per_batch": index.last_gather_bytes, "identical_to_unsharded_index": check}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
