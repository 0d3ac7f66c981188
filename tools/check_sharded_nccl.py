#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/check_sharded_nccl.py
Sentence-id sharded matching over NCCL must be bit-identical to the unsharded CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fuzzy_match_b200 import synth  # noqa: E402
from fuzzy_match_b200.sharded import ShardedIndex  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tm, off, V = synth.make_tm(30000, vocab=3000, len_lo=0, len_hi=30, seed=201)
    q, qo = synth.make_queries(tm, off, 3000, vocab=3000, seed=202, len_lo=1, len_hi=30)
    idx = ShardedIndex(tm, off, V, max_tokens=28, device=torch.device("cuda", local))
    ok = True
    for params in (dict(fuzzy=0.5, n=4, ml=2), dict(fuzzy=0.4, n=3, ml=3, idf=1.0, costs=(1, 0, 1)), dict(fuzzy=0.7, n=1, ml=3),
                   # contrastive rerank across shards: the accepted sentences travel in one all-reduced token slab
                   dict(fuzzy=0.5, n=5, ml=2, contrast=0.5), dict(fuzzy=0.4, n=4, ml=2, idf=0.7, costs=(1, 0, 1), contrast=0.5, reduce=1, buffer=8),
                   dict(fuzzy=0.3, n=0, ml=3, contrast=0.8)):
        out, cnt = idx.match_batch(q, qo, cap=8, **params)
        if rank == 0:
            from oracle import binding as ob
            ob.build()
            oracle = ob.OracleIndex(tm, off, V, max_tokens=28)
            ro, oc = oracle.match_batch(q, qo, cap=8, nthreads=8, **params)
            same = (cnt == oc).all() and all(out[i, :cnt[i]].tobytes() == ro[i].tobytes() for i in range(len(oc)))
            print("params", params, "world", dist.get_world_size(), "identical to oracle:", bool(same), "found", int((oc > 0).sum()),
                  "allgather bytes", idx.last_gather_bytes)
            ok = ok and same
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
