"""world_size-2 gloo tests (CPU) of the host-side logic of the sharded path: shard ranges and s_id
bases, the sfreq all-reduce, handing the communicator id to the ranks, and the wire-block layout the
ranks exchange (include/fuzzy_match_b200.h: fm_wire_block_bytes)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fuzzy_match_b200 import sharded, synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tm, off, V = synth.make_tm(2001, vocab=300, len_lo=0, len_hi=40, seed=5)
        max_tokens = 30
        lo, hi = sharded.shard_range(len(off) - 1, rank, world)
        n_kept = sharded.kept_count(off, lo, hi, max_tokens)
        base, n_global = sharded.exchange_kept_counts(n_kept, "cpu")
        # local sfreq the way the index counts it (once per sentence, kept sentences only)
        sf = np.zeros(V, dtype=np.int64)
        for s in range(lo, hi):
            sent = tm[off[s]:off[s + 1]]
            if 0 < len(sent) <= max_tokens:
                sf[np.unique(sent)] += 1
        sf_global = sharded.allreduce_sfreq(sf, "cpu")
        # the 128-byte communicator id travels from rank 0 to everybody
        secret = bytes((7 * i + 3) % 256 for i in range(128))
        got = sharded.broadcast_bytes(secret if rank == 0 else b"", 128, "cpu")
        ret[rank] = dict(lo=lo, hi=hi, n_kept=n_kept, base=base, n_global=n_global, sf_global=sf_global, uid=got, secret=secret)
    finally:
        dist.destroy_process_group()


def test_sharded_host_logic_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    r = [ret[k] for k in range(world)]
    tm, off, V = synth.make_tm(2001, vocab=300, len_lo=0, len_hi=40, seed=5)
    lens = np.diff(off)
    kept = (lens > 0) & (lens <= 30)
    # contiguous, disjoint, complete ranges; s_id bases are the kept counts before the shard
    assert r[0]["lo"] == 0 and r[0]["hi"] == r[1]["lo"] and r[1]["hi"] == 2001
    assert r[0]["base"] == 0 and r[1]["base"] == int(kept[:r[1]["lo"]].sum())
    assert r[0]["n_global"] == r[1]["n_global"] == int(kept.sum())
    # global sfreq equals the unsharded count
    sf = np.zeros(V, dtype=np.int64)
    for s in np.nonzero(kept)[0]:
        sf[np.unique(tm[off[s]:off[s + 1]])] += 1
    assert (r[0]["sf_global"] == sf).all() and (r[1]["sf_global"] == sf).all()
    assert r[0]["uid"] == r[1]["uid"] == r[0]["secret"]


def test_wire_block_layout():
    """header (4 int32) | offsets (n_q + 1, padded to 4) | 16-byte records: the size every rank computes for the all-gather."""
    from fuzzy_match_b200 import capi
    assert capi.WIRE_DTYPE.itemsize == 16
    for n_q, cap in ((1, 1), (37, 2), (100000, 65536), (4096, 0)):
        assert capi.wire_block_bytes(n_q, cap) == 4 * (4 + (n_q + 1 + 3) // 4 * 4) + 16 * cap
        assert capi.wire_block_bytes(n_q, cap) % 16 == 0
