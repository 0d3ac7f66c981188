"""world_size-2 gloo tests (CPU) of the host-side logic of the sharded path: shard ranges and s_id
bases, the sfreq all-reduce, and the pack / all-gather / unpack of per-shard record buffers."""
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
        # per-shard record buffers of different sizes
        n_q = 37
        rng = np.random.default_rng(100 + rank)
        cnt = rng.integers(0, 4 + 3 * rank, size=n_q)
        rec_off = torch.zeros(n_q + 1, dtype=torch.int32)
        rec_off[1:] = torch.as_tensor(np.cumsum(cnt).astype(np.int32))
        n_rec = int(rec_off[-1])
        rec_words = torch.as_tensor(rng.integers(0, 1 << 30, size=n_rec * 8).astype(np.int32))
        allc = torch.empty(world, dtype=torch.int64)
        dist.all_gather_into_tensor(allc, torch.tensor([n_rec], dtype=torch.int64))
        max_rec = int(allc.max())
        buf = sharded.pack_records(rec_off, rec_words, n_q, max_rec)
        recv = sharded.gather_records(buf)
        hw = sharded.header_words(n_q)
        ret[rank] = dict(lo=lo, hi=hi, n_kept=n_kept, base=base, n_global=n_global, sf_global=sf_global,
                         rec_off=rec_off.numpy().copy(), rec_words=rec_words.numpy().copy(),
                         recv=recv.numpy().copy(), hw=hw, counts=allc.numpy().copy())
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
    # every rank received every shard's offsets and records at the documented positions
    for me in range(world):
        recv, hw = r[me]["recv"], r[me]["hw"]
        for k in range(world):
            n_q = len(r[k]["rec_off"]) - 1
            assert (recv[k, :n_q + 1] == r[k]["rec_off"]).all()
            nw = len(r[k]["rec_words"])
            assert (recv[k, hw:hw + nw] == r[k]["rec_words"]).all()
        assert hw % 8 == 0 and recv.shape[1] == hw + 8 * int(r[me]["counts"].max())
