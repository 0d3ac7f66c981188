"""Translation memory sharded by sentence-id range over the GPUs of one box (one process per GPU).

Layout (BASELINE.json north_star): rank r owns the contiguous sentence range
[r*n/world, (r+1)*n/world) with its own suffix array; IDF statistics (sfreq, N) are global (one
all-reduce at build time). Per batch every rank runs the whole pipeline on its shard INCLUDING the
candidate loop of src/fuzzy_match.cc:567-611 and keeps the records that loop accepts; ONE NCCL all-gather
moves them (16 bytes a record, one block per shard), and every rank replays the union in the reference's
candidate order -- bit-identical to an unsharded index. All of that is one C call
(fm_match_batch_sharded_device: kernels + ncclAllGather on the caller's stream, no torch on the data
path); torch.distributed is used here only at build time (s_id bases, the sfreq all-reduce, handing the
NCCL id to the ranks; the same helpers run on gloo/CPU tensors in tests/test_sharded_cpu.py).
The reference has no counterpart (single process, src/fuzzy_match.cc).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import capi


def shard_range(n_sent, rank, world):
    """Contiguous sentence range [lo, hi) of a rank."""
    return (rank * n_sent) // world, ((rank + 1) * n_sent) // world


def kept_count(tm_off, lo, hi, max_tokens):
    """Sentences of [lo, hi) that the index keeps (non-empty, <= max_tokens; suffix_array_index.cc:16)."""
    lens = np.diff(np.asarray(tm_off[lo:hi + 1], dtype=np.int64))
    return int(((lens > 0) & (lens <= max_tokens)).sum())


def exchange_kept_counts(n_kept_local, device, group=None):
    """All ranks learn (s_id base of this rank, global number of kept sentences): one tiny all_gather."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = torch.tensor([n_kept_local], dtype=torch.int64, device=device)
    allc = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allc, mine, group=group)
    counts = allc.cpu().numpy()
    return int(counts[:rank].sum()), int(counts.sum())


def allreduce_sfreq(sfreq_local, device, group=None):
    """Global word-in-sentence frequencies: shards hold disjoint sentences, so the per-shard counts
    (reference src/vocab_indexer.cc:73-90) add up exactly."""
    sf = torch.as_tensor(np.asarray(sfreq_local, dtype=np.int64), device=device)
    dist.all_reduce(sf, op=dist.ReduceOp.SUM, group=group)
    return sf.cpu().numpy().astype(np.uint32)


def broadcast_bytes(data, n, device, src=0, group=None):
    """Rank `src` hands `n` bytes to every rank (the NCCL unique id of fm_comm_create)."""
    buf = torch.zeros(n, dtype=torch.uint8, device=device)
    if dist.get_rank(group) == src:
        buf.copy_(torch.frombuffer(bytearray(data), dtype=torch.uint8))
    dist.broadcast(buf, src=src, group=group)
    return bytes(buf.cpu().numpy().tobytes())


class ShardedIndex:
    """One rank's shard and the communicator around it."""

    def __init__(self, tm_tokens, tm_off, vocab_size, max_tokens=300, device=None, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        n_sent = len(tm_off) - 1
        lo, hi = shard_range(n_sent, self.rank, self.world)
        tm_off = np.asarray(tm_off, dtype=np.int64)
        local_off = tm_off[lo:hi + 1] - tm_off[lo]
        local_tok = np.asarray(tm_tokens[tm_off[lo]:tm_off[hi]], dtype=np.int32)
        n_kept = kept_count(tm_off, lo, hi, max_tokens)
        base, n_global = exchange_kept_counts(n_kept, self.device, group)
        self.index = capi.Index(local_tok, local_off, vocab_size, max_tokens=max_tokens, s_id_base=base,
                                device=self.device.index)
        self.index.set_idf_stats(allreduce_sfreq(self.index.sfreq(), self.device, group), n_global)
        self.s_id_base, self.n_sent_global = base, n_global
        uid = capi.comm_unique_id() if self.rank == 0 else b""
        uid = broadcast_bytes(uid, 128, self.device, 0, group)
        self.comm = capi.Comm(uid, self.rank, self.world, self.device.index)

    @property
    def last_gather_bytes(self):
        return self.comm.last_gather_bytes

    @property
    def block_capacity(self):
        return self.comm.block_capacity

    def match_batch_device(self, d_q_tok, d_q_off, n_q, n_tok, d_out, d_out_count, cap, params, stream=None):
        """torch int32 CUDA tensors in, results in d_out (uint8/any tensor of n_q*cap*24 bytes) and
        d_out_count (int32[n_q]) on every rank. Collective."""
        st = torch.cuda.current_stream(self.device) if stream is None else stream
        self.index.match_batch_sharded_device(self.comm.h, d_q_tok.data_ptr(), d_q_off.data_ptr(), n_q, n_tok, d_out.data_ptr(),
                                              d_out_count.data_ptr(), cap, stream=st.cuda_stream, params=params)

    def submit_device(self, d_q_tok, d_q_off, n_q, n_tok, d_out, d_out_count, cap, params, stream):
        """Asynchronous form: returns a ticket for wait(); several batches may be in flight, each on its own stream
        (their kernels and all-gathers overlap). Every rank submits and waits in the same order."""
        return self.index.submit_sharded_device(self.comm.h, d_q_tok.data_ptr(), d_q_off.data_ptr(), n_q, n_tok, d_out.data_ptr(),
                                                d_out_count.data_ptr(), cap, stream.cuda_stream, params)

    def wait(self, ticket):
        self.index.wait(ticket)

    def match_batch(self, q_tokens, q_off, cap, **kw):
        """Host CSR in, numpy (matches[n_q, cap], counts[n_q]) out -- convenience for tests."""
        params = capi.Params.make(**kw)
        q_off = np.asarray(q_off, dtype=np.int64)
        n_q, n_tok = len(q_off) - 1, int(q_off[-1])
        d_tok = torch.as_tensor(np.asarray(q_tokens, dtype=np.int32), device=self.device)
        d_off = torch.as_tensor(q_off.astype(np.int32), device=self.device)
        d_out = torch.zeros(n_q * cap * capi.MATCH_DTYPE.itemsize, dtype=torch.uint8, device=self.device)
        d_cnt = torch.zeros(n_q, dtype=torch.int32, device=self.device)
        self.match_batch_device(d_tok, d_off, n_q, n_tok, d_out, d_cnt, cap, params)
        torch.cuda.synchronize(self.device)
        out = d_out.cpu().numpy().view(capi.MATCH_DTYPE).reshape(n_q, cap)
        return out, d_cnt.cpu().numpy()

    def close(self):
        self.comm.close()
        self.index.close()
