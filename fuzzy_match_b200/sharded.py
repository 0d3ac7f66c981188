"""Translation memory sharded by sentence-id range over the GPUs of one box (one process per GPU).

Layout (BASELINE.json north_star): rank r owns the contiguous sentence range
[r*n/world, (r+1)*n/world) with its own suffix array; IDF statistics (sfreq, N) are global
(one all-reduce at build time); every rank scores the whole query batch against its shard
(fm_shard_score_device), the per-shard scored candidates travel in ONE all-gather per batch, and
every rank replays the union exactly like the single-index candidate loop (fm_merge_replay_device),
so results are bit-identical to an unsharded index. The reference has no counterpart (single
process, src/fuzzy_match.cc); candidates are independent per TM sentence, only the bound heap /
top-N of src/fuzzy_match.cc:567-611,670-679 is global, and that is what the merge replays.

torch.distributed is plumbing only (NCCL on GPUs; the same code paths run on gloo/CPU tensors in
tests/test_sharded_cpu.py). No kernels here.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import capi

HEADER_ALIGN = 8  # int32 words; keeps the record block 32-byte aligned behind the offsets


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


def header_words(n_q):
    return ((n_q + 1 + HEADER_ALIGN - 1) // HEADER_ALIGN) * HEADER_ALIGN


def pack_records(rec_off, rec_words, n_q, max_rec, out=None):
    """[offsets (n_q+1, padded) | records (max_rec * 8 int32 words)] in one int32 buffer."""
    hw = header_words(n_q)
    total = hw + max_rec * 8
    if out is None or out.numel() < total:
        out = torch.empty(total, dtype=torch.int32, device=rec_off.device)
    buf = out[:total]
    buf[:n_q + 1].copy_(rec_off[:n_q + 1])
    n_words = rec_words.numel()
    buf[hw:hw + n_words].copy_(rec_words)
    return buf


def gather_records(buf, group=None, out=None):
    """The single data collective of a batch: all_gather of the packed per-shard buffers."""
    world = dist.get_world_size(group)
    if out is None or out.numel() < world * buf.numel():
        out = torch.empty(world * buf.numel(), dtype=torch.int32, device=buf.device)
    recv = out[:world * buf.numel()]
    dist.all_gather_into_tensor(recv, buf, group=group)
    return recv.view(world, buf.numel())


class _DevView:
    """Zero-copy view of a raw device pointer for torch.as_tensor (CUDA array interface v3)."""

    def __init__(self, ptr, n_words):
        self.__cuda_array_interface__ = {"shape": (n_words,), "typestr": "<i4", "data": (ptr, False), "version": 3}


class ShardedIndex:
    """One rank's shard plus the collectives around it."""

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
        self._send = None
        self._recv = None
        self._part = self._part_cnt = self._all = self._all_cnt = None
        self.last_gather_bytes = 0

    def match_batch_device(self, d_q_tok, d_q_off, n_q, n_tok, d_out, d_out_count, cap, params, stream=None):
        """torch int32 CUDA tensors in, results in d_out (uint8/any tensor of n_q*cap*24 bytes) and
        d_out_count (int32[n_q]) on every rank."""
        st = torch.cuda.current_stream(self.device) if stream is None else stream
        sp = st.cuda_stream
        off_ptr, rec_ptr, n_rec = self.index.shard_score_device(d_q_tok.data_ptr(), d_q_off.data_ptr(), n_q, n_tok,
                                                                stream=sp, params=params)
        with torch.cuda.stream(st):
            cnt = torch.tensor([n_rec], dtype=torch.int64, device=self.device)
            allc = torch.empty(self.world, dtype=torch.int64, device=self.device)
            dist.all_gather_into_tensor(allc, cnt, group=self.group)
            max_rec = int(allc.max().item())
            rec_off = torch.as_tensor(_DevView(off_ptr, n_q + 1), device=self.device)
            rec_words = torch.as_tensor(_DevView(rec_ptr, max(n_rec, 1) * 8), device=self.device)[:n_rec * 8]
            need = header_words(n_q) + max_rec * 8
            if self._send is None or self._send.numel() < need:
                self._send = torch.empty(need + need // 4, dtype=torch.int32, device=self.device)
            if self._recv is None or self._recv.numel() < self.world * need:
                self._recv = torch.empty(self.world * (need + need // 4), dtype=torch.int32, device=self.device)
            buf = pack_records(rec_off, rec_words, n_q, max_rec, self._send)
            recv = gather_records(buf, self.group, self._recv)
            self.last_gather_bytes = recv.numel() * 4
            hw = header_words(n_q)
            # Every rank holds every shard's records now; the replay of the union is split by query
            # range (rank r replays queries [r*per, (r+1)*per)), then two small all-gathers hand every
            # rank the complete result.
            per = (n_q + self.world - 1) // self.world
            q_lo = min(n_q, self.rank * per)
            q_cnt = min(n_q, q_lo + per) - q_lo
            msz = cap * capi.MATCH_DTYPE.itemsize
            if self._part is None or self._part.numel() < per * msz:
                self._part = torch.zeros(per * msz, dtype=torch.uint8, device=self.device)
                self._part_cnt = torch.zeros(per, dtype=torch.int32, device=self.device)
                self._all = torch.zeros(self.world * per * msz, dtype=torch.uint8, device=self.device)
                self._all_cnt = torch.zeros(self.world * per, dtype=torch.int32, device=self.device)
            if q_cnt > 0:
                offs = [recv[k].data_ptr() + q_lo * 4 for k in range(self.world)]
                recs = [recv[k].data_ptr() + hw * 4 for k in range(self.world)]
                self.index.merge_replay_device(offs, recs, d_q_off.data_ptr() + q_lo * 4, q_cnt, self._part.data_ptr(),
                                               self._part_cnt.data_ptr(), cap, stream=sp, params=params)
            dist.all_gather_into_tensor(self._all[:self.world * per * msz], self._part[:per * msz], group=self.group)
            dist.all_gather_into_tensor(self._all_cnt[:self.world * per], self._part_cnt[:per], group=self.group)
            d_out.view(torch.uint8).reshape(-1)[:n_q * msz].copy_(self._all[:n_q * msz])
            d_out_count[:n_q].copy_(self._all_cnt[:n_q])

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
