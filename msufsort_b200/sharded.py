"""One text sharded over the GPUs of one box: one process and one engine context per GPU, plumbing
through ``torch.distributed`` (NCCL over NVLink/NVSwitch on GPUs; gloo in the CPU tests, where the
kernels run under the emulator build).

Scheme (north_star item 4; include/b200sa.h "sharded building blocks"):
  * every rank holds the whole text and a replica of the ISA;
  * round 0 is partitioned by key range — each rank packs all n initial keys, derives the same
    G-1 splitters from a sorted regular sample (identical on all ranks, no communication) and
    radix-sorts only the suffixes whose key falls into its range;
  * a range consists of whole groups, so every doubling round sorts locally; after each round the
    ranks all-gather their (suffix, new rank) ISA updates and apply them to their replicas;
  * a rank ends up owning a contiguous slice of the suffix array and of the BWT.
The exchange step is the only collective on the data path; counts and the termination test are
tiny all-gathers / all-reduces.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


class ShardedResult:
    def __init__(self):
        self.sa = None            # int32[n+1]; only rows [row_begin,row_end) are valid on this rank
        self.row_begin = 0
        self.row_end = 0
        self.bwt = None           # uint8[n]; only bytes [out_begin,out_end) are valid on this rank
        self.out_begin = 0
        self.out_end = 0
        self.sentinel = 0
        self.rounds = 0
        self.exchanged_bytes = 0  # bytes this rank received in update all-gathers
        self.counts = []          # suffixes owned by every rank


class ShardedSorter:
    def __init__(self, engine, group: Optional[dist.ProcessGroup] = None):
        self.eng = engine
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    # -- tiny collectives ------------------------------------------------------------------------
    def _gather_int(self, value: int, device) -> list:
        t = torch.tensor([value], dtype=torch.int64, device=device)
        out = torch.empty(self.world, dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(out, t, group=self.group)
        return [int(x) for x in out.cpu().tolist()]

    def _exchange_updates(self, device, stream: int, res: ShardedResult, n: int) -> None:
        _, _, cnt = self.eng.shard_updates()
        counts = self._gather_int(cnt, device)
        maxc = max(counts)
        if maxc == 0:
            return
        # Every rank contributes maxc pairs: its own updates padded with (suffix = n, rank = 0) — writing
        # rank[n] = 0 is a no-op (the sentinel row is 0 by definition) — so the gathered arrays can be applied
        # to the ISA replica with ONE bucketed update instead of one per peer.
        send_idx = torch.full((maxc,), n, dtype=torch.int32, device=device)  # n <= 2^31-2 fits int32
        send_rank = torch.zeros(maxc, dtype=torch.int32, device=device)
        if cnt:
            self.eng.shard_copy_updates(send_idx, send_rank, maxc, stream)
        recv_idx = torch.empty(self.world * maxc, dtype=torch.int32, device=device)
        recv_rank = torch.empty(self.world * maxc, dtype=torch.int32, device=device)
        dist.all_gather_into_tensor(recv_idx, send_idx, group=self.group)
        dist.all_gather_into_tensor(recv_rank, send_rank, group=self.group)
        if device.type == "cuda":
            torch.cuda.current_stream().synchronize()
        res.exchanged_bytes += (self.world - 1) * 2 * 4 * maxc
        self.eng.shard_apply_updates(recv_idx, recv_rank, self.world * maxc, stream)

    # -- the sharded SA + BWT ----------------------------------------------------------------------
    def suffix_array_bwt(self, d_text: torch.Tensor, want_bwt: bool = True) -> ShardedResult:
        n = d_text.numel()
        device = d_text.device
        stream = torch.cuda.current_stream().cuda_stream if device.type == "cuda" else 0
        res = ShardedResult()
        res.sa = torch.empty(n + 1, dtype=torch.int32, device=device)
        n_local = self.eng.shard_begin(d_text, n, res.sa, self.rank, self.world, stream)
        res.counts = self._gather_int(n_local, device)
        assert sum(res.counts) == n, "key-range parts do not cover the text"
        slot_base = sum(res.counts[: self.rank])
        m_local = self.eng.shard_round0(slot_base, stream)
        self._exchange_updates(device, stream, res, n)
        res.rounds = 1
        while True:
            t = torch.tensor([m_local], dtype=torch.int64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            if int(t.item()) == 0:
                break
            m_local = self.eng.shard_round(stream)
            self._exchange_updates(device, stream, res, n)
            res.rounds += 1
        # rows of the (n+1)-row suffix array owned here; row 0 (the empty suffix) belongs to rank 0
        res.row_begin = 0 if self.rank == 0 else slot_base + 1
        res.row_end = slot_base + n_local + 1
        if want_bwt:
            res.bwt = torch.empty(n, dtype=torch.uint8, device=device)
            res.out_begin, res.out_end, res.sentinel = self.eng.shard_bwt(res.row_begin, res.row_end, res.bwt, stream)
        return res

    # -- the sharded inverse BWT ------------------------------------------------------------------------
    def inverse_bwt(self, d_bwt: torch.Tensor, sentinel_index: int) -> torch.Tensor:
        """psi table replicated, walkers partitioned: every rank measures and emits the segments of its slice
        of the walkers; the (length, successor) entries are all-gathered for the list ranking and the disjoint
        output slices are combined with a sum all-reduce.  Returns the whole text on every rank."""
        n = d_bwt.numel()
        device = d_bwt.device
        stream = torch.cuda.current_stream().cuda_stream if device.type == "cuda" else 0
        W = self.eng.unbwt_shard_build(d_bwt, n, sentinel_index, stream)
        per = (W + self.world - 1) // self.world
        spans = [(min(W, p * per), min(W, (p + 1) * per)) for p in range(self.world)]
        wb, we = spans[self.rank]
        self.eng.unbwt_shard_measure(wb, we, stream)
        send = torch.zeros((2, per), dtype=torch.int32, device=device)
        if we > wb:
            self.eng.unbwt_shard_segments(0, wb, we, send[0], send[1], stream)
        recv = torch.empty((self.world, 2, per), dtype=torch.int32, device=device)
        dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=self.group)
        if device.type == "cuda":
            torch.cuda.current_stream().synchronize()
        for p, (b, e) in enumerate(spans):
            if p != self.rank and e > b:
                self.eng.unbwt_shard_segments(1, b, e, recv[p, 0], recv[p, 1], stream)
        out = torch.zeros(n, dtype=torch.uint8, device=device)
        self.eng.unbwt_shard_finish(wb, we, out, stream)
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.group)
        return out

    # -- helpers for callers that want the whole result on every rank -------------------------------
    def gather_sa(self, res: ShardedResult) -> torch.Tensor:
        n1 = res.sa.numel()
        device = res.sa.device
        spans = self._gather_pairs(res.row_begin, res.row_end, device)
        maxlen = max(e - b for b, e in spans)
        send = torch.zeros(maxlen, dtype=torch.int32, device=device)
        send[: res.row_end - res.row_begin] = res.sa[res.row_begin:res.row_end]
        recv = torch.empty((self.world, maxlen), dtype=torch.int32, device=device)
        dist.all_gather_into_tensor(recv.view(-1), send, group=self.group)
        full = torch.empty(n1, dtype=torch.int32, device=device)
        for p, (b, e) in enumerate(spans):
            full[b:e] = recv[p, : e - b]
        return full

    def gather_bwt(self, res: ShardedResult) -> torch.Tensor:
        n = res.bwt.numel()
        device = res.bwt.device
        spans = self._gather_pairs(res.out_begin, res.out_end, device)
        maxlen = max(max(e - b for b, e in spans), 1)
        send = torch.zeros(maxlen, dtype=torch.uint8, device=device)
        send[: res.out_end - res.out_begin] = res.bwt[res.out_begin:res.out_end]
        recv = torch.empty((self.world, maxlen), dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(recv.view(-1), send, group=self.group)
        full = torch.empty(n, dtype=torch.uint8, device=device)
        for p, (b, e) in enumerate(spans):
            full[b:e] = recv[p, : e - b]
        return full

    def _gather_pairs(self, a: int, b: int, device) -> list:
        t = torch.tensor([a, b], dtype=torch.int64, device=device)
        out = torch.empty(2 * self.world, dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(out, t, group=self.group)
        v = out.cpu().tolist()
        return [(int(v[2 * p]), int(v[2 * p + 1])) for p in range(self.world)]
