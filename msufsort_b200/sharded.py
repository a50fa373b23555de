"""One text sharded over the GPUs of one box: one process and one engine context per GPU, plumbing
through ``torch.distributed`` (NCCL over NVLink/NVSwitch on GPUs; gloo in the CPU tests, where the
kernels run under the emulator build).

Scheme (north_star item 4; include/b200sa.h "sharded building blocks"):
  * every rank holds the whole text and a replica of the ISA;
  * round 0 is partitioned by key range — each rank packs all n initial keys, derives the same
    G-1 splitters from a sorted regular sample (identical on all ranks, no communication) and
    radix-sorts only the suffixes whose key falls into its range;
  * a range consists of whole groups, so every doubling round sorts locally;
  * the ISA is what travels.  ``isa="owner"`` (default): rank g is the authority for the ranks of the text
    positions [g*B, (g+1)*B); after a round every new (suffix, rank) pair is routed to the owner of the suffix
    (one radix sweep by owner + all-to-all), and before the next round every rank asks the owners for the
    ranks it is about to read (all-to-all of positions, all-to-all of values) and caches the replies in its
    own array — all three exchanges and all ISA work scale with 1/G.  ``isa="replicated"``: the pairs are
    all-gathered and every rank applies all of them (simpler, but the ISA update is replicated work);
    ``isa="peer"`` (the NVLink-native variant, what bench.py runs): the ISA is sharded the same way, but every rank
    maps the arrays of all peers (CUDA IPC) and the kernels load rank[suffix + h] from, and store new ranks into, the
    owner's HBM directly in bulk (an inbox per GPU, applied locally).  The round loop of this variant lives in C++
    (b200sa_shard_sort, csrc/engine_shard.inl); its control plane — two barriers and a sum per round — is a
    shared-memory segment the ranks of the node map (csrc/comm.cuh), so no collective library call is left on the
    path; torch.distributed only hands out the segment's name once;
  * a rank ends up owning a contiguous slice of the suffix array and of the BWT.
The exchanges are the only collectives on the data path; counts and the termination test are tiny.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

import os

from .api import Comm, torch_stream_handle


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
        self.counts = None        # suffixes owned by every rank (ShardedSorter.owned_counts)
        self.n_local = 0          # suffixes owned by this rank


class ShardedSorter:
    def __init__(self, engine, group: Optional[dist.ProcessGroup] = None, isa: str = "peer"):
        assert isa in ("owner", "replicated", "peer")
        self.eng = engine
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.isa = isa
        self.comm: Optional[Comm] = None
        assert self.world <= 256, "the owner routing sweep has 256 buckets"
        assert isa != "peer" or self.world <= 16, "the peer table holds 16 GPUs (one NVSwitch domain)"

    def _get_comm(self) -> Comm:
        """the shared-memory control plane of the C++ round loop; rank 0 picks a fresh segment name"""
        if self.comm is None:
            name = [None]
            if self.rank == 0:
                name[0] = "/b200sa_%d_%s" % (os.getpid(), os.urandom(6).hex())
            src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
            dist.broadcast_object_list(name, src=src, group=self.group)
            self.comm = Comm.shared_memory(name[0], self.rank, self.world, library=self.eng.lib)
        return self.comm

    def release(self) -> None:
        """Frees the engine's device workspace on every rank — collectively: buffers that peers have mapped through CUDA IPC
        must not be freed before every importer has closed its mapping (all ranks detach, meet, and only then free)."""
        self.eng.shard_peer_detach()
        if self.world > 1:
            dist.barrier(group=self.group)
        self.eng.release_workspace()

    def close(self) -> None:
        """Collective: closes the peer mappings on every rank before any rank goes on to destroy its engine."""
        self.eng.shard_peer_detach()
        if self.world > 1:
            dist.barrier(group=self.group)
        if self.comm is not None:
            self.comm.close()
            self.comm = None

    # -- owner-sharded ISA ---------------------------------------------------------------------------
    def _owner_shift(self, n: int) -> int:
        per = (n + self.world - 1) // self.world
        return max(0, (per - 1).bit_length())        # B = 2^shift >= ceil(n/G): owner(p) = p >> shift < G for p < n

    def _a2a(self, send: torch.Tensor, send_counts: list, recv_counts: list) -> torch.Tensor:
        out = torch.empty(sum(recv_counts), dtype=send.dtype, device=send.device)
        dist.all_to_all_single(out, send, recv_counts, send_counts, group=self.group)
        return out

    def _exchange_counts(self, counts: list, device) -> list:
        sc = torch.tensor(counts, dtype=torch.int64, device=device)
        rc = torch.empty(self.world, dtype=torch.int64, device=device)
        dist.all_to_all_single(rc, sc, group=self.group)
        return [int(x) for x in rc.cpu().tolist()]

    def _route_updates(self, device, stream: int, res: "ShardedResult", shift: int) -> None:
        """new (suffix, rank) pairs -> the owners of the suffixes"""
        pi, pr, cnt = self.eng.shard_updates()
        kout = torch.empty(max(cnt, 1), dtype=torch.int32, device=device)
        vout = torch.empty(max(cnt, 1), dtype=torch.int32, device=device)
        counts = self.eng.shard_partition(pi, pr, cnt, shift, kout, vout, stream)[: self.world]
        rcounts = self._exchange_counts(counts, device)
        ridx = self._a2a(kout[:cnt], counts, rcounts)
        rrank = self._a2a(vout[:cnt], counts, rcounts)
        if device.type == "cuda":
            torch.cuda.current_stream().synchronize()
        res.exchanged_bytes += 8 * (sum(rcounts) - rcounts[self.rank])
        if ridx.numel():
            self.eng.shard_apply_updates(ridx, rrank, ridx.numel(), stream)

    def _fetch_lookups(self, m_local: int, device, stream: int, res: "ShardedResult", shift: int) -> None:
        """ranks the next round reads: ask the owners, cache the replies in the local array"""
        pos = torch.empty(max(m_local, 1), dtype=torch.int32, device=device)
        cnt = self.eng.shard_requests(pos, m_local, stream) if m_local else 0
        kout = torch.empty(max(cnt, 1), dtype=torch.int32, device=device)
        vout = torch.empty(max(cnt, 1), dtype=torch.int32, device=device)
        counts = self.eng.shard_partition(pos, None, cnt, shift, kout, vout, stream)[: self.world]
        rcounts = self._exchange_counts(counts, device)
        rpos = self._a2a(kout[:cnt], counts, rcounts)                 # positions other ranks want from my shard
        vals = torch.empty(max(rpos.numel(), 1), dtype=torch.int32, device=device)
        if device.type == "cuda":
            torch.cuda.current_stream().synchronize()
        if rpos.numel():
            self.eng.shard_gather_ranks(rpos, rpos.numel(), vals, stream)
        replies = self._a2a(vals[: rpos.numel()], rcounts, counts)    # come back in the order of kout
        if device.type == "cuda":
            torch.cuda.current_stream().synchronize()
        res.exchanged_bytes += 4 * (sum(rcounts) - rcounts[self.rank]) + 4 * (cnt - counts[self.rank])
        if cnt:
            self.eng.shard_apply_updates(kout[:cnt], replies, cnt, stream)

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
        stream = torch_stream_handle() if device.type == "cuda" else 0
        res = ShardedResult()
        res.sa = torch.empty(n + 1, dtype=torch.int32, device=device)
        shift = self._owner_shift(n)
        if self.isa == "peer" and self.world > 1:
            if want_bwt:
                res.bwt = torch.empty(n, dtype=torch.uint8, device=device)
            info = self.eng.shard_sort(self._get_comm(), d_text, n, res.sa, res.bwt, stream)
            res.row_begin, res.row_end = info["row_begin"], info["row_end"]
            res.out_begin, res.out_end = info["out_begin"], info["out_end"]
            res.sentinel, res.rounds, res.exchanged_bytes = info["sentinel"], info["rounds"], info["sent_bytes"]
            res.n_local = info["n_local"]   # owned_counts() gathers these outside the hot path
            return res
        n_local = self.eng.shard_begin(d_text, n, res.sa, self.rank, self.world, stream)
        res.n_local = n_local
        res.counts = self._gather_int(n_local, device)
        assert sum(res.counts) == n, "key-range parts do not cover the text"
        slot_base = sum(res.counts[: self.rank])
        m_local = self.eng.shard_round0(slot_base, stream)
        res.rounds = 1
        while True:
            if self.isa == "owner":
                self._route_updates(device, stream, res, shift)
            else:
                self._exchange_updates(device, stream, res, n)
            t = torch.tensor([m_local], dtype=torch.int64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            if int(t.item()) == 0:
                break
            if self.isa == "owner":
                self._fetch_lookups(m_local, device, stream, res, shift)
            m_local = self.eng.shard_round(stream)
            res.rounds += 1
        if self.isa == "owner":
            # the sentinel row s = rank[0] lives on the owner of position 0
            t = torch.zeros(1, dtype=torch.int32, device=device)
            if self.rank == 0:
                self.eng.shard_gather_ranks(torch.zeros(1, dtype=torch.int32, device=device), 1, t, stream)
            src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
            dist.broadcast(t, src=src, group=self.group)
            if self.rank != 0:
                self.eng.shard_apply_updates(torch.zeros(1, dtype=torch.int32, device=device), t, 1, stream)
        # rows of the (n+1)-row suffix array owned here; row 0 (the empty suffix) belongs to rank 0
        res.row_begin = 0 if self.rank == 0 else slot_base + 1
        res.row_end = slot_base + n_local + 1
        if want_bwt:
            res.bwt = torch.empty(n, dtype=torch.uint8, device=device)
            res.out_begin, res.out_end, res.sentinel = self.eng.shard_bwt(res.row_begin, res.row_end, res.bwt, stream)
        return res

    # -- the sharded inverse BWT ------------------------------------------------------------------------
    def inverse_bwt(self, d_bwt: torch.Tensor, sentinel_index: int, gather_all: bool = True):
        """Every rank passes the same BWT.  C++ path (isa="peer"): the psi table is built on every GPU, the walkers are split
        over the ranks and the bytes are stored into the owner of their text position over NVLink.  gather_all=True: the
        other slices are pulled from the peers as well and the whole text is returned on every rank.  gather_all=False:
        returns (text tensor, begin, end) — only bytes [begin, end), the slice this rank owns, are valid (the analogue of
        owning a slice of the suffix array after suffix_array_bwt).
        NCCL path (other ISA modes): psi replicated, walkers partitioned, (length, successor) entries all-gathered, disjoint
        output slices combined with a sum all-reduce; always returns the whole text."""
        n = d_bwt.numel()
        device = d_bwt.device
        stream = torch_stream_handle() if device.type == "cuda" else 0
        if self.isa == "peer" and self.world > 1:
            out = torch.empty(n, dtype=torch.uint8, device=device)
            b, e = self.eng.shard_unbwt(self._get_comm(), d_bwt, n, sentinel_index, out, gather_all, stream)
            return out if gather_all else (out, b, e)
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

    def suffix_array_bwt_host(self, h_text: torch.Tensor, h_sa: torch.Tensor, h_bwt: Optional[torch.Tensor]) -> ShardedResult:
        """Host buffers in, host buffers out (pinned tensors; every rank passes the same text): every rank uploads 1/G of
        the text over its own PCIe link and the slices are all-gathered over NVLink; the suffix array and the BWT leave
        as this rank's rows / bytes only — the caller's host arrays are filled by the ranks together (shared pinned
        memory) or gathered by the caller."""
        n = h_text.numel()
        if self.world > 1 and n >= 16 * self.world:
            # every rank uploads one slice over its own PCIe link; the slices are all-gathered over NVLink
            per = -(-n // self.world)
            d_all = torch.empty(per * self.world, dtype=torch.uint8, device="cuda")
            lo, hi = min(n, self.rank * per), min(n, (self.rank + 1) * per)
            mine = d_all[self.rank * per: self.rank * per + (hi - lo)]
            mine.copy_(h_text[lo:hi], non_blocking=True)
            dist.all_gather_into_tensor(d_all, d_all[self.rank * per:(self.rank + 1) * per], group=self.group)
            d_text = d_all[:n]
        else:
            d_text = h_text.to("cuda", non_blocking=True)
        res = self.suffix_array_bwt(d_text, want_bwt=h_bwt is not None)
        h_sa[res.row_begin:res.row_end].copy_(res.sa[res.row_begin:res.row_end], non_blocking=True)
        if h_bwt is not None and res.out_end > res.out_begin:
            h_bwt[res.out_begin:res.out_end].copy_(res.bwt[res.out_begin:res.out_end], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return res

    def owned_counts(self, res: ShardedResult) -> list:
        if res.counts is None:
            res.counts = self._gather_int(res.n_local, res.sa.device)
        return res.counts

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
