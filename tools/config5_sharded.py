#!/usr/bin/env python
"""BASELINE.json configs[4] (pathological deep-doubling strings) with ONE text sharded over the GPUs of a box.
usage (torchrun, one process per GPU): config5_sharded.py family n [repeats]
The suffix array is validated by the O(n) GPU validator on rank 0's GPU after an all-gather of the slices (outside the
timed region); the BWT by the sentinel row and a sample of its definition."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
from msufsort_b200.api import Engine, torch_stream_handle
from msufsort_b200.sharded import ShardedSorter
from msufsort_b200 import textgen

family = sys.argv[1]; n = int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:
    eng = Engine(local)
    sorter = ShardedSorter(eng, isa="peer")
    x = textgen.GENERATORS[family](n)
    d_text = torch.from_numpy(x).cuda()
    del x
    res = sorter.suffix_array_bwt(d_text)                 # warm-up: allocates the workspace, maps the peers
    times = []
    for _ in range(reps):
        dist.barrier(); torch.cuda.synchronize()
        eng.profile_reset(); eng.set_profiling(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = sorter.suffix_array_bwt(d_text)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    prof = eng.profile(); eng.set_profiling(False)
    counts = sorter.owned_counts(res)
    # validation outside the timed region; the workspace goes first (the gathered suffix array needs the room)
    sa_slice = res.sa[res.row_begin:res.row_end].clone()
    bwt_slice = res.bwt[res.out_begin:res.out_end].clone()
    rb, re, ob, oe, s = res.row_begin, res.row_end, res.out_begin, res.out_end, res.sentinel
    res.sa = None; res.bwt = None
    sorter.release(); torch.cuda.empty_cache()
    full = torch.empty(n + 1, dtype=torch.int32, device="cuda") if rank == 0 else None
    # slices travel to rank 0 one after the other (point-to-point: no G-fold staging buffers)
    for src in range(world):
        meta = torch.tensor([rb, re], dtype=torch.int64, device="cuda")
        dist.broadcast(meta, src=src)
        b, e = int(meta[0]), int(meta[1])
        if src == 0:
            if rank == 0:
                full[b:e] = sa_slice
        elif rank == src:
            dist.send(sa_slice, dst=0)
        elif rank == 0:
            dist.recv(full[b:e], src=src)
    bad = -1
    if rank == 0:
        bad = eng.check_suffix_array_dev(d_text, n, full, torch_stream_handle())
        ok_sentinel = int(full[s]) == 0 and int(full[0]) == n
        rows = torch.randint(max(1, rb), re, (4096,), device="cuda")
        rows = rows[rows != s]
        out_idx = rows - (rows > s).long()
        ok_bwt = bool((bwt_slice[out_idx - ob] == d_text[(full[rows].long() - 1)]).all())
        ms = min(times)
        print(json.dumps({"family": family, "n": n, "n_gpus": world, "sa_bwt_ms": ms, "sa_bwt_MBps": n / ms / 1e3, "all_ms": [round(t, 1) for t in times],
                          "rounds": res.rounds, "owned_suffixes_per_rank": counts, "bad_rows": bad, "sentinel_ok": ok_sentinel, "bwt_sample_ok": ok_bwt,
                          "nvlink_bytes_stored_by_rank0": res.exchanged_bytes,
                          "phases_rank0_ms": {k: round(v["ms"] / reps, 1) for k, v in prof["phases"].items() if v["launches"]}}))
    dist.barrier()
    sorter.close(); eng.close()
finally:
    dist.destroy_process_group()
