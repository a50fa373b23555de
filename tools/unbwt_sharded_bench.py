#!/usr/bin/env python
"""Inverse BWT of ONE text over N GPUs (walkers partitioned, psi replicated); run under torchrun.
usage: torchrun --nproc-per-node N tools/unbwt_sharded_bench.py [n] [steps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from msufsort_b200.api import torch_stream_handle, Engine
from msufsort_b200.sharded import ShardedSorter
from msufsort_b200 import textgen

n = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 28)
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
eng = Engine(lr)
stream = torch_stream_handle()
d_text = torch.from_numpy(textgen.markov3(n)).cuda()
d_bwt = torch.empty(n, dtype=torch.uint8, device="cuda")
s = eng.bwt_dev(d_text, n, d_bwt, None, stream)
eng.release_workspace()
sorter = ShardedSorter(eng)
out = sorter.inverse_bwt(d_bwt, s)
assert torch.equal(out, d_text), "sharded inverse BWT mismatch"
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    out = sorter.inverse_bwt(d_bwt, s)
e1.record(); dist.barrier(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(t.item())
    print(json.dumps({"metric": "unbwt_input_throughput", "value": n / ms / 1e3, "unit": "MB/s", "n_gpus": world, "n_bytes": n,
                      "ms_per_step": ms, "scaling": "strong", "parallelism": "psi replicated, walkers partitioned, sum all-reduce of the output"}))
dist.destroy_process_group()
