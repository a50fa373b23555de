#!/usr/bin/env python
"""debug driver: sharded SA+BWT of one generated text under torchrun, verbose errors. usage: shard_debug.py family n isa"""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch, torch.distributed as dist
from msufsort_b200.api import torch_stream_handle, Engine
from msufsort_b200.sharded import ShardedSorter
from msufsort_b200 import textgen
family, n, isa = sys.argv[1], int(sys.argv[2]), sys.argv[3]
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
try:
    eng = Engine(lr)
    x = textgen.GENERATORS[family](n)
    d_text = torch.from_numpy(x).cuda()
    sorter = ShardedSorter(eng, isa=isa)
    res = sorter.suffix_array_bwt(d_text)
    full = sorter.gather_sa(res)
    bad = eng.check_suffix_array_dev(d_text, n, full, torch_stream_handle())
    print(f"[rank {rank}] {family} n={n} isa={isa}: rounds={res.rounds} counts={res.counts} bad_rows={bad} rx={res.exchanged_bytes}", flush=True)
except Exception:
    print(f"[rank {rank}] EXCEPTION", flush=True)
    traceback.print_exc()
    sys.stdout.flush(); sys.stderr.flush()
    os._exit(1)
dist.destroy_process_group()
