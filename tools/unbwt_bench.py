#!/usr/bin/env python
"""Inverse-BWT benchmark (BASELINE.json configs[3]): BWT of the Markov text -> text, resident in HBM.
usage: python tools/unbwt_bench.py [n] [steps]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from msufsort_b200.api import torch_stream_handle, Engine
from msufsort_b200 import textgen

n = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 28)
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
eng = Engine(0)
stream = torch_stream_handle()
x = textgen.markov3(n)
d_text = torch.from_numpy(x).cuda()
d_bwt = torch.empty(n, dtype=torch.uint8, device="cuda")
t0 = time.time(); s = eng.bwt_dev(d_text, n, d_bwt, None, stream); torch.cuda.synchronize()
print(f"forward: n={n} sentinel={s} {time.time()-t0:.3f}s (first call, includes workspace allocation)")
eng.release_workspace()
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
eng.unbwt_dev(d_bwt, n, s, d_out, stream)
assert torch.equal(d_out, d_text), "inverse BWT mismatch"
eng.profile_reset(); eng.set_profiling(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(steps):
    eng.unbwt_dev(d_bwt, n, s, d_out, stream)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
p = eng.profile()
print(json.dumps({"metric": "unbwt_input_throughput", "value": n / ms / 1e3, "unit": "MB/s", "n_bytes": n, "ms_per_step": ms,
                  "phases": {k: round(v["ms"] / steps, 3) for k, v in p["phases"].items() if v["launches"]}}))
