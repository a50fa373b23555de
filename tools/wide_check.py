#!/usr/bin/env python
"""Wide-index run at n > 2^31 on one GPU: SA (uint32) + BWT, O(n) validator, timings.  usage: wide_check.py [family] [n]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msufsort_b200.api import Engine
from msufsort_b200 import textgen

family = sys.argv[1] if len(sys.argv) > 1 else "periodic1009"
n = int(sys.argv[2]) if len(sys.argv) > 2 else (1 << 31) + 4099
eng = Engine(0)
x = textgen.GENERATORS[family](n)
d_text = torch.from_numpy(x).cuda()
d_sa = torch.empty(n + 1, dtype=torch.int32, device="cuda")
d_bwt = torch.empty(n, dtype=torch.uint8, device="cuda")
eng.set_profiling(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
s = eng.bwt_u32_dev(d_text, n, d_bwt, d_sa)
e1.record(); torch.cuda.synchronize()
p = eng.profile()
bad = eng.check_suffix_array_u32_dev(d_text, n, d_sa)
eng.release_workspace()  # the index arithmetic below needs 8 bytes per row of scratch
u = lambda t: t.long() & 0xffffffff
rows = torch.randint(1, n + 1, (65536,), device="cuda"); rows = rows[rows != s]
ok_bwt = bool((d_bwt[rows - (rows > s).long()] == d_text[u(d_sa[rows]) - 1]).all())
print(json.dumps({"family": family, "n": n, "sa_bwt_ms_cold": e0.elapsed_time(e1), "rounds": p["rounds"], "sweeps": p["sort_passes"], "bad_rows": bad,
                  "sentinel": s, "sa0": int(u(d_sa[0])), "sa_at_sentinel": int(u(d_sa[s])), "bwt_sample_ok": ok_bwt,
                  "suffixes_ge_2_31": int((u(d_sa) >= (1 << 31)).sum()), "mem_GB": torch.cuda.max_memory_allocated() / 1e9}))
