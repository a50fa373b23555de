#!/usr/bin/env python
"""Large single-GPU configurations (BASELINE.json configs[2] and [4]): SA + BWT + inverse BWT, judged by the
O(n) GPU validator (no CPU oracle at these sizes) and the round trip.  usage: big_check.py family n"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from msufsort_b200.api import torch_stream_handle, Engine
from msufsort_b200 import textgen

family = sys.argv[1]; n = int(sys.argv[2])
eng = Engine(0)
stream = torch_stream_handle()
t0 = time.time(); x = textgen.GENERATORS[family](n); tg = time.time() - t0
d_text = torch.from_numpy(x).cuda()
d_sa = torch.empty(n + 1, dtype=torch.int32, device="cuda")
d_bwt = torch.empty(n, dtype=torch.uint8, device="cuda")
eng.bwt_dev(d_text, n, d_bwt, d_sa, stream)            # warm-up: allocates the workspace
eng.profile_reset(); eng.set_profiling(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
s = eng.bwt_dev(d_text, n, d_bwt, d_sa, stream)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
p = eng.profile(); eng.set_profiling(False)
bad = eng.check_suffix_array_dev(d_text, n, d_sa, stream)
del d_sa
eng.release_workspace()
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
eng.unbwt_dev(d_bwt, n, s, d_out, stream)
torch.cuda.synchronize(); u0.record()
eng.unbwt_dev(d_bwt, n, s, d_out, stream)
u1.record(); torch.cuda.synchronize()
ok = bool(torch.equal(d_out, d_text))
sp = p["phases"]["sort_pass"]
print(json.dumps({"family": family, "n": n, "sa_bwt_ms": ms, "sa_bwt_MBps": n / ms / 1e3, "rounds": p["rounds"], "sweeps": p["sort_passes"],
                  "bad_rows": bad, "unbwt_ms": u0.elapsed_time(u1), "unbwt_MBps": n / u0.elapsed_time(u1) / 1e3, "roundtrip_ok": ok,
                  "sweep_GBps": sp["alg_bytes"] / sp["ms"] / 1e6, "gen_s": round(tg, 1), "mem_GB": torch.cuda.max_memory_allocated() / 1e9,
                  "phases_ms": {k: round(v["ms"], 1) for k, v in p["phases"].items() if v["launches"]}}))
