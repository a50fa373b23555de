#!/usr/bin/env python
"""Radix-sort microbenchmark: (u64 key, u32 value) pairs resident in HBM, per-sweep time and HBM fraction.
usage: python tools/sort_bench.py [log2_m] [bits] [kind]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msufsort_b200.api import torch_stream_handle, Engine

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 28
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 64
kind = sys.argv[3] if len(sys.argv) > 3 else "random"
m = 1 << lg
eng = Engine(0)
g = torch.Generator(device="cuda"); g.manual_seed(1)
keys0 = torch.randint(-(1 << 62), 1 << 62, (m,), dtype=torch.int64, device="cuda", generator=g)
if bits < 64:
    keys0 &= (1 << bits) - 1
if kind == "skew":      # few distinct values per digit
    keys0 &= 0x0303030303030303
ka = torch.empty_like(keys0); v = torch.empty(m, dtype=torch.int32, device="cuda"); va = torch.empty_like(v)
stream = torch_stream_handle()
peak = 6547.2
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
for it in range(4):
    k = keys0.clone()
    eng.profile_reset(); eng.set_profiling(True)
    eng.radix_sort_pairs_dev(k, ka, None, va, m, 0, bits, stream)
    p = eng.profile()
sp, sh = p["phases"]["sort_pass"], p["phases"]["sort_hist"]
gbs = sp["alg_bytes"] / sp["ms"] / 1e6
print(f"m=2^{lg} bits={bits} {kind}: sweeps {sp['launches']} avg {sp['ms']/sp['launches']:.3f} ms  {gbs:.0f} GB/s = {gbs/peak:.3f} of measured peak; hist {sh['ms']:.3f} ms")
