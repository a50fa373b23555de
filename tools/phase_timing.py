#!/usr/bin/env python
"""needs a -DB200SA_PHASE_TIMING build: per-phase cycles of the scatter sweep (warp 0 of every 16th tile)"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msufsort_b200.api import Engine, torch_stream_handle
eng = Engine(0)
m = 1 << 28
g = torch.Generator(device="cuda"); g.manual_seed(1)
keys0 = torch.randint(-(1 << 62), 1 << 62, (m,), dtype=torch.int64, device="cuda", generator=g)
ka = torch.empty_like(keys0); va = torch.empty(m, dtype=torch.int32, device="cuda")
out = (C.c_ulonglong * 8)()
fn = eng.lib.cdll.b200sa_debug_phase_cycles
for it in range(2):
    k = keys0.clone()
    fn(out, 1)
    eng.radix_sort_pairs_dev(k, ka, None, va, m, 0, 64, torch_stream_handle())
fn(out, 0)
names = ["wait for keys", "ranking (16 rows)", "combine + scan + barriers", "staging", "look-back", "barrier after look-back", "write-out issue"]
n = out[7] or 1
tot = sum(out[i] for i in range(7))
print(f"sampled tiles {n}, mean cycles per tile (warp 0): {tot / n:.0f}")
for i, nm in enumerate(names):
    print(f"  {nm:28s} {out[i] / n:8.0f} cycles  {100 * out[i] / tot:5.1f} %")
