"""One SA + one LCP call at 256 MiB (for an ncu launch list)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from msufsort_b200 import textgen as t
from msufsort_b200.api import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
eng = Engine(0)
d_text = torch.from_numpy(t.markov3(n, t.SEED_MARKOV)).cuda()
d_sa = torch.empty(n + 1, dtype=torch.int32, device="cuda")
d_lcp = torch.empty(n + 1, dtype=torch.int32, device="cuda")
eng.suffix_array_dev(d_text, n, d_sa)
eng.lcp_dev(d_text, n, d_sa, d_lcp)
torch.cuda.synchronize()
print("lcp max", int(d_lcp.max().item()))
