#!/usr/bin/env python
"""LCP, batched forward / inverse transforms on small inputs through the host entry points — compute-sanitizer payload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from msufsort_b200.api import Engine
from msufsort_b200 import textgen
from conftest import Oracle
o = Oracle()
eng = Engine(0)
for fam, n in [("markov3", 40000), ("zeros", 9000), ("fib", 28657), ("abcabca", 12001), ("rand", 4097)]:
    x = textgen.GENERATORS[fam](n + 1)[1:]          # unaligned text pointer on the host side is irrelevant; staging aligns it
    lcp, sa = eng.make_lcp_array(x, return_sa=True)
    assert np.array_equal(lcp, o.lcp(x, sa, kasai=True)), fam
    print("ok lcp", fam, n, flush=True)
blocks = [textgen.GENERATORS[f](m) for f, m in [("markov3", 20000), ("zeros", 3000), ("rand", 17), ("fib", 4097), ("acgt_rep", 9000), ("rand", 1)]]
blocks.insert(2, np.empty(0, np.uint8))
sas = eng.suffix_array_batch(blocks)
bw, sent = eng.bwt_batch(blocks)
for b, x in enumerate(blocks):
    if x.size:
        want = o.sa(x)
        assert np.array_equal(sas[b], want), b
        wb, ws = o.bwt_from_sa(x, want)
        assert ws == sent[b] and np.array_equal(bw[b], wb), b
back = eng.unbwt_batch(bw, sent)
for b, x in enumerate(blocks):
    assert np.array_equal(back[b], x), b
print("ok batch", len(blocks), flush=True)
eng.close()
