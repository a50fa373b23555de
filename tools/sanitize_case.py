#!/usr/bin/env python
"""Small runs of every path through the host entry points (no torch) — the payload for compute-sanitizer:
SA + BWT + inverse BWT (one context), the resident-suffix-array reuse, corrupted inverse-BWT input, LCP, batches, and one
text sharded over three contexts on one device (group API: peer-pointer ISA, inbox sends, bucket-major apply, sharded
inverse BWT)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from msufsort_b200.api import B200SAError, Engine, Group
from msufsort_b200 import textgen
from conftest import Oracle
o = Oracle()
eng = Engine(0)
for fam, n in [("markov3", 60000), ("abcabca", 20000), ("zeros", 9000), ("fib", 30000), ("rand", 4097)]:
    x = textgen.GENERATORS[fam](n)
    sa, bwt, s = eng.suffix_array_and_bwt(x)
    want = o.sa(x)
    assert np.array_equal(sa, want), fam
    b = bwt.copy(); eng.reverse_burrows_wheeler_transform(b, s)
    assert np.array_equal(b, x), fam
    # the drop-in sequence: SA, then BWT of the same bytes (reuses the resident sort)
    assert np.array_equal(eng.make_suffix_array(x), want)
    b2 = x.copy(); assert eng.forward_burrows_wheeler_transform(b2) == s and np.array_equal(b2, bwt)
    # corrupted input must be rejected without touching memory outside the buffers
    bad = bwt.copy(); bad[n // 2] ^= 1
    try:
        eng.reverse_burrows_wheeler_transform(bad, s)
    except B200SAError:
        pass
    assert np.array_equal(eng.make_lcp_array(x, sa), o.lcp(x, want, kasai=True)), fam
    print("ok", fam, n, flush=True)
blocks = [textgen.GENERATORS["markov3"](30000), textgen.GENERATORS["rand"](257), textgen.GENERATORS["zeros"](999)]
bw, sent = eng.bwt_batch(blocks)
for got, blk in zip(eng.unbwt_batch(bw, sent), blocks):
    assert np.array_equal(got, blk)
print("ok batch", flush=True)
eng.close()
g = Group([0, 0, 0])
for fam, n in [("markov3", 60000), ("abcabca", 20000), ("fib", 30000)]:
    x = textgen.GENERATORS[fam](n)
    sa, bwt, s = g.suffix_array_and_bwt(x)
    assert np.array_equal(sa, o.sa(x)), fam
    b = bwt.copy(); g.reverse_burrows_wheeler_transform(b, s)
    assert np.array_equal(b, x), fam
    print("ok group", fam, n, flush=True)
g.close()
