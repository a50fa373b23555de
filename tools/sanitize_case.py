#!/usr/bin/env python
"""Small SA + BWT + inverse BWT through the host entry points (no torch) — the payload for compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from msufsort_b200.api import Engine
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
    print("ok", fam, n, flush=True)
eng.close()
