#!/usr/bin/env python
"""Generates tests/golden/kat_lcp.json from the UNMODIFIED reference: suffix array by the reference library, LCP array by
the reference demo's own construction (src/executable/msufsort/main.cpp:16-105, compiled into oracle/_ref by ref_shim.cpp),
mapped to this repository's convention (n+1 entries aligned with the SA, lcp[0] = lcp[1] = 0; the demo's output[i] is
lcp[i+2]).  Run in the build container only:

    make oracle && python tests/golden/make_golden_lcp.py

The demo's construction is O(n * LCP): repetitive families stay small here.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import Oracle  # noqa: E402
from cases import gen  # noqa: E402

CASES = [("rand", 1 << 20), ("rand", 1 << 22), ("markov3", 1 << 20), ("markov3", 1 << 22), ("acgt_rep", 1 << 20), ("acgt_rep", 1 << 22),
         ("sigma2", 1 << 20), ("sigma4", 1 << 20), ("zero_tail", 100003), ("fib", 1 << 14), ("periodic1009", 1 << 15),
         ("abcabca", 1 << 13), ("zeros", 1 << 12), ("rand", 2), ("rand", 3), ("rand", 17), ("rand", 257), ("rand", 4097)]


def main():
    o = Oracle()
    assert o.ref is not None, "build oracle/_ref first (make oracle)"
    out = []
    for family, n in CASES:
        x = gen(family, n)
        t0 = time.time()
        sa = o.ref_sa(x, 1)
        lcp = o.ref_lcp(x, sa, 1)
        e = {"family": family, "n": n, "text_fnv": f"{o.fnv(x):016x}", "lcp_fnv": f"{o.fnv(lcp):016x}", "lcp_max": int(lcp.max()),
             "lcp_sum": int(lcp.astype("int64").sum())}
        out.append(e)
        print(e, f"{time.time() - t0:.1f}s", flush=True)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat_lcp.json"), "w") as f:
        json.dump({"source": "unmodified reference: library SA + demo LCP (oracle/_ref/libmsufsort_ref.so)", "fnv": "FNV-1a-64 of the int32 array",
                   "convention": "n+1 entries, lcp[0] = lcp[1] = 0, lcp[r] = lcp(SA[r-1], SA[r])", "cases": out}, f, indent=1)


if __name__ == "__main__":
    main()
