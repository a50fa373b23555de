#!/usr/bin/env python
"""Generates tests/golden/kat.json by running the UNMODIFIED reference (oracle/_ref, built from
/root/reference by oracle/Makefile) on closed-form inputs.  Run in the build container only:

    make oracle && python tests/golden/make_golden.py

Each entry: generator name + n (inputs are reproducible from msufsort_b200.textgen), FNV-1a-64 of
the little-endian int32 suffix array (all n+1 entries), the BWT's sentinel index and FNV-1a-64 of
its n bytes.  Includes the inputs of SURVEY.md §4's known-answer table (re-derived here).
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

CASES = [
    # (family, n, threads)  — runs of one byte need threads=1 (SURVEY.md §8c hazard 3)
    ("zeros", 1 << 20, 1),
    ("abcabca", 1 << 20, 1),
    ("abcabca", 1 << 24, 1),
    ("fib", 1 << 20, 8),
    ("fib", 1 << 21, 8),
    ("fib", 1 << 22, 8),
    ("rand", 1 << 20, 8),
    ("rand", 1 << 24, 8),
    ("markov3", 1 << 20, 8),
    ("markov3", 1 << 24, 8),
    ("acgt_rep", 1 << 20, 8),
    ("acgt_rep", 1 << 24, 8),
    ("periodic7", 1 << 20, 1),
    ("periodic1009", 1 << 22, 1),
    ("sigma2", 1 << 20, 8),
    ("sigma3", 1 << 18, 8),
    ("sigma4", 1 << 22, 8),
    ("zero_tail", 100003, 1),
    ("rand", 1, 1), ("rand", 2, 1), ("rand", 3, 1), ("rand", 17, 1), ("rand", 257, 1), ("rand", 4097, 1),
]


def main():
    o = Oracle()
    assert o.ref is not None, "build oracle/_ref first (make oracle)"
    out = []
    for family, n, threads in CASES:
        x = gen(family, n)
        t0 = time.time()
        sa = o.ref_sa(x, threads)
        bwt, s = o.ref_bwt(x, threads)
        back = o.ref_unbwt(bwt, s, max(1, threads))
        assert (back == x).all()
        e = {"family": family, "n": n, "text_fnv": f"{o.fnv(x):016x}", "sa_fnv": f"{o.fnv(sa):016x}",
             "bwt_sentinel": s, "bwt_fnv": f"{o.fnv(bwt):016x}"}
        out.append(e)
        print(e, f"{time.time() - t0:.1f}s", flush=True)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat.json"), "w") as f:
        json.dump({"source": "unmodified reference (oracle/_ref/libmsufsort_ref.so)", "fnv": "FNV-1a-64", "cases": out}, f, indent=1)


if __name__ == "__main__":
    main()
