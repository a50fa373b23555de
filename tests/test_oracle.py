"""CPU tier: pins the oracle (oracle/oracle.c) before anything is compared against it.

  (i)   against brute force (the ordering of the reference's validator, main.cpp:210-232),
  (ii)  against the unmodified reference built from /root/reference (oracle/_ref), when present,
  (iii) against tests/golden/kat.json, digests produced by that reference build
        (tests/golden/make_golden.py) — these travel to boxes where /root/reference does not exist.
"""
import json
import os

import numpy as np
import pytest

from cases import EDGE_SIZES, FAMILIES, gen, small_alphabet_exhaustive

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "kat.json")) as f:
    KAT = json.load(f)["cases"]


def test_bruteforce_exhaustive_small_alphabets(oracle):
    # every string over {0,1} up to length 10 and over {0,1,2} up to length 6
    for sigma, max_len in ((2, 10), (3, 6)):
        for x in small_alphabet_exhaustive(max_len, sigma):
            assert np.array_equal(oracle.sa(x), oracle.sa_bruteforce(x)), x.tolist()


@pytest.mark.parametrize("family", FAMILIES)
def test_oracle_vs_bruteforce_edge_sizes(oracle, family):
    for n in EDGE_SIZES:
        if family in ("zeros", "abcabca", "periodic7", "fib") and n > 300:
            continue  # brute force is O(n^2 log n) on periodic inputs
        x = gen(family, n)
        assert np.array_equal(oracle.sa(x), oracle.sa_bruteforce(x)), (family, n)


@pytest.mark.parametrize("family", FAMILIES)
def test_oracle_vs_reference_build(ref, family):
    for n in EDGE_SIZES + [65536 + 17, 300001]:
        x = gen(family, n)
        threads = 1 if family in ("zeros", "zero_tail") else 3
        want = ref.ref_sa(x, threads)
        assert np.array_equal(ref.sa(x), want), (family, n)
        rb, rs = ref.ref_bwt(x, threads)
        ob, os_ = ref.bwt(x)
        assert os_ == rs and np.array_equal(ob, rb), (family, n)
        assert np.array_equal(ref.unbwt(rb, rs), x), (family, n)
        assert np.array_equal(ref.ref_unbwt(rb, rs, 2), x), (family, n)


def test_reference_selftest_inputs(ref):
    """a sample of the reference's hidden self-test grid (main.cpp:389-435): rand()%sym inputs"""
    from msufsort_b200 import textgen
    for sym in (1, 2, 3, 7, 64, 255):
        for size in (1, 2, 31, 64, 500, 1023):
            x = textgen.reference_selftest(size, sym, sym * size)
            want = ref.ref_sa(x, 1)
            assert np.array_equal(ref.sa(x), want), (sym, size)
            assert ref.check_sa(x, want) == 0


@pytest.mark.parametrize("case", [c for c in KAT if c["n"] <= (1 << 22)], ids=lambda c: f"{c['family']}-{c['n']}")
def test_oracle_matches_golden(oracle, case):
    x = gen(case["family"], case["n"])
    assert f"{oracle.fnv(x):016x}" == case["text_fnv"], "generator drifted from the golden input"
    sa = oracle.sa(x)
    assert f"{oracle.fnv(sa):016x}" == case["sa_fnv"]
    bwt, s = oracle.bwt_from_sa(x, sa)
    assert s == case["bwt_sentinel"]
    assert f"{oracle.fnv(bwt):016x}" == case["bwt_fnv"]
    assert np.array_equal(oracle.unbwt(bwt, s), x)
    assert oracle.check_sa(x, sa) == 0


def test_checker_rejects_wrong_arrays(oracle):
    x = gen("markov3", 5000)
    sa = oracle.sa(x)
    assert oracle.check_sa(x, sa) == 0
    bad = sa.copy(); bad[[10, 11]] = bad[[11, 10]]
    assert oracle.check_sa(x, bad) > 0
    bad = sa.copy(); bad[5] = bad[6]
    assert oracle.check_sa(x, bad) > 0
    bad = sa.copy(); bad[0] = 0
    assert oracle.check_sa(x, bad) > 0
