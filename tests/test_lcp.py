"""LCP array: oracle pinned against the reference demo's own LCP code, kernel logic
under the emulator (CPU tier) and bit-exact parity on the GPU through the C ABI (GPU tier)."""
import os

import numpy as np
import pytest

from cases import EDGE_SIZES, FAMILIES, gen

LCP_FAMILIES = ["rand", "markov3", "acgt_rep", "periodic7", "periodic1009", "fib", "zeros", "abcabca", "sigma2", "zero_tail"]


# ---- oracle -----------------------------------------------------------------------------------
@pytest.mark.parametrize("family", FAMILIES)
def test_oracle_lcp_pinned_to_reference_demo(ref, family):
    """restated recursion == Kasai == the unmodified main.cpp:16-105 compiled into oracle/_ref"""
    for n in [1, 2, 3, 5, 17, 257, 4097, 30011]:
        x = gen(family, n)
        sa = ref.sa(x)
        a = ref.lcp(x, sa)
        assert np.array_equal(a, ref.lcp(x, sa, kasai=True)), (family, n)
        assert np.array_equal(a, ref.ref_lcp(x, sa)), (family, n)
        if n >= 4:
            assert np.array_equal(a, ref.ref_lcp(x, sa, threads=3)), (family, n)


def test_oracle_lcp_bruteforce_small(oracle):
    rng = np.random.default_rng(5)
    for n in range(1, 40):
        x = rng.integers(0, 3, size=n, dtype=np.uint8)
        sa = oracle.sa(x)
        lcp = oracle.lcp(x, sa)
        assert lcp[0] == 0 and lcp[1] == 0
        for r in range(2, n + 1):
            a, b = bytes(x[sa[r - 1]:]), bytes(x[sa[r]:])
            l = 0
            while l < min(len(a), len(b)) and a[l] == b[l]:
                l += 1
            assert lcp[r] == l


def _golden():
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kat_lcp.json")) as f:
        return json.load(f)["cases"]


def test_oracle_lcp_matches_golden_vectors(oracle):
    """digests produced by the reference (library SA + demo LCP) with tests/golden/make_golden_lcp.py; they travel to the
    GPU box, where /root/reference does not exist"""
    for c in _golden():
        x = gen(c["family"], c["n"])
        assert f"{oracle.fnv(x):016x}" == c["text_fnv"]
        sa = oracle.sa(x)
        for kasai in (False, True):
            if not kasai and c["lcp_max"] > 1000 and c["n"] > 20000:
                continue
            lcp = oracle.lcp(x, sa, kasai=kasai)
            assert f"{oracle.fnv(lcp):016x}" == c["lcp_fnv"], (c["family"], c["n"], kasai)
            assert int(lcp.max()) == c["lcp_max"] and int(lcp.astype(np.int64).sum()) == c["lcp_sum"]


# ---- emulator ---------------------------------------------------------------------------------
@pytest.mark.parametrize("family", LCP_FAMILIES)
def test_emu_lcp(emu_engine, oracle, family):
    for n in EDGE_SIZES + [20011]:
        x = gen(family, n)
        sa = oracle.sa(x)
        assert np.array_equal(emu_engine.make_lcp_array(x, sa), oracle.lcp(x, sa, kasai=True)), (family, n)


def test_emu_lcp_long_matches_and_unaligned_text(emu_engine, oracle):
    """matches far beyond the per-thread budget (CTA compare), text pointer at every alignment mod 4"""
    for family, n in [("zeros", 20001), ("periodic7", 15000), ("fib", 17711), ("abcabca", 12001)]:
        buf = gen(family, n + 3)
        for shift in range(4):
            x = buf[shift:shift + n]
            sa = oracle.sa(x)
            lcp, sa2 = emu_engine.make_lcp_array(x, return_sa=True)
            assert np.array_equal(sa2, sa)
            assert np.array_equal(lcp, oracle.lcp(x, sa, kasai=True)), (family, shift)
            # device-pointer entry point straight on the (unaligned) view
            out = np.empty(n + 1, dtype=np.int32)
            emu_engine.lcp_dev(x, n, sa, out)
            assert np.array_equal(out, lcp), (family, shift)


def test_emu_lcp_bucketed_phi(oracle, monkeypatch):
    """phi scattered through the bucketed (radix sweep + L2-window) path"""
    from conftest import ROOT
    from msufsort_b200.api import Engine, Library
    monkeypatch.setenv("B200SA_ISA_DIRECT_BYTES", "0")
    monkeypatch.setenv("B200SA_ISA_MIN_UPDATES", "1")
    eng = Engine(0, library=Library(os.path.join(ROOT, "tests", "emu", "libb200sa_emu.so")))
    try:
        for family, n in [("markov3", 70011), ("acgt_rep", 40000), ("zeros", 35000), ("fib", 46368), ("rand", 1)]:
            x = gen(family, n)
            sa = oracle.sa(x)
            assert np.array_equal(eng.make_lcp_array(x, sa), oracle.lcp(x, sa, kasai=True)), (family, n)
    finally:
        eng.close()


def test_emu_lcp_direct_route(oracle, monkeypatch):
    """B200SA_LCP_DIRECT=1: budgeted row-wise comparison; few long rows are finished by the CTA-wide compare, many of them
    send the call down the PLCP route"""
    from conftest import ROOT
    from msufsort_b200.api import Engine, Library
    monkeypatch.setenv("B200SA_LCP_DIRECT", "1")
    eng = Engine(0, library=Library(os.path.join(ROOT, "tests", "emu", "libb200sa_emu.so")))
    try:
        for family, n in [("markov3", 50011), ("acgt_rep", 60000), ("rand", 3000), ("zeros", 9000), ("fib", 17711), ("rand", 1), ("rand", 2)]:
            buf = gen(family, n + 1)
            x = buf[1:]
            sa = oracle.sa(x)
            before = eng.launch_count()
            assert np.array_equal(eng.make_lcp_array(x, sa), oracle.lcp(x, sa, kasai=True)), (family, n)
            launches = eng.launch_count() - before
            if family in ("markov3", "rand") and n > 2:
                assert launches <= 4, launches          # SA validation (2) + direct pass (+ finish), no PLCP levels
            if family in ("zeros", "fib"):
                assert launches > 10, launches          # fell through to the PLCP route
        # long matches in a few rows only: one repeated 5000-byte segment inside random text
        rng = np.random.default_rng(9)
        x = rng.integers(0, 256, size=80000, dtype=np.uint8)
        x[60000:61000] = x[1000:2000]
        sa = oracle.sa(x)
        before = eng.launch_count()
        assert np.array_equal(eng.make_lcp_array(x, sa), oracle.lcp(x, sa, kasai=True))
        assert eng.launch_count() - before == 4         # validator (2) + direct pass + CTA-wide finish
    finally:
        eng.close()


def test_emu_lcp_empty_and_errors(emu_engine):
    from msufsort_b200.api import B200SAError
    assert emu_engine.make_lcp_array(np.empty(0, dtype=np.uint8)).tolist() == [0]
    with pytest.raises(ValueError):
        emu_engine.make_lcp_array(np.zeros(4, dtype=np.uint8), np.zeros(3, dtype=np.int32))
    with pytest.raises(B200SAError):
        emu_engine.lcp_dev(None, 5, None, None)


# ---- GPU --------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("family", FAMILIES)
def test_gpu_lcp_parity(gpu_engine, oracle, family):
    for n in EDGE_SIZES + [65536, 1 << 20]:
        x = gen(family, n)
        lcp, sa = gpu_engine.make_lcp_array(x, return_sa=True)
        want_sa = oracle.sa(x)
        assert np.array_equal(sa, want_sa), (family, n)
        assert np.array_equal(lcp, oracle.lcp(x, want_sa, kasai=True)), (family, n)


@pytest.mark.gpu
def test_gpu_lcp_matches_golden_vectors(gpu_engine, oracle):
    for c in _golden():
        x = gen(c["family"], c["n"])
        lcp = gpu_engine.make_lcp_array(x)
        assert f"{oracle.fnv(lcp):016x}" == c["lcp_fnv"], (c["family"], c["n"])


@pytest.mark.gpu
def test_gpu_lcp_16MiB_device_resident(gpu_engine, oracle):
    """device-resident entry point at 16 MiB (above the direct-scatter threshold: bucketed phi), unaligned text"""
    import torch
    n = 1 << 24
    for family in ["markov3", "acgt_rep"]:
        buf = gen(family, n + 1)
        x = buf[1:]
        d_buf = torch.from_numpy(buf).cuda()
        d_text = d_buf[1:]
        d_sa = torch.empty(n + 1, dtype=torch.int32, device="cuda")
        d_lcp = torch.empty(n + 1, dtype=torch.int32, device="cuda")
        gpu_engine.suffix_array_dev(d_text, n, d_sa)
        gpu_engine.lcp_dev(d_text, n, d_sa, d_lcp)
        torch.cuda.synchronize()
        sa = d_sa.cpu().numpy()
        want_sa = oracle.sa(x)
        assert np.array_equal(sa, want_sa)
        assert np.array_equal(d_lcp.cpu().numpy(), oracle.lcp(x, want_sa, kasai=True)), family


@pytest.mark.gpu
def test_gpu_lcp_deep_repeats(gpu_engine, oracle):
    """lcp ~ n inputs: the CTA-wide compare keeps these at a few streaming passes"""
    for family, n in [("zeros", 1 << 22), ("fib", 1 << 22), ("periodic1009", 1 << 23), ("abcabca", 1 << 22)]:
        x = gen(family, n)
        lcp, sa = gpu_engine.make_lcp_array(x, return_sa=True)
        assert np.array_equal(lcp, oracle.lcp(x, sa, kasai=True)), family
