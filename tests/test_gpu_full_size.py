"""GPU tier at BASELINE.json's full sizes.  Where the reference itself finishes in seconds on the box's host cores
(256 MiB Markov text: SA + BWT; 2^30-2 Markov text: BWT) the results are compared with it bit for bit; beyond that
correctness is judged by size-independent properties: the O(n) GPU validator (SA[0]=n, permutation, order of neighbouring
rows via the ISA — which, the SA being unique, proves bit-exactness with the reference), the sentinel index being the row
of suffix 0, and the BWT -> inverse BWT round trip reproducing the text byte for byte."""
import numpy as np
import pytest

from cases import gen

pytestmark = pytest.mark.gpu


def _free_hbm_gb() -> float:
    import torch
    free, _ = torch.cuda.mem_get_info()
    return free / 2**30


def _run(gpu_engine, family, n):
    import torch
    x = gen(family, n)
    d_text = torch.from_numpy(x).cuda()
    d_sa = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    d_bwt = torch.empty(n, dtype=torch.uint8, device="cuda")
    s = gpu_engine.bwt_dev(d_text, n, d_bwt, d_sa)
    assert gpu_engine.check_suffix_array_dev(d_text, n, d_sa) == 0
    assert int(d_sa[s]) == 0 and int(d_sa[0]) == n                     # sentinel index = row of suffix 0
    # spot-check the BWT definition on a sample of rows
    rows = torch.randint(1, n + 1, (4096,), device="cuda")
    rows = rows[rows != s]
    out_idx = rows - (rows > s).long()
    assert bool((d_bwt[out_idx] == d_text[(d_sa[rows].long() - 1)]).all())
    del d_sa
    gpu_engine.release_workspace()
    d_back = torch.empty(n, dtype=torch.uint8, device="cuda")
    gpu_engine.unbwt_dev(d_bwt, n, s, d_back)
    assert bool(torch.equal(d_back, d_text))
    gpu_engine.release_workspace()


def test_config2_markov_256mib(gpu_engine):
    _run(gpu_engine, "markov3", 1 << 28)


def test_config2_markov_256mib_bit_exact_vs_reference(gpu_engine, oracle):
    """BASELINE.json configs[1] against the UNMODIFIED reference build (oracle/_ref): suffix array and BWT bit for bit"""
    import torch
    if oracle.ref is None:
        pytest.skip("oracle/_ref was not built (no reference tree at build time)")
    n = 1 << 28
    x = gen("markov3", n)
    d_text = torch.from_numpy(x).cuda()
    d_sa = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    d_bwt = torch.empty(n, dtype=torch.uint8, device="cuda")
    s = gpu_engine.bwt_dev(d_text, n, d_bwt, d_sa)
    gpu_engine.release_workspace()
    import os
    threads = os.cpu_count() or 1
    want_sa = oracle.ref_sa(x, threads)
    assert bool(torch.equal(d_sa.cpu(), torch.from_numpy(want_sa)))
    del d_sa, want_sa
    want_bwt, want_s = oracle.ref_bwt(x, threads)
    assert s == want_s and bool(torch.equal(d_bwt.cpu(), torch.from_numpy(want_bwt)))


def test_config4_unbwt_of_the_references_bwt_of_the_markov_text_1gib(gpu_engine, oracle):
    """BASELINE.json configs[3] literally: the inverse BWT of the BWT the REFERENCE produced for the 2^30-2 byte Markov
    text gives the text back; the reference's BWT also equals ours bit for bit"""
    import torch
    if oracle.ref is None:
        pytest.skip("oracle/_ref was not built (no reference tree at build time)")
    n = (1 << 30) - 2
    x = gen("markov3", n)
    import os
    ref_bwt, ref_s = oracle.ref_bwt(x, os.cpu_count() or 1)
    d_bwt = torch.from_numpy(ref_bwt).cuda()
    d_back = torch.empty(n, dtype=torch.uint8, device="cuda")
    gpu_engine.unbwt_dev(d_bwt, n, ref_s, d_back)
    gpu_engine.release_workspace()
    d_text = torch.from_numpy(x).cuda()
    assert bool(torch.equal(d_back, d_text))
    del d_back
    d_ours = torch.empty(n, dtype=torch.uint8, device="cuda")
    s = gpu_engine.bwt_dev(d_text, n, d_ours, None)
    gpu_engine.release_workspace()
    assert s == ref_s and bool(torch.equal(d_ours, d_bwt))


def test_config3_and_4_acgt_repeats_1gib(gpu_engine):
    _run(gpu_engine, "acgt_rep", (1 << 30) - 2)          # the largest n the reference itself handles correctly


def test_config5_periodic_2gib(gpu_engine):
    """2 GiB deep-doubling run (≈130 GB of HBM, ≈10 s)"""
    if _free_hbm_gb() < 150:
        pytest.skip("needs 150 GB of free HBM")
    _run(gpu_engine, "periodic7", (1 << 31) - 2)


def test_config5_periodic_exactly_2gib_wide_index(gpu_engine):
    """n = 2^31 + 4099: beyond every int32 suffix index — the uint32 entry points (SA + BWT + inverse BWT), judged by the
    O(n) validator and the round trip"""
    import torch
    if _free_hbm_gb() < 150:
        pytest.skip("needs 150 GB of free HBM")
    n = (1 << 31) + 4099
    x = gen("periodic1009", n)
    d_text = torch.from_numpy(x).cuda()
    d_sa = torch.empty(n + 1, dtype=torch.int32, device="cuda")      # uint32 payload in an int32 tensor
    d_bwt = torch.empty(n, dtype=torch.uint8, device="cuda")
    s = gpu_engine.bwt_u32_dev(d_text, n, d_bwt, d_sa)
    assert gpu_engine.check_suffix_array_u32_dev(d_text, n, d_sa) == 0
    gpu_engine.release_workspace()
    as_u32 = lambda t: t.long() & 0xffffffff
    assert int(as_u32(d_sa[s])) == 0 and int(as_u32(d_sa[0])) == n
    rows = torch.randint(1, n + 1, (4096,), device="cuda")
    rows = torch.cat([rows[rows != s], torch.tensor([n, n - 1, 1], device="cuda")])
    rows = rows[rows != s]
    out_idx = rows - (rows > s).long()
    assert bool((d_bwt[out_idx] == d_text[as_u32(d_sa[rows]) - 1]).all())
    assert int((as_u32(d_sa) >= (1 << 31)).sum()) == n - (1 << 31) + 1   # every suffix start >= 2^31 appears exactly once
    # wide inverse BWT (int64 sentinel index, rows beyond 2^31 in the psi table)
    del d_sa, rows, out_idx
    torch.cuda.empty_cache()
    d_back = torch.empty(n, dtype=torch.uint8, device="cuda")
    gpu_engine.unbwt_u32_dev(d_bwt, n, s, d_back)
    gpu_engine.release_workspace()
    assert bool(torch.equal(d_back, d_text))
