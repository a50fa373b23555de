"""Wide-index entry points: uint32 suffix arrays / int64 sentinel rows for texts beyond the
int32 limit.  CPU tier: same results as the int32 calls at small n through the emulator, descriptor packing for counts
>= 2^31, argument limits.  GPU tier: parity at small n; the 2^31-byte text runs in tests/test_gpu_full_size.py."""
import ctypes as C

import numpy as np
import pytest

from cases import EDGE_SIZES, gen


def _lcp_u32(eng, x, sa):
    """b200sa_lcp_u32_dev on host arrays (emulator) or device tensors (GPU)"""
    n = x.size
    if eng.lib.path.endswith("libb200sa_emu.so"):
        out = np.empty(n + 1, dtype=np.uint32)
        eng.lcp_u32_dev(x, n, sa, out)
        return out
    import torch
    d_x = torch.from_numpy(x).cuda()
    d_sa = torch.from_numpy(sa.view(np.int32)).cuda()
    d_out = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    eng.lcp_u32_dev(d_x, n, d_sa, d_out)
    return d_out.cpu().numpy().view(np.uint32)


def _check_small(eng, oracle):
    for family in ["markov3", "zeros", "fib", "rand", "abcabca"]:
        for n in EDGE_SIZES[:12] + [4097, 30011]:
            x = gen(family, n)
            sa, bwt, s = eng.suffix_array_and_bwt_u32(x)
            want = oracle.sa(x)
            assert sa.dtype == np.uint32 and np.array_equal(sa.astype(np.int64), want.astype(np.int64)), (family, n)
            wb, ws = oracle.bwt_from_sa(x, want)
            assert s == ws and np.array_equal(bwt, wb), (family, n)
            # wide inverse transform (int64 sentinel index) and wide LCP array (uint32 entries)
            back = bwt.copy()
            eng.reverse_burrows_wheeler_transform_u32(back, s)
            assert np.array_equal(back, x), (family, n)
            lcp = _lcp_u32(eng, x, sa)
            assert np.array_equal(lcp.astype(np.int64), oracle.lcp(x, want, kasai=True).astype(np.int64)), (family, n)
    sa, bwt, s = eng.suffix_array_and_bwt_u32(np.empty(0, np.uint8))
    assert sa.tolist() == [0] and s == 0


def test_emu_wide_matches_int32_calls(emu_engine, oracle):
    _check_small(emu_engine, oracle)
    x = gen("markov3", 20000)
    sa = oracle.sa(x).astype(np.uint32)
    out = np.empty(x.size + 1, dtype=np.uint32)
    emu_engine.suffix_array_u32_dev(x, x.size, out)
    assert np.array_equal(out, sa)
    assert emu_engine.check_suffix_array_u32_dev(x, x.size, out) == 0
    out[[5, 6]] = out[[6, 5]]
    assert emu_engine.check_suffix_array_u32_dev(x, x.size, out) > 0
    bwt = np.empty(x.size, dtype=np.uint8)
    s = emu_engine.bwt_u32_dev(x, x.size, bwt)
    assert (bwt.tolist(), s) == (oracle.bwt(x)[0].tolist(), oracle.bwt(x)[1])


def test_rerank_descriptors_hold_counts_beyond_2_31(emu_engine):
    """the kept count of a rerank tile prefix may need all 32 bits (n up to 2^32 - 8194)"""
    out = (C.c_uint64 * 5)()
    cases = [(0, 0, 0), (4096, 2048, 4096), (0x7fffffff, 0x3fffffff, 0x7fffffff), (0x80000000, 0x40000000, 0x80000001),
             (0xffffdffe, 0x7fffefff, 0xffffdffe), (0x80000001, 0, 0xffffffff), (0xffffffff, 0x7fffffff, 0xffffffff)]
    for kept, kheads, lh in cases:
        assert emu_engine.lib.cdll.b200sa_debug_rerank_descriptor(kept, kheads, lh, out) == 0
        assert list(out) == [kept, kheads, lh, 2, 2], (kept, kheads, lh, list(out))


def test_wide_limits(emu_engine):
    from msufsort_b200.api import B200SAError
    x = np.zeros(8, dtype=np.uint8)
    sa = np.zeros(9, dtype=np.uint32)
    with pytest.raises(B200SAError):
        emu_engine.suffix_array_u32_dev(x, (1 << 32) - 8193, sa)      # above B200SA_MAX_N_UINT32
    with pytest.raises(B200SAError):
        emu_engine.suffix_array_dev(x, (1 << 31) - 1, sa)             # the int32 entry point keeps its limit
    with pytest.raises(B200SAError):
        emu_engine.check_suffix_array_dev(x, 1 << 31, sa)


@pytest.mark.gpu
def test_gpu_wide_matches_int32_calls(gpu_engine, oracle):
    _check_small(gpu_engine, oracle)
