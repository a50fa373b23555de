"""GPU tier: the nvcc-built library on a real B200, through the C ABI, against the oracle.

Bit-exact bar (integer/byte work): SA, BWT bytes, sentinel index and inverse BWT must equal the
oracle's (= the reference's, see test_oracle.py) on the same bytes.
"""
import numpy as np
import pytest

from cases import EDGE_SIZES, FAMILIES, gen, small_alphabet_exhaustive

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    return torch


# ---------------------------------------------------------------------------------------------
def test_product_library_is_cuda_build(gpu_engine):
    import os
    assert os.path.basename(gpu_engine.lib.path) == "libb200sa.so"
    assert gpu_engine.lib.cdll.b200sa_device_count() >= 1


@pytest.mark.parametrize("m,bits", [(1, 64), (5, 8), (4096, 64), (4097, 17), (100000, 64), (1 << 20, 40), (3000000, 64)])
def test_radix_sort_pairs(gpu_engine, m, bits):
    torch = _torch()
    rng = np.random.default_rng(m * 131 + bits)
    keys = rng.integers(0, 1 << 63, size=m, dtype=np.uint64)
    if bits < 64:
        keys &= np.uint64((1 << bits) - 1)
    keys[: m // 3] = keys[m // 2]  # plenty of duplicates: stability matters
    vals = rng.permutation(m).astype(np.uint32)
    order = np.argsort(keys, kind="stable")
    dk = torch.from_numpy(keys.view(np.int64)).cuda()
    dka = torch.empty_like(dk)
    dv = torch.from_numpy(vals.view(np.int32)).cuda()
    dva = torch.empty_like(dv)
    side = gpu_engine.radix_sort_pairs_dev(dk, dka, dv, dva, m, 0, bits)
    torch.cuda.synchronize()
    ok = (dka if side else dk).cpu().numpy().view(np.uint64)
    ov = (dva if side else dv).cpu().numpy().view(np.uint32)
    assert np.array_equal(ok, keys[order])
    assert np.array_equal(ov, vals[order])
    # generated values (element indices)
    dk = torch.from_numpy(keys.view(np.int64)).cuda()
    side = gpu_engine.radix_sort_pairs_dev(dk, dka, None, dva, m, 0, bits)
    torch.cuda.synchronize()
    assert np.array_equal(dva.cpu().numpy().view(np.uint32), order.astype(np.uint32))


@pytest.mark.parametrize("family", FAMILIES)
def test_sa_bwt_unbwt_edge_sizes(gpu_engine, oracle, family):
    for n in EDGE_SIZES:
        x = gen(family, n)
        sa = gpu_engine.make_suffix_array(x)
        want = oracle.sa(x)
        assert np.array_equal(sa, want), (family, n)
        b = x.copy()
        s = gpu_engine.forward_burrows_wheeler_transform(b)
        wb, ws = oracle.bwt(x)
        assert s == ws and np.array_equal(b, wb), (family, n)
        gpu_engine.reverse_burrows_wheeler_transform(b, s)
        assert np.array_equal(b, x), (family, n)


def test_exhaustive_binary_strings(gpu_engine, oracle):
    # all strings over {0x00, 0x01} up to length 8: the byte-0-versus-sentinel rule
    for x in small_alphabet_exhaustive(8, 2):
        assert np.array_equal(gpu_engine.make_suffix_array(x), oracle.sa_bruteforce(x)), x.tolist()


@pytest.mark.parametrize("family", FAMILIES)
@pytest.mark.parametrize("n", [65536 + 17, (1 << 20) + 3])
def test_sa_bwt_unbwt_medium(gpu_engine, oracle, family, n):
    x = gen(family, n)
    sa, bwt, s = gpu_engine.suffix_array_and_bwt(x)
    want = oracle.sa(x)
    assert np.array_equal(sa, want)
    wb, ws = oracle.bwt_from_sa(x, want)
    assert s == ws and np.array_equal(bwt, wb)
    u = bwt.copy()
    gpu_engine.reverse_burrows_wheeler_transform(u, s)
    assert np.array_equal(u, x)


@pytest.mark.parametrize("family", ["rand", "markov3", "acgt_rep"])
def test_config1_16mib_vs_reference(gpu_engine, ref, family):
    """BASELINE.json configs[0]: 16 MiB buffer, SA + forward BWT bit-exact against the unmodified
    reference library (oracle/_ref) on host cores; plus the inverse transform."""
    n = 1 << 24
    x = gen(family, n)
    sa = gpu_engine.make_suffix_array(x)
    want = ref.ref_sa(x, threads=8)
    assert np.array_equal(sa, want)
    b = x.copy()
    s = gpu_engine.forward_burrows_wheeler_transform(b)
    rb, rs = ref.ref_bwt(x, threads=8)
    assert s == rs and np.array_equal(b, rb)
    gpu_engine.reverse_burrows_wheeler_transform(b, s)
    assert np.array_equal(b, x)


@pytest.mark.parametrize("family,n", [("fib", 1 << 22), ("abcabca", 1 << 24), ("zeros", 1 << 22), ("periodic1009", 1 << 23)])
def test_pathological_against_checker_and_oracle(gpu_engine, oracle, family, n):
    """deep-doubling inputs: GPU validator says 0 bad rows, and the SA equals the oracle's"""
    torch = _torch()
    x = gen(family, n)
    dt = torch.from_numpy(x).cuda()
    dsa = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    gpu_engine.suffix_array_dev(dt, n, dsa)
    assert gpu_engine.check_suffix_array_dev(dt, n, dsa) == 0
    assert np.array_equal(dsa.cpu().numpy(), oracle.sa(x))


def test_checker_detects_corruption(gpu_engine):
    torch = _torch()
    n = 100000
    x = gen("markov3", n)
    dt = torch.from_numpy(x).cuda()
    dsa = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    gpu_engine.suffix_array_dev(dt, n, dsa)
    assert gpu_engine.check_suffix_array_dev(dt, n, dsa) == 0
    bad = dsa.clone()
    bad[[500, 501]] = bad[[501, 500]]
    assert gpu_engine.check_suffix_array_dev(dt, n, bad) > 0
    bad = dsa.clone()
    bad[7] = bad[8]
    assert gpu_engine.check_suffix_array_dev(dt, n, bad) > 0


def test_device_entry_points_match_host_entry_points(gpu_engine, oracle):
    torch = _torch()
    n = 300001
    x = gen("acgt_rep", n)
    dt = torch.from_numpy(x).cuda()
    dsa = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    dbwt = torch.empty(n, dtype=torch.uint8, device="cuda")
    s = gpu_engine.bwt_dev(dt, n, dbwt, dsa)
    want = oracle.sa(x)
    wb, ws = oracle.bwt(x)
    assert s == ws
    assert np.array_equal(dsa.cpu().numpy(), want)
    assert np.array_equal(dbwt.cpu().numpy(), wb)
    dout = torch.empty(n, dtype=torch.uint8, device="cuda")
    gpu_engine.unbwt_dev(dbwt, n, s, dout)
    assert np.array_equal(dout.cpu().numpy(), x)


def test_errors(gpu_engine):
    from msufsort_b200 import B200SAError
    x = gen("rand", 100)
    b = x.copy()
    with pytest.raises(B200SAError):
        gpu_engine.reverse_burrows_wheeler_transform(b, 0)      # sentinel must be in [1, n]
    with pytest.raises(B200SAError):
        gpu_engine.reverse_burrows_wheeler_transform(b, 101)
    assert gpu_engine.make_suffix_array(np.empty(0, dtype=np.uint8)).tolist() == [0]


def test_profile_counts_launches(gpu_engine):
    gpu_engine.profile_reset()
    gpu_engine.set_profiling(True)
    x = gen("markov3", 1 << 20)
    before = gpu_engine.launch_count()
    gpu_engine.make_suffix_array(x)
    p = gpu_engine.profile()
    gpu_engine.set_profiling(False)
    assert gpu_engine.launch_count() > before
    assert p["phases"]["sort_pass"]["launches"] >= 8
    assert p["phases"]["sort_pass"]["ms"] > 0
    assert p["rounds"] >= 2
