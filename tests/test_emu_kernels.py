"""CPU tier: the CUDA sources compiled against the SIMT emulator (tests/emu) — exercises the kernel
LOGIC (tile indexing, warp ranking, look-back chaining, scans, walkers) through the same C ABI.
This is test infrastructure: the product never loads the emulator build."""
import numpy as np
import pytest

from cases import EDGE_SIZES, FAMILIES, gen, small_alphabet_exhaustive


@pytest.mark.parametrize("family", FAMILIES)
def test_emu_sa_bwt_unbwt_edge_sizes(emu_engine, oracle, family):
    for n in EDGE_SIZES:
        x = gen(family, n)
        want = oracle.sa(x)
        assert np.array_equal(emu_engine.make_suffix_array(x), want), (family, n)
        b = x.copy()
        s = emu_engine.forward_burrows_wheeler_transform(b)
        wb, ws = oracle.bwt_from_sa(x, want)
        assert s == ws and np.array_equal(b, wb), (family, n)
        emu_engine.reverse_burrows_wheeler_transform(b, s)
        assert np.array_equal(b, x), (family, n)


def test_emu_exhaustive_binary_strings(emu_engine, oracle):
    for x in small_alphabet_exhaustive(7, 2):
        assert np.array_equal(emu_engine.make_suffix_array(x), oracle.sa_bruteforce(x)), x.tolist()


@pytest.mark.parametrize("family", FAMILIES)
def test_emu_medium(emu_engine, oracle, family):
    n = 50021
    x = gen(family, n)
    sa, bwt, s = emu_engine.suffix_array_and_bwt(x)
    want = oracle.sa(x)
    assert np.array_equal(sa, want)
    wb, ws = oracle.bwt_from_sa(x, want)
    assert s == ws and np.array_equal(bwt, wb)
    u = bwt.copy()
    emu_engine.reverse_burrows_wheeler_transform(u, s)
    assert np.array_equal(u, x)
    assert emu_engine.check_suffix_array_dev(x, n, sa) == 0


@pytest.mark.parametrize("m,bits", [(1, 64), (7, 8), (4096, 64), (4097, 17), (20000, 64), (70001, 33)])
def test_emu_radix_sort_pairs(emu_engine, m, bits):
    rng = np.random.default_rng(m + bits)
    keys = rng.integers(0, 1 << 63, size=m, dtype=np.uint64)
    if bits < 64:
        keys &= np.uint64((1 << bits) - 1)
    keys[: m // 3] = keys[m // 2]
    vals = rng.permutation(m).astype(np.uint32)
    order = np.argsort(keys, kind="stable")
    k, ka, v, va = keys.copy(), np.empty_like(keys), vals.copy(), np.empty_like(vals)
    side = emu_engine.radix_sort_pairs_dev(k, ka, v, va, m, 0, bits)
    assert np.array_equal(ka if side else k, keys[order])
    assert np.array_equal(va if side else v, vals[order])
    k = keys.copy()
    emu_engine.radix_sort_pairs_dev(k, ka, None, va, m, 0, bits)
    assert np.array_equal(va, order.astype(np.uint32))


def test_emu_checker_detects_corruption(emu_engine, oracle):
    x = gen("markov3", 20000)
    sa = oracle.sa(x)
    assert emu_engine.check_suffix_array_dev(x, x.size, sa) == 0
    bad = sa.copy(); bad[[100, 101]] = bad[[101, 100]]
    assert emu_engine.check_suffix_array_dev(x, x.size, bad) > 0
    bad = sa.copy(); bad[3] = bad[4]
    assert emu_engine.check_suffix_array_dev(x, x.size, bad) > 0


def test_emu_host_validator(emu_engine, oracle):
    x = gen("acgt_rep", 30000)
    sa = oracle.sa(x)
    assert emu_engine.check_suffix_array(x, sa) == 0
    bad = sa.copy(); bad[[7, 9]] = bad[[9, 7]]
    assert emu_engine.check_suffix_array(x, bad) > 0
    assert emu_engine.check_suffix_array(np.empty(0, np.uint8), np.zeros(1, np.int32)) == 0
    with pytest.raises(ValueError):
        emu_engine.check_suffix_array(x, sa[:-1])


def test_emu_profile_accounting(emu_engine):
    emu_engine.profile_reset()
    x = gen("markov3", 30000)
    emu_engine.make_suffix_array(x)
    p = emu_engine.profile()
    assert p["rounds"] >= 2
    assert p["phases"]["sort_pass"]["launches"] == p["sort_passes"]
    # first sweep of round 0 moves 20 B per tuple (generated values), all others 24 B
    assert p["phases"]["sort_pass"]["alg_bytes"] == 24 * p["sorted_tuples"] - 4 * x.size


@pytest.mark.parametrize("env", [
    {"B200SA_GROUPSORT_TINY": "2", "B200SA_GROUPSORT_MEDIUM": "8", "B200SA_GROUPSORT_AVG": "1000000"},   # tiny/medium/huge + fallback
    {"B200SA_GROUPSORT_TINY": "3", "B200SA_GROUPSORT_MEDIUM": "4096", "B200SA_GROUPSORT_AVG": "1000000"},  # CTA bitonic path
    {"B200SA_GROUPSORT_AVG": "0"},                                                                        # radix rounds only
    {"B200SA_ISA_DIRECT_BYTES": "0", "B200SA_ISA_MIN_UPDATES": "1"},                                      # bucketed ISA update
    {"B200SA_PACK_RADIX": "0"},                                                                           # bit-packed round-0 keys (mixed radix is the default)
    {"B200SA_MAX_KEY_BITS": "24"},                                                                        # narrow round-0 keys
], ids=["groups-small-thresholds", "groups-cta", "radix-only", "bucketed-isa", "bit-packed-keys", "narrow-keys"])
def test_emu_round_variants(oracle, env, monkeypatch):
    """every way a doubling round can run (in-place group sort: thread / CTA / per-group radix / fallback; radix
    rounds; direct and bucketed ISA update) gives the oracle's suffix array"""
    import os
    from conftest import ROOT
    from msufsort_b200.api import Engine, Library
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    eng = Engine(0, library=Library(os.path.join(ROOT, "tests", "emu", "libb200sa_emu.so")))
    try:
        for family, n in [("markov3", 30011), ("acgt_rep", 20000), ("zeros", 3000), ("abcabca", 5000), ("fib", 6000),
                          ("sigma2", 4097), ("sigma3", 5000), ("periodic1009", 9000), ("zero_tail", 777), ("rand", 2000)]:
            x = gen(family, n)
            assert np.array_equal(eng.make_suffix_array(x), oracle.sa(x)), (family, n, env)
        # the same variants through the batched sort (block number above the symbol part of the key)
        blocks = [gen("markov3", 5000), gen("sigma3", 700), gen("zeros", 300), np.empty(0, np.uint8), gen("rand", 40)]
        for got, x in zip(eng.suffix_array_batch(blocks), blocks):
            assert np.array_equal(got, oracle.sa(x)) if x.size else got.tolist() == [0], env
    finally:
        eng.close()


@pytest.mark.parametrize("cap_mult", ["1", "4"])
def test_emu_inverse_bwt_window_overflow(oracle, cap_mult, monkeypatch):
    """inverse BWT with small decode windows (many walkers outgrow them and finish in the placement pass)"""
    import os
    from conftest import ROOT
    from msufsort_b200.api import Engine, Library
    monkeypatch.setenv("B200SA_UNBWT_CAP_MULT", cap_mult)
    eng = Engine(0, library=Library(os.path.join(ROOT, "tests", "emu", "libb200sa_emu.so")))
    try:
        for family, n in [("markov3", 100003), ("zeros", 20000), ("abcabca", 30000), ("rand", 70000), ("fib", 50000), ("rand", 5)]:
            x = gen(family, n)
            bwt, s = oracle.bwt(x)
            b = bwt.copy()
            eng.reverse_burrows_wheeler_transform(b, s)
            assert np.array_equal(b, x), (family, n, cap_mult)
    finally:
        eng.close()


def test_emu_resident_suffix_array_is_reused(emu_engine, oracle):
    """make_suffix_array then forward_burrows_wheeler_transform of the same bytes: one sort; of other bytes: two"""
    x = gen("markov3", 20011)
    y = x.copy()
    y[777] ^= 1
    emu_engine.make_suffix_array(x)
    p0 = emu_engine.profile()["rounds"]
    b = x.copy()
    s = emu_engine.forward_burrows_wheeler_transform(b)
    assert emu_engine.profile()["rounds"] == p0, "the same text must not be sorted again"
    wb, ws = oracle.bwt(x)
    assert s == ws and np.array_equal(b, wb)
    emu_engine.make_suffix_array(x)
    p1 = emu_engine.profile()["rounds"]
    b = y.copy()
    s = emu_engine.forward_burrows_wheeler_transform(b)
    assert emu_engine.profile()["rounds"] > p1, "a different text of the same length must be sorted"
    wb, ws = oracle.bwt(y)
    assert s == ws and np.array_equal(b, wb)
    # any other call in between drops the resident result
    emu_engine.make_suffix_array(x)
    emu_engine.check_suffix_array(x, oracle.sa(x))
    p2 = emu_engine.profile()["rounds"]
    b = x.copy()
    emu_engine.forward_burrows_wheeler_transform(b)
    assert emu_engine.profile()["rounds"] > p2


def test_emu_bwt_windowed_gather(oracle, monkeypatch):
    """forward BWT in passes over text windows (taken for texts larger than L2 on the GPU; forced here with tiny windows),
    aligned and unaligned output buffers, whole transforms and the row ranges of a sharded run"""
    import os
    from conftest import ROOT
    from msufsort_b200.api import Engine, Group, Library
    monkeypatch.setenv("B200SA_BWT_WINDOW_BYTES", "1000")
    monkeypatch.setenv("B200SA_BWT_MAX_PASSES", "64")
    lib = Library(os.path.join(ROOT, "tests", "emu", "libb200sa_emu.so"))
    eng = Engine(0, library=lib)
    try:
        for family in FAMILIES:
            for n in (63, 64, 65, 4097, 20011):
                x = gen(family, n)
                want = oracle.sa(x)
                wb, ws = oracle.bwt_from_sa(x, want)
                sa, bwt, s = eng.suffix_array_and_bwt(x)
                assert np.array_equal(sa, want) and s == ws and np.array_equal(bwt, wb), (family, n)
                for shift in (1, 3):                     # output buffer at every alignment
                    buf = np.zeros(n + 8, dtype=np.uint8)
                    view = buf[shift:shift + n]
                    view[:] = x
                    assert eng.forward_burrows_wheeler_transform(view) == ws and np.array_equal(view, wb), (family, n, shift)
    finally:
        eng.close()
    g = Group([0, 0, 0], library=lib)
    try:
        for family, n in (("markov3", 40003), ("fib", 20000), ("zeros", 9000)):
            x = gen(family, n)
            sa, bwt, s = g.suffix_array_and_bwt(x)
            wb, ws = oracle.bwt(x)
            assert s == ws and np.array_equal(bwt, wb), (family, n)
    finally:
        g.close()
