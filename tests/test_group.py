"""One text sharded over several contexts through the C++ round loop (b200sa_group_*, b200sa_shard_sort): the ranks are
host threads of this process, the control plane is comm.cuh, the ISA lives in (peer) memory of the contexts.

CPU tier: the emulator build, 2..5 contexts.  GPU tier: several contexts on ONE device (what a single-GPU box can run:
the same kernels, the same peer-pointer loads and stores, the same barriers as on eight GPUs) and, when the box has them,
one context per GPU."""
import os

import numpy as np
import pytest

from cases import gen
from conftest import ROOT
from msufsort_b200.api import B200SAError, Group, Library

CASES = [("markov3", 40000), ("acgt_rep", 30011), ("abcabca", 9000), ("zeros", 5000), ("fib", 10000), ("periodic7", 8191), ("rand", 20000)]


def _check(group, oracle, x):
    sa, bwt, s = group.suffix_array_and_bwt(x)
    want = oracle.sa(x)
    assert np.array_equal(sa, want)
    wb, ws = oracle.bwt_from_sa(x, want)
    assert s == ws and np.array_equal(bwt, wb)
    assert np.array_equal(group.make_suffix_array(x), want)
    b = x.copy()
    assert group.forward_burrows_wheeler_transform(b) == ws and np.array_equal(b, wb)
    group.reverse_burrows_wheeler_transform(b, ws)
    assert np.array_equal(b, x)


@pytest.mark.parametrize("world", [2, 3, 5])
def test_emu_group(oracle, world, monkeypatch):
    monkeypatch.setenv("B200SA_GROUPSORT_TINY", "4")      # reach the CTA and the radix paths at these sizes too
    monkeypatch.setenv("B200SA_GROUPSORT_MEDIUM", "64")
    if world != 3:
        monkeypatch.setenv("B200SA_ISA_DIRECT_BYTES", "0")
        monkeypatch.setenv("B200SA_ISA_MIN_UPDATES", "1")
    if world == 5:
        monkeypatch.setenv("B200SA_ISA_PULL_FRACTION", "1000000000")   # radix rounds always pull the peers' ISA shards in bulk
    g = Group([0] * world, library=Library(os.path.join(ROOT, "tests", "emu", "libb200sa_emu.so")))
    try:
        for family, n in CASES:
            _check(g, oracle, gen(family, n))
        _check(g, oracle, gen("rand", 100))      # fewer than 4096 bytes per GPU: one context does it
    finally:
        g.close()


def test_emu_group_untrusted_bwt_and_recovery(oracle):
    g = Group([0, 0, 0], library=Library(os.path.join(ROOT, "tests", "emu", "libb200sa_emu.so")))
    try:
        x = gen("markov3", 30000)
        bwt, s = oracle.bwt(x)
        bad = bwt.copy()
        bad[1234] ^= np.uint8(4)
        keep = bad.copy()
        with pytest.raises(B200SAError) as ei:
            g.reverse_burrows_wheeler_transform(bad, s)
        assert ei.value.code == 1 and np.array_equal(bad, keep)
        _check(g, oracle, x)                      # the group (its comm included) stays usable after a failed call
    finally:
        g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3, 8])
def test_gpu_group_contexts_on_one_device(oracle, world):
    """the sharded path on a single-GPU box: `world` contexts on cuda:0"""
    g = Group([0] * world)
    try:
        for family, n in [("markov3", (1 << 22) + 5), ("acgt_rep", 1 << 22), ("rand", 1 << 20), ("abcabca", 1 << 20), ("fib", 1 << 19),
                          ("zeros", 1 << 18), ("periodic7", 300007)]:
            before = g.launch_count()
            _check(g, oracle, gen(family, n))
            assert g.launch_count() > before
    finally:
        g.close()


@pytest.mark.gpu
def test_gpu_group_one_context_per_gpu(oracle):
    import torch
    ng = torch.cuda.device_count()
    if ng < 2:
        pytest.skip("one context per GPU needs at least 2 GPUs (the sharded path itself is covered by test_gpu_group_contexts_on_one_device)")
    g = Group(list(range(min(ng, 8))))
    try:
        for family, n in [("markov3", (1 << 24) + 5), ("acgt_rep", 1 << 23), ("fib", 1 << 20)]:
            _check(g, oracle, gen(family, n))
    finally:
        g.close()


def _gpus_calls(lib, oracle, n):
    """the context-free calls with a GPU count (b200sa_*_gpus)"""
    import ctypes as C
    x = gen("markov3", n)
    sa = np.empty(n + 1, dtype=np.int32)
    lib.check(lib.cdll.b200sa_suffix_array_gpus(x.ctypes.data, n, sa.ctypes.data, 1))
    assert np.array_equal(sa, oracle.sa(x))
    b = x.copy()
    s = C.c_int32(0)
    lib.check(lib.cdll.b200sa_bwt_gpus(b.ctypes.data, n, C.byref(s), 0))      # 0 = every GPU present
    wb, ws = oracle.bwt(x)
    assert s.value == ws and np.array_equal(b, wb)
    lib.check(lib.cdll.b200sa_unbwt_gpus(b.ctypes.data, n, s.value, 0))
    assert np.array_equal(b, x)


def test_emu_calls_with_gpu_count(oracle):
    _gpus_calls(Library(os.path.join(ROOT, "tests", "emu", "libb200sa_emu.so")), oracle, 30011)


@pytest.mark.gpu
def test_gpu_calls_with_gpu_count(gpu_engine, oracle):
    _gpus_calls(gpu_engine.lib, oracle, (1 << 20) + 7)


def test_emu_group_fuzz(oracle):
    """random small texts over tiny alphabets (deep rounds, huge groups, empty key ranges) sharded over 2..6 contexts:
    suffix array, BWT and the sharded inverse against the oracle"""
    from hypothesis import HealthCheck, given, settings, strategies as st
    lib = Library(os.path.join(ROOT, "tests", "emu", "libb200sa_emu.so"))
    groups = {}

    @settings(max_examples=12, deadline=None, suppress_health_check=[HealthCheck.too_slow])
    @given(seed=st.integers(0, 2**31), sigma=st.integers(1, 4), n=st.integers(4096 * 6, 4096 * 6 + 3000), world=st.integers(2, 6))
    def run(seed, sigma, n, world):
        rng = np.random.default_rng(seed)
        x = rng.integers(0, sigma, size=n, dtype=np.uint8)
        if seed % 3 == 0:
            x[n // 2:] = x[: n - n // 2]                      # a long repeat: many rounds
        g = groups.get(world) or groups.setdefault(world, Group([0] * world, library=lib))
        sa, bwt, s = g.suffix_array_and_bwt(x)
        want = oracle.sa(x)
        assert np.array_equal(sa, want)
        wb, ws = oracle.bwt_from_sa(x, want)
        assert s == ws and np.array_equal(bwt, wb)
        b = bwt.copy()
        g.reverse_burrows_wheeler_transform(b, s)
        assert np.array_equal(b, x)

    try:
        run()
    finally:
        for g in groups.values():
            g.close()


def test_emu_group_reuses_the_resident_sort(oracle):
    """make_suffix_array followed by forward_burrows_wheeler_transform of the same bytes on one group costs ONE sharded sort (as on
    a single GPU); another text of the same size, or a call on one of the group's contexts in between, is sorted afresh"""
    from msufsort_b200.api import Engine
    lib = Library(os.path.join(ROOT, "tests", "emu", "libb200sa_emu.so"))
    g = Group([0, 0, 0], library=lib)
    try:
        x = gen("markov3", 40000)
        want = oracle.sa(x)
        wb, ws = oracle.bwt_from_sa(x, want)
        l0 = g.launch_count()
        assert np.array_equal(g.make_suffix_array(x), want)
        sort_launches = g.launch_count() - l0
        b = x.copy()
        l1 = g.launch_count()
        assert g.forward_burrows_wheeler_transform(b) == ws and np.array_equal(b, wb)
        reuse_launches = g.launch_count() - l1
        assert reuse_launches <= 12 and reuse_launches * 4 < sort_launches       # compare + BWT rows per context, no sort
        sa2, b2, s2 = g.suffix_array_and_bwt(x)                                  # both results from the resident sort
        assert np.array_equal(sa2, want) and np.array_equal(b2, wb) and s2 == ws
        # a different text of the same size must not be mistaken for the resident one
        y = x.copy()
        y[31337] ^= np.uint8(1)
        wy = oracle.sa(y)
        wyb, wys = oracle.bwt_from_sa(y, wy)
        by = y.copy()
        assert g.forward_burrows_wheeler_transform(by) == wys and np.array_equal(by, wyb)
        assert np.array_equal(g.make_suffix_array(y), wy)
        # a call on one of the group's contexts drops the resident result: the next group call sorts again
        ctx = lib.cdll.b200sa_group_context(g._g, 1)
        assert lib.cdll.b200sa_release_workspace(ctx) == 0
        l2 = g.launch_count()
        by = y.copy()
        assert g.forward_burrows_wheeler_transform(by) == wys and np.array_equal(by, wyb)
        assert g.launch_count() - l2 > reuse_launches * 2
    finally:
        g.close()


def _batch_blocks(rng, total, count):
    """`count` blocks of random sizes summing to `total`: mixed families, duplicates, empty blocks"""
    cuts = np.sort(rng.integers(0, total + 1, size=count - 1))
    sizes = np.diff(np.concatenate([[0], cuts, [total]]))
    fams = ["markov3", "acgt_rep", "rand", "zeros", "abcabca", "fib"]
    blocks = [gen(fams[i % len(fams)], int(s)) if s else np.empty(0, np.uint8) for i, s in enumerate(sizes)]
    blocks[-1] = blocks[0].copy()                                             # a duplicate block in another GPU's run
    return blocks


def _check_batch(group, oracle, blocks):
    sas = group.suffix_array_batch(blocks)
    bw, sent = group.bwt_batch(blocks)
    for b, blk in enumerate(blocks):
        if blk.size:
            want = oracle.sa(blk)
            assert np.array_equal(sas[b], want), b
            wb, ws = oracle.bwt_from_sa(blk, want)
            assert sent[b] == ws and np.array_equal(bw[b], wb), b
        else:
            assert sas[b].tolist() == [0] and sent[b] == 0
    back = group.unbwt_batch(bw, sent)
    for b, blk in enumerate(blocks):
        assert np.array_equal(back[b], blk), b


@pytest.mark.parametrize("world", [2, 3, 5])
def test_emu_group_batch(oracle, world):
    """batches of independent blocks over the contexts of a group: one run of blocks per context, results at the blocks' own places"""
    rng = np.random.default_rng(world)
    g = Group([0] * world, library=Library(os.path.join(ROOT, "tests", "emu", "libb200sa_emu.so")))
    try:
        _check_batch(g, oracle, _batch_blocks(rng, 60000, 37))
        _check_batch(g, oracle, [gen("markov3", 50000)] + [gen("rand", 10)] * (world + 3))   # one huge block: some runs are tiny or empty
        _check_batch(g, oracle, [gen("rand", 100), gen("zeros", 50)])                         # fewer blocks than contexts: one context does it
        # a corrupted block in the LAST run: the inverse fails and nothing of the caller's buffer is written, earlier runs included
        blocks = _batch_blocks(rng, 50000, 20)
        bw, sent = g.bwt_batch(blocks)
        packed = np.concatenate(bw)
        offsets = np.zeros(len(bw) + 1, dtype=np.int64)
        np.cumsum([b.size for b in bw], out=offsets[1:])
        big = max(range(len(bw) - 5, len(bw)), key=lambda b: bw[b].size)
        assert bw[big].size > 64
        packed[int(offsets[big]) + bw[big].size // 2] ^= np.uint8(1)
        keep = packed.copy()
        sent_arr = np.asarray(sent + [0], dtype=np.int32)
        rc = g.lib.cdll.b200sa_group_unbwt_batch(g._g, packed.ctypes.data, offsets.ctypes.data, len(bw), sent_arr.ctypes.data)
        assert rc == 1 and np.array_equal(packed, keep)                                        # B200SA_EINVAL, buffer untouched
        _check_batch(g, oracle, blocks)                                                        # the group stays usable
    finally:
        g.close()
