"""Inverse BWT of untrusted input: bytes + sentinel index that are not the BWT of any text must fail with
B200SA_EINVAL and must not touch memory outside the buffers (a flipped bit made the list ranking sum
lengths around a stray cycle of the LF mapping and the placement pass wrote out of bounds).  The unmodified reference
returns garbage of the right length on such input (msufsort.cpp:1821-2096); failing loudly is the safer contract.

CPU tier: the kernels under the emulator (an out-of-bounds write is a real heap overrun there, caught by the guard
pages / canaries below).  GPU tier: the same cases through the nvcc build."""
import numpy as np
import pytest

from cases import gen
from msufsort_b200.api import B200SAError


def _corruptions(bwt: np.ndarray, s: int):
    n = bwt.size
    rng = np.random.default_rng(n)
    for pos in (0, n // 3, n - 1):
        for bit in (0, 3, 7):
            b = bwt.copy()
            b[pos] ^= np.uint8(1 << bit)
            yield f"bit{bit}@{pos}", b, s
    yield "sentinel+1", bwt.copy(), s + 1 if s < n else s - 1
    yield "sentinel=1", bwt.copy(), 1 if s != 1 else 2
    yield "random", rng.integers(0, 256, n, dtype=np.uint8), int(rng.integers(1, n + 1))
    yield "two-symbols", rng.integers(97, 99, n, dtype=np.uint8), int(rng.integers(1, n + 1))
    yield "constant-wrong-sentinel", np.full(n, 97, np.uint8), max(1, n // 2)


def _is_bwt(oracle, b: np.ndarray, s: int) -> bool:
    """(b, s) is the BWT of some text iff its LF mapping is ONE cycle over the n+1 rows (with a unique smallest end
    marker every such cycle spells a text whose transform is (b, s)).  Checked with numpy pointer jumping: psi = stable
    sort of the rows by byte, every row must reach row 0."""
    n = b.size
    rows = np.arange(n, dtype=np.int64)
    rows += rows >= s
    psi = np.empty(n + 1, dtype=np.int64)
    psi[0] = 0                                  # terminal: self-loop
    psi[1:] = rows[np.argsort(b, kind="stable")]
    nxt = psi
    for _ in range(int(n + 1).bit_length() + 1):
        nxt = nxt[nxt]
    return bool((nxt == 0).all())


def _run(engine, oracle, family, n):
    x = gen(family, n)
    bwt, s = oracle.bwt(x)
    good = bwt.copy()
    engine.reverse_burrows_wheeler_transform(good, s)
    assert np.array_equal(good, x)
    rejected = 0
    for name, b, sent in _corruptions(bwt, s):
        guard = np.full(b.size + 256, 0xA5, np.uint8)          # canaries on both sides of the caller's buffer
        guard[128:128 + b.size] = b
        view = guard[128:128 + b.size]
        valid = _is_bwt(oracle, b, sent)
        if valid:
            engine.reverse_burrows_wheeler_transform(view, sent)
            assert np.array_equal(view, oracle.unbwt(b, sent)), (family, n, name)
        else:
            with pytest.raises(B200SAError) as ei:
                engine.reverse_burrows_wheeler_transform(view, sent)
            assert ei.value.code == 1 and "not a Burrows-Wheeler transform" in str(ei.value), (family, n, name)
            assert np.array_equal(view, b), "a rejected call must leave the caller's buffer unchanged"
            rejected += 1
        assert (guard[:128] == 0xA5).all() and (guard[128 + b.size:] == 0xA5).all(), (family, n, name)
    assert rejected >= 8
    # the context stays usable
    again = bwt.copy()
    engine.reverse_burrows_wheeler_transform(again, s)
    assert np.array_equal(again, x)


def _run_batch(engine, oracle):
    blocks = [gen("markov3", 3000), gen("rand", 257), gen("acgt_rep", 5000)]
    pairs = [oracle.bwt(b) for b in blocks]
    bw = [p[0].copy() for p in pairs]
    sent = [p[1] for p in pairs]
    out = engine.unbwt_batch(bw, sent)
    for o, b in zip(out, blocks):
        assert np.array_equal(o, b)
    bad = [b.copy() for b in bw]
    bad[1][100] ^= np.uint8(1)
    if not _is_bwt(oracle, bad[1], sent[1]):
        with pytest.raises(B200SAError) as ei:
            engine.unbwt_batch(bad, sent)
        assert ei.value.code == 1
    bad = [np.random.default_rng(5).integers(0, 256, b.size, dtype=np.uint8) for b in bw]
    with pytest.raises(B200SAError):
        engine.unbwt_batch(bad, sent)
    out = engine.unbwt_batch(bw, sent)
    for o, b in zip(out, blocks):
        assert np.array_equal(o, b)


@pytest.mark.parametrize("family,n", [("markov3", 3000), ("rand", 1000), ("zeros", 500), ("abcabca", 777), ("acgt_rep", 20011)])
def test_emu_untrusted_bwt(emu_engine, oracle, family, n):
    _run(emu_engine, oracle, family, n)


def test_emu_untrusted_bwt_batch(emu_engine, oracle):
    _run_batch(emu_engine, oracle)


@pytest.mark.gpu
@pytest.mark.parametrize("family,n", [("markov3", 3000), ("rand", 1000), ("zeros", 500), ("acgt_rep", 200003), ("markov3", (1 << 22) + 1)])
def test_gpu_untrusted_bwt(gpu_engine, oracle, family, n):
    _run(gpu_engine, oracle, family, n)


@pytest.mark.gpu
def test_gpu_untrusted_bwt_batch(gpu_engine, oracle):
    _run_batch(gpu_engine, oracle)
