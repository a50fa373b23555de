"""Batches of independent blocks: all blocks are suffix-sorted together by one launch
sequence; every block's SA / BWT / sentinel index must equal what a per-block call (= the oracle = the
reference) gives, whatever the other blocks of the batch contain."""
import numpy as np
import pytest

from cases import FAMILIES, gen


def _check(eng, oracle, blocks):
    sas = eng.suffix_array_batch(blocks)
    bw, sent = eng.bwt_batch(blocks)
    assert len(sas) == len(bw) == len(sent) == len(blocks)
    for b, x in enumerate(blocks):
        x = np.asarray(x, dtype=np.uint8)
        if x.size == 0:
            assert sas[b].tolist() == [0] and sent[b] == 0 and bw[b].size == 0
            continue
        want = oracle.sa(x)
        assert np.array_equal(sas[b], want), (b, x.size)
        wb, ws = oracle.bwt_from_sa(x, want)
        assert sent[b] == ws and np.array_equal(bw[b], wb), (b, x.size)
    back = eng.unbwt_batch(bw, sent)
    for b, x in enumerate(blocks):
        assert np.array_equal(back[b], np.asarray(x, dtype=np.uint8)), b


def _cases(scale):
    rng = np.random.default_rng(11)
    yield "single", [gen("markov3", 1000 * scale)]
    yield "mixed", [gen("markov3", 1000 * scale), gen("rand", 17), gen("zeros", 300), np.empty(0, np.uint8), gen("fib", 4097),
                    gen("acgt_rep", 5000 * scale), gen("zeros", 1), np.empty(0, np.uint8)]
    x = gen("markov3", 3000 * scale)
    yield "identical-blocks", [x] * 9                      # cross-block repeats must not deepen the rounds
    yield "identical-runs", [gen("zeros", 50)] * 40
    yield "tiny", [rng.integers(0, 3, size=int(rng.integers(0, 12)), dtype=np.uint8) for _ in range(700)]
    yield "families", [gen(f, int(rng.integers(1, 9000 * scale))) for f in FAMILIES for _ in range(3)]
    yield "two-random", [gen("rand", 30000 * scale), gen("rand", 20000 * scale)]
    yield "all-empty", [np.empty(0, np.uint8)] * 3
    yield "zero-bytes", [np.zeros(5, np.uint8), np.array([0, 0, 1, 0], np.uint8), np.array([1, 0, 0], np.uint8)]


@pytest.mark.parametrize("name", [n for n, _ in _cases(1)])
def test_emu_batch(emu_engine, oracle, name):
    _check(emu_engine, oracle, dict(_cases(1))[name])


def test_emu_batch_profile_one_sort(emu_engine):
    """the whole batch is ONE sort: launches do not grow with the number of blocks"""
    a = [gen("markov3", 2000)] * 4
    b = [gen("markov3", 250)] * 32
    counts = []
    for blocks in (a, b):
        before = emu_engine.launch_count()
        emu_engine.bwt_batch(blocks)
        counts.append(emu_engine.launch_count() - before)
    assert counts[1] <= counts[0] + 40, counts


@pytest.mark.parametrize("cap_mult", ["1", "4"])
def test_emu_batch_inverse_window_overflow(oracle, cap_mult, monkeypatch):
    """batched inverse with small decode windows; blocks much longer and much shorter than the walker spacing"""
    import os
    from conftest import ROOT
    from msufsort_b200.api import Engine, Library
    monkeypatch.setenv("B200SA_UNBWT_CAP_MULT", cap_mult)
    eng = Engine(0, library=Library(os.path.join(ROOT, "tests", "emu", "libb200sa_emu.so")))
    try:
        blocks = [gen("markov3", 50003), gen("zeros", 9000), gen("rand", 63), gen("rand", 64), gen("rand", 65), gen("fib", 20000),
                  np.empty(0, np.uint8), gen("abcabca", 127), gen("abcabca", 128), gen("rand", 1)]
        bw, sent = [], []
        for x in blocks:
            if x.size:
                b, s = oracle.bwt(x)
            else:
                b, s = x, 0
            bw.append(b); sent.append(s)
        back = eng.unbwt_batch(bw, sent)
        for b, x in enumerate(blocks):
            assert np.array_equal(back[b], x), (b, cap_mult)
        before = eng.launch_count()
        eng.unbwt_batch(bw * 8, sent * 8)
        assert eng.launch_count() - before < 60  # one sort + one walk, whatever the number of blocks
    finally:
        eng.close()


def test_emu_batch_errors(emu_engine):
    from msufsort_b200.api import B200SAError
    assert emu_engine.suffix_array_batch([]) == []
    with pytest.raises(ValueError):
        emu_engine.unbwt_batch([np.zeros(3, np.uint8)], [1, 2])
    with pytest.raises(B200SAError):
        emu_engine.unbwt_batch([np.zeros(3, np.uint8)], [7])  # sentinel outside [1, n]
    with pytest.raises(B200SAError):
        emu_engine.batch_dev(np.zeros(4, np.uint8), np.array([0, 3, 2], dtype=np.int64))  # decreasing offsets


@pytest.mark.gpu
@pytest.mark.parametrize("name", [n for n, _ in _cases(8)])
def test_gpu_batch(gpu_engine, oracle, name):
    _check(gpu_engine, oracle, dict(_cases(8))[name])


@pytest.mark.gpu
def test_gpu_batch_many_blocks_device_resident(gpu_engine, oracle):
    """4096 blocks of 16 KiB .. 48 KiB through the device-resident entry point; spot-checked against the oracle,
    all blocks round-tripped through the inverse"""
    import torch
    rng = np.random.default_rng(3)
    sizes = rng.integers(16 << 10, 48 << 10, size=4096)
    text = gen("markov3", int(sizes.sum()))
    offsets = np.zeros(sizes.size + 1, dtype=np.int64)
    np.cumsum(sizes, out=offsets[1:])
    total, count = int(offsets[-1]), sizes.size
    d_blocks = torch.from_numpy(text).cuda()
    d_bwt = torch.empty(total, dtype=torch.uint8, device="cuda")
    d_sa = torch.empty(total + count, dtype=torch.int32, device="cuda")
    sent = gpu_engine.batch_dev(d_blocks, offsets, d_bwt, d_sa)
    torch.cuda.synchronize()
    bwt, sa = d_bwt.cpu().numpy(), d_sa.cpu().numpy()
    for b in list(range(0, count, 257)) + [count - 1]:
        x = text[offsets[b]:offsets[b + 1]]
        want = oracle.sa(x)
        assert np.array_equal(sa[offsets[b] + b: offsets[b + 1] + b + 1], want), b
        wb, ws = oracle.bwt_from_sa(x, want)
        assert int(sent[b]) == ws and np.array_equal(bwt[offsets[b]:offsets[b + 1]], wb), b
    d_back = torch.zeros(total, dtype=torch.uint8, device="cuda")
    gpu_engine.unbwt_batch_dev(d_bwt, offsets, sent, d_back)
    torch.cuda.synchronize()
    assert torch.equal(d_back, d_blocks)


# ---- streaming pipeline ---------------------------------------------------------------------------
def _pipeline_roundtrip(lib, oracle, nbatches, blocks_per_batch, block_len, depth, devices=None):
    from msufsort_b200.api import Pipeline
    rng = np.random.default_rng(depth)
    jobs = []
    with Pipeline(0, depth, library=lib, devices=devices) as pipe:
        for j in range(nbatches):
            blocks = [gen(["markov3", "rand", "acgt_rep", "zeros"][(j + b) % 4], int(rng.integers(1, block_len))) for b in range(blocks_per_batch)]
            packed = np.concatenate(blocks)
            offsets = np.zeros(len(blocks) + 1, dtype=np.int64)
            np.cumsum([x.size for x in blocks], out=offsets[1:])
            sent = np.zeros(len(blocks), dtype=np.int32)
            orig = packed.copy()
            jobs.append((pipe.submit_bwt(packed, offsets, sent), packed, offsets, sent, orig, blocks))
        for t, packed, offsets, sent, orig, blocks in jobs:
            pipe.wait(t)
            for b, x in enumerate(blocks):
                wb, ws = oracle.bwt(x)
                assert ws == sent[b] and np.array_equal(packed[offsets[b]:offsets[b + 1]], wb), b
        # inverse through the same pipeline, all in flight at once, then drain
        for t, packed, offsets, sent, orig, blocks in jobs:
            pipe.submit_unbwt(packed, offsets, sent)
        pipe.drain()
        for t, packed, offsets, sent, orig, blocks in jobs:
            assert np.array_equal(packed, orig)
        # suffix arrays
        packed, offsets = jobs[0][4], jobs[0][2]
        sa = np.empty(packed.size + offsets.size - 1, dtype=np.int32)
        pipe.wait(pipe.submit_suffix_array(packed, offsets, sa))
        for b, x in enumerate(jobs[0][5]):
            assert np.array_equal(sa[offsets[b] + b: offsets[b + 1] + b + 1], oracle.sa(x))


def test_emu_pipeline(oracle):
    import os
    from conftest import ROOT
    from msufsort_b200.api import B200SAError, Library, Pipeline
    lib = Library(os.path.join(ROOT, "tests", "emu", "libb200sa_emu.so"))
    _pipeline_roundtrip(lib, oracle, nbatches=5, blocks_per_batch=6, block_len=3000, depth=3)
    with Pipeline(0, 2, library=lib) as pipe:
        bad = np.zeros(4, dtype=np.uint8)
        t = pipe.submit_unbwt(bad, np.array([0, 4], dtype=np.int64), np.array([9], dtype=np.int32))  # sentinel outside [1, n]
        with pytest.raises(B200SAError):
            pipe.wait(t)
        with pytest.raises(B200SAError):
            pipe.wait(12345)
        with pytest.raises(B200SAError):
            pipe.wait(t)                     # a ticket can be collected once
    with pytest.raises(B200SAError):
        Pipeline(0, 0, library=lib)
    # two contexts on each of three listed devices behind the one queue (b200sa_pipeline_create_devices)
    _pipeline_roundtrip(lib, oracle, nbatches=7, blocks_per_batch=4, block_len=2500, depth=2, devices=[0, 0, 0])
    with pytest.raises(B200SAError):
        Pipeline(library=lib, devices=[])


def test_emu_pipeline_waiters_never_hang(oracle):
    """two threads waiting for one ticket, or a drain while a thread waits, must not leave a waiter blocked:
    exactly one waiter collects the result, the others return EINVAL"""
    import os
    import threading
    from conftest import ROOT
    from msufsort_b200.api import B200SAError, Library, Pipeline
    lib = Library(os.path.join(ROOT, "tests", "emu", "libb200sa_emu.so"))
    with Pipeline(0, 1, library=lib) as pipe:
        for round_ in range(3):
            blocks = [gen("markov3", 20000 + 17 * round_) for _ in range(4)]
            packed = np.concatenate(blocks)
            offsets = np.concatenate([[0], np.cumsum([b.size for b in blocks])]).astype(np.int64)
            sent = np.zeros(len(blocks), dtype=np.int32)
            t = pipe.submit_bwt(packed, offsets, sent)
            outcomes = []

            def waiter():
                try:
                    pipe.wait(t)
                    outcomes.append("ok")
                except B200SAError as e:
                    outcomes.append("err%d" % e.code)

            threads = [threading.Thread(target=waiter) for _ in range(3)]
            for th in threads:
                th.start()
            if round_ == 2:
                try:
                    pipe.drain()             # may collect the ticket before any waiter does
                except B200SAError:
                    pass
            for th in threads:
                th.join(timeout=120)
                assert not th.is_alive(), "a waiter is blocked forever"
            assert outcomes.count("ok") <= 1 and all(o in ("ok", "err1") for o in outcomes), outcomes
            if round_ < 2:
                assert outcomes.count("ok") == 1
                want = [oracle.bwt(b) for b in blocks]
                assert [w[1] for w in want] == sent.tolist()


@pytest.mark.gpu
def test_gpu_pipeline(gpu_engine, oracle):
    _pipeline_roundtrip(gpu_engine.lib, oracle, nbatches=8, blocks_per_batch=24, block_len=200000, depth=3)
