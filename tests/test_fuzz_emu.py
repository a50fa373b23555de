"""Property-based checks of the kernel logic under the emulator (CPU tier): random batches of small blocks over tiny
alphabets (many equal blocks, empty blocks, blocks that are prefixes of each other), random texts for the LCP array.
Small alphabets make long repeats, deep doubling rounds and cross-block ties likely."""
import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

block = st.lists(st.integers(0, 2), min_size=0, max_size=40).map(lambda v: np.array(v, dtype=np.uint8))
SETTINGS = dict(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])


@settings(**SETTINGS)
@given(blocks=st.lists(block, min_size=1, max_size=12), dup=st.integers(0, 3))
def test_fuzz_batch_equals_per_block_oracle(emu_engine, oracle, blocks, dup):
    blocks = blocks + blocks[:dup]          # exact duplicates inside one batch
    sas = emu_engine.suffix_array_batch(blocks)
    bw, sent = emu_engine.bwt_batch(blocks)
    for b, x in enumerate(blocks):
        if x.size == 0:
            assert sas[b].tolist() == [0] and sent[b] == 0 and bw[b].size == 0
            continue
        want = oracle.sa_bruteforce(x)
        assert np.array_equal(sas[b], want), (b, x.tolist())
        wb, ws = oracle.bwt_from_sa(x, want)
        assert sent[b] == ws and np.array_equal(bw[b], wb), (b, x.tolist())
    back = emu_engine.unbwt_batch(bw, sent)
    for b, x in enumerate(blocks):
        assert np.array_equal(back[b], x), (b, x.tolist())


@settings(**SETTINGS)
@given(text=st.lists(st.integers(0, 1), min_size=1, max_size=300), shift=st.integers(0, 3))
def test_fuzz_lcp_against_kasai_and_bruteforce_sa(emu_engine, oracle, text, shift):
    buf = np.array([7] * shift + text, dtype=np.uint8)
    x = buf[shift:]                          # every alignment of the text pointer
    sa = oracle.sa_bruteforce(x)
    lcp, sa2 = emu_engine.make_lcp_array(x, return_sa=True)
    assert np.array_equal(sa2, sa)
    assert np.array_equal(lcp, oracle.lcp(x, sa, kasai=True))
    assert np.array_equal(lcp, oracle.lcp(x, sa, kasai=False))
    out = np.empty(x.size + 1, dtype=np.int32)
    emu_engine.lcp_dev(x, x.size, sa, out)
    assert np.array_equal(out, lcp)


@settings(**SETTINGS)
@given(text=st.lists(st.integers(0, 255), min_size=1, max_size=200), wide=st.booleans())
def test_fuzz_sa_bwt_roundtrip_any_bytes(emu_engine, oracle, text, wide):
    x = np.array(text, dtype=np.uint8)
    want = oracle.sa_bruteforce(x)
    if wide:
        sa, bwt, s = emu_engine.suffix_array_and_bwt_u32(x)
    else:
        sa, bwt, s = emu_engine.suffix_array_and_bwt(x)
    assert np.array_equal(sa.astype(np.int64), want.astype(np.int64))
    wb, ws = oracle.bwt_from_sa(x, want)
    assert s == ws and np.array_equal(bwt, wb)
    back = bwt.copy()
    emu_engine.reverse_burrows_wheeler_transform(back, s)
    assert np.array_equal(back, x)
