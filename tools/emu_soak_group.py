"""Randomised soak of the SHARDED path under the CPU emulator (test infrastructure): random families / tiny alphabets / long
repeats, sizes up to 150 k, 2..8 contexts in one group; suffix array, BWT (also served from the resident sort), the sharded inverse BWT and a random cut into blocks as one batch
over the contexts, all against the oracle.
Knobs come from the environment (B200SA_*), duration and seed from SECS / SEED.  B200SA_EMU_SANITIZER=asan|ubsan (with the
matching LD_PRELOAD, see tools/emu_asan.sh) runs it over a sanitizer build.
    SECS=600 SEED=3 python tools/emu_soak_group.py"""
import os, sys, time
import numpy as np
ROOT_ = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT_); sys.path.insert(0, os.path.join(ROOT_, "tests"))
from conftest import Oracle, ROOT
from cases import FAMILIES, gen
from msufsort_b200.api import Group, Library

o = Oracle()
lib = Library(os.path.join(ROOT, "tests", "emu", "libb200sa_emu.so"))
rng = np.random.default_rng(int(os.environ.get("SEED", "1")))
groups = {}
t0 = time.time(); it = 0
while time.time() - t0 < float(os.environ.get("SECS", "500")):
    it += 1
    world = int(rng.integers(2, 9))
    n = int(rng.integers(4096 * world, 150000))
    mode = int(rng.integers(3))
    if mode == 0:
        fam = FAMILIES[int(rng.integers(len(FAMILIES)))]
        x = gen(fam, n)
    else:
        sigma = int(rng.integers(1, 5)); fam = "sigma%d" % sigma
        x = rng.integers(0, sigma, size=n, dtype=np.uint8)
        if mode == 2:
            x[n // 2:] = x[: n - n // 2]; fam += "+repeat"
    g = groups.get(world) or groups.setdefault(world, Group([0] * world, library=lib))
    sa, bwt, s = g.suffix_array_and_bwt(x)
    want = o.sa(x)
    assert np.array_equal(sa, want), ("sa", fam, n, world)
    wb, ws = o.bwt_from_sa(x, want)
    assert s == ws and np.array_equal(bwt, wb), ("bwt", fam, n, world)
    b = x.copy()                                   # the same bytes again: served from the resident sharded sort
    assert g.forward_burrows_wheeler_transform(b) == ws and np.array_equal(b, wb), ("bwt reuse", fam, n, world)
    if it % 3 == 0:                                # another text of the same size must not be mistaken for it
        y = x.copy(); y[int(rng.integers(n))] ^= np.uint8(1)
        wy = o.sa(y); wyb, wys = o.bwt_from_sa(y, wy)
        by = y.copy()
        assert g.forward_burrows_wheeler_transform(by) == wys and np.array_equal(by, wyb), ("bwt other", fam, n, world)
    b = bwt.copy(); g.reverse_burrows_wheeler_transform(b, s)
    assert np.array_equal(b, x), ("unbwt", fam, n, world)
    # the text cut at random into blocks, as ONE batch over the group's contexts
    k = int(rng.integers(1, 60))
    cuts = np.sort(rng.integers(0, n + 1, size=k)); cuts = np.concatenate([[0], cuts, [n]])
    blocks = [x[cuts[i]:cuts[i + 1]] for i in range(len(cuts) - 1)]
    sas = g.suffix_array_batch(blocks); bw, sent = g.bwt_batch(blocks)
    for bi, blk in enumerate(blocks):
        if blk.size:
            w = o.sa(blk); assert np.array_equal(sas[bi], w), ("bsa", fam, n, world, bi)
            bb, ss = o.bwt_from_sa(blk, w); assert sent[bi] == ss and np.array_equal(bw[bi], bb), ("bbwt", fam, n, world, bi)
    for bi, blk in enumerate(g.unbwt_batch(bw, sent)): assert np.array_equal(blk, blocks[bi]), ("bunbwt", fam, n, world, bi)
    print(it, fam, n, world, round(time.time() - t0), flush=True)
for g in groups.values():
    g.close()
print("group soak ok", it)
