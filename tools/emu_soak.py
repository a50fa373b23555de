"""Randomised soak of the kernel logic under the CPU emulator (test infrastructure): random families and sizes up to 400 k,
SA + BWT + LCP + inverse on the whole text, then a random cut into blocks through the batch entry points; every result
is compared with the oracle.  Knobs come from the environment (B200SA_*), duration and seed from SECS / SEED.
    SECS=600 SEED=3 B200SA_PACK_RADIX=1 python tools/emu_soak.py"""
import sys, os, numpy as np, time
ROOT_ = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT_); sys.path.insert(0, os.path.join(ROOT_, "tests"))
from conftest import Oracle, ROOT
from cases import FAMILIES, gen
from msufsort_b200.api import Engine, Library
o = Oracle()
eng = Engine(0, library=Library(os.path.join(ROOT,'tests','emu','libb200sa_emu.so')))
rng = np.random.default_rng(int(os.environ.get("SEED","1")))
t0=time.time(); it=0
while time.time()-t0 < float(os.environ.get("SECS","500")):
    it+=1
    fam = FAMILIES[int(rng.integers(len(FAMILIES)))]
    n = int(rng.integers(1, 400000))
    x = gen(fam, n)
    sa = o.sa(x)
    # single text: SA+BWT, LCP, wide
    s1, b1, z1 = eng.suffix_array_and_bwt(x)
    assert np.array_equal(s1, sa), ("sa", fam, n)
    wb, ws = o.bwt_from_sa(x, sa); assert z1==ws and np.array_equal(b1, wb), ("bwt", fam, n)
    lcp = eng.make_lcp_array(x, sa); assert np.array_equal(lcp, o.lcp(x, sa, kasai=True)), ("lcp", fam, n)
    back = b1.copy(); eng.reverse_burrows_wheeler_transform(back, z1); assert np.array_equal(back, x), ("unbwt", fam, n)
    # batch: random cut of the text into blocks (+ duplicates, empties)
    k = int(rng.integers(1, 40))
    cuts = np.sort(rng.integers(0, n+1, size=k)); cuts = np.concatenate([[0], cuts, [n]])
    blocks = [x[cuts[i]:cuts[i+1]] for i in range(len(cuts)-1)]
    blocks += [blocks[0], np.empty(0,np.uint8)]
    sas = eng.suffix_array_batch(blocks); bw, sent = eng.bwt_batch(blocks)
    for b, blk in enumerate(blocks):
        if blk.size:
            w = o.sa(blk); assert np.array_equal(sas[b], w), ("bsa", fam, n, b)
            bb, ss = o.bwt_from_sa(blk, w); assert sent[b]==ss and np.array_equal(bw[b], bb), ("bbwt", fam, n, b)
    bk = eng.unbwt_batch(bw, sent)
    for b, blk in enumerate(blocks): assert np.array_equal(bk[b], blk), ("bunbwt", fam, n, b)
    print(it, fam, n, len(blocks), round(time.time()-t0), flush=True)
print("soak ok", it)
