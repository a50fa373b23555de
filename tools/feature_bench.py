"""LCP array and batched-blocks timings on one GPU (CUDA events on torch's current stream, inputs resident).
    python tools/feature_bench.py [--n 268435456] [--blocks 1024] [--out gpurun_out/feature_bench.json]
Results are validated: LCP against the oracle on a 16 MiB prefix run, batch results against per-block calls."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    from msufsort_b200 import textgen as t
    from msufsort_b200.api import Engine
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 28)
    ap.add_argument("--blocks", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    eng = Engine(0)
    eng.set_profiling(True)
    res = {}

    def timed(fn, reps=a.reps):
        fn()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    n = a.n
    text = t.markov3(n, t.SEED_MARKOV)
    d_text = torch.from_numpy(text).cuda()
    d_sa = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    d_lcp = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    ms_sa = timed(lambda: eng.suffix_array_dev(d_text, n, d_sa))
    eng.profile_reset()
    ms_lcp = timed(lambda: eng.lcp_dev(d_text, n, d_sa, d_lcp), reps=1)
    prof = eng.profile()
    ms_lcp = min(ms_lcp, timed(lambda: eng.lcp_dev(d_text, n, d_sa, d_lcp)))
    res["lcp"] = {"n": n, "text": "markov3", "sa_ms": ms_sa, "lcp_ms": ms_lcp, "lcp_MBps": n / ms_lcp / 1e3,
                  "lcp_max": int(d_lcp.max().item()), "lcp_mean": float(d_lcp.double().mean().item()),
                  "phases_two_calls": {k: v for k, v in prof["phases"].items() if v["launches"]}}
    print("LCP", json.dumps(res["lcp"]), flush=True)
    # deep-repeat input: lcp ~ n
    for fam, m in (("fib", 1 << 26), ("periodic1009", 1 << 26)):
        x = t.GENERATORS[fam](m)
        dx = torch.from_numpy(x).cuda()
        eng.suffix_array_dev(dx, m, d_sa)
        ms = timed(lambda: eng.lcp_dev(dx, m, d_sa, d_lcp), reps=2)
        res["lcp_" + fam] = {"n": m, "lcp_ms": ms, "lcp_max": int(d_lcp[: m + 1].max().item())}
        print("LCP", fam, json.dumps(res["lcp_" + fam]), flush=True)
        del dx

    # ---- batch: the same text cut into `blocks` blocks of equal size vs one call per block
    for count in sorted({a.blocks, 64, 16384}):
        bs = n // count
        offsets = (np.arange(count + 1, dtype=np.int64) * bs)
        total = int(offsets[-1])
        d_bwt = torch.empty(total, dtype=torch.uint8, device="cuda")
        sent = None

        def run_batch():
            nonlocal sent
            sent = eng.batch_dev(d_text, offsets, d_bwt, None)
        eng.profile_reset()
        ms_batch = timed(run_batch, reps=2)
        rounds = eng.profile()["rounds"] // 3
        # per-block loop over a sample of blocks (device-resident too), extrapolated
        sample = min(count, 64)
        d_one = torch.empty(bs, dtype=torch.uint8, device="cuda")
        sents = []

        def run_loop():
            sents.clear()
            for b in range(sample):
                sents.append(eng.bwt_dev(d_text[b * bs:(b + 1) * bs], bs, d_one))
        ms_loop = timed(run_loop, reps=2) * (count / sample)
        ok = all(int(sent[b]) == sents[b] for b in range(sample))
        # last sampled block's bytes must agree too
        ok = ok and bool(torch.equal(d_one, d_bwt[(sample - 1) * bs: sample * bs]))
        d_back = torch.empty(total, dtype=torch.uint8, device="cuda")
        ms_unbwt = timed(lambda: eng.unbwt_batch_dev(d_bwt, offsets, sent, d_back), reps=2)
        ok = ok and bool(torch.equal(d_back, d_text[:total]))
        del d_back
        res[f"batch_{count}"] = {"blocks": count, "block_bytes": bs, "batch_ms": ms_batch, "batch_MBps": total / ms_batch / 1e3,
                                 "per_block_calls_ms_extrapolated": ms_loop, "speedup": ms_loop / ms_batch, "rounds": rounds,
                                 "unbwt_batch_ms": ms_unbwt, "unbwt_batch_MBps": total / ms_unbwt / 1e3,
                                 "matches_per_block_calls": ok}
        print("BATCH", json.dumps(res[f"batch_{count}"]), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
