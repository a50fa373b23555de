"""Streaming throughput of the batch path from PINNED host memory: a stream of batches through
(a) one context, one b200sa_bwt_batch call after the other, (b) the pipeline with depth 2 and 3.
    python tools/pipeline_bench.py [--batches 16] [--batch-mib 64] [--block-kib 256]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from msufsort_b200 import textgen as t
    from msufsort_b200.api import Engine, Pipeline
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", type=int, default=16)
    ap.add_argument("--batch-mib", type=int, default=64)
    ap.add_argument("--block-kib", type=int, default=256)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    bsz = a.batch_mib << 20
    blk = a.block_kib << 10
    count = bsz // blk
    offsets = np.arange(count + 1, dtype=np.int64) * blk
    src = torch.from_numpy(t.markov3(bsz * a.batches, t.SEED_MARKOV)).pin_memory()
    work = torch.empty_like(src).pin_memory()
    sent = torch.zeros(a.batches * count, dtype=torch.int32).pin_memory()
    res = {"batches": a.batches, "batch_bytes": bsz, "block_bytes": blk, "blocks_per_batch": count}

    eng = Engine(0)

    def serial():
        for j in range(a.batches):
            eng.lib.check(eng.lib.cdll.b200sa_bwt_batch(eng._ctx, work.data_ptr() + j * bsz, offsets.ctypes.data, count, sent.data_ptr() + 4 * j * count))

    for name, fn in [("serial", serial)]:
        work.copy_(src); fn()  # warm-up: workspace allocation
        work.copy_(src)
        t0 = time.perf_counter(); fn(); dt = time.perf_counter() - t0
        res[name] = {"s": dt, "MBps": bsz * a.batches / dt / 1e6}
        print(name, json.dumps(res[name]), flush=True)
    ref_sent = sent.clone()
    ref_out = work.clone()
    eng.release_workspace()
    for depth in (2, 3):
        with Pipeline(0, depth) as pipe:
            def piped():
                for j in range(a.batches):
                    pipe.submit_bwt_ptr(work.data_ptr() + j * bsz, offsets, sent.data_ptr() + 4 * j * count)
                pipe.drain()
            work.copy_(src); piped()
            work.copy_(src); sent.zero_()
            t0 = time.perf_counter(); piped(); dt = time.perf_counter() - t0
            ok = bool(torch.equal(work, ref_out) and torch.equal(sent, ref_sent))
            res[f"pipeline_depth{depth}"] = {"s": dt, "MBps": bsz * a.batches / dt / 1e6, "identical_to_serial": ok}
            print(f"depth {depth}", json.dumps(res[f"pipeline_depth{depth}"]), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
