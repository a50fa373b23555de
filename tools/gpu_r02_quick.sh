#!/bin/bash
# quick single-GPU bench (no CPU arm / extras) + the 1 GiB ACGT configuration on one GPU
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --no-facade --no-unbwt > gpurun_out/r02_quick.json 2> gpurun_out/r02_quick.err || tail -5 gpurun_out/r02_quick.err
python tools/bench_summary.py gpurun_out/r02_quick.json
timeout 900 python bench.py --workload acgt_1GiB --steps 3 --warmup 1 --no-cpu-baseline --no-extras --no-facade --no-unbwt > gpurun_out/r02_acgt_1GiB_n1.json 2> gpurun_out/r02_acgt_1GiB_n1.err || tail -5 gpurun_out/r02_acgt_1GiB_n1.err
python tools/bench_summary.py gpurun_out/r02_acgt_1GiB_n1.json
