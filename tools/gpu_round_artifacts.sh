#!/bin/bash
# Collects the round's evidence on one B200: tests, bench (ours + reference arm), ncu launch list + full capture, ubench.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; echo "ref rc=$?"
timeout 600 python bench.py --workload rand_16MiB --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_rand16.json 2>> gpurun_out/${TAG}_bench.err
./tools/ubench/ubench > gpurun_out/${TAG}_ubench.txt 2>&1
timeout 600 python tools/sort_bench.py 28 64 random > gpurun_out/${TAG}_sort_bench.txt 2>&1
timeout 600 python tools/unbwt_bench.py 1073741822 2 > gpurun_out/${TAG}_unbwt_1GiB.txt 2>&1
bash tools/gpu_profile.sh ${TAG} k_onesweep_pass k_rerank
cat gpurun_out/${TAG}_bench.json | cut -c1-400
