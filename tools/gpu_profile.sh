#!/bin/bash
# ncu evidence (B200_PROFILING.md recipe): launch list of one bench step + full-set capture of chosen kernels.
# usage: tools/gpu_profile.sh <tag> [kernel-regex ...]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-prof}; shift
BENCH="python bench.py --no-cpu-baseline --no-extras --no-facade --no-unbwt"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    $BENCH --steps 1 --warmup 1 > gpurun_out/${TAG}_launches.out 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/${TAG}_launches.csv | cut -c1-300
for K in "$@"; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s ${SKIP:-6} -c ${COUNT:-3} -o gpurun_out/${TAG}_$K -f \
      $BENCH --steps 1 --warmup 0 > gpurun_out/${TAG}_$K.out 2>&1
  echo "full $K rc=$?"; ls -la gpurun_out/${TAG}_$K.ncu-rep
done
