#!/bin/bash
# builds tile-shape / ranking variants of the scatter sweep on the GPU box and times each
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for V in "$@"; do
  DEFS=$(echo "$V" | tr ',' ' ')
  touch msufsort_b200/csrc/b200sa.cu
  make -s lib NVCC_DEFS="$DEFS" > /dev/null 2>&1 || { echo "build failed: $DEFS"; continue; }
  R=$(grep -A3 "k_onesweep_passIm" msufsort_b200/lib/ptxas.log | grep -oE "Used [0-9]+ registers|[0-9]+ bytes spill stores" | tr '\n' ' ')
  echo "== $DEFS  [$R]"
  timeout 300 python tools/sort_bench.py 28 64 random 2>&1 | tail -1
done
