#!/bin/bash
# builds compile-time variants on the GPU box and runs the headline bench for each (phases printed)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for V in "$@"; do
  DEFS=$(echo "$V" | tr ',' ' ')
  touch msufsort_b200/csrc/b200sa.cu
  make -s lib NVCC_DEFS="$DEFS" > /dev/null 2>&1 || { echo "build failed: $DEFS"; continue; }
  echo "== [$DEFS]"
  timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | python tools/bench_summary.py | grep -E "value|rerank|sort_pass|isa|bwt|build"
done
