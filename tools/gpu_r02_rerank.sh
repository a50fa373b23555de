#!/bin/bash
# k_rerank variants: suffix prefetch before the look-back, CTAs per SM
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
cp msufsort_b200/lib/libb200sa.so /tmp/libb200sa_default.so
for V in "base -DB200SA_RR_PREFETCH=0" "prefetch6 -DB200SA_RR_PREFETCH=1" "prefetch5 -DB200SA_RR_PREFETCH=1 -DB200SA_RR_MIN_BLOCKS=5" "prefetch4 -DB200SA_RR_PREFETCH=1 -DB200SA_RR_MIN_BLOCKS=4" "base5 -DB200SA_RR_PREFETCH=0 -DB200SA_RR_MIN_BLOCKS=5"; do
  set -- $V; NAME=$1; shift
  nvcc -gencode arch=compute_100a,code=sm_100a $@ -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr -shared -o msufsort_b200/lib/libb200sa.so msufsort_b200/csrc/b200sa.cu 2> gpurun_out/r02_rr_$NAME.build || { tail -3 gpurun_out/r02_rr_$NAME.build; continue; }
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --no-facade --no-unbwt > gpurun_out/r02_rr_$NAME.json 2> gpurun_out/r02_rr_$NAME.err || tail -5 gpurun_out/r02_rr_$NAME.err
  python - <<PY
import json
d = json.loads([l for l in open('gpurun_out/r02_rr_$NAME.json') if l.startswith('{')][-1])
print('$NAME', 'step', round(d['ms_per_step'], 3), 'rerank', round(d['phases']['rerank']['ms_per_step'], 3))
PY
done
cp /tmp/libb200sa_default.so msufsort_b200/lib/libb200sa.so
