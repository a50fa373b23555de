#!/bin/bash
# sweep with four resident CTAs per SM (16-bit per-warp counters, 64 registers) against the default (three CTAs, 80 registers)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() {
  timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --no-facade --no-unbwt > gpurun_out/r02_4cta_$1.json 2>/dev/null
  python - <<PY
import json
d = json.loads([l for l in open('gpurun_out/r02_4cta_$1.json') if l.startswith('{')][-1])
r = d['roofline']
print('$1', 'step', round(d['ms_per_step'], 2), 'sweep ms/launch', round(r['avg_launch_ms'], 4), 'frac', round(r['frac'], 3), 'isa', round(d['phases']['isa']['ms_per_step'], 3))
PY
}
cp msufsort_b200/lib/libb200sa.so /tmp/libb200sa_default.so
nvcc -gencode arch=compute_100a,code=sm_100a -DB200SA_RS_WHIST_U16=1 -DB200SA_RS_MIN_BLOCKS=4 -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr -shared -o /tmp/libb200sa_4cta.so msufsort_b200/csrc/b200sa.cu 2> gpurun_out/r02_4cta_build.err &
run default
wait
cp /tmp/libb200sa_4cta.so msufsort_b200/lib/libb200sa.so
run four_ctas
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
cp /tmp/libb200sa_default.so msufsort_b200/lib/libb200sa.so
