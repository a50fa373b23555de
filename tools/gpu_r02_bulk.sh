#!/bin/bash
# sweep with bulk asynchronous tile loads (cp.async.bulk + mbarrier) against the default sweep, same box, same run
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() {
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --no-facade --no-unbwt > gpurun_out/r02_sweep_$1.json 2> gpurun_out/r02_sweep_$1.err || tail -5 gpurun_out/r02_sweep_$1.err
  python - <<PY
import json
d = json.loads([l for l in open('gpurun_out/r02_sweep_$1.json') if l.startswith('{')][-1])
r = d['roofline']
print('$1', 'step', round(d['ms_per_step'], 2), 'sweep ms/launch', round(r['avg_launch_ms'], 4), 'frac', round(r['frac'], 3), 'isa', round(d['phases']['isa']['ms_per_step'], 3), 'pack', round(d['phases']['pack']['ms_per_step'], 3))
PY
}
run default
cp msufsort_b200/lib/libb200sa.so /tmp/libb200sa_default.so
if nvcc -gencode arch=compute_100a,code=sm_100a -DB200SA_RS_BULK_LOAD=1 -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr -shared -o msufsort_b200/lib/libb200sa.so msufsort_b200/csrc/b200sa.cu 2> gpurun_out/r02_sweep_bulk_build.err; then
  run bulk
  cuobjdump -sass msufsort_b200/lib/libb200sa.so | grep -E "UBLKCP|SYNCS" | sed 's/\/\*[0-9a-f]*\*\///' | sort | uniq -c > gpurun_out/r02_sweep_bulk_sass.txt
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
fi
cp /tmp/libb200sa_default.so msufsort_b200/lib/libb200sa.so
