#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_group.py -m gpu -x -q 2>&1 | tail -2
for W in 100663296 67108864 0; do
  B200SA_BWT_WINDOW_BYTES=$W timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --no-facade --no-unbwt > gpurun_out/r02_bwtvec_$W.json 2>/dev/null
  python - <<PY
import json
d = json.loads([l for l in open('gpurun_out/r02_bwtvec_$W.json') if l.startswith('{')][-1])
print('window $W', 'step', round(d['ms_per_step'], 2), 'bwt', round(d['phases']['bwt']['ms_per_step'], 3), 'launches', d['phases']['bwt']['launches_per_step'], 'e2e', round(d['e2e']['ms_per_step'], 1))
PY
done
B200SA_BWT_MAX_PASSES=16 B200SA_BWT_WINDOW_BYTES=100663296 timeout 300 python bench.py --workload acgt_1GiB --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-facade --no-unbwt > gpurun_out/r02_bwtvec_1g.json 2>/dev/null
python - <<PY
import json
d = json.loads([l for l in open('gpurun_out/r02_bwtvec_1g.json') if l.startswith('{')][-1])
print('acgt_1GiB 96MiB windows', 'step', round(d['ms_per_step'], 2), 'bwt', round(d['phases']['bwt']['ms_per_step'], 3), 'launches', d['phases']['bwt']['launches_per_step'])
PY
timeout 300 python bench.py --workload acgt_1GiB --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-facade --no-unbwt > gpurun_out/r02_bwtvec_1g0.json 2>/dev/null
python - <<PY
import json
d = json.loads([l for l in open('gpurun_out/r02_bwtvec_1g0.json') if l.startswith('{')][-1])
print('acgt_1GiB default', 'step', round(d['ms_per_step'], 2), 'bwt', round(d['phases']['bwt']['ms_per_step'], 3), 'launches', d['phases']['bwt']['launches_per_step'])
PY
