#!/bin/bash
# windowed BWT gather: window size sweep (0 = one pass)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() {  # workload window steps
  B200SA_BWT_WINDOW_BYTES=$2 timeout 300 python bench.py --workload $1 --steps $3 --warmup 2 --no-cpu-baseline --no-extras --no-facade --no-unbwt > gpurun_out/r02_bwtwin_$1_$2.json 2>/dev/null
  python - <<PY
import json
d = json.loads([l for l in open('gpurun_out/r02_bwtwin_$1_$2.json') if l.startswith('{')][-1])
print('$1 window $2', 'step', round(d['ms_per_step'], 2), 'bwt', round(d['phases']['bwt']['ms_per_step'], 3), 'launches', d['phases']['bwt']['launches_per_step'], 'e2e', round(d['e2e']['ms_per_step'], 1))
PY
}
for W in 0 33554432 50331648 67108864 100663296; do run markov3_256MiB $W 5; done
for W in 0 67108864 134217728; do run acgt_1GiB $W 2; done
