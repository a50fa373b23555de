#!/bin/bash
# round 2: GPU test tier + one short bench run on one B200
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -2
grep -c processor /proc/cpuinfo; free -g | head -2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/r02_pytest_gpu.log
echo "== bench"
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err || tail -20 gpurun_out/r02_bench.err
python tools/bench_summary.py gpurun_out/r02_bench.json
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/r02_bench.json') if l.startswith('{')][-1])
    for k in ('e2e', 'e2e_facade', 'unbwt', 'cpu_baseline', 'extras'):
        print(k, json.dumps(d.get(k))[:900])
except Exception as e:
    print('no bench line', e)
PY
