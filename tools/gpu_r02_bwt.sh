#!/bin/bash
# BWT routes on one B200: row-order gather, text-order scatter with byte stores, with 32-bit OR reductions
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for V in "gather B200SA_BWT_SCATTER_MIN=99999999999" "scatter_bytes B200SA_BWT_SCATTER_BYTES=1" "scatter_words B200SA_DUMMY=1"; do
  set -- $V
  env $2 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras --no-facade --no-unbwt > gpurun_out/r02_bwt_$1.json 2> gpurun_out/r02_bwt_$1.err || tail -5 gpurun_out/r02_bwt_$1.err
  python - <<PY
import json
d = json.loads([l for l in open('gpurun_out/r02_bwt_$1.json') if l.startswith('{')][-1])
print('$1', 'step', round(d['ms_per_step'], 2), 'bwt', d['phases']['bwt'], 'e2e', round(d['e2e']['ms_per_step'], 1))
PY
done
