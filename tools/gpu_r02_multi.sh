#!/bin/bash
# round 2: N-GPU checks — sharded parity tests, then bench.py exactly as the driver launches it
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
nvidia-smi topo -m | head -12
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1500 python -m pytest tests/test_sharded_gpu.py tests/test_group.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r02_pytest_multi_n$N.log
fi
echo "== bench --gpus $N (torchrun, default mode)"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-5} --warmup ${WARMUP:-3} \
  > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err || tail -30 gpurun_out/r02_bench_n$N.err
python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/r02_bench_n$N.json') if l.startswith('{')][-1])
    print('value', round(d['value']), d['unit'], 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'], 1), 'launches', d['gpu_launches'])
    print('config', json.dumps(d['config'])[:500])
    print('limiter', d['limiter'])
    print('nvlink', {k: v for k, v in d['nvlink'].items() if k != 'note'})
    for k, v in d['phases_rank0'].items(): print('   %-12s %8.3f ms %5.1f' % (k, v['ms_per_step'], v['launches_per_step']))
    print('sharded_1GiB', json.dumps(d.get('sharded_1GiB'))[:1200])
    print('unbwt', json.dumps(d.get('unbwt'))[:900])
except Exception as e:
    print('no bench line', e)
PY
