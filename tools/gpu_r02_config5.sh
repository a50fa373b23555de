#!/bin/bash
# BASELINE configs[4] sharded over N GPUs: periodic (p = 7) and Fibonacci strings of 2^31 - 2 bytes
N=${1:-8}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for F in ${FAMILIES:-periodic7 fib}; do
  timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/config5_sharded.py $F ${SIZE:-2147483646} 1 \
     > gpurun_out/r02_config5_${F}_n$N.json 2> gpurun_out/r02_config5_${F}_n$N.err || tail -15 gpurun_out/r02_config5_${F}_n$N.err
  grep "^{" gpurun_out/r02_config5_${F}_n$N.json | cut -c1-900
done
