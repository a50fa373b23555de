#!/bin/bash
# ISA in peer memory on N GPUs: NCCL tests of the peer mode, then sharded benches peer vs owner.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}
TESTS=${2:-yes}
mkdir -p gpurun_out
if [ "$TESTS" = yes ]; then
echo "== sharded peer tests"; timeout 500 python -m pytest tests/test_sharded_gpu.py -x -q --timeout 200 -k "peer and (markov3 or zeros or fib)" 2>&1 | tail -5
fi
for ISA in peer owner; do
echo "== bench sharded $ISA N=$N"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 2 --mode sharded --isa $ISA 2> gpurun_out/bench_shard_${ISA}_$N.err | grep "^{" | tee gpurun_out/bench_shard_${ISA}_$N.json | cut -c1-260
tail -2 gpurun_out/bench_shard_${ISA}_$N.err
done
