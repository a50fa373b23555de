#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-8}
mkdir -p gpurun_out
echo "== bench sharded peer N=$N"; timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 2 --mode sharded --isa peer 2> gpurun_out/bench_shard_peer_$N.err | grep "^{" | tee gpurun_out/bench_shard_peer_${N}b.json | cut -c1-200
tail -2 gpurun_out/bench_shard_peer_$N.err
