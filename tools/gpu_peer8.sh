#!/bin/bash
# ISA in peer memory on N GPUs: the 256 MiB text and the 2^30-2 ACGT text, sharded.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-8}
mkdir -p gpurun_out
echo "== bench sharded peer N=$N"; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 2 --mode sharded --isa peer 2> gpurun_out/bench_shard_peer_$N.err | grep "^{" | tee gpurun_out/bench_shard_peer_$N.json | cut -c1-200
tail -2 gpurun_out/bench_shard_peer_$N.err
echo "== acgt 1GiB sharded peer N=$N"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 2 --warmup 1 --mode sharded --isa peer --workload acgt_1GiB 2> gpurun_out/bench_shard_peer_acgt_$N.err | grep "^{" | tee gpurun_out/bench_shard_peer_acgt_$N.json | cut -c1-200
tail -2 gpurun_out/bench_shard_peer_acgt_$N.err
