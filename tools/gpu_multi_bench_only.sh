#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-8}
mkdir -p gpurun_out
for ISA in owner replicated; do
echo "== bench sharded $ISA N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 2 --mode sharded --isa $ISA 2> gpurun_out/bench_shard_${ISA}_$N.err | grep "^{" | tee gpurun_out/bench_shard_${ISA}_$N.json | cut -c1-200
done
echo "== unbwt sharded N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/unbwt_sharded_bench.py 1073741822 2 2>gpurun_out/unbwt_shard_$N.err | tee gpurun_out/unbwt_shard_$N.json | tail -1
echo "== acgt 1GiB sharded owner N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 2 --warmup 1 --mode sharded --isa owner --workload acgt_1GiB 2> gpurun_out/bench_shard_acgt_$N.err | grep "^{" | tee gpurun_out/bench_shard_acgt_$N.json | cut -c1-200
