#!/bin/bash
# multi-GPU checks under `gpurun --gpus N`: sharded NCCL tests, weak-scaling bench, sharded bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}
mkdir -p gpurun_out
echo "== sharded nccl tests"; timeout 600 python -m pytest tests/test_sharded_gpu.py -x -q --timeout 200 2>&1 | tail -5
echo "== bench independent N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 2 2> gpurun_out/bench_ind_$N.err | tee gpurun_out/bench_ind_$N.json | tail -1 | cut -c1-600
tail -3 gpurun_out/bench_ind_$N.err
echo "== bench sharded N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 2 --mode sharded 2> gpurun_out/bench_shard_$N.err | tee gpurun_out/bench_shard_$N.json | tail -1 | cut -c1-1500
tail -3 gpurun_out/bench_shard_$N.err
echo "== bench sharded replicated N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 3 --warmup 2 --mode sharded --isa replicated 2> gpurun_out/bench_shardrep_$N.err | tee gpurun_out/bench_shardrep_$N.json | tail -1 | cut -c1-300
echo "== unbwt sharded N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/unbwt_sharded_bench.py 1073741822 2 2>gpurun_out/unbwt_shard_$N.err | tee gpurun_out/unbwt_shard_$N.json | tail -1
tail -2 gpurun_out/unbwt_shard_$N.err
