#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}; shift
for SPEC in "$@"; do
  set -- $(echo $SPEC | tr ',' ' ')
  echo "== $1 $2 $3"
  CUDA_LAUNCH_BLOCKING=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 tools/shard_debug.py $1 $2 $3 2>&1 | grep -vE "^\*|OMP_NUM|^$|frozen|torch/distributed|^  File \"/opt|elastic|^    |^Traceback|^=====|Root Cause|time |host |rank  |exitcode|error_file|traceback :|^\[1\]|^\[0\]|^-----" | tail -14
done
