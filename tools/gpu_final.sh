#!/bin/bash
# Final single-GPU evidence of the round: smoke, every GPU test, the bench (ours + reference arm).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG=${1:-r01z}
mkdir -p gpurun_out
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/${TAG}_pytest_gpu.log
echo "== bench"; timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "rc=$?"; tail -2 gpurun_out/${TAG}_bench.err; python tools/bench_summary.py gpurun_out/${TAG}_bench.json; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print('extras', json.dumps(d.get('extras')))"
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; echo "rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench_reference.json
