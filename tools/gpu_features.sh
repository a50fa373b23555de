#!/bin/bash
# LCP + batched-blocks evidence on one B200: their GPU tests, then the feature bench.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG=${1:-r01f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lcp.py tests/test_batch.py -m gpu -x -q --timeout 300 > gpurun_out/${TAG}_pytest_features.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest_features.log
timeout 600 python tools/feature_bench.py --out gpurun_out/${TAG}_feature_bench.json 2>&1 | tail -12
