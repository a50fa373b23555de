#!/bin/bash
# Round-0 key width experiment: fewer key bits = fewer sweeps in round 0, more work in the doubling rounds.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for KB in 64 56 48; do
  echo "== B200SA_MAX_KEY_BITS=$KB"
  B200SA_MAX_KEY_BITS=$KB timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/keybits_$KB.json 2> gpurun_out/keybits_$KB.err
  python tools/bench_summary.py gpurun_out/keybits_$KB.json
done
