#!/bin/bash
# (historical: the first GPU call of round 2.  The persistent sweep it measures was removed afterwards; results: profiles/r02_knobs.txt)
# First GPU call of the next round: the experiments that are correct under the emulator but still unmeasured.
#   B200SA_PACK_RADIX=1     mixed-radix round-0 keys (more symbols per key)
#   B200SA_RS_PERSISTENT=1  persistent sweep with next-tile key prefetch
#   B200SA_LCP_DIRECT=1     budgeted row-wise LCP comparison before the PLCP route
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() {  # name, env...
  local name=$1; shift
  echo "== $name"
  env "$@" timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/knob_$name.json 2> gpurun_out/knob_$name.err || tail -3 gpurun_out/knob_$name.err
  python tools/bench_summary.py gpurun_out/knob_$name.json | head -12
}
run baseline B200SA_DUMMY=0
run pack_radix B200SA_PACK_RADIX=1
run persistent B200SA_RS_PERSISTENT=1
run both B200SA_PACK_RADIX=1 B200SA_RS_PERSISTENT=1
echo "== compile-time variant: within-warp ranks packed two per register (sweep: 80 B of spills -> 0)"
cp msufsort_b200/lib/libb200sa.so /tmp/libb200sa_default.so
if nvcc -gencode arch=compute_100a,code=sm_100a -DB200SA_RS_PACK_POS=1 -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr -shared -o msufsort_b200/lib/libb200sa.so msufsort_b200/csrc/b200sa.cu 2> gpurun_out/knob_packpos_build.err; then
  run packpos B200SA_DUMMY=0
  run packpos_persistent B200SA_RS_PERSISTENT=1
fi
cp /tmp/libb200sa_default.so msufsort_b200/lib/libb200sa.so
echo "== LCP: PLCP route vs budgeted direct route (extras.lcp of the bench line)"
for D in 0 1; do
  B200SA_LCP_DIRECT=$D timeout 200 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/knob_lcp_direct$D.json 2> gpurun_out/knob_lcp_direct$D.err
  python -c "import json; d=json.load(open('gpurun_out/knob_lcp_direct$D.json')); print('LCP_DIRECT=$D', d['extras']['lcp'])"
done
echo "== parity with the knobs on"
B200SA_PACK_RADIX=1 B200SA_RS_PERSISTENT=1 B200SA_LCP_DIRECT=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_batch.py tests/test_lcp.py -m gpu -x -q 2>&1 | tail -3
