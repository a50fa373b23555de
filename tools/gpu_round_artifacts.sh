#!/bin/bash
# Round artefacts on one B200: GPU test log, the default bench line, the reference arm, the ncu launch list of one step and
# a full-set capture of the dominant kernel.  Everything lands in gpurun_out/ with the given tag.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/${TAG}_pytest_gpu.log; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err || tail -20 gpurun_out/${TAG}_bench_n1.err
python tools/bench_summary.py gpurun_out/${TAG}_bench_n1.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_reference_arm.err
cut -c1-700 gpurun_out/${TAG}_bench_reference_arm.json
BENCH="python bench.py --no-cpu-baseline --no-extras --no-facade --no-unbwt"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
    $BENCH --steps 1 --warmup 1 > gpurun_out/${TAG}_launches.out 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_onesweep_pass -s 10 -c 2 -o gpurun_out/${TAG}_k_onesweep_pass -f \
    $BENCH --steps 1 --warmup 0 > gpurun_out/${TAG}_k_onesweep_pass.out 2>&1
echo "full rc=$?"; ls -la gpurun_out/${TAG}_k_onesweep_pass.ncu-rep
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
