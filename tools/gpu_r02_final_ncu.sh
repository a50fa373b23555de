#!/bin/bash
# The last 45 GPU-seconds of round 2: ncu --set full of the dominant kernel as it is built NOW (four CTAs per SM), and the launch
# list of one suffix-array call, both through the command line tool (no Python / torch start-up inside the budget).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 10 python -c "
import sys; sys.path.insert(0, '.')
from msufsort_b200 import textgen as t
t.markov3(1 << 28, t.SEED_MARKOV).tofile('/dev/shm/t.bin')" || exit 1
timeout 22 ncu --set full --clock-control none --import-source on -k regex:k_onesweep_pass -s 1 -c 1 -o gpurun_out/r02j_k_onesweep_pass -f \
    msufsort_b200/lib/msufsort s /dev/shm/t.bin > gpurun_out/r02j_full.out 2>&1
echo "full rc=$?"
timeout 9 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r02j_launches.csv \
    msufsort_b200/lib/msufsort s /dev/shm/t.bin > gpurun_out/r02j_launches.out 2>&1
echo "list rc=$?"
tail -n 3 gpurun_out/r02j_full.out
