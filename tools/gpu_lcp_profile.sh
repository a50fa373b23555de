#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for F in 0 1; do
B200SA_LCP_FUSED=$F timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:'k_plcp|k_lcp|k_scatter_pairs' -c 200 --csv --log-file gpurun_out/lcp_launches_fused$F.csv python tools/lcp_profile.py > gpurun_out/lcp_profile_$F.out 2>&1
tail -2 gpurun_out/lcp_profile_$F.out
done
