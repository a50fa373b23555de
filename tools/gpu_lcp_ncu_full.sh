#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 80 ncu --set full --clock-control none --import-source on -k regex:'k_plcp_level|k_lcp_gather' -s 15 -c 3 -o gpurun_out/r01_lcp_full -f python tools/lcp_profile.py 67108864 > gpurun_out/r01_lcp_full.out 2>&1
echo "rc=$?"; tail -2 gpurun_out/r01_lcp_full.out; ls -la gpurun_out/r01_lcp_full.ncu-rep
