#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for TOOL in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $TOOL (LCP + batch)"
  B200SA_UNBWT_CAP_MULT=1 timeout 400 compute-sanitizer --tool $TOOL --print-limit 20 python tools/sanitize_features.py > gpurun_out/sanitize_features_$TOOL.log 2>&1
  echo "rc=$?"; grep -E "^ok|ERROR SUMMARY|RACECHECK SUMMARY|Error|error:|hazard" gpurun_out/sanitize_features_$TOOL.log | sort | uniq -c | sort -rn | head -12
done
