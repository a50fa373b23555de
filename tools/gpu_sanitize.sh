#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for TOOL in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $TOOL (default thresholds)"
  B200SA_GROUPSORT_TINY=${TINY:-32} timeout 900 compute-sanitizer --tool $TOOL --print-limit 20 python tools/sanitize_case.py > gpurun_out/sanitize_$TOOL.log 2>&1
  echo "rc=$?"; grep -E "^ok|ERROR SUMMARY|RACECHECK SUMMARY|Error|error:|hazard" gpurun_out/sanitize_$TOOL.log | sort | uniq -c | sort -rn | head -12
done
echo "== memcheck with forced round variants (bucketed ISA, small group thresholds)"
B200SA_ISA_DIRECT_BYTES=0 B200SA_ISA_MIN_UPDATES=1 B200SA_GROUPSORT_TINY=2 B200SA_GROUPSORT_MEDIUM=8 B200SA_GROUPSORT_AVG=1000000 B200SA_UNBWT_CAP_MULT=1 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_case.py > gpurun_out/sanitize_memcheck_variants.log 2>&1
echo "rc=$?"; grep -E "^ok|ERROR SUMMARY|Error|error:" gpurun_out/sanitize_memcheck_variants.log | sort | uniq -c | sort -rn | head -8
