#!/bin/bash
# compute-sanitizer over the small cases of tools/sanitize_cases.py (run under gpurun, one GPU).
# Usage: tools/run_sanitizers.sh [tag]   -> gpurun_out/sanitizer_<tag>_<tool>_<case>.log + a one-line-per-run summary
tag=${1:-r2}
mkdir -p gpurun_out
sum=gpurun_out/sanitizer_${tag}_summary.txt
: > $sum
for tool in ${SAN_TOOLS:-memcheck racecheck synccheck}; do
  for c in ${SAN_CASES:-c2 c3 c3cl c3r c3rs c4 c4t c5 c5redo}; do
    log=gpurun_out/sanitizer_${tag}_${tool}_${c}.log
    start=$(date +%s)
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py $c > $log 2>&1
    rc=$?
    echo "$tool $c rc=$rc $(( $(date +%s) - start ))s | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1) | $(grep -E ': ok,' $log | tail -1)" | tee -a $sum
  done
done
