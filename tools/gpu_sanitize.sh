#!/usr/bin/env bash
# compute-sanitizer over the small kernel cases (tools/sanitize_cases.py); summaries -> gpurun_out/sanitize_*.txt
set -x
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 40 python tools/sanitize_cases.py $([ $tool = memcheck ] && echo --mem) \
    > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/sanitize_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok |^FAIL|ALL OK|FAILURES|Error|hazard" gpurun_out/sanitize_$tool.log | sort | uniq -c | sort -rn | head -40 > gpurun_out/sanitize_$tool.txt
  cat gpurun_out/sanitize_$tool.txt
done
