#!/bin/bash
# compute-sanitizer over profiles/sanitize_target.py; logs under gpurun_out/ (copy the summaries to profiles/r2/).
set -x
for tool in ${SANITIZE_TOOLS:-memcheck racecheck}; do
  SANITIZE_N=${SANITIZE_N:-40000} timeout 1500 compute-sanitizer --tool $tool --print-limit 20 \
     python profiles/sanitize_target.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/sanitize_$tool.log
  tail -12 gpurun_out/sanitize_$tool.log
done
