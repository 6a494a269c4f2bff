#!/bin/bash
# compute-sanitizer over small runs of every kernel family (scratch/sanitize.py)
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 10 python scratch/sanitize.py > gpurun_out/r02t_$tool.txt 2>&1
  echo "$tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02t_$tool.txt | tail -1)"
done
