#!/bin/bash
# compute-sanitizer memcheck + racecheck over the bundled-row parity run (tools/sanitize_run.py); logs -> gpurun_out/
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  for kinds in "linear gnb kmeans forest" "knn" "svc"; do
    tag=$(echo $kinds | tr ' ' '_')
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py $kinds > gpurun_out/sanitizer_${tool}_${tag}.log 2>&1
    echo "$tool [$kinds] rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run done' gpurun_out/sanitizer_${tool}_${tag}.log | tr '\n' ' ')"
  done
done
