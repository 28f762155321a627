#!/bin/bash
# compute-sanitizer over every kernel (small batches), after the kernel rework of round 2b
mkdir -p gpurun_out
for tool in memcheck initcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 400 compute-sanitizer --tool $tool --log-file gpurun_out/r02e_sanitizer_$tool.log python tests/gpu_sanitize.py > gpurun_out/r02e_sanitizer_$tool.out 2>&1
  tail -1 gpurun_out/r02e_sanitizer_$tool.out
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r02e_sanitizer_$tool.log | tail -1
  grep -E "Invalid|Uninitialized|hazard|Race" gpurun_out/r02e_sanitizer_$tool.log | sort | uniq -c | head -5
done
