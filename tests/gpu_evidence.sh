#!/bin/bash
# Evidence run (1 GPU, under gpurun): smoke() under ncu, compute-sanitizer over every kernel,
# launch list and full ncu capture of the tokenize pipeline.  Outputs go to gpurun_out/.
mkdir -p gpurun_out
echo "== smoke() under ncu (the driver's third execution of smoke() in round 1 decoded wrong bytes)"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_smoke_launches.csv \
    python -c "import __graft_entry__ as g; g.smoke(); print('__SMOKE_OK__')" > gpurun_out/r02_smoke_under_ncu.log 2>&1
tail -4 gpurun_out/r02_smoke_under_ncu.log
for tool in memcheck initcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/r02_sanitizer_$tool.log python tests/gpu_sanitize.py > gpurun_out/r02_sanitizer_$tool.out 2>&1
  tail -1 gpurun_out/r02_sanitizer_$tool.out
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r02_sanitizer_$tool.log | tail -1
done
echo "== launch list of the bench command"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tmgpu|DeviceScan|DeviceRadixSort" -c 400 --csv \
    --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r02_bench_under_ncu.log 2>&1
echo "== full capture: table + chain, c2 1M strings"
ncu --set full --clock-control none --import-source on -k regex:"table_kernel|chain_kernel" -s 2 -c 2 -o gpurun_out/r02_final_full \
    python tests/gpu_profile_pipeline.py c2 1000000 2 > gpurun_out/r02_final_full.log 2>&1
tail -1 gpurun_out/r02_final_full.log
