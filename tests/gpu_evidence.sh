#!/bin/bash
# Evidence run (1 GPU, under gpurun).  Outputs go to gpurun_out/; what is kept is copied to profiles/.
#   bash tests/gpu_evidence.sh            suite, bench line, decode timings, launch list, full captures
#   bash tests/gpu_evidence.sh sanitize   compute-sanitizer (memcheck / initcheck / racecheck / synccheck)
#   bash tests/gpu_evidence.sh smoke      smoke() under ncu, as the driver runs it
# (Round 2 ran these pieces as six separate calls, r02b .. r02g in profiles/, to fit the budget.)
mkdir -p gpurun_out
set -o pipefail
TAG=${TAG:-r02}
if [ "$1" = "sanitize" ]; then
  for tool in memcheck initcheck racecheck synccheck; do
    echo "== compute-sanitizer --tool $tool"
    timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/${TAG}_sanitizer_$tool.log python tests/gpu_sanitize.py > gpurun_out/${TAG}_sanitizer_$tool.out 2>&1
    tail -1 gpurun_out/${TAG}_sanitizer_$tool.out
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/${TAG}_sanitizer_$tool.log | tail -1
  done
  exit 0
fi
if [ "$1" = "smoke" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_smoke_launches.csv \
      python -c "import __graft_entry__ as g; g.smoke(); print('__SMOKE_OK__')" > gpurun_out/${TAG}_smoke_under_ncu.log 2>&1
  tail -4 gpurun_out/${TAG}_smoke_under_ncu.log
  exit 0
fi
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/${TAG}_pytest_gpu.log
rc=${PIPESTATUS[0]}
echo "pytest exit $rc"
[ $rc -ne 0 ] && exit $rc
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_c2_n1.json 2> gpurun_out/${TAG}_bench_c2_n1.err
head -c 600 gpurun_out/${TAG}_bench_c2_n1.json; echo
python tests/gpu_bench_decode.py > gpurun_out/${TAG}_decode.log 2>&1; tail -2 gpurun_out/${TAG}_decode.log
for C in "c3 1000000" "c5 300000"; do python tests/gpu_profile_pipeline.py $C 4 2>&1 | tail -1; done | tee gpurun_out/${TAG}_stage_times.log
echo "== launch list of the bench command"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"kernel|DeviceScan|DeviceRadixSort" -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
echo "== full captures: table + chain (c2, 1M strings), decode kernels"
ncu --set full --clock-control none --import-source on -k regex:"table_kernel|chain_kernel" -s 2 -c 2 -o gpurun_out/${TAG}_pipeline_full \
    python tests/gpu_profile_pipeline.py c2 1000000 2 > gpurun_out/${TAG}_pipeline_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"capcode_decode_scan|decode_copy|compact_bytes|decode_size" -c 4 -o gpurun_out/${TAG}_decode_full \
    python tests/gpu_bench_decode.py 1000000 > gpurun_out/${TAG}_decode_full.log 2>&1
# then, here:  python tools/ncu_kernels.py gpurun_out/${TAG}_pipeline_full.ncu-rep > profiles/${TAG}_pipeline_raw.txt
#              ncu -i ... --page source --csv --kernel-name regex:table_kernel > src.csv; python tools/ncu_by_line.py <mangled name> src.csv
