#!/bin/bash
# Round-2 second half, one gpurun call: the whole GPU test suite, then (only if it is green) the
# bench line, the decode timings, the launch list and a full ncu capture of the reworked kernels.
mkdir -p gpurun_out
set -o pipefail
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/r02b_pytest.log
rc=${PIPESTATUS[0]}
echo "pytest exit $rc"
[ $rc -ne 0 ] && exit $rc
python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench_c2_n1.json 2> gpurun_out/r02b_bench_c2_n1.err
echo "bench exit $?"; head -c 1500 gpurun_out/r02b_bench_c2_n1.json; echo
python tests/gpu_bench_decode.py > gpurun_out/r02b_decode.log 2>&1; tail -4 gpurun_out/r02b_decode.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tmgpu|DeviceScan|DeviceRadixSort" -c 400 --csv \
    --log-file gpurun_out/r02b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r02b_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"table_kernel|chain_kernel" -s 2 -c 2 -o gpurun_out/r02b_pipeline_full \
    python tests/gpu_profile_pipeline.py c2 1000000 2 > gpurun_out/r02b_pipeline_full.log 2>&1
tail -1 gpurun_out/r02b_pipeline_full.log
ncu --set full --clock-control none --import-source on -k regex:"capcode_decode_scan|decode_copy|compact_bytes" -c 3 -o gpurun_out/r02b_decode_full \
    python tests/gpu_bench_decode.py 1000000 > gpurun_out/r02b_decode_full.log 2>&1
tail -2 gpurun_out/r02b_decode_full.log
