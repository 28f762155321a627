#!/bin/bash
# Last GPU call of the round: suite, final bench line, stage times, full capture, launch list.
mkdir -p gpurun_out
set -o pipefail
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/r02g_pytest.log
rc=${PIPESTATUS[0]}
echo "pytest exit $rc"
[ $rc -ne 0 ] && exit $rc
python bench.py --steps 10 --warmup 3 > gpurun_out/r02g_bench_c2_n1.json 2> gpurun_out/r02g_bench_c2_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02g_bench_c2_n1.json'))
print(d['value'], d['ms_per_step'], d['roofline']['stages_ms'], 'e2e', d['e2e']['value'], d['e2e']['pageable']['value'], 'decode', d['decode']['ms'], d['parity']['match'])
PY
for C in "c3 1000000" "c5 300000"; do python tests/gpu_profile_pipeline.py $C 4 2>&1 | tail -1; done | tee gpurun_out/r02g_stage_times.log
ncu --set full --clock-control none --import-source on -k regex:"table_kernel|chain_kernel" -s 2 -c 2 -o gpurun_out/r02g_pipeline_full \
    python tests/gpu_profile_pipeline.py c2 1000000 2 > gpurun_out/r02g_pipeline_full.log 2>&1
tail -1 gpurun_out/r02g_pipeline_full.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tmgpu|DeviceScan|DeviceRadixSort" -c 400 --csv \
    --log-file gpurun_out/r02g_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r02g_bench_under_ncu.log 2>&1
