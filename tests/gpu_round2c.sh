#!/bin/bash
# Second GPU call of the session: the suite on the new decode kernels, the bench line, the host-path
# chunk-size experiment (TMGPU_CHUNK_LANES), ncu capture of the decode kernels.
mkdir -p gpurun_out
set -o pipefail
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/r02c_pytest.log
rc=${PIPESTATUS[0]}
echo "pytest exit $rc"
[ $rc -ne 0 ] && exit $rc
python bench.py --steps 10 --warmup 3 > gpurun_out/r02c_bench_c2_n1.json 2> gpurun_out/r02c_bench_c2_n1.err
echo "bench exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02c_bench_c2_n1.json'))
print(d['value'], d['ms_per_step'], d['roofline']['stages_ms'], 'e2e', d['e2e']['value'], d['e2e']['pageable']['value'], 'decode', d['decode']['ms'], d['parity']['match'])
PY
python tests/gpu_bench_decode.py > gpurun_out/r02c_decode.log 2>&1; tail -3 gpurun_out/r02c_decode.log
for L in 2048 1536 1280; do
  echo "== TMGPU_CHUNK_LANES=$L"
  TMGPU_CHUNK_LANES=$L python tests/gpu_e2e_trace.py c2 1000000 2>&1 | grep -E "call [23]|chunk" | tail -8
done > gpurun_out/r02c_chunk_lanes.log 2>&1
cat gpurun_out/r02c_chunk_lanes.log
ncu --set full --clock-control none --import-source on -k regex:"capcode_decode_scan|decode_copy|compact_bytes|decode_size" -c 4 -o gpurun_out/r02c_decode_full \
    python tests/gpu_bench_decode.py 1000000 > gpurun_out/r02c_decode_full.log 2>&1
tail -2 gpurun_out/r02c_decode_full.log
