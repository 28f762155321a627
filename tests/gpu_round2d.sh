#!/bin/bash
# Third GPU call: table-size experiment (TMGPU_DFA_PER_BUCKET), chunk-size check, c3 bench line.
mkdir -p gpurun_out
for P in 2 1; do
  for C in "c2 1000000" "c3 1000000" "c5 300000"; do
    echo "== TMGPU_DFA_PER_BUCKET=$P $C"
    TMGPU_DFA_PER_BUCKET=$P python tests/gpu_profile_pipeline.py $C 4 2>&1 | tail -1
  done
done > gpurun_out/r02d_dfa_per_bucket.log 2>&1
cat gpurun_out/r02d_dfa_per_bucket.log
for L in 1280 1024; do
  echo "== TMGPU_CHUNK_LANES=$L"
  TMGPU_CHUNK_LANES=$L python tests/gpu_e2e_trace.py c2 1000000 2>&1 | grep -E "call [123]:" | tail -6
done > gpurun_out/r02d_chunk_lanes.log 2>&1
cat gpurun_out/r02d_chunk_lanes.log
python bench.py --config c3 --steps 10 --warmup 3 > gpurun_out/r02d_bench_c3_n1.json 2> gpurun_out/r02d_bench_c3_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02d_bench_c3_n1.json'))
print('c3', d['value'], d['ms_per_step'], d['roofline']['stages_ms'], 'e2e', d['e2e']['value'], 'decode', d['decode']['ms'], d['parity']['match'])
PY
