#!/bin/bash
# chain kernel with sector-wise id stores: suite, then stage times (before: c2 chain 2.63 ms, c3 4.11 ms)
mkdir -p gpurun_out
set -o pipefail
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/r02f_pytest.log
rc=${PIPESTATUS[0]}
echo "pytest exit $rc"
[ $rc -ne 0 ] && exit $rc
for C in "c2 1000000" "c3 1000000" "c5 300000"; do python tests/gpu_profile_pipeline.py $C 4 2>&1 | tail -1; done | tee gpurun_out/r02f_stage_times.log
