#!/bin/bash
# one short gpurun call: A/B of the step-kernel variants at 1M beads (bit-identity + timings), then the GPU parity
# tests of the step kernels and the core suites under the fastest identical variant
mkdir -p gpurun_out
rm -f gpurun_out/step_ab.txt gpurun_out/step_ab_best.txt
timeout ${AB_TIMEOUT:-70} python scripts/step_ab.py 1000000 > gpurun_out/step_ab.log 2>&1; echo "ab rc=$?"; tail -16 gpurun_out/step_ab.log
BEST=$(cat gpurun_out/step_ab_best.txt 2>/dev/null || echo 0)
LE_STEP_VARIANT=$BEST timeout ${PT_TIMEOUT:-60} python -m pytest tests/test_gpu_step2.py tests/test_gpu_parity.py tests/test_gpu_md.py -x -q > gpurun_out/pytest_best.log 2>&1
echo "pytest(step variant $BEST) rc=$?"; tail -4 gpurun_out/pytest_best.log
