#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/step_ab.py 1000000 4:3 4:5 > gpurun_out/r2_ab7.log 2>&1; grep variant gpurun_out/r2_ab7.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_step3.py -m gpu -q -x > gpurun_out/r2_b5_tests.log 2>&1; tail -5 gpurun_out/r2_b5_tests.log
timeout 200 python scripts/perf_1m.py 1000000 400 > gpurun_out/r2_b5_perf.log 2>&1; tail -16 gpurun_out/r2_b5_perf.log
