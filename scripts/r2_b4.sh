#!/bin/bash
# k_build4 first run: list/force parity tests, then per-kernel timing at 1M beads for both list builds, then the slab tests on one device
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_step3.py -m gpu -q -x > gpurun_out/r2_b4_tests.log 2>&1; tail -5 gpurun_out/r2_b4_tests.log
LE_BUILD_VARIANT=4 timeout 200 python scripts/perf_1m.py 1000000 400 > gpurun_out/r2_b4_perf4.log 2>&1; tail -14 gpurun_out/r2_b4_perf4.log
LE_BUILD_VARIANT=3 timeout 200 python scripts/perf_1m.py 1000000 400 > gpurun_out/r2_b4_perf3.log 2>&1; tail -14 gpurun_out/r2_b4_perf3.log
timeout 1500 python -m pytest tests/test_gpu_dd.py -m gpu -q -x > gpurun_out/r2_b4_dd.log 2>&1; tail -5 gpurun_out/r2_b4_dd.log
