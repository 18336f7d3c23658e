#!/bin/bash
# quick check: list/force parity tests + per-kernel timing at 1M beads
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_step3.py -m gpu -q -x > gpurun_out/r2_quick_tests.log 2>&1; tail -5 gpurun_out/r2_quick_tests.log
timeout 200 python scripts/perf_1m.py 1000000 400 > gpurun_out/r2_perf2.log 2>&1; tail -14 gpurun_out/r2_perf2.log
