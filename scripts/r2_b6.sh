#!/bin/bash
# k_build6 against k_build3: identical lists (tests), then the per-kernel table at 10^6 beads; $1 = variants to time, $2 = "ncu" for a full capture of k_build6
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_step3.py -m gpu -q -x -k "build_variants" > gpurun_out/r2_b6_tests.log 2>&1; tail -15 gpurun_out/r2_b6_tests.log
for v in ${1:-3 6}; do
  LE_BUILD_VARIANT=$v timeout 300 python scripts/perf_1m.py 1000000 400 > gpurun_out/r2_b6_perf_v$v.log 2>&1
  grep -E "graphs:|direct:|in-graph|k_build" gpurun_out/r2_b6_perf_v$v.log | tail -8
done
if [ "$2" = ncu ]; then LE_BUILD_VARIANT=6 bash scripts/r2_prof.sh "k_build6" prof_r2_b6 1; fi
