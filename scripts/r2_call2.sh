#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rxXs > gpurun_out/r2_gputests2.log 2>&1; tail -40 gpurun_out/r2_gputests2.log
timeout 200 python scripts/perf_1m.py 1000000 400 > gpurun_out/r2_perf2.log 2>&1; tail -40 gpurun_out/r2_perf2.log
