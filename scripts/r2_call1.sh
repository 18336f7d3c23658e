#!/bin/bash
# round 2, GPU call 1: measure what round 1 left unmeasured (step kernel variants, fp32 pair terms), full gpu test run with reasons
mkdir -p gpurun_out
timeout 60 python scripts/step_ab.py 1000000 0 33 545 1057 > gpurun_out/r2_ab.log 2>&1; tail -8 gpurun_out/r2_ab.log
LE_PAIR_FP32=1 timeout 60 python scripts/step_ab.py 1000000 0 33 545 1057 > gpurun_out/r2_ab_p32.log 2>&1; tail -8 gpurun_out/r2_ab_p32.log
timeout 900 python -m pytest tests -m gpu -q -rxXs > gpurun_out/r2_gputests.log 2>&1; tail -25 gpurun_out/r2_gputests.log
