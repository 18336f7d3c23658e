#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/step_ab.py 1000000 4:3 45:3 46:3 48:3 > gpurun_out/r2_ab5.log 2>&1; grep variant gpurun_out/r2_ab5.log
bash scripts/r2_prof.sh "k_step4" prof_r2e 3
