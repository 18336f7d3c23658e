#!/bin/bash
# full GPU suite under the new default step kernel, then ncu --set full of it, then a launch list
mkdir -p gpurun_out
timeout 70 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r1i.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r1i.log
export LE_B200_DIRECT=1
timeout 55 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_step2 -s 2 -c 2 -f -o gpurun_out/prof_r1i python scripts/prof_target2.py > gpurun_out/ncu_full_r1i.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full_r1i.log
timeout 40 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1i.csv python scripts/prof_target2.py 1000000 100 > gpurun_out/ncu_ll_r1i.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/launches_r1i.csv
