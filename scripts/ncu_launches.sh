#!/bin/bash
# launch list (per-kernel device time) of a short 1M-bead bench; plain run first, as the recipe requires
CMD="env LE_B200_DIRECT=1 python bench.py --steps 1 --warmup 3 --md-steps 100 --relax 300 --no-cpu"
$CMD > gpurun_out/plain_short.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s ${1:-4500} -c ${2:-1500} --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_short.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/plain_short.log; wc -l gpurun_out/launches.csv
