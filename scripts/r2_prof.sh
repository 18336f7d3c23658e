#!/bin/bash
# ncu --set full of the rebuild chain + step kernels (direct launches) at 1M beads
PAT=${1:-"k_step3|k_build3|k_scan_cells|k_permute|k_cell_count|k_cell_scatter"}; OUT=${2:-prof_r2a}; CNT=${3:-13}
mkdir -p gpurun_out
LE_B200_DIRECT=1 python scripts/prof_target2.py 1000000 12 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$PAT" -c $CNT -f -o gpurun_out/$OUT env LE_B200_DIRECT=1 python scripts/prof_target2.py 1000000 12 > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/prof_plain.log; tail -3 gpurun_out/ncu_full.log
