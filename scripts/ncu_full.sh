#!/bin/bash
# ncu --set full capture of the kernels matching $1 (regex) inside scripts/prof_target.py's profiled region
PAT=${1:-k_step}; OUT=${2:-prof}; CNT=${3:-3}; SKIP=${4:-2}
LE_B200_DIRECT=1 python scripts/prof_target.py > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$PAT -s $SKIP -c $CNT -f -o gpurun_out/$OUT env LE_B200_DIRECT=1 python scripts/prof_target.py > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/prof_plain.log; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/
