#!/bin/bash
# fix bond/create against the compiled reference (+ the neighbours of the code it shares: bond/break, the USER-LE replay)
mkdir -p gpurun_out
timeout 800 python -m pytest tests/test_gpu_md.py tests/test_gpu_parity.py -m gpu -q -x -k "bond_create or bond_break or le_replay" > gpurun_out/r2_bc_tests.log 2>&1; tail -25 gpurun_out/r2_bc_tests.log
