#!/bin/bash
mkdir -p gpurun_out
time timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k one_million > gpurun_out/r2_1m_test.log 2>&1; tail -30 gpurun_out/r2_1m_test.log | cut -c1-300
