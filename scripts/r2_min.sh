#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_minimize.py tests/test_gpu_deck.py -m gpu -q -x > gpurun_out/r2_min_test.log 2>&1; tail -40 gpurun_out/r2_min_test.log | cut -c1-400
