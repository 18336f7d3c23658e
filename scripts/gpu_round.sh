#!/bin/bash
# one gpurun call: GPU parity tests, smoke, the 1M-bead bench, then an ncu launch list of a short bench
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
python bench.py --steps 4 --warmup 3 > gpurun_out/bench_1m.json 2> gpurun_out/bench_1m.err; echo "bench rc=$?"; cat gpurun_out/bench_1m.json; tail -3 gpurun_out/bench_1m.err
