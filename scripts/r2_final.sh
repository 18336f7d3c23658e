#!/bin/bash
# end-of-round evidence in one call: all GPU tests, smoke, the bench line, an ncu launch list of a short bench, ncu --set full of the step kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2f_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2f_smoke.log
python bench.py --steps 4 --warmup 3 > gpurun_out/r2f_bench_1m.json 2> gpurun_out/r2f_bench_1m.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2f_bench_1m.json
bash scripts/ncu_launches.sh 1900 1200
bash scripts/r2_prof.sh "k_step4" prof_r2h 2
