#!/bin/bash
# round-2 evidence in one call: smoke, the default bench line, an ncu launch list of a short bench, ncu --set full of the step kernel and the rebuild chain
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
python bench.py --steps 4 --warmup 3 > gpurun_out/r2_bench_1m.json 2> gpurun_out/r2_bench_1m.err; echo "bench rc=$?"; cat gpurun_out/r2_bench_1m.json; tail -3 gpurun_out/r2_bench_1m.err
bash scripts/ncu_launches.sh 1900 1200
bash scripts/r2_prof.sh "k_step4|k_build3|k_scan_cells|k_permute|k_cell_count|k_cell_scatter|k_decide" prof_r2g 16
