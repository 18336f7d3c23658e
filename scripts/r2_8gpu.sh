#!/bin/bash
# one 8-GPU call: multi-GPU parity log, weak + strong bench lines, BASELINE configs[4]
mkdir -p gpurun_out
N=8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 scripts/dd_check.py 480000 200 > gpurun_out/r2_dd_check_8.log 2>&1; echo "dd_check rc=$?"; grep -E " ok | FAIL|PASSED|FAILED|ms/step" gpurun_out/r2_dd_check_8.log | tail -14
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus $N --steps 4 --warmup 3 --no-cpu > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench weak rc=$?"; cut -c1-400 gpurun_out/r2_bench_n8.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29703 bench.py --gpus $N --steps 4 --warmup 3 --no-cpu --scaling strong > gpurun_out/r2_bench_strong_n8.json 2> gpurun_out/r2_bench_strong_n8.err; echo "bench strong rc=$?"; cut -c1-400 gpurun_out/r2_bench_strong_n8.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29704 scripts/config5.py 10000000 20 300 > gpurun_out/r2_config5_8.log 2>&1; echo "config5 rc=$?"; grep -v "^\*\|^$\|OMP_NUM\|NCCL\|W1018\|W10" gpurun_out/r2_config5_8.log | tail -8
