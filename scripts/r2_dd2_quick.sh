#!/bin/bash
# 2-GPU slab timing, MD only: graphs + per-kernel marks
mkdir -p gpurun_out
N=2
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29705 scripts/dd_perf.py 1000000 300 6.0 > gpurun_out/r2d_dd_md_$N.log 2>&1; grep "world" gpurun_out/r2d_dd_md_$N.log
LE_B200_TIMING=1 LE_B200_DIRECT=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29703 scripts/dd_perf.py 1000000 100 6.0 > gpurun_out/r2d_dd_direct_$N.log 2>&1; grep "k_step4<0" gpurun_out/r2d_dd_direct_$N.log | tail -2
