#!/bin/bash
# weak-scaling timing of the slab engine: N ranks, 1M beads per GPU: graphs (MD only / with USER-LE) and direct launches with per-kernel marks
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29704 scripts/dd_perf.py 1000000 400 6.0 le > gpurun_out/r2b_dd_perf_$N.log 2>&1; grep "world" gpurun_out/r2b_dd_perf_$N.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29705 scripts/dd_perf.py 1000000 400 6.0 > gpurun_out/r2b_dd_perf_md_$N.log 2>&1; grep "world" gpurun_out/r2b_dd_perf_md_$N.log
LE_B200_TIMING=1 LE_B200_DIRECT=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29703 scripts/dd_perf.py 1000000 200 6.0 le > gpurun_out/r2b_dd_perf_direct_$N.log 2>&1; grep -A45 "timing rank 0" gpurun_out/r2b_dd_perf_direct_$N.log | tail -47
