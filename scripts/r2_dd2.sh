#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 scripts/dd_check.py $((60000*N)) 200 > gpurun_out/r2_dd_check_$N.log 2>&1; echo "dd_check rc=$?"; grep -E "ok|FAIL|PASSED|FAILED|ms/step|Error|error" gpurun_out/r2_dd_check_$N.log | tail -16
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29702 scripts/dd_check.py 0 300 melt > gpurun_out/r2_dd_melt_$N.log 2>&1; echo "dd_melt rc=$?"; grep -E "ok|FAIL|PASSED|FAILED|ms/step|Error|error" gpurun_out/r2_dd_melt_$N.log | tail -10
LE_B200_TIMING=1 LE_B200_DIRECT=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29703 scripts/dd_perf.py 1000000 200 6.0 le > gpurun_out/r2_dd_perf_direct_$N.log 2>&1; echo "perf direct rc=$?"; grep -A22 "timing rank 0" gpurun_out/r2_dd_perf_direct_$N.log | tail -24
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29704 scripts/dd_perf.py 1000000 400 6.0 le > gpurun_out/r2_dd_perf_$N.log 2>&1; grep "world" gpurun_out/r2_dd_perf_$N.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29705 scripts/dd_perf.py 1000000 400 6.0 > gpurun_out/r2_dd_perf_md_$N.log 2>&1; grep "world" gpurun_out/r2_dd_perf_md_$N.log
