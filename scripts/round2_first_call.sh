#!/bin/bash
# First GPU call of round 2 (one GPU, about two minutes): everything that was written after the last GPU minute of
# round 1 and is still unmeasured.
#   1. step kernels: default (33) vs fused reneighbor decision (545) vs dynamic tile fetch (1057), bit-identity + timing at 1M beads
#   2. fp32 pair terms (LE_PAIR_FP32=1): k_step vs k_step2p identity + timing (identity is relative to variant 0 under the same switch)
#   3. the pending tests (non-strict xfail): fused decide, fp32 pair terms, USER-LE barrier variants
#   4. slabs on ONE GPU (LE_DD_SHARE_GPU=1, two ranks time-slicing device 0): k_step (0), k_step2<1,256> (17), k_step2p<1,256> (49)
mkdir -p gpurun_out
timeout 40 python scripts/step_ab.py 1000000 0 33 545 1057 > gpurun_out/r2_ab.log 2>&1; tail -5 gpurun_out/r2_ab.log
LE_PAIR_FP32=1 timeout 40 python scripts/step_ab.py 1000000 0 33 545 1057 > gpurun_out/r2_ab_p32.log 2>&1; tail -5 gpurun_out/r2_ab_p32.log
timeout 120 python -m pytest tests/test_gpu_step2.py tests/test_gpu_parity.py -q -rxX 2>&1 | tail -12
for v in 0 17 49; do
  LE_DD_SHARE_GPU=1 LE_STEP_VARIANT=$v timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 2971$((v % 10)) scripts/dd_check.py 60000 120 > gpurun_out/r2_dd_$v.log 2>&1
  echo "dd_check variant $v rc=$?"; grep -E "PASSED|FAILED|ms/step|differ" gpurun_out/r2_dd_$v.log | tail -6
done
