#!/bin/bash
mkdir -p gpurun_out
for mode in nccl_first nccl_late; do
  echo "=== $mode" >> gpurun_out/mgpu_diag.log
  DIAG_MODE=$mode NCCL_DEBUG=WARN timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) tools/mgpu_diag.py > gpurun_out/mgpu_diag_$mode.log 2>&1
  echo "rc=$?" >> gpurun_out/mgpu_diag.log
  grep -E "^\[[01] |File|Thread|Error|error" gpurun_out/mgpu_diag_$mode.log | head -60 >> gpurun_out/mgpu_diag.log
done
cat gpurun_out/mgpu_diag.log
