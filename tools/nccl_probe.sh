#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  echo "=== $name" >> gpurun_out/nccl_probe.log
  env "$@" NCCL_DEBUG=INFO timeout 75 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) tools/nccl_probe.py > gpurun_out/nccl_$name.log 2>&1
  echo "rc=$?" >> gpurun_out/nccl_probe.log
  grep -E "^\[[01]\]|rc=" gpurun_out/nccl_$name.log >> gpurun_out/nccl_probe.log
  tail -4 gpurun_out/nccl_$name.log | cut -c1-300 >> gpurun_out/nccl_probe.log
}
nvidia-smi topo -m > gpurun_out/topo.log 2>&1
run default PROBE_EAGER=1
run lazy PROBE_EAGER=0
run lo NCCL_SOCKET_IFNAME=lo PROBE_EAGER=1
run nop2p NCCL_P2P_DISABLE=1 NCCL_SOCKET_IFNAME=lo PROBE_EAGER=1
cat gpurun_out/nccl_probe.log
