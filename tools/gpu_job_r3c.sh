#!/bin/bash
# Source-level ncu captures of EpiAttn (id 28), LN(mlp2) (id 31) and the fine q/k/v projection (id 123).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off --kernel-name-base demangled --set full --import-source on"
for id in 28 31 123; do
  timeout 900 $NCU --launch-skip $id --launch-count 1 -f -o gpurun_out/r3c_id$id python tools/profile_step.py > /dev/null 2>&1
  echo "id $id rc=$?"
done
ls -la gpurun_out/*.ncu-rep; du -sm gpurun_out
