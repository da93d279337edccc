#!/bin/bash
# ncu --set full captures of the coarse-transformer kernels of the final library (first self pass of layer 1):
# launch ids 25..31 of a step = EpiKvProj, kv_gemm, kv_tile_merge, EpiAttn, LN(merge), mlp1, LN(mlp2).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off --kernel-name-base demangled --set full"
timeout 900 $NCU --import-source on --launch-skip 25 --launch-count 2 -f -o gpurun_out/r3a_kvproj_kvgemm python tools/profile_step.py > /dev/null 2>&1
echo "rc=$?"
timeout 900 $NCU --launch-skip 28 --launch-count 4 -f -o gpurun_out/r3a_attn_ln_mlp python tools/profile_step.py > /dev/null 2>&1
echo "rc=$?"
ls -la gpurun_out/*.ncu-rep; du -sm gpurun_out
