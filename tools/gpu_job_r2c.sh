#!/bin/bash
# Round-2 third GPU job: main-loop / epilogue probe + size-bounded ncu captures with source.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/gemm_probe.py > gpurun_out/r2c_probe_mode0.txt 2>&1
LOFTR_B200_MODE=2 python tools/gemm_probe.py > gpurun_out/r2c_probe_mode2.txt 2>&1
NCU="ncu --clock-control none --profile-from-start off --kernel-name-base demangled --set full --import-source on"
timeout 900 $NCU -k regex:"EpiKv|EpiAttn|EpiLayerNorm|EpiPlanes" -c 6 -f -o gpurun_out/r2c_tf python tools/profile_step.py > gpurun_out/r2c_tf.out 2>&1
LOFTR_B200_LIB=nocoal timeout 900 $NCU -k regex:"EpiConv" -s 4 -c 5 -f -o gpurun_out/r2c_conv_nocoal python tools/profile_step.py > gpurun_out/r2c_conv.out 2>&1
ls -la gpurun_out/
sz=$(du -sm gpurun_out | cut -f1)
if [ "$sz" -gt 60 ]; then echo "too big ($sz MB): dropping the conv capture"; rm -f gpurun_out/r2c_conv_nocoal.ncu-rep; fi
cat gpurun_out/r2c_probe_mode0.txt gpurun_out/r2c_probe_mode2.txt
