#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/gemm_probe.py > gpurun_out/r2e_probe.txt 2>&1
cat gpurun_out/r2e_probe.txt
