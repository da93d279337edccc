#!/bin/bash
# Round-2 job y: programmatic dependent launch of the tensor-core kernels (LOFTR_B200_PDL=1): parity subset + alternating A/B.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOFTR_B200_PDL=1 tests/run_gpu_tests.sh "tensor_core_backbone|transformer_matches|reference_golden|batch8_640x480_ds|832_masked|duplicate" > gpurun_out/r2y_tests.txt 2>&1
echo "tests with PDL rc=$?"
grep -E "passed|failed|Error|error|^E " gpurun_out/r2y_tests.txt | tail -20
bash tools/gpu_job_ab.sh r2y LOFTR_B200_PDL 4
