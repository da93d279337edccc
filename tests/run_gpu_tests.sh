#!/bin/bash
# Run the GPU parity tests group by group in separate processes (a device-side trap poisons the CUDA
# context of its process), keeping per-group logs under gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
groups=(
  "tests/test_engine_gpu.py::test_gemm_split_matches_fp64"
  "tests/test_engine_gpu.py::test_tensor_core_backbone_matches_torch"
  "tests/test_engine_gpu.py::test_transformer_matches_oracle"
  "tests/test_engine_gpu.py::test_coarse_matching_matches_reference_golden"
  "tests/test_engine_gpu.py::test_coarse_matching_full_size_vs_oracle"
  "tests/test_engine_gpu.py::test_coarse_matching_large_logit_spread"
  "tests/test_engine_gpu.py::test_fine_level_matches_oracle"
  "tests/test_engine_gpu.py::test_end_to_end_matches_reference_golden"
  "tests/test_engine_gpu.py::test_end_to_end_640x480_vs_oracle"
  "tests/test_engine_gpu.py::test_outdoor_832_masked_vs_oracle"
  "tests/test_engine_gpu.py::test_sinkhorn_640x480_vs_oracle"
  "tests/test_engine_gpu.py::test_resolution_sweep_properties"
  "tests/test_engine_gpu.py::test_no_cpu_fallback"
)
rc=0
: > gpurun_out/gpu_tests.log
rm -f gpurun_out/parity_stats.jsonl
for g in "${groups[@]}"; do
  echo "=== $g" >> gpurun_out/gpu_tests.log
  timeout 600 python -m pytest "$g" -q -m gpu -x --no-header -p no:cacheprovider >> gpurun_out/gpu_tests.log 2>&1 || rc=1
done
grep -E "^(===|[0-9]+ (passed|failed)|FAILED|ERROR|E  )" gpurun_out/gpu_tests.log | head -80
exit $rc
