#!/bin/bash
# Round-2 second GPU job: coalesced epilogue I/O (A/B against the -DLB_COALESCE=0 build), ncu captures with source.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=gpurun_out/r2b_summary.txt
: > $R
tests/run_gpu_tests.sh "tensor_core_backbone|transformer_matches|fine_level|reference_golden|batch8_640x480_ds|sinkhorn_640|packed|gemm_split|832_masked" > gpurun_out/r2b_tests.txt 2>&1
echo "tests rc=$?" >> $R
cp gpurun_out/gpu_tests.log gpurun_out/r2b_tests.log
timeout 600 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2b_bench_coal.json 2> gpurun_out/r2b_bench_coal.err; echo "bench coal rc=$?" >> $R
LOFTR_B200_LIB=nocoal timeout 600 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2b_bench_nocoal.json 2> gpurun_out/r2b_bench_nocoal.err; echo "bench nocoal rc=$?" >> $R
NCU_TF_COUNT=16 tools/ncu_capture.sh r2b > gpurun_out/r2b_ncu.txt 2>&1; echo "ncu rc=$?" >> $R
cat $R
grep -E "passed|failed" gpurun_out/r2b_tests.txt | tail -30
python - <<'PY'
import json
for f in ("gpurun_out/r2b_bench_coal.json", "gpurun_out/r2b_bench_nocoal.json"):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), round(d["ms_per_step"], 2), {k: round(v["total_ms_per_step"], 3) for k, v in d["kernels"].items() if isinstance(v, dict)})
    except Exception as e:
        print(f, "unreadable", e)
PY
