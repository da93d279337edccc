#!/bin/bash
# Round-2 job d: shared-pointer fix, L2 prefetch of residuals, single-pass LayerNorm moments; coalescing A/B again.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=gpurun_out/r2d_summary.txt
: > $R
tests/run_gpu_tests.sh "tensor_core_backbone|transformer_matches|fine_level|reference_golden|batch8_640x480_ds|sinkhorn_640|gemm_split|832_masked|large_logit" > gpurun_out/r2d_tests.txt 2>&1
echo "tests rc=$?" >> $R
cp gpurun_out/gpu_tests.log gpurun_out/r2d_tests.log
timeout 600 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2d_bench_coal.json 2> gpurun_out/r2d_bench_coal.err; echo "bench coal rc=$?" >> $R
LOFTR_B200_LIB=nocoal timeout 600 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2d_bench_nocoal.json 2> gpurun_out/r2d_bench_nocoal.err; echo "bench nocoal rc=$?" >> $R
python tools/gemm_probe.py > gpurun_out/r2d_probe_mode0.txt 2>&1
NCU="ncu --clock-control none --profile-from-start off --kernel-name-base demangled"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2d_launches.csv python tools/profile_step.py > gpurun_out/r2d_launches.out 2>&1
LOFTR_B200_LIB=nocoal timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2d_launches_nocoal.csv python tools/profile_step.py > gpurun_out/r2d_launches_nocoal.out 2>&1
cat $R
grep -E "passed|failed" gpurun_out/r2d_tests.txt | tail -30
cat gpurun_out/r2d_probe_mode0.txt
python - <<'PY'
import json
for f in ("gpurun_out/r2d_bench_coal.json", "gpurun_out/r2d_bench_nocoal.json"):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), round(d["ms_per_step"], 2), {k: round(v["total_ms_per_step"], 3) for k, v in d["kernels"].items() if isinstance(v, dict)})
    except Exception as e:
        print(f, "unreadable", e)
PY
