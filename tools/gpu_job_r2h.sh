#!/bin/bash
# Round-2 job h: packed fp16 split conversions, EpiKv per-quarter partials, BN staged once per CTA.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=gpurun_out/r2h_summary.txt
: > $R
tests/run_gpu_tests.sh "tensor_core_backbone|transformer_matches|fine_level|reference_golden|batch8_640x480_ds|gemm_split|832_masked|sinkhorn_640" > gpurun_out/r2h_tests.txt 2>&1
echo "tests rc=$?" >> $R
cp gpurun_out/gpu_tests.log gpurun_out/r2h_tests.log
for i in 1 2; do
  timeout 600 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2h_bench_$i.json 2> gpurun_out/r2h_bench_$i.err; echo "bench $i rc=$?" >> $R
done
NCU="ncu --clock-control none --profile-from-start off --kernel-name-base demangled"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2h_launches.csv python tools/profile_step.py > gpurun_out/r2h_launches.out 2>&1
cat $R
grep -E "passed|failed" gpurun_out/r2h_tests.txt | tail -30
python - <<'PY'
import json
for f in ("gpurun_out/r2h_bench_1.json", "gpurun_out/r2h_bench_2.json"):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), round(d["ms_per_step"], 2), round(d["e2e"]["value"], 1), d["clocks"], {k: round(v["total_ms_per_step"], 3) for k, v in d["kernels"].items() if isinstance(v, dict)})
    except Exception as e:
        print(f, "unreadable", e)
PY
