#!/bin/bash
# Round-2 job n: EpiConv without local-memory round trips (predicated tails, three upsample modes), window_attn with cp.async double buffering.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=gpurun_out/r2n_summary.txt
: > $R
tests/run_gpu_tests.sh "tensor_core_backbone|fine_level|reference_golden|end_to_end_640|batch8_640x480|832_masked|sweep|duplicate|large_logit|copy_pickle" > gpurun_out/r2n_tests.txt 2>&1
echo "tests rc=$?" >> $R
cp gpurun_out/gpu_tests.log gpurun_out/r2n_tests.log
cp gpurun_out/parity_stats.jsonl gpurun_out/r2n_parity_stats.jsonl 2>/dev/null
for v in 1 0; do
  LOFTR_B200_WINDOW_ATTN=$v timeout 600 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2n_bench_winattn$v.json 2> gpurun_out/r2n_bench_winattn$v.err; echo "bench window_attn=$v rc=$?" >> $R
done
NCU="ncu --clock-control none --profile-from-start off --kernel-name-base demangled"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2n_launches.csv python tools/profile_step.py > gpurun_out/r2n_launches.out 2>&1
cat $R
grep -E "passed|failed|Error|error" gpurun_out/r2n_tests.txt | tail -30
python - <<'PY'
import json
for f in ("gpurun_out/r2n_bench_winattn1.json", "gpurun_out/r2n_bench_winattn0.json"):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), round(d["ms_per_step"], 2), round(d["e2e"]["value"], 1), d["clocks"]["sm_mhz"], d["gpu_launches_per_step"])
    except Exception as e:
        print(f, "unreadable", e)
PY
grep -i "window_attn" gpurun_out/r2n_launches.csv | cut -c1-200 | head -3
