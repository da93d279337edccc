#!/bin/bash
# Round-2 job k: FPN lateral convolutions with the TMA-staged upsample window, A/B against per-thread loads.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=gpurun_out/r2k_summary.txt
: > $R
tests/run_gpu_tests.sh "tensor_core_backbone|reference_golden|batch8_640x480_ds|832_masked|sweep" > gpurun_out/r2k_tests.txt 2>&1
echo "tests rc=$?" >> $R
cp gpurun_out/gpu_tests.log gpurun_out/r2k_tests.log
cp gpurun_out/parity_stats.jsonl gpurun_out/r2k_parity_stats.jsonl 2>/dev/null
for v in 1 0; do
  LOFTR_B200_UP_STAGE=$v timeout 600 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2k_bench_upstage$v.json 2> gpurun_out/r2k_bench_upstage$v.err; echo "bench up_stage=$v rc=$?" >> $R
done
NCU="ncu --clock-control none --profile-from-start off --kernel-name-base demangled"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2k_launches.csv python tools/profile_step.py > gpurun_out/r2k_launches.out 2>&1
cat $R
grep -E "passed|failed|Error|error" gpurun_out/r2k_tests.txt | tail -30
python - <<'PY'
import json
for f in ("gpurun_out/r2k_bench_upstage1.json", "gpurun_out/r2k_bench_upstage0.json"):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), round(d["ms_per_step"], 2), round(d["e2e"]["value"], 1), d["clocks"]["sm_mhz"], d["gpu_launches_per_step"])
    except Exception as e:
        print(f, "unreadable", e)
PY
grep -i "EpiConv<2[05][68], true\|EpiConv<2[05][68], 1" gpurun_out/r2k_launches.csv | cut -c1-200 | head -3
