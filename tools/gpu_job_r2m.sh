#!/bin/bash
# Round-2 job m: fine-level attention in one kernel per pass (window_attn_kernel) + vectorised window gather, A/B.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=gpurun_out/r2m_summary.txt
: > $R
tests/run_gpu_tests.sh "fine_level|reference_golden|end_to_end_640|batch8_640x480_ds|832_masked|sweep|duplicate" > gpurun_out/r2m_tests.txt 2>&1
echo "tests rc=$?" >> $R
cp gpurun_out/gpu_tests.log gpurun_out/r2m_tests.log
cp gpurun_out/parity_stats.jsonl gpurun_out/r2m_parity_stats.jsonl 2>/dev/null
for v in 1 0; do
  LOFTR_B200_WINDOW_ATTN=$v timeout 600 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2m_bench_winattn$v.json 2> gpurun_out/r2m_bench_winattn$v.err; echo "bench window_attn=$v rc=$?" >> $R
done
NCU="ncu --clock-control none --profile-from-start off --kernel-name-base demangled"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2m_launches.csv python tools/profile_step.py > gpurun_out/r2m_launches.out 2>&1
cat $R
grep -E "passed|failed|Error|error" gpurun_out/r2m_tests.txt | tail -30
python - <<'PY'
import json
for f in ("gpurun_out/r2m_bench_winattn1.json", "gpurun_out/r2m_bench_winattn0.json"):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), round(d["ms_per_step"], 2), round(d["e2e"]["value"], 1), d["clocks"]["sm_mhz"], d["gpu_launches_per_step"])
    except Exception as e:
        print(f, "unreadable", e)
PY
grep -i "window_attn\|fine_gather\|attn_apply" gpurun_out/r2m_launches.csv | cut -c1-200 | head -3
# source-level captures (kept under 64 MiB in total): 1x1 stride-2 downsample (pure epilogue), l1 lateral with the staged
# upsample window, and the fine-level CUDA-core kernels
NCUF="ncu --clock-control none --profile-from-start off --kernel-name-base demangled --set full --import-source on"
for id in 7 20; do
  timeout 600 $NCUF --launch-skip $id --launch-count 1 -f -o gpurun_out/r2m_conv_id$id python tools/profile_step.py > /dev/null 2>&1
  echo "ncu id $id rc=$?"
done
timeout 900 $NCUF -k regex:"fine_gather|fine_bias|window_attn|stem" -c 4 -f -o gpurun_out/r2m_simt python tools/profile_step.py > /dev/null 2>&1
echo "ncu simt rc=$?"
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
