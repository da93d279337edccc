#!/bin/bash
# Round-2 job p: 1x1 convolutions with a single accumulator (two TMEM stages); source-level capture of a residual 3x3 layer.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=gpurun_out/r2p_summary.txt
: > $R
tests/run_gpu_tests.sh "tensor_core_backbone|reference_golden|end_to_end_640|batch8_640x480_ds|832_masked|sweep" > gpurun_out/r2p_tests.txt 2>&1
echo "tests rc=$?" >> $R
cp gpurun_out/parity_stats.jsonl gpurun_out/r2p_parity_stats.jsonl 2>/dev/null
timeout 600 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?" >> $R
NCU="ncu --clock-control none --profile-from-start off --kernel-name-base demangled"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2p_launches.csv python tools/profile_step.py > gpurun_out/r2p_launches.out 2>&1
cat $R
grep -E "passed|failed|Error|error" gpurun_out/r2p_tests.txt | tail -30
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2p_bench.json"))
print(round(d["value"], 1), round(d["ms_per_step"], 2), round(d["e2e"]["value"], 1), d["clocks"]["sm_mhz"], d["gpu_launches_per_step"])
PY
LOFTR_B200_CONV1X1_DUAL=1 timeout 600 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2p_bench_dual1x1.json 2> gpurun_out/r2p_bench_dual1x1.err
python -c "
import json; d=json.load(open('gpurun_out/r2p_bench_dual1x1.json')); print('dual1x1', round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
NCUF="ncu --clock-control none --profile-from-start off --kernel-name-base demangled --set full --import-source on"
timeout 600 $NCUF --launch-skip 8 --launch-count 1 -f -o gpurun_out/r2p_conv_id8 python tools/profile_step.py > /dev/null 2>&1
echo "ncu id 8 rc=$?"
ls -la gpurun_out/*.ncu-rep
