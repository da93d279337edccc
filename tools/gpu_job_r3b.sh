#!/bin/bash
# Branch-free fast elu+1 in the attention epilogues: parity subset, bench line, launch list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tests/run_gpu_tests.sh "transformer_matches|fine_level|reference_golden|end_to_end_640|batch8_640x480|832_masked|sweep|duplicate" > gpurun_out/r3b_tests.txt 2>&1
echo "tests rc=$?"
cp gpurun_out/parity_stats.jsonl gpurun_out/r3b_parity_stats.jsonl 2>/dev/null
grep -E "passed|failed|Error|error|^E " gpurun_out/r3b_tests.txt | tail -20
timeout 600 python bench.py --no-extra --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r3b_bench.json 2> gpurun_out/r3b_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r3b_bench.json')); print(round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['clocks']['sm_mhz'])"
NCU="ncu --clock-control none --profile-from-start off --kernel-name-base demangled"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r3b_launches.csv python tools/profile_step.py > gpurun_out/r3b_launches.out 2>&1
