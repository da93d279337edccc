#!/bin/bash
# Round-2 job r: K^T V of the coarse attention on the tensor cores (EpiKvProj + kv_gemm_kernel + kv_assemble_kernel), A/B.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=gpurun_out/r2r_summary.txt
: > $R
tests/run_gpu_tests.sh "transformer_matches|reference_golden|end_to_end_640|batch8_640x480|832_masked|sweep" > gpurun_out/r2r_tests.txt 2>&1
echo "tests rc=$?" >> $R
cp gpurun_out/parity_stats.jsonl gpurun_out/r2r_parity_stats.jsonl 2>/dev/null
cp gpurun_out/gpu_tests.log gpurun_out/r2r_tests.log
for v in 1 0; do
  LOFTR_B200_KV_GEMM=$v timeout 600 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2r_bench_kvgemm$v.json 2> gpurun_out/r2r_bench_kvgemm$v.err; echo "bench kv_gemm=$v rc=$?" >> $R
done
NCU="ncu --clock-control none --profile-from-start off --kernel-name-base demangled"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2r_launches.csv python tools/profile_step.py > gpurun_out/r2r_launches.out 2>&1
cat $R
grep -E "passed|failed|Error|error|^E " gpurun_out/r2r_tests.txt | tail -40
python - <<'PY'
import json
for v in (1, 0):
    try:
        d = json.load(open(f"gpurun_out/r2r_bench_kvgemm{v}.json"))
        print(v, round(d["value"], 1), round(d["ms_per_step"], 2), round(d["e2e"]["value"], 1), d["clocks"]["sm_mhz"], d["gpu_launches_per_step"])
    except Exception as e:
        print(v, "unreadable", e)
PY
grep -i "kv_gemm\|EpiKvProj\|kv_assemble" gpurun_out/r2r_launches.csv | cut -c60-150,330-420 | head -6
