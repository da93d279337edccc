#!/bin/bash
# Round-2 job e: TMA-store epilogues (A/B through LOFTR_B200_TMA_STORE), paired TMEM loads; tests, bench, probe.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=gpurun_out/r2e_summary.txt
: > $R
LOFTR_B200_TMA_STORE=0 python tools/gemm_probe.py > gpurun_out/r2e_probe_plain.txt 2>&1
LOFTR_B200_TMA_STORE=1 python tools/gemm_probe.py > gpurun_out/r2e_probe_tma.txt 2>&1
tests/run_gpu_tests.sh "tensor_core_backbone|transformer_matches|fine_level|reference_golden|batch8_640x480_ds|gemm_split|832_masked" > gpurun_out/r2e_tests.txt 2>&1
echo "tests (TMA store on) rc=$?" >> $R
cp gpurun_out/gpu_tests.log gpurun_out/r2e_tests.log
for v in 1 0; do
  LOFTR_B200_TMA_STORE=$v timeout 600 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2e_bench_tma$v.json 2> gpurun_out/r2e_bench_tma$v.err; echo "bench tma=$v rc=$?" >> $R
done
LOFTR_B200_MODE=2 timeout 600 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2e_bench_mode2.json 2> gpurun_out/r2e_bench_mode2.err; echo "bench mode2 rc=$?" >> $R
NCU="ncu --clock-control none --profile-from-start off --kernel-name-base demangled"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2e_launches.csv python tools/profile_step.py > gpurun_out/r2e_launches.out 2>&1
cat $R
grep -E "passed|failed" gpurun_out/r2e_tests.txt | tail -30
cat gpurun_out/r2e_probe_plain.txt gpurun_out/r2e_probe_tma.txt
python - <<'PY'
import json
for f in ("gpurun_out/r2e_bench_tma1.json", "gpurun_out/r2e_bench_tma0.json", "gpurun_out/r2e_bench_mode2.json"):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), round(d["ms_per_step"], 2), {k: round(v["total_ms_per_step"], 3) for k, v in d["kernels"].items() if isinstance(v, dict)})
    except Exception as e:
        print(f, "unreadable", e)
PY
