#!/bin/bash
# Round-2 job f: full parity suite, full bench line (extra workloads, CPU baseline), size-bounded ncu captures.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=gpurun_out/r2f_summary.txt
: > $R
tests/run_gpu_tests.sh > gpurun_out/r2f_tests.txt 2>&1; echo "full suite (test_engine_gpu) rc=$?" >> $R
cp gpurun_out/gpu_tests.log gpurun_out/r2f_tests.log
cp gpurun_out/parity_stats.jsonl gpurun_out/r2f_parity_stats.jsonl 2>/dev/null
timeout 600 python -m pytest tests/test_evaluation.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r2f_eval_tests.txt 2>&1; echo "evaluation tests rc=$?" >> $R
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?" >> $R
LOFTR_B200_FUSED_UPSAMPLE=1 timeout 600 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2f_bench_fused_upsample.json 2> gpurun_out/r2f_bench_fused_upsample.err; echo "bench fused-upsample rc=$?" >> $R
NCU="ncu --clock-control none --profile-from-start off --kernel-name-base demangled"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2f_launches.csv python tools/profile_step.py > gpurun_out/r2f_launches.out 2>&1
FULL="--set full --import-source on"
timeout 900 $NCU $FULL -k regex:"EpiKv|EpiAttn|EpiLayerNorm|EpiPlanes" -c 6 -f -o gpurun_out/r2f_tf python tools/profile_step.py > gpurun_out/r2f_tf.out 2>&1
timeout 900 $NCU $FULL -k regex:"EpiScore" -c 2 -f -o gpurun_out/r2f_score python tools/profile_step.py > gpurun_out/r2f_score.out 2>&1
timeout 900 $NCU $FULL -k regex:"EpiConv|upsample" -s 16 -c 5 -f -o gpurun_out/r2f_conv python tools/profile_step.py > gpurun_out/r2f_conv.out 2>&1
du -sm gpurun_out; ls -la gpurun_out | grep ncu-rep
sz=$(du -sm gpurun_out | cut -f1)
if [ "$sz" -gt 60 ]; then echo "too big ($sz MB): dropping the conv capture" >> $R; rm -f gpurun_out/r2f_conv.ncu-rep; fi
cat $R
grep -E "passed|failed" gpurun_out/r2f_tests.txt | tail -30
tail -3 gpurun_out/r2f_eval_tests.txt
python - <<'PY'
import json
for f in ("gpurun_out/r2f_bench.json", "gpurun_out/r2f_bench_fused_upsample.json"):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), round(d["ms_per_step"], 2), round(d["e2e"]["value"], 1), {k: round(v["total_ms_per_step"], 3) for k, v in d["kernels"].items() if isinstance(v, dict)})
        for e in d.get("extra_workloads", []):
            print("   EX", e["workload"][:70], round(e["ms_per_step"], 2), round(e["pairs_per_s"], 1), e["matches_per_step_rank0"], e.get("score_lse", {}).get("frac"))
    except Exception as e:
        print(f, "unreadable", e)
PY
