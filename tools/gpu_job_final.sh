#!/bin/bash
# Final single-GPU validation of a build: the driver's own test command, smoke(), the per-function parity suite, the
# evaluation tests, the full bench line (extra workloads + CPU arm) and compact ncu evidence for every kernel kind
# (per-kernel metric captures summarised on the box; one `--set full` capture of the score kernel).
# Usage: tools/gpu_job_final.sh <tag>
tag="${1:-r2z}"
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=gpurun_out/${tag}_summary.txt
: > $R
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/${tag}_driver_pytest.txt 2>&1; echo "driver pytest -m gpu rc=$?" >> $R
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.txt 2>&1; echo "smoke rc=$?" >> $R
tests/run_gpu_tests.sh > gpurun_out/${tag}_tests.txt 2>&1; echo "per-function suite rc=$?" >> $R
cp gpurun_out/parity_stats.jsonl gpurun_out/${tag}_parity_stats.jsonl 2>/dev/null
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" >> $R
NCU="ncu --clock-control none --profile-from-start off --kernel-name-base demangled"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py > gpurun_out/${tag}_launches.out 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_active.avg,sm__cycles_elapsed.max,l1tex__throughput.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,launch__registers_per_thread"
# every launch of one step with DRAM bytes and pipe utilisation (129 kernels, metrics only: small report)
timeout 1500 $NCU --metrics $M --section WarpStateStats -f -o gpurun_out/${tag}_step_metrics python tools/profile_step.py > gpurun_out/${tag}_step_metrics.out 2>&1
echo "step metrics rc=$?" >> $R
python tools/ncu_summary.py gpurun_out/${tag}_step_metrics.ncu-rep gpurun_out/${tag}_step_kernels_ncu_summary.csv >> $R 2>&1
timeout 900 $NCU --set full -k regex:"EpiScoreLse" -c 1 -f -o gpurun_out/${tag}_score_lse_full python tools/profile_step.py > gpurun_out/${tag}_score.out 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_score_lse_full.ncu-rep gpurun_out/${tag}_score_lse_ncu_full_summary.csv >> $R 2>&1
sz=$(du -sm gpurun_out | cut -f1)
if [ "$sz" -gt 55 ]; then echo "gpurun_out $sz MB: dropping the per-step metric report (its CSV summary stays)" >> $R; rm -f gpurun_out/${tag}_step_metrics.ncu-rep; fi
cat $R
tail -3 gpurun_out/${tag}_driver_pytest.txt
tail -3 gpurun_out/${tag}_smoke.txt
grep -E "passed|failed" gpurun_out/${tag}_tests.txt | tail -30
python - "$tag" <<'PY'
import json, sys
f = f"gpurun_out/{sys.argv[1]}_bench.json"
try:
    d = json.load(open(f))
    print(f, round(d["value"], 1), round(d["ms_per_step"], 2), round(d["e2e"]["value"], 1), d["clocks"], d["roofline"], d["cpu_baseline"])
    for e in d.get("extra_workloads", []):
        print("   EX", e["workload"][:70], round(e["ms_per_step"], 2), round(e["pairs_per_s"], 1), e["matches_per_step_rank0"], e.get("score_lse", {}).get("frac"))
except Exception as e:
    print(f, "unreadable", e)
PY
