#!/bin/bash
# Round-2 first GPU job: gate the new conv layout / fused attention with fallbacks, full parity suite, bench, ncu.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=gpurun_out/r2a_summary.txt
: > $R
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv >> $R
gate() {  # $1 = label, $2 = test filter ; env from caller
  tests/run_gpu_tests.sh "$2" > gpurun_out/r2a_gate_$1.txt 2>&1
  local rc=$?
  cp gpurun_out/gpu_tests.log gpurun_out/r2a_gate_$1.log
  echo "gate $1 ($2) env[REM=$LOFTR_B200_CONV_REM N208=$LOFTR_B200_CONV_N208 FUSED=$LOFTR_B200_FUSED_ATTN] rc=$rc" >> $R
  return $rc
}
if ! gate conv_new "tensor_core_backbone"; then
  if LOFTR_B200_CONV_REM=0 gate conv_n208only "tensor_core_backbone"; then export LOFTR_B200_CONV_REM=0;
  elif LOFTR_B200_CONV_N208=0 gate conv_remonly "tensor_core_backbone"; then export LOFTR_B200_CONV_N208=0;
  else export LOFTR_B200_CONV_REM=0 LOFTR_B200_CONV_N208=0; gate conv_old "tensor_core_backbone"; fi
fi
if ! gate attn_new "transformer_matches"; then
  export LOFTR_B200_FUSED_ATTN=0
  gate attn_old "transformer_matches"
fi
echo "config for the rest: REM=$LOFTR_B200_CONV_REM N208=$LOFTR_B200_CONV_N208 FUSED=$LOFTR_B200_FUSED_ATTN" >> $R
tests/run_gpu_tests.sh > gpurun_out/r2a_tests.txt 2>&1; echo "full suite rc=$?" >> $R
cp gpurun_out/gpu_tests.log gpurun_out/r2a_tests.log
cp gpurun_out/parity_stats.jsonl gpurun_out/r2a_parity_stats.jsonl 2>/dev/null
timeout 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?" >> $R
# A/B of the attention path on the same box
LOFTR_B200_FUSED_ATTN=0 timeout 600 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2a_bench_unfused.json 2> gpurun_out/r2a_bench_unfused.err
echo "bench unfused rc=$?" >> $R
tools/ncu_capture.sh r2a > gpurun_out/r2a_ncu.txt 2>&1; echo "ncu rc=$?" >> $R
cat $R
grep -E "passed|failed" gpurun_out/r2a_tests.txt | tail -30
head -c 1500 gpurun_out/r2a_bench.json
