#!/bin/bash
# A/B of kernel variants: LOFTR_B200_LIB in {"", k32} x LOFTR_B200_MODE in {0, 2}
mkdir -p gpurun_out
: > gpurun_out/ab.log
for lib in "" k32; do
  for mode in 0 2; do
    tag="lib=${lib:-k64}_mode=$mode"
    if [ "$lib" = "k32" ]; then
      LOFTR_B200_LIB=$lib LOFTR_B200_MODE=$mode timeout 300 python -m pytest tests/test_engine_gpu.py -q -m gpu --no-header -p no:cacheprovider -x -k "gemm_split or golden or backbone or 640x480_vs_oracle or large_logit" > gpurun_out/ab_tests_$tag.log 2>&1
      echo "$tag tests: $(tail -1 gpurun_out/ab_tests_$tag.log)" >> gpurun_out/ab.log
    fi
    LOFTR_B200_LIB=$lib LOFTR_B200_MODE=$mode timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab_bench_$tag.json 2> gpurun_out/ab_bench_$tag.err
    python - "$tag" <<'PY' >> gpurun_out/ab.log
import json, sys
tag = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/ab_bench_{tag}.json"))
    k = d["kernels"]
    print(tag, "pairs/s %.1f  ms/step %.2f  conv %.2f ms  proj %.4f merge %.4f mlp1 %.4f mlp2 %.4f lse %.4f argmax %.4f fine %.4f" % (
        d["value"], d["ms_per_step"], k["backbone_conv"]["total_ms_per_step"], k["proj_act"]["avg_ms"], k["merge_ln"]["avg_ms"],
        k["mlp1_relu"]["avg_ms"], k["mlp2_ln_res"]["avg_ms"], k["score_lse"]["avg_ms"], k["score_argmax"]["avg_ms"], k["fine_merge"]["avg_ms"]))
except Exception as e:
    print(tag, "FAILED", e)
PY
  done
done
cat gpurun_out/ab.log
