#!/bin/bash
# Alternating A/B of one environment switch: tools/gpu_job_ab.sh <tag> <ENV_NAME> [pairs]   (A: =1, B: =0)
tag="$1"; name="$2"; n="${3:-4}"
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/${tag}_ab.txt
: > $out
for i in $(seq 1 $n); do
  for v in 1 0; do
    env $name=$v timeout 600 python bench.py --no-extra --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$name=$v', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['clocks']['samples'])" >> $out
  done
done
cat $out
python - "$out" <<'PY'
import sys, statistics
rows = [l.split() for l in open(sys.argv[1]) if l.strip()]
for v in ("=1", "=0"):
    xs = [float(r[2]) for r in rows if r[0].endswith(v)]
    print(v, "median ms/step", statistics.median(xs), "min", min(xs), "n", len(xs))
PY
