#!/bin/bash
# 2-GPU job (r2g, r2j): multi-GPU parity test (library NCCL communicator, overflow protocol) + 2-GPU bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2j_summary.txt
timeout 900 python -m pytest tests/test_multigpu.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r2j_mgpu_tests.txt 2>&1
echo "multigpu tests rc=$?" >> gpurun_out/r2j_summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2j_bench_2gpu.json 2> gpurun_out/r2j_bench_2gpu.err
echo "bench 2gpu rc=$?" >> gpurun_out/r2j_summary.txt
cat gpurun_out/r2j_summary.txt; tail -5 gpurun_out/r2j_mgpu_tests.txt; tail -3 gpurun_out/r2j_bench_2gpu.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2j_bench_2gpu.json"))
    print(round(d["value"], 1), round(d["ms_per_step"], 2), round(d["e2e"]["value"], 1), d["rank_ms"], d["gather_check"], d["gpu_launches_per_step"])
    for e in d.get("extra_workloads", []):
        print("   EX", e["workload"][:90], round(e["ms_per_step"], 2), round(e["pairs_per_s"], 1), round(e["e2e_pairs_per_s"], 1))
except Exception as e:
    print("bench line unreadable", e)
PY
