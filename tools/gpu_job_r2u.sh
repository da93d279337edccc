#!/bin/bash
# Round-2 job u: per-kernel-kind CTA-pair mode A/B (ncu launch lists).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off --kernel-name-base demangled"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2u_launches_default.csv python tools/profile_step.py > gpurun_out/r2u_a.out 2>&1
LOFTR_B200_MODE_TAGS="merge_ln=2,mlp2_ln_res=2,tf_kv_proj_fused=2,tf_q_attn_fused=2,score_lse=2,score_argmax=2,proj_act=2" timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2u_launches_pair.csv python tools/profile_step.py > gpurun_out/r2u_b.out 2>&1
tail -2 gpurun_out/r2u_a.out gpurun_out/r2u_b.out
