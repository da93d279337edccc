#!/usr/bin/env python
"""Run exactly one profiled matcher(batch) step (after warm-up) between cudaProfilerStart/Stop, for
`ncu --profile-from-start off ...`.  Same workload as bench.py (batch of 640x480 pairs, indoor_ds, thr 0)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import loftr_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--h", type=int, default=480)
ap.add_argument("--w", type=int, default=640)
ap.add_argument("--cfg", default="indoor_ds")
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
model = loftr_b200.LoFTR(loftr_b200.get_cfg(args.cfg, thr=0.0)).eval().cuda()
g = torch.Generator().manual_seed(1000)
i0 = torch.rand(args.batch, 1, args.h, args.w, generator=g).cuda()
i1 = torch.rand(args.batch, 1, args.h, args.w, generator=g).cuda()
for _ in range(args.warmup):
    model({"image0": i0, "image1": i1})
torch.cuda.synchronize()
torch.cuda.profiler.start()
d = {"image0": i0, "image1": i1}
model(d)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("matches:", int(d["mconf"].shape[0]))
