"""Minimal 2-rank NCCL all-gather probe (diagnostics for the multi-GPU bench)."""
import os
import sys
import time

import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
t0 = time.time()
eager = os.environ.get("PROBE_EAGER", "1") == "1"
if eager:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
else:
    dist.init_process_group("nccl")
print(f"[{rank}] init done {time.time()-t0:.1f}s", flush=True)
x = torch.full((1024, 6), float(rank), device="cuda")
out = torch.empty(world, 1024, 6, device="cuda")
dist.all_gather_into_tensor(out, x)
torch.cuda.synchronize()
print(f"[{rank}] all_gather ok {out[:, 0, 0].tolist()} {time.time()-t0:.1f}s", flush=True)
dist.barrier()
torch.cuda.synchronize()
print(f"[{rank}] barrier ok {time.time()-t0:.1f}s", flush=True)
dist.destroy_process_group()
