"""Stage-by-stage 2-rank diagnostic of the sharded matcher + match all-gather (prints where it stalls)."""
import faulthandler
import os
import sys
import time

faulthandler.dump_traceback_later(50, repeat=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
t0 = time.time()


def say(msg):
    print(f"[{rank} {time.time()-t0:6.1f}s] {msg}", flush=True)


torch.cuda.set_device(local)
dev = torch.device("cuda", local)
mode = os.environ.get("DIAG_MODE", "nccl_first")
if mode == "nccl_first":
    dist.init_process_group("nccl", device_id=dev)
    say("pg init (eager)")
import loftr_b200
from loftr_b200 import parallel
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(0)
model = loftr_b200.LoFTR(loftr_b200.get_cfg("indoor_ds", thr=0.0)).eval().to(dev)
say("model built")
B, H, W = 2, 240, 320
g = torch.Generator().manual_seed(rank)
i0 = torch.rand(B, 1, H, W, generator=g).to(dev); i1 = torch.rand(B, 1, H, W, generator=g).to(dev)
data = {"image0": i0, "image1": i1}
with torch.no_grad():
    fc, ff = model.backbone(torch.cat([i0, i1]))
torch.cuda.synchronize(); say("backbone alone ok")
model(data)
torch.cuda.synchronize(); say(f"matcher ok M={int(data['mconf'].shape[0])}")
if mode != "nccl_first":
    dist.init_process_group("nccl", device_id=dev)
    say("pg init (after first forward)")
cap = B * (H // 8) * (W // 8)
lo, _ = parallel.shard_range(B * world, rank, world)
out = parallel.all_gather_matches(data, lo, cap)
torch.cuda.synchronize(); say(f"all_gather ok total={int(out['mconf'].shape[0])} counts={out['counts']}")
model(data); torch.cuda.synchronize(); say("second forward ok")
dist.barrier(); torch.cuda.synchronize(); say("barrier ok")
dist.destroy_process_group()
say("done")
