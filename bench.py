#!/usr/bin/env python
"""Benchmark of the matching hot path: image-pairs/sec @640x480, indoor_ds dual-softmax (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one `matcher(batch)` call on a batch of 8 synthetic 640x480 grayscale pairs per GPU
(BASELINE.json configs[1]; weak scaling: every rank processes its own 8 pairs, then ONE NCCL all-gather of the
match lists).  Random-init weights (torch.manual_seed(0)), uniform-random images; thr = 0.0 so the fine level
actually runs (with random weights conf.max < the cfg default 0.2 and the fine path would be dead code,
SURVEY.md finding 3) -- recorded in `config`.

Prints ONE JSON line (rank 0).  Keys follow the driver's contract:
  value      pairs/s, inputs resident in HBM, CUDA-event timed, max over ranks, L2 flushed between steps
  e2e        pairs/s through the public API with pinned-host inputs (H2D + D2H inside the timed region)
  roofline   the score-matrix kernel (EpiScoreLse pass of gemm_split_kernel): algorithmic 2*N*L*S*C flops per
             launch / its CUDA-event duration, against the measured bf16 peak of MEASURED_PEAKS.json
  cpu_baseline  the oracle port (PyTorch-CPU backbone + numpy restatement of the reference) on the host cores
`--impl reference` times that CPU port as the whole arm (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W_IMG = 480, 640
BATCH_PER_GPU = 8
METRIC = "image-pairs/sec @640x480 indoor_ds dual-softmax"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--thr", type=float, default=0.0)
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--backbone", default="b200", choices=["b200", "torch"],
                    help="b200: implicit-GEMM convolutions on tcgen05 (default); torch: PyTorch/cuDNN fp32 backbone")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_pairs_per_sec(thr, steps, warmup, pairs_per_step=1):
    """PyTorch-CPU backbone + numpy oracle hot path on all host cores; each step = `pairs_per_step` pairs."""
    import numpy as np
    import torch
    import loftr_b200
    from oracle import loftr_oracle as O
    # all host cores up to 32 threads: beyond that the ~1-10 ms numpy / BLAS calls of this workload only
    # oversubscribe (measured on the 128-core GPU host: 12.1 s/pair with 128 threads)
    cores = min(os.cpu_count() or 1, 32)
    torch.set_num_threads(cores)
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)
    except Exception:
        pass
    torch.manual_seed(0)
    cfg = loftr_b200.get_cfg("indoor_ds", thr=thr)
    model = loftr_b200.LoFTR(cfg).eval()
    state = {k: v.detach().numpy() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(0)
    img0 = torch.rand(pairs_per_step, 1, H, W_IMG, generator=g)
    img1 = torch.rand(pairs_per_step, 1, H, W_IMG, generator=g)
    times, m = [], 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        with torch.no_grad():
            fc, ff = model.backbone(torch.cat([img0, img1], 0))
        (c0, c1), (f0, f1) = fc.split(pairs_per_step), ff.split(pairs_per_step)
        out = O.hot_path(c0.numpy(), c1.numpy(), f0.numpy(), f1.numpy(), state, cfg, (H, W_IMG), (H, W_IMG))
        dt = time.perf_counter() - t0
        m = len(out["b_ids"])
        if it >= warmup:
            times.append(dt)
    return pairs_per_step * len(times) / sum(times), cores, m, sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    v, cores, m, s_per_step = cpu_pairs_per_sec(args.thr, steps, warmup)
    sample = f"{steps} timed steps of 1 pair 640x480 (of the batch of {args.batch}), M={m} matches/pair"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"batch={args.batch} 640x480 pairs, indoor_ds dual-softmax, thr={args.thr}",
                   "global_batch": args.batch * args.gpus, "thr": args.thr, "weights": "random-init seed 0"},
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None, "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import loftr_b200
    from loftr_b200 import _lib, parallel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.backends.cudnn.allow_tf32 = False          # the backbone stays fp32 for parity (SURVEY.md §7 hard part 9)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True

    B = args.batch
    torch.manual_seed(0)
    cfg = loftr_b200.get_cfg("indoor_ds", thr=args.thr)
    model = loftr_b200.LoFTR(cfg, backbone_impl=args.backbone).eval().to(dev)
    g = torch.Generator().manual_seed(1000 + rank)
    h_img0 = torch.rand(B, 1, H, W_IMG, generator=g).pin_memory()
    h_img1 = torch.rand(B, 1, H, W_IMG, generator=g).pin_memory()
    d_img0, d_img1 = h_img0.to(dev), h_img1.to(dev)
    hc, wc = H // 8, W_IMG // 8
    cap = B * hc * wc
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    lib = _lib.load()
    stream = torch.cuda.current_stream()

    t_start = time.perf_counter()

    def note(msg):
        print(f"[bench rank {rank} +{time.perf_counter() - t_start:6.1f}s] {msg}", file=sys.stderr, flush=True)

    def local_step(i0, i1):
        data = {"image0": i0, "image1": i1}
        model(data)
        return data

    def timed(fn, n):
        """sum of per-step CUDA-event times (ms), L2 flushed (untimed) before every step"""
        evs = []
        for _ in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            evs.append((e0, e1))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs)

    # algorithmic FLOPs of the backbone for one step, counted on a meta-device copy (no kernels are launched)
    import copy
    from torch.utils.flop_counter import FlopCounterMode
    with FlopCounterMode(display=False) as fcm:
        copy.deepcopy(model.backbone).to("meta")(torch.empty(2 * B, 1, H, W_IMG, device="meta"))
    backbone_flops = float(fcm.get_total_flops())

    # Phase A -- everything that loads CUDA kernels runs BEFORE the NCCL communicator exists.  With the
    # communicator created first, the first launch of every not-yet-loaded kernel module stalls for tens of
    # seconds on this image (measured with tools/mgpu_diag.py: first cuDNN convolution 54 s after an eager
    # `init_process_group`, 0.5 s before it) -- slow module loading, not a deadlock.
    K, Wm = max(1, args.steps), max(3, args.warmup)
    lo_pair, _ = parallel.shard_range(B * world, rank, world)
    last = None
    for _ in range(Wm):
        last = local_step(d_img0, d_img1)
        parallel.unpack_matches(parallel.pack_matches(last, lo_pair, cap).unsqueeze(0))
    m_per_step = int(last["mconf"].shape[0])
    out_host = {}

    def e2e_local():
        i0 = h_img0.to(dev, non_blocking=True)
        i1 = h_img1.to(dev, non_blocking=True)
        d = local_step(i0, i1)
        if world > 1 and dist.is_initialized():
            d = parallel.all_gather_matches(d, lo_pair, cap)
        for k in ("mkpts0_f", "mkpts1_f", "mconf", "m_bids"):
            out_host[k] = d[k].cpu()

    timed(e2e_local, 1)
    torch.tensor([1.0], dtype=torch.float64, device=dev).max().item()
    note("single-process warm-up done")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to stdout while the communicator is created; stdout must carry exactly one
        # JSON line, so fd 1 points at stderr during initialisation and the collective warm-up.
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
        note("process group up")

    def step(i0, i1):
        data = local_step(i0, i1)
        if world > 1:
            parallel.all_gather_matches(data, lo_pair, cap)
        return data

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(2 if world > 1 else 0):   # collective warm-up (NCCL channels, all-gather kernel)
        step(d_img0, d_img1)
    max_over_ranks(0.0)
    barrier()
    if world > 1:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    note("collective warm-up done")

    # ---- device-resident throughput
    barrier()
    launches0 = lib.lb_launch_count()
    with ClockSampler(local) as clk:
        ms_total = timed(lambda: step(d_img0, d_img1), K)
        barrier()
    launches = lib.lb_launch_count() - launches0
    ms_step = max_over_ranks(ms_total / K)
    value = B * world / (ms_step * 1e-3)

    note(f"device-resident timing done: {ms_step:.2f} ms/step")
    # ---- end to end through the public API: pinned host images in, host match lists out
    for _ in range(2):
        e2e_local()
    barrier()
    t0 = time.perf_counter()
    ms_e2e_dev = timed(e2e_local, K)
    barrier()
    wall_e2e = (time.perf_counter() - t0) * 1e3
    ms_e2e = max_over_ranks(ms_e2e_dev / K)
    h2d = 2 * B * H * W_IMG * 4
    d2h = sum(v.numel() * v.element_size() for v in out_host.values())
    e2e = {"value": B * world / (ms_e2e * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e, "wall_ms_per_step_incl_l2_flush": wall_e2e / K}
    note("e2e timing done")

    # ---- per-kernel CUDA-event timing of the tensor-core kernels (rank 0), separate pass
    roof, kernels = None, {}
    if rank == 0:
        _lib.timing_enable(True)
        nprof = 3
        for _ in range(nprof):
            flush.zero_()
            local_step(d_img0, d_img1)   # rank-local: the other ranks are already waiting at the final barrier
        torch.cuda.synchronize()
        rec = _lib.timing_collect()
        _lib.timing_enable(False)
        pk = peaks()
        L = S = hc * wc
        C = cfg["coarse"]["d_model"]
        for tag, (ms, cnt) in rec.items():
            if cnt:
                kernels[tag] = {"launches_per_step": cnt / nprof, "avg_ms": ms / cnt}
        if "score_lse" in kernels:
            flops = 2.0 * B * L * S * C                      # algorithmic, counted once
            avg_ms = kernels["score_lse"]["avg_ms"]
            achieved = flops / (avg_ms * 1e-3) * 1e-12
            peak = pk["bf16_tflops_sustained"]               # timed inside a long step -> sustained figure
            traffic = None
            tp = os.path.join(ROOT, "profiles", "score_lse_traffic.json")
            if os.path.exists(tp):
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            roof = {"kernel": "gemm_split_kernel<256, EpiScoreLse<rows,cols>> (score matrix + dual-softmax statistics)",
                    "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": traffic, "peak_source": pk["source"] + " (bf16_tflops_sustained)",
                    "flops_per_launch": flops, "avg_launch_ms": avg_ms,
                    "issued_flops_factor": 3, "note": "three fp16 MMAs per product (hi*hi+hi*lo+lo*hi) for fp32-level accuracy; "
                    "frac counts algorithmic flops once"}

        if "backbone_conv" in kernels:
            conv_flops = backbone_flops
            tot_ms = kernels["backbone_conv"]["avg_ms"] * kernels["backbone_conv"]["launches_per_step"]
            ach = conv_flops / (tot_ms * 1e-3) * 1e-12
            kernels["backbone_conv"].update({"algorithmic_flops_per_step": conv_flops, "total_ms_per_step": tot_ms,
                                             "achieved_tflops": ach, "frac_of_measured_bf16_sustained": ach / pk["bf16_tflops_sustained"]})

    # ---- CPU baseline (rank 0, single GPU run only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, m_cpu, s = cpu_pairs_per_sec(args.thr, steps=3, warmup=1)
        cpu = {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port",
               "sample": f"3 timed forwards of 1 pair 640x480 after 1 warm-up ({s:.2f} s each, M={m_cpu}); "
                         "PyTorch-CPU backbone + numpy oracle of the reference hot path"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (backbone fp32; hot-path products = 3x fp16 tcgen05 MMA with fp32 accumulate)",
            "data": "synthetic",
            "config": {"workload": f"batch={B} 640x480 pairs per GPU, indoor_ds dual-softmax, thr={args.thr}",
                       "global_batch": B * world, "thr": args.thr, "weights": "random-init seed 0",
                       "matches_per_step_rank0": m_per_step, "l2": "256 MiB flush buffer written before every timed step",
                       "backbone": ("ResNetFPN_8_2 as implicit-GEMM convolutions on tcgen05 (3x fp16 split, fp32 accumulate)"
                                    if args.backbone == "b200" else "PyTorch/cuDNN fp32 (TF32 off)"),
                       "parallelism": f"pairs sharded over {world} GPU(s), one NCCL all-gather of match lists"},
            "clocks": clk.summary(),
            "e2e": e2e, "gpu_launches": int(launches),
            "gpu_launches_per_step": launches / K,
            "roofline": roof, "kernels": kernels, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
