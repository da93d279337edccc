"""Pair-sharded multi-GPU matching: one process per GPU, pairs split by rank, and ONE all-gather of the
match lists at the end (BASELINE.json north_star; replaces the reference's two-step pickled-object gather on a
gloo side group, src/utils/comm.py:113-176, with a single static-shape collective on the compute stream).

Wire format per rank: a fixed-capacity float32 buffer [1 + capacity, 6]
    row 0      : (count, 0, 0, 0, 0, 0)
    row 1 + k  : (x0, y0, x1, y1, mconf, global_pair_id)         k < count
so the collective has static shapes (CUDA-graph friendly) and needs no size pre-exchange.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_pairs: int, rank: int, world: int):
    """Contiguous block partition of `n_pairs` pairs: rank r gets [lo, hi)."""
    base, rem = divmod(n_pairs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_matches(data: dict, pair_offset: int, capacity: int) -> torch.Tensor:
    """Matcher outputs of this rank -> [1 + capacity, 6] float32 on the matches' device."""
    mk0, mk1, conf, bid = data["mkpts0_f"], data["mkpts1_f"], data["mconf"], data["m_bids"]
    m = int(mk0.shape[0])
    if m > capacity:
        raise RuntimeError(f"{m} matches exceed the gather capacity {capacity}")
    buf = torch.zeros(1 + capacity, 6, dtype=torch.float32, device=mk0.device)
    buf[0, 0] = float(m)
    if m:
        buf[1:1 + m, 0:2] = mk0
        buf[1:1 + m, 2:4] = mk1
        buf[1:1 + m, 4] = conf
        buf[1:1 + m, 5] = (bid + pair_offset).to(torch.float32)
    return buf


def unpack_matches(gathered: torch.Tensor) -> dict:
    """[world, 1 + capacity, 6] -> concatenated lists in global (pair, i) order."""
    world = gathered.shape[0]
    counts = gathered[:, 0, 0].round().to(torch.int64).tolist()
    rows = [gathered[r, 1:1 + counts[r]] for r in range(world)]
    allm = torch.cat(rows, 0) if rows else gathered.new_zeros(0, 6)
    return {"mkpts0_f": allm[:, 0:2], "mkpts1_f": allm[:, 2:4], "mconf": allm[:, 4],
            "m_bids": allm[:, 5].round().to(torch.int64), "counts": counts}


def all_gather_matches(data: dict, pair_offset: int, capacity: int, group=None) -> dict:
    """Every rank ends with the global match list (ranks hold contiguous pair blocks, so concatenating in
    rank order preserves the reference's ascending (b, i) ordering).  `capacity` MUST be the same on every rank
    (static-shape collective): use max_pairs_per_rank * min(L, S)."""
    buf = pack_matches(data, pair_offset, capacity)
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return unpack_matches(buf.unsqueeze(0))
    world = dist.get_world_size(group)
    out = torch.empty(world, *buf.shape, dtype=buf.dtype, device=buf.device)
    if buf.is_cuda:
        dist.all_gather_into_tensor(out, buf, group=group)       # NCCL over NVLink / NVSwitch
    else:                                                        # gloo (CPU tests of the host logic)
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf, group=group)
        out = torch.stack(parts, 0)
    return unpack_matches(out)
