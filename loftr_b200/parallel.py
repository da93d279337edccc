"""Pair-sharded multi-GPU matching: one process per GPU, pairs split by rank, and ONE all-gather of the
match lists at the end (BASELINE.json north_star; replaces the reference's two-step pickled-object gather on a
gloo side group, src/utils/comm.py:113-176, with a single static-shape collective on the compute stream).

Wire format per rank: a fixed-capacity float32 buffer [1 + capacity, 6]
    row 0      : (count, 0, 0, 0, 0, 0)
    row 1 + k  : (x0, y0, x1, y1, mconf, global_pair_id)         k < count
so the collective has static shapes (CUDA-graph friendly) and needs no size pre-exchange.

On CUDA tensors the exchange runs through the library's C ABI (`lb_pack_matches` -> `lb_allgather_matches`
(ncclAllGather over NVLink / NVSwitch) -> `lb_unpack_matches`; include/loftr_b200.h): three launches, one small
device->host read of the counts.  `MatchGatherer` owns the communicator.  CPU tensors (the gloo tests of the host
logic) go through torch.distributed with the same wire format.
"""
from __future__ import annotations

import ctypes as C

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
    if mk0.is_cuda:
        from . import _lib
        lib = _lib.load()
        buf = torch.empty(1 + capacity, 6, dtype=torch.float32, device=mk0.device)
        mk0, mk1, conf, bid = (t.contiguous() for t in (mk0.float(), mk1.float(), conf.float(), bid.long()))
        st = C.c_void_p(torch.cuda.current_stream(buf.device).cuda_stream)
        _lib.check(lib.lb_pack_matches(mk0.data_ptr(), mk1.data_ptr(), conf.data_ptr(), bid.data_ptr(), m, int(pair_offset),
                                       buf.data_ptr(), capacity, st))
        return buf
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
    world, rows, _ = gathered.shape
    capacity = rows - 1
    if gathered.is_cuda:
        from . import _lib
        lib = _lib.load()
        dev = gathered.device
        gathered = gathered.contiguous()
        out_cap = world * capacity
        mk0 = torch.empty(max(out_cap, 1), 2, dtype=torch.float32, device=dev)
        mk1 = torch.empty(max(out_cap, 1), 2, dtype=torch.float32, device=dev)
        conf = torch.empty(max(out_cap, 1), dtype=torch.float32, device=dev)
        bids = torch.empty(max(out_cap, 1), dtype=torch.int64, device=dev)
        counts = torch.empty(world + 1, dtype=torch.int32, device=dev)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.lb_unpack_matches(gathered.data_ptr(), world, capacity, mk0.data_ptr(), mk1.data_ptr(),
                                         conf.data_ptr(), bids.data_ptr(), out_cap, counts.data_ptr(), st))
        c = counts.tolist()           # the one device->host read: sizes the returned lists
        per_rank, total = c[:world], c[world]
        if max(per_rank) > capacity:
            raise OverflowError(max(per_rank))
        return {"mkpts0_f": mk0[:total], "mkpts1_f": mk1[:total], "mconf": conf[:total], "m_bids": bids[:total],
                "counts": per_rank}
    counts = gathered[:, 0, 0].round().to(torch.int64).tolist()
    rows = [gathered[r, 1:1 + counts[r]] for r in range(world)]
    allm = torch.cat(rows, 0) if rows else gathered.new_zeros(0, 6)
    return {"mkpts0_f": allm[:, 0:2], "mkpts1_f": allm[:, 2:4], "mconf": allm[:, 4],
            "m_bids": allm[:, 5].round().to(torch.int64), "counts": counts}


class MatchGatherer:
    """Owns the library-side NCCL communicator of this process (one per GPU) and a running capacity bound.

    The rendezvous id is created by rank 0 (`lb_comm_unique_id`) and shipped through the already initialised
    torch.distributed group; after that torch.distributed plays no part in the exchange.
    `capacity`: rows per rank on the wire.  `None` starts from `initial` and grows (identically on every rank: the
    counts of all ranks are part of the gathered data) whenever a rank's list approaches it; a list that does not fit
    is re-sent once with the capacity every rank derives from the reported counts.
    """

    def __init__(self, device, group=None, capacity: int | None = None, initial: int = 4096):
        from . import _lib
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.fixed = capacity is not None
        self.capacity = int(capacity) if self.fixed else int(initial)
        idbuf = C.create_string_buffer(_lib.NCCL_UNIQUE_ID_BYTES)
        path = self._nccl_path()
        if self.rank == 0:
            _lib.check(self.lib.lb_comm_unique_id(idbuf, path))
        box = [bytes(idbuf.raw)]
        dist.broadcast_object_list(box, src=0, group=group)
        self.comm = C.c_void_p()
        _lib.check(self.lib.lb_comm_init(box[0], self.rank, self.world, self.device.index or 0, path, C.byref(self.comm)))

    @staticmethod
    def _nccl_path():
        try:   # the NCCL that PyTorch itself loaded (pip wheel layout); otherwise the loader's default libnccl.so.2
            import os
            import nvidia.nccl as n
            p = os.path.join(os.path.dirname(n.__file__), "lib", "libnccl.so.2")
            return p.encode() if os.path.exists(p) else None
        except Exception:
            return None

    def all_gather(self, data: dict, pair_offset: int) -> dict:
        m = int(data["mkpts0_f"].shape[0])
        cap = self.capacity
        # a rank whose own list does not fit still takes part with a truncated payload and its TRUE count in row 0
        send = data if m <= cap else {k: data[k][:cap] for k in ("mkpts0_f", "mkpts1_f", "mconf", "m_bids")}
        out = self._exchange_counted(send, pair_offset, cap, m)
        try:
            res = unpack_matches(out)
        except OverflowError as e:
            if self.fixed:
                raise RuntimeError(f"{e.args[0]} matches exceed the fixed gather capacity {cap}") from None
            cap = self._grow(int(e.args[0]))
            res = unpack_matches(self._exchange_counted(data, pair_offset, cap, m))
        if not self.fixed and 2 * max(res["counts"]) > self.capacity:
            self._grow(max(res["counts"]))
        return res

    def _exchange_counted(self, send, pair_offset, cap, true_count):
        from . import _lib
        buf = pack_matches(send, pair_offset, cap)
        if true_count != int(send["mkpts0_f"].shape[0]):
            buf[0, 0] = float(true_count)
        out_buf = torch.empty(self.world, 1 + cap, 6, dtype=torch.float32, device=buf.device)
        st = C.c_void_p(torch.cuda.current_stream(buf.device).cuda_stream)
        _lib.check(self.lib.lb_allgather_matches(self.comm, buf.data_ptr(), out_buf.data_ptr(), cap, st))
        return out_buf

    def _grow(self, need: int) -> int:
        cap = self.capacity
        while cap < 2 * need:
            cap *= 2
        self.capacity = cap
        return cap

    def close(self):
        if self.comm:
            self.lib.lb_comm_destroy(self.comm)
            self.comm = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def all_gather_matches(data: dict, pair_offset: int, capacity: int, group=None, gatherer: MatchGatherer | None = None) -> dict:
    """Every rank ends with the global match list (ranks hold contiguous pair blocks, so concatenating in
    rank order preserves the reference's ascending (b, i) ordering).  `capacity` MUST be the same on every rank
    (static-shape collective).  CUDA tensors need a `MatchGatherer` (library NCCL communicator); CPU tensors use
    torch.distributed (gloo)."""
    buf_is_cuda = data["mkpts0_f"].is_cuda
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return unpack_matches(pack_matches(data, pair_offset, capacity).unsqueeze(0))
    if buf_is_cuda:
        if gatherer is None:
            raise RuntimeError("loftr_b200.parallel: gathering CUDA match lists needs a MatchGatherer (library-side NCCL "
                               "communicator); create one per process after init_process_group")
        return gatherer.all_gather(data, pair_offset)
    buf = pack_matches(data, pair_offset, capacity)
    world = dist.get_world_size(group)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return unpack_matches(torch.stack(parts, 0))
