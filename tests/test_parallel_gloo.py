"""world_size-2 gloo test of the pair-sharding + match all-gather host logic (no GPU needed)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from loftr_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_outputs(lo, hi, seed):
    g = torch.Generator().manual_seed(seed)
    counts = [int(torch.randint(0, 7, (1,), generator=g)) for _ in range(lo, hi)]
    m = sum(counts)
    bids = torch.cat([torch.full((c,), b, dtype=torch.int64) for b, c in enumerate(counts)]) if m else torch.zeros(0, dtype=torch.int64)
    return {"mkpts0_f": torch.rand(m, 2, generator=g), "mkpts1_f": torch.rand(m, 2, generator=g),
            "mconf": torch.rand(m, generator=g), "m_bids": bids}


def _worker(rank, world, port, n_pairs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = parallel.shard_range(n_pairs, rank, world)
    local = _fake_outputs(lo, hi, 100 + rank)
    out = parallel.all_gather_matches(local, lo, capacity=64)
    q.put((rank, {k: (v.numpy().copy() if torch.is_tensor(v) else v) for k, v in out.items()}))  # plain numpy: no shared-memory handles
    dist.barrier()
    dist.destroy_process_group()


def test_all_gather_matches_world2():
    world, n_pairs = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # expected: concatenation of the ranks' lists with global pair ids
    exp = []
    for r in range(world):
        lo, hi = parallel.shard_range(n_pairs, r, world)
        d = _fake_outputs(lo, hi, 100 + r)
        d["m_bids"] = d["m_bids"] + lo
        exp.append(d)
    for key in ("mkpts0_f", "mkpts1_f", "mconf", "m_bids"):
        want = torch.cat([e[key] for e in exp], 0)
        for r in range(world):
            assert (results[r][key] == want.numpy()).all(), key
    ids = results[0]["m_bids"].tolist()
    assert ids == sorted(ids)


def test_shard_range_partitions():
    for n in (1, 5, 8, 32, 33):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_pack_capacity_error():
    import pytest
    d = _fake_outputs(0, 4, 1)
    with pytest.raises(RuntimeError):
        parallel.pack_matches(d, 0, capacity=0 if d["mconf"].numel() else -1)
