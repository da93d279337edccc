"""Pair sharding over 2 GPUs + the NCCL all-gather of match lists reproduces the single-GPU result
(SURVEY.md §8(e)).  Needs >= 2 GPUs: skipped on the single-GPU test box."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_pairs, q):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), here, os.path.join(here, "golden")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import util
    from cases import build_inputs
    from loftr_b200 import parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    case = {"name": "mg", "n": n_pairs, "hw0": (96, 128), "hw1": (96, 128), "thr": 0.0, "images": "smooth"}
    model, _, _ = util.build_model(case, f"cuda:{rank}")
    inp = build_inputs(case)
    lo, hi = parallel.shard_range(n_pairs, rank, world)
    data = {k: torch.from_numpy(v[lo:hi]).to(f"cuda:{rank}") for k, v in inp.items()}
    model(data)   # first forward before the communicator exists (cuDNN module loading is ~55 s slower after it)
    torch.cuda.synchronize()
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    # library-side communicator (C ABI: lb_comm_init / lb_allgather_matches); starts with a deliberately small
    # capacity so that the overflow -> grow -> re-send protocol is exercised as well
    gatherer = parallel.MatchGatherer(torch.device("cuda", rank), initial=16)
    out = parallel.all_gather_matches(data, lo, capacity=0, gatherer=gatherer)
    assert gatherer.capacity >= 2 * max(out["counts"])
    out2 = parallel.all_gather_matches(data, lo, capacity=0, gatherer=gatherer)   # steady state: no re-send
    assert all(torch.equal(out[k], out2[k]) for k in ("mkpts0_f", "mkpts1_f", "mconf", "m_bids"))
    res = {k: (v.cpu().numpy() if torch.is_tensor(v) else v) for k, v in out.items()}
    if rank == 0:  # single-GPU run of the whole batch for comparison
        full = {k: torch.from_numpy(v).to("cuda:0") for k, v in inp.items()}
        model(full)
        res["full"] = {k: full[k].cpu().numpy() for k in ("mkpts0_f", "mkpts1_f", "mconf", "m_bids")}
    q.put((rank, res))
    dist.barrier()
    gatherer.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_shard_and_gather_equals_single_gpu():
    import torch.multiprocessing as mp
    world, n_pairs = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    full = results[0]["full"]
    assert len(full["mconf"]) > 50
    for r in range(world):
        for k in ("mkpts0_f", "mkpts1_f", "mconf", "m_bids"):
            assert results[r][k].shape == full[k].shape, (r, k)
            # batch composition does not enter any per-pair computation of the hot path; cuDNN, however, picks
            # its convolution algorithm per batch size, so backbone features differ in the last bits
            if k == "m_bids":
                np.testing.assert_array_equal(results[r][k], full[k], err_msg=f"rank {r} {k}")
            else:
                np.testing.assert_allclose(results[r][k], full[k], rtol=1e-3, atol=2e-3, err_msg=f"rank {r} {k}")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_matcher_on_second_device_of_one_process():
    """The library must follow the buffers' device (it has its own CUDA runtime instance)."""
    import util
    from cases import build_inputs
    case = {"name": "dev1", "n": 1, "hw0": (96, 128), "hw1": (96, 128), "thr": 0.0, "images": "smooth"}
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        model, _, _ = util.build_model(case, dev)
        data = {k: torch.from_numpy(v).to(dev) for k, v in build_inputs(case).items()}
        with torch.cuda.device(dev):
            model(data)
        assert data["mconf"].device == torch.device(dev)
        outs.append({k: data[k].cpu().numpy() for k in ("mkpts0_f", "mkpts1_f", "mconf")})
    for k in outs[0]:
        np.testing.assert_array_equal(outs[0][k], outs[1][k])
