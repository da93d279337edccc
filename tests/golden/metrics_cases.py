"""Synthetic two-view geometry for the evaluation-metric tests (shared by make_golden_metrics.py and the tests)."""
from __future__ import annotations

import numpy as np


def _rot(rs, scale):
    w = rs.standard_normal(3) * scale
    th = np.linalg.norm(w)
    k = w / max(th, 1e-12)
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def make_scene_batch(seed, n_pairs, width=640, height=480):
    """-> dict of float32 arrays: mkpts0_f / mkpts1_f [M, 2] (pixel coordinates of projected 3-D points, noisy, with
    outliers), m_bids [M] int64 (grouped by pair), T_0to1 [N, 4, 4], K0 / K1 [N, 3, 3]."""
    rs = np.random.RandomState(seed)
    mk0, mk1, bids, Ts, K0s, K1s = [], [], [], [], [], []
    for b in range(n_pairs):
        n = int(rs.randint(60, 400))
        K0 = np.array([[570 + 20 * rs.rand(), 0, 320 + 5 * rs.randn()], [0, 575 + 20 * rs.rand(), 240 + 5 * rs.randn()], [0, 0, 1]])
        K1 = np.array([[580 + 20 * rs.rand(), 0, 318 + 5 * rs.randn()], [0, 572 + 20 * rs.rand(), 242 + 5 * rs.randn()], [0, 0, 1]])
        R = _rot(rs, 0.25)
        t = rs.standard_normal(3) * 0.4
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = R, t
        X = np.stack([rs.uniform(-2, 2, n), rs.uniform(-1.5, 1.5, n), rs.uniform(3, 8, n)], 1)
        x0 = (K0 @ X.T).T
        x0 = x0[:, :2] / x0[:, 2:]
        X1 = (R @ X.T).T + t
        x1 = (K1 @ X1.T).T
        x1 = x1[:, :2] / x1[:, 2:]
        x1 += rs.standard_normal(x1.shape) * 0.4                      # sub-pixel noise
        bad = rs.rand(n) < 0.2                                         # 20 % gross outliers
        x1[bad] = np.stack([rs.uniform(0, width, bad.sum()), rs.uniform(0, height, bad.sum())], 1)
        mk0.append(x0)
        mk1.append(x1)
        bids.append(np.full(n, b))
        Ts.append(T)
        K0s.append(K0)
        K1s.append(K1)
    f = np.float32
    return {"mkpts0_f": np.concatenate(mk0).astype(f), "mkpts1_f": np.concatenate(mk1).astype(f),
            "m_bids": np.concatenate(bids).astype(np.int64), "T_0to1": np.stack(Ts).astype(f),
            "K0": np.stack(K0s).astype(f), "K1": np.stack(K1s).astype(f)}


def AGG_CASE():
    """A metrics dict as accumulated over a test set (incl. a duplicated identifier and a failed pair)."""
    rs = np.random.RandomState(5)
    n = 40
    ids = [f"scene{i // 4:02d}#{i % 4}" for i in range(n)]
    ids[7] = ids[3]                                                   # a duplicate: the later entry wins
    R = np.abs(rs.standard_normal(n)) * 8
    t = np.abs(rs.standard_normal(n)) * 12
    R[11], t[11] = np.inf, np.inf                                     # pose estimation failed
    epi = [np.abs(rs.standard_normal(int(rs.randint(0, 50)))) * 1e-3 for _ in range(n)]
    return {"identifiers": ids, "R_errs": list(R), "t_errs": list(t), "epi_errs": epi}
