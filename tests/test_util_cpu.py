"""CPU checks of the test infrastructure itself: the per-pair / row-blocked oracle helpers used by the full-size
GPU parity tests must agree with the plain oracle and with a direct fp64 evaluation."""
import numpy as np

import util
from cases import build_cfg


def test_per_pair_oracle_equals_batched_oracle():
    case = {"name": "pp", "n": 3, "hw0": (64, 96), "hw1": (64, 96), "thr": 0.0, "images": "smooth",
            "valid0": [(64, 96), (48, 96), (64, 72)], "valid1": [(64, 80), (64, 96), (56, 96)], "scales": True}
    whole = util.oracle_forward(case)
    parts, gold = util.oracle_forward_per_pair(case)
    assert len(whole["b_ids"]) > 20
    for k in ["b_ids", "i_ids", "j_ids"]:
        np.testing.assert_array_equal(whole[k], parts[k])
    for k in ["mconf", "mkpts0_c", "mkpts1_c", "mkpts0_f", "mkpts1_f"]:
        np.testing.assert_allclose(whole[k], parts[k], rtol=2e-5, atol=2e-5)
    assert gold["row_top2_f64"].shape == (3, 8 * 12, 2) and gold["col_top2_f64"].shape == (3, 8 * 12, 2)


def test_blocked_fp64_top2_equals_direct_evaluation():
    rs = np.random.RandomState(0)
    L, S, c = 150, 131, 256
    x0 = (rs.standard_normal((L, c)) * 1.5 + 2).astype(np.float32)
    x1 = (rs.standard_normal((S, c)) * 1.5 + 2).astype(np.float32)
    m0 = rs.uniform(size=L) > 0.1
    m1 = rs.uniform(size=S) > 0.1
    cfg = build_cfg({})["match_coarse"]
    for masks in ((None, None), (m0, m1)):
        r, cc = util.near_tie_top2_f64(x0, x1, cfg, *masks, block=37)
        sim = (x0.astype(np.float64) / 16) @ (x1.astype(np.float64) / 16).T / 0.1
        if masks[0] is not None:
            sim[~(m0[:, None] & m1[None, :])] = -1e9
        e1 = np.exp(sim - sim.max(1, keepdims=True))
        e0 = np.exp(sim - sim.max(0, keepdims=True))
        conf = (e1 / e1.sum(1, keepdims=True)) * (e0 / e0.sum(0, keepdims=True))
        np.testing.assert_allclose(r, -np.sort(-conf, 1)[:, :2], rtol=1e-6, atol=1e-300)
        np.testing.assert_allclose(cc, (-np.sort(-conf, 0)[:2]).T, rtol=1e-6, atol=1e-300)
