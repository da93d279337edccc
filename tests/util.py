"""Shared helpers of the parity tests."""
from __future__ import annotations

import os

import numpy as np
import torch

import weights as W
from cases import build_cfg, build_cm_inputs, build_inputs  # noqa: F401

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def build_model(case, device="cpu"):
    """loftr_b200.LoFTR with the deterministic RandomState weights of tests/golden/weights.py."""
    import loftr_b200
    cfg = build_cfg(case)
    model = loftr_b200.LoFTR(cfg).eval()
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    state = W.make_state(shapes, seed=case.get("wseed", 0))
    if "bin_score" in case and "coarse_matching.bin_score" in state:
        state["coarse_matching.bin_score"] = np.asarray(case["bin_score"], np.float32)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()})
    return model.to(device), cfg, state


def oracle_forward(case, backbone_device="cpu"):
    """PyTorch backbone (the part that stays PyTorch) + numpy oracle for everything after it."""
    from oracle import loftr_oracle as O
    model, cfg, state = build_model(case, backbone_device)
    inp = build_inputs(case)
    with torch.no_grad():
        i0 = torch.from_numpy(inp["image0"]).to(backbone_device)
        i1 = torch.from_numpy(inp["image1"]).to(backbone_device)
        if i0.shape == i1.shape:
            fc, ff = model.backbone(torch.cat([i0, i1], 0))
            (c0, c1), (f0, f1) = fc.split(i0.shape[0]), ff.split(i0.shape[0])
        else:
            (c0, f0), (c1, f1) = model.backbone(i0), model.backbone(i1)
    npf = lambda t: t.float().cpu().numpy()
    out = O.hot_path(npf(c0), npf(c1), npf(f0), npf(f1), state, cfg, inp["image0"].shape[2:], inp["image1"].shape[2:],
                     inp.get("mask0"), inp.get("mask1"), inp.get("scale0"), inp.get("scale1"))
    return out


def match_keys(b, i, j):
    return {(int(x), int(y), int(z)) for x, y, z in zip(b, i, j)}


def compare_matches(got, ref, gold=None, conf_rtol=1e-3, px_tol=0.5, min_overlap=0.995, label=""):
    """Compare two match lists by (b, i, j) key (SURVEY.md §7 hard part 2).

    - the key sets must overlap >= min_overlap (relative to the larger set);
    - every non-shared match must be a near-tie / near-threshold case according to the fp64 reference
      statistics in `gold` (row_top2_f64 / col_top2_f64) when they are available;
    - on the intersection: mconf within rtol, keypoints within px_tol pixels.
    """
    kg = {k: n for n, k in enumerate(zip(got["b_ids"].tolist(), got["i_ids"].tolist(), got["j_ids"].tolist()))}
    kr = {k: n for n, k in enumerate(zip(ref["b_ids"].tolist(), ref["i_ids"].tolist(), ref["j_ids"].tolist()))}
    common = sorted(set(kg) & set(kr))
    big = max(len(kg), len(kr))
    if big == 0:
        return {"n": 0, "overlap": 1.0}
    overlap = len(common) / big
    assert overlap >= min_overlap, f"{label}: match-set overlap {overlap:.4f} ({len(kg)} vs {len(kr)} matches)"
    if gold is not None and "row_top2_f64" in gold:
        for (b, i, j) in set(kg) ^ set(kr):
            r, c = gold["row_top2_f64"][b, i], gold["col_top2_f64"][b, j]
            tie = min(abs(r[0] - r[1]) / max(r[0], 1e-30), abs(c[0] - c[1]) / max(c[0], 1e-30))
            assert tie < 1e-2, f"{label}: match {(b, i, j)} differs and is not a near-tie (gap {tie:.3e})"
    ig = np.array([kg[k] for k in common])
    ir = np.array([kr[k] for k in common])
    stats = {"n": len(common), "overlap": overlap}
    # the list order must be ascending (b, i) like torch.where
    order = [k[:2] for k in sorted(kg, key=kg.get)]
    assert order == sorted(order), f"{label}: match list is not ordered by (b, i)"
    for key, tol in (("mconf", None), ("mkpts0_c", px_tol), ("mkpts1_c", px_tol), ("mkpts0_f", px_tol),
                     ("mkpts1_f", px_tol)):
        if key not in got or key not in ref:
            continue
        g, r = np.asarray(got[key])[ig], np.asarray(ref[key])[ir]
        if key == "mconf":
            rel = np.abs(g - r) / np.maximum(np.abs(r), 1e-30)
            stats["mconf_rel_max"] = float(rel.max())
            assert rel.max() <= conf_rtol, f"{label}: mconf rel err {rel.max():.3e} > {conf_rtol}"
        else:
            d = np.abs(g - r).max() if len(g) else 0.0
            stats[key + "_max"] = float(d)
            assert d < tol, f"{label}: {key} differs by {d:.4f} px"
    return stats


def record(name, stats):
    """Append parity statistics to gpurun_out/parity_stats.jsonl (kept as evidence; prints are captured by pytest)."""
    import json
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_stats.jsonl"), "a") as f:
            f.write(json.dumps({"test": name, **{k: (float(v) if isinstance(v, (int, float)) else v) for k, v in stats.items()}}) + "\n")
    except OSError:
        pass
