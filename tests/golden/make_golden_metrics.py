"""Golden vectors for the evaluation metrics: runs the UNMODIFIED reference functions of
src/utils/metrics.py (oracle/ref_import.py: kornia stand-ins, np.bool alias) on synthetic two-view geometry and stores
inputs + outputs in tests/golden/metrics_scenes.npz.  Authoring container only.

    python tests/golden/make_golden_metrics.py
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import ref_import  # noqa: E402
from metrics_cases import make_scene_batch, AGG_CASE  # noqa: E402


def main():
    import cv2
    M = ref_import.load_reference_metrics()
    out = {}
    for tag, seed, n_pairs in (("a", 0, 3), ("b", 1, 2)):
        sc = make_scene_batch(seed, n_pairs)
        data = {k: torch.from_numpy(v) for k, v in sc.items()}
        M.compute_symmetrical_epipolar_errors(data)
        cfg = types.SimpleNamespace(TRAINER=types.SimpleNamespace(RANSAC_PIXEL_THR=0.5, RANSAC_CONF=0.99999))
        cv2.setRNGSeed(0)
        M.compute_pose_errors(data, cfg)
        out[f"{tag}_epi_errs"] = data["epi_errs"].numpy()
        out[f"{tag}_R_errs"] = np.asarray(data["R_errs"], np.float64)
        out[f"{tag}_t_errs"] = np.asarray(data["t_errs"], np.float64)
        out[f"{tag}_n_inliers"] = np.asarray([int(np.sum(i)) for i in data["inliers"]], np.int64)
        # relative_pose_error on a perturbed ground truth
        rs = np.random.RandomState(seed + 10)
        for b in range(n_pairs):
            T = sc["T_0to1"][b].astype(np.float64)
            Rp, _ = cv2.Rodrigues(rs.standard_normal(3) * 0.05)
            te, re_ = M.relative_pose_error(T, Rp @ T[:3, :3], T[:3, 3] + 0.05 * rs.standard_normal(3))
            out[f"{tag}_rpe_{b}"] = np.asarray([te, re_], np.float64)
    # aggregation
    m = {k: (list(v) if not isinstance(v, list) else v) for k, v in AGG_CASE().items()}
    agg = M.aggregate_metrics(m, epi_err_thr=5e-4)
    for k, v in agg.items():
        out["agg_" + k] = np.asarray(v, np.float64)
    auc = M.error_auc(np.asarray(AGG_CASE()["R_errs"]), [5, 10, 20])
    for k, v in auc.items():
        out["aucR_" + k] = np.asarray(v, np.float64)
    path = os.path.join(HERE, "metrics_scenes.npz")
    np.savez_compressed(path, **out)
    print({k: (v.shape if v.ndim else float(v)) for k, v in out.items()})
    print("->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
