"""CPU oracle of the evaluation metrics -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as loftr_oracle.py).

numpy (float64 unless noted) restatement of the reference's src/utils/metrics.py (zju3dv/LoFTR), pinned by the golden
vectors of tests/golden/metrics_*.npz, which tests/golden/make_golden_metrics.py produced by calling the reference's
own functions.  RANSAC pose estimation itself is OpenCV (`cv2.findEssentialMat` / `cv2.recoverPose`, a third-party
dependency of the reference, metrics.py:88-103); its call sequence is restated in `estimate_pose`.
"""
from __future__ import annotations

import numpy as np


def relative_pose_error(T_0to1, R, t, ignore_gt_t_thr=0.0):
    """metrics.py:12-27: angular errors (degrees) of a recovered (R, t) against the ground-truth 4x4 pose."""
    t_gt = T_0to1[:3, 3]
    n = np.linalg.norm(t) * np.linalg.norm(t_gt)
    t_err = np.rad2deg(np.arccos(np.clip(np.dot(t, t_gt) / n, -1.0, 1.0)))
    t_err = np.minimum(t_err, 180 - t_err)                  # the essential matrix leaves the sign of t open  :17
    if np.linalg.norm(t_gt) < ignore_gt_t_thr:              # :18-19
        t_err = 0
    R_gt = T_0to1[:3, :3]
    cos = np.clip((np.trace(np.dot(R.T, R_gt)) - 1) / 2, -1.0, 1.0)   # :23-24
    return t_err, np.rad2deg(np.abs(np.arccos(cos)))


def essential_from_pose(T_0to1):
    """E = [t]_x R (metrics.py:56-57: kornia cross_product_matrix(t) @ R)."""
    t = T_0to1[:3, 3]
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]], dtype=T_0to1.dtype)
    return tx @ T_0to1[:3, :3]


def symmetric_epipolar_distance(pts0, pts1, E, K0, K1):
    """metrics.py:30-48: squared symmetric epipolar distance of [N, 2] pixel coordinates (normalised by K)."""
    p0 = (pts0 - K0[[0, 1], [2, 2]][None]) / K0[[0, 1], [0, 1]][None]     # :37
    p1 = (pts1 - K1[[0, 1], [2, 2]][None]) / K1[[0, 1], [0, 1]][None]     # :38
    p0 = np.concatenate([p0, np.ones_like(p0[:, :1])], 1)                  # :39-40
    p1 = np.concatenate([p1, np.ones_like(p1[:, :1])], 1)
    Ep0 = p0 @ E.T                                                         # :42
    p1Ep0 = (p1 * Ep0).sum(-1)                                             # :43
    Etp1 = p1 @ E                                                          # :44
    return p1Ep0 ** 2 * (1.0 / (Ep0[:, 0] ** 2 + Ep0[:, 1] ** 2) + 1.0 / (Etp1[:, 0] ** 2 + Etp1[:, 1] ** 2))   # :46


def symmetrical_epipolar_errors(mkpts0_f, mkpts1_f, m_bids, T_0to1, K0, K1):
    """compute_symmetrical_epipolar_errors (metrics.py:51-72): per-match errors, pair by pair, in match order
    (matches arrive grouped by pair, so concatenating per pair preserves the list order)."""
    out = np.zeros(len(m_bids), dtype=mkpts0_f.dtype)
    for b in range(T_0to1.shape[0]):
        sel = m_bids == b
        out[sel] = symmetric_epipolar_distance(mkpts0_f[sel], mkpts1_f[sel], essential_from_pose(T_0to1[b]), K0[b], K1[b])
    return out


def estimate_pose(kpts0, kpts1, K0, K1, thresh, conf=0.99999):
    """metrics.py:75-103 (OpenCV RANSAC on K-normalised points, threshold normalised by the mean focal length)."""
    import cv2
    if len(kpts0) < 5:
        return None
    k0 = (kpts0 - K0[[0, 1], [2, 2]][None]) / K0[[0, 1], [0, 1]][None]
    k1 = (kpts1 - K1[[0, 1], [2, 2]][None]) / K1[[0, 1], [0, 1]][None]
    ransac_thr = thresh / np.mean([K0[0, 0], K1[1, 1], K0[0, 0], K1[1, 1]])       # :83 (sic: K0 fx and K1 fy, twice)
    E, mask = cv2.findEssentialMat(k0, k1, np.eye(3), threshold=ransac_thr, prob=conf, method=cv2.RANSAC)
    if E is None:
        return None
    best, ret = 0, None
    for _E in np.split(E, len(E) / 3):                                            # :96
        n, R, t, _ = cv2.recoverPose(_E, k0, k1, np.eye(3), 1e9, mask=mask)
        if n > best:
            ret, best = (R, t[:, 0], mask.ravel() > 0), n
    return ret


def error_auc(errors, thresholds=(5, 10, 20)):
    """metrics.py:143-161: area under the recall-vs-error curve up to each threshold (the reference ignores its
    `thresholds` argument and always uses 5/10/20)."""
    errors = [0] + sorted(list(errors))
    recall = list(np.linspace(0, 1, len(errors)))
    out = {}
    for thr in (5, 10, 20):
        last = np.searchsorted(errors, thr)
        y = recall[:last] + [recall[last - 1]]
        x = errors[:last] + [thr]
        out[f"auc@{thr}"] = np.trapezoid(y, x) / thr if hasattr(np, "trapezoid") else np.trapz(y, x) / thr
    return out


def epidist_prec(errors, thresholds):
    """metrics.py:164-176: mean over pairs of the fraction of matches below each threshold."""
    out = {}
    for thr in thresholds:
        per_pair = [np.mean(e < thr) if len(e) > 0 else 0 for e in errors]
        out[f"prec@{thr:.0e}"] = np.mean(per_pair) if len(per_pair) > 0 else 0
    return out


def aggregate_metrics(metrics, epi_err_thr=5e-4):
    """metrics.py:179-200: de-duplicate by identifier (last occurrence wins), pose AUC of max(R_err, t_err),
    matching precision at `epi_err_thr`."""
    last = {}
    for idx, iden in enumerate(metrics["identifiers"]):
        last[iden] = idx
    order = list(dict.fromkeys(metrics["identifiers"]))          # first-seen order of the unique identifiers ...
    unq = [last[i] for i in order]                               # ... holding the LAST index of each (OrderedDict update)
    pose_err = np.max(np.stack([metrics["R_errs"], metrics["t_errs"]]), axis=0)[unq]
    epi = [metrics["epi_errs"][i] for i in unq]
    return {**error_auc(pose_err), **epidist_prec(epi, [epi_err_thr])}
