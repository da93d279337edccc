"""Evaluation harness: epipolar errors, RANSAC relative pose, AUC@5/10/20 and matching precision
(SURVEY.md §8(f) rank 3).  Mirrors the reference's `src/utils/metrics.py` (function names, argument meaning,
the keys written into `data`) so that `src/lightning/lightning_loftr.py:_compute_metrics` (:101-121) and
`test_epoch_end` (:232-253) can call it unchanged:

    compute_symmetrical_epipolar_errors(batch)     # -> batch['epi_errs']            CUDA kernel lb_epipolar_errors
    compute_pose_errors(batch, config)             # -> batch['R_errs' / 't_errs' / 'inliers']   OpenCV RANSAC (CPU)
    aggregate_metrics(metrics, epi_err_thr)        # -> {'auc@5', 'auc@10', 'auc@20', 'prec@5e-04'}

`evaluate_pairs` is the end-to-end loop of `test.py` over a pair list such as the reference's
`assets/scannet_test_1500` (`load_scannet_pair_list` reads its `test.npz` / `intrinsics.npz` layout).
The epipolar errors run on the GPU (one thread per match, no CPU fallback); pose estimation is OpenCV's
`findEssentialMat` / `recoverPose` exactly as in the reference (a CPU library call there too).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib

__all__ = ["compute_symmetrical_epipolar_errors", "estimate_pose", "relative_pose_error", "compute_pose_errors",
           "error_auc", "epidist_prec", "aggregate_metrics", "load_scannet_pair_list", "evaluate_pairs"]


# ------------------------------------------------------------------------------------------------ per-batch metrics
@torch.no_grad()
def compute_symmetrical_epipolar_errors(data: dict) -> None:
    """Writes data['epi_errs'] [M] (squared symmetric epipolar distance in normalised image coordinates).
    Needs mkpts0_f, mkpts1_f, m_bids and the ground truth T_0to1 [N,4,4], K0, K1 [N,3,3]
    (reference metrics.py:51-72)."""
    mk0, mk1, bids = data["mkpts0_f"], data["mkpts1_f"], data["m_bids"]
    if not mk0.is_cuda:
        raise RuntimeError("loftr_b200: the epipolar-error kernel needs CUDA tensors (no CPU implementation)")
    dev = mk0.device
    m = int(mk0.shape[0])
    errs = torch.empty(m, dtype=torch.float32, device=dev)
    if m:
        f = lambda t: t.to(dev, torch.float32).contiguous()
        T, K0, K1 = f(data["T_0to1"]), f(data["K0"]), f(data["K1"])
        mk0, mk1, bids = f(mk0), f(mk1), bids.to(dev, torch.int64).contiguous()
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(_lib.load().lb_epipolar_errors(mk0.data_ptr(), mk1.data_ptr(), bids.data_ptr(), m, int(T.shape[0]),
                                                  T.data_ptr(), K0.data_ptr(), K1.data_ptr(), errs.data_ptr(), st))
    data.update({"epi_errs": errs})


def relative_pose_error(T_0to1, R, t, ignore_gt_t_thr=0.0):
    """(t_err, R_err) in degrees of a recovered pose against the ground truth (metrics.py:12-27); the translation
    error is direction-only and folded to [0, 90] because the essential matrix leaves the sign of t open."""
    t_gt = T_0to1[:3, 3]
    cos_t = float(np.dot(t, t_gt)) / (np.linalg.norm(t) * np.linalg.norm(t_gt))
    t_err = np.degrees(np.arccos(np.clip(cos_t, -1.0, 1.0)))
    t_err = min(t_err, 180.0 - t_err)
    if np.linalg.norm(t_gt) < ignore_gt_t_thr:
        t_err = 0
    cos_r = np.clip((np.trace(R.T @ T_0to1[:3, :3]) - 1.0) / 2.0, -1.0, 1.0)
    return t_err, np.degrees(abs(np.arccos(cos_r)))


def estimate_pose(kpts0, kpts1, K0, K1, thresh, conf=0.99999):
    """RANSAC essential matrix + cheirality test on K-normalised points; -> (R, t, inlier mask) or None
    (metrics.py:75-103).  The pixel threshold is normalised by mean(K0 fx, K1 fy) as in the reference (:83)."""
    import cv2
    if len(kpts0) < 5:
        return None
    c0, f0 = K0[[0, 1], [2, 2]], K0[[0, 1], [0, 1]]
    c1, f1 = K1[[0, 1], [2, 2]], K1[[0, 1], [0, 1]]
    n0, n1 = (kpts0 - c0[None]) / f0[None], (kpts1 - c1[None]) / f1[None]
    thr = thresh / np.mean([K0[0, 0], K1[1, 1], K0[0, 0], K1[1, 1]])
    E, mask = cv2.findEssentialMat(n0, n1, np.eye(3), threshold=thr, prob=conf, method=cv2.RANSAC)
    if E is None:
        return None
    best, found = 0, None
    for k in range(E.shape[0] // 3):   # findEssentialMat may return several stacked candidates
        n, R, t, _ = cv2.recoverPose(E[3 * k:3 * k + 3], n0, n1, np.eye(3), 1e9, mask=mask)
        if n > best:
            best, found = n, (R, t[:, 0], mask.ravel() > 0)
    return found


def compute_pose_errors(data: dict, config=None, pixel_thr=None, conf=None) -> None:
    """Writes data['R_errs'], data['t_errs'] (lists of floats, inf when no pose was found) and data['inliers']
    (metrics.py:106-138).  `config` may be the reference's yacs node (TRAINER.RANSAC_PIXEL_THR / RANSAC_CONF)."""
    if pixel_thr is None:
        pixel_thr = config.TRAINER.RANSAC_PIXEL_THR if config is not None else 0.5
    if conf is None:
        conf = config.TRAINER.RANSAC_CONF if config is not None else 0.99999
    g = lambda k: data[k].detach().cpu().numpy()
    bids, p0, p1, K0, K1, T = g("m_bids"), g("mkpts0_f"), g("mkpts1_f"), g("K0"), g("K1"), g("T_0to1")
    R_errs, t_errs, inliers = [], [], []
    for b in range(K0.shape[0]):
        sel = bids == b
        ret = estimate_pose(p0[sel], p1[sel], K0[b], K1[b], pixel_thr, conf=conf)
        if ret is None:
            R_errs.append(np.inf)
            t_errs.append(np.inf)
            inliers.append(np.zeros(0, dtype=bool))
        else:
            R, t, inl = ret
            te, re_ = relative_pose_error(T[b], R, t, ignore_gt_t_thr=0.0)
            R_errs.append(re_)
            t_errs.append(te)
            inliers.append(inl)
    data.update({"R_errs": R_errs, "t_errs": t_errs, "inliers": inliers})


# ------------------------------------------------------------------------------------------------ aggregation
def error_auc(errors, thresholds=(5, 10, 20)):
    """Area under the cumulative pose-error curve up to 5 / 10 / 20 degrees, normalised by the threshold
    (metrics.py:143-161; like the reference, always those three thresholds)."""
    e = np.concatenate([[0.0], np.sort(np.asarray(errors, dtype=np.float64))])
    recall = np.linspace(0.0, 1.0, len(e))
    out = {}
    for thr in (5, 10, 20):
        k = int(np.searchsorted(e, thr))                       # points strictly below the threshold
        x = np.concatenate([e[:k], [thr]])
        y = np.concatenate([recall[:k], [recall[k - 1]]])      # the curve is held flat up to the threshold
        out[f"auc@{thr}"] = float(np.sum((x[1:] - x[:-1]) * (y[1:] + y[:-1]) * 0.5) / thr)
    return out


def epidist_prec(errors, thresholds, ret_dict=False):
    """Mean over pairs of the fraction of matches whose epipolar error is below each threshold (metrics.py:164-176)."""
    precs = []
    for thr in thresholds:
        per_pair = [float(np.mean(np.asarray(e) < thr)) if len(e) > 0 else 0 for e in errors]
        precs.append(float(np.mean(per_pair)) if per_pair else 0)
    return {f"prec@{t:.0e}": p for t, p in zip(thresholds, precs)} if ret_dict else precs


def aggregate_metrics(metrics: dict, epi_err_thr=5e-4):
    """Whole-dataset numbers (metrics.py:179-200): duplicates (same identifier) count once -- the LAST occurrence, in
    first-seen order -- then pose AUC of max(R_err, t_err) and the matching precision at `epi_err_thr`
    (5e-4 ScanNet, 1e-4 MegaDepth)."""
    last = {}
    for i, iden in enumerate(metrics["identifiers"]):
        last[iden] = i
    keep = [last[i] for i in dict.fromkeys(metrics["identifiers"])]
    pose = np.maximum(np.asarray(metrics["R_errs"], dtype=np.float64), np.asarray(metrics["t_errs"], dtype=np.float64))[keep]
    epi = [metrics["epi_errs"][i] for i in keep]
    return {**error_auc(pose), **epidist_prec(epi, [epi_err_thr], True)}


# ------------------------------------------------------------------------------------------------ pair lists / driver
def load_scannet_pair_list(npz_path, intrinsics_path):
    """The reference's test-pair format (`assets/scannet_test_1500/test.npz` + `intrinsics.npz`,
    src/datasets/scannet.py:43-48,77-80,96): -> list of dicts with scene, the two colour-image paths relative to the
    ScanNet root, K (3x3) and -- when the file carries it -- T_0to1 from `rel_pose` (3x4 row-major)."""
    with np.load(npz_path) as d:
        names = d["name"]
        rel = d["rel_pose"] if "rel_pose" in d.files else None
    intr = dict(np.load(intrinsics_path))
    pairs = []
    for i, (scene, sub, s0, s1) in enumerate(names):
        scene_name = f"scene{int(scene):04d}_{int(sub):02d}"
        item = {"scene_id": scene_name, "pair_id": i, "K": intr[scene_name].astype(np.float32).reshape(3, 3),
                "pair_names": (os.path.join(scene_name, "color", f"{int(s0)}.jpg"), os.path.join(scene_name, "color", f"{int(s1)}.jpg")),
                "pose_names": (os.path.join(scene_name, "pose", f"{int(s0)}.txt"), os.path.join(scene_name, "pose", f"{int(s1)}.txt"))}
        if rel is not None:
            T = np.eye(4, dtype=np.float32)
            T[:3] = rel[i].reshape(3, 4)
            item["T_0to1_from_list"] = T
        pairs.append(item)
    return pairs


def _read_gray(path, size=(640, 480)):
    import cv2
    img = cv2.imread(path, cv2.IMREAD_GRAYSCALE)           # src/utils/dataset.py read_scannet_gray: resize to 640x480, /255
    if img is None:
        raise FileNotFoundError(path)
    img = cv2.resize(img, size)
    return torch.from_numpy(img).float()[None] / 255.0


def _rel_pose_from_files(root, pose_names):
    # cam-to-world 4x4 text matrices; ScanNetDataset._compute_rel_pose inverts them first (src/utils/dataset.py
    # read_scannet_pose returns world-to-cam): T_0to1 = pose1 @ inv(pose0)
    w2c = [np.linalg.inv(np.loadtxt(os.path.join(root, p), delimiter=" ")) for p in pose_names]
    return (w2c[1] @ np.linalg.inv(w2c[0])).astype(np.float32)


@torch.no_grad()
def evaluate_pairs(matcher, pairs, root_dir, batch_size=8, device="cuda", pose_dir=None, ransac_pixel_thr=0.5,
                   ransac_conf=0.99999, epi_err_thr=5e-4, rel_pose_from_list=False):
    """`test.py` as a function: match every pair of `pairs` (load_scannet_pair_list) with `matcher`, accumulate the
    per-pair metrics like PL_LoFTR._compute_metrics (lightning_loftr.py:101-121) and aggregate them.
    Ground-truth poses come from the ScanNet pose files under `pose_dir` (default `root_dir`), or from the pair
    list's `rel_pose` column with rel_pose_from_list=True."""
    metrics = {"R_errs": [], "t_errs": [], "inliers": [], "epi_errs": [], "identifiers": []}
    pose_dir = pose_dir or root_dir
    for lo in range(0, len(pairs), batch_size):
        chunk = pairs[lo:lo + batch_size]
        batch = {"image0": torch.stack([_read_gray(os.path.join(root_dir, p["pair_names"][0])) for p in chunk]).to(device),
                 "image1": torch.stack([_read_gray(os.path.join(root_dir, p["pair_names"][1])) for p in chunk]).to(device)}
        T = [p["T_0to1_from_list"] if rel_pose_from_list else _rel_pose_from_files(pose_dir, p["pose_names"]) for p in chunk]
        batch["T_0to1"] = torch.from_numpy(np.stack(T)).to(device)
        batch["K0"] = batch["K1"] = torch.from_numpy(np.stack([p["K"] for p in chunk])).to(device)
        matcher(batch)
        compute_symmetrical_epipolar_errors(batch)
        compute_pose_errors(batch, pixel_thr=ransac_pixel_thr, conf=ransac_conf)
        bids, epi = batch["m_bids"].cpu().numpy(), batch["epi_errs"].cpu().numpy()
        for b, p in enumerate(chunk):
            metrics["identifiers"].append("#".join(p["pair_names"]))
            metrics["epi_errs"].append(epi[bids == b])
        for k in ("R_errs", "t_errs", "inliers"):
            metrics[k] += list(batch[k])
    return aggregate_metrics(metrics, epi_err_thr), metrics
