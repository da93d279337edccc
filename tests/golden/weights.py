"""Deterministic, torch-independent weights and inputs for parity tests.

numpy's legacy RandomState stream is frozen across numpy versions, so the authoring container (where the
reference runs and the golden vectors are made) and the GPU box build bit-identical tensors without
shipping 46 MB of parameters.  Initialisation scales follow the reference's own init rules (xavier for
the transformers transformer.py:75-78, kaiming fan_out for convs resnet_fpn.py:88-93), but BatchNorm /
LayerNorm get non-trivial statistics and affine terms so those code paths are really exercised.
"""
from __future__ import annotations

import numpy as np


def make_state(shapes: dict, seed: int = 0) -> dict:
    """shapes: name -> tuple (e.g. from model.state_dict()).  Returns name -> np.ndarray (float32 / int64)."""
    rs = np.random.RandomState(seed)
    out = {}
    for name in sorted(shapes):
        shp = tuple(shapes[name])
        if name.endswith("num_batches_tracked"):
            out[name] = np.zeros(shp, np.int64)
        elif name.endswith("running_mean"):
            out[name] = (0.1 * rs.standard_normal(shp)).astype(np.float32)
        elif name.endswith("running_var"):
            out[name] = (1.0 + 0.1 * np.abs(rs.standard_normal(shp))).astype(np.float32)
        elif name.endswith("bin_score"):
            out[name] = np.asarray(1.0, np.float32).reshape(shp)
        elif len(shp) == 4:  # conv [out, in, kh, kw]
            fan_out = shp[0] * shp[2] * shp[3]
            out[name] = (rs.standard_normal(shp) * np.sqrt(2.0 / fan_out)).astype(np.float32)
        elif len(shp) == 2:  # linear [out, in]
            a = np.sqrt(6.0 / (shp[0] + shp[1]))
            out[name] = rs.uniform(-a, a, shp).astype(np.float32)
        elif len(shp) == 1 and name.endswith("weight"):  # BN / LN gain
            out[name] = (1.0 + 0.1 * rs.standard_normal(shp)).astype(np.float32)
        elif len(shp) == 1:  # biases
            out[name] = (0.1 * rs.standard_normal(shp)).astype(np.float32)
        else:
            raise ValueError(f"no rule for {name} {shp}")
    return out


def make_images(n, h, w, seed):
    rs = np.random.RandomState(seed)
    return rs.uniform(0, 1, (n, 1, h, w)).astype(np.float32), rs.uniform(0, 1, (n, 1, h, w)).astype(np.float32)


def smooth_images(n, h, w, seed, shift=(3, 5)):
    """A more realistic pair: low-pass noise and a shifted, slightly perturbed copy (gives real matches)."""
    rs = np.random.RandomState(seed)
    base = rs.uniform(0, 1, (n, 1, h + 32, w + 32)).astype(np.float32)
    k = np.ones((5, 5), np.float32) / 25.0
    sm = np.zeros_like(base)
    for dy in range(5):
        for dx in range(5):
            sm += k[dy, dx] * np.roll(np.roll(base, dy - 2, 2), dx - 2, 3)
    sm = (sm - sm.min()) / (sm.max() - sm.min())
    im0 = sm[:, :, 16:16 + h, 16:16 + w]
    im1 = sm[:, :, 16 + shift[0]:16 + shift[0] + h, 16 + shift[1]:16 + shift[1] + w]
    im1 = im1 + 0.02 * rs.standard_normal(im1.shape).astype(np.float32)
    return np.ascontiguousarray(im0), np.ascontiguousarray(np.clip(im1, 0, 1).astype(np.float32))
