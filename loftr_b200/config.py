"""Matcher configuration dictionaries (lower-case keys, the form `LoFTR(config)` consumes).

`default_cfg` mirrors src/loftr/utils/cvpr_ds_config.py:9-50 (what `from src.loftr import default_cfg`
exports); `get_cfg(name)` builds the matcher part of the experiment configs under configs/loftr/
on top of src/config/default.py:5-44 (indoor_ds / outdoor_ds / indoor_ot / outdoor_ot).
"""
from __future__ import annotations

import copy


def _base(temp_bug_fix, thr=0.2, train_coarse_percent=0.2, prefilter=False):
    return {
        "backbone_type": "ResNetFPN",
        "resolution": (8, 2),
        "fine_window_size": 5,
        "fine_concat_coarse_feat": True,
        "resnetfpn": {"initial_dim": 128, "block_dims": [128, 196, 256]},
        "coarse": {"d_model": 256, "d_ffn": 256, "nhead": 8, "layer_names": ["self", "cross"] * 4,
                   "attention": "linear", "temp_bug_fix": temp_bug_fix},
        "match_coarse": {"thr": thr, "border_rm": 2, "match_type": "dual_softmax", "dsmax_temperature": 0.1,
                         "skh_iters": 3, "skh_init_bin_score": 1.0, "skh_prefilter": prefilter,
                         "train_coarse_percent": train_coarse_percent, "train_pad_num_gt_min": 200,
                         "sparse_spvs": True},
        "fine": {"d_model": 128, "d_ffn": 128, "nhead": 8, "layer_names": ["self", "cross"], "attention": "linear"},
    }


# cvpr_ds_config.py: TEMP_BUG_FIX False, SKH_PREFILTER True, TRAIN_COARSE_PERCENT 0.4, no SPARSE_SPVS key
default_cfg = _base(temp_bug_fix=False, train_coarse_percent=0.4, prefilter=True)
del default_cfg["match_coarse"]["sparse_spvs"]


def get_cfg(name: str = "indoor_ds", thr: float | None = None):
    """Matcher dict of configs/loftr/{indoor,outdoor}/loftr_{ds,ot}.py over src/config/default.py."""
    cfg = _base(temp_bug_fix=True)
    scene, kind = name.split("_")
    if scene not in ("indoor", "outdoor") or kind not in ("ds", "ot"):
        raise KeyError(name)
    cfg["match_coarse"]["match_type"] = "dual_softmax" if kind == "ds" else "sinkhorn"
    if scene == "outdoor":
        cfg["match_coarse"]["train_coarse_percent"] = 0.3  # configs/loftr/outdoor/loftr_{ds,ot}.py; unused at inference
    if thr is not None:
        cfg["match_coarse"]["thr"] = thr
    return copy.deepcopy(cfg)
