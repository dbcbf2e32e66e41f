"""Config surface of the drop-in: the `cfg.sf_mdl` keys the reference reads on this path
(SlowFast/slowfast/config/defaults.py and
configs/vsitu_mdl_cfgs/*.yaml), restated as plain Python data.

The reference builds `cfg.sf_mdl` with yacs (vidsitu_code/extended_config.py:145-195);
only attribute access is ever used on it, so either a yacs CfgNode or the `AttrDict`
below works with `vidsitu_b200.SFBase`.  `sf_mdl_cfg(name)` accepts the same
`mdl.sf_mdl_name` keys as `sf_mdl_to_cfg_fpath_dct` (extended_config.py:14-20) plus
`slow_fast_r101_16x8` (SlowFast/configs/Kinetics/c2/SLOWFAST_16x8_R101_50_50.yaml).
"""
from __future__ import annotations

import copy
from typing import Any, Dict


class AttrDict(dict):
    """dict with attribute access, recursively (the subset of yacs.CfgNode used here)."""

    def __init__(self, d: Dict[str, Any] | None = None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = AttrDict(v) if isinstance(v, dict) and not isinstance(v, AttrDict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def merge(self, other: Dict[str, Any]) -> "AttrDict":
        for k, v in other.items():
            if isinstance(v, dict) and isinstance(self.get(k), dict):
                self[k].merge(v)
            else:
                self[k] = AttrDict(v) if isinstance(v, dict) else copy.deepcopy(v)
        return self


# Defaults of the keys on the hot path (SlowFast/slowfast/config/defaults.py).
_DEFAULTS = {
    "BN": {"NORM_TYPE": "batchnorm", "NUM_SPLITS": 1, "NUM_SYNC_DEVICES": 1},   # defaults.py:36-52
    "RESNET": {
        "TRANS_FUNC": "bottleneck_transform", "NUM_GROUPS": 1, "WIDTH_PER_GROUP": 64,
        "INPLACE_RELU": True, "STRIDE_1X1": False, "ZERO_INIT_FINAL_BN": False, "DEPTH": 50,
        "NUM_BLOCK_TEMP_KERNEL": [[3], [4], [6], [3]],                   # defaults.py:137
        "SPATIAL_STRIDES": [[1], [2], [2], [2]],                         # defaults.py:140
        "SPATIAL_DILATIONS": [[1], [1], [1], [1]],                       # defaults.py:143
    },
    "NONLOCAL": {
        "LOCATION": [[[]], [[]], [[]], [[]]], "GROUP": [[1], [1], [1], [1]], "INSTANTIATION": "dot_product",
        "POOL": [[[1, 2, 2], [1, 2, 2]]] * 4,                            # defaults.py:162-171
    },
    "MODEL": {
        "ARCH": "slowfast", "MODEL_NAME": "SlowFast", "NUM_CLASSES": 400, "DROPOUT_RATE": 0.5,
        "FC_INIT_STD": 0.01, "HEAD_ACT": "softmax",
        "SINGLE_PATHWAY_ARCH": ["c2d", "i3d", "slow"], "MULTI_PATHWAY_ARCH": ["slowfast"],
    },
    "SLOWFAST": {"BETA_INV": 8, "ALPHA": 8, "FUSION_CONV_CHANNEL_RATIO": 2, "FUSION_KERNEL_SZ": 5},
    "DATA": {
        "NUM_FRAMES": 8, "SAMPLING_RATE": 8, "TARGET_FPS": 30, "CROP_SIZE": 224,
        "MEAN": [0.45, 0.45, 0.45], "STD": [0.225, 0.225, 0.225],        # defaults.py:248,254
        "INPUT_CHANNEL_NUM": [3, 3], "REVERSE_INPUT_CHANNEL": False,
    },
    "DETECTION": {"ENABLE": False},
    "MULTIGRID": {"SHORT_CYCLE": False},
}

_R50_SINGLE = {
    "DATA": {"NUM_FRAMES": 8, "SAMPLING_RATE": 8, "INPUT_CHANNEL_NUM": [3]},
    "RESNET": {"ZERO_INIT_FINAL_BN": True, "WIDTH_PER_GROUP": 64, "NUM_GROUPS": 1, "DEPTH": 50,
               "STRIDE_1X1": False, "NUM_BLOCK_TEMP_KERNEL": [[3], [4], [6], [3]]},
}

_SLOWFAST_COMMON = {
    "RESNET": {"ZERO_INIT_FINAL_BN": True, "WIDTH_PER_GROUP": 64, "NUM_GROUPS": 1, "STRIDE_1X1": False,
               "NUM_BLOCK_TEMP_KERNEL": [[3, 3], [4, 4], [6, 6], [3, 3]],
               "SPATIAL_STRIDES": [[1, 1], [2, 2], [2, 2], [2, 2]],
               "SPATIAL_DILATIONS": [[1, 1], [1, 1], [1, 1], [1, 1]]},
    "NONLOCAL": {"LOCATION": [[[], []], [[], []], [[], []], [[], []]],
                 "GROUP": [[1, 1], [1, 1], [1, 1], [1, 1]], "INSTANTIATION": "dot_product"},
    "MODEL": {"ARCH": "slowfast", "MODEL_NAME": "SlowFast"},
}


def _single(arch: str, **nonlocal_kw) -> Dict[str, Any]:
    d = copy.deepcopy(_R50_SINGLE)
    d["MODEL"] = {"ARCH": arch, "MODEL_NAME": "ResNet"}
    d["NONLOCAL"] = {"LOCATION": [[[]], [[]], [[]], [[]]], "GROUP": [[1], [1], [1], [1]],
                     "INSTANTIATION": "softmax"}
    d["NONLOCAL"].update(nonlocal_kw)
    return d


def _slowfast(depth: int, frames: int, fusion_k: int) -> Dict[str, Any]:
    d = copy.deepcopy(_SLOWFAST_COMMON)
    d["RESNET"]["DEPTH"] = depth
    d["DATA"] = {"NUM_FRAMES": frames, "SAMPLING_RATE": 2, "INPUT_CHANNEL_NUM": [3, 3]}
    d["SLOWFAST"] = {"ALPHA": 4, "BETA_INV": 8, "FUSION_CONV_CHANNEL_RATIO": 2, "FUSION_KERNEL_SZ": fusion_k}
    return d


# `mdl.sf_mdl_name` -> overrides (configs/vsitu_mdl_cfgs/*.yaml restated)
SF_MDL_PRESETS: Dict[str, Dict[str, Any]] = {
    "slow_fast_nl_r50_8x8": _slowfast(50, 32, 7),          # Kinetics_c2_SLOWFAST_8x8_R50.yaml:11-32
    "slow_nl_r50_8x8": _single("slow", INSTANTIATION="dot_product"),   # Kinetics_c2_SLOW_8x8_R50.yaml
    "c2d_r50_8x8": _single("c2d"),                         # Kinetics_C2D_8x8_R50.yaml
    "i3d_r50_8x8": _single("i3d"),                         # Kinetics_c2_I3D_8x8_R50.yaml
    "i3d_r50_nl_8x8": _single("i3d", LOCATION=[[[]], [[1, 3]], [[1, 3, 5]], [[]]]),  # ..._I3D_NLN_8x8_R50.yaml:25-28
    "slow_fast_r101_16x8": _slowfast(101, 64, 5),          # SlowFast/configs/Kinetics/c2/SLOWFAST_16x8_R101_50_50.yaml
}


def sf_mdl_cfg(sf_mdl_name: str, **overrides) -> AttrDict:
    """`cfg.sf_mdl` for one of the reference's backbone names."""
    if sf_mdl_name not in SF_MDL_PRESETS:
        raise KeyError(f"unknown sf_mdl_name {sf_mdl_name!r}; known: {sorted(SF_MDL_PRESETS)}")
    cfg = AttrDict(copy.deepcopy(_DEFAULTS))
    cfg.merge(SF_MDL_PRESETS[sf_mdl_name])
    cfg.merge(overrides)
    return cfg


def make_cfg(sf_mdl_name: str = "slow_fast_nl_r50_8x8", **sf_overrides) -> AttrDict:
    """Minimal top-level cfg with the keys SFBase reads (`cfg.sf_mdl`, `cfg.mdl`)."""
    return AttrDict({
        "sf_mdl": sf_mdl_cfg(sf_mdl_name, **sf_overrides),
        "mdl": {"sf_mdl_name": sf_mdl_name, "mdl_name": "sf_base"},
        "task_type": "vb",
    })


def make_comm(sf_cfg, num_verbs: int = 1560) -> AttrDict:
    """The `comm` fields SFBase reads (dat_loader.py:69-79,133-138): path_type and a
    verb vocabulary whose only use on this path is len()."""
    arch = sf_cfg.MODEL.ARCH
    if arch in sf_cfg.MODEL.MULTI_PATHWAY_ARCH:
        path_type = "multi"
    elif arch in sf_cfg.MODEL.SINGLE_PATHWAY_ARCH:
        path_type = "single"
    else:
        raise NotImplementedError(arch)
    return AttrDict({"path_type": path_type, "vb_id_vocab": list(range(num_verbs))})
