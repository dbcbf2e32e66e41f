"""Checkpoint ingestion for the hot path.

Two formats reach `feat_extractor.main` (vidsitu_code/feat_extractor.py:147-161):

* a VidSitu `.pth` written by the trainer: `{"model_state_dict": {"module.<key>": tensor}}`
  (DataParallel prefix stripped by `rem_mdl`, feat_extractor.py:115-117) -> `load_vidsitu_checkpoint`;
* a PySlowFast Caffe2 `.pkl` (`{"blobs": {c2_name: ndarray}}`, latin-1 pickle) loaded into
  `mdl.sf_mdl` with `convert_from_caffe2=True`
  (SlowFast/slowfast/utils/checkpoint.py:203-259, name table in c2_model_loading.py:9-120)
  -> `load_caffe2_checkpoint`.

The Caffe2 blob grammar is parsed structurally here (pathway prefix, stage/block, branch, tensor
suffix) instead of through the reference's ordered regex table; tests/test_host.py checks the two
agree on every blob name of every supported backbone (tests/golden/c2_name_pairs.json, generated
by running the reference's own converter).
"""
from __future__ import annotations

import pickle
import re
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

_BN_LEAF = {"s": "weight", "b": "bias", "rm": "running_mean", "riv": "running_var"}
_LEAF = {"w": "weight", "b": "bias"}

_RE_NONLOCAL = re.compile(r"^nonlocal_conv(\d+)_(\d+)_(theta|phi|g|out|bn)_(\w+)$")
_RE_FUSE_POOL = re.compile(r"^t_pool1_subsample_(bn_)?(\w+)$")
_RE_FUSE_RES = re.compile(r"^t_res(\d+)_(\d+)_branch2c_bn_subsample_(bn_)?(\w+)$")
_RE_BLOCK = re.compile(r"^(t_)?res(\d+)_(\d+)_branch(1|2[abc])_(bn_)?(\w+)$")
_RE_STEM_BN = re.compile(r"^(t_)?res_conv1_bn_(\w+)$")
_RE_STEM = re.compile(r"^(t_)?(?:res_)?conv1_(\w+)$")
_RE_PRED = re.compile(r"^pred_(\w+)$")


def convert_caffe2_name(name: str) -> Optional[str]:
    """Caffe2 blob name -> key of `SFBase.sf_mdl.state_dict()`; None for blobs that are not model
    tensors (solver state: `*_momentum`, `lr`, `model_iter`, ...)."""

    def leaf(tok: str, bn: bool) -> Optional[str]:
        return (_BN_LEAF if bn else _LEAF).get(tok)

    m = _RE_NONLOCAL.match(name)
    if m:
        stage, idx, part, tok = m.groups()
        lf = leaf(tok, part == "bn")
        if lf is None:
            return None
        sub = "bn" if part == "bn" else f"conv_{part}"
        return f"s{stage}.pathway0_nonlocal{idx}.{sub}.{lf}"
    m = _RE_FUSE_POOL.match(name)
    if m:
        bn, tok = m.groups()
        lf = leaf(tok, bool(bn))
        return None if lf is None else f"s1_fuse.{'bn' if bn else 'conv_f2s'}.{lf}"
    m = _RE_FUSE_RES.match(name)
    if m:
        stage, _, bn, tok = m.groups()
        lf = leaf(tok, bool(bn))
        return None if lf is None else f"s{stage}_fuse.{'bn' if bn else 'conv_f2s'}.{lf}"
    m = _RE_BLOCK.match(name)
    if m:
        fast, stage, blk, branch, bn, tok = m.groups()
        lf = leaf(tok, bool(bn))
        if lf is None:
            return None
        p = 1 if fast else 0
        mod = "branch1" if branch == "1" else f"branch2.{branch[1]}"
        return f"s{stage}.pathway{p}_res{blk}.{mod}{'_bn' if bn else ''}.{lf}"
    m = _RE_STEM_BN.match(name)
    if m:
        fast, tok = m.groups()
        lf = leaf(tok, True)
        return None if lf is None else f"s1.pathway{1 if fast else 0}_stem.bn.{lf}"
    m = _RE_STEM.match(name)
    if m:
        fast, tok = m.groups()
        lf = leaf(tok, False)
        return None if lf is None else f"s1.pathway{1 if fast else 0}_stem.conv.{lf}"
    m = _RE_PRED.match(name)
    if m:
        lf = leaf(m.group(1), False)
        return None if lf is None else f"head.projection.{lf}"
    return None


def convert_caffe2_blobs(blobs: Dict[str, np.ndarray], target: Dict[str, torch.Tensor]
                         ) -> Tuple["OrderedDict[str, torch.Tensor]", List[str], List[str]]:
    """The conversion loop of checkpoint.py:210-257: a blob is taken when its converted name exists in `target`
    with the same shape.  Sub-BN models (BN.NORM_TYPE sub_batchnorm): running statistics go to
    `<bn>.split_bn.running_*` (c2_normal_to_sub_bn, checkpoint.py:331-348) tiled NUM_SPLITS times
    (checkpoint.py:215-232).  Returns (state_dict, mismatched, skipped)."""
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    mismatched, skipped = [], []
    for c2_name, blob in blobs.items():
        key = convert_caffe2_name(c2_name)
        if key is not None and key not in target and "bn.running_" in key:
            sub = key.replace("bn.running_", "bn.split_bn.running_")
            if sub in target:
                key = sub
        if key is None or key not in target:
            skipped.append(c2_name)
            continue
        arr = np.asarray(blob)
        tshape = tuple(target[key].shape)
        if len(tshape) == 1 and arr.ndim == 1 and tshape[0] > arr.shape[0] and tshape[0] % arr.shape[0] == 0:
            arr = np.concatenate([arr] * (tshape[0] // arr.shape[0]))
        if tuple(arr.shape) == tuple(target[key].shape):
            out[key] = torch.tensor(arr).clone()
        else:
            mismatched.append(c2_name)
    return out, mismatched, skipped


def load_caffe2_checkpoint(path, sf_mdl) -> Dict[str, List[str]]:
    """`load_checkpoint(path, model=mdl.sf_mdl, data_parallel=False, convert_from_caffe2=True)`
    (feat_extractor.py:156-161).  Loads non-strictly, like the reference; reports what was not used."""
    with open(path, "rb") as f:
        ckpt = pickle.load(f, encoding="latin1")
    target = sf_mdl.state_dict()
    sd, mismatched, skipped = convert_caffe2_blobs(ckpt["blobs"], target)
    res = sf_mdl.load_state_dict(sd, strict=False)
    if any(".split_bn." in k for k in sd):
        # eval-mode SubBatchNorm3d (and our BN folding) reads the aggregated `<bn>.bn.running_*`
        # (batchnorm_helper.py:78-95): derive them from the split statistics just loaded
        from .model import aggregate_sub_bn_stats
        aggregate_sub_bn_stats(sf_mdl)
    stat_blobs_skipped = [n for n in skipped if n.endswith(("_rm", "_riv"))]
    if stat_blobs_skipped and any(".split_bn." in k for k in target):
        raise ValueError(f"BatchNorm statistics blobs were not loaded into the sub-BN model: {stat_blobs_skipped[:4]} ...")
    owner = getattr(sf_mdl, "_owner", None)
    if owner is not None and owner() is not None:
        owner().invalidate_engines()   # kernels read prepared copies of the weights
    return {"loaded": list(sd.keys()), "mismatched": mismatched, "skipped": skipped,
            "missing": list(res.missing_keys)}


def strip_module_prefix(key: str) -> str:
    """`rem_mdl` (feat_extractor.py:115-117): text after the first 'module.'."""
    return key.split("module.", 1)[1]


def load_vidsitu_checkpoint(path, model, map_location="cpu") -> None:
    """`mdl.load_state_dict({rem_mdl(k): v for k, v in torch.load(p)["model_state_dict"].items()})`
    (feat_extractor.py:148-154): strict, DataParallel prefix removed."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    sd = ckpt["model_state_dict"]
    model.load_state_dict({strip_module_prefix(k): v for k, v in sd.items()})
