"""ORACLE -- test infrastructure only.  Never imported by the product (vidsitu_b200/).

CPU fp32 restatement of the reference's event-clip forward, used as the checker in
tests/, in __graft_entry__.smoke() and as bench.py's `cpu_baseline` / `--impl reference`
arm.  It is a floating-point path, so the restatement is functional PyTorch fp32
(torch.nn.functional on the CPU) driven directly by a reference-layout state_dict; the
network structure is read from the state_dict keys and weight shapes themselves, not
from vidsitu_b200's spec builder, so the two cannot share a structural bug.

Parity status: the reference holds NO test, golden vector or fixture for this path
 -- "parity unpinned" by the reference's own tests.  It is pinned
instead against outputs of the reference itself: tests/golden/make_golden.py imports the
unmodified `vidsitu_code.mdl_sf_base.SFBase` from /root/reference (behind stub modules
for its missing third-party imports), runs it on seeded inputs/weights and commits the
outputs under tests/golden/; tests/test_oracle.py checks this file against them.

Reference functions restated (file:line under /root/reference):
  get_sequence            utils/video_utils.py:18-38
  tensor_normalize        utils/video_utils.py:147-164
  pack_pathway_output     utils/video_utils.py:41-74
  SFBase.get_feats        vidsitu_code/mdl_sf_base.py:169-180
  forward_features        vidsitu_code/mdl_sf_base.py:21-34 (SlowFast), 46-55 (ResNet)
  VideoModelStem          SlowFast/slowfast/models/stem_helper.py:92-99, 173-178
  FuseFastToSlow.forward  SlowFast/slowfast/models/video_model_builder.py:124-131
  ResStage.forward        SlowFast/slowfast/models/resnet_helper.py:530-561
  ResBlock.forward        resnet_helper.py:352-358
  BottleneckTransform     resnet_helper.py:225-240 (stride on the 3x3 unless STRIDE_1X1)
  Nonlocal.forward        SlowFast/slowfast/models/nonlocal_helper.py:105-148
  ResNetBasicHead_Trimmed vidsitu_code/mdl_sf_base.py:103-113
  forward_decoder         vidsitu_code/mdl_sf_base.py:189-211
  EvalB top-5             vidsitu_code/evl_vsitu.py:41-42
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm3d constructor default; cfg.BN.EPSILON is never passed

POOL1 = {"c2d": [2, 1, 1], "i3d": [2, 1, 1], "slow": [1, 1, 1], "slowfast": [1, 1, 1],
         "c2d_nopool": [1, 1, 1], "i3d_nopool": [1, 1, 1]}


# ----------------------------------------------------------------------------- data side
def get_sequence(center_idx: int, half_len: int, sample_rate: int, max_num_frames: int) -> List[int]:
    seq = list(range(center_idx - half_len, center_idx + half_len, sample_rate))
    return [min(max(i, 0), max_num_frames - 1) for i in seq]


def event_frame_indices(num_frames: int, sampling_rate: int, fps: int = 30, max_frames: int = 300):
    """Frame indices of the five 2-second events (dat_loader.py:69-79, 454-472)."""
    half = num_frames * sampling_rate // 2
    return [get_sequence(int((ev + 0.5) * fps * 2), half, sampling_rate, max_frames) for ev in range(5)]


def tensor_normalize(frames_u8: torch.Tensor, mean: Sequence[float], std: Sequence[float]) -> torch.Tensor:
    x = frames_u8.float()
    x = x / 255.0
    x = x - torch.tensor(mean)
    x = x / torch.tensor(std)
    return x


def pack_pathways(frames_cthw: torch.Tensor, multi: bool, alpha: int, reverse: bool = False) -> List[torch.Tensor]:
    if reverse:
        frames_cthw = frames_cthw[[2, 1, 0]]
    if not multi:
        return [frames_cthw]
    t = frames_cthw.shape[1]
    idx = torch.linspace(0, t - 1, t // alpha).long()
    return [torch.index_select(frames_cthw, 1, idx), frames_cthw]


def clips_from_frames(frames_u8: torch.Tensor, cfg) -> List[torch.Tensor]:
    """uint8 [N, T, H, W, 3] -> list of per-pathway fp32 [N, 3, T_p, H, W] (what get_feats yields)."""
    multi = cfg.MODEL.ARCH in cfg.MODEL.MULTI_PATHWAY_ARCH
    alpha = cfg.SLOWFAST.ALPHA if multi else 1
    per_clip = []
    for f in frames_u8:
        x = tensor_normalize(f, cfg.DATA.MEAN, cfg.DATA.STD).permute(3, 0, 1, 2)
        per_clip.append(pack_pathways(x, multi, alpha, bool(cfg.DATA.REVERSE_INPUT_CHANNEL)))
    return [torch.stack([c[p] for c in per_clip]).float() for p in range(len(per_clip[0]))]


# --------------------------------------------------------------------------- model side
def _bn(sd: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor) -> torch.Tensor:
    if prefix + ".bn.running_mean" in sd:
        # SubBatchNorm3d in eval mode (batchnorm_helper.py:97-109): affine-less BN on the aggregated statistics,
        # then the single weight / bias pair
        x = F.batch_norm(x, sd[prefix + ".bn.running_mean"], sd[prefix + ".bn.running_var"], None, None,
                         training=False, eps=BN_EPS)
        return x * sd[prefix + ".weight"].view(-1, 1, 1, 1) + sd[prefix + ".bias"].view(-1, 1, 1, 1)
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], training=False, eps=BN_EPS)


def _conv(sd, key: str, x, stride, pad):
    return F.conv3d(x, sd[key + ".weight"], sd.get(key + ".bias"), stride=stride, padding=pad)


def _stem(sd, prefix: str, x):
    kt = sd[prefix + ".conv.weight"].shape[2]
    x = _conv(sd, prefix + ".conv", x, (1, 2, 2), (kt // 2, 3, 3))
    x = F.relu(_bn(sd, prefix + ".bn", x))
    return F.max_pool3d(x, kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1))


def _fuse(sd, prefix: str, x_s, x_f, alpha: int):
    k = sd[prefix + ".conv_f2s.weight"].shape[2]
    f = _conv(sd, prefix + ".conv_f2s", x_f, (alpha, 1, 1), (k // 2, 0, 0))
    f = F.relu(_bn(sd, prefix + ".bn", f))
    return torch.cat([x_s, f], 1)


def _block(sd, prefix: str, x, stride: int, stride_1x1: bool, dilation: int):
    s1, s3 = (stride, 1) if stride_1x1 else (1, stride)
    kt = sd[prefix + ".branch2.a.weight"].shape[2]
    y = _conv(sd, prefix + ".branch2.a", x, (1, s1, s1), (kt // 2, 0, 0))
    y = F.relu(_bn(sd, prefix + ".branch2.a_bn", y))
    w = sd[prefix + ".branch2.b.weight"]
    y = F.conv3d(y, w, None, stride=(1, s3, s3), padding=(0, dilation, dilation), dilation=(1, dilation, dilation))
    y = F.relu(_bn(sd, prefix + ".branch2.b_bn", y))
    y = _bn(sd, prefix + ".branch2.c_bn", _conv(sd, prefix + ".branch2.c", y, 1, 0))
    if prefix + ".branch1.weight" in sd:
        x = _bn(sd, prefix + ".branch1_bn", _conv(sd, prefix + ".branch1", x, (1, stride, stride), 0))
    return F.relu(x + y)


def _nonlocal(sd, prefix: str, x, pool, instantiation: str):
    n, c, t, h, w = x.shape
    inner = sd[prefix + ".conv_theta.weight"].shape[0]
    theta = _conv(sd, prefix + ".conv_theta", x, 1, 0)
    xp = F.max_pool3d(x, kernel_size=pool, stride=pool) if pool is not None and any(s > 1 for s in pool) else x
    phi = _conv(sd, prefix + ".conv_phi", xp, 1, 0)
    g = _conv(sd, prefix + ".conv_g", xp, 1, 0)
    theta, phi, g = theta.view(n, inner, -1), phi.view(n, inner, -1), g.view(n, inner, -1)
    tp = torch.einsum("nct,ncp->ntp", theta, phi)
    if instantiation == "softmax":
        tp = F.softmax(tp * (inner ** -0.5), dim=2)
    elif instantiation == "dot_product":
        tp = tp / tp.shape[2]
    else:
        raise NotImplementedError(instantiation)
    y = torch.einsum("ntg,ncg->nct", tp, g).view(n, inner, t, h, w)
    return x + _bn(sd, prefix + ".bn", _conv(sd, prefix + ".conv_out", y, 1, 0))


def _stage(sd, name: str, xs: List[torch.Tensor], cfg, si: int) -> List[torch.Tensor]:
    out = []
    stride_1x1 = bool(cfg.RESNET.STRIDE_1X1) if cfg.MODEL.MODEL_NAME == "ResNet" else False
    for p, x in enumerate(xs):
        i = 0
        while f"{name}.pathway{p}_res{i}.branch2.a.weight" in sd:
            x = _block(sd, f"{name}.pathway{p}_res{i}", x, cfg.RESNET.SPATIAL_STRIDES[si][p] if i == 0 else 1,
                       stride_1x1, cfg.RESNET.SPATIAL_DILATIONS[si][p])
            nl = f"{name}.pathway{p}_nonlocal{i}"
            if nl + ".conv_theta.weight" in sd:
                if cfg.NONLOCAL.GROUP[si][p] != 1:
                    raise NotImplementedError("NONLOCAL.GROUP > 1")
                x = _nonlocal(sd, nl, x, cfg.NONLOCAL.POOL[si][p], cfg.NONLOCAL.INSTANTIATION)
            i += 1
        out.append(x)
    return out


@torch.no_grad()
def forward_features(sd: Dict[str, torch.Tensor], cfg, xs: List[torch.Tensor]) -> List[torch.Tensor]:
    """sd: state_dict of `SFBase.sf_mdl` (keys without the 'sf_mdl.' prefix)."""
    multi = len(xs) == 2
    alpha = cfg.SLOWFAST.ALPHA if multi else 1
    xs = [_stem(sd, f"s1.pathway{p}_stem", x) for p, x in enumerate(xs)]
    for si in range(4):
        if multi:
            xs = [_fuse(sd, f"s{si + 1}_fuse", xs[0], xs[1], alpha), xs[1]]
        xs = _stage(sd, f"s{si + 2}", xs, cfg, si)
        if si == 0:
            k = POOL1[cfg.MODEL.ARCH]
            if any(v != 1 for v in k):
                xs = [F.max_pool3d(x, kernel_size=k, stride=k) for x in xs]
    return xs


def head(feats: List[torch.Tensor]) -> torch.Tensor:
    return torch.cat([F.adaptive_avg_pool3d(f, (1, 1, 1)) for f in feats], 1)


@torch.no_grad()
def sfbase_forward(full_sd: Dict[str, torch.Tensor], cfg, xs: List[torch.Tensor]):
    """full_sd: state_dict of the whole SFBase.  Returns (feature maps, pooled [N,D], logits [N,V])."""
    sd = {k[len("sf_mdl."):]: v for k, v in full_sd.items() if k.startswith("sf_mdl.")}
    fmaps = forward_features(sd, cfg, xs)
    pooled = head(fmaps).flatten(1)
    h = F.relu(F.linear(pooled, full_sd["proj_head.0.weight"], full_sd["proj_head.0.bias"]))
    logits = F.linear(h, full_sd["proj_head.2.weight"], full_sd["proj_head.2.bias"])
    return fmaps, pooled, logits


def top5(logits: torch.Tensor) -> torch.Tensor:
    """EvalB.forward_one_batch: softmax, full descending sort, first five ids."""
    probs = F.softmax(logits, dim=-1)
    return probs.sort(dim=-1, descending=True)[1][..., :5]


def topk_probs(logits: torch.Tensor, k: int = 5):
    """EvalB.forward_one_batch (evl_vsitu.py:41-47, 57-60): (ids, scores) of the k most probable verbs."""
    probs = F.softmax(logits, dim=-1)
    s, i = probs.sort(dim=-1, descending=True, stable=True)
    return i[..., :k], s[..., :k]
