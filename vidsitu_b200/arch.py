"""Network description of the SlowFast / ResNet(C2D, I3D, Slow) backbones on the hot path.

`build_spec(cfg.sf_mdl)` turns the reference's config keys into a flat, explicit
description (every conv with its state_dict names, shapes, strides and what follows
it).  It restates WHAT the reference builds --
  SlowFast._construct_network   SlowFast/slowfast/models/video_model_builder.py:161-377
  ResNet._construct_network     video_model_builder.py:432-557
  VideoModelStem/ResNetBasicStem stem_helper.py:9-178
  ResStage / ResBlock / BottleneckTransform  resnet_helper.py:110-561
  FuseFastToSlow                video_model_builder.py:74-131
  Nonlocal                      nonlocal_helper.py:10-148
-- as data; the same spec drives the parameter container (state_dict key set,
SURVEY.md section 8 a20) and the kernel plan (vidsitu_b200/engine.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

Triple = Tuple[int, int, int]

STAGE_DEPTH = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3)}        # video_model_builder.py:16

# temporal kernel of [stem, res2, res3, res4, res5] per pathway        video_model_builder.py:19-61
TEMPORAL_KERNELS = {
    "c2d": [[[1]], [[1]], [[1]], [[1]], [[1]]],
    "c2d_nopool": [[[1]], [[1]], [[1]], [[1]], [[1]]],
    "i3d": [[[5]], [[3]], [[3, 1]], [[3, 1]], [[1, 3]]],
    "i3d_nopool": [[[5]], [[3]], [[3, 1]], [[3, 1]], [[1, 3]]],
    "slow": [[[1]], [[1]], [[1]], [[3]], [[3]]],
    "slowfast": [[[1], [5]], [[1], [3]], [[1], [3]], [[3], [3]], [[3], [3]]],
}

# pathway{p}_pool kernel (= stride) after res2                         video_model_builder.py:64-71
POOL1 = {
    "c2d": [[2, 1, 1]], "c2d_nopool": [[1, 1, 1]], "i3d": [[2, 1, 1]], "i3d_nopool": [[1, 1, 1]],
    "slow": [[1, 1, 1]], "slowfast": [[1, 1, 1], [1, 1, 1]],
}


@dataclass
class ConvSpec:
    key: str                    # state_dict prefix: weight at f"{key}.weight" ([cout,cin,kt,kh,kw])
    cin: int
    cout: int
    kernel: Triple
    stride: Triple
    pad: Triple
    bn: Optional[str] = None    # state_dict prefix of the BatchNorm3d that follows
    relu: bool = False
    has_bias: bool = False      # conv bias (Nonlocal convs only)
    final_bn: bool = False      # gamma zero-initialised when RESNET.ZERO_INIT_FINAL_BN

    @property
    def flops_per_out_pixel(self) -> int:
        kt, kh, kw = self.kernel
        return 2 * self.cin * self.cout * kt * kh * kw


@dataclass
class NonlocalSpec:
    prefix: str
    dim: int
    dim_inner: int
    theta: ConvSpec
    phi: ConvSpec
    g: ConvSpec
    out: ConvSpec
    pool: Optional[Triple]
    softmax: bool


@dataclass
class BlockSpec:
    prefix: str
    branch1: Optional[ConvSpec]
    a: ConvSpec
    b: ConvSpec
    c: ConvSpec
    nonlocal_: Optional[NonlocalSpec] = None


@dataclass
class StemSpec:
    conv: ConvSpec
    pool_kernel: Triple = (1, 3, 3)
    pool_stride: Triple = (1, 2, 2)
    pool_pad: Triple = (0, 1, 1)


@dataclass
class NetSpec:
    model_name: str
    arch: str
    num_pathways: int
    num_frames: int             # DATA.NUM_FRAMES (frames of the fast / single pathway)
    alpha: int
    stems: List[StemSpec]
    fuses: List[Optional[ConvSpec]]            # after s1..s4 (None for single-pathway nets)
    stages: List[List[List[BlockSpec]]]        # [s2..s5][pathway][block]
    pool1: List[Triple]
    feat_dims: List[int]
    num_classes: int            # of the backbone's own (unused) head.projection
    zero_init_final_bn: bool
    fc_init_std: float
    crop: int = 224
    mean: Tuple[float, ...] = (0.45, 0.45, 0.45)
    std: Tuple[float, ...] = (0.225, 0.225, 0.225)
    reverse_input_channel: bool = False
    norm_type: str = "batchnorm"   # BN.NORM_TYPE
    num_splits: int = 1            # BN.NUM_SPLITS (sub_batchnorm)
    sampling_rate: int = 2      # DATA.SAMPLING_RATE: frame step inside an event window
    target_fps: int = 30        # DATA.TARGET_FPS of the extracted frames

    def pathway_frames(self) -> List[int]:
        if self.num_pathways == 2:
            return [self.num_frames // self.alpha, self.num_frames]
        return [self.num_frames]

    def all_convs(self) -> List[ConvSpec]:
        out: List[ConvSpec] = [s.conv for s in self.stems]
        out += [f for f in self.fuses if f is not None]
        for stage in self.stages:
            for blocks in stage:
                for b in blocks:
                    out += [c for c in (b.branch1, b.a, b.b, b.c) if c is not None]
                    if b.nonlocal_ is not None:
                        nl = b.nonlocal_
                        out += [nl.theta, nl.phi, nl.g, nl.out]
        return out


def _bottleneck(prefix: str, dim_in: int, dim_out: int, dim_inner: int, temp_k: int, stride: int,
                stride_1x1: bool, dilation: int) -> BlockSpec:
    """resnet_helper.py:243-358 (ResBlock) + :110-240 (BottleneckTransform)."""
    if dilation != 1:
        # resnet_helper.py:196-207 dilates the 3x3 (padding = dilation); ConvSpec / the kernels have no dilation, so a
        # config with RESNET.SPATIAL_DILATIONS > 1 must fail loudly instead of computing a different conv
        raise NotImplementedError(f"{prefix}: RESNET.SPATIAL_DILATIONS = {dilation} (only 1 is implemented)")
    str1, str3 = (stride, 1) if stride_1x1 else (1, stride)
    branch1 = None
    if dim_in != dim_out or stride != 1:
        branch1 = ConvSpec(f"{prefix}.branch1", dim_in, dim_out, (1, 1, 1), (1, stride, stride), (0, 0, 0),
                           bn=f"{prefix}.branch1_bn")
    a = ConvSpec(f"{prefix}.branch2.a", dim_in, dim_inner, (temp_k, 1, 1), (1, str1, str1), (temp_k // 2, 0, 0),
                 bn=f"{prefix}.branch2.a_bn", relu=True)
    b = ConvSpec(f"{prefix}.branch2.b", dim_inner, dim_inner, (1, 3, 3), (1, str3, str3), (0, dilation, dilation),
                 bn=f"{prefix}.branch2.b_bn", relu=True)
    c = ConvSpec(f"{prefix}.branch2.c", dim_inner, dim_out, (1, 1, 1), (1, 1, 1), (0, 0, 0),
                 bn=f"{prefix}.branch2.c_bn", relu=False, final_bn=True)
    return BlockSpec(prefix, branch1, a, b, c)


def _nonlocal(prefix: str, dim: int, pool, instantiation: str) -> NonlocalSpec:
    """nonlocal_helper.py:60-103: dim_inner = dim // 2 (resnet_helper.py:520-524)."""
    inner = dim // 2
    one, z = (1, 1, 1), (0, 0, 0)
    use_pool = pool is not None and any(s > 1 for s in pool)
    if instantiation not in ("softmax", "dot_product"):
        raise NotImplementedError(f"Unknown norm type {instantiation}")
    return NonlocalSpec(
        prefix, dim, inner,
        theta=ConvSpec(f"{prefix}.conv_theta", dim, inner, one, one, z, has_bias=True),
        phi=ConvSpec(f"{prefix}.conv_phi", dim, inner, one, one, z, has_bias=True),
        g=ConvSpec(f"{prefix}.conv_g", dim, inner, one, one, z, has_bias=True),
        out=ConvSpec(f"{prefix}.conv_out", inner, dim, one, one, z, bn=f"{prefix}.bn", has_bias=True,
                     final_bn=True),
        pool=tuple(pool) if use_pool else None,
        softmax=instantiation == "softmax",
    )


def _stage(name: str, dim_in, dim_out, dim_inner, temp_kernels, strides, num_blocks, num_block_temp_kernel,
           nonlocal_inds, nonlocal_pool, instantiation, stride_1x1, dilations, trans_func, groups,
           nonlocal_group) -> List[List[BlockSpec]]:
    """resnet_helper.py:361-528 (ResStage)."""
    if trans_func != "bottleneck_transform":
        raise NotImplementedError(f"RESNET.TRANS_FUNC={trans_func!r} (only bottleneck_transform is on the VidSitu path)")
    pathways = []
    for p in range(len(num_blocks)):
        if groups[p] != 1:
            raise NotImplementedError("grouped 3x3 convolutions (ResNeXt) are outside the VidSitu configs")
        if nonlocal_inds[p] and nonlocal_group[p] != 1:
            raise NotImplementedError("NONLOCAL.GROUP > 1 is outside the VidSitu configs")
        assert num_block_temp_kernel[p] <= num_blocks[p]
        tks = (list(temp_kernels[p]) * num_blocks[p])[: num_block_temp_kernel[p]]
        tks += [1] * (num_blocks[p] - num_block_temp_kernel[p])
        blocks = []
        for i in range(num_blocks[p]):
            prefix = f"{name}.pathway{p}_res{i}"
            blk = _bottleneck(prefix, dim_in[p] if i == 0 else dim_out[p], dim_out[p], dim_inner[p], tks[i],
                              strides[p] if i == 0 else 1, stride_1x1, dilations[p])
            if i in nonlocal_inds[p]:
                blk.nonlocal_ = _nonlocal(f"{name}.pathway{p}_nonlocal{i}", dim_out[p], nonlocal_pool[p],
                                          instantiation)
            blocks.append(blk)
        pathways.append(blocks)
    return pathways


def build_spec(cfg) -> NetSpec:
    """cfg = `cfg.sf_mdl` (yacs CfgNode or vidsitu_b200.config.AttrDict)."""
    # get_norm (batchnorm_helper.py:15-34): all three are frozen per-channel affines in eval mode; sub_batchnorm
    # keeps its statistics under `<bn>.bn.*` / `<bn>.split_bn.*` (SubBatchNorm3d, batchnorm_helper.py:37-109)
    if cfg.BN.NORM_TYPE not in ("batchnorm", "sub_batchnorm", "sync_batchnorm"):
        raise NotImplementedError(f"Norm type {cfg.BN.NORM_TYPE} is not supported")
    if cfg.DETECTION.ENABLE:
        raise NotImplementedError("DETECTION.ENABLE is not on the VidSitu path")
    name = cfg.MODEL.MODEL_NAME
    arch = cfg.MODEL.ARCH
    if arch not in POOL1 or cfg.RESNET.DEPTH not in STAGE_DEPTH:
        raise NotImplementedError(f"arch {arch!r} / depth {cfg.RESNET.DEPTH}")
    if name == "SlowFast":
        npw = 2
    elif name == "ResNet":
        npw = 1
    else:
        raise NotImplementedError(name)
    if len(POOL1[arch]) != npw:
        raise ValueError(f"MODEL.ARCH={arch!r} does not have {npw} pathway(s)")
    depths = STAGE_DEPTH[cfg.RESNET.DEPTH]
    w = cfg.RESNET.WIDTH_PER_GROUP
    groups = cfg.RESNET.NUM_GROUPS
    dim_inner = groups * w
    tk = TEMPORAL_KERNELS[arch]
    beta_inv = cfg.SLOWFAST.BETA_INV if npw == 2 else 1
    ratio = cfg.SLOWFAST.FUSION_CONV_CHANNEL_RATIO if npw == 2 else 0
    alpha = cfg.SLOWFAST.ALPHA if npw == 2 else 1
    fusion_k = cfg.SLOWFAST.FUSION_KERNEL_SZ if npw == 2 else 0
    cin = list(cfg.DATA.INPUT_CHANNEL_NUM)
    if len(cin) != npw:
        raise ValueError("DATA.INPUT_CHANNEL_NUM does not match the number of pathways")

    def widths(mult: int) -> List[int]:          # per-pathway channel count for a slow width w*mult
        return [w * mult] if npw == 1 else [w * mult, w * mult // beta_inv]

    # --- stems (video_model_builder.py:185-195, 453-460; stem_helper.py:102-178)
    stem_prefix = ["s1.pathway0_stem", "s1.pathway1_stem"]
    stems = []
    for p in range(npw):
        kt = tk[0][p][0]
        stems.append(StemSpec(ConvSpec(f"{stem_prefix[p]}.conv", cin[p], widths(1)[p], (kt, 7, 7), (1, 2, 2),
                                       (kt // 2, 3, 3), bn=f"{stem_prefix[p]}.bn", relu=True)))

    # --- lateral connections (video_model_builder.py:196-202 etc.)
    def fuse(stage_name: str, fast_dim: int) -> Optional[ConvSpec]:
        if npw == 1:
            return None
        return ConvSpec(f"{stage_name}_fuse.conv_f2s", fast_dim, fast_dim * ratio, (fusion_k, 1, 1), (alpha, 1, 1),
                        (fusion_k // 2, 0, 0), bn=f"{stage_name}_fuse.bn", relu=True)

    fuses = [fuse("s1", widths(1)[-1]), fuse("s2", widths(4)[-1]), fuse("s3", widths(8)[-1]),
             fuse("s4", widths(16)[-1])]

    # --- res stages (video_model_builder.py:204-351, 462-552).  SlowFast's ResStage calls never pass
    # stride_1x1 (default False, resnet_helper.py:376); ResNet passes RESNET.STRIDE_1X1 (:476).
    stride_1x1 = bool(cfg.RESNET.STRIDE_1X1) if name == "ResNet" else False
    stages = []
    in_mult = [1, 4, 8, 16]
    out_mult = [4, 8, 16, 32]
    for si in range(4):
        d_in = widths(in_mult[si])
        if npw == 2:
            d_in = [d_in[0] + d_in[1] * ratio, d_in[1]]     # slow input = slow + fused channels
        inner = [dim_inner * (2 ** si)] if npw == 1 else [dim_inner * (2 ** si), dim_inner * (2 ** si) // beta_inv]
        stages.append(_stage(
            f"s{si + 2}", d_in, widths(out_mult[si]), inner, tk[si + 1], cfg.RESNET.SPATIAL_STRIDES[si],
            [depths[si]] * npw, cfg.RESNET.NUM_BLOCK_TEMP_KERNEL[si], cfg.NONLOCAL.LOCATION[si],
            cfg.NONLOCAL.POOL[si], cfg.NONLOCAL.INSTANTIATION, stride_1x1,
            cfg.RESNET.SPATIAL_DILATIONS[si], cfg.RESNET.TRANS_FUNC, [groups] * npw, cfg.NONLOCAL.GROUP[si]))

    return NetSpec(
        model_name=name, arch=arch, num_pathways=npw, num_frames=cfg.DATA.NUM_FRAMES, alpha=alpha,
        stems=stems, fuses=fuses, stages=stages, pool1=[tuple(p) for p in POOL1[arch]], feat_dims=widths(32),
        num_classes=cfg.MODEL.NUM_CLASSES, zero_init_final_bn=cfg.RESNET.ZERO_INIT_FINAL_BN,
        fc_init_std=cfg.MODEL.FC_INIT_STD, crop=int(getattr(cfg.DATA, "CROP_SIZE", 224)),
        mean=tuple(cfg.DATA.MEAN), std=tuple(cfg.DATA.STD),
        reverse_input_channel=bool(cfg.DATA.REVERSE_INPUT_CHANNEL),
        sampling_rate=int(cfg.DATA.SAMPLING_RATE), target_fps=int(getattr(cfg.DATA, "TARGET_FPS", 30)),
        norm_type=str(cfg.BN.NORM_TYPE), num_splits=int(getattr(cfg.BN, "NUM_SPLITS", 1)),
    )
