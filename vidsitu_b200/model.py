"""Parameter containers with the reference's state_dict key set.

`BackboneParams` is what `SFBase.sf_mdl` is in this package: an nn.Module tree whose
parameter / buffer names equal those of the reference's `SlowFast_FeatModel` /
`ResNet_FeatModel` (vidsitu_code/mdl_sf_base.py:20-62), so checkpoints load with
`load_state_dict` / `load_checkpoint(model=mdl.sf_mdl)` unchanged
(vidsitu_code/feat_extractor.py:147-161).  It holds no torch compute modules -- the
forward is the kernel plan in engine.py.
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch
from torch import nn

from .arch import ConvSpec, NetSpec


class ConvParams(nn.Module):
    """weight (+ bias) of one nn.Conv3d, reference layout [cout, cin, kt, kh, kw]."""

    def __init__(self, spec: ConvSpec):
        super().__init__()
        self.weight = nn.Parameter(torch.empty((spec.cout, spec.cin) + tuple(spec.kernel)))
        if spec.has_bias:
            self.bias = nn.Parameter(torch.zeros(spec.cout))
        else:
            self.register_parameter("bias", None)


class BNParams(nn.Module):
    """Parameters and running statistics of one frozen nn.BatchNorm3d (eps 1e-5)."""

    eps = 1e-5  # constructor default; cfg.BN.EPSILON is never passed (resnet_helper.py:126-127)

    def __init__(self, c: int):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class _Node(nn.Module):
    pass


class _BNStats(nn.Module):
    """Running statistics of an affine-less nn.BatchNorm3d (the inner `bn` / `split_bn` of SubBatchNorm3d)."""

    def __init__(self, c: int):
        super().__init__()
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class SubBNParams(nn.Module):
    """state_dict layout of SubBatchNorm3d (SlowFast/slowfast/models/batchnorm_helper.py:37-109): one weight /
    bias pair, `bn.*` = the aggregated statistics eval mode uses, `split_bn.*` = the per-split statistics
    training maintains (num_splits * c).  Frozen: only `bn.*`, weight and bias reach the kernels."""

    eps = 1e-5

    def __init__(self, c: int, num_splits: int):
        super().__init__()
        self.num_splits = int(num_splits)
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.bn = _BNStats(c)
        self.split_bn = _BNStats(c * self.num_splits)

    @torch.no_grad()
    def aggregate_stats(self) -> None:
        """`SubBatchNorm3d.aggregate_stats` (batchnorm_helper.py:64-95): mean of the split means, mean of the split
        variances plus the variance of the split means."""
        n = self.num_splits
        means = self.split_bn.running_mean.view(n, -1)
        mean = means.sum(0) / n
        var = self.split_bn.running_var.view(n, -1).sum(0) / n + ((means - mean) ** 2).sum(0) / n
        self.bn.running_mean.copy_(mean)
        self.bn.running_var.copy_(var)


def aggregate_sub_bn_stats(module: nn.Module) -> int:
    """Call `aggregate_stats()` on every sub-batchnorm of `module` (what the reference's trainer does before
    evaluation); works on this package's SubBNParams and on the reference's SubBatchNorm3d alike."""
    count = 0
    for m in module.modules():
        if hasattr(m, "aggregate_stats") and hasattr(m, "split_bn"):
            m.aggregate_stats()
            count += 1
    return count


def _attach(root: nn.Module, dotted: str, leaf: nn.Module) -> None:
    parts = dotted.split(".")
    node = root
    for p in parts[:-1]:
        if not hasattr(node, p):
            node.add_module(p, _Node())
        node = getattr(node, p)
    node.add_module(parts[-1], leaf)


class BackboneParams(nn.Module):
    def __init__(self, spec: NetSpec):
        super().__init__()
        self.spec = spec
        self.num_pathways = spec.num_pathways
        for conv in spec.all_convs():
            _attach(self, conv.key, ConvParams(conv))
            if conv.bn is not None:
                _attach(self, conv.bn, SubBNParams(conv.cout, spec.num_splits) if spec.norm_type == "sub_batchnorm"
                        else BNParams(conv.cout))
        # the backbone's own classification head exists in the reference state_dict
        # (head.projection.{weight,bias}, head_helper.py:185) although SFBase never calls it
        head = _Node()
        head.add_module("projection", nn.Linear(sum(spec.feat_dims), spec.num_classes, bias=True))
        self.add_module("head", head)
        self.reset_parameters()

    @torch.no_grad()
    def reset_parameters(self) -> None:
        """Random init of the same distributions as init_weights
        (SlowFast/slowfast/utils/weight_init_helper.py:10-43): conv ~ N(0, sqrt(2/fan_out)),
        BN gamma 1 (0 for a block's final BN when ZERO_INIT_FINAL_BN), beta 0,
        Linear ~ N(0, FC_INIT_STD), bias 0."""
        spec = self.spec
        for conv in spec.all_convs():
            m = self.get_submodule(conv.key)
            fan_out = conv.cout * conv.kernel[0] * conv.kernel[1] * conv.kernel[2]
            m.weight.normal_(0.0, math.sqrt(2.0 / fan_out))
            if m.bias is not None:
                m.bias.zero_()
            if conv.bn is not None:
                bn = self.get_submodule(conv.bn)
                bn.weight.fill_(0.0 if (conv.final_bn and spec.zero_init_final_bn) else 1.0)
                bn.bias.zero_()
                for st in ([bn.bn, bn.split_bn] if isinstance(bn, SubBNParams) else [bn]):
                    st.running_mean.zero_()
                    st.running_var.fill_(1.0)
        self.head.projection.weight.normal_(0.0, spec.fc_init_std)
        self.head.projection.bias.zero_()

    def named_tensors(self) -> Dict[str, torch.Tensor]:
        return dict(self.state_dict())


@torch.no_grad()
def randomize_bn_(module: nn.Module, seed: int = 1) -> None:
    """Parity harness helper (hard part 1): the stock init zeroes the
    last BN of every bottleneck, which would hide broken a/b/c convs.  Overwrite every
    BN's gamma/beta/mean/var with seeded non-trivial values; works on this package's
    BNParams and on torch.nn.BatchNorm3d alike (it only touches tensors by name)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = module.state_dict()
    for k in sorted(sd.keys()):
        if not k.endswith("running_var"):
            continue
        base = k[: -len("running_var")]
        c = sd[k].numel()
        if base.endswith(".split_bn."):
            # sub-batchnorm: randomise the per-split statistics; the aggregated `bn.*` ones follow below
            sd[base + "running_mean"].copy_(torch.randn(c, generator=g) * 0.1)
            sd[base + "running_var"].copy_(torch.rand(c, generator=g) + 0.5)
            continue
        if base.endswith(".bn.") and (base[:-3] + "weight") in sd and (base[:-3] + "split_bn.running_var") in sd:
            outer = base[:-3]           # sub-batchnorm: weight / bias live one level up, statistics are aggregated
            if ".c_bn." in outer or "_nonlocal" in outer:
                sd[outer + "weight"].copy_(torch.rand(c, generator=g) * 0.3 + 0.2)
            else:
                sd[outer + "weight"].copy_(torch.rand(c, generator=g) + 0.5)
            sd[outer + "bias"].copy_(torch.randn(c, generator=g) * 0.1)
            continue
        if ".c_bn." in base or "_nonlocal" in base:
            # last BN of a residual branch: keep the branch gain below 1 so 16 stacked blocks stay O(1)
            sd[base + "weight"].copy_(torch.rand(c, generator=g) * 0.3 + 0.2)    # gamma ~ U[0.2, 0.5]
        else:
            sd[base + "weight"].copy_(torch.rand(c, generator=g) + 0.5)      # gamma ~ U[0.5, 1.5]
        sd[base + "bias"].copy_(torch.randn(c, generator=g) * 0.1)           # beta  ~ N(0, 0.1)
        sd[base + "running_mean"].copy_(torch.randn(c, generator=g) * 0.1)   # mean  ~ N(0, 0.1)
        sd[base + "running_var"].copy_(torch.rand(c, generator=g) + 0.5)     # var   ~ U[0.5, 1.5]
    aggregate_sub_bn_stats(module)
