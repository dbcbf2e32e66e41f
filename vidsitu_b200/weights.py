"""One-time weight preparation (host side, torch as plumbing): reference-layout
parameters -> the layouts the kernels read.

  * conv weights  [cout, cin, kt, kh, kw] fp32 (nn.Conv3d, state_dict layout
    SURVEY.md section 8 a20)  ->  K-major [cout_store][kt*kh*kw][cin_store];
  * frozen BatchNorm3d (SlowFast/slowfast/models/batchnorm_helper.py:15-24, eps
    1e-5 constructor default) -> per-channel fp32 (scale, bias) applied in the conv
    epilogue -- kept in fp32 instead of being folded into bf16 weights;
  * the Cin=3 stems (stem_helper.py:157-178) -> "quad view" weights, see
    `stem_quad_weight`.
"""
from __future__ import annotations

import torch

CH_QUANTUM = 16  # stored channel counts are multiples of 16 (UMMA K granularity for bf16)


def round_up(x: int, q: int = CH_QUANTUM) -> int:
    return (x + q - 1) // q * q


def pack_conv_weight(w: torch.Tensor, cin_store: int, cout_store: int, dtype: torch.dtype) -> torch.Tensor:
    cout, cin, kt, kh, kw = w.shape
    out = torch.zeros((cout_store, kt * kh * kw, cin_store), dtype=torch.float32, device=w.device)
    out[:cout, :, :cin] = w.permute(0, 2, 3, 4, 1).reshape(cout, kt * kh * kw, cin)
    return out.to(dtype).contiguous()


def bn_tensor_keys(tensors, bn_key: str):
    """(weight, bias, running_mean, running_var) key names of the frozen norm layer `bn_key` in a reference-layout
    state_dict: nn.BatchNorm3d / NaiveSyncBatchNorm3d keep all four side by side, SubBatchNorm3d keeps the
    statistics its eval forward uses under `<bn_key>.bn.*` (batchnorm_helper.py:55-62, 97-109)."""
    stats = bn_key + ".bn" if (bn_key + ".bn.running_mean") in tensors else bn_key
    return bn_key + ".weight", bn_key + ".bias", stats + ".running_mean", stats + ".running_var"


def fold_bn(gamma, beta, mean, var, eps: float, cout_store: int, conv_bias=None):
    """y = gamma * (x + conv_bias - mean) / sqrt(var + eps) + beta  ==  scale * x + bias."""
    scale = gamma.double() / torch.sqrt(var.double() + eps)
    bias = beta.double() - mean.double() * scale
    if conv_bias is not None:
        bias = bias + conv_bias.double() * scale
    s = torch.zeros(cout_store, dtype=torch.float32, device=gamma.device)
    b = torch.zeros(cout_store, dtype=torch.float32, device=gamma.device)
    s[: scale.numel()] = scale.float()
    b[: bias.numel()] = bias.float()
    return s, b


def identity_affine(cout: int, cout_store: int, conv_bias, device):
    s = torch.zeros(cout_store, dtype=torch.float32, device=device)
    b = torch.zeros(cout_store, dtype=torch.float32, device=device)
    s[:cout] = 1.0
    if conv_bias is not None:
        b[:cout] = conv_bias.float()
    return s, b


def stem_quad_weight(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """Stem conv [cout, 3, kt, 7, 7], stride (1,2,2), pad (kt//2,3,3) restated on the
    "quad view" of the packed frames.

    The pack kernel writes 4 channels per pixel, so four neighbouring pixels are 16
    contiguous bf16 values: [n,t,224,224,4] *is* [n,t,224,56,16].  Two neighbouring
    outputs (wo = 2j, 2j+1) read input columns 4j-3 .. 4j+5, i.e. quads j-1, j, j+1.
    The stem therefore equals a conv with cin'=16, cout'=2*cout, kernel (kt,7,3),
    stride (1,2,1), pad (kt//2,3,1) whose [.., 56, 2*cout] output is memory-identical
    to the reference's [.., 112, cout]:
        W'[p*cout+co, (kt,kh,qt), r*4+ci] = W[co, ci, kt, kh, kw],  4*qt + r = 2p + 1 + kw.
    """
    cout, cin, kt, kh, kw = w.shape
    assert cin <= 4 and kh == 7 and kw == 7
    out = torch.zeros((2, cout, kt, kh, 3, 4, 4), dtype=torch.float32, device=w.device)  # p,co,kt,kh,qt,r,ci
    for p in range(2):
        for k in range(kw):
            pos = 2 * p + 1 + k
            qt, r = pos // 4, pos % 4
            out[p, :, :, :, qt, r, :cin] = w[:, :, :, :, k].permute(0, 2, 3, 1)
    return out.reshape(2 * cout, kt * kh * 3, 16).to(dtype).contiguous()


def group_conv_weight(w: torch.Tensor, cin_store: int, cout_store: int, J: int, stride_w: int, pad_w: int,
                      dtype: torch.dtype):
    """Restate a conv along W on "pixel groups": J neighbouring output pixels become one GEMM row
    with J*cout_store channels, G = J*stride_w neighbouring input pixels one row of G*cin_store
    channels.  Memory is untouched -- [.., W, C] *is* [.., W/G, G*C] -- only the weights are
    expanded (block-Toeplitz along W):

        output pixel  J*j + p   reads input pixel  G*j + off,   off = p*stride_w - pad_w + kw
        W'[p*cout_store + co, (kt, kh, gt - gmin), r*cin_store + ci] = W[co, ci, kt, kh, kw]
        with gt = floor(off / G), r = off mod G.

    Why: the im2col TMA unit issues one request per pixel row; rows of 16-32 channels (32-64 B) are
    request-rate bound, 128 B rows are not.  The extra zero MACs are free on layers that are
    nowhere near the tensor-pipe roofline (stems, the 8..64-channel Fast pathway).

    Returns (packed [J*cout_store, kt*kh*ngt, G*cin_store], ngt, pad_lo) where the grouped conv has
    kernel width ngt, stride 1 and leading pad pad_lo along W'."""
    cout, cin, kt, kh, kw = w.shape
    G = J * stride_w
    offs = [p * stride_w - pad_w + k for p in range(J) for k in range(kw)]
    gmin, gmax = min(offs) // G, max(offs) // G
    ngt = gmax - gmin + 1
    out = torch.zeros((J, cout_store, kt, kh, ngt, G, cin_store), dtype=torch.float32, device=w.device)
    wp = w.permute(0, 2, 3, 1, 4)  # cout, kt, kh, cin, kw
    for p in range(J):
        for k in range(kw):
            off = p * stride_w - pad_w + k
            out[p, :cout, :, :, off // G - gmin, off % G, :cin] = wp[..., k]
    return out.reshape(J * cout_store, kt * kh * ngt, G * cin_store).to(dtype).contiguous(), ngt, -gmin


def group_tap_ranges(kw: int, cin_store: int, J: int, stride_w: int, pad_w: int):
    """Per grouped tap gt of `group_conv_weight`: the 16-aligned channel range [lo, hi) of the
    G*cin_store-channel input group that carries non-zero weights (the outer groups of a
    block-Toeplitz tap row only contribute their first / last few pixels)."""
    G = J * stride_w
    offs = [p * stride_w - pad_w + k for p in range(J) for k in range(kw)]
    gmin, gmax = min(offs) // G, max(offs) // G
    ranges = []
    for gt in range(gmin, gmax + 1):
        rs = [o % G for o in offs if o // G == gt]
        lo = (min(rs) * cin_store) // 16 * 16
        hi = -(-((max(rs) + 1) * cin_store) // 16) * 16
        ranges.append((lo, min(hi, G * cin_store)))
    return ranges


def slice_tap_channels(w: torch.Tensor, taps_outer: int, kw: int, ranges) -> torch.Tensor:
    """[cout, taps_outer*kw, cin] -> [cout, taps_outer * sum(hi-lo)] keeping, for every kw tap, only
    its channel range (the layout the window conv kernel expects when ranges are partial)."""
    cout, taps, cin = w.shape
    assert taps == taps_outer * kw and len(ranges) == kw
    v = w.view(cout, taps_outer, kw, cin)
    parts = [v[:, :, k, lo:hi] for k, (lo, hi) in enumerate(ranges)]
    return torch.cat(parts, dim=2).reshape(cout, -1).contiguous()
