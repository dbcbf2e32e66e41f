"""Kernel plan of one event-clip forward (the body of the hot path).

`ClipEngine` turns a `NetSpec` + a reference-layout state_dict into a static program
of C-ABI calls over a fixed set of device buffers for a fixed clip count `n`:

  inputs   [n, T_p, 224, 224, 4]   filled by the pack kernel (uint8 frames) or by the
                                   NCTHW->NTHWC converter (reference fp32 tensors)
  trunk    stems -> (lateral) -> res2..res5 (+ non-local), every conv one launch with
           BN/ReLU/residual folded into its epilogue; the Fast->Slow lateral conv
           writes straight into the channel slice of the slow tensor it is
           concatenated to (video_model_builder.py:124-131), so no torch.cat pass
  head     global average pool -> feats [n, D] fp32 -> Linear/ReLU/Linear -> logits

It mirrors the stage order of SlowFast_FeatModel / ResNet_FeatModel.forward_features
(vidsitu_code/mdl_sf_base.py:21-34, 46-55).  The whole program is stream-ordered and
allocation-free, so it is captured once into a CUDA graph and replayed.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from .arch import BlockSpec, ConvSpec, NetSpec, NonlocalSpec
from .lib import VSB_BF16, VSB_F32, VsbError
from .ops import Act, BottleneckPlan, ConvPlan
from .weights import (bn_tensor_keys, fold_bn, group_conv_weight, group_tap_ranges, identity_affine, pack_conv_weight, round_up,
                      slice_tap_channels)


_PLAN_KNOBS = ("block_n", "kchunk", "stages", "epi_n", "epi_bufs", "flags")
_TUNE_TABLE: Optional[dict] = None


def _tune_table() -> dict:
    """Per-shape plan tuning measured on B200 (tools/autotune.py); missing file = all automatic."""
    global _TUNE_TABLE
    if _TUNE_TABLE is None:
        import json
        p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tune_table.json")
        _TUNE_TABLE = json.load(open(p)).get("entries", {}) if os.path.exists(p) else {}
    return _TUNE_TABLE


class _Pool:
    """Stream-ordered buffer reuse: activations die as soon as their last reader has been
    enqueued, and every op of a pool runs on one stream, so a freed buffer can be
    handed to the next producer without a fence."""

    def __init__(self, device, elem_dtype: torch.dtype):
        self.device = device
        self.dtype = elem_dtype
        self.free: List[torch.Tensor] = []
        self.all: List[torch.Tensor] = []     # every buffer ever handed out (scratch regions of a saved program)
        self.total_bytes = 0

    def take(self, numel: int) -> torch.Tensor:
        best = None
        for i, b in enumerate(self.free):
            if b.numel() >= numel and (best is None or b.numel() < self.free[best].numel()):
                best = i
        if best is not None and self.free[best].numel() <= 2 * numel:
            return self.free.pop(best)
        buf = torch.empty(numel, dtype=self.dtype, device=self.device)
        self.total_bytes += buf.numel() * buf.element_size()
        self.all.append(buf)
        return buf

    def give(self, buf: torch.Tensor) -> None:
        self.free.append(buf)


class _SlotLaunch:
    """The stem conv of one pathway: one plan per input slot (the slots differ only in the input buffer), the
    launch runs the plan of the slot selected at launch / capture / program-build time."""

    def __init__(self, engine: "ClipEngine", plans: list, launches: int = 1):
        self.engine = engine
        self.plans = plans
        self.launches = launches    # kernels per run (the fused stem zero-fills its output first)

    def __call__(self) -> None:
        self.plans[self.engine._slot].run()

    def emit(self, prog, lane: int, name: str) -> None:
        self.plans[self.engine._slot].emit(prog, lane, name)

    @property
    def _keep(self):
        return self.plans[self.engine._slot]._keep


class ClipEngine:
    def __init__(self, spec: NetSpec, tensors: Dict[str, torch.Tensor], n: int, dtype: int = VSB_BF16,
                 device: Optional[torch.device] = None, proj_head: Optional[Sequence[torch.Tensor]] = None,
                 tune: Optional[dict] = None, bn_eps: float = 1e-5, input_slots: int = 1,
                 prep_cache: Optional[dict] = None):
        if not torch.cuda.is_available():
            raise VsbError("ClipEngine needs a CUDA device: the forward is made of sm_100a kernels only")
        self.spec = spec
        self.n = int(n)
        self.dtype = dtype
        self.tdt = ops.TORCH_DTYPE[dtype]
        self.device = torch.device(device if device is not None else "cuda")
        self.tune = dict(tune or {})
        self.bn_eps = bn_eps
        self._t = tensors
        # prepared (folded / packed / uploaded) weights, keyed by layer and layout: shared by the engines of one model
        # (one per batch size), so only the first engine pays for the preparation
        self._cache = prep_cache if prep_cache is not None else {}
        self._keep: List[torch.Tensor] = []
        self.trunk_ops: List[Tuple[str, Callable[[], None], float]] = []   # (name, launch, flops)
        self.head_ops: List[Tuple[str, Callable[[], None], float]] = []
        # one buffer pool per pathway: the two pathways may run on two streams, and a pool hands buffers
        # back out in stream order only
        self._pools = [_Pool(self.device, self.tdt) for _ in range(max(1, spec.num_pathways))]
        self._pool = self._pools[0]
        self._graph = None
        self._programs = None      # one captured clip program per input slot (replay through the C ABI)
        self.frames_in = None      # static uint8 frame buffer of an exported program
        # replay path: "program" = ONE C call per forward (vsb_program_run on a CUDA graph captured inside the
        # library), "torch" = a torch.cuda.CUDAGraph of the Python launch loop
        self.replay_mode = str(self.tune.get("*", {}).get("replay", os.environ.get("VSB_REPLAY", "program")))
        self._main_stream = None
        self.two_streams = (spec.num_pathways == 2 and dtype == VSB_BF16
                            and str(self.tune.get("*", {}).get("streams", os.environ.get("VSB_STREAMS", "2"))) == "2")
        self._side_stream = None
        # Lateral convs write a channel slice that no slow-pathway kernel touches, so with dedicated concat
        # buffers the Fast stream never waits for the Slow one (only the next slow stage waits for the lateral).
        self.early_lateral = self.two_streams and str(self.tune.get("*", {}).get(
            "early_lateral", os.environ.get("VSB_EARLY_LATERAL", "1"))) == "1"
        # Direction of the tile walk, alternating along each pathway's chain a -> b -> c -> a' ...: a kernel that
        # starts where its producer finished finds the most recently written part of its input still in L2.
        self.alternate = str(self.tune.get("*", {}).get("alternate", os.environ.get("VSB_ALTERNATE", "1"))) == "1"
        self._rev = [True] * max(1, spec.num_pathways)   # the stem pools write ascending: the first conv walks down
        self._pw = 0
        self._dedicated: List[torch.Tensor] = []
        self.dedicated_bytes = 0
        self.op_bytes: Dict[str, float] = {}
        self.op_sig: Dict[str, str] = {}
        self.fused_shortcuts: List[str] = []   # branch1 convs computed inside their block's last conv
        self.chained: List[tuple] = []  # (last conv of a block, first conv of the next) run as chained launches
        self.fused_blocks: List[str] = []      # identity blocks that run as one fused a-b-c launch
        self.fused_stems: List[str] = []       # stems that run conv + BN + ReLU + max-pool as one kernel
        self.crop = spec.crop
        if self.crop % 16:
            raise VsbError("crop size must be a multiple of 16")
        frames = spec.pathway_frames()
        # ---- inputs: 4 channels per pixel.  On the tensor-core path every row carries its own zero
        # border (3 pixels left = the stem's W padding, 13 right), so the stem conv runs on 16-pixel
        # groups without any W padding (weights.group_conv_weight); the border is zeroed once here and
        # never written again.
        self.x_off = 3 if dtype == VSB_BF16 else 0
        self.w_buf = self.crop + 16 if dtype == VSB_BF16 else self.crop
        # `input_slots` > 1: several input buffer sets, so that the frames of batch k+1 can be packed (on another
        # stream) while the trunk still reads batch k; everything behind the stems is shared
        self.input_sets: List[List[Act]] = []
        self._slot = 0
        for _ in range(max(1, int(input_slots))):
            acts = []
            for t in frames:
                buf = torch.zeros(self.n * t * self.crop * self.w_buf * 4, dtype=self.tdt, device=self.device)
                acts.append(Act(buf, self.n, t, self.crop, self.w_buf, 4, 4, c_real=3))
            self.input_sets.append(acts)
        # slow-pathway temporal indices, exactly as utils/video_utils.py:62-69
        t_fast = spec.num_frames
        self.fast_idx = list(range(t_fast))
        if spec.num_pathways == 2:
            self.slow_idx = torch.linspace(0, t_fast - 1, t_fast // spec.alpha).long().tolist()
        # ---- trunk
        self.trunk_out = self._build_trunk()
        # ---- head
        d = sum(spec.feat_dims)
        self.feats = torch.zeros((self.n, d), dtype=torch.float32, device=self.device)
        off = 0
        for p, x in enumerate(self.trunk_out):
            self.head_ops.append((f"head.pathway{p}_avgpool",
                                  ops.global_avgpool_call(x, self.feats, off, self.dtype), 0.0))
            off += x.c_real
        self.logits = None
        if proj_head is not None:
            w0, b0, w1, b1 = self._memo(("proj_head",), lambda: tuple(
                t.detach().to(self.device, torch.float32).contiguous() for t in proj_head))
            self._keep += [w0, b0, w1, b1]
            self.hidden = torch.zeros((self.n, w0.shape[0]), dtype=torch.float32, device=self.device)
            self.logits = torch.zeros((self.n, w1.shape[0]), dtype=torch.float32, device=self.device)
            self.head_ops.append(("proj_head.0", ops.linear_call(self.feats, w0, b0, self.hidden, True),
                                  2.0 * self.n * w0.numel()))
            self.head_ops.append(("proj_head.2", ops.linear_call(self.hidden, w1, b1, self.logits, False),
                                  2.0 * self.n * w1.numel()))
        self._t = None  # drop the reference-layout tensors

    # ------------------------------------------------------------------ building blocks
    def _store(self, c: int) -> int:
        """Stored channel count.  8-channel tensors (the Fast pathway's res2 bottleneck) stay 8 wide on the
        tensor-core path: they are only ever touched through pixel-group views (8 pixels x 8 channels = one
        128-byte GEMM row), so padding them to the 16-channel MMA quantum would just double their traffic."""
        if self.dtype == VSB_BF16 and c <= 8 and self.tune.get("*", {}).get("pad8", False) is False:
            return round_up(c, 8)
        return round_up(c)

    def _alloc(self, n, t, h, w, c_real, pitch: Optional[int] = None, min_c: int = 0) -> Act:
        c = max(self._store(c_real), min_c)
        pitch = pitch or c
        if pitch > c and self.early_lateral:
            # a slow tensor with room for the lateral channels: the Fast pathway's stream writes its slice
            # without waiting for the slow stage, so the buffer is never shared with another tensor
            buf = torch.empty(n * t * h * w * pitch, dtype=self.tdt, device=self.device)
            self._dedicated.append(buf)
            self.dedicated_bytes += buf.numel() * buf.element_size()
        else:
            buf = self._pool.take(n * t * h * w * pitch)
        return Act(buf, n, t, h, w, c, pitch, 0, c_real)

    def _free(self, a: Act) -> None:
        if a.buf is None or any(a.buf is i.buf for acts in self.input_sets for i in acts) or \
                any(a.buf is d for d in self._dedicated):
            return
        self._pool.give(a.buf)

    def _tensor(self, key: str) -> torch.Tensor:
        """Reference-layout parameter as an fp32 HOST tensor: BatchNorm folding and weight packing are done on
        the CPU (one upload per prepared tensor instead of a dozen tiny device kernels per layer)."""
        if key not in self._t:
            raise KeyError(f"state_dict has no {key!r}")
        return self._t[key].detach().float().cpu()

    def _memo(self, key: tuple, make):
        """Prepared tensors of one layer / layout, built once per model (`prep_cache`)."""
        if key not in self._cache:
            self._cache[key] = make()
        return self._cache[key]

    def _up(self, t: torch.Tensor) -> torch.Tensor:
        return t.contiguous().to(self.device)

    def _affine(self, cs: ConvSpec, cout_store: int):
        """Folded frozen BatchNorm (or conv bias) of a conv as HOST fp32 (scale, bias)."""
        bias = self._tensor(cs.key + ".bias") if cs.has_bias else None
        if cs.bn is not None:
            kw, kb, km, kv = bn_tensor_keys(self._t, cs.bn)
            return fold_bn(self._tensor(kw), self._tensor(kb), self._tensor(km), self._tensor(kv),
                           self.bn_eps, cout_store, bias)
        return identity_affine(cs.cout, cout_store, bias, "cpu")

    def _sb(self, cs: ConvSpec, cout_store: int, j: int = 1):
        """Device (scale, bias) of a conv, repeated for j-pixel groups."""
        def make():
            s, b = self._affine(cs, cout_store)
            return self._up(s.repeat(j)), self._up(b.repeat(j))
        return self._memo((cs.key, "sb", cout_store, j), make)

    def _out_dims(self, x: Act, cs: ConvSpec):
        to = (x.t + 2 * cs.pad[0] - cs.kernel[0]) // cs.stride[0] + 1
        ho = (x.h + 2 * cs.pad[1] - cs.kernel[1]) // cs.stride[1] + 1
        wo = (x.w + 2 * cs.pad[2] - cs.kernel[2]) // cs.stride[2] + 1
        return to, ho, wo

    def _tune(self, key: str, sig: Optional[str] = None) -> dict:
        """Tuning of one conv, lowest to highest priority: the shipped per-shape table
        (vidsitu_b200/tune_table.json, measured on B200 by tools/autotune.py), the caller's "*"
        entry, the caller's per-layer entry."""
        tune = dict(_tune_table().get(sig, {})) if (sig and self.tune.get("*", {}).get("table", True)) else {}
        tune.update(self.tune.get("*", {}))
        tune.update(self.tune.get(key, {}))
        return tune

    @staticmethod
    def _sig(cs: ConvSpec, x: Act, out: Act, residual: Optional[Act], x2: Optional[Act] = None) -> str:
        """Shape signature of a conv launch (the key of the tuning table)."""
        k, st = cs.kernel, cs.stride
        return (f"{x.c}>{out.c}|k{k[0]}{k[1]}{k[2]}|s{st[0]}{st[1]}{st[2]}|o{out.t}x{out.h}x{out.w}"
                f"|p{x.pitch}>{out.pitch}|r{int(residual is not None)}|x{x2.c if x2 is not None else 0}")

    def _group_factor(self, cs: ConvSpec, x: Act, out: Act, residual: Optional[Act]) -> int:
        """Pixels per GEMM row (weights.group_conv_weight): thin-channel layers are re-viewed so that
        one im2col row is >= 128 bytes."""
        if self.dtype != VSB_BF16 or x.c >= 64:
            return 1
        dense = lambda a: a is None or (a.pitch == a.c and a.c_off == 0)
        if not (dense(x) and dense(out) and dense(residual)):
            return 1
        j = self._tune(cs.key).get("group", 64 // x.c)
        sw = cs.stride[2]
        while j > 1 and (out.w % j or x.w % (j * sw) or out.w // j != x.w // (j * sw)):
            j //= 2
        return max(j, 1)

    def _window_plan(self, cs: ConvSpec, x: Act, out: Act, residual: Optional[Act], relu: bool, wt,
                     pad_w: Optional[int] = None, j: Optional[int] = None, rev_flag: int = 0) -> Optional[ConvPlan]:
        """Shared-memory window algorithm (conv_win_sm100.cu) for convs with spatial taps and <= 64
        (grouped) input channels: returns the plan, or None when the layer is outside its domain."""
        tn = self._tune(cs.key, self._sig(cs, x, out, residual))
        if self.dtype != VSB_BF16 or tn.get("algo", "auto") == "im2col":
            return None
        if cs.stride[0] != 1 or x.h < 14:
            return None
        if cs.kernel[1] * cs.kernel[2] == 1:
            # kt x 1 x 1 (the Fast pathway's temporal `a` convs): the ring of frame windows reads every input frame
            # once, the im2col kernel once per temporal tap (through L2).  VSB_WIN_TEMPORAL=1 / tune win_temporal
            if cs.kernel[0] == 1 or str(tn.get("win_temporal", os.environ.get("VSB_WIN_TEMPORAL", "0"))) not in ("1", "True"):
                return None
        dense = lambda a: a is None or (a.pitch == a.c and a.c_off == 0)
        sw = cs.stride[2]
        if j is None:
            j = tn.get("win_group", max(1, (64 // x.c) // sw) if x.c <= 64 else 0)
        if j < 1:
            return None
        # stems pass pad_w: their input rows carry a zero border, so grouped widths need not match
        while j > 1 and (out.w % j or x.w % (j * sw) or (pad_w is None and out.w // j != x.w // (j * sw))):
            j //= 2
        g = j * sw
        if g * x.c > 64 or (g * x.c) % 16 or (j * out.c) % 16 or j * out.c > 256:
            return None
        if g > 1 and not (dense(x) and dense(out) and dense(residual)):
            return None
        if x.w % g or out.w % j:
            return None
        pad_w = cs.pad[2] if pad_w is None else pad_w

        def make():
            w, ngt, plo = group_conv_weight(wt, x.c, out.c, j, sw, pad_w, self.tdt)
            if ngt > 8 or plo < 0:
                return None, ngt, plo, None
            ranges = group_tap_ranges(cs.kernel[2], x.c, j, sw, pad_w)
            return self._up(slice_tap_channels(w, cs.kernel[0] * cs.kernel[1], ngt, ranges)), ngt, plo, ranges
        w, ngt, plo, ranges = self._memo((cs.key, "win", x.c, out.c, j, sw, pad_w), make)
        if w is None:
            return None
        scale, bias = self._sb(cs, out.c, j)
        xin = Act(x.buf, x.n, x.t, x.h, x.w // g, g * x.c, g * x.pitch if g == 1 else g * x.c, x.c_off if g == 1 else 0)
        yout = Act(out.buf, out.n, out.t, out.h, out.w // j, j * out.c, out.pitch if j == 1 else j * out.c,
                   out.c_off if j == 1 else 0)
        res = None
        if residual is not None:
            res = Act(residual.buf, out.n, out.t, out.h, out.w // j, j * out.c,
                      residual.pitch if j == 1 else j * out.c, residual.c_off if j == 1 else 0)
        phi = yout.w - 1 + ngt - xin.w - plo
        try:
            return ConvPlan(self.dtype, xin, w, j * out.c, (cs.kernel[0], cs.kernel[1], ngt),
                            (cs.stride[0], cs.stride[1], 1), (cs.pad[0], cs.pad[1], plo), (cs.pad[0], cs.pad[1], phi),
                            scale, bias, yout, res, relu, algo=2, kw_ranges=ranges,
                            **dict({k: v for k, v in tn.items() if k in ("stages", "epi_n", "epi_bufs")},
                                   flags=int(tn.get("flags", 0)) | rev_flag))
        except VsbError:
            if tn.get("algo") == "window":
                raise
            return None

    def _conv(self, cs: ConvSpec, x: Act, out: Act, residual: Optional[Act] = None, relu: Optional[bool] = None,
              reverse: bool = False, chain: Optional[dict] = None, out_f16: bool = False):
        """Plans one conv launch and appends it to the trunk.  `chain` (tile-granular chaining, _chain_kwargs):
        ConvPlan keywords tile_signal / tile_wait; the launch is then returned instead of appended (the caller
        emits producer and consumer as one trunk op)."""
        if x.c_real != cs.cin:
            raise VsbError(f"{cs.key}: input has {x.c_real} channels, conv expects {cs.cin}")
        sig = self._sig(cs, x, out, residual)
        self.op_sig[cs.key] = sig
        tune = {k: v for k, v in self._tune(cs.key, sig).items() if k in _PLAN_KNOBS}
        rev_flag = 128 if (reverse and self.alternate and self.dtype == VSB_BF16) else 0   # VSB_PLAN_REVERSE
        tune["flags"] = int(tune.get("flags", 0)) | rev_flag
        if chain:
            tune = self._chain_tune(tune, rev_flag)
            tune.update(chain)
        relu = cs.relu if relu is None else relu
        wt = self._tensor(cs.key + ".weight")
        plan = self._window_plan(cs, x, out, residual, relu, wt, rev_flag=rev_flag) if not (chain or out_f16) else None
        j = self._group_factor(cs, x, out, residual) if plan is None else 0
        if chain and j > 1:
            raise VsbError(f"{cs.key}: tile chaining is for ungrouped layers")
        bn = tune.get("block_n", 0)
        if bn and ((max(j, 1) * out.c) % bn or bn > max(j, 1) * out.c):
            tune.pop("block_n")     # a table / "*" entry that does not fit this layer: automatic
        if plan is not None:
            pass
        elif j > 1:
            sw = cs.stride[2]
            g = j * sw
            w, ngt, plo = self._memo((cs.key, "grp", x.c, out.c, j, sw, cs.pad[2]), lambda: (lambda r: (
                self._up(r[0]), r[1], r[2]))(group_conv_weight(wt, x.c, out.c, j, sw, cs.pad[2], self.tdt)))
            scale, bias = self._sb(cs, out.c, j)
            xin = Act(x.buf, x.n, x.t, x.h, x.w // g, g * x.c, g * x.c)
            yout = Act(out.buf, out.n, out.t, out.h, out.w // j, j * out.c, j * out.c)
            res = None if residual is None else Act(residual.buf, out.n, out.t, out.h, out.w // j, j * out.c,
                                                    j * out.c)
            phi = yout.w - 1 + ngt - xin.w - plo
            plan = ConvPlan(self.dtype, xin, w, j * out.c, (cs.kernel[0], cs.kernel[1], ngt),
                            (cs.stride[0], cs.stride[1], 1), (cs.pad[0], cs.pad[1], plo), (cs.pad[0], cs.pad[1], phi),
                            scale, bias, yout, res, relu, **tune)
        else:
            w = self._memo((cs.key, "plain", x.c, out.c), lambda: self._up(pack_conv_weight(wt, x.c, out.c, self.tdt)))
            scale, bias = self._sb(cs, out.c)
            plan = ConvPlan(self.dtype, x, w, out.c, cs.kernel, cs.stride, cs.pad, None, scale, bias, out, residual,
                            relu, out_f16=out_f16, **tune)
        if out_f16 and (j > 1 or self.dtype != VSB_BF16 or residual is not None):
            raise VsbError(f"{cs.key}: half outputs are for plain bf16-path convs without residual")
        self._keep.append(plan)
        m = out.pixels
        es = 2 if self.dtype == VSB_BF16 else 4
        self.op_bytes[cs.key] = es * (x.pixels * x.c + m * out.c * (2 if residual is not None else 1))
        op = (cs.key, plan.run, float(m) * cs.flops_per_out_pixel)
        if chain:
            return op, plan
        self.trunk_ops.append(op)

    @staticmethod
    def _chain_tune(tune: dict, rev_flag: int) -> dict:
        """Plan knobs of a chained launch: the one-SM kernel, column blocks of at most 128 channels (two CTAs of
        different kernels share the 512 TMEM columns of an SM)."""
        t = {k: v for k, v in tune.items() if k in ("stages", "epi_n", "epi_bufs")}
        t["flags"] = rev_flag | 32 | (int(tune.get("flags", 0)) & 1)   # VSB_PLAN_ONE_SM, keep STREAM_WEIGHTS
        if tune.get("block_n", 0) and tune["block_n"] <= 128:
            t["block_n"] = tune["block_n"]
        return t

    def _conv_with_shortcut(self, c: ConvSpec, b: Act, br: ConvSpec, x: Act, out: Act, reverse: bool = False,
                            chain: Optional[dict] = None):
        """relu(BN1(branch1(x)) + BN_c(c(b)))  (resnet_helper.py:352-358) as ONE launch: the strided 1x1x1
        projection shortcut is a second K segment of the block's last conv (vsb_conv_desc.in2), so its
        [M, 4*dim_inner] result is never written to HBM and never re-read as a residual.  Both frozen
        BatchNorms must share the epilogue's per-channel scale: per channel the larger of the two scales
        stays in the fp32 epilogue and the ratio (|.| <= 1) is folded into the other weight block before
        its single bf16 rounding.  Returns False when the layer is outside the fused kernel's domain."""
        if self.dtype != VSB_BF16 or self._tune(c.key).get("fuse_shortcut", True) is False:
            return False
        if tuple(c.kernel) != (1, 1, 1) or tuple(br.kernel) != (1, 1, 1) or tuple(c.stride) != (1, 1, 1) or br.stride[0] != 1:
            return False
        if b.pitch != b.c or b.c_off or x.c_off or x.c_real != br.cin or out.c % 16:
            return False
        kchunk = 64
        # Thin layers (the Fast pathway: 8 .. 32 bottleneck channels) are restated on J-pixel groups like the
        # single convs (weights.group_conv_weight): J * b.c = 64 channels = one K chunk; the strided shortcut
        # reads groups of J * stride_w block-input pixels (zero weights on the pixels the stride skips).
        j, sw = 1, br.stride[2]
        if b.c % 64 or x.c < 64:
            dense = lambda a: a.pitch == a.c and a.c_off == 0
            j = 64 // b.c if b.c and 64 % b.c == 0 else 0
            if (j < 2 or self._tune(c.key).get("fuse_shortcut_grouped", True) is False
                    or not (dense(b) and dense(x) and dense(out)) or j * out.c > 256
                    or out.w % j or x.w % (j * sw) or out.w // j != x.w // (j * sw) or b.w != out.w):
                return False
        g = j * sw
        k2 = round_up(g * x.c if j > 1 else x.c, kchunk)

        def make():
            s_c, b_c = self._affine(c, out.c)
            s_1, b_1 = self._affine(br, out.c)
            use_c = s_c.abs() >= s_1.abs()
            s = torch.where(use_c, s_c, s_1)
            safe = torch.where(s == 0, torch.ones_like(s), s)
            r_c = torch.where(s == 0, torch.zeros_like(s), s_c / safe)
            r_1 = torch.where(s == 0, torch.zeros_like(s), s_1 / safe)
            if j == 1:
                w_c = pack_conv_weight(self._tensor(c.key + ".weight"), b.c, out.c, torch.float32).reshape(out.c, b.c)
                w_1 = pack_conv_weight(self._tensor(br.key + ".weight"), k2, out.c, torch.float32).reshape(out.c, k2)
                w = torch.cat([w_c * r_c[:, None], w_1 * r_1[:, None]], dim=1).to(self.tdt)
                return self._up(w), self._up(s), self._up(b_c + b_1)
            w_c, ngt_c, plo_c = group_conv_weight(self._tensor(c.key + ".weight"), b.c, out.c, j, 1, 0, torch.float32)
            w_1, ngt_1, plo_1 = group_conv_weight(self._tensor(br.key + ".weight"), x.c, out.c, j, sw, 0, torch.float32)
            if ngt_c != 1 or ngt_1 != 1 or plo_c or plo_1:
                return None
            w_c = w_c.reshape(j * out.c, j * b.c) * r_c.repeat(j)[:, None]
            w_1 = torch.nn.functional.pad(w_1.reshape(j * out.c, g * x.c), (0, k2 - g * x.c)) * r_1.repeat(j)[:, None]
            w = torch.cat([w_c, w_1], dim=1).to(self.tdt)
            return self._up(w), self._up(s.repeat(j)), self._up((b_c + b_1).repeat(j))
        made = self._memo((c.key, "dual", b.c, k2, out.c, j), make)
        if made is None:
            return False
        w, s, b_sum = made
        sig = self._sig(c, b, out, None, x)
        self.op_sig[c.key] = sig
        tune = {k: v for k, v in self._tune(c.key, sig).items() if k in _PLAN_KNOBS and k != "kchunk"}
        if tune.get("block_n", 0) and ((j * out.c) % tune["block_n"] or tune["block_n"] > j * out.c):
            tune.pop("block_n")
        tune["flags"] = int(tune.get("flags", 0)) | (128 if (reverse and self.alternate) else 0)
        if chain:
            if j > 1:
                return False
            tune = self._chain_tune(tune, 128 if (reverse and self.alternate) else 0)
            tune.update(chain)
        if j == 1:
            bg, xg, og, stride2 = b, x, out, br.stride
        else:
            bg = Act(b.buf, b.n, b.t, b.h, b.w // j, j * b.c, j * b.c)
            xg = Act(x.buf, x.n, x.t, x.h, x.w // g, g * x.c, g * x.c)
            og = Act(out.buf, out.n, out.t, out.h, out.w // j, j * out.c, j * out.c)
            stride2 = (br.stride[0], br.stride[1], 1)
        try:
            plan = ConvPlan(self.dtype, bg, w, og.c, c.kernel, c.stride, c.pad, None, s, b_sum, og, None, True,
                            kchunk=kchunk, x2=xg, stride2=stride2, **tune)
        except VsbError:
            if self._tune(c.key).get("fuse_shortcut") is True:
                raise
            return False
        self._keep.append(plan)
        m = out.pixels
        self.op_bytes[c.key] = 2 * (b.pixels * b.c + m * x.c + m * out.c)
        op = (c.key, plan.run, float(m) * (c.flops_per_out_pixel + br.flops_per_out_pixel))
        self.fused_shortcuts.append(br.key)
        if chain:
            return op, plan
        self.trunk_ops.append(op)
        return True

    def _stem(self, p: int, x: Act) -> Act:
        """Stem conv of pathway p.  One plan per input slot (the slots differ only in the input buffer): the op
        runs the plan of the slot selected at launch / capture time."""
        cs = self.spec.stems[p].conv
        n, t = x.n, x.t
        to, ho, wo = self._out_dims(Act(None, n, t, self.crop, self.crop, 4, 4), cs)
        y = self._alloc(n, to, ho, wo, cs.cout, pitch=cs.cout)   # dense [.., cout] (no channel padding yet)
        y.c = cs.cout
        plans = [self._stem_plan(cs, inputs[p], y, to, ho, wo) for inputs in self.input_sets]
        self._keep += plans
        es = 2 if self.dtype == VSB_BF16 else 4
        self.op_bytes[cs.key] = es * (x.pixels * 4 + y.pixels * cs.cout)
        self.trunk_ops.append((cs.key, _SlotLaunch(self, plans), float(y.pixels) * cs.flops_per_out_pixel))
        return y

    def _fused_stem(self, p: int, pitch: Optional[int]) -> Optional[Act]:
        """[1,7,7] / [5,7,7] stems with 64 output channels (Slow pathway, Slow-only, C2D / I3D): conv + BN + ReLU + max-pool as ONE
        kernel (vsb_stem_pool_*, stem_pool_sm100.cu) - the conv output never reaches HBM.  Returns the pooled
        activation, or None when the stem is outside the kernel's domain (then conv and pool run separately)."""
        st = self.spec.stems[p]
        cs = st.conv
        if self.dtype != VSB_BF16 or self.x_off != 3 or self.crop % 32:
            return None
        if str(self._tune(cs.key).get("fuse_stem", os.environ.get("VSB_FUSE_STEM", "1"))) not in ("1", "True"):
            return None
        kt = cs.kernel[0]
        if (kt not in (1, 5) or tuple(cs.kernel[1:]) != (7, 7) or tuple(cs.stride) != (1, 2, 2)
                or tuple(cs.pad) != (kt // 2, 3, 3) or cs.cout != 64
                or cs.cin != 3 or tuple(st.pool_kernel) != (1, 3, 3) or tuple(st.pool_stride) != (1, 2, 2)
                or tuple(st.pool_pad) != (0, 1, 1) or cs.has_bias):
            return None
        x0 = self.input_sets[0][p]
        n, t = x0.n, x0.t
        if kt == 5 and (t < 3 or (t > 8 and t % 8)):
            return None

        def make():
            w = self._tensor(cs.key + ".weight")                     # [64, 3, kt, 7, 7]
            order = (0,) if kt == 1 else (0, 2, 1, 4, 3)             # the kernel multiplies taps (2,1), (4,3) as pairs
            q = torch.zeros((len(order), 64, 7, 8, 4), dtype=torch.float32)
            for i, k in enumerate(order):
                q[i, :, :, :7, :3] = w[:, :, k].permute(0, 2, 3, 1)  # [co, kh, kw, c]
            return self._up(q.to(self.tdt))
        wq = self._memo((cs.key, "stem_pool"), make)
        scale, bias = self._sb(cs, 64)
        y = self._alloc(n, t, self.crop // 4, self.crop // 4, 64, pitch=pitch, min_c=16)
        plans = [ops.StemPoolPlan(inputs[p], self.x_off, wq, scale, bias, y, self.crop, kt=kt)
                 for inputs in self.input_sets]
        self._keep += plans
        name = cs.key + "+pool"
        self.op_bytes[name] = 2.0 * (x0.pixels * 4 + y.pixels * 64)
        self.trunk_ops.append((name, _SlotLaunch(self, plans, launches=2), float(n * t * (self.crop // 2) ** 2) * cs.flops_per_out_pixel))
        self.fused_stems.append(cs.key)
        return y

    def _stem_plan(self, cs: ConvSpec, x: Act, y: Act, to: int, ho: int, wo: int) -> ConvPlan:
        n, t = x.n, x.t
        wt = self._tensor(cs.key + ".weight")
        tune = {k: v for k, v in self._tune(cs.key).items() if k in _PLAN_KNOBS}
        plan = None
        algo = self._tune(cs.key).get("algo", "auto")
        if self.dtype == VSB_BF16 and (algo == "window" or (algo == "auto" and cs.kernel[0] > 1)):
            # window algorithm on J-pixel groups (2J input pixels x 4 channels = one slot).  Stems with
            # temporal taps (the Fast pathway's 5x7x7) run in its temporal-scatter mode: J = 4 -> 64-byte
            # slots, N = kt * 4 * cout per MMA, every frame window read from shared memory once.
            # (kt == 1 stems: off by default -- 32-byte slots make the TMA box loads request-rate bound.)
            xw = Act(x.buf, n, t, x.h, self.w_buf, 4, 4)
            yw = Act(y.buf, n, to, ho, wo, cs.cout, cs.cout)
            plan = self._window_plan(cs, xw, yw, None, True, wt, pad_w=cs.pad[2] - self.x_off,
                                     j=self._tune(cs.key).get("win_group", 4 if cs.kernel[0] > 1 else 2))
        if plan is not None:
            pass
        elif self.dtype == VSB_BF16:
            # J output pixels per GEMM row on the zero-bordered input rows (x' = x + x_off, so the
            # conv needs no W padding): cin' = 2J*4, cout' = J*cout, kernel (kt,7,ngt), stride (1,2,1)
            sw = cs.stride[2]
            j = self._tune(cs.key).get("group", 8 if cs.cout * 8 <= 64 else 4)
            while j > 1 and (wo % j or self.w_buf % (j * sw) or (j * cs.cout) % 16):
                j //= 2
            g = j * sw
            if (j * cs.cout) % 16 or self.w_buf % g or g * 4 < 16:
                raise VsbError("stem geometry outside the pixel-group restatement")
            wq, ngt, plo = self._memo((cs.key, "grp", 4, cs.cout, j, sw, cs.pad[2] - self.x_off), lambda: (lambda r: (
                self._up(r[0]), r[1], r[2]))(group_conv_weight(wt, 4, cs.cout, j, sw, cs.pad[2] - self.x_off, self.tdt)))
            scale, bias = self._sb(cs, cs.cout, j)
            xin = Act(x.buf, n, t, x.h, self.w_buf // g, g * 4, g * 4)
            yout = Act(y.buf, n, to, ho, wo // j, j * cs.cout, j * cs.cout)
            phi = yout.w - 1 + ngt - xin.w - plo
            if plo < 0 or xin.w + plo + phi - ngt + 1 != yout.w:
                raise VsbError("stem pixel-group geometry does not close")
            plan = ConvPlan(self.dtype, xin, wq, j * cs.cout, (cs.kernel[0], cs.kernel[1], ngt),
                            (cs.stride[0], cs.stride[1], 1), (cs.pad[0], cs.pad[1], plo),
                            (cs.pad[0], cs.pad[1], phi), scale, bias, yout, None, True, **tune)
        else:
            wp = self._memo((cs.key, "plain", 4, cs.cout), lambda: self._up(pack_conv_weight(wt, 4, cs.cout, self.tdt)))
            scale, bias = self._sb(cs, cs.cout)
            plan = ConvPlan(self.dtype, x, wp, cs.cout, cs.kernel, cs.stride, cs.pad, None, scale, bias, y, None,
                            True)
        return plan

    def _maxpool(self, name: str, x: Act, kernel, stride, pad, pitch: Optional[int] = None, min_c: int = 0) -> Act:
        to = (x.t + 2 * pad[0] - kernel[0]) // stride[0] + 1
        ho = (x.h + 2 * pad[1] - kernel[1]) // stride[1] + 1
        wo = (x.w + 2 * pad[2] - kernel[2]) // stride[2] + 1
        y = self._alloc(x.n, to, ho, wo, x.c_real, pitch=pitch, min_c=min_c)
        es = 2 if self.dtype == VSB_BF16 else 4
        self.op_bytes[name] = es * (x.pixels * x.c_real + y.pixels * y.c)
        self.trunk_ops.append((name, ops.maxpool3d_call(x, y, kernel, stride, pad, self.dtype), 0.0))
        return y

    def _fused_block(self, x: Act, blk: BlockSpec, out_pitch: Optional[int]) -> Optional[Act]:
        """Identity ResBlock (no projection shortcut, unit strides) as ONE launch of the fused bottleneck kernel
        (vsb_bottleneck_*, bottleneck_fused_sm100.cu): a's and b's outputs never reach HBM.  Thin layers are
        restated on J-pixel groups (J * d = 16 channels at least) exactly like the single convs.  Returns None
        when the block is outside the kernel's domain (then the three-launch path runs)."""
        tn = self._tune(blk.prefix)
        if self.dtype != VSB_BF16 or blk.nonlocal_ is not None:
            return None
        thin = self._thin_block(x, blk, out_pitch, tn)
        if thin is not None:
            return thin
        if blk.branch1 is not None:
            return None
        # opt-in (tune fuse_block / VSB_FUSE_BLOCK=1): measured on B200 the fused launch is at parity with the three
        # launches on the thin Fast pathway (it is bound by the TMA unit and the single MMA-issuing thread, not by HBM)
        # and, owning the whole SM, it no longer overlaps the Slow pathway's kernels: 11.53 vs 11.39 ms per step
        env = os.environ.get("VSB_FUSE_BLOCK", "0")   # "1" = every eligible block, or a list of block-name prefixes
        env_on = env == "1" or any(blk.prefix.startswith(pre) for pre in env.split(",") if len(pre) > 1)
        if str(tn.get("fuse_block", "1" if env_on else "0")) not in ("1", "True"):
            return None
        a, b, c = blk.a, blk.b, blk.c
        if (tuple(a.kernel[1:]) != (1, 1) or a.kernel[0] not in (1, 3) or tuple(b.kernel) != (1, 3, 3)
                or tuple(c.kernel) != (1, 1, 1) or any(s != 1 for cs in (a, b, c) for s in cs.stride)
                or tuple(a.pad) != (a.kernel[0] // 2, 0, 0) or tuple(b.pad) != (0, 1, 1) or tuple(c.pad) != (0, 0, 0)
                or c.cout != a.cin or x.c_real != a.cin or x.c_off or x.pitch != x.c):
            return None
        if out_pitch is not None and out_pitch != self._store(c.cout):
            return None   # channel-slice outputs (the slow pathway's concat buffers) stay on the three-launch path
        d_store = self._store(a.cout)
        j = 1
        while j * d_store < 16:
            j *= 2
        if x.w % j or (j * d_store) % 16 or j * d_store > 64 or (j * x.c) % 16 or j * x.c > 256:
            return None
        kt = a.kernel[0]

        def make():
            wa, _, _ = group_conv_weight(self._tensor(a.key + ".weight"), x.c, d_store, j, 1, 0, self.tdt)
            wb, ngt, plo = group_conv_weight(self._tensor(b.key + ".weight"), d_store, d_store, j, 1, 1, self.tdt)
            wc, _, _ = group_conv_weight(self._tensor(c.key + ".weight"), d_store, x.c, j, 1, 0, self.tdt)
            if ngt != 3 or plo != 1:
                return None
            return self._up(wa), self._up(wb), self._up(wc)
        w = self._memo((blk.prefix, "fused", x.c, d_store, j), make)
        if w is None:
            return None
        sa, ba = self._sb(a, d_store, j)
        sb_, bb = self._sb(b, d_store, j)
        sc, bc = self._sb(c, x.c, j)
        y = self._alloc(x.n, x.t, x.h, x.w, c.cout)
        xg = Act(x.buf, x.n, x.t, x.h, x.w // j, j * x.c, j * x.c)
        yg = Act(y.buf, y.n, y.t, y.h, y.w // j, j * y.c, j * y.c)
        try:
            plan = BottleneckPlan(xg, yg, j * d_store, kt, w[0], w[1], w[2], sa, ba, sb_, bb, sc, bc)
        except VsbError:
            if tn.get("fuse_block") is True:
                raise
            self._free(y)
            return None
        self._keep.append(plan)
        name = blk.prefix + ".fused_abc"
        self.op_bytes[name] = 2.0 * (x.pixels * x.c + y.pixels * y.c)
        self.trunk_ops.append((name, plan.run, float(x.pixels) * (a.flops_per_out_pixel + b.flops_per_out_pixel
                                                                  + c.flops_per_out_pixel)))
        self.fused_blocks.append(blk.prefix)
        self._free(x)
        return y

    def _thin_block(self, x: Act, blk: BlockSpec, out_pitch: Optional[int], tn: dict) -> Optional[Act]:
        """ResBlock of a THIN stage (stored bottleneck width 8 or 16, block width 4x that: the Fast pathway's
        res2 / res3) with unit strides as ONE launch of the warp-MMA walk kernel (vsb_bottleneck_* algo 1,
        bottleneck_thin_sm100.cu): ungrouped weights, every frame of a row strip fetched once, a's and b's outputs
        never reach HBM.  Identity blocks, and - for width 8 - block 0 with its 1x1x1 projection shortcut folded
        into conv c as a second K half (same folding as _conv_with_shortcut).  On by default (VSB_THIN_BLOCK=0 /
        tune {"thin_block": False} keeps the three launches)."""
        if str(tn.get("thin_block", os.environ.get("VSB_THIN_BLOCK", "1"))) not in ("1", "True"):
            return None
        a, b, c, br = blk.a, blk.b, blk.c, blk.branch1
        d_store = self._store(a.cout)
        c_out = self._store(c.cout)
        if d_store == 16 and str(tn.get("thin_d16", os.environ.get("VSB_THIN_D16", "1"))) not in ("1", "True"):
            return None
        if (d_store not in (8, 16) or c_out != 4 * d_store or x.c_real != a.cin or x.c_off or x.pitch != x.c
                or tuple(a.kernel[1:]) != (1, 1) or a.kernel[0] not in (1, 3) or tuple(b.kernel) != (1, 3, 3)
                or tuple(c.kernel) != (1, 1, 1) or any(s != 1 for cs in (a, b, c) for s in cs.stride)
                or tuple(a.pad) != (a.kernel[0] // 2, 0, 0) or tuple(b.pad) != (0, 1, 1) or tuple(c.pad) != (0, 0, 0)):
            return None
        if br is None:
            if x.c != c_out or c.cout != a.cin:
                return None
        else:
            if (str(tn.get("thin_proj", os.environ.get("VSB_THIN_PROJ", "1"))) not in ("1", "True")
                    or d_store != 8 or x.c not in (8, 16) or x.c_real != 8 or br.cin != a.cin or br.cout != c.cout
                    or tuple(br.kernel) != (1, 1, 1) or any(s != 1 for s in br.stride) or any(br.pad)):
                return None
        if out_pitch is not None and out_pitch != c_out:
            return None

        def make():
            cin = x.c if br is None else 8   # projection blocks read the 8 real channels of a (possibly 16-wide) pixel
            wa = pack_conv_weight(self._tensor(a.key + ".weight"), cin, d_store, self.tdt)
            wb = pack_conv_weight(self._tensor(b.key + ".weight"), d_store, d_store, self.tdt)
            if br is None:
                wc = pack_conv_weight(self._tensor(c.key + ".weight"), d_store, c_out, self.tdt)
                s_, b_ = self._affine(c, c_out)
            else:
                # one accumulator for conv c and the shortcut: the two BatchNorm scales become ratios to the
                # larger one (the common epilogue scale), the biases add
                s_c, b_c = self._affine(c, c_out)
                s_1, b_1 = self._affine(br, c_out)
                s_ = torch.where(s_c.abs() >= s_1.abs(), s_c, s_1)
                safe = torch.where(s_ == 0, torch.ones_like(s_), s_)
                r_c = torch.where(s_ == 0, torch.zeros_like(s_), s_c / safe)
                r_1 = torch.where(s_ == 0, torch.zeros_like(s_), s_1 / safe)
                w_c = pack_conv_weight(self._tensor(c.key + ".weight"), d_store, c_out, torch.float32).reshape(c_out, d_store)
                w_1 = pack_conv_weight(self._tensor(br.key + ".weight"), cin, c_out, torch.float32).reshape(c_out, cin)
                wc = torch.cat([w_c * r_c[:, None], w_1 * r_1[:, None]], dim=1).to(self.tdt)
                b_ = b_c + b_1
            return self._up(wa), self._up(wb), self._up(wc), self._up(s_), self._up(b_)
        w = self._memo((blk.prefix, "thin", x.c, d_store, br is not None), make)
        sa, ba = self._sb(a, d_store)
        sb_, bb = self._sb(b, d_store)
        y = self._alloc(x.n, x.t, x.h, x.w, c.cout)
        try:
            plan = BottleneckPlan(x, y, d_store, a.kernel[0], w[0], w[1], w[2], sa, ba, sb_, bb, w[3], w[4], algo=1,
                                  cin=(8 if br is not None else 0),
                                  stages=int(tn.get("thin_slots", 0)), walk_len=int(tn.get("thin_rows", 0)))
        except VsbError:
            if tn.get("thin_block") is True:
                raise
            self._free(y)
            return None
        self._keep.append(plan)
        name = blk.prefix + ".thin_abc"
        self.op_bytes[name] = 2.0 * (x.pixels * x.c + y.pixels * y.c)
        flops = a.flops_per_out_pixel + b.flops_per_out_pixel + c.flops_per_out_pixel
        if br is not None:
            flops += br.flops_per_out_pixel
            self.fused_shortcuts.append(br.key)
        self.trunk_ops.append((name, plan.run, float(x.pixels) * flops))
        self.fused_blocks.append(blk.prefix)
        self._free(x)
        return y

    def _chain_ok(self, blk: BlockSpec, nxt: Optional[BlockSpec], out_pitch: Optional[int]) -> bool:
        """Tile-granular chaining of this block's last conv with the next block's first (vsb_conv_desc.tile_signal /
        tile_wait): the next `a` must be a 1x1x1 stride-1 conv of at most 128 stored channels over this block's
        dense output.  Opt-in (VSB_CHAIN=1 / tune {"*": {"chain": True}}): measured on B200 the chained pairs
        lose to two full-width launches (12.1 vs 11.2 ms per SF50 batch-64 step, profiles/r02_ab_knobs.txt):
        the consumer's CTAs take shared memory and TMEM from the producer, and the tile it waits for is
        ~3000 clk of store latency away, so the SM slots it owns mostly spin."""
        if nxt is None or self.dtype != VSB_BF16 or blk.nonlocal_ is not None or out_pitch is not None:
            return False
        if str(self._tune(blk.prefix).get("chain", os.environ.get("VSB_CHAIN", "0"))) not in ("1", "True"):
            return False
        a = nxt.a
        if tuple(a.kernel) != (1, 1, 1) or tuple(a.stride) != (1, 1, 1) or tuple(a.pad) != (0, 0, 0):
            return False
        if tuple(blk.c.kernel) != (1, 1, 1) or tuple(blk.c.stride) != (1, 1, 1):
            return False
        # pixel-grouped (thin) layers stay on their own plans: _group_factor groups inputs below 64 channels
        return (self._store(a.cout) <= 128 and self._store(a.cout) >= 64 and self._store(blk.c.cin) >= 64
                and self._store(blk.c.cout) >= 64 and self._store(blk.c.cout) % 128 == 0)

    def _block(self, x: Act, blk: BlockSpec, out_pitch: Optional[int], nxt: Optional[BlockSpec] = None,
               a_pre: Optional[Act] = None) -> Tuple[Act, Optional[Act]]:
        """One ResBlock.  Returns (block output, a_next): a_next is the output of the NEXT block's conv `a` when
        it was chained to this block's last conv (then the next call gets it as a_pre and skips its own `a`)."""
        n = x.n
        pw = self._pools.index(self._pool)
        d = self._rev[pw]              # a and c walk in direction d, b against it; the next block flips
        self._rev[pw] = not d
        if a_pre is None:
            fused = self._fused_block(x, blk, out_pitch)
            if fused is not None:
                return fused, None
            ta, ha, wa = self._out_dims(x, blk.a)
            a = self._alloc(n, ta, ha, wa, blk.a.cout)
            self._conv(blk.a, x, a, reverse=d)
        else:
            a = a_pre
        tb, hb, wb = self._out_dims(a, blk.b)
        b = self._alloc(n, tb, hb, wb, blk.b.cout)
        self._conv(blk.b, a, b, reverse=not d)
        self._free(a)
        # chained pair: the next block's `a` runs next to this block's `c`, tile by tile (its buffer is taken
        # BEFORE b is released: it is written while c still reads b)
        chain_on = self._chain_ok(blk, nxt, out_pitch)
        flags = a_next = None
        if chain_on:
            flags = torch.zeros((n * tb * hb * wb + 127) // 128, dtype=torch.int32, device=self.device)
            a_next = self._alloc(n, tb, hb, wb, nxt.a.cout)
        # CTA slots (two per SM) are split between producer and consumer: the producer moves ~2x the bytes
        slots = 2 * torch.cuda.get_device_properties(self.device).multi_processor_count
        share = float(self._tune(blk.prefix).get("chain_share", os.environ.get("VSB_CHAIN_SHARE", "0.72")))
        g_c = max(8, int(slots * share) // 8 * 8)
        chain = {"tile_signal": flags, "grid_limit": g_c} if chain_on else None
        sc = None
        y = None
        c_op = None
        if blk.branch1 is not None:
            y = self._alloc(n, tb, hb, wb, blk.c.cout, pitch=out_pitch)
            r = self._conv_with_shortcut(blk.c, b, blk.branch1, x, y, reverse=d, chain=chain)
            if r:
                res = None
                if chain_on:
                    c_op = r
            else:
                self._free(y)
                y = None
                sc = self._alloc(n, tb, hb, wb, blk.branch1.cout)
                self._conv(blk.branch1, x, sc, reverse=d)
                res = sc
        else:
            res = x
        if y is None:
            y = self._alloc(n, tb, hb, wb, blk.c.cout, pitch=out_pitch)
        # relu(shortcut + BN(c(.)))   resnet_helper.py:352-358
        if res is not None:
            r = self._conv(blk.c, b, y, residual=res, relu=True, reverse=d, chain=chain)
            if chain_on:
                c_op = r
        if chain_on:
            (c_name, c_run, c_flops), c_plan = c_op
            count = 8 * (self._store(blk.c.cout) // c_plan.info()["block_n"])
            (a_name, a_run, a_flops), _ = self._conv(nxt.a, y, a_next, reverse=d,
                                                     chain={"tile_wait": flags, "tile_wait_count": count,
                                                            "grid_limit": slots - c_plan.info()["grid"]})
            # one trunk op: the consumer must directly follow its producer on the stream (it consumes the tile
            # counters the producer raises and resets them)
            self.op_bytes[c_name + "+" + a_name] = self.op_bytes[c_name] + self.op_bytes[a_name]
            self.trunk_ops.append((c_name + "+" + a_name, lambda: (c_run(), a_run()), c_flops + a_flops))
            self.chained.append((c_name, a_name, c_run, a_run, c_plan))
        self._free(b)
        if sc is not None:
            self._free(sc)
        self._free(x)
        return y, a_next

    def _nonlocal(self, x: Act, nl: NonlocalSpec, out_pitch: Optional[int]) -> Act:
        n = x.n
        if nl.pool is not None:
            xp = self._maxpool(nl.prefix + ".pool", x, nl.pool, nl.pool, (0, 0, 0))
        else:
            xp = x
        # tensor-core route: theta and phi are written (and multiplied) as IEEE half - they enter the exponent of the
        # softmax, where their bf16 rounding would be a relative error of every attention weight
        half = (self._nl_gemm_ok(nl, self._store(nl.dim_inner), xp.t * xp.h * xp.w)
                and self._tune(nl.prefix).get("nl_f16", True) is not False)
        theta = self._alloc(n, x.t, x.h, x.w, nl.dim_inner)
        self._conv(nl.theta, x, theta, out_f16=half)
        phi = self._alloc(n, xp.t, xp.h, xp.w, nl.dim_inner)
        if self.dtype == VSB_BF16:
            # the tensor-core route reads phi_i as a weight matrix of round_up(keys, 16) rows: slack after the last clip
            self._free(phi)
            phi = Act(self._pool.take(phi.pixels * phi.pitch + 16 * phi.pitch), n, xp.t, xp.h, xp.w, phi.c, phi.pitch, 0,
                      phi.c_real)
        g = self._alloc(n, xp.t, xp.h, xp.w, nl.dim_inner)
        self._conv(nl.phi, xp, phi, out_f16=half)
        self._conv(nl.g, xp, g)
        if xp is not x:
            self._free(xp)
        att = self._alloc(n, x.t, x.h, x.w, nl.dim_inner)
        tq, tk = x.t * x.h * x.w, xp.t * xp.h * xp.w
        if not self._nonlocal_gemms(nl, theta, phi, g, att, tq, tk, half):
            self.trunk_ops.append((nl.prefix + ".attention",
                                   ops.nonlocal_attention_call(theta, phi, g, att, nl.softmax, self.dtype),
                                   4.0 * n * tq * tk * nl.dim_inner))
        y = self._alloc(n, x.t, x.h, x.w, nl.dim, pitch=out_pitch)
        # x + BN(conv_out(.)), no ReLU   nonlocal_helper.py:145-148
        self._conv(nl.out, att, y, residual=x, relu=False)
        for a in (theta, phi, g, att, x):
            self._free(a)
        return y

    def _nl_gemm_ok(self, nl: NonlocalSpec, c: int, tk: int) -> bool:
        """Is this non-local block inside the tensor-core route (_nonlocal_gemms)?"""
        if self.dtype != VSB_BF16 or self._tune(nl.prefix).get("nl_gemm", True) is False:
            return False
        if c % 64 or tk % 2:
            return False
        if not nl.softmax:
            # "dot_product" instantiation: no VidSitu config places such a block (Kinetics_c2_SLOW_8x8_R50.yaml has
            # none), so the batched route was never measured with it - it stays on the CUDA-core kernel
            return False
        # few keys (small crops): bf16 probabilities do not average out (max error 0.0100 vs 0.0070 of the feature
        # scale on the crop-64 fixture) and the CUDA-core kernel costs nothing at this size
        return tk >= 128

    def _nonlocal_gemms(self, nl: NonlocalSpec, theta: Act, phi: Act, g: Act, att: Act, tq: int, tk: int,
                        half: bool = False) -> bool:
        """The two einsums of Nonlocal.forward (nonlocal_helper.py:123-141) on the tensor cores: per clip,
        scores_i = theta_i . phi_i^T is a 1x1x1 conv of theta_i whose WEIGHTS are phi_i ([keys, c], the softmax
        scale c^-1/2 - or 1/keys - in the epilogue), vsb_score_rows normalises the rows and zeroes the K padding,
        and out_i = P_i . g_i is a 1x1x1 conv of P_i whose weights are g_i^T ([c, keys], vsb_transpose_pad).
        `half`: theta and phi hold IEEE half (vsb_conv_desc.in_f16).
        Returns False when the block is outside this route (fp32 mode, odd shapes): the CUDA-core kernel runs."""
        c = theta.c
        dense = all(a.pitch == a.c and a.c_off == 0 for a in (theta, phi, g, att))
        if not self._nl_gemm_ok(nl, c, tk) or not dense or phi.c != c or g.c != c or att.c != c:
            if half:
                raise VsbError(f"{nl.prefix}: theta / phi were planned as half but the block left the tensor-core route")
            return False
        n = theta.n
        kc, kp = round_up(tk, 16), round_up(tk, 64)
        scores = self._pool.take(n * tq * kp)
        g_t = self._pool.take(n * c * kp)
        dev = self.device
        s_scale = torch.full((kc,), float(c) ** -0.5 if nl.softmax else 1.0 / tk, dtype=torch.float32, device=dev)
        s_bias = torch.zeros(kc, dtype=torch.float32, device=dev)
        y_scale = torch.ones(c, dtype=torch.float32, device=dev)
        y_bias = torch.zeros(c, dtype=torch.float32, device=dev)
        t, h, w = theta.t, theta.h, theta.w
        # one launch per product: clip i's weight matrix starts tk (resp. c) rows after clip i-1's
        plan_s = ConvPlan(self.dtype, theta, phi.buf, kc, (1, 1, 1), (1, 1, 1), (0, 0, 0), None, s_scale, s_bias,
                          Act(scores, n, t, h, w, kc, kp), None, False, out_f16=True,   # half scores: 11 mantissa bits into exp()
                          wgt_clip_rows=tk, in_f16=half)
        plan_y = ConvPlan(self.dtype, Act(scores, n, t, h, w, kp, kp), g_t, c, (1, 1, 1), (1, 1, 1), (0, 0, 0), None,
                          y_scale, y_bias, att, None, False, wgt_clip_rows=c)
        self._keep += [plan_s, plan_y, s_scale, s_bias, y_scale, y_bias, scores, g_t]
        run_scores, run_out = plan_s.run, plan_y.run

        fl = 2.0 * n * tq * tk * c
        self.trunk_ops.append((nl.prefix + ".g_transpose", ops.transpose_pad_call(g, g_t, kp), 0.0))
        self.trunk_ops.append((nl.prefix + ".scores", run_scores, fl))
        self.trunk_ops.append((nl.prefix + ".softmax",
                               ops.score_rows_call(scores, n * tq, tk, kp, kp, nl.softmax), 0.0))
        self.trunk_ops.append((nl.prefix + ".attention_out", run_out, fl))
        self._pool.give(scores)
        self._pool.give(g_t)
        self.op_bytes[nl.prefix + ".scores"] = 2.0 * n * (tq * c + tk * c + tq * kc)
        self.op_bytes[nl.prefix + ".attention_out"] = 2.0 * n * (tq * kp + tk * c + tq * c)
        return True

    def _stage(self, x: Act, blocks: List[BlockSpec], out_pitch: Optional[int]) -> Act:
        a_pre = None
        for i, blk in enumerate(blocks):
            last = i == len(blocks) - 1
            x, a_pre = self._block(x, blk, out_pitch if (last and blk.nonlocal_ is None) else None,
                                   nxt=None if last else blocks[i + 1], a_pre=a_pre)
            if blk.nonlocal_ is not None:
                x = self._nonlocal(x, blk.nonlocal_, out_pitch if last else None)
        return x

    def _build_trunk(self) -> List[Act]:
        spec = self.spec
        npw = spec.num_pathways
        xs: List[Act] = []
        # s1: stem conv + BN + ReLU + max-pool per pathway (stem_helper.py:173-178)
        for p in range(npw):
            self._pool = self._pools[p]
            st = spec.stems[p]
            fuse = spec.fuses[0] if p == 0 else None
            pitch = self._store(st.conv.cout) + fuse.cout if fuse is not None else None
            fused = self._fused_stem(p, pitch)
            if fused is not None:
                xs.append(fused)
                continue
            y = self._stem(p, self.inputs[p])
            # (16 channels at least: the lateral conv reads this tensor un-grouped, 16 = one MMA K step)
            xs.append(self._maxpool(f"s1.pathway{p}_stem.pool_layer", y, st.pool_kernel, st.pool_stride, st.pool_pad,
                                    pitch, min_c=16))
            self._free(y)
        for si in range(4):
            fuse = spec.fuses[si]
            if fuse is not None:
                xs[0] = self._lateral(fuse, xs[0], xs[1])
            stage = spec.stages[si]
            nxt_fuse = spec.fuses[si + 1] if si + 1 < 4 else None
            for p in range(npw):
                self._pool = self._pools[p]
                out_c = stage[p][-1].c.cout
                pitch = self._store(out_c) + nxt_fuse.cout if (nxt_fuse is not None and p == 0) else None
                xs[p] = self._stage(xs[p], stage[p], pitch)
            if si == 0:
                # pathway{p}_pool after res2 (mdl_sf_base.py:26-28): identity [1,1,1] pools are skipped
                for p in range(npw):
                    self._pool = self._pools[p]
                    k = spec.pool1[p]
                    if any(v != 1 for v in k):
                        if nxt_fuse is not None and p == 0:
                            raise VsbError("pool1 on a pathway that carries a lateral concat is not planned")
                        y = self._maxpool(f"pathway{p}_pool", xs[p], k, k, (0, 0, 0))
                        self._free(xs[p])
                        xs[p] = y
        return xs

    def _lateral(self, fuse: ConvSpec, x_s: Act, x_f: Act) -> Act:
        """FuseFastToSlow: conv(7x1x1, stride alpha)+BN+ReLU written into channels [C_s, C_s+C_fuse)
        of the slow tensor; returns the widened slow activation (the torch.cat result)."""
        c_s = x_s.c
        if x_s.pitch != c_s + fuse.cout or c_s != x_s.c_real:
            raise VsbError("slow activation was not allocated with room for the lateral channels")
        dst = Act(x_s.buf, x_s.n, x_s.t, x_s.h, x_s.w, fuse.cout, x_s.pitch, c_off=c_s, c_real=fuse.cout)
        to, ho, wo = self._out_dims(x_f, fuse)
        if (to, ho, wo) != (x_s.t, x_s.h, x_s.w):
            raise VsbError("lateral conv output does not match the slow pathway extent")
        self._conv(fuse, x_f, dst)
        return Act(x_s.buf, x_s.n, x_s.t, x_s.h, x_s.w, x_s.pitch, x_s.pitch, 0, x_s.pitch)

    # --------------------------------------------------------------------------- running
    @property
    def inputs(self) -> List[Act]:
        """Input activations of the current slot."""
        return self.input_sets[self._slot]

    def select_slot(self, slot: int) -> None:
        if not 0 <= slot < len(self.input_sets):
            raise VsbError(f"input slot {slot} outside [0, {len(self.input_sets)})")
        self._slot = slot

    def load_frames(self, frames: torch.Tensor, slot: Optional[int] = None) -> None:
        """uint8 [n, T, H, W, 3] frames of the fast/single pathway window (dat_loader.py:474-476), packed into
        input slot `slot` (default: the current one)."""
        if slot is not None:
            self.select_slot(slot)
        spec = self.spec
        if frames.device != self.device:
            raise VsbError(f"frames are on {frames.device}, the engine runs on {self.device}")
        if (frames.shape[0] != self.n or frames.shape[1] != spec.num_frames
                or tuple(frames.shape[2:4]) != (self.crop, self.crop)):
            raise VsbError(f"expected frames [{self.n}, {spec.num_frames}, {self.crop}, {self.crop}, 3], got "
                           f"{tuple(frames.shape)}")
        if spec.num_pathways == 2:
            ops.pack_frames(frames, self.slow_idx, spec.mean, spec.std, self.inputs[0], self.dtype,
                            spec.reverse_input_channel, self.x_off)
            ops.pack_frames(frames, self.fast_idx, spec.mean, spec.std, self.inputs[1], self.dtype,
                            spec.reverse_input_channel, self.x_off)
        else:
            ops.pack_frames(frames, self.fast_idx, spec.mean, spec.std, self.inputs[0], self.dtype,
                            spec.reverse_input_channel, self.x_off)

    def load_videos(self, videos: torch.Tensor) -> None:
        """uint8 [n/5, F, H, W, 3] whole videos (F = 300 extracted frames, dat_loader.py:454-458) resident
        on the GPU: the five event windows of every video (events.event_frame_indices) are gathered by
        the pack kernel straight from the video tensor.  Clips are laid out EVENT-major in the batch
        (clip = event * n_videos + video): one pack launch per event and pathway."""
        from .events import EVENTS_PER_VIDEO, event_frame_indices
        spec = self.spec
        if (videos.dim() != 5 or videos.shape[0] * EVENTS_PER_VIDEO != self.n
                or tuple(videos.shape[2:4]) != (self.crop, self.crop)):
            raise VsbError(f"expected videos [{self.n // EVENTS_PER_VIDEO}, F, {self.crop}, {self.crop}, 3] for "
                           f"an engine of {self.n} clips")
        n_vid, f = int(videos.shape[0]), int(videos.shape[1])
        windows = event_frame_indices(spec.num_frames, spec.sampling_rate, spec.target_fps, f)
        for ev, win in enumerate(windows):
            for p, a in enumerate(self.inputs):
                idx = [win[i] for i in self.slow_idx] if (spec.num_pathways == 2 and p == 0) else win
                per_clip = a.t * a.h * a.w * a.pitch
                sub = Act(a.buf[ev * n_vid * per_clip: (ev + 1) * n_vid * per_clip], n_vid, a.t, a.h, a.w, 4, 4,
                          c_real=3)
                ops.pack_frames(videos, idx, spec.mean, spec.std, sub, self.dtype, spec.reverse_input_channel,
                                self.x_off)

    def load_ncthw(self, xs: Sequence[torch.Tensor]) -> None:
        """The reference's already-normalised fp32 [n, 3, T_p, H, W] pathway tensors."""
        if len(xs) != len(self.inputs):
            raise VsbError(f"expected {len(self.inputs)} pathway tensors")
        for x, a in zip(xs, self.inputs):
            if tuple(x.shape) != (self.n, 3, a.t, a.h, self.crop):
                raise VsbError(f"pathway tensor {tuple(x.shape)} != {(self.n, 3, a.t, a.h, self.crop)}")
            ops.ncthw_to_act(x.contiguous(), a, self.dtype, self.x_off)

    def run_trunk(self) -> None:
        for name, fn, _ in self.trunk_ops:
            with self._range(name):
                fn()

    def run_head(self) -> None:
        for name, fn, _ in self.head_ops:
            with self._range(name):
                fn()

    @staticmethod
    def _range(name: str):
        """NVTX range per launch (VSB_NVTX=1): the reference's layer names on the profiler timeline."""
        if os.environ.get("VSB_NVTX") == "1":
            return torch.cuda.nvtx.range(name)
        import contextlib
        return contextlib.nullcontext()

    @staticmethod
    def _op_stream(name: str) -> int:
        """0 = slow-pathway stream, 1 = fast-pathway stream (which also runs the lateral convs)."""
        return 1 if ("pathway1" in name or "_fuse" in name) else 0

    def _guard(self):
        """Launches take torch's CURRENT stream and encode TMA maps in the CURRENT context: pin both to the engine's
        device (a model on cuda:1 while cuda:0 is current would otherwise launch on the wrong device)."""
        return torch.cuda.device(self.device)

    def run(self) -> None:
        with self._guard():
            self._run()

    def _schedule(self):
        """The forward as a sequence of ("op", lane, name, launch) and ("sync", from_lane, to_lane) items - consumed
        by the eager / torch-graph run (_run: lanes are torch streams) and by build_program (lanes of a C clip
        program).  SlowFast nets run their two pathways on two lanes (fork after the inputs are packed, join before
        the projection head): the pathways only meet at the lateral convs (video_model_builder.py:124-131), which
        wait for the slow stage they write into and are waited for by the next slow stage.  Order inside a lane
        is the program order."""
        if not self.two_streams:
            for name, fn, _ in self.trunk_ops + self.head_ops:
                yield ("op", 0, name, fn)
            return
        yield ("sync", 0, 1)
        prev_fuse = False
        for name, fn, _ in self.trunk_ops:
            sid = self._op_stream(name)
            is_fuse = "_fuse" in name
            if is_fuse and not self.early_lateral:
                yield ("sync", 0, 1)          # pooled concat buffer: wait until the slow stage has made it live
            elif sid == 0 and prev_fuse:
                yield ("sync", 1, 0)          # the next slow stage reads the concatenated channels
            yield ("op", sid, name, fn)
            if sid == 0:
                prev_fuse = False
            elif is_fuse:
                prev_fuse = True
        joined = False
        for name, fn, _ in self.head_ops:
            if name.startswith("proj_head") and not joined:
                yield ("sync", 1, 0)
                joined = True
            yield ("op", 1 if (self._op_stream(name) == 1 and not joined) else 0, name, fn)
        if not joined:
            yield ("sync", 1, 0)

    def _run(self) -> None:
        """One forward, launch by launch from Python (eager runs, per-op timing, torch graph capture)."""
        if not self.two_streams:
            for _, _, name, fn in self._schedule():
                with self._range(name):
                    fn()
            return
        origin = torch.cuda.current_stream()
        prio = int(self.tune.get("*", {}).get("prio", os.environ.get("VSB_STREAM_PRIO", "0")))
        if self._side_stream is None:
            # prio 1: the Slow pathway (the critical path) runs on a high-priority stream, the Fast pathway
            # fills in; prio 2: the other way round; 0: Slow on the caller's stream, both default priority
            self._side_stream = torch.cuda.Stream(self.device, priority=-1 if prio == 2 else 0)
            self._main_stream = torch.cuda.Stream(self.device, priority=-1) if prio == 1 else None
        main = self._main_stream if self._main_stream is not None else origin
        if main is not origin:
            main.wait_stream(origin)
        streams = (main, self._side_stream)
        for item in self._schedule():
            if item[0] == "sync":
                ev = torch.cuda.Event()
                ev.record(streams[item[1]])
                streams[item[2]].wait_event(ev)
            else:
                _, lane, name, fn = item
                with torch.cuda.stream(streams[lane]), self._range(name):
                    fn()
        if main is not origin:
            origin.wait_stream(main)

    # ------------------------------------------------------------------ clip programs (C ABI v7)
    @staticmethod
    def _emit(fn, prog: "ops.Program", lane: int, name: str) -> None:
        target = fn if hasattr(fn, "emit") else getattr(fn, "__self__", None)
        if target is None or not hasattr(target, "emit"):
            raise VsbError(f"{name}: this launch cannot be recorded into a clip program")
        target.emit(prog, lane, name)

    def _pack_calls(self, frames: torch.Tensor) -> List[Tuple[str, "ops.Call"]]:
        spec = self.spec
        idxs = [self.slow_idx, self.fast_idx] if spec.num_pathways == 2 else [self.fast_idx]
        return [(f"pack.pathway{p}", ops.pack_frames_call(frames, idx, spec.mean, spec.std, self.inputs[p], self.dtype,
                                                          spec.reverse_input_channel, self.x_off))
                for p, idx in enumerate(idxs)]

    def build_program(self, slot: Optional[int] = None, with_pack: bool = False) -> "ops.Program":
        """The forward of this engine as ONE C handle (include/vidsitu_b200.h: vsb_program_*): every launch of
        _schedule() recorded with its lane, plans borrowed from the engine.  `with_pack`: the program starts with
        the pack launches reading the engine's static uint8 frame buffer (self.frames_in, region "frames")."""
        with self._guard():
            prog = ops.Program()
            keep = self._slot
            if slot is not None:
                self.select_slot(slot)
            try:
                if with_pack:
                    if self.frames_in is None:
                        self.frames_in = torch.zeros((self.n, self.spec.num_frames, self.crop, self.crop, 3),
                                                     dtype=torch.uint8, device=self.device)
                    for name, call in self._pack_calls(self.frames_in):
                        call.emit(prog, 0, name)
                for item in self._schedule():
                    if item[0] == "sync":
                        prog.sync(item[1], item[2])
                    else:
                        self._emit(item[3], prog, item[1], item[2])
            finally:
                self._slot = keep
            prog.keep(self)
            return prog

    def export_program(self, path: str) -> "ops.Program":
        """Save the whole forward (pack -> trunk -> head) as a relocatable program file that any host process can
        load and run through the C ABI alone (vsb_program_load / vsb_program_region / vsb_program_run): packed
        weights and folded BatchNorm travel as constant regions, activations as sized scratch regions; the host
        writes uint8 frames into region "frames" and reads fp32 region "feats" (and "logits")."""
        prog = self.build_program(slot=0, with_pack=True)
        scratch: Dict[int, Tuple[str, torch.Tensor]] = {}

        def reg(name: str, t: Optional[torch.Tensor]) -> None:
            if t is not None:
                scratch.setdefault(t.untyped_storage().data_ptr(), (name, t))

        reg("frames", self.frames_in)
        reg("feats", self.feats)
        reg("logits", self.logits)
        reg("hidden", getattr(self, "hidden", None))
        for p, a in enumerate(self.input_sets[0]):
            reg(f"input.pathway{p}", a.buf)
        for i, b in enumerate(b for pool in self._pools for b in pool.all):
            reg(f"act{i}", b)
        for i, b in enumerate(self._dedicated):
            reg(f"concat{i}", b)
        consts: Dict[int, torch.Tensor] = {}
        launches = [fn for item in self._schedule() if item[0] == "op" for fn in (item[3],)]
        launches += [c for _, c in self._pack_calls(self.frames_in)]
        for fn in launches:
            owner = fn if hasattr(fn, "_keep") else getattr(fn, "__self__", None)
            for t in getattr(owner, "_keep", ()):
                if isinstance(t, torch.Tensor) and t.is_cuda:
                    sp = t.untyped_storage().data_ptr()
                    if sp not in scratch:
                        consts.setdefault(sp, t)

        def whole(t: torch.Tensor) -> torch.Tensor:
            st = t.untyped_storage()
            return torch.empty(0, dtype=torch.uint8, device=t.device).set_(st, 0, (st.nbytes(),), (1,))

        for name, t in scratch.values():
            prog.add_region(name, whole(t), ops.Program.SCRATCH)
        for i, t in enumerate(consts.values()):
            prog.add_region(f"const{i}", whole(t), ops.Program.CONST)
        prog.save(path)
        return prog

    def _use_program(self) -> bool:
        return self.replay_mode == "program" and not self.chained

    def capture(self) -> None:
        """Record the forward once for replay(): as clip programs captured inside the library (default), or as
        torch CUDA graphs of the Python launch loop (replay_mode "torch")."""
        with self._guard():
            if self._use_program():
                if self._programs is None:
                    progs = [self.build_program(slot=i) for i in range(len(self.input_sets))]
                    for p in progs:
                        p.capture()
                    self._programs = progs
            elif self._graph is None:
                self._capture()

    def _capture(self) -> None:
        """Capture trunk + head once into a CUDA graph (inputs/outputs are static buffers)."""
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._run()          # warm-up outside capture (lazy module loading, func attributes)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        keep = self._slot
        self._graph = []
        for slot in range(len(self.input_sets)):   # one graph per input slot (only the stem launches differ)
            self._slot = slot
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._run()
            self._graph.append(g)
        self._slot = keep

    def replay(self, slot: Optional[int] = None) -> None:
        if slot is not None:
            self.select_slot(slot)
        self.capture()
        with self._guard():
            if self._use_program():
                self._programs[self._slot].run()     # ONE C call: vsb_program_run
            else:
                self._graph[self._slot].replay()

    @property
    def num_launches(self) -> int:
        return sum(getattr(fn, "launches", 1) for _, fn, _ in self.trunk_ops + self.head_ops)

    @property
    def conv_flops(self) -> float:
        """Algorithmic conv FLOPs (2*MAC on the reference's un-padded shapes) of one run."""
        return sum(f for name, _, f in self.trunk_ops
                   if not name.endswith((".attention", ".scores", ".attention_out")))

    def features_ncthw(self) -> List[torch.Tensor]:
        return [ops.act_to_ncthw(x, self.dtype) for x in self.trunk_out]

    def time_ops(self, iters: int = 3) -> List[Tuple[str, float, float]]:
        """Per-launch device time (ms, best of `iters`) -- tuning / profiling aid."""
        out = []
        for name, fn, flops in self.trunk_ops + self.head_ops:
            best = float("inf")
            for _ in range(iters):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                e1.synchronize()
                best = min(best, e0.elapsed_time(e1))
            out.append((name, best, flops))
        return out
