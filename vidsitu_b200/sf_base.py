"""Drop-in for `vidsitu_code.mdl_sf_base.SFBase` (reference lines 116-216).

Same constructor `SFBase(cfg, comm)`, same attributes (`sf_mdl`, `head`, `proj_head`),
same state_dict key set, same methods and tensor contracts:

    forward_encoder(inp) -> [ [N,2048,T,7,7], [N,256,T',7,7] ]   (fp32 NCTHW, as the reference)
    head(feat_list)      -> [N, D, 1, 1, 1]
    forward_decoder(enc_out, inp) -> [B, 5, V]
    forward(inp)         -> {"mdl_out": [B, 5, V]}

but every arithmetic op runs in libvidsitu_b200.so (hand-written sm_100a kernels).
There is no torch.nn / cuDNN / CPU path behind these methods: on a machine without a
CUDA device, or without the built library, they raise.

Beyond the reference surface, `extract_features(frames_uint8)` is the fused fast path
of feat_extractor.py:90-112: uint8 NTHWC frames in, pooled [N, D] features (+ logits)
out, one CUDA-graph replay per micro-batch.

Precision: `precision="bf16"` (default) = tcgen05 tensor-core kernels, bf16 storage,
fp32 accumulate/epilogue; `precision="fp32"` = CUDA-core verification kernels.
It can also be set through `cfg.mdl.vsb_precision`.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
from torch import nn

from .arch import build_spec
from .engine import ClipEngine
from .lib import VSB_BF16, VSB_F32, VsbError
from .model import BackboneParams
from . import ops

_PRECISIONS = {"bf16": VSB_BF16, "fp32": VSB_F32}


def combine_first_ax(t: torch.Tensor) -> torch.Tensor:
    """[B, 5, ...] -> [B*5, ...] (utils/misc_utils.py:1-5)."""
    return t.reshape((t.shape[0] * t.shape[1],) + tuple(t.shape[2:]))


class BackboneModel(BackboneParams):
    """`SFBase.sf_mdl`: parameters named as the reference backbone + `forward_features`."""

    def __init__(self, spec):
        super().__init__(spec)
        self._owner = None

    def forward_features(self, x: List[torch.Tensor]) -> List[torch.Tensor]:
        """x = [slow, fast] (or [fast]) fp32 NCTHW -> per-pathway fp32 NCTHW feature maps
        (mdl_sf_base.py:21-34, 46-55)."""
        if self._owner is None:
            raise VsbError("forward_features needs the owning SFBase (engine cache)")
        return self._owner()._forward_features(x)

    # the kernels read prepared copies of the parameters: every way of changing them through `sf_mdl` alone
    # (load_checkpoint(model=mdl.sf_mdl, ...), sf_mdl.load_state_dict, sf_mdl.to / .half / ...) drops those copies
    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        owner = self._owner() if self._owner is not None else None
        if owner is not None:
            owner.invalidate_engines()
        return out

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        owner = self._owner() if getattr(self, "_owner", None) is not None else None
        if owner is not None and hasattr(owner, "_engines"):
            owner.invalidate_engines()
        return out

    def forward(self, x, bboxes=None):
        raise NotImplementedError("the reference never calls sf_mdl.forward on this path "
                                  "(mdl_sf_base.py:36-42 is broken upstream); use forward_features")


class PooledHead(nn.Module):
    """ResNetBasicHead_Trimmed (mdl_sf_base.py:65-113): AdaptiveAvgPool3d((1,1,1)) per pathway, cat on C."""

    def __init__(self, dim_in: Sequence[int], pool_size: Sequence[Optional[Sequence[int]]]):
        super().__init__()
        assert len({len(pool_size), len(dim_in)}) == 1, "pathway dimensions are not consistent."
        if any(p is not None for p in pool_size):
            raise NotImplementedError("only pool_size=None (global average pool) is used by SFBase")
        self.num_pathways = len(pool_size)
        self.dim_in = list(dim_in)

    def forward(self, inputs: List[torch.Tensor]) -> torch.Tensor:
        assert len(inputs) == self.num_pathways, \
            "Input tensor does not contain {} pathway".format(self.num_pathways)
        n = inputs[0].shape[0]
        feats = torch.zeros((n, sum(x.shape[1] for x in inputs)), dtype=torch.float32, device=inputs[0].device)
        off = 0
        for x in inputs:
            if not x.is_cuda:
                raise VsbError("head() runs on CUDA tensors only")
            _, c, t, h, w = x.shape
            if c % 4:
                raise VsbError("channel count must be a multiple of 4")
            # channels-last copy is layout plumbing; the reduction itself is vsb_global_avgpool
            cl = x.float().permute(0, 2, 3, 4, 1).contiguous()
            ops.global_avgpool(ops.Act(cl, n, t, h, w, c, c), feats, off, VSB_F32)
            off += c
        return feats.view(n, -1, 1, 1, 1)


class _KernelLinear(nn.Linear):
    def forward(self, x):
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1]).float().contiguous()
        y = torch.empty((x2.shape[0], self.out_features), dtype=torch.float32, device=x2.device)
        ops.linear(x2, self.weight.detach().float().contiguous(),
                   self.bias.detach().float().contiguous() if self.bias is not None else None, y,
                   getattr(self, "_fuse_relu", False))
        return y.view(*lead, self.out_features)


class _FusedReLU(nn.Module):
    """Placeholder keeping index 1 of proj_head (state_dict keys proj_head.{0,2}.*): the ReLU
    itself is applied inside the preceding vsb_linear launch."""

    def forward(self, x):
        return x


class SFBase(nn.Module):
    def __init__(self, cfg, comm, precision: Optional[str] = None, micro_batch: int = 40,
                 tune: Optional[dict] = None):
        super().__init__()
        self.full_cfg = cfg
        self.sf_cfg = cfg.sf_mdl
        self.cfg = cfg.mdl
        self.comm = comm
        if precision is None:
            precision = getattr(self.cfg, "vsb_precision", "bf16") if self.cfg is not None else "bf16"
        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
        self.precision = precision
        self.micro_batch = int(micro_batch)
        self.tune = tune
        self._engines: Dict[tuple, ClipEngine] = {}
        self._prep: Dict[tuple, dict] = {}      # prepared weights per (precision, device), shared by all batch sizes
        self._host_tensors: Optional[dict] = None
        self._weights_version = 0
        self.build_model()

    # ------------------------------------------------------------------ construction
    def build_model(self):
        self.build_sf_model(self.sf_cfg)
        self.build_head(self.sf_cfg)
        self.build_projection_head(self.sf_cfg)

    def build_sf_model(self, cfg):
        mdl_name = cfg.MODEL.MODEL_NAME
        if mdl_name not in ("SlowFast", "ResNet"):
            raise NotImplementedError
        self.spec = build_spec(cfg)
        self.sf_mdl = BackboneModel(self.spec)
        import weakref
        self.sf_mdl._owner = weakref.ref(self)

    def build_head(self, cfg):
        width_per_group = cfg.RESNET.WIDTH_PER_GROUP
        if self.comm.path_type == "multi":
            self.head = PooledHead(
                dim_in=[width_per_group * 32, width_per_group * 32 // cfg.SLOWFAST.BETA_INV],
                pool_size=[None, None])
        elif self.comm.path_type == "single":
            self.head = PooledHead(dim_in=[width_per_group * 32], pool_size=[None])
        else:
            raise NotImplementedError
        if self.head.num_pathways != self.spec.num_pathways:
            raise ValueError(f"comm.path_type={self.comm.path_type!r} does not match MODEL.ARCH={self.spec.arch!r}")

    def build_projection_head(self, cfg, out_dim=None):
        if out_dim is None:
            out_dim = len(self.comm.vb_id_vocab)
        din = sum(self.head.dim_in)
        first = _KernelLinear(din, din // 2)
        first._fuse_relu = True
        self.proj_head = nn.Sequential(first, _FusedReLU(), _KernelLinear(din // 2, out_dim))

    # ------------------------------------------------------------------ weight tracking
    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self.invalidate_engines()
        return out

    def invalidate_engines(self) -> None:
        """Call after mutating parameters in place: kernels read prepared copies of the weights."""
        self._engines.clear()
        self._prep.clear()
        self._host_tensors = None
        self._weights_version += 1

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        if hasattr(self, "_engines"):
            self.invalidate_engines()
        return out

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError("vidsitu_b200.SFBase is inference-only (frozen BatchNorm folded into the convs); "
                                      "training (utils/trn_utils.py:583-628) is out of scope")
        return super().train(False)

    # ------------------------------------------------------------------ engines
    def _engine(self, n: int, device) -> ClipEngine:
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        slots = getattr(self, "input_slots", 1)
        key = (n, self.precision, str(device), slots)
        eng = self._engines.get(key)
        if eng is None:
            if self._host_tensors is None:
                # one device -> host copy of the reference-layout parameters; BatchNorm folding and weight packing
                # run on the CPU and are cached per (precision, device): an engine for another batch size (a tail
                # batch) re-uses the prepared tensors and only builds its plans and buffers
                self._host_tensors = {k: v.detach().float().cpu() for k, v in self.sf_mdl.state_dict().items()}
            ph = (self.proj_head[0].weight, self.proj_head[0].bias, self.proj_head[2].weight, self.proj_head[2].bias)
            with torch.cuda.device(device):
                eng = ClipEngine(self.spec, self._host_tensors, n, _PRECISIONS[self.precision], device, proj_head=ph,
                                 tune=self.tune, input_slots=slots,
                                 prep_cache=self._prep.setdefault((self.precision, str(device)), {}))
            self._engines[key] = eng
        return eng

    def _device(self) -> torch.device:
        return next(self.parameters()).device

    # ------------------------------------------------------------------ reference surface
    def get_feats(self, inp):
        if self.comm.path_type == "multi":
            return [combine_first_ax(inp["frms_ev_slow_tensor"]), combine_first_ax(inp["frms_ev_fast_tensor"])]
        elif self.comm.path_type == "single":
            return [combine_first_ax(inp["frms_ev_fast_tensor"])]
        raise NotImplementedError

    @torch.no_grad()
    def _forward_features(self, xs: List[torch.Tensor]) -> List[torch.Tensor]:
        if len(xs) != self.spec.num_pathways:
            raise VsbError(f"expected {self.spec.num_pathways} pathway tensors, got {len(xs)}")
        if not xs[0].is_cuda:
            raise VsbError("inputs must be CUDA tensors: vidsitu_b200 has no CPU path")
        n = xs[0].shape[0]
        outs: List[List[torch.Tensor]] = [[] for _ in xs]
        for s in range(0, n, self.micro_batch):
            e = min(n, s + self.micro_batch)
            eng = self._engine(e - s, xs[0].device)
            eng.load_ncthw([x[s:e].float() for x in xs])
            eng.run_trunk()
            for p, t in enumerate(eng.features_ncthw()):
                outs[p].append(t)
        return [torch.cat(o, dim=0) if len(o) > 1 else o[0] for o in outs]

    def forward_encoder(self, inp):
        feats_used = self.get_feats(inp)
        nfeats_used = len(feats_used)
        feat_out = self.sf_mdl.forward_features(feats_used)
        assert len(feat_out) == nfeats_used
        return feat_out

    def forward_decoder(self, enc_out, inp):
        head_out = self.head(enc_out)
        head_out = head_out.permute((0, 2, 3, 4, 1))
        proj_out = self.proj_head(head_out)
        B = len(inp["vseg_idx"])
        out = proj_out.view(B, 5, -1)
        assert out.size(-1) == len(self.comm.vb_id_vocab)
        return out

    def forward(self, inp: Dict):
        feat_out = self.forward_encoder(inp)
        mdl_out = self.forward_decoder(feat_out, inp)
        return {"mdl_out": mdl_out}

    @torch.no_grad()
    def predict_verbs(self, inp: Dict, topk_save: int = 5) -> List[dict]:
        """`EvalB.forward_one_batch` (vidsitu_code/evl_vsitu.py:39-75): forward, softmax, descending sort,
        top-5 verbs and scores per event, as the list of {"pred_vbs_ev", "pred_scores_ev", "ann_idx"}
        dicts the evaluator pickles.  The softmax / ranking run on the GPU (vsb_softmax_topk); only
        5 ids + 5 scores per event cross PCIe instead of the [B, 5, V] logits."""
        mdl_out = self.forward(inp)["mdl_out"]
        idx, prob = ops.softmax_topk(mdl_out.float(), topk_save)
        idx_l, prob_l = idx.tolist(), prob.tolist()
        ann_lst = inp["vseg_idx"].tolist()
        symbols = getattr(self.comm.vb_id_vocab, "symbols", self.comm.vb_id_vocab)   # fairseq Dictionary.symbols
        out = []
        for pred_vbs, pred_scores, ann_idx in zip(idx_l, prob_l, ann_lst):
            assert len(pred_vbs) == 5 and len(pred_scores) == 5
            out.append({"pred_vbs_ev": [[symbols[pv] for pv in pvb] for pvb in pred_vbs],
                        "pred_scores_ev": [list(pvs) for pvs in pred_scores], "ann_idx": ann_idx})
        return out

    # ------------------------------------------------------------------ fused fast paths
    @torch.no_grad()
    def extract_features(self, frames: torch.Tensor, want_logits: bool = False, use_graph: bool = True):
        """frames: uint8 [N, T, 224, 224, 3] (T = DATA.NUM_FRAMES window of the fast/single pathway,
        dat_loader.py:454-476) on the GPU.  Returns feats [N, D] fp32 (and logits [N, V])."""
        if not frames.is_cuda:
            raise VsbError("frames must be a CUDA tensor (pinned-host staging is the caller's H2D copy)")
        n = frames.shape[0]
        feats, logits = [], []
        for s in range(0, n, self.micro_batch):
            e = min(n, s + self.micro_batch)
            eng = self._engine(e - s, frames.device)
            eng.load_frames(frames[s:e])
            if use_graph:
                eng.replay()
            else:
                eng.run()
            # always a copy: eng.feats / eng.logits are the engine's static CUDA-graph output buffers, the next
            # call with the same batch size overwrites them
            feats.append(eng.feats.clone())
            if want_logits:
                logits.append(eng.logits.clone())
        f = torch.cat(feats, 0) if len(feats) > 1 else feats[0]
        if want_logits:
            return f, (torch.cat(logits, 0) if len(logits) > 1 else logits[0])
        return f

    @torch.no_grad()
    def extract_video_features(self, videos: torch.Tensor, want_logits: bool = False, use_graph: bool = True):
        """videos: uint8 [B, F, H, W, 3] on the GPU, F = the 300 frames `get_frms_all` lists per video
        (dat_loader.py:454-458).  The five event windows (dat_loader.py:69-79, 459-472) are cut on the
        device by the pack kernel; returns feats [B, 5, D] fp32 - the array `FeatExtract.forward_all`
        saves per video (feat_extractor.py:98-111) - and, on request, logits [B, 5, V]."""
        if not videos.is_cuda:
            raise VsbError("videos must be a CUDA tensor (pinned-host staging is the caller's H2D copy)")
        b = videos.shape[0]
        per = max(1, self.micro_batch // 5)
        feats, logits = [], []
        for s in range(0, b, per):
            e = min(b, s + per)
            eng = self._engine(5 * (e - s), videos.device)
            eng.load_videos(videos[s:e])
            if use_graph:
                eng.replay()
            else:
                eng.run()
            # engine rows are event-major (event * n_videos + video) -> [videos, 5, D]
            # (copied out: the engine's output buffers are reused by the next chunk)
            feats.append(torch.empty((e - s, 5, eng.feats.shape[1]), dtype=torch.float32, device=videos.device)
                         .copy_(eng.feats.view(5, e - s, -1).transpose(0, 1)))
            if want_logits:
                logits.append(torch.empty((e - s, 5, eng.logits.shape[1]), dtype=torch.float32, device=videos.device)
                              .copy_(eng.logits.view(5, e - s, -1).transpose(0, 1)))
        f = torch.cat(feats, 0) if len(feats) > 1 else feats[0]
        if want_logits:
            return f, (torch.cat(logits, 0) if len(logits) > 1 else logits[0])
        return f

    @torch.no_grad()
    def forward_pooled(self, inp: Dict, want_logits: bool = True):
        """Same inputs as forward() (reference fp32 NCTHW tensors) but never materialises the
        [N,2048,T,7,7] maps: returns (feats [B,5,D], logits [B,5,V])."""
        xs = self.get_feats(inp)
        n = xs[0].shape[0]
        feats, logits = [], []
        for s in range(0, n, self.micro_batch):
            e = min(n, s + self.micro_batch)
            eng = self._engine(e - s, xs[0].device)
            eng.load_ncthw([x[s:e].float() for x in xs])
            eng.replay()
            feats.append(eng.feats.clone())
            logits.append(eng.logits.clone())
        f = torch.cat(feats, 0).view(n // 5, 5, -1)
        lg = torch.cat(logits, 0).view(n // 5, 5, -1)
        return (f, lg) if want_logits else f
