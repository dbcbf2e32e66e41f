"""Operator-level host wrappers: torch tensors in (device memory + current stream are
the only things torch provides), C-ABI calls out.  Each function mirrors one family
of reference calls; see include/vidsitu_b200.h for the file:line citations.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import lib as _l
from .lib import VSB_BF16, VSB_F32, BottleneckDesc, ConvDesc, StemPoolDesc, VsbError, check

TORCH_DTYPE = {VSB_BF16: torch.bfloat16, VSB_F32: torch.float32}


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise VsbError("vidsitu_b200 ops run on CUDA tensors only (there is no CPU fallback)")


@dataclass
class Act:
    """A channels-last activation: `buf` holds [n, t, h, w, pitch] elements, the
    logical tensor is its channel slice [c_off, c_off + c)."""

    buf: torch.Tensor
    n: int
    t: int
    h: int
    w: int
    c: int            # stored channels of this view (padded to the dtype's channel quantum)
    pitch: int
    c_off: int = 0
    c_real: int = -1  # logical (reference) channel count

    def __post_init__(self):
        if self.c_real < 0:
            self.c_real = self.c

    @property
    def ptr(self) -> int:
        return self.buf.data_ptr() + self.c_off * self.buf.element_size()

    @property
    def pixels(self) -> int:
        return self.n * self.t * self.h * self.w

    def nthwc(self) -> torch.Tensor:
        """[n, t, h, w, c_real] torch view of the logical tensor (no copy)."""
        v = self.buf.view(-1)[: self.pixels * self.pitch].view(self.n, self.t, self.h, self.w, self.pitch)
        return v[..., self.c_off: self.c_off + self.c_real]


class Call:
    """One C-ABI launch with its arguments marshalled once (include/vidsitu_b200.h: vsb_<op>(args..., stream)).
    Calling it launches on torch's current stream; emit() appends it to a clip program
    (vsb_program_add_<op>(program, args..., lane, name))."""

    def __init__(self, op: str, args: tuple, keep: tuple = ()):
        self.op = op
        self.args = args
        self._keep = keep          # tensors / ctypes arrays the arguments point into
        self._fn = getattr(_l.load(), "vsb_" + op)

    def __call__(self) -> None:
        check(self._fn(*self.args, _stream_ptr()), "vsb_" + self.op)

    def emit(self, prog: "Program", lane: int, name: str) -> None:
        check(getattr(_l.load(), "vsb_program_add_" + self.op)(prog.handle, *self.args, lane, name.encode()),
              "vsb_program_add_" + self.op)
        prog.keep(self)


class Program:
    """A clip program (include/vidsitu_b200.h, ABI v7): the whole forward as ONE C handle - built by
    ClipEngine.build_program, replayed by vsb_program_run (optionally as a CUDA graph captured inside the library),
    saved to / loaded from a relocatable file that any host process can run without Python."""

    CONST, SCRATCH = 0, 1

    def __init__(self, handle: Optional[int] = None):
        self._lib = _l.load()
        self._h = C.c_void_p(handle)
        self._keep: list = []
        if handle is None:
            check(self._lib.vsb_program_create(C.byref(self._h)), "vsb_program_create")

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    def keep(self, obj) -> None:
        self._keep.append(obj)     # plans and buffers are borrowed by the C side

    def sync(self, src: int, dst: int) -> None:
        check(self._lib.vsb_program_add_sync(self._h, src, dst), "vsb_program_add_sync")

    def add_region(self, name: str, t: torch.Tensor, kind: int) -> None:
        _require_cuda(t)
        check(self._lib.vsb_program_add_region(self._h, name.encode(), t.data_ptr(), t.numel() * t.element_size(), kind),
              "vsb_program_add_region")
        self._keep.append(t)

    def region(self, name: str):
        """(device pointer, bytes) of a named region."""
        ptr, nbytes = C.c_void_p(), C.c_ulonglong()
        check(self._lib.vsb_program_region(self._h, name.encode(), C.byref(ptr), C.byref(nbytes)), "vsb_program_region")
        return int(ptr.value), int(nbytes.value)

    @property
    def num_ops(self) -> int:
        return int(self._lib.vsb_program_num_ops(self._h))

    @property
    def num_launches(self) -> int:
        return int(self._lib.vsb_program_num_launches(self._h))

    @property
    def device_bytes(self) -> int:
        return int(self._lib.vsb_program_device_bytes(self._h))

    def run(self) -> None:
        check(self._lib.vsb_program_run(self._h, _stream_ptr()), "vsb_program_run")

    def capture(self) -> None:
        check(self._lib.vsb_program_capture(self._h, _stream_ptr()), "vsb_program_capture")

    def save(self, path: str) -> None:
        check(self._lib.vsb_program_save(self._h, str(path).encode()), "vsb_program_save")

    @classmethod
    def load(cls, path: str) -> "Program":
        """Rebuild a saved program in library-owned device memory (current device)."""
        lib = _l.load()
        h = C.c_void_p()
        check(lib.vsb_program_load(str(path).encode(), None, 0, C.byref(h)), "vsb_program_load")
        return cls(h.value)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._lib.vsb_program_destroy(h)
            h.value = None


class ConvPlan:
    """One planned conv launch (TMA descriptors are encoded once, at plan time)."""

    def __init__(self, dtype: int, x: Act, wgt: torch.Tensor, cout: int, kernel: Sequence[int],
                 stride: Sequence[int], pad_lo: Sequence[int], pad_hi: Optional[Sequence[int]],
                 scale: torch.Tensor, bias: torch.Tensor, out: Act, residual: Optional[Act] = None,
                 relu: bool = False, block_n: int = 0, kchunk: int = 0, stages: int = 0, algo: int = 0,
                 kw_ranges: Optional[Sequence[Sequence[int]]] = None, x2: Optional[Act] = None,
                 stride2: Sequence[int] = (1, 1, 1), epi_n: int = 0, epi_bufs: int = 0, flags: int = 0, out_f16: bool = False,
                 wgt_clip_rows: int = 0, tile_signal: Optional[torch.Tensor] = None,
                 tile_wait: Optional[torch.Tensor] = None, tile_wait_count: int = 0, grid_limit: int = 0,
                 in_f16: bool = False):
        _require_cuda(x.buf, wgt, scale, bias, out.buf, residual.buf if residual is not None else None,
                      x2.buf if x2 is not None else None, tile_signal, tile_wait)
        pad_hi = pad_lo if pad_hi is None else pad_hi
        d = ConvDesc()
        d.dtype = dtype
        d.inp = x.ptr
        d.n, d.t, d.h, d.w, d.cin, d.in_pitch = x.n, x.t, x.h, x.w, x.c, x.pitch
        d.wgt = wgt.data_ptr()
        d.cout = cout
        d.kt, d.kh, d.kw = kernel
        d.st, d.sh, d.sw = stride
        d.pt_lo, d.ph_lo, d.pw_lo = pad_lo
        d.pt_hi, d.ph_hi, d.pw_hi = pad_hi
        d.scale = scale.data_ptr()
        d.bias = bias.data_ptr()
        d.residual = residual.ptr if residual is not None else None
        d.res_pitch = residual.pitch if residual is not None else 0
        d.relu = int(relu)
        d.out = out.ptr
        d.out_pitch = out.pitch
        d.block_n, d.kchunk, d.stages = block_n, kchunk, stages
        d.algo = algo
        d.epi_n, d.epi_bufs, d.flags = epi_n, epi_bufs, flags
        d.out_f16 = int(out_f16)
        d.wgt_clip_rows = int(wgt_clip_rows)
        # tile-granular chaining with the next / previous launch (vsb_conv_desc.tile_signal / tile_wait)
        for t in (tile_signal, tile_wait):
            if t is not None and (t.dtype != torch.int32 or t.numel() * 128 < out.pixels):
                raise VsbError("tile counters: one int32 per 128 output rows")
        d.tile_signal = tile_signal.data_ptr() if tile_signal is not None else None
        d.tile_wait = tile_wait.data_ptr() if tile_wait is not None else None
        d.tile_wait_count = int(tile_wait_count)
        d.grid_limit = int(grid_limit)
        d.in_f16 = int(in_f16)
        k_per_tap_row = x.c * d.kw
        if kw_ranges is not None:
            if len(kw_ranges) != d.kw or d.kw > 8:
                raise VsbError("kw_ranges needs one (lo, hi) pair per kw tap (kw <= 8)")
            for k, (lo, hi) in enumerate(kw_ranges):
                d.kw_c_lo[k], d.kw_c_hi[k] = int(lo), int(hi)
            k_per_tap_row = sum(int(hi) - int(lo) for lo, hi in kw_ranges)
        k2 = 0
        if x2 is not None:
            # second source (fused shortcut projection): strided 1x1x1 over x2, weights appended along K
            if not kchunk:
                raise VsbError("a second source needs an explicit kchunk (its weights are padded to it)")
            d.in2 = x2.ptr
            d.t2, d.h2, d.w2, d.cin2, d.in2_pitch = x2.t, x2.h, x2.w, x2.c, x2.pitch
            d.st2, d.sh2, d.sw2 = stride2
            k2 = -(-x2.c // kchunk) * kchunk
            if x2.n != x.n or x2.buf.dtype != x.buf.dtype:
                raise VsbError("second source must have the batch size and dtype of the first")
        expect = torch.bfloat16 if dtype == VSB_BF16 else torch.float32
        if x.buf.dtype != expect or wgt.dtype != expect or out.buf.dtype != expect:
            raise VsbError(f"conv tensors must be {expect}")
        if scale.dtype != torch.float32 or bias.dtype != torch.float32:
            raise VsbError("scale/bias must be float32")
        if wgt_clip_rows:
            if wgt.numel() < ((x.n - 1) * wgt_clip_rows + cout) * x.c:
                raise VsbError("per-clip weights: wgt holds fewer than (n-1)*wgt_clip_rows + cout rows")
        elif wgt.numel() != cout * (d.kt * d.kh * k_per_tap_row + k2):
            raise VsbError(f"packed weight has {wgt.numel()} elements, expected {cout}x({d.kt * d.kh}x{k_per_tap_row}+{k2})")
        self._keep = (x.buf, wgt, scale, bias, out.buf, residual.buf if residual is not None else None,
                      x2.buf if x2 is not None else None, tile_signal, tile_wait)
        self._h = C.c_void_p()
        self._lib = _l.load()
        check(self._lib.vsb_conv3d_plan_create(C.byref(d), C.byref(self._h)), "vsb_conv3d_plan_create")
        to, ho, wo = C.c_int(), C.c_int(), C.c_int()
        check(self._lib.vsb_conv3d_plan_out_shape(self._h, C.byref(to), C.byref(ho), C.byref(wo)), "out_shape")
        self.out_shape = (to.value, ho.value, wo.value)
        self.flops = float(self._lib.vsb_conv3d_plan_flops(self._h))
        if x.n * to.value * ho.value * wo.value != out.pixels:
            raise VsbError(f"conv output extent {x.n}x{self.out_shape} does not match the output buffer "
                           f"({out.n},{out.t},{out.h},{out.w})")

    def run(self) -> None:
        check(self._lib.vsb_conv3d_run(self._h, _stream_ptr()), "vsb_conv3d_run")

    def emit(self, prog: "Program", lane: int, name: str) -> None:
        check(self._lib.vsb_program_add_conv(prog.handle, self._h, lane, name.encode()), "vsb_program_add_conv")
        prog.keep(self)

    def info(self) -> dict:
        """How the plan runs (debug / tests): algorithm, mode, pipeline depth, grid, shared memory."""
        out = (C.c_longlong * 8)()
        check(self._lib.vsb_debug_conv_plan_info(self._h, out), "vsb_debug_conv_plan_info")
        keys = ("algo", "tsc", "stages", "nacc", "grid", "smem_bytes", "block_n", "ctas_per_sm")
        return dict(zip(keys, [int(v) for v in out]))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._lib.vsb_conv3d_plan_destroy(h)
            h.value = None


class BottleneckPlan:
    """One planned fused identity-bottleneck launch (vsb_bottleneck_*): out = relu(x + c(b(a(x)))) with the three
    frozen BatchNorms folded, a's and b's outputs never leaving the SM.  All tensors bf16, weights K-major and
    zero-padded to the stored widths: wa [d, kt, c], wb [d, 9, d], wc [c, d]."""

    def __init__(self, x: Act, out: Act, d: int, kt: int, wa: torch.Tensor, wb: torch.Tensor, wc: torch.Tensor,
                 sa: torch.Tensor, ba: torch.Tensor, sb: torch.Tensor, bb: torch.Tensor, sc: torch.Tensor,
                 bc: torch.Tensor, stages: int = 0, walk_len: int = 0, grid: int = 0, algo: int = 0, cin: int = 0):
        _require_cuda(x.buf, out.buf, wa, wb, wc, sa, ba, sb, bb, sc, bc)
        c = out.c if cin else x.c   # cin: projection block (algo 1), x has cin channels, wc = [Wc' | W1']
        if (out.n, out.t, out.h, out.w) != (x.n, x.t, x.h, x.w) or out.c != c or (cin and x.c < cin):
            raise VsbError("fused bottleneck: output extent / width must equal the input's")
        for t in (x.buf, out.buf, wa, wb, wc):
            if t.dtype != torch.bfloat16:
                raise VsbError("fused bottleneck tensors must be bfloat16")
        ca, kc = (cin, d + cin) if cin else (c, d)
        if wa.numel() != d * kt * ca or wb.numel() != d * 9 * d or wc.numel() != c * kc:
            raise VsbError(f"fused bottleneck weights must be [{d},{kt},{ca}], [{d},9,{d}], [{c},{kc}]")
        for t, nn in ((sa, d), (ba, d), (sb, d), (bb, d), (sc, c), (bc, c)):
            if t.dtype != torch.float32 or t.numel() != nn or not t.is_contiguous():
                raise VsbError("fused bottleneck scale / bias must be contiguous float32 of the stored widths")
        dsc = BottleneckDesc()
        dsc.x = x.ptr
        dsc.n, dsc.t, dsc.h, dsc.w, dsc.c, dsc.x_pitch = x.n, x.t, x.h, x.w, c, x.pitch
        dsc.out = out.ptr
        dsc.out_pitch = out.pitch
        dsc.d, dsc.kt = d, kt
        dsc.wa, dsc.wb, dsc.wc = wa.data_ptr(), wb.data_ptr(), wc.data_ptr()
        dsc.sa, dsc.ba, dsc.sb, dsc.bb, dsc.sc, dsc.bc = (t.data_ptr() for t in (sa, ba, sb, bb, sc, bc))
        dsc.stages, dsc.walk_len, dsc.grid = stages, walk_len, grid
        dsc.cin = cin
        dsc.algo = algo   # 0 = tcgen05 flat-raster kernel, 1 = warp-MMA walk kernel (thin blocks: d = 8 / 16, c = 4 d)
        self.algo = algo
        self._keep = (x.buf, out.buf, wa, wb, wc, sa, ba, sb, bb, sc, bc)
        self._desc = dsc
        self._h = C.c_void_p()
        self._lib = _l.load()
        check(self._lib.vsb_bottleneck_plan_create(C.byref(dsc), C.byref(self._h)), "vsb_bottleneck_plan_create")

    def run(self) -> None:
        check(self._lib.vsb_bottleneck_run(self._h, _stream_ptr()), "vsb_bottleneck_run")

    def emit(self, prog: "Program", lane: int, name: str) -> None:
        check(self._lib.vsb_program_add_bottleneck(prog.handle, self._h, C.byref(self._desc), lane, name.encode()),
              "vsb_program_add_bottleneck")
        prog.keep(self)

    def info(self) -> dict:
        out = (C.c_longlong * 8)()
        check(self._lib.vsb_bottleneck_plan_info(self._h, out), "vsb_bottleneck_plan_info")
        if self.algo == 1:
            keys = ("rows_per_strip", "strips", "ring_slots", "steps_per_cta", "grid", "smem_bytes", "a_x1000_bc_tiles", "_")
        else:
            keys = ("rp", "fp", "stages_x100_cps", "tiles_per_cta", "grid", "smem_bytes", "tiles_per_clip", "tmem_cols")
        return dict(zip(keys, [int(v) for v in out]))

    def stats(self) -> list:
        """Role timeline counters (plan created under VSB_FUSED_DEBUG=1); clears them."""
        out = (C.c_longlong * 32)()
        check(self._lib.vsb_debug_bottleneck_stats(self._h, out), "vsb_debug_bottleneck_stats")
        return [int(v) for v in out]

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._lib.vsb_bottleneck_plan_destroy(h)
            h.value = None


class StemPoolPlan:
    """One planned fused [1,7,7] stem (vsb_stem_pool_*): conv + frozen BN + ReLU + 1x3x3/s2 max-pool of the packed
    bf16 frames in one kernel (stem_helper.py:157-178); `out` receives the pooled [frames, h/4, w/4, 64] channels."""

    def __init__(self, x: Act, x_off: int, wgt: torch.Tensor, scale: torch.Tensor, bias: torch.Tensor, out: Act, w: int,
                 kt: int = 1):
        _require_cuda(x.buf, wgt, scale, bias, out.buf)
        if x_off != 3 or x.pitch != 4 or x.c_off:
            raise VsbError("fused stem: input must be the packed 4-channel frames with a 3-pixel left border")
        if wgt.dtype != torch.bfloat16 or wgt.numel() != kt * 64 * 7 * 32 or not wgt.is_contiguous():
            raise VsbError("fused stem: weights must be contiguous bf16 [kt * 64, 7, 8, 4]")
        for t in (scale, bias):
            if t.dtype != torch.float32 or t.numel() != 64 or not t.is_contiguous():
                raise VsbError("fused stem: scale / bias must be contiguous float32 [64]")
        if (out.n, out.t, out.h, out.w) != (x.n, x.t, x.h // 4, w // 4) or out.c != 64:
            raise VsbError("fused stem: output must be [n, t, h/4, w/4] with 64 stored channels")
        d = StemPoolDesc()
        d.in_ = x.ptr
        d.frames, d.h, d.w, d.w_buf = x.n * x.t, x.h, w, x.w
        d.wgt, d.scale, d.bias = wgt.data_ptr(), scale.data_ptr(), bias.data_ptr()
        d.out, d.out_pitch = out.ptr, out.pitch
        d.kt, d.t = kt, x.t
        self._keep = (x.buf, wgt, scale, bias, out.buf)
        self._h = C.c_void_p()
        self._lib = _l.load()
        check(self._lib.vsb_stem_pool_plan_create(C.byref(d), C.byref(self._h)), "vsb_stem_pool_plan_create")
        self.flops = 2.0 * d.frames * (d.h // 2) * (w // 2) * 64 * 7 * 32 * kt

    def run(self) -> None:
        check(self._lib.vsb_stem_pool_run(self._h, _stream_ptr()), "vsb_stem_pool_run")

    def emit(self, prog: "Program", lane: int, name: str) -> None:
        check(self._lib.vsb_program_add_stem_pool(prog.handle, self._h, lane, name.encode()), "vsb_program_add_stem_pool")
        prog.keep(self)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._lib.vsb_stem_pool_plan_destroy(h)
            h.value = None


def pack_frames(frames: torch.Tensor, idx: Sequence[int], mean: Sequence[float], std: Sequence[float],
                out: Act, dtype: int, reverse_channels: bool = False, x_off: int = 0) -> None:
    """frames uint8 [n, t_in, h, w, 3] -> out [n, len(idx), h, out.w, 4] (normalised, temporally subsampled),
    frame columns written at pixel offset x_off of each (possibly wider, zero-bordered) output row."""
    pack_frames_call(frames, idx, mean, std, out, dtype, reverse_channels, x_off)()


def pack_frames_call(frames: torch.Tensor, idx: Sequence[int], mean: Sequence[float], std: Sequence[float],
                     out: Act, dtype: int, reverse_channels: bool = False, x_off: int = 0) -> Call:
    _require_cuda(frames, out.buf)
    if frames.dtype != torch.uint8 or frames.dim() != 5 or frames.shape[-1] != 3 or not frames.is_contiguous():
        raise VsbError("frames must be a contiguous uint8 [n, t, h, w, 3] tensor")
    n, t_in, h, w, _ = frames.shape
    if (out.n, out.t, out.h, out.pitch, out.c_off) != (n, len(idx), h, 4, 0) or out.w < w + x_off:
        raise VsbError("pack output must be a dense [n, len(idx), h, >= w + x_off, 4] activation")
    idx_arr = (C.c_int * len(idx))(*[int(i) for i in idx])
    m = (C.c_float * 3)(*mean)
    s = (C.c_float * 3)(*std)
    return Call("pack_frames", (frames.data_ptr(), n, t_in, h, w, idx_arr, len(idx), m, s, int(reverse_channels),
                                out.ptr, 4, out.w, x_off, dtype), (frames, out.buf, idx_arr, m, s))


def ncthw_to_act(x: torch.Tensor, out: Act, dtype: int, x_off: int = 0) -> None:
    """fp32 NCTHW clip tensor (reference layout) -> 4-channel NTHWC activation."""
    _require_cuda(x, out.buf)
    if x.dtype != torch.float32 or x.dim() != 5 or not x.is_contiguous():
        raise VsbError("expected a contiguous float32 [n, c, t, h, w] tensor")
    n, c, t, h, w = x.shape
    if (out.n, out.t, out.h, out.pitch) != (n, t, h, 4) or out.w < w + x_off:
        raise VsbError("output activation must be dense [n, t, h, >= w + x_off, 4]")
    check(_l.load().vsb_ncthw_f32_to_nthwc(x.data_ptr(), n, c, t * h * w, w, out.ptr, 4, out.w, x_off, dtype,
                                           _stream_ptr()), "vsb_ncthw_f32_to_nthwc")


def maxpool3d(x: Act, out: Act, kernel, stride, pad, dtype: int) -> None:
    maxpool3d_call(x, out, kernel, stride, pad, dtype)()


def maxpool3d_call(x: Act, out: Act, kernel, stride, pad, dtype: int) -> Call:
    _require_cuda(x.buf, out.buf)
    return Call("maxpool3d", (x.ptr, x.n, x.t, x.h, x.w, x.c_real, x.pitch, out.ptr, out.pitch, out.c,
                              kernel[0], kernel[1], kernel[2], stride[0], stride[1], stride[2],
                              pad[0], pad[1], pad[2], dtype), (x.buf, out.buf))


def global_avgpool(x: Act, feats: torch.Tensor, feat_off: int, dtype: int) -> None:
    global_avgpool_call(x, feats, feat_off, dtype)()


def global_avgpool_call(x: Act, feats: torch.Tensor, feat_off: int, dtype: int) -> Call:
    _require_cuda(x.buf, feats)
    if feats.dtype != torch.float32 or feats.dim() != 2 or feats.shape[0] != x.n:
        raise VsbError("feats must be float32 [n, D]")
    return Call("global_avgpool", (x.ptr, x.n, x.t * x.h * x.w, x.c_real, x.pitch, feats.data_ptr(), feats.stride(0),
                                   feat_off, dtype), (x.buf, feats))


def linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], y: torch.Tensor, relu: bool) -> None:
    linear_call(x, w, b, y, relu)()


def linear_call(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], y: torch.Tensor, relu: bool) -> Call:
    _require_cuda(x, w, y)
    for t in (x, w, y):
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise VsbError("linear operands must be contiguous float32")
    n, din = x.shape
    dout = w.shape[0]
    if w.shape[1] != din or tuple(y.shape) != (n, dout):
        raise VsbError("linear shape mismatch")
    return Call("linear", (x.data_ptr(), n, din, w.data_ptr(), b.data_ptr() if b is not None else None,
                           y.data_ptr(), dout, int(relu)), (x, w, b, y))


def softmax_topk(logits: torch.Tensor, k: int = 5):
    """softmax + descending sort + [:k] of EvalB.forward_one_batch (evl_vsitu.py:39-47) on the last dim of
    fp32 logits [..., V]: returns (idx int32 [..., k], prob fp32 [..., k])."""
    _require_cuda(logits)
    if logits.dtype != torch.float32:
        raise VsbError("softmax_topk takes float32 logits")
    lead = tuple(logits.shape[:-1])
    x = logits.reshape(-1, logits.shape[-1])
    if x.stride(-1) != 1 or x.stride(0) < x.shape[1]:
        x = x.contiguous()
    n, v = x.shape
    idx = torch.empty((n, k), dtype=torch.int32, device=x.device)
    prob = torch.empty((n, k), dtype=torch.float32, device=x.device)
    check(_l.load().vsb_softmax_topk(x.data_ptr(), n, v, x.stride(0), k, idx.data_ptr(), prob.data_ptr(),
                                     _stream_ptr()), "vsb_softmax_topk")
    return idx.view(*lead, k), prob.view(*lead, k)


def nonlocal_attention(theta: Act, phi: Act, g: Act, out: Act, softmax: bool, dtype: int) -> None:
    nonlocal_attention_call(theta, phi, g, out, softmax, dtype)()


def nonlocal_attention_call(theta: Act, phi: Act, g: Act, out: Act, softmax: bool, dtype: int) -> Call:
    _require_cuda(theta.buf, phi.buf, g.buf, out.buf)
    tq = theta.t * theta.h * theta.w
    tk = phi.t * phi.h * phi.w
    return Call("nonlocal_attention", (theta.ptr, theta.pitch, phi.ptr, phi.pitch, g.ptr, g.pitch, out.ptr, out.pitch,
                                       theta.n, tq, tk, theta.c_real, int(softmax), dtype),
                (theta.buf, phi.buf, g.buf, out.buf))


def act_to_ncthw(x: Act, dtype: int) -> torch.Tensor:
    """Materialise an activation as the reference's float32 [n, c, t, h, w] tensor."""
    _require_cuda(x.buf)
    out = torch.empty((x.n, x.c_real, x.t, x.h, x.w), dtype=torch.float32, device=x.buf.device)
    check(_l.load().vsb_nthwc_to_ncthw_f32(x.ptr, x.n, x.t * x.h * x.w, x.c_real, x.pitch, out.data_ptr(), dtype,
                                           _stream_ptr()), "vsb_nthwc_to_ncthw_f32")
    return out


def score_rows(scores: torch.Tensor, rows: int, valid: int, width: int, pitch: int, softmax: bool) -> None:
    """In place on bf16 scores [rows, pitch]: softmax over columns [0, valid) (or unchanged), zeros in [valid, width)."""
    score_rows_call(scores, rows, valid, width, pitch, softmax)()


def score_rows_call(scores: torch.Tensor, rows: int, valid: int, width: int, pitch: int, softmax: bool) -> Call:
    _require_cuda(scores)
    return Call("score_rows", (scores.data_ptr(), rows, valid, width, pitch, int(softmax)), (scores,))


def transpose_pad(src: Act, out: torch.Tensor, out_pitch: int) -> None:
    """bf16 [n, keys, c] (pitch) -> out [n, c, out_pitch], zero-padded along keys."""
    transpose_pad_call(src, out, out_pitch)()


def transpose_pad_call(src: Act, out: torch.Tensor, out_pitch: int) -> Call:
    _require_cuda(src.buf, out)
    keys = src.t * src.h * src.w
    return Call("transpose_pad", (src.ptr, src.pitch, out.data_ptr(), src.n, keys, src.c, out_pitch), (src.buf, out))
