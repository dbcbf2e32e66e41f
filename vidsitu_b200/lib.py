"""ctypes binding of libvidsitu_b200.so (the C ABI declared in include/vidsitu_b200.h).

The library is the product: if it is missing, or a call fails, this module raises --
there is no PyTorch / CPU fallback behind any of these wrappers.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

VSB_BF16 = 0
VSB_F32 = 1

_LIB_NAME = "libvidsitu_b200.so"
_lib = None


class VsbError(RuntimeError):
    """A libvidsitu_b200 entry point returned a non-zero status."""


class ConvDesc(C.Structure):
    """Mirror of `vsb_conv_desc` (include/vidsitu_b200.h)."""

    _fields_ = [
        ("dtype", C.c_int),
        ("inp", C.c_void_p),
        ("n", C.c_int), ("t", C.c_int), ("h", C.c_int), ("w", C.c_int), ("cin", C.c_int), ("in_pitch", C.c_int),
        ("wgt", C.c_void_p),
        ("cout", C.c_int),
        ("kt", C.c_int), ("kh", C.c_int), ("kw", C.c_int),
        ("st", C.c_int), ("sh", C.c_int), ("sw", C.c_int),
        ("pt_lo", C.c_int), ("ph_lo", C.c_int), ("pw_lo", C.c_int),
        ("pt_hi", C.c_int), ("ph_hi", C.c_int), ("pw_hi", C.c_int),
        ("scale", C.c_void_p),
        ("bias", C.c_void_p),
        ("residual", C.c_void_p),
        ("res_pitch", C.c_int),
        ("relu", C.c_int),
        ("out", C.c_void_p),
        ("out_pitch", C.c_int),
        ("block_n", C.c_int), ("kchunk", C.c_int), ("stages", C.c_int),
        ("algo", C.c_int),
        ("kw_c_lo", C.c_int * 8),
        ("kw_c_hi", C.c_int * 8),
        ("in2", C.c_void_p),
        ("t2", C.c_int), ("h2", C.c_int), ("w2", C.c_int), ("cin2", C.c_int), ("in2_pitch", C.c_int),
        ("st2", C.c_int), ("sh2", C.c_int), ("sw2", C.c_int),
        ("epi_n", C.c_int), ("epi_bufs", C.c_int), ("flags", C.c_int), ("out_f16", C.c_int), ("wgt_clip_rows", C.c_int),
        ("tile_signal", C.c_void_p), ("tile_wait", C.c_void_p), ("tile_wait_count", C.c_int),
        ("grid_limit", C.c_int), ("in_f16", C.c_int),
    ]


class BottleneckDesc(C.Structure):
    """Mirror of `vsb_bottleneck_desc` (include/vidsitu_b200.h)."""

    _fields_ = [
        ("x", C.c_void_p),
        ("n", C.c_int), ("t", C.c_int), ("h", C.c_int), ("w", C.c_int), ("c", C.c_int), ("x_pitch", C.c_int),
        ("out", C.c_void_p),
        ("out_pitch", C.c_int),
        ("d", C.c_int),
        ("kt", C.c_int),
        ("wa", C.c_void_p), ("wb", C.c_void_p), ("wc", C.c_void_p),
        ("sa", C.c_void_p), ("ba", C.c_void_p),
        ("sb", C.c_void_p), ("bb", C.c_void_p),
        ("sc", C.c_void_p), ("bc", C.c_void_p),
        ("stages", C.c_int), ("walk_len", C.c_int), ("grid", C.c_int),
        ("algo", C.c_int), ("cin", C.c_int),
    ]


class StemPoolDesc(C.Structure):
    """Mirror of `vsb_stem_pool_desc` (include/vidsitu_b200.h)."""

    _fields_ = [
        ("in_", C.c_void_p),
        ("frames", C.c_int), ("h", C.c_int), ("w", C.c_int), ("w_buf", C.c_int),
        ("wgt", C.c_void_p), ("scale", C.c_void_p), ("bias", C.c_void_p),
        ("out", C.c_void_p),
        ("out_pitch", C.c_int),
        ("kt", C.c_int), ("t", C.c_int),
    ]


def lib_path() -> Path:
    env = os.environ.get("VIDSITU_B200_LIB")
    return Path(env) if env else Path(__file__).resolve().parent / _LIB_NAME


def load() -> C.CDLL:
    """Load the shared library once; raise (never fall back) if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not path.exists():
        raise VsbError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(vidsitu_b200/csrc/build.sh). There is no fallback implementation."
        )
    lib = C.CDLL(str(path))
    i, vp, f32p, ll = C.c_int, C.c_void_p, C.c_void_p, C.c_longlong
    lib.vsb_abi_version.restype = i
    lib.vsb_last_error.restype = C.c_char_p
    lib.vsb_launch_count.restype = C.c_uint64
    lib.vsb_pack_frames.argtypes = [vp, i, i, i, i, C.POINTER(C.c_int), i, C.POINTER(C.c_float),
                                    C.POINTER(C.c_float), i, vp, i, i, i, i, vp]
    lib.vsb_conv3d_plan_create.argtypes = [C.POINTER(ConvDesc), C.POINTER(vp)]
    lib.vsb_conv3d_run.argtypes = [vp, vp]
    lib.vsb_conv3d_plan_destroy.argtypes = [vp]
    lib.vsb_conv3d_plan_destroy.restype = None
    lib.vsb_conv3d_plan_out_shape.argtypes = [vp, C.POINTER(i), C.POINTER(i), C.POINTER(i)]
    lib.vsb_conv3d_plan_flops.argtypes = [vp]
    lib.vsb_conv3d_plan_flops.restype = C.c_double
    lib.vsb_bottleneck_plan_create.argtypes = [C.POINTER(BottleneckDesc), C.POINTER(vp)]
    lib.vsb_bottleneck_run.argtypes = [vp, vp]
    lib.vsb_bottleneck_plan_destroy.argtypes = [vp]
    lib.vsb_bottleneck_plan_destroy.restype = None
    lib.vsb_bottleneck_plan_info.argtypes = [vp, C.POINTER(C.c_longlong)]
    lib.vsb_maxpool3d.argtypes = [vp, i, i, i, i, i, i, vp, i, i, i, i, i, i, i, i, i, i, i, i, vp]
    lib.vsb_global_avgpool.argtypes = [vp, i, i, i, i, f32p, i, i, i, vp]
    lib.vsb_linear.argtypes = [f32p, i, i, f32p, f32p, f32p, i, i, vp]
    lib.vsb_softmax_topk.argtypes = [f32p, i, i, i, i, vp, f32p, vp]
    lib.vsb_nonlocal_attention.argtypes = [vp, i, vp, i, vp, i, vp, i, i, i, i, i, i, i, vp]
    lib.vsb_score_rows.argtypes = [vp, ll, i, i, i, i, vp]
    lib.vsb_transpose_pad.argtypes = [vp, i, vp, i, i, i, i, vp]
    lib.vsb_nthwc_to_ncthw_f32.argtypes = [vp, i, i, i, i, f32p, i, vp]
    lib.vsb_ncthw_f32_to_nthwc.argtypes = [f32p, i, i, ll, i, vp, i, i, i, i, vp]
    lib.vsb_debug_im2col_probe.argtypes = [vp] + [i] * 25 + [vp, vp]
    lib.vsb_debug_umma_semantics.argtypes = [vp, i, vp, i, i, i, i, i, i, vp, vp]
    lib.vsb_debug_umma_rate.argtypes = [i, i, i, i, i, i, vp, vp]
    lib.vsb_debug_umma_rate2.argtypes = [i, i, i, i, i, i, i, i, vp, vp]
    lib.vsb_debug_conv_stats.argtypes = [vp, C.POINTER(C.c_longlong)]
    lib.vsb_debug_conv_plan_info.argtypes = [vp, C.POINTER(C.c_longlong)]
    lib.vsb_debug_tma_rate.argtypes = [vp, i, i, i, i, i, i, i, i, i, i, i, vp, vp]
    lib.vsb_debug_bottleneck_stats.argtypes = [vp, C.POINTER(C.c_longlong)]
    for name in ("vsb_pack_frames", "vsb_bottleneck_plan_create", "vsb_bottleneck_run", "vsb_bottleneck_plan_info",
                 "vsb_conv3d_plan_create", "vsb_conv3d_run", "vsb_conv3d_plan_out_shape",
                 "vsb_maxpool3d", "vsb_global_avgpool", "vsb_linear", "vsb_softmax_topk", "vsb_nonlocal_attention",
                 "vsb_score_rows", "vsb_transpose_pad",
                 "vsb_nthwc_to_ncthw_f32", "vsb_ncthw_f32_to_nthwc", "vsb_debug_im2col_probe",
                 "vsb_debug_umma_semantics", "vsb_debug_umma_rate", "vsb_debug_umma_rate2", "vsb_debug_conv_stats", "vsb_debug_conv_plan_info",
                 "vsb_debug_bottleneck_stats", "vsb_debug_tma_rate"):
        getattr(lib, name).restype = i
    # clip programs (ABI v7): every vsb_program_add_<op> takes (program, <the op's arguments without the stream>,
    # lane, name)
    cp, ull = C.c_char_p, C.c_ulonglong
    lib.vsb_program_create.argtypes = [C.POINTER(vp)]
    lib.vsb_program_destroy.argtypes = [vp]
    lib.vsb_program_destroy.restype = None
    lib.vsb_program_add_region.argtypes = [vp, cp, vp, ull, i]
    lib.vsb_program_region.argtypes = [vp, cp, C.POINTER(vp), C.POINTER(ull)]
    lib.vsb_program_num_ops.argtypes = [vp]
    lib.vsb_program_num_launches.argtypes = [vp]
    lib.vsb_program_device_bytes.argtypes = [vp]
    lib.vsb_program_device_bytes.restype = ull
    lib.vsb_program_add_conv.argtypes = [vp, vp, i, cp]
    lib.vsb_program_add_bottleneck.argtypes = [vp, vp, C.POINTER(BottleneckDesc), i, cp]
    for op in ("pack_frames", "maxpool3d", "global_avgpool", "linear", "nonlocal_attention", "score_rows",
               "transpose_pad"):
        getattr(lib, "vsb_program_add_" + op).argtypes = [vp] + list(getattr(lib, "vsb_" + op).argtypes[:-1]) + [i, cp]
    lib.vsb_stem_pool_plan_create.argtypes = [C.POINTER(StemPoolDesc), C.POINTER(vp)]
    lib.vsb_stem_pool_run.argtypes = [vp, vp]
    lib.vsb_stem_pool_plan_destroy.argtypes = [vp]
    lib.vsb_stem_pool_plan_destroy.restype = None
    lib.vsb_stem_pool_plan_desc.argtypes = [vp, C.POINTER(StemPoolDesc)]
    lib.vsb_program_add_stem_pool.argtypes = [vp, vp, i, cp]
    lib.vsb_jpeg_info.argtypes = [vp, ull, C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(i)]
    lib.vsb_jpeg_decoder_create.argtypes = [i, i, C.POINTER(vp)]
    lib.vsb_jpeg_decoder_destroy.argtypes = [vp]
    lib.vsb_jpeg_decoder_destroy.restype = None
    lib.vsb_jpeg_decode_resize.argtypes = [vp, vp, ull, vp, i, i, vp]
    lib.vsb_resize_bicubic_u8.argtypes = [vp, vp, i, i, vp, i, i, vp]
    lib.vsb_jpeg_batch_create.argtypes = [C.POINTER(vp)]
    lib.vsb_jpeg_batch_destroy.argtypes = [vp]
    lib.vsb_jpeg_batch_destroy.restype = None
    lib.vsb_jpeg_batch_decode_resize.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(ull), i, C.POINTER(vp), i, i,
                                                 C.POINTER(i), vp]
    lib.vsb_program_add_sync.argtypes = [vp, i, i]
    lib.vsb_program_run.argtypes = [vp, vp]
    lib.vsb_program_capture.argtypes = [vp, vp]
    lib.vsb_program_save.argtypes = [vp, cp]
    lib.vsb_program_file_device_bytes.argtypes = [cp, C.POINTER(ull)]
    lib.vsb_program_load.argtypes = [cp, vp, ull, C.POINTER(vp)]
    for name in ("vsb_jpeg_batch_create", "vsb_jpeg_batch_decode_resize", "vsb_jpeg_info", "vsb_jpeg_decoder_create", "vsb_jpeg_decode_resize", "vsb_resize_bicubic_u8",
                 "vsb_stem_pool_plan_create", "vsb_stem_pool_run", "vsb_stem_pool_plan_desc", "vsb_program_add_stem_pool",
                 "vsb_program_create", "vsb_program_add_region", "vsb_program_region", "vsb_program_num_ops",
                 "vsb_program_num_launches", "vsb_program_add_conv", "vsb_program_add_bottleneck",
                 "vsb_program_add_pack_frames", "vsb_program_add_maxpool3d", "vsb_program_add_global_avgpool",
                 "vsb_program_add_linear", "vsb_program_add_nonlocal_attention", "vsb_program_add_score_rows",
                 "vsb_program_add_transpose_pad", "vsb_program_add_sync", "vsb_program_run", "vsb_program_capture",
                 "vsb_program_save", "vsb_program_file_device_bytes", "vsb_program_load"):
        getattr(lib, name).restype = i
    if lib.vsb_abi_version() != 8:
        raise VsbError(f"ABI mismatch: library reports {lib.vsb_abi_version()}, binding expects 8")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().vsb_last_error().decode("utf-8", "replace")
        raise VsbError(f"{what} failed (status {rc}): {msg}")


def launch_count() -> int:
    return int(load().vsb_launch_count())
