"""Frame ingest on the GPU: JPEG bytes -> uint8 [S, S, 3] in device memory, bit-identical
to the reference's `VsituDS.read_img` (vidsitu_code/dat_loader.py:183-191: PIL open, RGB, `resize((224, 224))`).

Host side of `vsb_jpeg_*` (include/vidsitu_b200.h): the library decodes the Huffman segment on the host and runs
dequantisation, IDCT, chroma upsampling, colour conversion and Pillow's BICUBIC resampling as CUDA kernels.  There is
no fallback inside: a file the GPU path does not decode (progressive, CMYK, ...) raises `VsbError`; callers that want
the reference's PIL reader for such files catch it (frames_io.load_video_device does, and counts them)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import lib as _l
from .lib import VsbError, check


def jpeg_info(data: bytes) -> Tuple[int, int, int, int, int]:
    """(width, height, components, h_samp, v_samp) of a JPEG the GPU path can decode (no GPU needed)."""
    w, h, n, hs, vs = (C.c_int() for _ in range(5))
    buf = (C.c_ubyte * len(data)).from_buffer_copy(data)
    check(_l.load().vsb_jpeg_info(buf, len(data), C.byref(w), C.byref(h), C.byref(n), C.byref(hs), C.byref(vs)),
          "vsb_jpeg_info")
    return w.value, h.value, n.value, hs.value, vs.value


class JpegDecoder:
    """One decoder = one set of workspaces (pinned staging + device planes) for frames up to max_width x max_height.
    Not thread-safe: one per host thread."""

    def __init__(self, max_width: int = 1920, max_height: int = 1088, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise VsbError("JpegDecoder needs a CUDA device (the host-only reader is frames_io.read_img)")
        self.device = torch.device(device if device is not None else "cuda")
        self._lib = _l.load()
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self._lib.vsb_jpeg_decoder_create(max_width, max_height, C.byref(self._h)), "vsb_jpeg_decoder_create")

    def decode_resize(self, data: bytes, out: torch.Tensor) -> None:
        """Decode `data` into `out` (CUDA uint8 [S_h, S_w, 3], contiguous) on the current stream."""
        if not out.is_cuda or out.dtype != torch.uint8 or out.dim() != 3 or out.shape[2] != 3 or not out.is_contiguous():
            raise VsbError("out must be a contiguous CUDA uint8 [h, w, 3] tensor")
        buf = (C.c_ubyte * len(data)).from_buffer_copy(data)
        with torch.cuda.device(self.device):
            check(self._lib.vsb_jpeg_decode_resize(self._h, buf, len(data), out.data_ptr(), out.shape[0], out.shape[1],
                                                   torch.cuda.current_stream().cuda_stream), "vsb_jpeg_decode_resize")

    def resize(self, img: torch.Tensor, out: torch.Tensor) -> None:
        """PIL-exact BICUBIC resize of a CUDA uint8 [h, w, 3] image into `out` [S_h, S_w, 3]."""
        for t in (img, out):
            if not t.is_cuda or t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3 or not t.is_contiguous():
                raise VsbError("resize takes contiguous CUDA uint8 [h, w, 3] tensors")
        with torch.cuda.device(self.device):
            check(self._lib.vsb_resize_bicubic_u8(self._h, img.data_ptr(), img.shape[0], img.shape[1], out.data_ptr(),
                                                  out.shape[0], out.shape[1], torch.cuda.current_stream().cuda_stream),
                  "vsb_resize_bicubic_u8")

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._lib.vsb_jpeg_decoder_destroy(h)
            h.value = None
