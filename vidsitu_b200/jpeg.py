"""Frame ingest on the GPU: JPEG bytes -> uint8 [S, S, 3] in device memory, bit-identical
to the reference's `VsituDS.read_img` (vidsitu_code/dat_loader.py:183-191: PIL open, RGB, `resize((224, 224))`).

Host side of `vsb_jpeg_*` (include/vidsitu_b200.h): the library decodes the Huffman segment on the host and runs
dequantisation, IDCT, chroma upsampling, colour conversion and Pillow's BICUBIC resampling as CUDA kernels.  There is
no fallback inside: a file the GPU path does not decode (progressive, CMYK, ...) raises `VsbError`; callers that want
the reference's PIL reader for such files catch it (frames_io.load_video_device does, and counts them)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

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


class JpegBatchDecoder:
    """The fully-on-device path (vsb_jpeg_batch_*): the Huffman segments of a whole batch of files are decoded on the
    GPU, one warp per frame; the host parses headers and stages bytes.  Same pixels as JpegDecoder / Pillow."""

    def __init__(self, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise VsbError("JpegBatchDecoder needs a CUDA device (the host-only reader is frames_io.read_img)")
        self.device = torch.device(device if device is not None else "cuda")
        self._lib = _l.load()
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self._lib.vsb_jpeg_batch_create(C.byref(self._h)), "vsb_jpeg_batch_create")

    def decode_resize(self, datas: Sequence[bytes], outs: Sequence[torch.Tensor]) -> List[bool]:
        """Decode datas[i] into outs[i] (CUDA uint8 [h, w, 3] views of one size, each contiguous).  Returns per-file
        success; a refused / corrupt file leaves its output untouched.  Completes before returning."""
        n = len(datas)
        if n == 0:
            return []
        if len(outs) != n:
            raise VsbError("one output per file")
        oh, ow = int(outs[0].shape[0]), int(outs[0].shape[1])
        for t in outs:
            if (not t.is_cuda or t.dtype != torch.uint8 or tuple(t.shape) != (oh, ow, 3) or not t.is_contiguous()):
                raise VsbError("outs must be contiguous CUDA uint8 [h, w, 3] tensors of one size")
        ptrs = (C.c_char_p * n)(*datas)
        sizes = (C.c_ulonglong * n)(*[len(d) for d in datas])
        dsts = (C.c_void_p * n)(*[t.data_ptr() for t in outs])
        status = (C.c_int * n)()
        with torch.cuda.device(self.device):
            check(self._lib.vsb_jpeg_batch_decode_resize(self._h, ptrs, sizes, n, dsts, oh, ow, status,
                                                         torch.cuda.current_stream().cuda_stream),
                  "vsb_jpeg_batch_decode_resize")
        return [status[i] == 0 for i in range(n)]

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._lib.vsb_jpeg_batch_destroy(h)
            h.value = None
