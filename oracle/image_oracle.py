"""ORACLE -- test infrastructure only.  Never imported by the product (vidsitu_b200/).

CPU restatement (numpy, integer arithmetic) of the frame ingest in front of the pack kernel
(SURVEY.md section 8 row f3): `VsituDS.read_img` (vidsitu_code/dat_loader.py:183-191)

    img = Image.open(img_fpath).convert("RGB");  img = img.resize((224, 224));  np.array(img)

The algorithms live in third-party code that is not vendored under /root/reference:
  * `Image.resize` -- Pillow (the reference pins pillow=7.2.0, vsitu_pyt_env.yml:151): default filter BICUBIC
    (a = -0.5), `src/libImaging/Resample.c`: per output pixel a window of `support * max(scale, 1)` source pixels,
    double-precision weights normalised to sum 1, rounded to 22-bit fixed point, a horizontal pass into an 8-bit
    intermediate image, then a vertical pass; each pass rounds with +2^21 and clips to [0, 255].  Restated here from
    the published algorithm.
Pinned by tests/test_io.py against Pillow itself (the installed version, live) on synthetic images.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2
BICUBIC_SUPPORT = 2.0


def bicubic_filter(x: float) -> float:
    """Keys' cubic convolution kernel with a = -0.5, evaluated as Resample.c does (same operation order)."""
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int) -> Tuple[List[Tuple[int, int]], np.ndarray]:
    """(bounds, kk): bounds[xx] = (first source index, count); kk[xx, :count] = 22-bit fixed-point weights
    (Resample.c precompute_coeffs + normalize_coeffs_8bpc, box = the whole image)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = BICUBIC_SUPPORT * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = []
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds.append((xmin, xmax))
    return bounds, kk


def _pass(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One resampling pass along `axis` (1 = horizontal, 0 = vertical) of an [H, W, C] uint8 image."""
    bounds, kk = precompute_coeffs(img.shape[axis], out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for xx, (xmin, cnt) in enumerate(bounds):
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for k in range(cnt):
            acc += src[xmin + k] * int(kk[xx, k])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bicubic_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """`PIL.Image.resize((out_w, out_h))` (default BICUBIC) of an [H, W, 3] uint8 image: horizontal pass first, then
    vertical, each only when the size along it changes (ImagingResample)."""
    if img.dtype != np.uint8 or img.ndim != 3:
        raise ValueError("expected an [H, W, C] uint8 image")
    out = img
    if out.shape[1] != out_w:
        out = _pass(out, out_w, 1)
    if out.shape[0] != out_h:
        out = _pass(out, out_h, 0)
    return np.ascontiguousarray(out)


# ----------------------------------------------------------------------------------------------- JPEG decode
# `Image.open(path).convert("RGB")` -- Pillow hands the file to libjpeg(-turbo) with its defaults: baseline Huffman
# decode, dequantisation, the "islow" integer inverse DCT (jidctint.c, CONST_BITS = 13, PASS1_BITS = 2), "fancy"
# (triangle-filter) chroma upsampling (jdsample.c: h2v1 / h2v2), integer YCbCr -> RGB tables (jdcolor.c, 16-bit
# fixed point).  libjpeg-turbo's SIMD paths are bit-exact with these C routines.  Restated from the published
# algorithms (ITU T.81 + the IJG code's documented arithmetic); frames of the dataset are ffmpeg `-q:v 1` MJPEG
# (prep_data/dwn_yt.py:229-250): baseline, YCbCr 4:2:0, no restart markers.
ZIGZAG = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14,
          21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53,
          60, 61, 54, 47, 55, 62, 63]


class JpegUnsupported(ValueError):
    pass


def parse_jpeg(data: bytes) -> dict:
    """Headers of a baseline JPEG: frame geometry, quantisation tables (natural order), Huffman tables, scan."""
    if data[:2] != b"\xff\xd8":
        raise JpegUnsupported("not a JPEG (no SOI)")
    pos = 2
    info = {"qt": {}, "dc": {}, "ac": {}, "restart": 0}
    while True:
        if data[pos] != 0xFF:
            raise JpegUnsupported("marker expected")
        while data[pos + 1] == 0xFF:
            pos += 1
        m = data[pos + 1]
        pos += 2
        if m in (0xD8, 0x01) or 0xD0 <= m <= 0xD7:
            continue
        ln = (data[pos] << 8) | data[pos + 1]
        seg = data[pos + 2: pos + ln]
        if m == 0xDB:
            i = 0
            while i < len(seg):
                pq, tq = seg[i] >> 4, seg[i] & 15
                i += 1
                tab = [0] * 64
                for k in range(64):
                    if pq:
                        v = (seg[i] << 8) | seg[i + 1]
                        i += 2
                    else:
                        v = seg[i]
                        i += 1
                    tab[ZIGZAG[k]] = v
                info["qt"][tq] = tab
        elif m in (0xC0, 0xC1):
            if seg[0] != 8:
                raise JpegUnsupported("only 8-bit samples")
            info["h"], info["w"] = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4]
            info["comps"] = [{"id": seg[6 + 3 * c], "hs": seg[7 + 3 * c] >> 4, "vs": seg[7 + 3 * c] & 15,
                              "tq": seg[8 + 3 * c]} for c in range(seg[5])]
        elif 0xC2 <= m <= 0xCF and m not in (0xC4, 0xC8, 0xCC):
            raise JpegUnsupported("only baseline / extended sequential Huffman JPEGs (SOF0 / SOF1)")
        elif m == 0xC4:
            i = 0
            while i < len(seg):
                tc, th = seg[i] >> 4, seg[i] & 15
                counts = list(seg[i + 1: i + 17])
                n = sum(counts)
                info["ac" if tc else "dc"][th] = (counts, list(seg[i + 17: i + 17 + n]))
                i += 17 + n
        elif m == 0xDD:
            info["restart"] = (seg[0] << 8) | seg[1]
        elif m == 0xDA:
            ns = seg[0]
            if ns != len(info["comps"]):
                raise JpegUnsupported("only single-scan (interleaved) JPEGs")
            for s in range(ns):
                cid, tabs = seg[1 + 2 * s], seg[2 + 2 * s]
                comp = next(c for c in info["comps"] if c["id"] == cid)
                comp["td"], comp["ta"] = tabs >> 4, tabs & 15
            info["scan_start"] = pos + ln
            return info
        elif m == 0xD9:
            raise JpegUnsupported("no scan")
        pos += ln


def _huff_lookup(counts, symbols):
    """code -> symbol maps per length (T.81 annex C)."""
    table, code, k = {}, 0, 0
    for length in range(1, 17):
        for _ in range(counts[length - 1]):
            table[(length, code)] = symbols[k]
            code += 1
            k += 1
        code <<= 1
    return table


def entropy_decode(data: bytes, info: dict) -> List[np.ndarray]:
    """Quantised DCT coefficients per component: int16 [blocks_h, blocks_w, 64] in natural order, the MCU-padded
    block grid (T.81 annex F.2.2)."""
    comps = info["comps"]
    hmax, vmax = max(c["hs"] for c in comps), max(c["vs"] for c in comps)
    mcux, mcuy = -(-info["w"] // (8 * hmax)), -(-info["h"] // (8 * vmax))
    out = [np.zeros((mcuy * c["vs"], mcux * c["hs"], 64), dtype=np.int16) for c in comps]
    dc_tabs = {k: _huff_lookup(*v) for k, v in info["dc"].items()}
    ac_tabs = {k: _huff_lookup(*v) for k, v in info["ac"].items()}
    pos, bitbuf, nbits = info["scan_start"], 0, 0

    def getbit():
        nonlocal pos, bitbuf, nbits
        if nbits == 0:
            b = data[pos] if pos < len(data) else 0
            pos += 1
            if b == 0xFF:
                nxt = data[pos] if pos < len(data) else 0
                if nxt == 0:
                    pos += 1
                else:                      # a marker inside the scan: feed zeros (libjpeg does the same)
                    pos -= 1
                    b = 0
            bitbuf, nbits = b, 8
        nbits -= 1
        return (bitbuf >> nbits) & 1

    def decode(tab):
        code = 0
        for length in range(1, 17):
            code = (code << 1) | getbit()
            s = tab.get((length, code))
            if s is not None:
                return s
        raise JpegUnsupported("bad Huffman code")

    def receive_extend(s):
        v = 0
        for _ in range(s):
            v = (v << 1) | getbit()
        return v if v >= (1 << (s - 1)) else v - (1 << s) + 1

    pred = [0] * len(comps)
    restart, count = info["restart"], 0
    for my in range(mcuy):
        for mx in range(mcux):
            if restart and count and count % restart == 0:
                nbits = 0                                   # byte-align, skip the RSTn marker, reset predictors
                while not (data[pos] == 0xFF and 0xD0 <= data[pos + 1] <= 0xD7):
                    pos += 1
                pos += 2
                pred = [0] * len(comps)
            count += 1
            for ci, c in enumerate(comps):
                for by in range(c["vs"]):
                    for bx in range(c["hs"]):
                        blk = out[ci][my * c["vs"] + by, mx * c["hs"] + bx]
                        s = decode(dc_tabs[c["td"]])
                        pred[ci] += receive_extend(s) if s else 0
                        blk[0] = pred[ci]
                        k = 1
                        while k < 64:
                            rs = decode(ac_tabs[c["ta"]])
                            r, s = rs >> 4, rs & 15
                            if s == 0:
                                if r != 15:
                                    break
                                k += 16
                                continue
                            k += r
                            blk[ZIGZAG[k]] = receive_extend(s)
                            k += 1
    return out


_F = dict(c0298=2446, c0390=3196, c0541=4433, c0765=6270, c0899=7373, c1175=9633, c1501=12299, c1847=15137,
          c1961=16069, c2053=16819, c2562=20995, c3072=25172)


def _idct_1d(x, shift):
    """One pass of jpeg_idct_islow over the LAST axis of an int64 array [..., 8]."""
    f = _F
    z2, z3 = x[..., 2], x[..., 6]
    z1 = (z2 + z3) * f["c0541"]
    tmp2 = z1 + z3 * (-f["c1847"])
    tmp3 = z1 + z2 * f["c0765"]
    z2, z3 = x[..., 0], x[..., 4]
    tmp0 = (z2 + z3) << 13
    tmp1 = (z2 - z3) << 13
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    tmp0, tmp1, tmp2, tmp3 = x[..., 7], x[..., 5], x[..., 3], x[..., 1]
    z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
    z5 = (z3 + z4) * f["c1175"]
    tmp0, tmp1, tmp2, tmp3 = tmp0 * f["c0298"], tmp1 * f["c2053"], tmp2 * f["c3072"], tmp3 * f["c1501"]
    z1, z2, z3, z4 = z1 * -f["c0899"], z2 * -f["c2562"], z3 * -f["c1961"], z4 * -f["c0390"]
    z3, z4 = z3 + z5, z4 + z5
    tmp0, tmp1, tmp2, tmp3 = tmp0 + z1 + z3, tmp1 + z2 + z4, tmp2 + z2 + z3, tmp3 + z1 + z4
    rnd = 1 << (shift - 1)
    return np.stack([(tmp10 + tmp3 + rnd) >> shift, (tmp11 + tmp2 + rnd) >> shift, (tmp12 + tmp1 + rnd) >> shift,
                     (tmp13 + tmp0 + rnd) >> shift, (tmp13 - tmp0 + rnd) >> shift, (tmp12 - tmp1 + rnd) >> shift,
                     (tmp11 - tmp2 + rnd) >> shift, (tmp10 - tmp3 + rnd) >> shift], axis=-1)


def range_limit(v):
    """libjpeg's post-IDCT range-limit table (index masked to 10 bits, centred on 128)."""
    v = v & 1023
    return np.where(v < 128, v + 128, np.where(v < 512, 255, np.where(v < 896, 0, v - 896))).astype(np.uint8)


def idct_islow(coefs: np.ndarray, qt) -> np.ndarray:
    """[bh, bw, 64] quantised coefficients -> [bh * 8, bw * 8] uint8 samples (dequantise + jpeg_idct_islow)."""
    bh, bw, _ = coefs.shape
    x = coefs.astype(np.int64) * np.asarray(qt, dtype=np.int64)
    x = x.reshape(bh, bw, 8, 8)                                  # [.., row, col]
    ws = _idct_1d(x.transpose(0, 1, 3, 2), 13 - 2).transpose(0, 1, 3, 2)     # pass 1: down the columns
    px = range_limit(_idct_1d(ws, 13 + 2 + 3))                                # pass 2: along the rows
    return px.transpose(0, 2, 1, 3).reshape(bh * 8, bw * 8)


def upsample_h2v1_fancy(plane: np.ndarray, width: int) -> np.ndarray:
    """jdsample.c h2v1_fancy_upsample on the first `width` columns: [rows, width] -> [rows, 2 * width]."""
    p = plane[:, :width].astype(np.int32)
    left = np.concatenate([p[:, :1], p[:, :-1]], axis=1)
    right = np.concatenate([p[:, 1:], p[:, -1:]], axis=1)
    out = np.empty((p.shape[0], 2 * width), dtype=np.int32)
    out[:, 0::2] = (p * 3 + left + 1) >> 2
    out[:, 1::2] = (p * 3 + right + 2) >> 2
    out[:, 0] = p[:, 0]
    out[:, -1] = p[:, -1]
    return out.astype(np.uint8)


def upsample_h2v2_fancy(plane: np.ndarray, width: int, height: int) -> np.ndarray:
    """jdsample.c h2v2_fancy_upsample: [>= height, >= width] -> [2 * height, 2 * width]; the row above the first /
    below the last REAL row is that row itself (jdmainct.c context rows)."""
    p = plane[:height, :width].astype(np.int32)
    up = np.concatenate([p[:1], p[:-1]], axis=0)
    dn = np.concatenate([p[1:], p[-1:]], axis=0)
    out = np.empty((2 * height, 2 * width), dtype=np.int32)
    for v, other in ((0, up), (1, dn)):
        colsum = p * 3 + other                                   # thiscolsum of every column
        last = np.concatenate([colsum[:, :1], colsum[:, :-1]], axis=1)
        nxt = np.concatenate([colsum[:, 1:], colsum[:, -1:]], axis=1)
        even = (colsum * 3 + last + 8) >> 4
        odd = (colsum * 3 + nxt + 7) >> 4
        even[:, 0] = (colsum[:, 0] * 4 + 8) >> 4
        odd[:, -1] = (colsum[:, -1] * 4 + 7) >> 4
        out[v::2, 0::2] = even
        out[v::2, 1::2] = odd
    return out.astype(np.uint8)


def ycc_to_rgb(y: np.ndarray, cb: np.ndarray, cr: np.ndarray) -> np.ndarray:
    """jdcolor.c ycc_rgb_convert: 16-bit fixed-point tables."""
    def fix(v):
        return int(v * 65536 + 0.5)
    xcb, xcr = cb.astype(np.int64) - 128, cr.astype(np.int64) - 128
    half = 1 << 15
    r = y.astype(np.int64) + ((fix(1.40200) * xcr + half) >> 16)
    b = y.astype(np.int64) + ((fix(1.77200) * xcb + half) >> 16)
    g = y.astype(np.int64) + ((-fix(0.34414) * xcb + half - fix(0.71414) * xcr) >> 16)
    return np.clip(np.stack([r, g, b], axis=-1), 0, 255).astype(np.uint8)


def decode_jpeg(data: bytes) -> np.ndarray:
    """[H, W, 3] uint8 RGB, as `np.array(Image.open(..).convert("RGB"))`."""
    info = parse_jpeg(data)
    coefs = entropy_decode(data, info)
    comps, h, w = info["comps"], info["h"], info["w"]
    planes = [idct_islow(c, info["qt"][comp["tq"]]) for c, comp in zip(coefs, comps)]
    if len(comps) == 1:
        g = planes[0][:h, :w]
        return np.stack([g, g, g], axis=-1)
    if len(comps) != 3:
        raise JpegUnsupported("1 or 3 components")
    hmax, vmax = comps[0]["hs"], comps[0]["vs"]
    if any((c["hs"], c["vs"]) != (1, 1) for c in comps[1:]) or (hmax, vmax) not in ((1, 1), (2, 1), (2, 2)):
        raise JpegUnsupported("sampling factors other than 4:4:4 / 4:2:2 / 4:2:0")
    cw, chh = -(-w // hmax), -(-h // vmax)
    up = []
    for p in planes[1:]:
        if (hmax, vmax) == (1, 1):
            up.append(p)
        elif (hmax, vmax) == (2, 1):
            up.append(upsample_h2v1_fancy(p[:h], cw))
        else:
            up.append(upsample_h2v2_fancy(p, cw, chh))
    return ycc_to_rgb(planes[0][:h, :w], up[0][:h, :w], up[1][:h, :w])


def read_img(data: bytes, size: int = 224) -> np.ndarray:
    """VsituDS.read_img (dat_loader.py:183-191) on the bytes of a JPEG file."""
    return resize_bicubic_u8(decode_jpeg(data), size, size)
