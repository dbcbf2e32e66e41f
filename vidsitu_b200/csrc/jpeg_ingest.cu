// Frame ingest on the GPU: JPEG file bytes -> uint8 [S, S, 3] RGB in device memory,
// bit-identical to the reference's reader
//     Image.open(path).convert("RGB").resize((224, 224))          vidsitu_code/dat_loader.py:183-191
// for the frames the dataset holds (ffmpeg -q:v 1 MJPEG: baseline Huffman, YCbCr 4:2:0; prep_data/dwn_yt.py:229-250;
// 4:4:4, 4:2:2 and grayscale are decoded too).
//
// Split: the entropy-coded segment is sequential by construction (every Huffman code starts where the previous one
// ended), so it is decoded on the HOST by a table-driven decoder into quantised coefficient blocks in pinned memory;
// everything that touches pixels runs on the device - dequantisation + the "islow" integer inverse DCT of libjpeg
// (jidctint.c: CONST_BITS 13, PASS1_BITS 2), "fancy" triangle-filter chroma upsampling (jdsample.c), the fixed-point
// YCbCr -> RGB tables (jdcolor.c), and Pillow's two-pass fixed-point BICUBIC resampling (libImaging/Resample.c:
// 22-bit weights, 8-bit intermediate image).  All of it is integer arithmetic: the result equals Pillow's bit for bit.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "common.h"

using namespace vsb;

namespace {

const unsigned char kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                   41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                   30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// ------------------------------------------------------------------------------------------------ host: parse
struct HuffTable {
  bool defined = false;
  // 9-bit lookahead: (length << 8) | symbol for codes of at most 9 bits, 0 otherwise
  unsigned short look[512];
  // canonical decoding for longer codes (T.81 annex F.2.2.3)
  int maxcode[18];
  int valptr[17];
  int mincode[17];
  unsigned char symbols[256];
};

struct Component {
  int id = 0, hs = 1, vs = 1, tq = 0, td = 0, ta = 0;
};

struct JpegHeader {
  int width = 0, height = 0, ncomp = 0;
  Component comp[3];
  unsigned short qt[4][64];  // natural order
  bool qt_defined[4] = {false, false, false, false};
  HuffTable dc[4], ac[4];
  int restart = 0;
  size_t scan_start = 0;
  int hmax = 1, vmax = 1, mcux = 0, mcuy = 0;
};

bool build_huff(HuffTable& t, const unsigned char* counts, const unsigned char* syms, int nsyms) {
  if (nsyms > 256) return false;
  memcpy(t.symbols, syms, (size_t)nsyms);
  memset(t.look, 0, sizeof(t.look));
  int code = 0, k = 0;
  for (int len = 1; len <= 16; ++len) {
    t.valptr[len] = k;
    t.mincode[len] = code;
    for (int i = 0; i < counts[len - 1]; ++i, ++k, ++code) {
      if (len <= 9) {
        const int first = code << (9 - len), n = 1 << (9 - len);
        for (int j = 0; j < n; ++j) t.look[first + j] = (unsigned short)((len << 8) | syms[k]);
      }
    }
    t.maxcode[len] = counts[len - 1] ? code - 1 : -1;
    if (code > (1 << len)) return false;
    code <<= 1;
  }
  t.maxcode[17] = 0x7fffffff;
  t.defined = true;
  return true;
}

int parse_header(const unsigned char* d, size_t n, JpegHeader& h) {
#define JP_FAIL(...)        \
  do {                      \
    set_error(__VA_ARGS__); \
    return VSB_ERR_INVALID; \
  } while (0)
  if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) JP_FAIL("not a JPEG (no SOI marker)");
  size_t pos = 2;
  bool have_frame = false;
  for (;;) {
    if (pos + 4 > n) JP_FAIL("truncated JPEG headers");
    if (d[pos] != 0xFF) JP_FAIL("JPEG marker expected at byte %zu", pos);
    while (pos + 1 < n && d[pos + 1] == 0xFF) ++pos;
    const int m = d[pos + 1];
    pos += 2;
    if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
    if (m == 0xD9) JP_FAIL("JPEG has no scan");
    if (pos + 2 > n) JP_FAIL("truncated JPEG headers");
    const size_t len = ((size_t)d[pos] << 8) | d[pos + 1];
    if (len < 2 || pos + len > n) JP_FAIL("truncated JPEG segment");
    const unsigned char* s = d + pos + 2;
    const size_t sl = len - 2;
    if (m == 0xDB) {
      size_t i = 0;
      while (i < sl) {
        const int pq = s[i] >> 4, tq = s[i] & 15;
        ++i;
        if (tq > 3 || i + (pq ? 128 : 64) > sl) JP_FAIL("bad quantisation table");
        for (int k = 0; k < 64; ++k) {
          h.qt[tq][kZigzag[k]] = pq ? (unsigned short)((s[i] << 8) | s[i + 1]) : s[i];
          i += pq ? 2 : 1;
        }
        h.qt_defined[tq] = true;
      }
    } else if (m == 0xC0 || m == 0xC1) {
      if (sl < 6 || s[0] != 8) JP_FAIL("only 8-bit JPEGs are decoded");
      h.height = (s[1] << 8) | s[2];
      h.width = (s[3] << 8) | s[4];
      h.ncomp = s[5];
      if ((h.ncomp != 1 && h.ncomp != 3) || sl < (size_t)(6 + 3 * h.ncomp)) JP_FAIL("only 1- and 3-component JPEGs are decoded");
      for (int c = 0; c < h.ncomp; ++c) {
        h.comp[c].id = s[6 + 3 * c];
        h.comp[c].hs = s[7 + 3 * c] >> 4;
        h.comp[c].vs = s[7 + 3 * c] & 15;
        h.comp[c].tq = s[8 + 3 * c];
        if (h.comp[c].tq > 3) JP_FAIL("bad quantisation table index");
      }
      have_frame = true;
    } else if (m >= 0xC2 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
      JP_FAIL("only baseline / extended-sequential Huffman JPEGs are decoded on the GPU path (SOF%d found)", m - 0xC0);
    } else if (m == 0xC4) {
      size_t i = 0;
      while (i < sl) {
        if (i + 17 > sl) JP_FAIL("bad Huffman table");
        const int tc = s[i] >> 4, th = s[i] & 15;
        int nsyms = 0;
        for (int k = 0; k < 16; ++k) nsyms += s[i + 1 + k];
        if (th > 3 || tc > 1 || i + 17 + nsyms > sl) JP_FAIL("bad Huffman table");
        if (!build_huff(tc ? h.ac[th] : h.dc[th], s + i + 1, s + i + 17, nsyms)) JP_FAIL("bad Huffman table");
        i += 17 + (size_t)nsyms;
      }
    } else if (m == 0xDD) {
      if (sl < 2) JP_FAIL("bad DRI segment");
      h.restart = (s[0] << 8) | s[1];
    } else if (m == 0xDA) {
      if (!have_frame) JP_FAIL("scan before frame header");
      if (sl < 1 || s[0] != h.ncomp || sl < (size_t)(1 + 2 * h.ncomp)) JP_FAIL("only single-scan (interleaved) JPEGs are decoded");
      for (int k = 0; k < h.ncomp; ++k) {
        const int cid = s[1 + 2 * k], tabs = s[2 + 2 * k];
        int c = -1;
        for (int j = 0; j < h.ncomp; ++j)
          if (h.comp[j].id == cid) c = j;
        if (c < 0) JP_FAIL("scan names an unknown component");
        h.comp[c].td = tabs >> 4;
        h.comp[c].ta = tabs & 15;
        if (h.comp[c].td > 3 || h.comp[c].ta > 3 || !h.dc[h.comp[c].td].defined || !h.ac[h.comp[c].ta].defined)
          JP_FAIL("scan uses an undefined Huffman table");
        if (!h.qt_defined[h.comp[c].tq]) JP_FAIL("component uses an undefined quantisation table");
      }
      h.scan_start = pos + len;
      break;
    }
    pos += len;
  }
  if (h.width <= 0 || h.height <= 0) JP_FAIL("empty JPEG frame");
  h.hmax = h.vmax = 1;
  for (int c = 0; c < h.ncomp; ++c) {
    if (h.comp[c].hs < 1 || h.comp[c].vs < 1) JP_FAIL("bad sampling factors");
    if (h.comp[c].hs > h.hmax) h.hmax = h.comp[c].hs;
    if (h.comp[c].vs > h.vmax) h.vmax = h.comp[c].vs;
  }
  if (h.ncomp == 1) {
    h.comp[0].hs = h.comp[0].vs = 1;  // a single-component scan is never interleaved: one block per MCU
    h.hmax = h.vmax = 1;
  } else {
    const bool chroma11 = h.comp[1].hs == 1 && h.comp[1].vs == 1 && h.comp[2].hs == 1 && h.comp[2].vs == 1;
    const bool ok = chroma11 && ((h.hmax == 1 && h.vmax == 1) || (h.hmax == 2 && h.vmax == 1) || (h.hmax == 2 && h.vmax == 2)) &&
                    h.comp[0].hs == h.hmax && h.comp[0].vs == h.vmax;
    if (!ok) JP_FAIL("only 4:4:4, 4:2:2 and 4:2:0 sampling is decoded on the GPU path");
  }
  h.mcux = (h.width + 8 * h.hmax - 1) / (8 * h.hmax);
  h.mcuy = (h.height + 8 * h.vmax - 1) / (8 * h.vmax);
  return VSB_OK;
#undef JP_FAIL
}

// ------------------------------------------------------------------------------------ host: entropy decode
struct BitReader {
  const unsigned char* d;
  size_t n, pos;
  unsigned long long buf = 0;
  int bits = 0;
  bool hit_marker = false;
  int fake_bytes = 0;  // zero bytes fed past the end of the entropy-coded data
  // true when bits that do not exist in the file were consumed: the scan is shorter than the frame needs
  bool overran() const { return bits < 8 * fake_bytes; }
  void fill() {
    while (bits <= 56) {
      unsigned b = 0;
      if (!hit_marker && pos < n) {
        b = d[pos];
        if (b == 0xFF) {
          const unsigned nx = pos + 1 < n ? d[pos + 1] : 0xD9;
          if (nx == 0) {
            pos += 2;
          } else {
            hit_marker = true;  // a marker: feed zero bits from here on (as libjpeg does)
            b = 0;
            ++fake_bytes;
          }
        } else {
          ++pos;
        }
      } else {
        ++fake_bytes;
      }
      buf |= (unsigned long long)b << (56 - bits);
      bits += 8;
    }
  }
  inline unsigned peek(int k) { return (unsigned)(buf >> (64 - k)); }
  inline void drop(int k) {
    buf <<= k;
    bits -= k;
  }
  inline int get(int k) {
    if (k == 0) return 0;
    const int v = (int)peek(k);
    drop(k);
    return v;
  }
};

inline int huff_decode(BitReader& br, const HuffTable& t) {
  if (br.bits < 16) br.fill();
  const unsigned short e = t.look[br.peek(9)];
  if (e) {
    br.drop(e >> 8);
    return e & 255;
  }
  int code = (int)br.peek(10), len = 10;
  while (len <= 16 && code > t.maxcode[len]) {
    ++len;
    code = (int)br.peek(len);
  }
  if (len > 16) return -1;
  br.drop(len);
  return t.symbols[t.valptr[len] + code - t.mincode[len]];
}

inline int extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

// coefficient layout: component c at coef + off[c], blocks in raster order of the MCU-padded grid, 64 int16 each (natural order)
int entropy_decode(const unsigned char* d, size_t n, const JpegHeader& h, short* coef, const size_t* off) {
  BitReader br{d, n, h.scan_start};
  int pred[3] = {0, 0, 0};
  long long count = 0;
  for (int my = 0; my < h.mcuy; ++my) {
    for (int mx = 0; mx < h.mcux; ++mx, ++count) {
      if (h.restart && count && count % h.restart == 0) {
        // byte-align, skip to just behind the RSTn marker, reset the DC predictors
        size_t p = br.pos;
        while (p + 1 < n && !(d[p] == 0xFF && d[p + 1] >= 0xD0 && d[p + 1] <= 0xD7)) ++p;
        if (p + 1 >= n) {
          set_error("JPEG restart marker missing");
          return VSB_ERR_INVALID;
        }
        if (br.overran()) {
          set_error("truncated JPEG data (a restart interval ends early)");
          return VSB_ERR_INVALID;
        }
        br.pos = p + 2;
        br.buf = 0;
        br.bits = 0;
        br.hit_marker = false;
        br.fake_bytes = 0;
        pred[0] = pred[1] = pred[2] = 0;
      }
      for (int c = 0; c < h.ncomp; ++c) {
        const Component& cm = h.comp[c];
        const HuffTable& dct = h.dc[cm.td];
        const HuffTable& act = h.ac[cm.ta];
        const int bw = h.mcux * cm.hs;
        for (int by = 0; by < cm.vs; ++by)
          for (int bx = 0; bx < cm.hs; ++bx) {
            short* blk = coef + off[c] + ((size_t)(my * cm.vs + by) * bw + (size_t)(mx * cm.hs + bx)) * 64;
            int s = huff_decode(br, dct);
            if (s < 0 || s > 15) {
              set_error("corrupt JPEG data (bad DC Huffman code)");
              return VSB_ERR_INVALID;
            }
            if (s) {
              if (br.bits < s) br.fill();
              pred[c] += extend(br.get(s), s);
            }
            blk[0] = (short)pred[c];
            for (int k = 1; k < 64;) {
              const int rs = huff_decode(br, act);
              if (rs < 0) {
                set_error("corrupt JPEG data (bad AC Huffman code)");
                return VSB_ERR_INVALID;
              }
              const int r = rs >> 4;
              s = rs & 15;
              if (s == 0) {
                if (r != 15) break;
                k += 16;
                continue;
              }
              k += r;
              if (k > 63) {
                set_error("corrupt JPEG data (coefficient index past 63)");
                return VSB_ERR_INVALID;
              }
              if (br.bits < s) br.fill();
              blk[kZigzag[k]] = (short)extend(br.get(s), s);
              ++k;
            }
          }
      }
    }
  }
  if (br.overran()) {  // Pillow raises "image file is truncated" here; so do we
    set_error("truncated JPEG data (the scan ends before the last MCU)");
    return VSB_ERR_INVALID;
  }
  return VSB_OK;
}

// ------------------------------------------------------------------------------------------ device kernels
// libjpeg's post-IDCT range limit (index masked to 10 bits, centred on 128)
__device__ __forceinline__ unsigned char range_limit(int v) {
  v &= 1023;
  return (unsigned char)(v < 128 ? v + 128 : (v < 512 ? 255 : (v < 896 ? 0 : v - 896)));
}

// one pass of jpeg_idct_islow over 8 values; rnd / shift = the DESCALE of the pass
__device__ __forceinline__ void idct8(const int (&x)[8], int (&o)[8], int shift) {
  typedef long long L;  // libjpeg computes in JLONG (64 bits on LP64): identical for any coefficient / table values
  L z2 = x[2], z3 = x[6];
  L z1 = (z2 + z3) * 4433;
  L tmp2 = z1 + z3 * (-15137);
  L tmp3 = z1 + z2 * 6270;
  z2 = x[0];
  z3 = x[4];
  L tmp0 = (z2 + z3) * 8192;
  L tmp1 = (z2 - z3) * 8192;
  const L tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  tmp0 = x[7];
  tmp1 = x[5];
  tmp2 = x[3];
  tmp3 = x[1];
  z1 = tmp0 + tmp3;
  z2 = tmp1 + tmp2;
  z3 = tmp0 + tmp2;
  L z4 = tmp1 + tmp3;
  const L z5 = (z3 + z4) * 9633;
  tmp0 *= 2446;
  tmp1 *= 16819;
  tmp2 *= 25172;
  tmp3 *= 12299;
  z1 *= -7373;
  z2 *= -20995;
  z3 *= -16069;
  z4 *= -3196;
  z3 += z5;
  z4 += z5;
  tmp0 += z1 + z3;
  tmp1 += z2 + z4;
  tmp2 += z2 + z3;
  tmp3 += z1 + z4;
  const L rnd = 1ll << (shift - 1);
  o[0] = (int)((tmp10 + tmp3 + rnd) >> shift);
  o[7] = (int)((tmp10 - tmp3 + rnd) >> shift);
  o[1] = (int)((tmp11 + tmp2 + rnd) >> shift);
  o[6] = (int)((tmp11 - tmp2 + rnd) >> shift);
  o[2] = (int)((tmp12 + tmp1 + rnd) >> shift);
  o[5] = (int)((tmp12 - tmp1 + rnd) >> shift);
  o[3] = (int)((tmp13 + tmp0 + rnd) >> shift);
  o[4] = (int)((tmp13 - tmp0 + rnd) >> shift);
}

struct PlaneDesc {
  long long coef_off;   // in int16 elements
  long long plane_off;  // in bytes
  int bw, bh;           // blocks
  int qt;               // index into the quantisation tables
};
struct IdctParams {
  PlaneDesc pl[3];
  int ncomp;
  int block_start[4];  // prefix sums of bw * bh
};

// 8 threads per block: thread j runs column j (pass 1), then row j (pass 2) through shared memory
__global__ void __launch_bounds__(256)
jpeg_idct_kernel(const short* __restrict__ coef, const unsigned short* __restrict__ qt, unsigned char* __restrict__ planes,
                 IdctParams p) {
  __shared__ int ws[32][8][9];
  const int lb = threadIdx.x >> 3, j = threadIdx.x & 7;
  const int b = blockIdx.x * 32 + lb;
  if (b < p.block_start[p.ncomp]) {
    const int c = b >= p.block_start[2] ? 2 : (b >= p.block_start[1] ? 1 : 0);
    const PlaneDesc pd = p.pl[c];
    const int bi = b - p.block_start[c];
    const short* blk = coef + pd.coef_off + (long long)bi * 64;
    const unsigned short* q = qt + pd.qt * 64;
    int x[8], o[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) x[r] = (int)blk[r * 8 + j] * (int)q[r * 8 + j];
    idct8(x, o, 13 - 2);
#pragma unroll
    for (int r = 0; r < 8; ++r) ws[lb][r][j] = o[r];
  }
  __syncthreads();
  if (b < p.block_start[p.ncomp]) {
    const int c = b >= p.block_start[2] ? 2 : (b >= p.block_start[1] ? 1 : 0);
    const PlaneDesc pd = p.pl[c];
    const int bi = b - p.block_start[c];
    const int by = bi / pd.bw, bx = bi - by * pd.bw;
    int x[8], o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = ws[lb][j][k];
    idct8(x, o, 13 + 2 + 3);
    unsigned char px[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) px[k] = range_limit(o[k]);
    unsigned char* dst = planes + pd.plane_off + ((long long)(by * 8 + j) * (pd.bw * 8) + bx * 8);
    *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(px);
  }
}

struct ColorParams {
  int w, h;            // image
  int ncomp, hmax, vmax;
  long long off[3];    // plane byte offsets
  int pitch[3];        // plane row pitch (bytes)
  int cw, ch;          // real (downsampled) chroma extent
};

__device__ __forceinline__ int colsum(const unsigned char* row0, const unsigned char* row1, int c) {
  return 3 * (int)row0[c] + (int)row1[c];
}

// fancy upsampling (jdsample.c h2v1 / h2v2) + YCbCr -> RGB (jdcolor.c); one thread per output pixel
__global__ void __launch_bounds__(256)
jpeg_color_kernel(const unsigned char* __restrict__ planes, unsigned char* __restrict__ rgb, ColorParams p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= p.w) return;
  const int yv = planes[p.off[0] + (long long)y * p.pitch[0] + x];
  unsigned char* dst = rgb + ((long long)y * p.w + x) * 3;
  if (p.ncomp == 1) {
    dst[0] = dst[1] = dst[2] = (unsigned char)yv;
    return;
  }
  int cc[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const unsigned char* pl = planes + p.off[1 + k];
    const int pitch = p.pitch[1 + k];
    if (p.hmax == 1) {
      cc[k] = pl[(long long)y * pitch + x];
    } else if (p.vmax == 1) {
      const unsigned char* row = pl + (long long)y * pitch;
      const int c = x >> 1, v = row[c];
      if (x & 1)
        cc[k] = c == p.cw - 1 ? v : (3 * v + row[c + 1] + 2) >> 2;
      else
        cc[k] = c == 0 ? v : (3 * v + row[c - 1] + 1) >> 2;
    } else {
      const int r = y >> 1;
      int ro = (y & 1) ? r + 1 : r - 1;  // the nearer neighbour row; past the real rows: the row itself
      ro = ro < 0 ? 0 : (ro > p.ch - 1 ? p.ch - 1 : ro);
      const unsigned char* row0 = pl + (long long)r * pitch;
      const unsigned char* row1 = pl + (long long)ro * pitch;
      const int c = x >> 1, cs = colsum(row0, row1, c);
      if (x & 1)
        cc[k] = c == p.cw - 1 ? (cs * 4 + 7) >> 4 : (cs * 3 + colsum(row0, row1, c + 1) + 7) >> 4;
      else
        cc[k] = c == 0 ? (cs * 4 + 8) >> 4 : (cs * 3 + colsum(row0, row1, c - 1) + 8) >> 4;
    }
  }
  const int cb = cc[0] - 128, cr = cc[1] - 128;
  const int r = yv + ((91881 * cr + 32768) >> 16);
  const int b = yv + ((116130 * cb + 32768) >> 16);
  const int g = yv + ((-22554 * cb + 32768 - 46802 * cr) >> 16);
  dst[0] = (unsigned char)min(max(r, 0), 255);
  dst[1] = (unsigned char)min(max(g, 0), 255);
  dst[2] = (unsigned char)min(max(b, 0), 255);
}

// One pass of Pillow's resampling: out[o, i, ch] = clip8((2^21 + sum_k in[bounds[o].first + k, i, ch] * kk[o, k]) >> 22)
// along the axis with stride `in_stride_o` (elements); `inner` = elements orthogonal to it.
__global__ void __launch_bounds__(256)
resample_pass_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out, const int* __restrict__ bounds,
                     const int* __restrict__ kk, int ksize, int out_size, long long inner, long long in_stride_o,
                     long long in_stride_i, long long out_stride_o, long long out_stride_i, int channels) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= inner * channels) return;
  const int o = blockIdx.y;
  const long long i = idx / channels;
  const int ch = (int)(idx - i * channels);
  const int first = bounds[2 * o], cnt = bounds[2 * o + 1];
  const int* k = kk + (long long)o * ksize;
  const unsigned char* src = in + first * in_stride_o + i * in_stride_i + ch;
  int acc = 1 << 21;
  for (int t = 0; t < cnt; ++t) acc += (int)src[t * in_stride_o] * k[t];
  acc >>= 22;
  out[o * out_stride_o + i * out_stride_i + ch] = (unsigned char)min(max(acc, 0), 255);
}

// -------------------------------------------------------------------------------- host: Pillow's coefficients
double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

// Resample.c precompute_coeffs + normalize_coeffs_8bpc for the whole image as the box
int precompute_coeffs(int in_size, int out_size, std::vector<int>& bounds, std::vector<int>& kk) {
  const double scale = (double)in_size / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 2.0 * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  bounds.assign((size_t)out_size * 2, 0);
  kk.assign((size_t)out_size * ksize, 0);
  std::vector<double> w((size_t)ksize);
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = 0.0 + (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      w[x] = bicubic_filter((x + xmin - center + 0.5) * ss);
      ww += w[x];
    }
    for (int x = 0; x < xmax; ++x) {
      const double v = ww != 0.0 ? w[x] / ww : w[x];
      kk[(size_t)xx * ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << 22)) : (int)(0.5 + v * (1 << 22));
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  return ksize;
}

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

struct ResizeTables {
  int in_size = 0, out_size = 0, ksize = 0;
  int* bounds = nullptr;  // device
  int* kk = nullptr;      // device
};

struct vsb_jpeg_decoder {
  size_t max_pixels = 0;
  // pinned host staging of the coefficient blocks + quantisation tables (two slots: the host decodes frame k + 1
  // while the copy of frame k is in flight)
  short* h_coef[2] = {nullptr, nullptr};
  unsigned short* h_qt[2] = {nullptr, nullptr};
  cudaEvent_t copied[2] = {nullptr, nullptr};
  size_t coef_cap = 0;  // int16 elements per slot
  int slot = 0;
  // device workspaces
  short* d_coef = nullptr;
  unsigned short* d_qt = nullptr;
  unsigned char* d_planes = nullptr;
  size_t planes_cap = 0;
  unsigned char* d_rgb = nullptr;
  unsigned char* d_tmp = nullptr;
  ResizeTables th, tv;
};

namespace {

int ensure_tables(ResizeTables& t, int in_size, int out_size, cudaStream_t s) {
  if (t.in_size == in_size && t.out_size == out_size) return VSB_OK;
  std::vector<int> bounds, kk;
  const int ksize = precompute_coeffs(in_size, out_size, bounds, kk);
  // the previous tables may still be read by kernels in flight: wait for the stream before replacing them
  VSB_CHECK_CUDA(cudaStreamSynchronize(s));
  if (t.bounds) (void)cudaFree(t.bounds);
  if (t.kk) (void)cudaFree(t.kk);
  t.bounds = t.kk = nullptr;
  t.in_size = t.out_size = 0;
  VSB_CHECK_CUDA(cudaMalloc((void**)&t.bounds, bounds.size() * sizeof(int)));
  VSB_CHECK_CUDA(cudaMalloc((void**)&t.kk, kk.size() * sizeof(int)));
  VSB_CHECK_CUDA(cudaMemcpy(t.bounds, bounds.data(), bounds.size() * sizeof(int), cudaMemcpyHostToDevice));
  VSB_CHECK_CUDA(cudaMemcpy(t.kk, kk.data(), kk.size() * sizeof(int), cudaMemcpyHostToDevice));
  t.in_size = in_size;
  t.out_size = out_size;
  t.ksize = ksize;
  return VSB_OK;
}

// [h, w, 3] -> [out_h, out_w, 3]: horizontal pass into tmp [h, out_w, 3], then vertical (each only when the size changes)
int resize_rgb(vsb_jpeg_decoder* dec, const unsigned char* in, int h, int w, unsigned char* out, int out_h, int out_w,
               cudaStream_t s) {
  const unsigned char* cur = in;
  if (w != out_w) {
    int rc = ensure_tables(dec->th, w, out_w, s);
    if (rc != VSB_OK) return rc;
    unsigned char* dst = h != out_h ? dec->d_tmp : out;
    const long long work = (long long)h * 3;
    dim3 grid((unsigned)((work + 255) / 256), (unsigned)out_w);
    resample_pass_kernel<<<grid, 256, 0, s>>>(cur, dst, dec->th.bounds, dec->th.kk, dec->th.ksize, out_w, h, 3, (long long)w * 3,
                                              3, (long long)out_w * 3, 3);
    VSB_CHECK_LAUNCH("resample_pass_kernel");
    cur = dst;
  }
  if (h != out_h) {
    int rc = ensure_tables(dec->tv, h, out_h, s);
    if (rc != VSB_OK) return rc;
    const long long work = (long long)out_w * 3;
    dim3 grid((unsigned)((work + 255) / 256), (unsigned)out_h);
    resample_pass_kernel<<<grid, 256, 0, s>>>(cur, out, dec->tv.bounds, dec->tv.kk, dec->tv.ksize, out_h, out_w, (long long)out_w * 3,
                                              3, (long long)out_w * 3, 3, 3);
    VSB_CHECK_LAUNCH("resample_pass_kernel");
  } else if (cur == in) {
    VSB_CHECK_CUDA(cudaMemcpyAsync(out, in, (size_t)h * w * 3, cudaMemcpyDeviceToDevice, s));
  }
  return VSB_OK;
}

}  // namespace

extern "C" int vsb_jpeg_info(const uint8_t* jpeg, unsigned long long bytes, int* width, int* height, int* components,
                             int* h_samp, int* v_samp) {
  VSB_CHECK_ARG(jpeg && bytes > 0, "null argument");
  JpegHeader h;
  int rc = parse_header(jpeg, (size_t)bytes, h);
  if (rc != VSB_OK) return rc;
  if (width) *width = h.width;
  if (height) *height = h.height;
  if (components) *components = h.ncomp;
  if (h_samp) *h_samp = h.hmax;
  if (v_samp) *v_samp = h.vmax;
  return VSB_OK;
}

extern "C" int vsb_jpeg_decoder_create(int max_width, int max_height, vsb_jpeg_decoder** out) {
  VSB_CHECK_ARG(out && max_width > 0 && max_height > 0 && max_width <= 16384 && max_height <= 16384, "bad decoder extent");
  *out = nullptr;
  vsb_jpeg_decoder* dec = new (std::nothrow) vsb_jpeg_decoder();
  VSB_CHECK_ARG(dec, "out of host memory");
  // MCU-padded worst case (4:4:4): three full planes, 16-pixel padding either way
  const size_t pw = ((size_t)max_width + 15) & ~(size_t)15, ph = ((size_t)max_height + 15) & ~(size_t)15;
  dec->max_pixels = (size_t)max_width * max_height;
  dec->coef_cap = 3 * pw * ph;
  dec->planes_cap = 3 * pw * ph;
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
    e = cudaMallocHost((void**)&dec->h_coef[i], dec->coef_cap * sizeof(short));
    if (e == cudaSuccess) e = cudaMallocHost((void**)&dec->h_qt[i], 4 * 64 * sizeof(unsigned short));
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&dec->copied[i], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaMalloc((void**)&dec->d_coef, dec->coef_cap * sizeof(short));
  if (e == cudaSuccess) e = cudaMalloc((void**)&dec->d_qt, 4 * 64 * sizeof(unsigned short));
  if (e == cudaSuccess) e = cudaMalloc((void**)&dec->d_planes, dec->planes_cap);
  if (e == cudaSuccess) e = cudaMalloc((void**)&dec->d_rgb, dec->max_pixels * 3);
  if (e == cudaSuccess) e = cudaMalloc((void**)&dec->d_tmp, dec->max_pixels * 3);
  if (e != cudaSuccess) {
    set_error("jpeg decoder: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    vsb_jpeg_decoder_destroy(dec);
    return VSB_ERR_CUDA;
  }
  *out = dec;
  return VSB_OK;
}

extern "C" void vsb_jpeg_decoder_destroy(vsb_jpeg_decoder* dec) {
  if (!dec) return;
  for (int i = 0; i < 2; ++i) {
    if (dec->h_coef[i]) (void)cudaFreeHost(dec->h_coef[i]);
    if (dec->h_qt[i]) (void)cudaFreeHost(dec->h_qt[i]);
    if (dec->copied[i]) (void)cudaEventDestroy(dec->copied[i]);
  }
  void* dev[] = {dec->d_coef, dec->d_qt, dec->d_planes, dec->d_rgb, dec->d_tmp, dec->th.bounds, dec->th.kk, dec->tv.bounds, dec->tv.kk};
  for (void* p : dev)
    if (p) (void)cudaFree(p);
  delete dec;
}

extern "C" int vsb_jpeg_decode_resize(vsb_jpeg_decoder* dec, const uint8_t* jpeg, unsigned long long bytes, uint8_t* out,
                                      int out_h, int out_w, void* stream) {
  VSB_CHECK_ARG(dec && jpeg && bytes > 0 && out && out_h > 0 && out_w > 0, "null argument or empty output");
  cudaStream_t s = (cudaStream_t)stream;
  JpegHeader h;
  int rc = parse_header(jpeg, (size_t)bytes, h);
  if (rc != VSB_OK) return rc;
  VSB_CHECK_ARG((size_t)h.width * h.height <= dec->max_pixels, "JPEG is %d x %d, the decoder was created for %zu pixels",
                h.width, h.height, dec->max_pixels);
  VSB_CHECK_ARG((size_t)h.height * out_w <= dec->max_pixels, "output wider than the decoder's workspace");
  // layout of the coefficient blocks and of the sample planes
  size_t coef_off[3] = {0, 0, 0}, plane_off[3] = {0, 0, 0}, ncoef = 0, nplane = 0;
  IdctParams ip;
  memset(&ip, 0, sizeof(ip));
  ip.ncomp = h.ncomp;
  for (int c = 0; c < h.ncomp; ++c) {
    const int bw = h.mcux * h.comp[c].hs, bh = h.mcuy * h.comp[c].vs;
    coef_off[c] = ncoef;
    plane_off[c] = nplane;
    ncoef += (size_t)bw * bh * 64;
    nplane += align256((size_t)bw * bh * 64);
    ip.pl[c].coef_off = (long long)coef_off[c];
    ip.pl[c].plane_off = (long long)plane_off[c];
    ip.pl[c].bw = bw;
    ip.pl[c].bh = bh;
    ip.pl[c].qt = h.comp[c].tq;
    ip.block_start[c + 1] = ip.block_start[c] + bw * bh;
  }
  for (int c = h.ncomp; c < 3; ++c) ip.block_start[c + 1] = ip.block_start[c];
  VSB_CHECK_ARG(ncoef <= dec->coef_cap && nplane <= dec->planes_cap, "JPEG larger than the decoder's workspace");
  // host: entropy decode into the next pinned slot (wait until its previous copy has left)
  const int slot = dec->slot;
  dec->slot ^= 1;
  VSB_CHECK_CUDA(cudaEventSynchronize(dec->copied[slot]));
  memset(dec->h_coef[slot], 0, ncoef * sizeof(short));
  rc = entropy_decode(jpeg, (size_t)bytes, h, dec->h_coef[slot], coef_off);
  if (rc != VSB_OK) return rc;
  memcpy(dec->h_qt[slot], h.qt, sizeof(h.qt));
  // device: everything that touches pixels
  VSB_CHECK_CUDA(cudaMemcpyAsync(dec->d_coef, dec->h_coef[slot], ncoef * sizeof(short), cudaMemcpyHostToDevice, s));
  VSB_CHECK_CUDA(cudaMemcpyAsync(dec->d_qt, dec->h_qt[slot], sizeof(h.qt), cudaMemcpyHostToDevice, s));
  VSB_CHECK_CUDA(cudaEventRecord(dec->copied[slot], s));
  const int nblocks = ip.block_start[h.ncomp];
  jpeg_idct_kernel<<<(nblocks + 31) / 32, 256, 0, s>>>(dec->d_coef, dec->d_qt, dec->d_planes, ip);
  VSB_CHECK_LAUNCH("jpeg_idct_kernel");
  ColorParams cp;
  memset(&cp, 0, sizeof(cp));
  cp.w = h.width;
  cp.h = h.height;
  cp.ncomp = h.ncomp;
  cp.hmax = h.hmax;
  cp.vmax = h.vmax;
  for (int c = 0; c < h.ncomp; ++c) {
    cp.off[c] = (long long)plane_off[c];
    cp.pitch[c] = ip.pl[c].bw * 8;
  }
  cp.cw = (h.width + h.hmax - 1) / h.hmax;
  cp.ch = (h.height + h.vmax - 1) / h.vmax;
  dim3 grid((unsigned)((h.width + 255) / 256), (unsigned)h.height);
  jpeg_color_kernel<<<grid, 256, 0, s>>>(dec->d_planes, dec->d_rgb, cp);
  VSB_CHECK_LAUNCH("jpeg_color_kernel");
  return resize_rgb(dec, dec->d_rgb, h.height, h.width, out, out_h, out_w, s);
}

extern "C" int vsb_resize_bicubic_u8(vsb_jpeg_decoder* dec, const uint8_t* in, int h, int w, uint8_t* out, int out_h, int out_w,
                                     void* stream) {
  VSB_CHECK_ARG(dec && in && out && h > 0 && w > 0 && out_h > 0 && out_w > 0, "null argument or empty image");
  VSB_CHECK_ARG((size_t)h * out_w <= dec->max_pixels, "image larger than the decoder's workspace");
  return resize_rgb(dec, in, h, w, out, out_h, out_w, (cudaStream_t)stream);
}

// ====================================================================================================================
// Batched, fully-on-device decode: the Huffman segment too.
//
// One frame's entropy-coded segment is sequential, but a feature-extraction batch holds hundreds to thousands of
// independent frames (5 events x 32 frames x videos): ONE WARP PER FRAME walks its bit stream (lane 0 decodes from
// shared-memory copies of the frame's Huffman tables, the whole warp fetches the tables), a few hundred warps in
// flight hide each other's latency.  The host only parses headers (tables, geometry) and copies the file bytes to
// pinned memory; no pixel and no coefficient crosses PCIe.  The pixel stages are the kernels above, launched once per
// batch over (frame, block / pixel) grids.  Same arithmetic, same bits as Pillow.
namespace {

struct HuffDev {  // a HuffTable as the device decoder reads it (4-byte aligned)
  unsigned short look[512];
  int maxcode[18];
  int valptr[17];
  int mincode[17];
  unsigned char symbols[256];
};

struct FrameDesc {
  // entropy decode
  long long data_off;    // first byte of the entropy-coded segment in the batch's byte buffer
  long long data_len;    // bytes from there to the end of the file
  int tab[3][2];         // per component: index of its DC / AC table among the frame's 4 uploaded tables
  int ncomp, mcux, mcuy, restart;
  int hs[3], vs[3];
  long long coef_off[3];  // int16 elements
  // pixel stages
  IdctParams ip;
  ColorParams cp;
  int qt_base;            // first of the frame's 4 quantisation tables (x 64 entries)
  long long rgb_off, tmp_off;
  unsigned char* out;
  const int* bounds_h;
  const int* kk_h;
  const int* bounds_v;
  const int* kk_v;
  int ksize_h, ksize_v;
  int valid;              // 0: refused by the host parser (nothing is launched for it)
};

struct DevBits {
  const unsigned char* d;
  long long n, pos;
  unsigned long long buf;
  int bits;
  bool hit_marker;
  int fake_bytes;
  __device__ __forceinline__ void fill() {
    // fast path: eight stream bytes in one go when none of them is 0xFF (stuffing / markers take the byte loop).
    // `d` is 8-byte aligned and the buffer has 16 bytes of slack behind the last segment.
    if (!hit_marker && pos + 8 <= n && bits <= 56) {
      const unsigned long long* w = reinterpret_cast<const unsigned long long*>(d) + (pos >> 3);
      const int sh = (int)(pos & 7) * 8;
      unsigned long long u = w[0] >> sh;
      if (sh) u |= w[1] << (64 - sh);
      const unsigned long long v = ~u;
      if (((v - 0x0101010101010101ull) & ~v & 0x8080808080808080ull) == 0) {
        const unsigned lo = (unsigned)u, hi = (unsigned)(u >> 32);
        const unsigned long long be = ((unsigned long long)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
        const int k = (64 - bits) >> 3;  // 1..8 whole bytes fit
        buf |= (be >> (64 - 8 * k)) << (64 - bits - 8 * k);
        bits += 8 * k;
        pos += k;
        return;
      }
    }
    while (bits <= 56) {
      unsigned b = 0;
      if (!hit_marker && pos < n) {
        b = d[pos];
        if (b == 0xFF) {
          const unsigned nx = pos + 1 < n ? d[pos + 1] : 0xD9;
          if (nx == 0) {
            pos += 2;
          } else {
            hit_marker = true;
            b = 0;
            ++fake_bytes;
          }
        } else {
          ++pos;
        }
      } else {
        ++fake_bytes;
      }
      buf |= (unsigned long long)b << (56 - bits);
      bits += 8;
    }
  }
  __device__ __forceinline__ unsigned peek(int k) const { return (unsigned)(buf >> (64 - k)); }
  __device__ __forceinline__ void drop(int k) {
    buf <<= k;
    bits -= k;
  }
  __device__ __forceinline__ bool overran() const { return bits < 8 * fake_bytes; }
};

__device__ __forceinline__ int dev_huff(DevBits& br, const HuffDev& t) {
  if (br.bits < 16) br.fill();
  const unsigned short e = t.look[br.peek(9)];
  if (e) {
    br.drop(e >> 8);
    return e & 255;
  }
  int code = (int)br.peek(10), len = 10;
  while (len <= 16 && code > t.maxcode[len]) {
    ++len;
    code = (int)br.peek(len);
  }
  if (len > 16) return -1;
  br.drop(len);
  return t.symbols[t.valptr[len] + code - t.mincode[len]];
}

__device__ __forceinline__ int dev_receive(DevBits& br, int s) {
  if (br.bits < s) br.fill();
  const int v = (int)br.peek(s);
  br.drop(s);
  return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
}

__constant__ unsigned char kZigzagDev[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                              41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                              30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

constexpr int kHuffWarps = 4;

// status: 0 = decoded, 1 = corrupt / truncated data
__global__ void __launch_bounds__(kHuffWarps * 32)
jpeg_huffman_kernel(const unsigned char* __restrict__ bytes, const HuffDev* __restrict__ tables, const FrameDesc* __restrict__ frames,
                    int nframes, short* __restrict__ coef, int* __restrict__ status) {
  __shared__ HuffDev tabs[kHuffWarps][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = blockIdx.x * kHuffWarps + warp;
  if (f >= nframes) return;
  const FrameDesc& fd = frames[f];
  if (!fd.valid) return;
  {
    const unsigned int* src = reinterpret_cast<const unsigned int*>(tables + (long long)f * 4);
    unsigned int* dst = reinterpret_cast<unsigned int*>(&tabs[warp][0]);
    for (int i = lane; i < (int)(4 * sizeof(HuffDev) / 4); i += 32) dst[i] = src[i];
  }
  __syncwarp();
  if (lane != 0) return;
  DevBits br;
  br.d = bytes + fd.data_off;
  br.n = fd.data_len;
  br.pos = 0;
  br.buf = 0;
  br.bits = 0;
  br.hit_marker = false;
  br.fake_bytes = 0;
  int pred[3] = {0, 0, 0};
  int bad = 0;
  long long count = 0;
  for (int my = 0; my < fd.mcuy && !bad; ++my) {
    for (int mx = 0; mx < fd.mcux && !bad; ++mx, ++count) {
      if (fd.restart && count && count % fd.restart == 0) {
        long long p = br.pos;
        while (p + 1 < br.n && !(br.d[p] == 0xFF && br.d[p + 1] >= 0xD0 && br.d[p + 1] <= 0xD7)) ++p;
        if (p + 1 >= br.n || br.overran()) {
          bad = 1;
          break;
        }
        br.pos = p + 2;
        br.buf = 0;
        br.bits = 0;
        br.hit_marker = false;
        br.fake_bytes = 0;
        pred[0] = pred[1] = pred[2] = 0;
      }
      for (int c = 0; c < fd.ncomp && !bad; ++c) {
        const HuffDev& dct = tabs[warp][fd.tab[c][0]];
        const HuffDev& act = tabs[warp][fd.tab[c][1]];
        const int bw = fd.mcux * fd.hs[c];
        for (int by = 0; by < fd.vs[c] && !bad; ++by)
          for (int bx = 0; bx < fd.hs[c] && !bad; ++bx) {
            short* blk = coef + fd.coef_off[c] + ((long long)(my * fd.vs[c] + by) * bw + (mx * fd.hs[c] + bx)) * 64;
            int s = dev_huff(br, dct);
            if (s < 0 || s > 15) {
              bad = 1;
              break;
            }
            if (s) pred[c] += dev_receive(br, s);
            blk[0] = (short)pred[c];
            for (int k = 1; k < 64;) {
              const int rs = dev_huff(br, act);
              if (rs < 0) {
                bad = 1;
                break;
              }
              const int r = rs >> 4;
              s = rs & 15;
              if (s == 0) {
                if (r != 15) break;
                k += 16;
                continue;
              }
              k += r;
              if (k > 63) {
                bad = 1;
                break;
              }
              blk[kZigzagDev[k]] = (short)dev_receive(br, s);
              ++k;
            }
          }
      }
    }
  }
  if (!bad && br.overran()) bad = 1;
  status[f] = bad;
}

// the pixel stages over a batch: blockIdx.y (idct) / blockIdx.z (colour, resize) = frame
__global__ void __launch_bounds__(256)
jpeg_idct_batch_kernel(const short* __restrict__ coef, const unsigned short* __restrict__ qt, unsigned char* __restrict__ planes,
                       const FrameDesc* __restrict__ frames, const int* __restrict__ status) {
  __shared__ int ws[32][8][9];
  const FrameDesc& fd = frames[blockIdx.y];
  const bool live = fd.valid && status[blockIdx.y] == 0;
  const IdctParams& p = fd.ip;
  const int lb = threadIdx.x >> 3, j = threadIdx.x & 7;
  const int b = blockIdx.x * 32 + lb;
  const bool on = live && b < p.block_start[p.ncomp];
  int c = 0, bi = 0;
  if (on) {
    c = b >= p.block_start[2] ? 2 : (b >= p.block_start[1] ? 1 : 0);
    bi = b - p.block_start[c];
    const short* blk = coef + p.pl[c].coef_off + (long long)bi * 64;
    const unsigned short* q = qt + (long long)(fd.qt_base + p.pl[c].qt) * 64;
    int x[8], o[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) x[r] = (int)blk[r * 8 + j] * (int)q[r * 8 + j];
    idct8(x, o, 13 - 2);
#pragma unroll
    for (int r = 0; r < 8; ++r) ws[lb][r][j] = o[r];
  }
  __syncthreads();
  if (on) {
    const PlaneDesc pd = p.pl[c];
    const int by = bi / pd.bw, bx = bi - by * pd.bw;
    int x[8], o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = ws[lb][j][k];
    idct8(x, o, 13 + 2 + 3);
    unsigned char px[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) px[k] = range_limit(o[k]);
    unsigned char* dst = planes + pd.plane_off + ((long long)(by * 8 + j) * (pd.bw * 8) + bx * 8);
    *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(px);
  }
}

__device__ __forceinline__ void color_pixel(const unsigned char* planes, unsigned char* dst, const ColorParams& p, int x, int y) {
  const int yv = planes[p.off[0] + (long long)y * p.pitch[0] + x];
  if (p.ncomp == 1) {
    dst[0] = dst[1] = dst[2] = (unsigned char)yv;
    return;
  }
  int cc[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const unsigned char* pl = planes + p.off[1 + k];
    const int pitch = p.pitch[1 + k];
    if (p.hmax == 1) {
      cc[k] = pl[(long long)y * pitch + x];
    } else if (p.vmax == 1) {
      const unsigned char* row = pl + (long long)y * pitch;
      const int c = x >> 1, v = row[c];
      if (x & 1)
        cc[k] = c == p.cw - 1 ? v : (3 * v + row[c + 1] + 2) >> 2;
      else
        cc[k] = c == 0 ? v : (3 * v + row[c - 1] + 1) >> 2;
    } else {
      const int r = y >> 1;
      int ro = (y & 1) ? r + 1 : r - 1;
      ro = ro < 0 ? 0 : (ro > p.ch - 1 ? p.ch - 1 : ro);
      const unsigned char* row0 = pl + (long long)r * pitch;
      const unsigned char* row1 = pl + (long long)ro * pitch;
      const int c = x >> 1, cs = colsum(row0, row1, c);
      if (x & 1)
        cc[k] = c == p.cw - 1 ? (cs * 4 + 7) >> 4 : (cs * 3 + colsum(row0, row1, c + 1) + 7) >> 4;
      else
        cc[k] = c == 0 ? (cs * 4 + 8) >> 4 : (cs * 3 + colsum(row0, row1, c - 1) + 8) >> 4;
    }
  }
  const int cb = cc[0] - 128, cr = cc[1] - 128;
  const int r = yv + ((91881 * cr + 32768) >> 16);
  const int b = yv + ((116130 * cb + 32768) >> 16);
  const int g = yv + ((-22554 * cb + 32768 - 46802 * cr) >> 16);
  dst[0] = (unsigned char)min(max(r, 0), 255);
  dst[1] = (unsigned char)min(max(g, 0), 255);
  dst[2] = (unsigned char)min(max(b, 0), 255);
}

__global__ void __launch_bounds__(256)
jpeg_color_batch_kernel(const unsigned char* __restrict__ planes, unsigned char* __restrict__ rgb,
                        const FrameDesc* __restrict__ frames, const int* __restrict__ status) {
  const FrameDesc& fd = frames[blockIdx.z];
  if (!fd.valid || status[blockIdx.z]) return;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= fd.cp.w || y >= fd.cp.h) return;
  color_pixel(planes, rgb + fd.rgb_off + ((long long)y * fd.cp.w + x) * 3, fd.cp, x, y);
}

// pass 0: horizontal (rgb [h, w, 3] -> tmp [h, out_w, 3], or straight to out when h == out_h); pass 1: vertical
__global__ void __launch_bounds__(256)
resample_batch_kernel(const unsigned char* __restrict__ rgb, unsigned char* __restrict__ tmp, const FrameDesc* __restrict__ frames,
                      const int* __restrict__ status, int pass, int out_h, int out_w) {
  const FrameDesc& fd = frames[blockIdx.z];
  if (!fd.valid || status[blockIdx.z]) return;
  const int h = fd.cp.h, w = fd.cp.w;
  const int o = blockIdx.y;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const unsigned char* in;
  unsigned char* out;
  const int* bounds;
  const int* kk;
  int ksize;
  long long inner, in_so, in_si, out_so, out_si;
  if (pass == 0) {
    if (w == out_w) return;
    in = rgb + fd.rgb_off;
    out = h != out_h ? tmp + fd.tmp_off : fd.out;
    bounds = fd.bounds_h, kk = fd.kk_h, ksize = fd.ksize_h;
    inner = h, in_so = 3, in_si = (long long)w * 3, out_so = 3, out_si = (long long)out_w * 3;
    if (o >= out_w) return;
  } else {
    if (h == out_h) {
      if (w == out_w) {  // neither pass resamples: copy
        if (o < out_h && idx < (long long)out_w * 3) fd.out[(long long)o * out_w * 3 + idx] = rgb[fd.rgb_off + (long long)o * w * 3 + idx];
      }
      return;
    }
    in = w != out_w ? tmp + fd.tmp_off : rgb + fd.rgb_off;
    out = fd.out;
    bounds = fd.bounds_v, kk = fd.kk_v, ksize = fd.ksize_v;
    inner = out_w, in_so = (long long)out_w * 3, in_si = 3, out_so = (long long)out_w * 3, out_si = 3;
    if (o >= out_h) return;
  }
  if (idx >= inner * 3) return;
  const long long i = idx / 3;
  const int ch = (int)(idx - i * 3);
  const int first = bounds[2 * o], cnt = bounds[2 * o + 1];
  const int* k = kk + (long long)o * ksize;
  const unsigned char* src = in + first * in_so + i * in_si + ch;
  int acc = 1 << 21;
  for (int t = 0; t < cnt; ++t) acc += (int)src[t * in_so] * k[t];
  acc >>= 22;
  out[o * out_so + i * out_si + ch] = (unsigned char)min(max(acc, 0), 255);
}

template <typename T>
int grow(T** ptr, size_t* cap, size_t need, bool pinned) {
  if (need <= *cap) return VSB_OK;
  if (*ptr) {
    if (pinned) (void)cudaFreeHost(*ptr);
    else (void)cudaFree(*ptr);
    *ptr = nullptr;
    *cap = 0;
  }
  need += need / 4;  // headroom: batches of one video differ a little in size
  cudaError_t e = pinned ? cudaMallocHost((void**)ptr, need * sizeof(T)) : cudaMalloc((void**)ptr, need * sizeof(T));
  if (e != cudaSuccess) {
    set_error("jpeg batch: allocating %zu bytes failed: %s", need * sizeof(T), cudaGetErrorString(e));
    (void)cudaGetLastError();
    return VSB_ERR_CUDA;
  }
  *cap = need;
  return VSB_OK;
}

}  // namespace

struct vsb_jpeg_batch {
  // pinned host staging
  unsigned char* h_bytes = nullptr; size_t h_bytes_cap = 0;
  FrameDesc* h_frames = nullptr;    size_t h_frames_cap = 0;
  HuffDev* h_tabs = nullptr;        size_t h_tabs_cap = 0;
  unsigned short* h_qt = nullptr;   size_t h_qt_cap = 0;
  int* h_status = nullptr;          size_t h_status_cap = 0;
  // device
  unsigned char* d_bytes = nullptr; size_t d_bytes_cap = 0;
  FrameDesc* d_frames = nullptr;    size_t d_frames_cap = 0;
  HuffDev* d_tabs = nullptr;        size_t d_tabs_cap = 0;
  unsigned short* d_qt = nullptr;   size_t d_qt_cap = 0;
  int* d_status = nullptr;          size_t d_status_cap = 0;
  short* d_coef = nullptr;          size_t d_coef_cap = 0;
  unsigned char* d_planes = nullptr; size_t d_planes_cap = 0;
  unsigned char* d_rgb = nullptr;   size_t d_rgb_cap = 0;
  unsigned char* d_tmp = nullptr;   size_t d_tmp_cap = 0;
  std::vector<ResizeTables> tables;  // resampling tables by (input size, output size)
};

namespace {
const ResizeTables* batch_tables(vsb_jpeg_batch* b, int in_size, int out_size) {
  for (const ResizeTables& t : b->tables)
    if (t.in_size == in_size && t.out_size == out_size) return &t;
  std::vector<int> bounds, kk;
  ResizeTables t;
  t.ksize = precompute_coeffs(in_size, out_size, bounds, kk);
  if (cudaMalloc((void**)&t.bounds, bounds.size() * sizeof(int)) != cudaSuccess ||
      cudaMalloc((void**)&t.kk, kk.size() * sizeof(int)) != cudaSuccess ||
      cudaMemcpy(t.bounds, bounds.data(), bounds.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(t.kk, kk.data(), kk.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("jpeg batch: resampling tables: %s", cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  t.in_size = in_size;
  t.out_size = out_size;
  b->tables.push_back(t);
  return &b->tables.back();
}

void copy_table(HuffDev& dst, const HuffTable& src) {
  memcpy(dst.look, src.look, sizeof(dst.look));
  memcpy(dst.maxcode, src.maxcode, sizeof(dst.maxcode));
  memcpy(dst.valptr, src.valptr, sizeof(dst.valptr));
  memcpy(dst.mincode, src.mincode, sizeof(dst.mincode));
  memcpy(dst.symbols, src.symbols, sizeof(dst.symbols));
}
}  // namespace

extern "C" int vsb_jpeg_batch_create(vsb_jpeg_batch** out) {
  VSB_CHECK_ARG(out, "null argument");
  *out = new (std::nothrow) vsb_jpeg_batch();
  VSB_CHECK_ARG(*out, "out of host memory");
  return VSB_OK;
}

extern "C" void vsb_jpeg_batch_destroy(vsb_jpeg_batch* b) {
  if (!b) return;
  void* pinned[] = {b->h_bytes, b->h_frames, b->h_tabs, b->h_qt, b->h_status};
  for (void* p : pinned)
    if (p) (void)cudaFreeHost(p);
  void* dev[] = {b->d_bytes, b->d_frames, b->d_tabs, b->d_qt, b->d_status, b->d_coef, b->d_planes, b->d_rgb, b->d_tmp};
  for (void* p : dev)
    if (p) (void)cudaFree(p);
  for (ResizeTables& t : b->tables) {
    if (t.bounds) (void)cudaFree(t.bounds);
    if (t.kk) (void)cudaFree(t.kk);
  }
  delete b;
}

extern "C" int vsb_jpeg_batch_decode_resize(vsb_jpeg_batch* b, const uint8_t* const* jpegs, const unsigned long long* bytes,
                                            int n, uint8_t* const* outs, int out_h, int out_w, int* status, void* stream) {
  VSB_CHECK_ARG(b && jpegs && bytes && outs && status && n > 0 && out_h > 0 && out_w > 0, "null argument or empty batch");
  cudaStream_t s = (cudaStream_t)stream;
  // ---- host: headers -> per-frame descriptors, workspace layout
  std::vector<JpegHeader> hdr((size_t)n);
  size_t nbytes = 0, ncoef = 0, nplanes = 0, nrgb = 0, ntmp = 0;
  int rc = grow(&b->h_frames, &b->h_frames_cap, (size_t)n, true);
  if (rc == VSB_OK) rc = grow(&b->h_tabs, &b->h_tabs_cap, (size_t)n * 4, true);
  if (rc == VSB_OK) rc = grow(&b->h_qt, &b->h_qt_cap, (size_t)n * 4 * 64, true);
  if (rc == VSB_OK) rc = grow(&b->h_status, &b->h_status_cap, (size_t)n, true);
  if (rc != VSB_OK) return rc;
  memset(b->h_frames, 0, (size_t)n * sizeof(FrameDesc));
  memset(b->h_tabs, 0, (size_t)n * 4 * sizeof(HuffDev));
  int max_blocks = 0, max_w = 0, max_h = 0, n_valid = 0;
  for (int f = 0; f < n; ++f) {
    FrameDesc& fd = b->h_frames[f];
    status[f] = VSB_ERR_INVALID;
    if (!jpegs[f] || !outs[f] || parse_header(jpegs[f], (size_t)bytes[f], hdr[f]) != VSB_OK) continue;  // refused: the caller's fallback
    const JpegHeader& h = hdr[f];
    fd.valid = 1;
    ++n_valid;
    nbytes = (nbytes + 7) & ~(size_t)7;  // the device bit reader fetches aligned 8-byte words
    const size_t frame_first = nbytes;
    fd.data_off = (long long)nbytes;
    fd.data_len = (long long)(bytes[f] - h.scan_start);
    nbytes += (size_t)fd.data_len;
    fd.ncomp = h.ncomp, fd.mcux = h.mcux, fd.mcuy = h.mcuy, fd.restart = h.restart;
    // the frame's (up to) four Huffman tables: slots 0 / 1 = DC / AC of component 0, 2 / 3 = of the chroma components
    for (int c = 0; c < h.ncomp; ++c) {
      fd.hs[c] = h.comp[c].hs, fd.vs[c] = h.comp[c].vs;
      const int slot = c == 0 ? 0 : 2;
      if (c == 2 && (h.comp[2].td != h.comp[1].td || h.comp[2].ta != h.comp[1].ta)) {
        fd.valid = 0;  // three distinct table pairs: outside the batch path (never produced by the usual encoders)
        break;
      }
      fd.tab[c][0] = slot, fd.tab[c][1] = slot + 1;
      copy_table(b->h_tabs[(size_t)f * 4 + slot], h.dc[h.comp[c].td]);
      copy_table(b->h_tabs[(size_t)f * 4 + slot + 1], h.ac[h.comp[c].ta]);
    }
    if (!fd.valid) {
      --n_valid;
      nbytes = frame_first;
      continue;
    }
    memcpy(b->h_qt + (size_t)f * 256, h.qt, sizeof(h.qt));
    fd.qt_base = f * 4;
    fd.ip.ncomp = h.ncomp;
    for (int c = 0; c < h.ncomp; ++c) {
      const int bw = h.mcux * h.comp[c].hs, bh = h.mcuy * h.comp[c].vs;
      fd.coef_off[c] = (long long)ncoef;
      fd.ip.pl[c].coef_off = (long long)ncoef;
      fd.ip.pl[c].plane_off = (long long)nplanes;
      fd.ip.pl[c].bw = bw, fd.ip.pl[c].bh = bh, fd.ip.pl[c].qt = h.comp[c].tq;
      fd.ip.block_start[c + 1] = fd.ip.block_start[c] + bw * bh;
      fd.cp.off[c] = (long long)nplanes;
      fd.cp.pitch[c] = bw * 8;
      ncoef += (size_t)bw * bh * 64;
      nplanes += align256((size_t)bw * bh * 64);
    }
    for (int c = h.ncomp; c < 3; ++c) fd.ip.block_start[c + 1] = fd.ip.block_start[c];
    if (fd.ip.block_start[h.ncomp] > max_blocks) max_blocks = fd.ip.block_start[h.ncomp];
    fd.cp.w = h.width, fd.cp.h = h.height, fd.cp.ncomp = h.ncomp, fd.cp.hmax = h.hmax, fd.cp.vmax = h.vmax;
    fd.cp.cw = (h.width + h.hmax - 1) / h.hmax;
    fd.cp.ch = (h.height + h.vmax - 1) / h.vmax;
    if (h.width > max_w) max_w = h.width;
    if (h.height > max_h) max_h = h.height;
    fd.rgb_off = (long long)nrgb;
    nrgb += align256((size_t)h.width * h.height * 3);
    fd.tmp_off = (long long)ntmp;
    ntmp += align256((size_t)h.height * out_w * 3);
    fd.out = outs[f];
    if (h.width != out_w) {
      const ResizeTables* t = batch_tables(b, h.width, out_w);
      if (!t) return VSB_ERR_CUDA;
      fd.bounds_h = t->bounds, fd.kk_h = t->kk, fd.ksize_h = t->ksize;
    }
    if (h.height != out_h) {
      const ResizeTables* t = batch_tables(b, h.height, out_h);
      if (!t) return VSB_ERR_CUDA;
      fd.bounds_v = t->bounds, fd.kk_v = t->kk, fd.ksize_v = t->ksize;
    }
  }
  if (n_valid == 0) return VSB_OK;  // every frame was refused: status[] says so
  // ---- workspaces (grow-only; a reallocation waits for the device)
  const bool realloc = nbytes + 64 > b->d_bytes_cap || (size_t)n > b->d_frames_cap || ncoef > b->d_coef_cap || nplanes > b->d_planes_cap ||
                       nrgb > b->d_rgb_cap || ntmp > b->d_tmp_cap || nbytes + 64 > b->h_bytes_cap;
  if (realloc) VSB_CHECK_CUDA(cudaDeviceSynchronize());
  nbytes += 64;  // slack behind the last segment (word-wise fetches)
  rc = grow(&b->h_bytes, &b->h_bytes_cap, nbytes, true);
  if (rc == VSB_OK) rc = grow(&b->d_bytes, &b->d_bytes_cap, nbytes, false);
  if (rc == VSB_OK) rc = grow(&b->d_frames, &b->d_frames_cap, (size_t)n, false);
  if (rc == VSB_OK) rc = grow(&b->d_tabs, &b->d_tabs_cap, (size_t)n * 4, false);
  if (rc == VSB_OK) rc = grow(&b->d_qt, &b->d_qt_cap, (size_t)n * 256, false);
  if (rc == VSB_OK) rc = grow(&b->d_status, &b->d_status_cap, (size_t)n, false);
  if (rc == VSB_OK) rc = grow(&b->d_coef, &b->d_coef_cap, ncoef, false);
  if (rc == VSB_OK) rc = grow(&b->d_planes, &b->d_planes_cap, nplanes, false);
  if (rc == VSB_OK) rc = grow(&b->d_rgb, &b->d_rgb_cap, nrgb, false);
  if (rc == VSB_OK) rc = grow(&b->d_tmp, &b->d_tmp_cap, ntmp ? ntmp : 1, false);
  if (rc != VSB_OK) return rc;
  for (int f = 0; f < n; ++f)
    if (b->h_frames[f].valid)
      memcpy(b->h_bytes + b->h_frames[f].data_off, jpegs[f] + hdr[f].scan_start, (size_t)b->h_frames[f].data_len);
  // ---- device
  VSB_CHECK_CUDA(cudaMemcpyAsync(b->d_bytes, b->h_bytes, nbytes, cudaMemcpyHostToDevice, s));
  VSB_CHECK_CUDA(cudaMemcpyAsync(b->d_frames, b->h_frames, (size_t)n * sizeof(FrameDesc), cudaMemcpyHostToDevice, s));
  VSB_CHECK_CUDA(cudaMemcpyAsync(b->d_tabs, b->h_tabs, (size_t)n * 4 * sizeof(HuffDev), cudaMemcpyHostToDevice, s));
  VSB_CHECK_CUDA(cudaMemcpyAsync(b->d_qt, b->h_qt, (size_t)n * 256 * sizeof(unsigned short), cudaMemcpyHostToDevice, s));
  VSB_CHECK_CUDA(cudaMemsetAsync(b->d_coef, 0, ncoef * sizeof(short), s));
  VSB_CHECK_CUDA(cudaMemsetAsync(b->d_status, 0, (size_t)n * sizeof(int), s));
  jpeg_huffman_kernel<<<(n + kHuffWarps - 1) / kHuffWarps, kHuffWarps * 32, 0, s>>>(b->d_bytes, b->d_tabs, b->d_frames, n, b->d_coef,
                                                                                    b->d_status);
  VSB_CHECK_LAUNCH("jpeg_huffman_kernel");
  jpeg_idct_batch_kernel<<<dim3((unsigned)((max_blocks + 31) / 32), (unsigned)n), 256, 0, s>>>(b->d_coef, b->d_qt, b->d_planes, b->d_frames,
                                                                                              b->d_status);
  VSB_CHECK_LAUNCH("jpeg_idct_batch_kernel");
  jpeg_color_batch_kernel<<<dim3((unsigned)((max_w + 255) / 256), (unsigned)max_h, (unsigned)n), 256, 0, s>>>(b->d_planes, b->d_rgb,
                                                                                                             b->d_frames, b->d_status);
  VSB_CHECK_LAUNCH("jpeg_color_batch_kernel");
  resample_batch_kernel<<<dim3((unsigned)(((long long)max_h * 3 + 255) / 256), (unsigned)out_w, (unsigned)n), 256, 0, s>>>(
      b->d_rgb, b->d_tmp, b->d_frames, b->d_status, 0, out_h, out_w);
  VSB_CHECK_LAUNCH("resample_batch_kernel");
  resample_batch_kernel<<<dim3((unsigned)(((long long)out_w * 3 + 255) / 256), (unsigned)out_h, (unsigned)n), 256, 0, s>>>(
      b->d_rgb, b->d_tmp, b->d_frames, b->d_status, 1, out_h, out_w);
  VSB_CHECK_LAUNCH("resample_batch_kernel");
  VSB_CHECK_CUDA(cudaMemcpyAsync(b->h_status, b->d_status, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, s));
  VSB_CHECK_CUDA(cudaStreamSynchronize(s));  // the per-frame verdicts are part of the result
  for (int f = 0; f < n; ++f)
    if (b->h_frames[f].valid) status[f] = b->h_status[f] ? VSB_ERR_INVALID : VSB_OK;
  return VSB_OK;
}
