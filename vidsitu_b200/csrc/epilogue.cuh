// Epilogue building blocks shared by the tensor-core conv kernels: TMEM accumulator rows ->
// frozen-BatchNorm scale/bias (+ residual) (+ ReLU) in fp32 -> one bf16 rounding -> swizzled
// shared-memory slab that a TMA store ships out.  Also a division-free tile cursor.
#pragma once
#include <cuda_fp16.h>

#include "ptx.cuh"

namespace vsb {

// 32 lanes x 32 consecutive fp32 columns in one instruction.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
      " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// explicit shared-space accesses (32-bit addresses: no generic-address resolution, fewer registers)
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ float4 lds128f(uint32_t saddr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// One thread's 16 accumulator columns -> its 32 bytes of the staging row (two swizzled 16-byte
// chunks).  buf_s / sb_s are SHARED-space addresses: the slab, and the (scale, bias) float2 pairs of
// these 16 columns (constant for the kernel: plain asm so the loads can be hoisted); relu_floor = 0 or -inf.
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// kF16: the 16-bit output is IEEE half instead of bf16 (attention scores: 3 more mantissa bits before the softmax)
template <bool kRes, bool kF16 = false>
__device__ __forceinline__ void epi_convert16(const uint32_t* v, uint32_t sb_s, uint32_t buf_s, uint32_t off0,
                                              uint32_t swz_mask, float relu_floor) {
  const uint32_t a0 = buf_s + (off0 ^ (((off0 >> 7) & swz_mask) << 4));
  const uint32_t off1 = off0 + 16;
  const uint32_t a1 = buf_s + (off1 ^ (((off1 >> 7) & swz_mask) << 4));
  uint32_t rr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (kRes) {
    const uint4 r0 = lds128(a0), r1 = lds128(a1);
    rr[0] = r0.x; rr[1] = r0.y; rr[2] = r0.z; rr[3] = r0.w;
    rr[4] = r1.x; rr[5] = r1.y; rr[6] = r1.z; rr[7] = r1.w;
  }
  uint32_t o[8];
#pragma unroll
  for (int qq = 0; qq < 8; ++qq) {
    const float4 p2 = lds128f(sb_s + 16 * qq);  // (scale, bias) of two columns, broadcast
    float x0 = fmaf(__uint_as_float(v[2 * qq]), p2.x, p2.y);
    float x1 = fmaf(__uint_as_float(v[2 * qq + 1]), p2.z, p2.w);
    if (kRes) {
      x0 += bf16_lo(rr[qq]);
      x1 += bf16_hi(rr[qq]);
    }
    x0 = fmaxf(x0, relu_floor);
    x1 = fmaxf(x1, relu_floor);
    o[qq] = kF16 ? pack_f16x2(x0, x1) : pack_bf16x2(x0, x1);
  }
  sts128(a0, o[0], o[1], o[2], o[3]);
  sts128(a1, o[4], o[5], o[6], o[7]);
}

// One warp converts its 32 rows x ncols (16 / 32 / 64) accumulator block into the slab at shared
// address buf_s (rows of row_bytes = ncols * 2).  taddr = TMEM address of (lane quarter, first column);
// sb_s = shared address of the (scale, bias) pair of the chunk's first column.
template <bool kRes, bool kF16 = false>
__device__ __forceinline__ void epi_convert_chunk(uint32_t taddr, int ncols, uint32_t buf_s, uint32_t row_bytes,
                                                  uint32_t swz_mask, int lane, uint32_t sb_s, float relu_floor) {
  if (ncols >= 32) {
    for (int j0 = 0; j0 < ncols; j0 += 32) {
      uint32_t v[32];
      tmem_ld32(taddr + j0, v);
      tmem_ld_wait();
      const uint32_t off0 = lane * row_bytes + j0 * 2;
      epi_convert16<kRes, kF16>(v, sb_s + j0 * 8, buf_s, off0, swz_mask, relu_floor);
      epi_convert16<kRes, kF16>(v + 16, sb_s + (j0 + 16) * 8, buf_s, off0 + 32, swz_mask, relu_floor);
    }
  } else {
    uint32_t v[16];
    tmem_ld16(taddr, v);
    tmem_ld_wait();
    epi_convert16<kRes, kF16>(v, sb_s, buf_s, lane * row_bytes, swz_mask, relu_floor);
  }
}

// Same conversion with the per-channel (scale, bias) read from global memory (read-only path, L1
// resident): scale == nullptr means the BatchNorm scale is already folded into the weights.
template <bool kRes>
__device__ __forceinline__ void epi_convert16_g(const uint32_t* v, const float* scale, const float* bias,
                                                uint32_t buf_s, uint32_t off0, uint32_t swz_mask, float relu_floor) {
  const uint32_t a0 = buf_s + (off0 ^ (((off0 >> 7) & swz_mask) << 4));
  const uint32_t off1 = off0 + 16;
  const uint32_t a1 = buf_s + (off1 ^ (((off1 >> 7) & swz_mask) << 4));
  uint32_t rr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (kRes) {
    const uint4 r0 = lds128(a0), r1 = lds128(a1);
    rr[0] = r0.x; rr[1] = r0.y; rr[2] = r0.z; rr[3] = r0.w;
    rr[4] = r1.x; rr[5] = r1.y; rr[6] = r1.z; rr[7] = r1.w;
  }
  float x[16];
#pragma unroll
  for (int qq = 0; qq < 4; ++qq) {
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias) + qq);
    if (scale) {
      const float4 s4 = __ldg(reinterpret_cast<const float4*>(scale) + qq);
      x[4 * qq + 0] = fmaf(__uint_as_float(v[4 * qq + 0]), s4.x, b4.x);
      x[4 * qq + 1] = fmaf(__uint_as_float(v[4 * qq + 1]), s4.y, b4.y);
      x[4 * qq + 2] = fmaf(__uint_as_float(v[4 * qq + 2]), s4.z, b4.z);
      x[4 * qq + 3] = fmaf(__uint_as_float(v[4 * qq + 3]), s4.w, b4.w);
    } else {
      x[4 * qq + 0] = __uint_as_float(v[4 * qq + 0]) + b4.x;
      x[4 * qq + 1] = __uint_as_float(v[4 * qq + 1]) + b4.y;
      x[4 * qq + 2] = __uint_as_float(v[4 * qq + 2]) + b4.z;
      x[4 * qq + 3] = __uint_as_float(v[4 * qq + 3]) + b4.w;
    }
  }
  uint32_t o[8];
#pragma unroll
  for (int qq = 0; qq < 8; ++qq) {
    float x0 = x[2 * qq], x1 = x[2 * qq + 1];
    if (kRes) {
      x0 += bf16_lo(rr[qq]);
      x1 += bf16_hi(rr[qq]);
    }
    o[qq] = pack_bf16x2(fmaxf(x0, relu_floor), fmaxf(x1, relu_floor));
  }
  sts128(a0, o[0], o[1], o[2], o[3]);
  sts128(a1, o[4], o[5], o[6], o[7]);
}

template <bool kRes>
__device__ __forceinline__ void epi_convert_chunk_g(uint32_t taddr, int ncols, uint32_t buf_s, uint32_t row_bytes,
                                                    uint32_t swz_mask, int lane, const float* scale,
                                                    const float* bias, float relu_floor) {
  if (ncols >= 32) {
    for (int j0 = 0; j0 < ncols; j0 += 32) {
      uint32_t v[32];
      tmem_ld32(taddr + j0, v);
      tmem_ld_wait();
      const uint32_t off0 = lane * row_bytes + j0 * 2;
      epi_convert16_g<kRes>(v, scale ? scale + j0 : nullptr, bias + j0, buf_s, off0, swz_mask, relu_floor);
      epi_convert16_g<kRes>(v + 16, scale ? scale + j0 + 16 : nullptr, bias + j0 + 16, buf_s, off0 + 32, swz_mask,
                            relu_floor);
    }
  } else {
    uint32_t v[16];
    tmem_ld16(taddr, v);
    tmem_ld_wait();
    epi_convert16_g<kRes>(v, scale, bias, buf_s, lane * row_bytes, swz_mask, relu_floor);
  }
}

// Division-free walk over run = run0, run0 + step, ... decomposed as run = (n * tdim + t) * yb_count + yb.
struct TileCursor {
  int run, yb, t, n;
  int step, dyb, dt, dn, yb_count, tdim;
  __device__ __forceinline__ void init(int run0, int step_, int yb_count_, int tdim_) {
    step = step_; yb_count = yb_count_; tdim = tdim_;
    run = run0;
    yb = run0 % yb_count_;
    int r = run0 / yb_count_;
    t = r % tdim_;
    n = r / tdim_;
    dyb = step_ % yb_count_;
    r = step_ / yb_count_;
    dt = r % tdim_;
    dn = r / tdim_;
  }
  __device__ __forceinline__ void next() {
    run += step;
    yb += dyb;
    if (yb >= yb_count) { yb -= yb_count; ++t; }
    t += dt;
    if (t >= tdim) { t -= tdim; ++n; }
    n += dn;
  }
};

}  // namespace vsb
