// Fused identity bottleneck for THIN blocks (bottleneck width d = 8 or 16, block width c = 4 d: the Fast pathway's
// res2 / res3 of SlowFast), algo 1 of vsb_bottleneck_*:  out = relu(x + BN_c(c(relu(BN_b(b(relu(BN_a(a(x)))))))))
// in ONE launch, with the matrix products on WARP-LEVEL tensor-core MMAs (mma.sync m16n8k16 / m16n8k8, bf16 in,
// fp32 accumulate) instead of tcgen05.
//
// Reference ops replaced: BottleneckTransform a -> b -> c (SlowFast/slowfast/models/resnet_helper.py:225-240) and the
// identity ResBlock around it (resnet_helper.py:352-358), as bottleneck_fused_sm100.cu (algo 0).
//
// Why not tcgen05 here.  With 8 / 16 output channels a tcgen05.mma is N <= 64 wide whatever the restatement, i.e. at
// the ~60-clk issue floor of one elected thread (profiles/r02_umma_commit_pitch_shift_rate.json), and the x tiles of
// the flat-raster kernel move as 5-D TMA boxes at 10 - 21 B/clk/SM: algo 0 is bound by issue, not by HBM, and ends at
// parity with the three launches.  A warp-level MMA is exactly 8 channels wide, sixteen warps issue independently,
// and the operands are plain shared-memory / register fragments.
//
// Schedule.  A CTA owns a strip of R output rows of one clip (full width) and WALKS it through time.  Per frame the
// strip plus one halo row above and below is ONE contiguous range of the NTHWC tensor, fetched by a single
// cp.async.bulk (no tensor map) into a ring of S frame slots; every frame is fetched once and used by the three
// temporal taps of `a` and by the residual.  Warp roles (16 warps of 128 registers; the a : bc split is a plan parameter):
//   a-warps  (7)  : a[t] = relu(BN(sum_taps x[t + tap - 1] . Wa[tap])) for the strip + halo rows, 16-pixel MMA tiles,
//                   Wa fragments resident in registers; result -> bf16 -> one of two a-tile buffers in shared memory
//                   laid out (R + 2) x (W + 2) pixels with a zero border (= the padding of the 3x3 conv)
//   bc-warps (8)  : b = relu(BN(sum_taps a[.. + dy, .. + dx] . Wb[tap])) from the a-tile buffer (every tap is a shifted
//                   shared-memory read), the b accumulator fragment IS the A fragment of c (no exchange), c 32 channels
//                   at a time, + BN + residual (x[t] from the ring) + ReLU -> 16-byte global stores
//   fetch warp    : one lane waits for a free slot, arms the slot's mbarrier and issues the bulk copy (merged into a
//                   compute warp it stalled that warp's tiles: 0.31 vs 0.22 ms per res2 block)
// The two a-tile buffers let the a-warps run one frame ahead of the bc-warps.  All handshakes are mbarriers on which
// every lane arrives for its own accesses, with bounded back-off waits (a protocol bug traps, it never hangs the GPU).
// Measured on B200 (batch 64): res2 identity block 0.189 ms (three launches: 0.385), res2 block 0 with its projection
// 0.144 ms (0.30), res3 identity block 0.181 ms (0.208); bound by instruction issue and shared-memory wavefronts (ncu:
// profiles/r02c_ncu_full_thin_bottleneck.txt), HBM time of the res2 block incl. the 1.25x halo re-reads: 0.142 ms.
//
// Channel permutation.  The K order of `a` and the N order of `c` are free as long as weights and activations agree,
// so thread (g, t) of a warp owns, for pixel rows g and g + 8, the 16-byte pieces [8 (4 q + t), + 8) of the pixel's
// channels (q = 0 .. c / 32 - 1): its x fragments are 128-bit shared loads, its residual is already in the layout of
// the c accumulator, and its output is one 16-byte store per 32 channels; the weight fragments are gathered with the
// same permutation once per CTA.  Pixel rows of 128 bytes and more make two pixel rows of a quarter-warp hit the
// same banks; the conflict-free variant (lanes with odd g fetch the pieces of q and q ^ 1 in swapped order and swap them
// back in registers, VSB_THIN_SWAP_PIECES) costs 8 selects per pair and pixel row and was measured no faster (res3 block 0.182 vs 0.181 ms).
#include <cuda.h>
#include <cuda_bf16.h>

#include <stdlib.h>

#include <new>

#include "bottleneck_thin.h"
#include "common.h"
#include "ptx.cuh"

namespace vsb {

namespace {
// a-warps + bc-warps (split chosen by the plan) + one fetch warp = 16 warps of 128 registers.  (24 warps of 80
// registers were measured slower for d = 8, 0.28 - 0.31 vs 0.22 ms per res2 block at batch 64: the kernel is bound by
// instruction issue and shared-memory wavefronts, not by exposed latency.)
__host__ __device__ constexpr int compute_warps(int) { return 15; }
__host__ __device__ constexpr int default_a_warps(int) { return 7; }
// 16-pixel tiles per warp and frame step (unrolled)
__host__ __device__ constexpr int max_a_tiles(int d) { return d == 8 ? 6 : 3; }
__host__ __device__ constexpr int max_bc_tiles(int d) { return d == 8 ? 4 : 2; }
constexpr int kMaxSlots = 8;
// mbarriers are spaced 64 bytes apart (uint64 units): packed as adjacent 8-byte words, three frame barriers armed back
// to back made compute-sanitizer report the first one as uninitialised (synccheck) and the ring fills as racing with
// the slot's readers (racecheck); spaced, both tools are clean
constexpr int kBarStride = 8;
constexpr uint32_t kRingOffset = (2 * kMaxSlots + 4) * kBarStride * 8;  // the mbarriers sit in front of the ring
}  // namespace

struct ThinParams {
  const uint8_t* x;
  uint8_t* out;
  const __nv_bfloat16 *wa, *wb, *wc;
  const float *sa, *ba, *sb, *bb, *sc, *bc;
  int n, T, H, W;
  int R, row_tiles;
  int a_warps;  // warps 0 .. a_warps-1 run conv a, the other compute warps conv b + c, the last warp fetches
  uint32_t out_pitch_bytes;
  int S;
  uint32_t slot_bytes, row_bytes;
  uint32_t x_px_bytes;  // bytes per x pixel (projection blocks: the pitch may exceed the 8 channels read)
  int a_px, a_tiles, bc_px, bc_tiles;
  uint32_t a_buf_bytes;
  uint32_t off_abuf, off_sbc, off_bar;
  uint32_t magic_w;  // floor(2^24 / W) + 1: px / W == (px * magic_w) >> 24 for px < 2^24 / W
  long long total_steps;
  int steps_per_cta;
};

namespace {

__device__ __forceinline__ uint32_t t_lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint4 t_lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float4 t_lds128f(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void t_sts32(uint32_t a, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void t_stg128(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// relu + one rounding to bf16 of two floats (lo -> bits 0-15): relu(round(v)) == round(relu(v))
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// D (16 x 8, fp32) += A (16 x 16, bf16, row) . B (16 x 8, bf16, col)
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// D = A . B (no accumulator input: no register initialisation)
__device__ __forceinline__ void mma16816_z(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                           uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%10, %10, %10, %10};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
}
// D (16 x 8, fp32) = A (16 x 8, bf16, row) . B (8 x 8, bf16, col)
__device__ __forceinline__ void mma1688_z(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%7, %7, %7, %7};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(a0), "r"(a1), "r"(b0), "f"(0.f));
}
// D (16 x 8, fp32) += A (16 x 8, bf16, row) . B (8 x 8, bf16, col)
__device__ __forceinline__ void mma1688(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(b0));
}
// contiguous global -> shared copy by the TMA unit (no tensor map), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst_s, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_s),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint32_t sel(bool c, uint32_t a, uint32_t b) { return c ? a : b; }

// mbarrier wait with back-off: a warp that spins on try_wait takes issue slots from the working warps of its
// scheduler, so a failed attempt sleeps before the next one.  Bounded: a protocol bug traps, it never hangs the GPU.
__device__ __forceinline__ void t_mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  for (;;) {
    __nanosleep(40);
    if (mbar_try_wait(bar, parity)) return;
    if (++spins > (1u << 24)) __trap();
  }
}

// Ring position of the frames around t: frame t sits in (slot, round) -- load index L = round * S + slot -- and
// advances by one slot per frame step; the neighbours t - 1 / t + 1 are one slot back / ahead.  No division per step.
struct RingPos {
  uint32_t slot, round;
  __device__ __forceinline__ void set(uint32_t L, uint32_t S) {
    round = L / S;
    slot = L - round * S;
  }
  __device__ __forceinline__ void step(uint32_t S) {
    if (++slot == S) {
      slot = 0;
      ++round;
    }
  }
  __device__ __forceinline__ RingPos at(int delta, uint32_t S) const {  // delta = -1, 0, +1
    RingPos r = *this;
    if (delta > 0) r.step(S);
    if (delta < 0) {
      if (r.slot == 0) {
        r.slot = S - 1;
        --r.round;
      } else {
        --r.slot;
      }
    }
    return r;
  }
};

// The 16-byte pieces (q = 0 .. NQ-1) of one pixel row for thread t: v[q] = bytes [64 q + 16 t, + 16) of the pixel;
// addr = pixel address + 16 t.
template <int NQ>
__device__ __forceinline__ void load_pieces(uint32_t addr, bool odd, uint4 (&v)[NQ]) {
  if (NQ == 1) {
    v[0] = t_lds128(addr);
  } else {
    const uint32_t swap = odd ? 64u : 0u;
#pragma unroll
    for (int m = 0; m < NQ / 2; ++m) {
      const uint4 first = t_lds128(addr + 128 * m + swap);
      const uint4 second = t_lds128(addr + 128 * m + (64u - swap));
      v[2 * m].x = sel(odd, second.x, first.x); v[2 * m].y = sel(odd, second.y, first.y);
      v[2 * m].z = sel(odd, second.z, first.z); v[2 * m].w = sel(odd, second.w, first.w);
      v[2 * m + 1].x = sel(odd, first.x, second.x); v[2 * m + 1].y = sel(odd, first.y, second.y);
      v[2 * m + 1].z = sel(odd, first.z, second.z); v[2 * m + 1].w = sel(odd, first.w, second.w);
    }
  }
}

__device__ __forceinline__ uint32_t piece_reg(const uint4& v, int i) {
  return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

// One walk = the frames [t0, t1) of one (clip, row strip) unit that a CTA processes back to back; the frames
// [f_lo, f_hi] are fetched for it (one halo frame on either side for the temporal taps), in this order.
struct ThinWalk {
  long long cur, end;
  int n, r0, t0, t1, f_lo, f_hi;
  uint32_t l0;  // load index of frame f_lo (running count of fetched frames of this CTA)
  __device__ __forceinline__ void init(const ThinParams& p) {
    cur = (long long)blockIdx.x * p.steps_per_cta;
    end = cur + p.steps_per_cta < p.total_steps ? cur + p.steps_per_cta : p.total_steps;
    l0 = 0;
  }
  template <int HT>
  __device__ __forceinline__ bool load(const ThinParams& p) {
    if (cur >= end) return false;
    const long long unit = cur / p.T;
    t0 = (int)(cur - unit * p.T);
    long long len = end - cur;
    if (len > p.T - t0) len = p.T - t0;
    t1 = t0 + (int)len;
    n = (int)(unit / p.row_tiles);
    r0 = (int)(unit - (long long)n * p.row_tiles) * p.R;
    f_lo = t0 - HT > 0 ? t0 - HT : 0;
    f_hi = t1 - 1 + HT < p.T - 1 ? t1 - 1 + HT : p.T - 1;
    return true;
  }
  __device__ __forceinline__ void next() {
    l0 += (uint32_t)(f_hi - f_lo + 1);
    cur += t1 - t0;
  }
};

}  // namespace

// PROJ: block 0 of a stage with a 1x1x1 projection shortcut and unit strides (d = 8 only: the Fast pathway's res2):
// x has d channels, conv c and the shortcut are ONE K = 16 MMA per column tile -- [b | x[t]] . [Wc' ; W1'] with the two
// BatchNorm scales folded into the weights as ratios to a common per-channel scale, exactly as the three-launch
// engine's fused-shortcut conv does (engine._conv_with_shortcut) -- and there is no residual.
template <int D, int KT, bool PROJ>
__global__ void __launch_bounds__((compute_warps(D) + 1) * 32, 1) bottleneck_thin_kernel(const ThinParams p) {
  static_assert(!PROJ || D == 8, "projection blocks: d = 8 only");
  constexpr int kComputeWarps = compute_warps(D), kThinThreads = (kComputeWarps + 1) * 32;
  constexpr int C = PROJ ? D : 4 * D;  // channels of x
  constexpr int NQ = D / 8;       // 32-channel groups of the output = 8-channel slices of the bottleneck width
  constexpr int NTD = D / 8;      // 8-column MMA tiles of the bottleneck width
  constexpr int HT = KT / 2;
  constexpr int CBc = C * 2;      // bytes per x pixel (identity blocks: dense)
  const uint32_t CB = PROJ ? p.x_px_bytes : (uint32_t)CBc;
  constexpr int NSL = 9 * NQ;     // 8-channel K slices of conv b (tap-major)
  constexpr int NP = (NSL + 1) / 2;
  constexpr int KSC = D >= 16 ? D / 16 : 1;  // K steps of conv c
  constexpr int AP = D == 8 ? 16 : 2 * D + 16;  // bytes per a-tile pixel (48 for d = 16: conflict-free 32-bit reads)
  constexpr int kMaxATiles = max_a_tiles(D), kMaxBCTiles = max_bc_tiles(D);

  extern __shared__ uint8_t thin_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(thin_smem_raw) + 127) & ~uintptr_t(127));
  // shared-memory layout: [mbarriers | frame ring | two a-tile buffers | tables]
  const uint32_t base_s = smem_u32(smem);
  const uint32_t ring_s = base_s + kRingOffset;
  const uint32_t abuf_s = base_s + p.off_abuf;
  const uint32_t sbc_s = base_s + p.off_sbc;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  // one mbarrier per 64 bytes
  uint64_t* empty = full + kMaxSlots * kBarStride;
  uint64_t* a_full = empty + kMaxSlots * kBarStride;
  uint64_t* a_empty = a_full + 2 * kBarStride;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
#ifndef VSB_THIN_SWAP_PIECES
#define VSB_THIN_SWAP_PIECES 0
#endif
  // conflict-free pair loads of conv a's x pieces (see the file comment) cost 8 selects per pair and pixel row; with
  // plain loads the two pixel rows of a quarter-warp collide (2 wavefronts more per load): measured no faster with the
  // swap (VSB_THIN_SWAP_PIECES=1), so the simpler loads are the default
  const bool odd = VSB_THIN_SWAP_PIECES ? (g & 1) != 0 : false;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.S; ++s) {
      mbar_init(&full[kBarStride * (s)], 1);
      mbar_init(&empty[kBarStride * (s)], kComputeWarps * 32);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&a_full[kBarStride * (b)], p.a_warps * 32);
      mbar_init(&a_empty[kBarStride * (b)], (kComputeWarps - p.a_warps) * 32);
    }
    fence_mbar_init();
  }
  __syncthreads();
  // zero both a-tile buffers once: the border columns are never written again
  for (uint32_t i = threadIdx.x * 16; i < 2 * p.a_buf_bytes; i += kThinThreads * 16)
    *reinterpret_cast<uint4*>(smem + p.off_abuf + i) = make_uint4(0, 0, 0, 0);
  // (scale, bias) of conv c in the order the c epilogue reads them: entry (q * 4 + i) * 4 + t = channels
  // 8 (4 q + t) + 2 i, + 1
  for (int e = threadIdx.x; e < NQ * 16; e += kThinThreads) {
    const int tt = e & 3, i = (e >> 2) & 3, q = e >> 4;
    const int ch = 8 * (4 * q + tt) + 2 * i;
    reinterpret_cast<float4*>(smem + p.off_sbc)[e] = make_float4(p.sc[ch], p.bc[ch], p.sc[ch + 1], p.bc[ch + 1]);
  }
  // (scale, bias) of conv b, entry nt * 4 + t = channels 8 nt + 2 t, + 1 (d = 16 reads them per tile: registers)
  for (int e = threadIdx.x; e < NTD * 4; e += kThinThreads) {
    const int ch = 8 * (e >> 2) + 2 * (e & 3);
    reinterpret_cast<float4*>(smem + p.off_sbc + 512)[e] = make_float4(p.sb[ch], p.bb[ch], p.sb[ch + 1], p.bb[ch + 1]);
  }
  // d = 16: the conv b weight fragments live in shared memory (36 registers per thread otherwise: spills), entry
  // ((pr * NTD + nt) * 32 + lane) = the (b0, b1) pair of K-slice pair pr, column tile nt for that lane
  if constexpr (D == 16) {
    for (int e = threadIdx.x; e < NP * NTD * 32; e += kThinThreads) {
      const int ln = e & 31, nt = (e >> 5) % NTD, pr = (e >> 5) / NTD;
      const int gg = ln >> 2, tt = ln & 3;
      const int s0 = 2 * pr, s1 = 2 * pr + 1;
      const __nv_bfloat16* row = p.wb + (size_t)(8 * nt + gg) * 9 * D;
      uint2 v;
      v.x = *reinterpret_cast<const uint32_t*>(row + (s0 / NQ) * D + 8 * (s0 % NQ) + 2 * tt);
      v.y = s1 < NSL ? *reinterpret_cast<const uint32_t*>(row + (s1 / NQ) * D + 8 * (s1 % NQ) + 2 * tt) : 0u;
      reinterpret_cast<uint2*>(smem + p.off_sbc + 1024)[e] = v;
    }
    // ... and so do conv c's: entry ((q * 4 + i) * 32 + lane), K = 16 in one step
    for (int e = threadIdx.x; e < NQ * 4 * 32; e += kThinThreads) {
      const int ln = e & 31, i = (e >> 5) & 3, q = e >> 7;
      const int gg = ln >> 2, tt = ln & 3;
      const int co = 8 * (4 * q + (gg >> 1)) + 2 * i + (gg & 1);
      const __nv_bfloat16* row = p.wc + (size_t)co * D + 2 * tt;
      uint2 v;
      v.x = *reinterpret_cast<const uint32_t*>(row);
      v.y = *reinterpret_cast<const uint32_t*>(row + 8);
      reinterpret_cast<uint2*>(smem + p.off_sbc + 1024 + 4608)[e] = v;
    }
  }
  __syncthreads();

  ThinWalk wk;
  wk.init(p);
  const int wp2 = p.W + 2;

  const int kAWarps = p.a_warps, kBCWarps = kComputeWarps - p.a_warps;
  if (warp == kComputeWarps) {
    // ------------------------------------------------------------------ frame fetches
    if (lane == 0) {
      uint32_t slot = 0, round = 0;  // load index L = round * S + slot
      while (wk.load<HT>(p)) {
        const int row_lo = wk.r0 - 1 > 0 ? wk.r0 - 1 : 0;
        const int row_hi = wk.r0 + p.R + 1 < p.H ? wk.r0 + p.R + 1 : p.H;
        const uint32_t bytes = (uint32_t)(row_hi - row_lo) * p.row_bytes;
        const uint32_t dst_off = (uint32_t)(row_lo - (wk.r0 - 1)) * p.row_bytes;
        for (int f = wk.f_lo; f <= wk.f_hi; ++f) {
          if (round > 0) t_mbar_wait(&empty[kBarStride * (slot)], (round - 1) & 1);
          const uint8_t* src = p.x + (((long long)wk.n * p.T + f) * p.H + row_lo) * (long long)p.row_bytes;
          mbar_expect_tx(&full[kBarStride * (slot)], bytes);
          bulk_g2s(ring_s + slot * p.slot_bytes + dst_off, src, bytes, &full[kBarStride * (slot)]);
          if (++slot == (uint32_t)p.S) {
            slot = 0;
            ++round;
          }
        }
        wk.next();
      }
    }
  } else if (warp < kAWarps) {
    // ------------------------------------------------------------------ conv a
    // weight fragments: step (tap, q, h) of K, column tile nt: b0 = Wa[8 nt + g][tap][8 (4 q + t) + 4 h + {0, 1}],
    // b1 = the next two channels
    uint2 wa_f[KT][NQ][2][NTD];
    uint32_t wa_p[KT];  // PROJ: x has 8 channels, tap's fragment = Wa[g][tap][2 t, + 1]
    if constexpr (PROJ) {
#pragma unroll
      for (int tap = 0; tap < KT; ++tap)
        wa_p[tap] = *reinterpret_cast<const uint32_t*>(p.wa + ((size_t)g * KT + tap) * C + 2 * t4);
    } else {
#pragma unroll
      for (int tap = 0; tap < KT; ++tap)
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int nt = 0; nt < NTD; ++nt)
              wa_f[tap][q][h][nt] = *reinterpret_cast<const uint2*>(
                  p.wa + ((size_t)(8 * nt + g) * KT + tap) * C + 8 * (4 * q + t4) + 4 * h);
    }
    float4 sa_f[NTD];  // (scale, bias) of channels 8 nt + 2 t, + 1
#pragma unroll
    for (int nt = 0; nt < NTD; ++nt) {
      const int ch = 8 * nt + 2 * t4;
      sa_f[nt] = make_float4(p.sa[ch], p.ba[ch], p.sa[ch + 1], p.ba[ch + 1]);
    }
    // tile geometry (the same tiles every frame step): per tile slot and pixel row, strip row << 16 | byte offset of
    // the pixel in the a-tile buffer (+ 4 t); 0xFFFF.... = past the strip
    uint32_t geo[kMaxATiles][2];
#pragma unroll
    for (int sl = 0; sl < kMaxATiles; ++sl)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int px = (warp + kAWarps * sl) * 16 + g + 8 * half;
        const int rr = (int)(((uint32_t)px * p.magic_w) >> 24);
        const int col = px - rr * p.W;
        geo[sl][half] = px < p.a_px ? ((uint32_t)rr << 16) | (uint32_t)((rr * wp2 + col + 1) * AP + 4 * t4) : 0xFFFF0000u;
      }
    uint32_t istep = 0;
    const uint32_t S = (uint32_t)p.S;
    while (wk.load<HT>(p)) {
      // strip rows [rr_lo, rr_hi) lie inside the image; the others are the zero padding of conv b
      const uint32_t rr_lo = wk.r0 == 0 ? 1u : 0u;
      const uint32_t rr_hi = (uint32_t)(p.H - wk.r0 + 1 < p.R + 2 ? p.H - wk.r0 + 1 : p.R + 2);
      RingPos pos;  // of frame t
      pos.set(wk.l0 + (uint32_t)(wk.t0 - wk.f_lo), S);
      for (int t = wk.t0; t < wk.t1; ++t, ++istep, pos.step(S)) {
        const uint32_t b = istep & 1;
        if (istep >= 2) t_mbar_wait(&a_empty[kBarStride * (b)], ((istep >> 1) - 1) & 1);
        uint32_t slot_s[KT];
        bool have[KT];
#pragma unroll
        for (int tap = 0; tap < KT; ++tap) {
          const int f = t + tap - HT;
          have[tap] = f >= 0 && f < p.T;
          slot_s[tap] = ring_s;
          if (have[tap]) {
            const RingPos rp = pos.at(tap - HT, S);
            t_mbar_wait(&full[kBarStride * (rp.slot)], rp.round & 1);
            slot_s[tap] = ring_s + rp.slot * p.slot_bytes + (PROJ ? 4 : 16) * t4;
          }
        }
        const uint32_t dst_s = abuf_s + b * p.a_buf_bytes;
#pragma unroll
        for (int sl = 0; sl < kMaxATiles; ++sl) {
          const int tile = warp + kAWarps * sl;
          if (tile >= p.a_tiles) break;
          const int px_g = tile * 16 + g, px_h = px_g + 8;
          const uint32_t off_g = (uint32_t)(px_g < p.a_px ? px_g : p.a_px - 1) * CB;
          const uint32_t off_h = (uint32_t)(px_h < p.a_px ? px_h : p.a_px - 1) * CB;
          float acc[NTD][4];
          if constexpr (PROJ) {
            // 8 channels per tap: taps 0 and 1 are the two halves of one K = 16 MMA, tap 2 a K = 8 MMA; a missing
            // frame (temporal padding) is a zero fragment
            uint32_t xg[KT], xh[KT];
#pragma unroll
            for (int tap = 0; tap < KT; ++tap) {
              xg[tap] = have[tap] ? t_lds32(slot_s[tap] + off_g) : 0u;
              xh[tap] = have[tap] ? t_lds32(slot_s[tap] + off_h) : 0u;
            }
            if (KT == 3) {
              mma16816_z(acc[0], xg[0], xh[0], xg[1], xh[1], wa_p[0], wa_p[1]);
              mma1688(acc[0], xg[KT - 1], xh[KT - 1], wa_p[KT - 1]);
            } else {
              mma1688_z(acc[0], xg[0], xh[0], wa_p[0]);
            }
          } else {
#pragma unroll
          for (int nt = 0; nt < NTD; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
          for (int tap = 0; tap < KT; ++tap) {
            if (!have[tap]) continue;
            uint4 vg[NQ], vh[NQ];
            load_pieces<NQ>(slot_s[tap] + off_g, odd, vg);
            load_pieces<NQ>(slot_s[tap] + off_h, odd, vh);
#pragma unroll
            for (int q = 0; q < NQ; ++q)
#pragma unroll
              for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int nt = 0; nt < NTD; ++nt)
                  mma16816(acc[nt], piece_reg(vg[q], 2 * h), piece_reg(vh[q], 2 * h), piece_reg(vg[q], 2 * h + 1),
                           piece_reg(vh[q], 2 * h + 1), wa_f[tap][q][h][nt].x, wa_f[tap][q][h][nt].y);
          }
          }
          // BN + ReLU -> bf16 -> a-tile buffer; rows outside the image are the zero padding of conv b
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint32_t ge = geo[sl][half], rr = ge >> 16;
            if (rr != 0xFFFFu) {
              const bool inside = rr >= rr_lo && rr < rr_hi;
              const uint32_t dst = dst_s + (ge & 0xFFFFu);
#pragma unroll
              for (int nt = 0; nt < NTD; ++nt) {
                const uint32_t v = pack_relu_bf16x2(fmaf(acc[nt][2 * half], sa_f[nt].x, sa_f[nt].y),
                                                    fmaf(acc[nt][2 * half + 1], sa_f[nt].z, sa_f[nt].w));
                t_sts32(dst + 16 * nt, inside ? v : 0u);
              }
            }
          }
          }
        // every lane arrives for its own shared-memory accesses (release / acquire per thread: nothing rests on
        // cumulativity through a warp barrier, and the race checker can follow it)
        {
          mbar_arrive(&a_full[kBarStride * (b)]);
          // frames conv a no longer needs
          const int f_done = t - HT;
          if (f_done >= wk.f_lo) mbar_arrive(&empty[kBarStride * (pos.at(-HT, S).slot)]);
          if (t == wk.t1 - 1)
            for (int f = (f_done + 1 > wk.f_lo ? f_done + 1 : wk.f_lo); f <= wk.f_hi; ++f)
              mbar_arrive(&empty[kBarStride * (pos.at(f - t, S).slot)]);
        }
      }
      wk.next();
    }
  } else {
    // ------------------------------------------------------------------ conv b, conv c, residual, store
    const int bw = warp - kAWarps;
    // conv b: K slices s = tap * NQ + q of 8 channels, two per MMA: b0 = Wb[8 nt + g][tap][8 q + 2 t, + 1]
    uint2 wb_f[NP][NTD];
    const uint32_t wb_s = sbc_s + 1024 + (uint32_t)lane * 8;
    if constexpr (D == 8)
#pragma unroll
    for (int pr = 0; pr < NP; ++pr)
#pragma unroll
      for (int nt = 0; nt < NTD; ++nt) {
        const int s0 = 2 * pr, s1 = 2 * pr + 1;
        const __nv_bfloat16* row = p.wb + (size_t)(8 * nt + g) * 9 * D;
        wb_f[pr][nt].x = *reinterpret_cast<const uint32_t*>(row + (s0 / NQ) * D + 8 * (s0 % NQ) + 2 * t4);
        wb_f[pr][nt].y = s1 < NSL ? *reinterpret_cast<const uint32_t*>(row + (s1 / NQ) * D + 8 * (s1 % NQ) + 2 * t4) : 0u;
      }
    // conv c: column tile (q, i) holds the output channels 8 (4 q + n / 2) + 2 i + n % 2 in its column n
    uint2 wc_f[NQ][4][KSC];
    if constexpr (D == 8)
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int ks = 0; ks < KSC; ++ks) {
          const int co = 8 * (4 * q + (g >> 1)) + 2 * i + (g & 1);
          // PROJ: rows of [Wc' | W1'], K = 8 + 8
          const __nv_bfloat16* row = p.wc + (size_t)co * (PROJ ? 2 * D : D) + 16 * ks + 2 * t4;
          wc_f[q][i][ks].x = *reinterpret_cast<const uint32_t*>(row);
          wc_f[q][i][ks].y = (D >= 16 || PROJ) ? *reinterpret_cast<const uint32_t*>(row + 8) : 0u;
        }
    float4 sb_f[NTD];  // d = 8 only: the (scale, bias) pairs of conv b stay in registers
    if (D == 8) {
#pragma unroll
      for (int nt = 0; nt < NTD; ++nt) sb_f[nt] = t_lds128f(sbc_s + 512 + (uint32_t)(nt * 4 + t4) * 16);
    }
    // tile geometry: per tile slot and pixel row, byte offset of the pixel's centre in the a-tile buffer (+ 4 t) |
    // pixel index in the strip << 16 | strip row << 26 (row 63 = past the strip)
    uint32_t geo[kMaxBCTiles][2];
#pragma unroll
    for (int sl = 0; sl < kMaxBCTiles; ++sl)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int px = (bw + kBCWarps * sl) * 16 + g + 8 * half;
        const int pc = px < p.bc_px ? px : p.bc_px - 1;
        const int rr = (int)(((uint32_t)pc * p.magic_w) >> 24);
        const int col = pc - rr * p.W;
        geo[sl][half] = (uint32_t)(((rr + 1) * wp2 + col + 1) * AP + 4 * t4) | ((uint32_t)pc << 16) |
                        ((uint32_t)(px < p.bc_px ? rr : 63) << 26);
      }
    const uint32_t res_base = (uint32_t)p.W * CB + (PROJ ? 4 : 16) * t4;  // the slot starts one row above the strip
    const uint32_t rowb = (uint32_t)wp2 * AP;
    uint32_t istep = 0;
    const uint32_t S = (uint32_t)p.S;
    while (wk.load<HT>(p)) {
      const uint32_t rr_hi = (uint32_t)(p.H - wk.r0 < p.R ? p.H - wk.r0 : p.R);
      RingPos pos;  // of frame t
      pos.set(wk.l0 + (uint32_t)(wk.t0 - wk.f_lo), S);
      for (int t = wk.t0; t < wk.t1; ++t, ++istep, pos.step(S)) {
        const uint32_t b = istep & 1;
        t_mbar_wait(&a_full[kBarStride * (b)], (istep >> 1) & 1);
        const uint32_t slot = pos.slot;
        t_mbar_wait(&full[kBarStride * (slot)], pos.round & 1);
        const uint32_t xs = ring_s + slot * p.slot_bytes;
        const uint32_t src_s = abuf_s + b * p.a_buf_bytes;
        uint8_t* out_frame = p.out + (((long long)wk.n * p.T + t) * p.H + wk.r0) * (long long)p.W * p.out_pitch_bytes + 16 * t4;
#pragma unroll
        for (int sl = 0; sl < kMaxBCTiles; ++sl) {
          if (bw + kBCWarps * sl >= p.bc_tiles) break;
          const uint32_t g0 = geo[sl][0], g1 = geo[sl][1];
          const uint32_t ctr0 = src_s + (g0 & 0xFFFFu), ctr1 = src_s + (g1 & 0xFFFFu);
          float accb[NTD][4];
#pragma unroll
          for (int pr = 0; pr < NP; ++pr) {
            const int s0 = 2 * pr, s1 = 2 * pr + 1;
            const int tap0 = s0 / NQ, q0 = s0 % NQ;
            const uint32_t sh0 = (uint32_t)(tap0 / 3) * rowb - rowb + (uint32_t)((tap0 % 3 - 1) * AP + 16 * q0);
            const uint32_t a0 = t_lds32(ctr0 + sh0), a1 = t_lds32(ctr1 + sh0);
            uint32_t a2 = 0u, a3 = 0u;
            if (s1 < NSL) {
              const int tap1 = s1 / NQ, q1 = s1 % NQ;
              const uint32_t sh1 = (uint32_t)(tap1 / 3) * rowb - rowb + (uint32_t)((tap1 % 3 - 1) * AP + 16 * q1);
              a2 = t_lds32(ctr0 + sh1);
              a3 = t_lds32(ctr1 + sh1);
            }
#pragma unroll
            for (int nt = 0; nt < NTD; ++nt) {
              uint2 wv;
              if constexpr (D == 8) {
                wv = wb_f[pr][nt];
              } else {
                asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(wv.x), "=r"(wv.y) : "r"(wb_s + (uint32_t)(pr * NTD + nt) * 256));
              }
              if (pr == 0) mma16816_z(accb[nt], a0, a1, a2, a3, wv.x, wv.y);
              else mma16816(accb[nt], a0, a1, a2, a3, wv.x, wv.y);
            }
          }
          // BN + ReLU -> bf16: the accumulator fragment of b is the A fragment of c
          uint32_t pb[NTD][2];
#pragma unroll
          for (int nt = 0; nt < NTD; ++nt) {
            const float4 s4 = D == 8 ? sb_f[nt] : t_lds128f(sbc_s + 512 + (uint32_t)(nt * 4 + t4) * 16);
            pb[nt][0] = pack_relu_bf16x2(fmaf(accb[nt][0], s4.x, s4.y), fmaf(accb[nt][1], s4.z, s4.w));
            pb[nt][1] = pack_relu_bf16x2(fmaf(accb[nt][2], s4.x, s4.y), fmaf(accb[nt][3], s4.z, s4.w));
          }
          // residual = x[t] at the tile's pixels, already in the c accumulator's layout (plain loads: two-way bank
          // conflicts for d = 16, but half the registers of the swapped pair loads conv a uses)
          const uint32_t pc0 = (g0 >> 16) & 0x3FFu, pc1 = (g1 >> 16) & 0x3FFu;
          const uint32_t res0 = xs + res_base + pc0 * CB, res1 = xs + res_base + pc1 * CB;
          const bool valid0 = (g0 >> 26) < rr_hi, valid1 = (g1 >> 26) < rr_hi;
          uint8_t* out0 = out_frame + (size_t)pc0 * p.out_pitch_bytes;
          uint8_t* out1 = out_frame + (size_t)pc1 * p.out_pitch_bytes;
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            uint32_t og[4], oh[4];
            uint4 rgq = make_uint4(0, 0, 0, 0), rhq = rgq;
            uint32_t xg = 0, xh = 0;  // PROJ: x[t] at the tile's pixels, channels 2 t, + 1 = the second K half of conv c
            if constexpr (PROJ) {
              xg = t_lds32(res0);
              xh = t_lds32(res1);
            } else {
              rgq = t_lds128(res0 + 64 * q);
              rhq = t_lds128(res1 + 64 * q);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float accc[4];
              if constexpr (PROJ) {
                mma16816_z(accc, pb[0][0], pb[0][1], xg, xh, wc_f[q][i][0].x, wc_f[q][i][0].y);
              } else if (D >= 16) {
                static_assert(D <= 16, "conv c fragments in shared memory: one K step");
                uint2 wv;
                asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(wv.x), "=r"(wv.y) : "r"(wb_s + 4608 + (uint32_t)(q * 4 + i) * 256));
                mma16816_z(accc, pb[0][0], pb[0][1], pb[1 % NTD][0], pb[1 % NTD][1], wv.x, wv.y);
              } else {
                mma1688_z(accc, pb[0][0], pb[0][1], wc_f[q][i][0].x);
              }
              const float4 s4 = t_lds128f(sbc_s + (uint32_t)((q * 4 + i) * 4 + t4) * 16);
              if constexpr (PROJ) {
                og[i] = pack_relu_bf16x2(fmaf(accc[0], s4.x, s4.y), fmaf(accc[1], s4.z, s4.w));
                oh[i] = pack_relu_bf16x2(fmaf(accc[2], s4.x, s4.y), fmaf(accc[3], s4.z, s4.w));
              } else {
                const uint32_t r_g = piece_reg(rgq, i), r_h = piece_reg(rhq, i);
                og[i] = pack_relu_bf16x2(fmaf(accc[0], s4.x, s4.y) + bf16_lo(r_g), fmaf(accc[1], s4.z, s4.w) + bf16_hi(r_g));
                oh[i] = pack_relu_bf16x2(fmaf(accc[2], s4.x, s4.y) + bf16_lo(r_h), fmaf(accc[3], s4.z, s4.w) + bf16_hi(r_h));
              }
            }
            if (valid0) t_stg128(out0 + 64 * q, og[0], og[1], og[2], og[3]);
            if (valid1) t_stg128(out1 + 64 * q, oh[0], oh[1], oh[2], oh[3]);
          }
        }
        {
          mbar_arrive(&a_empty[kBarStride * (b)]);
          mbar_arrive(&empty[kBarStride * (slot)]);
          // halo frames of the walk that no bc step visits
          if (t == wk.t0)
            for (int f = wk.f_lo; f < wk.t0; ++f) mbar_arrive(&empty[kBarStride * (pos.at(f - t, S).slot)]);
          if (t == wk.t1 - 1)
            for (int f = wk.t1; f <= wk.f_hi; ++f) mbar_arrive(&empty[kBarStride * (pos.at(f - t, S).slot)]);
        }
      }
      wk.next();
    }
  }
}

// ------------------------------------------------------------------ host side
struct ThinPlan {
  ThinParams params;
  size_t smem_bytes;
  unsigned grid;
  int d, kt, proj;
};

template <int D, int KT, bool PROJ>
static cudaError_t thin_set_attr() {
  return cudaFuncSetAttribute(bottleneck_thin_kernel<D, KT, PROJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

int thin_plan_create(const vsb_bottleneck_desc* d, ThinPlan** out_plan) {
  *out_plan = nullptr;
  VSB_CHECK_ARG(d->d == 8 || d->d == 16, "warp-MMA bottleneck: bottleneck width (stored) must be 8 or 16");
  VSB_CHECK_ARG(d->c == 4 * d->d, "warp-MMA bottleneck: block width must be 4 x the bottleneck width");
  const int cin = d->cin > 0 ? d->cin : d->c;  // cin != c: block with a projection shortcut (wc = [Wc' | W1'])
  const bool proj = cin != d->c;
  VSB_CHECK_ARG(!proj || (d->d == 8 && cin == 8), "warp-MMA bottleneck: projection blocks need d = cin = 8");
  VSB_CHECK_ARG(d->x_pitch == cin || (proj && d->x_pitch == 16),
                "warp-MMA bottleneck: x must be dense (pitch == channels; projection blocks: 8 or 16): frames are fetched as one range");
  VSB_CHECK_ARG(d->out_pitch >= d->c && d->out_pitch % 8 == 0, "out pitch must be >= c and a multiple of 8 elements");
  VSB_CHECK_ARG(d->w <= 256 && d->h <= 4096, "frame too large");
  ThinParams p{};
  p.x = static_cast<const uint8_t*>(d->x);
  p.out = static_cast<uint8_t*>(d->out);
  p.wa = static_cast<const __nv_bfloat16*>(d->wa);
  p.wb = static_cast<const __nv_bfloat16*>(d->wb);
  p.wc = static_cast<const __nv_bfloat16*>(d->wc);
  p.sa = d->sa; p.ba = d->ba; p.sb = d->sb; p.bb = d->bb; p.sc = d->sc; p.bc = d->bc;
  p.n = d->n; p.T = d->t; p.H = d->h; p.W = d->w;
  p.out_pitch_bytes = (uint32_t)d->out_pitch * 2;
  p.row_bytes = (uint32_t)d->w * d->x_pitch * 2;
  p.x_px_bytes = (uint32_t)d->x_pitch * 2;
  const uint32_t a_pitch = d->d == 8 ? 16u : (uint32_t)d->d * 2 + 16;  // as AP in the kernel
  p.magic_w = (1u << 24) / (uint32_t)d->w + 1;
  int a_warps = d->d == 8 ? default_a_warps(8) : default_a_warps(16);  // experiments: VSB_THIN_AWARPS
  if (getenv("VSB_THIN_AWARPS") && d->d == 8) a_warps = atoi(getenv("VSB_THIN_AWARPS"));
  if (getenv("VSB_THIN_AWARPS16") && d->d == 16) a_warps = atoi(getenv("VSB_THIN_AWARPS16"));
  if (a_warps < 4 || a_warps > compute_warps(d->d) - 4) a_warps = default_a_warps(d->d);
  p.a_warps = a_warps;
  const int min_slots = d->kt + 1;
  // rows per strip: the strip (+ 2 halo rows) is one ring slot; fewest fetched rows per frame among the strips that
  // leave room for min_slots + 1 slots (caller's override: walk_len)
  const long long budget = 227ll * 1024 - 4096;
  int best_r = 0, best_s = 0;
  long long best_cost = 0;
  for (int r = d->h; r >= 1; --r) {
    if (d->walk_len > 0 && r != d->walk_len) continue;
    const long long slot = (long long)(r + 2) * p.row_bytes;
    const long long abuf = (long long)(r + 2) * (d->w + 2) * a_pitch;
    const long long abuf_al = (abuf + 127) / 128 * 128;
    long long s = (budget - 2 * abuf_al - 10240) / slot;
    if (s > 6) s = 6;
    if (d->stages > 0 && s > d->stages) s = d->stages;
    if (s < min_slots) continue;
    if (abuf > 0xFFFF || slot > 0xFFFF || r * d->w > 1023 || r > 60) continue;  // tile geometry is bit-packed
    if (ceil_div((r + 2) * d->w, 16) > a_warps * max_a_tiles(d->d) || ceil_div(r * d->w, 16) > (compute_warps(d->d) - a_warps) * max_bc_tiles(d->d)) continue;
    // cost: rows fetched per frame (halo included), a small penalty for a short ring
    const long long cost = (long long)ceil_div(d->h, r) * (r + 2) * 16 + (s < min_slots + 1 ? 8 : 0);
    if (!best_r || cost < best_cost) {
      best_r = r;
      best_s = (int)s;
      best_cost = cost;
    }
  }
  if (!best_r) {
    set_error("warp-MMA bottleneck: no row strip of a %d x %d x %d frame fits %d ring slots in shared memory", d->h, d->w,
              d->c, min_slots);
    return VSB_ERR_INVALID;
  }
  p.R = best_r;
  p.S = best_s;
  p.row_tiles = ceil_div(d->h, p.R);
  p.slot_bytes = (uint32_t)(p.R + 2) * p.row_bytes;
  p.a_px = (p.R + 2) * d->w;
  p.a_tiles = ceil_div(p.a_px, 16);
  p.bc_px = p.R * d->w;
  p.bc_tiles = ceil_div(p.bc_px, 16);
  p.a_buf_bytes = (uint32_t)(((long long)(p.R + 2) * (d->w + 2) * a_pitch + 127) / 128 * 128);
  // a tile's tail rows may read up to 16 pixels past the last slot: keep the a-tile buffers behind the ring
  p.off_bar = 0;
  p.off_abuf = kRingOffset + (uint32_t)p.S * p.slot_bytes;
  p.off_abuf = (p.off_abuf + 127u) & ~127u;
  p.off_sbc = p.off_abuf + 2 * p.a_buf_bytes;
  const uint32_t off_end = p.off_sbc + 1024 + 4608 + 2048;  // conv c table at +0 (<= 512 B), conv b table at +512, d = 16: conv b / conv c weight fragments at +1024 (4608 B) / +5632 (2048 B)
  const size_t smem_bytes = (size_t)off_end + 128;
  VSB_CHECK_ARG(smem_bytes <= 227 * 1024, "warp-MMA bottleneck: shared memory plan exceeds 227 KiB");
  p.total_steps = (long long)d->n * p.row_tiles * d->t;
  int sms = 148;
  {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) (void)cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    (void)cudaGetLastError();
  }
  if (d->grid > 0 && d->grid < sms) sms = d->grid;
  const long long grid = p.total_steps < sms ? p.total_steps : sms;
  p.steps_per_cta = (int)ceil_div_ll(p.total_steps, grid);

  cudaError_t e = cudaSuccess;
  if (proj) e = d->kt == 3 ? thin_set_attr<8, 3, true>() : thin_set_attr<8, 1, true>();
  else if (d->d == 8 && d->kt == 3) e = thin_set_attr<8, 3, false>();
  else if (d->d == 8) e = thin_set_attr<8, 1, false>();
  else if (d->kt == 3) e = thin_set_attr<16, 3, false>();
  else e = thin_set_attr<16, 1, false>();
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(bottleneck_thin_kernel) failed: %s", cudaGetErrorString(e));
    return VSB_ERR_CUDA;
  }
  ThinPlan* plan = new (std::nothrow) ThinPlan();
  VSB_CHECK_ARG(plan, "out of host memory");
  plan->params = p;
  plan->smem_bytes = smem_bytes;
  plan->grid = (unsigned)ceil_div_ll(p.total_steps, p.steps_per_cta);
  plan->d = d->d;
  plan->kt = d->kt;
  plan->proj = proj ? 1 : 0;
  *out_plan = plan;
  return VSB_OK;
}

int thin_run(const ThinPlan* plan, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (plan->proj && plan->kt == 3)
    bottleneck_thin_kernel<8, 3, true><<<plan->grid, (compute_warps(8) + 1) * 32, plan->smem_bytes, s>>>(plan->params);
  else if (plan->proj)
    bottleneck_thin_kernel<8, 1, true><<<plan->grid, (compute_warps(8) + 1) * 32, plan->smem_bytes, s>>>(plan->params);
  else if (plan->d == 8 && plan->kt == 3)
    bottleneck_thin_kernel<8, 3, false><<<plan->grid, (compute_warps(8) + 1) * 32, plan->smem_bytes, s>>>(plan->params);
  else if (plan->d == 8)
    bottleneck_thin_kernel<8, 1, false><<<plan->grid, (compute_warps(8) + 1) * 32, plan->smem_bytes, s>>>(plan->params);
  else if (plan->kt == 3)
    bottleneck_thin_kernel<16, 3, false><<<plan->grid, (compute_warps(16) + 1) * 32, plan->smem_bytes, s>>>(plan->params);
  else
    bottleneck_thin_kernel<16, 1, false><<<plan->grid, (compute_warps(16) + 1) * 32, plan->smem_bytes, s>>>(plan->params);
  VSB_CHECK_LAUNCH("bottleneck_thin_kernel");
  return VSB_OK;
}

void thin_plan_destroy(ThinPlan* plan) { delete plan; }

void thin_plan_info(const ThinPlan* plan, long long* out8) {
  const ThinParams& p = plan->params;
  out8[0] = p.R;
  out8[1] = p.row_tiles;
  out8[2] = p.S;
  out8[3] = p.steps_per_cta;
  out8[4] = plan->grid;
  out8[5] = (long long)plan->smem_bytes;
  out8[6] = p.a_tiles * 1000 + p.bc_tiles;
  out8[7] = 0;
}

}  // namespace vsb
