// Fused 1x7x7 stride-2 stem: Conv3d + frozen BatchNorm + ReLU + MaxPool3d 1x3x3 stride 2 in ONE kernel.
//
// Reference ops replaced: ResNetBasicStem.forward (SlowFast/slowfast/models/stem_helper.py:157-178: conv [1,7,7]
// stride [1,2,2] pad [0,3,3] -> BN -> ReLU -> MaxPool3d [1,3,3] stride [1,2,2] pad [0,1,1]) of the Slow pathway
// (and of the Slow-only / C2D nets), on the packed bf16 frames vsb_pack_frames writes.  The conv output
// (822 MB per 64 SlowFast clips) never reaches HBM: it is pooled out of shared memory.
//
// Formulation.  The packed input keeps 4 channels per pixel and a 3-pixel zero border on the left of every row, so
// the 7 taps of one filter row of output column c start at buffer pixel 2c: 8 pixels x 4 channels = 32 bf16 = one
// 64-byte K slot, and consecutive output columns' slots start 16 bytes apart.  A TMA tensor map whose W stride
// (16 B) is smaller than its innermost extent (64 B) therefore delivers the im2col row of every output pixel
// straight into shared memory (the overlap costs nothing: the bytes come out of L2 once per slot).  K = 7 filter
// rows x 32, of which 7 x 7 x 3 = 147 carry weights (1.5x padding instead of the 4.6x of the pixel-group
// restatement), N = 64, weights (28 KB) resident.
//
// Tile = 8 x 16 output pixels of one frame (M = 128).  Input rows are split by parity into two sub-windows
// (odd rows feed the even filter rows kh = 0,2,4,6, even rows the odd ones) so that filter row kh of output row r
// is sub-window row r + kh/2: every tap is the same 128-row UMMA operand with its start address advanced by
// (kh/2) x 16 slots = a multiple of 1 KiB (the window trick of conv_win_sm100.cu in 2-D).  21 KB of shared memory
// per tile instead of the 56 KB an im2col tile of 128 pixels needs.
//
// Epilogue: TMEM -> scale/bias/ReLU -> bf16 -> the 8 x 16 x 64 conv tile in shared memory -> 3x3/s2 max over it.
// Pooled outputs whose window lies inside the tile (or hangs over the image edge: -inf padding) are stored;
// the ones on the tile's seams (first pooled row / column needs the previous tile's last conv row / column, and
// this tile's last conv row / column belongs to the next tile's first pooled output) are combined across tiles with
// red.global.max.bf16x2 on a zero-initialised output: post-ReLU values are >= 0, so 0 is the identity, and
// bf16 rounding is monotonic, so max(round(x)) == round(max(x)) bit for bit.
#include <cuda.h>
#include <cuda_bf16.h>

#include <stdlib.h>

#include <new>

#include "common.h"
#include "conv_plan.h"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace vsb {
namespace {

constexpr int kCout = 64;
constexpr int kTaps = 7;
constexpr int kTileH = 8, kTileW = 16;
constexpr int kStages = 4;
constexpr int kAcc = 4;
constexpr int kGroups = 4;                                 // epilogue groups of 4 warps, taking alternate tiles
constexpr int kEpiWarps = 4 * kGroups;
constexpr int kThreads = (2 + kEpiWarps) * 32;
constexpr uint32_t kWBytes = kTaps * kCout * 64;           // 28672
constexpr uint32_t kSub0Rows = 11, kSub1Rows = 10;         // odd-row window (4 taps), even-row window (3 taps)
constexpr uint32_t kSub0Bytes = kSub0Rows * kTileW * 64;   // 11264
constexpr uint32_t kSub1Bytes = kSub1Rows * kTileW * 64;   // 10240
constexpr uint32_t kStageBytes = 22528;                    // both sub-windows, 1 KiB-aligned
constexpr uint32_t kTileBytes = 128 * 128;                 // conv tile staging: 128 pixels x 64 bf16
constexpr uint32_t kOffW = 0;
constexpr uint32_t kOffE = kWBytes;
constexpr uint32_t kOffS = kOffE + kStages * kStageBytes;
constexpr uint32_t kOffSB = kOffS + kGroups * kTileBytes;
constexpr uint32_t kOffBar = kOffSB + kCout * 8;
constexpr uint32_t kSmemBytes = kOffBar + 256 + 1024;      // + alignment slack

struct StemParams {
  int frames, ho, wo;          // conv output extent per frame
  int ph, pw;                  // pooled extent per frame
  int tiles_h, tiles_w;        // tiles per frame
  long long total_tiles;
  const float* scale;
  const float* bias;
  __nv_bfloat16* out;
  int out_pitch;
  int t_clip;                  // kt = 5 kernel: frames per clip (temporal taps never cross clips)
  long long total_runs;        // kt = 5 kernel: clips * tiles per frame
};

__device__ __forceinline__ void red_max_bf16x2_v4(void* gptr, uint4 v) {
  asm volatile("red.global.max.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(gptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint32_t hmax2_u32(uint32_t a, uint32_t b) {
  __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162*>(&a), y = *reinterpret_cast<__nv_bfloat162*>(&b);
  __nv_bfloat162 m = __hmax2(x, y);
  return *reinterpret_cast<uint32_t*>(&m);
}
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// One 8 x 16 conv tile of output frame `f`: TMEM accumulator -> BN + ReLU -> bf16 tile in shared memory -> 3x3/s2
// max-pool -> global memory (stores inside the tile, red.max on its seams).  Run by the 128 threads of one
// epilogue group (`et` = thread of the group, `quarter` = TMEM lane quarter of the warp).
__device__ __forceinline__ void epilogue_tile(const StemParams& p, const float2* sb, uint32_t s_tile, uint32_t tacc,
                                              uint64_t* full_bar, uint32_t full_phase, uint64_t* empty_bar, int quarter,
                                              int lane, int et, int bar_id, int f, int th, int tw) {
  const int row = quarter * 32 + lane;       // GEMM row = conv pixel (row >> 4, row & 15) of the tile
  mbar_wait(full_bar, full_phase);
  tc_fence_after();
  const uint32_t taddr = tacc + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t v[32];
    tmem_ld32(taddr + half * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {   // 16-byte chunk = 8 channels
      uint32_t o[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int cidx = half * 32 + ch * 8 + jj * 2;
        const float2 s0 = sb[cidx], s1 = sb[cidx + 1];
        const float x0 = fmaxf(fmaf(__uint_as_float(v[ch * 8 + jj * 2]), s0.x, s0.y), 0.f);
        const float x1 = fmaxf(fmaf(__uint_as_float(v[ch * 8 + jj * 2 + 1]), s1.x, s1.y), 0.f);
        o[jj] = pack_bf16x2(x0, x1);
      }
      const int chunk = half * 4 + ch;
      sts128(s_tile + row * 128 + ((chunk ^ (row & 7)) << 4), o[0], o[1], o[2], o[3]);
    }
  }
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(empty_bar);
  named_bar_sync(bar_id, 128);   // the whole conv tile is in shared memory
  // pooled rows 4 th + pl (pl = 0..4), columns 8 tw + ql (ql = 0..8), 8 chunks of 8 channels each
  for (int item = et; item < 5 * 9 * 8; item += 128) {
    const int k = item & 7, pq = item >> 3;
    const int pl = pq / 9, ql = pq - pl * 9;
    const int pg = th * 4 + pl, qg = tw * 8 + ql;
    if (pg >= p.ph || qg >= p.pw) continue;
    uint4 m = make_uint4(0, 0, 0, 0);   // post-ReLU values are >= 0
#pragma unroll
    for (int dr = 0; dr < 3; ++dr) {
      const int r = 2 * pl - 1 + dr;
      if (r < 0 || r >= kTileH) continue;
#pragma unroll
      for (int dc = 0; dc < 3; ++dc) {
        const int c = 2 * ql - 1 + dc;
        if (c < 0 || c >= kTileW) continue;
        const int i = r * kTileW + c;
        const uint4 x = lds128(s_tile + i * 128 + ((k ^ (i & 7)) << 4));
        m.x = hmax2_u32(m.x, x.x);
        m.y = hmax2_u32(m.y, x.y);
        m.z = hmax2_u32(m.z, x.z);
        m.w = hmax2_u32(m.w, x.w);
      }
    }
    __nv_bfloat16* dst = p.out + (((long long)f * p.ph + pg) * p.pw + qg) * p.out_pitch + k * 8;
    // seam outputs get contributions from two or four tiles; image-edge outputs (pl == 0 in the first tile
    // row, ql == 0 in the first tile column) are complete: the missing row / column is -inf padding
    const bool seam = (pl == 0 && th > 0) || pl == 4 || (ql == 0 && tw > 0) || ql == 8;
    if (seam)
      red_max_bf16x2_v4(dst, m);
    else
      *reinterpret_cast<uint4*>(dst) = m;
  }
  named_bar_sync(bar_id, 128);   // the group's staging tile may be overwritten
}

__global__ void __launch_bounds__(kThreads, 1)
stem_pool_kernel(const __grid_constant__ CUtensorMap map_odd, const __grid_constant__ CUtensorMap map_even,
                 const __grid_constant__ CUtensorMap map_w, const __grid_constant__ StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* full = bars;                  // [kStages] window landed
  uint64_t* empty = bars + kStages;       // [kStages] window consumed by the MMAs
  uint64_t* tmem_full = bars + 2 * kStages;        // [kAcc]
  uint64_t* tmem_empty = bars + 2 * kStages + kAcc;  // [kAcc]
  uint64_t* wbar = bars + 2 * kStages + 2 * kAcc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2 * kAcc + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < kAcc; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    mbar_init(wbar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kAcc * kCout);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_frame = p.tiles_h * p.tiles_w;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      tma_prefetch_desc(&map_odd);
      tma_prefetch_desc(&map_even);
      tma_prefetch_desc(&map_w);
      mbar_expect_tx(wbar, kWBytes);
      for (int kh = 0; kh < kTaps; ++kh) tma_load_2d(smem + kOffW + kh * (kCout * 64), &map_w, wbar, kh * 32, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int f = (int)(tile / tiles_per_frame);
        const int rem = (int)(tile - (long long)f * tiles_per_frame);
        const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
        const int r0 = th * kTileH, c0 = tw * kTileW;
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* e = smem + kOffE + stage * kStageBytes;
        mbar_expect_tx(&full[stage], kSub0Bytes + kSub1Bytes);
        // odd input rows 2(r0 - 2 + m) + 1, m = 0..10 (filter rows 0,2,4,6); even rows 2(r0 - 1 + m), m = 0..9 (1,3,5)
        tma_load_4d(e, &map_odd, &full[stage], 0, c0, r0 - 2, f);
        tma_load_4d(e + kSub0Bytes, &map_even, &full[stage], 0, c0, r0 - 1, f);
        if (++stage == kStages) stage = 0, phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, kCout);
      mbar_wait(wbar, 0);
      tc_fence_after();
      const uint32_t w_s = smem_u32(smem + kOffW);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t e_s = smem_u32(smem + kOffE + stage * kStageBytes);
        const uint32_t d = tmem_base + acc * kCout;
#pragma unroll
        for (int kh = 0; kh < kTaps; ++kh) {
          const uint32_t a_s = e_s + ((kh & 1) ? kSub0Bytes : 0u) + (uint32_t)(kh >> 1) * (kTileW * 64);
          const uint32_t b_s = w_s + kh * (kCout * 64);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            umma_bf16(d, umma_smem_desc(a_s + ks * 32, 64), umma_smem_desc(b_s + ks * 32, 64), idesc, (kh | ks) != 0);
        }
        umma_commit(&empty[stage]);
        umma_commit(&tmem_full[acc]);
        if (++stage == kStages) stage = 0, phase ^= 1;
        if (++acc == kAcc) acc = 0, acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: BN + ReLU + 3x3/s2 max-pool
    // kGroups groups of four warps (one per TMEM lane quarter) take alternate tiles: one warp per scheduler cannot
    // hide the latency of the ~650 dependent-ish instructions a tile costs each thread (measured: 4260 clk per tile
    // with one group, the MMA side needs ~1200)
    const int ew = warp - 2;                   // 0 .. kEpiWarps-1
    const int group = ew >> 2;
    const int et = (ew & 3) * 32 + lane;       // thread of the group, 0..127
    const int quarter = warp & 3;              // TMEM lane quarter this warp may read
    float2* sb = reinterpret_cast<float2*>(smem + kOffSB);
    if (group == 0 && et < kCout) sb[et] = make_float2(p.scale[et], p.bias[et]);
    named_bar_sync(1, kEpiWarps * 32);
    const uint32_t s_tile = smem_u32(smem + kOffS) + group * kTileBytes;
    const int bar_id = 2 + group;
    long long j = group;                       // CTA-local tile counter: tile j uses accumulator j % kAcc
    for (long long tile = blockIdx.x + (long long)group * gridDim.x; tile < p.total_tiles;
         tile += (long long)kGroups * gridDim.x, j += kGroups) {
      const int acc = (int)(j % kAcc);
      const uint32_t acc_phase = (uint32_t)((j / kAcc) & 1);
      const int f = (int)(tile / tiles_per_frame);
      const int rem = (int)(tile - (long long)f * tiles_per_frame);
      const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
      epilogue_tile(p, sb, s_tile, tmem_base + acc * kCout, &tmem_full[acc], acc_phase, &tmem_empty[acc], quarter, lane, et,
                    bar_id, f, th, tw);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, kAcc * kCout);
  }
}

// ---------------------------------------------------------------------------------------------- kt = 5 variant
// [5,7,7] stems (I3D: stem_helper.py:157-178 with temporal kernel 5, pad 2).  Same tiles and windows; the five
// temporal taps are SCATTERED: the window of input frame f is multiplied by W[kt] into the accumulator of output
// frame f + 2 - kt, so a window is fetched once and lives in shared memory for one frame only (a gather would need
// five windows plus 140 KB of weights).  Output frame t owns TMEM slot t % 8 (8 slots x 64 columns = all of
// TMEM); slots of consecutive output frames are adjacent, so two taps run as ONE N = 128 MMA over the stacked
// weight blocks [W2; W1] and [W4; W3] (64 clk instead of 2 x the 60-clk floor of N = 64).  kt = 0 always opens a
// fresh accumulator and is issued alone (the accumulate flag is per instruction).  All 140 KB of weights resident.
constexpr int kStages5 = 2, kGroups5 = 2, kSlots5 = 8;
constexpr int kThreads5 = (2 + 4 * kGroups5) * 32;
constexpr uint32_t kWBytes5 = 5 * kWBytes;                 // 143360: block kt0 | pair [W2;W1] | pair [W4;W3]
constexpr uint32_t kOffE5 = kWBytes5;
constexpr uint32_t kOffS5 = kOffE5 + kStages5 * kStageBytes;
constexpr uint32_t kOffSB5 = kOffS5 + kGroups5 * kTileBytes;
constexpr uint32_t kOffBar5 = kOffSB5 + kCout * 8;
constexpr uint32_t kSmemBytes5 = kOffBar5 + 256 + 1024;

__global__ void __launch_bounds__(kThreads5, 1)
stem5_pool_kernel(const __grid_constant__ CUtensorMap map_odd, const __grid_constant__ CUtensorMap map_even,
                  const __grid_constant__ CUtensorMap map_w, const __grid_constant__ StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar5);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages5;
  uint64_t* tmem_full = bars + 2 * kStages5;
  uint64_t* tmem_empty = bars + 2 * kStages5 + kSlots5;
  uint64_t* wbar = bars + 2 * kStages5 + 2 * kSlots5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages5 + 2 * kSlots5 + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages5; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < kSlots5; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    mbar_init(wbar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kSlots5 * kCout);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_frame = p.tiles_h * p.tiles_w;
  const int T = p.t_clip;
  const int uses_per_run = T > kSlots5 ? T / kSlots5 : 1;   // times a slot is used by one run (T <= 8 or T % 8 == 0)

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&map_odd);
      tma_prefetch_desc(&map_even);
      tma_prefetch_desc(&map_w);
      mbar_expect_tx(wbar, kWBytes5);
      // global rows: [W0; W2; W1; W4; W3] x 64.  Shared memory: kt0 block [kh][64 rows], then per pair [kh][128 rows]
      for (int kh = 0; kh < kTaps; ++kh) {
        tma_load_2d(smem + kh * 4096, &map_w, wbar, kh * 32, 0);
        for (int pr = 0; pr < 2; ++pr)
          for (int half = 0; half < 2; ++half)
            tma_load_2d(smem + kWBytes + pr * (2 * kWBytes) + kh * 8192 + half * 4096, &map_w, wbar, kh * 32,
                        64 + pr * 128 + half * 64);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (long long run = blockIdx.x; run < p.total_runs; run += gridDim.x) {
        const int n = (int)(run / tiles_per_frame);
        const int rem = (int)(run - (long long)n * tiles_per_frame);
        const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
        const int r0 = th * kTileH, c0 = tw * kTileW;
        for (int f = 0; f < T; ++f) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* e = smem + kOffE5 + stage * kStageBytes;
          mbar_expect_tx(&full[stage], kSub0Bytes + kSub1Bytes);
          tma_load_4d(e, &map_odd, &full[stage], 0, c0, r0 - 2, n * T + f);
          tma_load_4d(e + kSub0Bytes, &map_even, &full[stage], 0, c0, r0 - 1, n * T + f);
          if (++stage == kStages5) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc64 = umma_idesc_bf16(128, 64), idesc128 = umma_idesc_bf16(128, 128);
      mbar_wait(wbar, 0);
      tc_fence_after();
      const uint32_t w0_s = smem_u32(smem);                       // kt = 0
      const uint32_t w21_s = w0_s + kWBytes;                      // [W2; W1]
      const uint32_t w43_s = w21_s + 2 * kWBytes;                 // [W4; W3]
      int stage = 0;
      uint32_t phase = 0;
      long long run_local = 0;
      // one tap set = 7 filter rows x 2 K steps over the window at e_s; b_s / b_kh_stride select the weight block
      auto tap_set = [&](uint32_t e_s, uint32_t b_s, uint32_t b_kh_stride, uint32_t idesc, int slot, bool fresh) {
        const uint32_t d = tmem_base + slot * kCout;
#pragma unroll
        for (int kh = 0; kh < kTaps; ++kh) {
          const uint32_t a_s = e_s + ((kh & 1) ? kSub0Bytes : 0u) + (uint32_t)(kh >> 1) * (kTileW * 64);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            umma_bf16(d, umma_smem_desc(a_s + ks * 32, 64), umma_smem_desc(b_s + kh * b_kh_stride + ks * 32, 64), idesc,
                      !(fresh && kh == 0 && ks == 0));
        }
      };
      auto wait_slot_free = [&](int t) {   // output frame t of this run opens its accumulator
        const long long use = run_local * uses_per_run + (t >> 3);
        mbar_wait(&tmem_empty[t & 7], (uint32_t)(use & 1) ^ 1);
      };
      for (long long run = blockIdx.x; run < p.total_runs; run += gridDim.x, ++run_local) {
        for (int f = 0; f < T; ++f) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t e_s = smem_u32(smem + kOffE5 + stage * kStageBytes);
          // kt = 0 -> output frame f + 2: always the first contribution
          if (f + 2 < T) {
            wait_slot_free(f + 2);
            tc_fence_after();
            tap_set(e_s, w0_s, 4096, idesc64, (f + 2) & 7, true);
          }
          // kt = 2 -> frame f, kt = 1 -> frame f + 1 (first contributions only for the run's first input frame)
          {
            const bool fresh = f == 0, v1 = f + 1 < T;
            if (fresh) {
              wait_slot_free(0);
              if (v1) wait_slot_free(1);
              tc_fence_after();
            }
            if (v1 && (f & 7) != 7) {
              tap_set(e_s, w21_s, 8192, idesc128, f & 7, fresh);
            } else {
              tap_set(e_s, w21_s, 8192, idesc64, f & 7, fresh);
              if (v1) tap_set(e_s, w21_s + 4096, 8192, idesc64, (f + 1) & 7, fresh);
            }
          }
          // kt = 4 -> frame f - 2, kt = 3 -> frame f - 1 (never first contributions)
          if (f >= 2) {
            if (((f - 2) & 7) != 7) {
              tap_set(e_s, w43_s, 8192, idesc128, (f - 2) & 7, false);
            } else {
              tap_set(e_s, w43_s, 8192, idesc64, (f - 2) & 7, false);
              tap_set(e_s, w43_s + 4096, 8192, idesc64, (f - 1) & 7, false);
            }
          } else if (f == 1) {
            tap_set(e_s, w43_s + 4096, 8192, idesc64, 0, false);
          }
          umma_commit(&empty[stage]);
          // output frames whose last contribution this was: f - 2, and at the end of the run the last two
          if (f >= 2) umma_commit(&tmem_full[(f - 2) & 7]);
          if (f == T - 1) {
            umma_commit(&tmem_full[(T - 2) & 7]);
            umma_commit(&tmem_full[(T - 1) & 7]);
          }
          if (++stage == kStages5) stage = 0, phase ^= 1;
        }
      }
    }
  } else {
    const int ew = warp - 2;
    const int group = ew >> 2;
    const int et = (ew & 3) * 32 + lane;
    const int quarter = warp & 3;
    float2* sb = reinterpret_cast<float2*>(smem + kOffSB5);
    if (group == 0 && et < kCout) sb[et] = make_float2(p.scale[et], p.bias[et]);
    named_bar_sync(1, 4 * kGroups5 * 32);
    const uint32_t s_tile = smem_u32(smem + kOffS5) + group * kTileBytes;
    const int bar_id = 2 + group;
    // output-frame tiles in completion order: j = run_local * T + t; the groups take alternate tiles
    for (long long j = group;; j += kGroups5) {
      const long long run_local = j / T;
      const int t = (int)(j - run_local * T);
      const long long run = blockIdx.x + run_local * gridDim.x;
      if (run >= p.total_runs) break;
      const int n = (int)(run / tiles_per_frame);
      const int rem = (int)(run - (long long)n * tiles_per_frame);
      const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
      const long long use = run_local * uses_per_run + (t >> 3);
      epilogue_tile(p, sb, s_tile, tmem_base + (t & 7) * kCout, &tmem_full[t & 7], (uint32_t)(use & 1), &tmem_empty[t & 7],
                    quarter, lane, et, bar_id, n * T + t, th, tw);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, kSlots5 * kCout);
  }
}

// Zero the first `chunks` 16-byte chunks of every SEAM pixel of the pooled [frames, ph, pw, pitch_chunks * 16 B] tensor:
// the pooled pixels that epilogue_tile combines from two or four conv tiles with red.max (pooled row a multiple of 4
// and >= 4, or pooled column a multiple of 8 and >= 8 - the same predicate as `seam` there).  Every other pixel is
// written exactly once by a plain store, so it needs no initial value (a third of the tensor is zeroed, not all of it).
// One warp per pooled row (32-bit index arithmetic, one division per row): a seam row is zeroed whole, any other row
// only at its seam columns.
__global__ void zero_seams_kernel(uint4* out, long long rows, int ph, int pw, int pitch_chunks) {
  constexpr int kChunks = kCout / 8;
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps) {
    const int pg = (int)(row % ph);
    uint4* base = out + row * pw * pitch_chunks;
    if (pg >= 4 && (pg & 3) == 0) {
      for (int i = lane; i < pw * kChunks; i += 32) base[(i / kChunks) * pitch_chunks + (i % kChunks)] = make_uint4(0, 0, 0, 0);
    } else {
      const int nseam = (pw - 1) >> 3;   // columns 8, 16, ... below pw
      for (int i = lane; i < nseam * kChunks; i += 32)
        base[((i / kChunks) + 1) * 8 * pitch_chunks + (i % kChunks)] = make_uint4(0, 0, 0, 0);
    }
  }
}

}  // namespace
}  // namespace vsb

using namespace vsb;

struct vsb_stem_pool_plan {
  vsb_stem_pool_desc desc;
  CUtensorMap map_odd, map_even, map_w;
  StemParams params;
  unsigned grid;
};

static int sp_encode(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                     const cuuint32_t* box, const char* what) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides,
                              box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(%s) failed (CUresult %d)", what, (int)r);
    return VSB_ERR_CUDA;
  }
  return VSB_OK;
}

extern "C" int vsb_stem_pool_plan_create(const vsb_stem_pool_desc* d, vsb_stem_pool_plan** out) {
  VSB_CHECK_ARG(d && out, "null argument");
  *out = nullptr;
  VSB_CHECK_ARG(d->in && d->wgt && d->scale && d->bias && d->out, "null tensor");
  VSB_CHECK_ARG(d->frames > 0 && d->h > 0 && d->w > 0, "bad extent");
  VSB_CHECK_ARG(d->kt == 1 || d->kt == 5, "fused stem: temporal kernel must be 1 or 5 (got %d)", d->kt);
  VSB_CHECK_ARG(d->kt == 1 || (d->t >= 3 && d->frames % d->t == 0 && (d->t <= 8 || d->t % 8 == 0)),
                "fused [5,7,7] stem: frames per clip (%d) must divide the frame count and be 3..8 or a multiple of 8", d->t);
  VSB_CHECK_ARG(d->h % 32 == 0 && d->w % 32 == 0, "fused stem: frame height and width must be multiples of 32 (got %d x %d)", d->h, d->w);
  VSB_CHECK_ARG(d->w_buf >= d->w + 8, "fused stem: input rows need 3 zero pixels before and 5 after the image (w_buf >= w + 8)");
  VSB_CHECK_ARG(d->out_pitch >= kCout && d->out_pitch % 8 == 0, "out_pitch must be a multiple of 8 and >= 64");
  VSB_CHECK_ARG(((uintptr_t)d->in | (uintptr_t)d->wgt | (uintptr_t)d->out) % 16 == 0, "tensors must be 16-byte aligned");
  int rc = load_driver_entry_points();
  if (rc != VSB_OK) return rc;
  vsb_stem_pool_plan* plan = new (std::nothrow) vsb_stem_pool_plan();
  VSB_CHECK_ARG(plan, "out of host memory");
  plan->desc = *d;
  const int ho = d->h / 2, wo = d->w / 2;
  const unsigned long long row_bytes = (unsigned long long)d->w_buf * 8, frame_bytes = row_bytes * d->h;
  const unsigned long long slots = ((unsigned long long)d->w_buf * 4 - 32) / 8 + 1;  // 64-byte slots at 16-byte steps
  {
    // overlapping view: [32 channels-of-8-pixels][slot, 16 B apart][row of one parity][frame]
    cuuint64_t dims[4] = {32, slots, (cuuint64_t)(d->h / 2), (cuuint64_t)d->frames};
    cuuint64_t strides[3] = {16, 2 * row_bytes, frame_bytes};
    cuuint32_t box0[4] = {32, (cuuint32_t)kTileW, kSub0Rows, 1}, box1[4] = {32, (cuuint32_t)kTileW, kSub1Rows, 1};
    rc = sp_encode(&plan->map_odd, (const char*)d->in + row_bytes, 4, dims, strides, box0, "stem input, odd rows");
    if (rc == VSB_OK) rc = sp_encode(&plan->map_even, d->in, 4, dims, strides, box1, "stem input, even rows");
  }
  if (rc == VSB_OK) {
    cuuint64_t dims[2] = {(cuuint64_t)kTaps * 32, (cuuint64_t)kCout * d->kt};
    cuuint64_t strides[1] = {(cuuint64_t)kTaps * 32 * 2};
    cuuint32_t box[2] = {32, (cuuint32_t)kCout};
    rc = sp_encode(&plan->map_w, d->wgt, 2, dims, strides, box, "stem weights");
  }
  if (rc != VSB_OK) {
    delete plan;
    return rc;
  }
  StemParams& p = plan->params;
  p.frames = d->frames, p.ho = ho, p.wo = wo;
  p.ph = ho / 2, p.pw = wo / 2;
  p.tiles_h = ho / kTileH, p.tiles_w = wo / kTileW;
  p.total_tiles = (long long)d->frames * p.tiles_h * p.tiles_w;
  p.scale = d->scale, p.bias = d->bias;
  p.out = (__nv_bfloat16*)d->out;
  p.out_pitch = d->out_pitch;
  p.t_clip = d->kt == 5 ? d->t : 1;
  p.total_runs = (long long)(d->frames / p.t_clip) * p.tiles_h * p.tiles_w;
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(stem5_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes5);
  if (e != cudaSuccess) {
    delete plan;
    set_error("fused stem plan: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return VSB_ERR_CUDA;
  }
  const long long units = d->kt == 5 ? p.total_runs : p.total_tiles;
  plan->grid = (unsigned)(units < sms ? units : sms);
  *out = plan;
  return VSB_OK;
}

extern "C" int vsb_stem_pool_run(const vsb_stem_pool_plan* plan, void* stream) {
  VSB_CHECK_ARG(plan, "null plan");
  cudaStream_t s = (cudaStream_t)stream;
  const StemParams& p = plan->params;
  zero_seams_kernel<<<1184, 256, 0, s>>>(reinterpret_cast<uint4*>(p.out), (long long)p.frames * p.ph, p.ph, p.pw, p.out_pitch / 8);
  VSB_CHECK_LAUNCH("zero_seams_kernel");
  if (plan->desc.kt == 5)
    stem5_pool_kernel<<<plan->grid, kThreads5, kSmemBytes5, s>>>(plan->map_odd, plan->map_even, plan->map_w, p);
  else
    stem_pool_kernel<<<plan->grid, kThreads, kSmemBytes, s>>>(plan->map_odd, plan->map_even, plan->map_w, p);
  VSB_CHECK_LAUNCH("stem_pool_kernel");
  return VSB_OK;
}

extern "C" void vsb_stem_pool_plan_destroy(vsb_stem_pool_plan* plan) { delete plan; }

extern "C" int vsb_stem_pool_plan_desc(const vsb_stem_pool_plan* plan, vsb_stem_pool_desc* out) {
  VSB_CHECK_ARG(plan && out, "null argument");
  *out = plan->desc;
  return VSB_OK;
}
