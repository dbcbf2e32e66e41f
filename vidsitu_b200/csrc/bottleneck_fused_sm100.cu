// Fused identity bottleneck block: out = relu(x + BN_c(c(relu(BN_b(b(relu(BN_a(a(x)))))))))  in ONE launch.
//
// Reference ops replaced: BottleneckTransform a -> b -> c (SlowFast/slowfast/models/resnet_helper.py:225-240:
// a = Conv3d [kt,1,1] + BN + ReLU, b = Conv3d [1,3,3] pad 1 + BN + ReLU, c = Conv3d 1x1x1 + BN) and the identity
// ResBlock around it (resnet_helper.py:352-358: relu(x + branch2(x))), for blocks without a projection shortcut
// and with unit strides.  Layer by layer this is three launches and 16 d bytes of HBM traffic per pixel
// (d = bottleneck width in bf16 bytes: x 4d in, a d out / d in, b d out / d in, residual 4d in, out 4d); here
// x is read (once from HBM, the temporal taps and the residual from L2) and out is written: 8 d.
//
// Geometry: "flat raster walk".  A clip is laid out as one flat array of SLOTS (slot = one GEMM row = one pixel,
// or one pixel group when the host restates the convs on groups): T frames x FP flat rows x RP slots, RP a power
// of two > W, FP > H, so every image row is followed by >= 1 zero slot and every frame by >= 1 zero row -- those
// zeros ARE the padding of the 3x3 conv.  M-tile m = flat slots [128 m, 128 m + 128).  For every tile of a walk
//   a-phase : a[m]   = relu(BN(x[m] . Wa))            x chunks by TMA (kt frames x channel chunks), acc in TMEM,
//                                                      epilogue -> bf16 -> shared-memory RING of a-tiles (zeros
//                                                      written at the padding slots)
//   b-phase : b[m]   = relu(BN(sum_taps a[m + dy RP + dx] . Wb[tap]))   every tap is a 128-row UMMA operand whose
//                                                      descriptor start address is shifted by whole slots inside
//                                                      the ring (hardware swizzle is a function of the absolute
//                                                      address: conv_win_sm100.cu), acc in TMEM, epilogue -> bf16
//                                                      -> shared-memory P tile
//   c-phase : out[m] = relu(BN(b[m] . Wc) + x[m])     acc in TMEM, residual by TMA into the staging slab, TMA store
// The ring holds 4 a-tiles plus a MIRROR of ring tile 0 behind tile 3, so that the 128-slot operand of any tap,
// which starts in tile (m-1) or m and runs into the next one, is contiguous in shared memory.  All weights are
// resident.  Warp roles (18 warps): 0-3 a-epilogue, 4-7 b-epilogue, 8-15 c-epilogue (two groups of four, alternate
// tiles), 16 TMA producer, 17 MMA issuer.  The MMA issuer runs the three phases software-pipelined over the
// CTA's whole tile sequence: step g issues a(g), b(centre g-2) and c(centre g-3); tcgen05.mma executes in issue
// order and a commit covers everything issued before it, which is what makes the ring reuse safe without extra
// barriers (a(g)'s epilogue overwrites the ring tile that b(centre g-5) was the last to read).
#include <cuda.h>

#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <new>

#include "bottleneck_thin.h"
#include "common.h"
#include "conv_plan.h"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace vsb {

namespace {
constexpr int kRing = 4;  // a-tiles in the ring (power of two) + 1 mirror tile
constexpr int kMaxBoxes = 16;
constexpr int kMaxStages = 24;
constexpr int kBEpiWarp0 = 4, kCEpiWarp0 = 8, kProducerWarp = 16, kMmaWarp = 17, kMmaAWarp = 18, kProducer2Warp0 = 19;
constexpr int kXProducers = 2;  // warps that copy x chunks in cp.async mode: kProducerWarp and one more (20 warps: 96 registers)
constexpr int kThreads = 20 * 32;
constexpr int kCEpiWarps = 8;
constexpr int kMaxEpiBufs = 3;
#ifndef FB_PRETEST
#define FB_PRETEST 0
#endif
}  // namespace

struct FusedParams {
  // flat slot space
  int W, H, T, n_clips;
  int RP, rp_shift, FP;
  int rows_per_tile;            // 128 / RP
  int box_rows, boxes_per_tile; // x TMA boxes: box_rows flat rows each
  uint32_t box_bytes;
  int tiles_per_clip;
  int walk_len;                 // longest contiguous tile run of one clip a CTA walks without re-priming the halo
  int tiles_per_cta, total_tiles;
  int row_bias;                 // multiple of FP added before the (t, y) split so that the dividend is >= 0
  // channels
  int kt, kc, cchunks, chunks_per_tile;
  int ksteps_x;                 // kc / 16
  int d16, ksteps_d;            // bottleneck width (multiple of 16), d16 / 16
  int cout;
  uint32_t x_row_bytes, a_row_bytes, chunk_bytes;
  int stages;                   // x ring: `stages` slots of `cps` chunks each (one full / empty barrier pair per slot)
  int cps, slots_per_tile;
  int pf_dist;                  // L2 prefetch distance of the x loads, in tiles (0 = off)
  int x_im2col;                 // x chunks by ONE im2col-mode TMA load each (else tiled boxes of box_rows flat rows)
  int x_cpasync;                // x chunks by per-thread cp.async copies (two producer warps), not by the TMA unit
  const uint8_t* x;             // block input (cp.async mode)
  int x_pitch;
  uint32_t stage_bytes;
  // shared-memory layout (bytes from the 1 KiB-aligned base)
  uint32_t off_wa, off_wb, off_wc, off_ring, off_p, off_epi, off_bar, off_tab;
  uint32_t wa_block_bytes, wb_block_bytes, wc_bytes, tile_bytes;
  // TMEM columns
  uint32_t tmem_cols, tm_a, tm_b, tm_c, tm_ab_stride, tm_c_stride;
  int c_bufs;                   // c accumulators per c-epilogue group (1 or 2): 2 * c_bufs in all
  uint32_t idesc_a, idesc_b, idesc_c;
  // c epilogue
  int epi_n, epi_chunks, epi_bufs, bw, bh;
  const float *sa, *ba, *sb, *bb, *sc, *bc;
  long long* dbg;  // role timeline counters (VSB_FUSED_DEBUG), else null
};

// Enumerates the a-tiles of one CTA.  The CTA owns the contiguous range [cur0, end) of the launch's flat tile index
// (clip-major); the range is cut at clip boundaries (and every walk_len tiles) into walks, and the a-tiles j = 0 .. L+1
// of a walk are the flat tiles m0-1 .. m0+L of clip n (the first and last one only feed the 3x3 halo of their
// neighbours).
struct ATileIter {
  int cur, end, j, L, n, m0;
  __device__ __forceinline__ void load_walk(const FusedParams& p) {
    if (cur < end) {
      n = cur / p.tiles_per_clip;
      m0 = cur - n * p.tiles_per_clip;
      L = end - cur;
      if (L > p.tiles_per_clip - m0) L = p.tiles_per_clip - m0;
      if (L > p.walk_len) L = p.walk_len;
    }
  }
  __device__ __forceinline__ void init(const FusedParams& p) {
    cur = blockIdx.x * p.tiles_per_cta;
    end = cur + p.tiles_per_cta < p.total_tiles ? cur + p.tiles_per_cta : p.total_tiles;
    j = 0;
    load_walk(p);
  }
  __device__ __forceinline__ bool valid(const FusedParams&) const { return cur < end; }
  __device__ __forceinline__ void next(const FusedParams& p) {
    if (++j == L + 2) {
      j = 0;
      cur += L;
      load_walk(p);
    }
  }
  __device__ __forceinline__ int m() const { return m0 - 1 + j; }
};

// 16 accumulator columns of one row -> bf16 -> 32 bytes of a swizzled K-major operand row, written to dst_s and
// (when dst2_s != 0) to its mirror; `keep` = 0 writes zeros (padding slots).  ReLU always.
__device__ __forceinline__ void fb_convert16(const uint32_t* v, uint32_t sb_s, uint32_t dst_s, uint32_t dst2_s,
                                             uint32_t off0, uint32_t swz_mask, uint32_t keep) {
  const uint32_t s0 = off0 ^ (((off0 >> 7) & swz_mask) << 4);
  const uint32_t off1 = off0 + 16;
  const uint32_t s1 = off1 ^ (((off1 >> 7) & swz_mask) << 4);
  uint32_t o[8];
#pragma unroll
  for (int qq = 0; qq < 8; ++qq) {
    const float4 p2 = lds128f(sb_s + 16 * qq);  // (scale, bias) of two columns
    const float x0 = fmaxf(fmaf(__uint_as_float(v[2 * qq]), p2.x, p2.y), 0.f);
    const float x1 = fmaxf(fmaf(__uint_as_float(v[2 * qq + 1]), p2.z, p2.w), 0.f);
    o[qq] = pack_bf16x2(x0, x1) & keep;
  }
  sts128(dst_s + s0, o[0], o[1], o[2], o[3]);
  sts128(dst_s + s1, o[4], o[5], o[6], o[7]);
  if (dst2_s) {
    sts128(dst2_s + s0, o[0], o[1], o[2], o[3]);
    sts128(dst2_s + s1, o[4], o[5], o[6], o[7]);
  }
}

// One warp converts its 32 rows x ncols accumulator block into rows [row0, row0 + 32) of an operand tile.
__device__ __forceinline__ void fb_convert_rows(uint32_t taddr, int ncols, uint32_t dst_s, uint32_t dst2_s,
                                                uint32_t row_bytes, uint32_t swz_mask, int row, uint32_t sb_s,
                                                uint32_t keep) {
  if (ncols >= 32) {
    for (int j0 = 0; j0 < ncols; j0 += 32) {
      uint32_t v[32];
      tmem_ld32(taddr + j0, v);
      tmem_ld_wait();
      const uint32_t off0 = (uint32_t)row * row_bytes + j0 * 2;
      fb_convert16(v, sb_s + j0 * 8, dst_s, dst2_s, off0, swz_mask, keep);
      fb_convert16(v + 16, sb_s + (j0 + 16) * 8, dst_s, dst2_s, off0 + 32, swz_mask, keep);
    }
  } else {
    uint32_t v[16];
    tmem_ld16(taddr, v);
    tmem_ld_wait();
    fb_convert16(v, sb_s, dst_s, dst2_s, (uint32_t)row * row_bytes, swz_mask, keep);
  }
}

// MMA issue: tight table-driven loops (the issuing warp is instruction-fetch / issue bound, not tensor-pipe bound: a
// fully unrolled issue sequence of a whole tile does not fit the instruction cache next to the four epilogue roles).
// tab[i] = (A offset, B offset) of MMA i in 16-byte descriptor units, relative to the operand bases.
__device__ __forceinline__ void fb_issue(uint32_t tm_d, uint64_t a_hi, uint64_t b_hi, uint32_t a_base, uint32_t b_base,
                                         const uint2* tab, int n, uint32_t idesc, uint32_t acc_first) {
  {
    const uint2 e = tab[0];
    umma_bf16(tm_d, a_hi | (uint64_t)(a_base + e.x), b_hi | (uint64_t)(b_base + e.y), idesc, acc_first);
  }
#pragma unroll 2
  for (int i = 1; i < n; ++i) {
    const uint2 e = tab[i];
    umma_bf16(tm_d, a_hi | (uint64_t)(a_base + e.x), b_hi | (uint64_t)(b_base + e.y), idesc, 1u);
  }
}

// Optional role timeline (VSB_FUSED_DEBUG=1 at plan time): cycles per role, summed over CTAs.
//  0 producer: wait free x slot          1 MMA: wait a-acc drained     2 MMA: wait x chunk        3 MMA: wait a-tile in ring
//  4 MMA: wait b-acc drained             5 MMA: wait P tile            6 MMA: wait c-acc drained  7 a-epi: wait a_full
//  8 a-epi: work                         9 b-epi: wait b_full         10 b-epi: work             11 c-epi (warp 8): wait c_full
// 12 c-epi: wait slab / residual        13 c-epi: work                14 c-epi: wait store read  15 CTA total (warp 0)
// 16 a-tiles                            17 c-tiles of warp 8
#define FB_T(idx, stmt)                              \
  do {                                               \
    if (kDbg) {                                      \
      const long long _t0 = clock64();               \
      stmt;                                          \
      dbg_acc[idx] += (uint32_t)(clock64() - _t0);   \
    } else {                                         \
      stmt;                                          \
    }                                                \
  } while (0)

template <bool kDbg>
__global__ void __launch_bounds__(kThreads, 1)
bottleneck_fused_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_wa,
                        const __grid_constant__ CUtensorMap map_wb, const __grid_constant__ CUtensorMap map_wc,
                        const __grid_constant__ CUtensorMap map_res, const __grid_constant__ CUtensorMap map_out,
                        const __grid_constant__ FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  uint64_t* x_full = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint64_t* x_empty = x_full + kMaxStages;
  uint64_t* a_full = x_empty + kMaxStages;
  uint64_t* a_done = a_full + 2;      // a-epilogue -> a-issuer: accumulator drained
  uint64_t* a_ready = a_done + 2;     // a-epilogue -> b/c-issuer: ring tile written (one per ring slot)
  uint64_t* ring_free = a_ready + kRing;  // b/c-issuer (commit) -> a-issuer: every MMA issued through step s has completed
  uint64_t* b_full = ring_free + kRing;
  uint64_t* b_done = b_full + 2;
  uint64_t* c_full = b_done + 2;      // [group * 2 + buffer]
  uint64_t* c_empty = c_full + 4;
  uint64_t* epi_ready = c_empty + 4;
  uint64_t* w_bar = epi_ready + kCEpiWarps * kMaxEpiBufs;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint32_t dbg_acc[24] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const long long dbg_start = kDbg ? clock64() : 0;

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_wa);
    tma_prefetch_desc(&map_wb);
    tma_prefetch_desc(&map_wc);
    tma_prefetch_desc(&map_res);
    tma_prefetch_desc(&map_out);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&x_full[s], p.x_cpasync ? 32 * kXProducers : 1);
      mbar_init(&x_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_done[i], 4);
      mbar_init(&a_ready[i], 4);
      mbar_init(&a_ready[i + 2], 4);
      mbar_init(&ring_free[i], 1);
      mbar_init(&ring_free[i + 2], 1);
      mbar_init(&b_full[i], 1);
      mbar_init(&b_done[i], 4);
      mbar_init(&c_full[i], 1);
      mbar_init(&c_full[i + 2], 1);
      mbar_init(&c_empty[i], 4);
      mbar_init(&c_empty[i + 2], 4);
    }
    for (int i = 0; i < kCEpiWarps * kMaxEpiBufs; ++i) mbar_init(&epi_ready[i], 1);
    mbar_init(w_bar, 1);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  // (scale, bias) pairs of the three convs
  float2* sb_a = reinterpret_cast<float2*>(smem + p.off_tab);
  float2* sb_b = sb_a + p.d16;
  float2* sb_c = sb_b + p.d16;
  for (int c = threadIdx.x; c < p.d16; c += blockDim.x) {
    sb_a[c] = make_float2(p.sa[c], p.ba[c]);
    sb_b[c] = make_float2(p.sb[c], p.bb[c]);
  }
  for (int c = threadIdx.x; c < p.cout; c += blockDim.x) sb_c[c] = make_float2(p.sc[c], p.bc[c]);
  // MMA offset tables (fb_issue): a-phase entries of one ring slot, the 9 taps of the b-phase, the c-phase
  uint2* tab_a = reinterpret_cast<uint2*>(sb_c + p.cout);
  uint2* tab_b = tab_a + p.cps * p.ksteps_x;
  uint2* tab_c = tab_b + 9 * p.ksteps_d;
  for (int i = threadIdx.x; i < p.cps * p.ksteps_x; i += blockDim.x) {
    const uint32_t c = i / p.ksteps_x, k = i % p.ksteps_x;
    tab_a[i] = make_uint2(c * (p.chunk_bytes >> 4) + 2 * k, c * (p.wa_block_bytes >> 4) + 2 * k);
  }
  for (int i = threadIdx.x; i < 9 * p.ksteps_d; i += blockDim.x) {
    const int tap = i / p.ksteps_d, k = i % p.ksteps_d;
    const int o = (tap / 3 - 1) * p.RP + (tap % 3 - 1);  // start slot relative to the centre tile (taps 0-3: < 0)
    tab_b[i] = make_uint2((((uint32_t)(o < 0 ? 128 + o : o) * p.a_row_bytes) >> 4) + 2 * k,
                          (uint32_t)tap * (p.wb_block_bytes >> 4) + 2 * k);
  }
  for (int i = threadIdx.x; i < p.ksteps_d; i += blockDim.x) tab_c[i] = make_uint2(2 * i, 2 * i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kProducerWarp || warp >= kProducer2Warp0) {
    if (warp == kProducerWarp && lane == 0) {
      // resident weights: Wa as chunks_per_tile blocks [d16 rows x kc], Wb as 9 tap blocks [d16 x d16], Wc [cout x d16]
      // (Wb / Wc boxes are a_row_bytes wide: when that is more than d16 channels the extra K columns are never read)
      mbar_expect_tx(w_bar, (uint32_t)p.chunks_per_tile * (uint32_t)p.d16 * p.x_row_bytes +
                                9u * (uint32_t)p.d16 * p.a_row_bytes + (uint32_t)p.cout * p.a_row_bytes);
      for (int q = 0; q < p.chunks_per_tile; ++q)
        tma_load_2d(smem + p.off_wa + q * p.wa_block_bytes, &map_wa, w_bar, q * p.kc, 0);
      for (int tap = 0; tap < 9; ++tap)
        tma_load_2d(smem + p.off_wb + tap * p.wb_block_bytes, &map_wb, w_bar, tap * p.d16, 0);
      tma_load_2d(smem + p.off_wc, &map_wc, w_bar, 0, 0);
    }
    __syncwarp();
    const int S = p.stages, cch = p.cchunks, kc = p.kc, cps = p.cps, cpt = p.chunks_per_tile;
    if (p.x_cpasync) {
      // ---------------------------------------------------------- x producer, cp.async mode (two whole warps)
      // The TMA unit is the bound of this kernel when it also has to move x (measured: ~25 B/clk/SM for x + residual
      // + output); here x goes through the LSU instead: every lane copies 16-byte pieces of chunk rows (coalesced:
      // 32 lanes = 32 / pieces-per-row consecutive slots), swizzling the destination as the TMA unit would and
      // zero-filling the padding slots (src-size 0).  Chunk q of a tile is copied by producer q % 2.
      const int pidx = warp == kProducerWarp ? 0 : warp - kProducer2Warp0 + 1;
      const uint32_t rb = p.x_row_bytes, swz_mask = rb == 128 ? 7u : (rb == 64 ? 3u : 1u);
      const int ppr = (int)(rb >> 4), rpi = 32 / ppr;          // 16-byte pieces per row, rows per warp instruction
      const int piece = lane % ppr;
      const int row_first = pidx * (128 / kXProducers) + lane / ppr;   // this producer copies rows [pidx * 64, +64) of every chunk
      const int iters = (128 / kXProducers) / rpi;
      const int pt = p.kt >> 1;
      const long long pitch2 = (long long)p.x_pitch * 2;
      const long long d_row = (long long)(p.W - p.RP) * pitch2;            // image row wrap (xs -= RP, ++y)
      const long long d_frame = (long long)(p.H - p.FP) * p.W * pitch2;    // frame wrap (y: FP -> 0, ++t)
      const long long d_step = (long long)rpi * pitch2;
      const long long frame_bytes = (long long)p.H * p.W * pitch2;
      const int RP = p.RP, FP = p.FP, W = p.W, H = p.H, T = p.T;
      int slot = 0;
      uint32_t parity = 1;
      ATileIter it;
      for (it.init(p); it.valid(p); it.next(p)) {
        const int m = it.m();
        const bool outside = m < 0 || m >= p.tiles_per_clip;
        // (xs, y, t) and byte offset of this lane's first row; later rows advance by rpi slots
        const int f0 = (outside ? 0 : m * 128) + row_first;
        const int xs0 = f0 & (RP - 1), r0 = f0 >> p.rp_shift;
        const int t0 = r0 / FP, y0 = r0 - t0 * FP;
        const long long g0 = ((((long long)it.n * T + t0) * H + y0) * W + xs0) * pitch2 + piece * 16;
        for (int q = 0; q < cpt; q += cps) {
          FB_T(0, mbar_wait(&x_empty[slot], parity));
          if (outside) {
            mbar_arrive(&x_full[slot]);   // every slot of the tile is padding: nothing to copy
          } else {
            const uint32_t slot_s = smem_u32(smem) + (uint32_t)slot * p.stage_bytes;
            for (int c = 0; c < cps; ++c) {
              const int qq = q + c;
              const int dt = qq / cch, cc = qq - dt * cch;
              const uint32_t dst_s = slot_s + (uint32_t)c * p.chunk_bytes;
              uint32_t off = (uint32_t)row_first * rb + (uint32_t)piece * 16u;
              int xs = xs0, y = y0, tt = t0 + dt - pt;
              long long g = g0 + (long long)(dt - pt) * frame_bytes + (long long)(cc * kc) * 2;
#pragma unroll 2
              for (int i = 0; i < iters; ++i) {
                const bool ok = xs < W && y < H && (unsigned)tt < (unsigned)T;
                cp_async16(dst_s + (off ^ (((off >> 7) & swz_mask) << 4)), p.x + (ok ? g : 0ll), ok ? 16u : 0u);
                off += (uint32_t)rpi * rb;
                xs += rpi;
                g += d_step;
                while (xs >= RP) {
                  xs -= RP;
                  g += d_row;
                  if (++y == FP) {
                    y = 0;
                    ++tt;
                    g += d_frame;
                  }
                }
              }
            }
            cp_async_arrive_noinc(&x_full[slot]);   // fires when this lane's copies of the slot have landed
          }
          if (++slot == S) {
            slot = 0;
            parity ^= 1;
          }
        }
      }
    } else if (warp == kProducerWarp && lane == 0) {
      // ------------------------------------------------------------ x producer, TMA mode (one thread)
      int slot = 0;
      uint32_t parity = 1;  // first pass over the ring: slots are free
      const int nbox = p.boxes_per_tile, FP = p.FP, pt = p.kt >> 1;
      ATileIter it;
      for (it.init(p); it.valid(p); it.next(p)) {
        const int r0 = it.m() * p.rows_per_tile + p.row_bias;  // >= 0
        int bt[kMaxBoxes], by[kMaxBoxes];
        for (int b = 0; b < nbox; ++b) {
          const int r = r0 + b * p.box_rows;
          const int tq = r / FP;
          bt[b] = tq - p.row_bias / FP;  // frame (may be -1 or T: zero-filled by the TMA unit)
          by[b] = r - tq * FP;           // row inside the frame (>= H: zero-filled)
        }
        int dt = 0, cc = 0;
        const int x0 = (it.m() * 128) & (p.RP - 1);   // first slot of the tile inside its flat row (im2col mode)
        // halo tiles outside the clip (m = -1, m = tiles_per_clip): every slot is padding, the a-epilogue writes zeros
        // whatever the accumulator holds -> nothing to load (im2col base coordinates must stay inside the bounding box)
        const bool outside = p.x_im2col && (it.m() < 0 || it.m() >= p.tiles_per_clip);
        for (int q = 0; q < cpt; q += cps) {
          FB_T(0, mbar_wait(&x_empty[slot], parity));
          if (outside) {
            mbar_arrive(&x_full[slot]);
            if (++slot == S) {
              slot = 0;
              parity ^= 1;
            }
            continue;
          }
          mbar_expect_tx(&x_full[slot], p.stage_bytes);
          uint8_t* dst = smem + (uint32_t)slot * p.stage_bytes;
          for (int c = 0; c < cps; ++c) {
            if (p.x_im2col) {
              // 128 consecutive flat slots from (slot 0 of flat row by[0], frame bt[0]): the traversal wraps W -> H -> T
              // inside the padded bounding box [0, RP) x [0, FP) x [-pt, T - 1 + pt], out-of-tensor positions zero-filled
              tma_load_im2col_5d(dst, &map_x, &x_full[slot], cc * kc, x0, by[0], bt[0] - pt, it.n, 0, 0, (uint16_t)dt);
            } else {
              for (int b = 0; b < nbox; ++b)
                tma_load_5d(dst + b * p.box_bytes, &map_x, &x_full[slot], cc * kc, 0, by[b], bt[b] + dt - pt, it.n);
            }
            dst += p.chunk_bytes;
            if (++cc == cch) {
              cc = 0;
              ++dt;
            }
          }
          if (++slot == S) {
            slot = 0;
            parity ^= 1;
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ---------------------------------------------------------------- MMA issuer
    // descriptor low words carry the layout flags: (flags | address >> 4) + offsets never carries into bit 14
    const uint64_t x_hi = umma_smem_desc(0, p.x_row_bytes) & 0xFFFFFFFF00000000ull;
    const uint32_t x_fl = (uint32_t)(umma_smem_desc(0, p.x_row_bytes) & 0xFFFFC000ull);
    const uint64_t d_hi = umma_smem_desc(0, p.a_row_bytes) & 0xFFFFFFFF00000000ull;
    const uint32_t d_fl = (uint32_t)(umma_smem_desc(0, p.a_row_bytes) & 0xFFFFC000ull);
    const uint32_t smem_lo = (smem_u32(smem) & 0x3FFFFu) >> 4;
    const uint32_t xring_lo = x_fl | smem_lo, wa_lo = x_fl | (smem_lo + (p.off_wa >> 4));
    const uint32_t wb_lo = d_fl | (smem_lo + (p.off_wb >> 4)), wc_lo = d_fl | (smem_lo + (p.off_wc >> 4));
    const uint32_t ring_lo = d_fl | (smem_lo + (p.off_ring >> 4)), p_lo = d_fl | (smem_lo + (p.off_p >> 4));
    const uint32_t chunk_lo = p.chunk_bytes >> 4, tile_lo = p.tile_bytes >> 4, stage_lo = p.stage_bytes >> 4;
    const uint32_t wa_blk_lo = p.wa_block_bytes >> 4, wb_blk_lo = p.wb_block_bytes >> 4;
    const int S = p.stages, spt = p.slots_per_tile, cps = p.cps, ksx = p.ksteps_x, ksd = p.ksteps_d;
    const uint32_t idesc_a = p.idesc_a, idesc_b = p.idesc_b, idesc_c = p.idesc_c;
    const int n_b_prev = 4 * ksd, n_b_cen = 5 * ksd;  // taps 0-3 start in the previous tile (RP >= 8)
    (void)xring_lo; (void)wa_lo; (void)stage_lo; (void)wa_blk_lo; (void)chunk_lo; (void)cps; (void)ksx; (void)S; (void)spt;
    (void)x_hi; (void)idesc_a;
    // Steps s = 0 .. G+1 (G = a-tiles of this CTA): step s issues the b-tile whose centre is a-tile s-2 and the c-tile
    // of the b-tile issued in step s-1, then commits ring_free[s & 3] (arrives once everything issued so far is done).
    int s_ = 0;
    int gb = 0, gc = 0;  // b-tiles / c-tiles issued so far
    int b_waited = 0;    // b-tiles whose epilogue this warp has observed (in order)
    int j1 = -1, j2 = -1;  // walk-local index of a-tiles s-1 and s-2 (-1: none)
    mbar_wait(w_bar, 0);
    ATileIter it;
    it.init(p);
    for (;; ++s_) {
      const int g = s_;
      const bool have_a = it.valid(p);
      if (!have_a && j1 < 0 && j2 < 2) break;
      // a-tile s-1 is in the ring (every tile is observed, in order, exactly once: its ring-slot barrier flips again
      // only four tiles later, which the a-issuer cannot reach before this warp has passed this step)
      if (j1 >= 0) FB_T(3, mbar_wait(&a_ready[(g - 1) & (kRing - 1)], ((g - 1) >> 2) & 1));
      if (j1 >= 2) {
        // b-tile whose centre is a-tile s-2 (previous s-3, next s-1)
        while (b_waited < gb - 1) {
          FB_T(4, mbar_wait(&b_done[b_waited & 1], (b_waited >> 1) & 1));
          ++b_waited;
        }
        const long long i0 = kDbg ? clock64() : 0;
        tc_fence_after();
        if (elect_one()) {
          const uint32_t prev_lo = ring_lo + (uint32_t)((g - 3) & (kRing - 1)) * tile_lo;
          const uint32_t cen_lo = ring_lo + (uint32_t)((g - 2) & (kRing - 1)) * tile_lo;
          const uint32_t tm_d = tmem_base + p.tm_b + (uint32_t)(gb & 1) * p.tm_ab_stride;
          fb_issue(tm_d, d_hi, d_hi, prev_lo, wb_lo, tab_b, n_b_prev, idesc_b, 0u);
          fb_issue(tm_d, d_hi, d_hi, cen_lo, wb_lo, tab_b + n_b_prev, n_b_cen, idesc_b, 1u);
          umma_commit(&b_full[gb & 1]);
        }
        __syncwarp();
        if (kDbg) dbg_acc[19] += (uint32_t)(clock64() - i0);
        ++gb;
      }
      if (j2 >= 2) {
        // c-tile gc: needs the P tile of b-tile gc and a drained accumulator gc & 1
        while (b_waited < gc + 1) {
          FB_T(5, mbar_wait(&b_done[b_waited & 1], (b_waited >> 1) & 1));
          ++b_waited;
        }
        // c accumulator of tile gc: group gc & 1, buffer (gc >> 1) % c_bufs of that group, use number (gc >> 1) / c_bufs
        const int cgrp = gc & 1, cuse = gc >> 1;
        const int cbuf = p.c_bufs == 2 ? (cuse & 1) : 0, cphase = p.c_bufs == 2 ? (cuse >> 1) : cuse;
        const int cidx = cgrp * 2 + cbuf;
        FB_T(6, mbar_wait(&c_empty[cidx], (cphase & 1) ^ 1));
        const long long i0 = kDbg ? clock64() : 0;
        tc_fence_after();
        if (elect_one()) {
          fb_issue(tmem_base + p.tm_c + (uint32_t)(cgrp * p.c_bufs + cbuf) * p.tm_c_stride, d_hi, d_hi,
                   p_lo + (uint32_t)(gc & 1) * tile_lo, wc_lo, tab_c, ksd, idesc_c, 0u);
          umma_commit(&c_full[cidx]);
        }
        __syncwarp();
        if (kDbg) dbg_acc[20] += (uint32_t)(clock64() - i0);
        ++gc;
      }
      if (elect_one()) umma_commit(&ring_free[g & (kRing - 1)]);
      __syncwarp();
      j2 = j1;
      j1 = have_a ? it.j : -1;
      if (have_a) it.next(p);
    }
  } else if (warp == kMmaAWarp) {
    // ---------------------------------------------------------------- a-phase MMA issuer (second issuing warp)
    // Two issuing warps: each one's barrier round trips and loop overhead overlap the other's tensor-pipe time (one
    // warp issuing all three phases is issue-latency bound at ~1.7x the pipe time).  tcgen05.commit only tracks the
    // issuing thread's own MMAs, so the ring-slot reuse (a-tile g overwrites a-tile g-4, last read by the b-tile of
    // step g-1) is ordered through ring_free, which the a-EPILOGUE waits for before it writes the ring: this warp
    // only needs a drained accumulator and its x chunks.
    const uint64_t x_hi = umma_smem_desc(0, p.x_row_bytes) & 0xFFFFFFFF00000000ull;
    const uint32_t x_fl = (uint32_t)(umma_smem_desc(0, p.x_row_bytes) & 0xFFFFC000ull);
    const uint32_t smem_lo = (smem_u32(smem) & 0x3FFFFu) >> 4;
    const uint32_t xring_lo = x_fl | smem_lo, wa_lo = x_fl | (smem_lo + (p.off_wa >> 4));
    const uint32_t stage_lo = p.stage_bytes >> 4, wa_blk_lo = p.wa_block_bytes >> 4;
    const int S = p.stages, spt = p.slots_per_tile, cps = p.cps, n_a = p.cps * p.ksteps_x;
    const uint32_t idesc_a = p.idesc_a;
    int slot = 0;
    uint32_t xpar = 0;
    int g = 0;
    mbar_wait(w_bar, 0);
    ATileIter it;
    for (it.init(p); it.valid(p); it.next(p), ++g) {
      if (g >= 2) FB_T(1, mbar_wait(&a_done[g & 1], ((g - 2) >> 1) & 1));  // accumulator drained by a-tile g-2's epilogue
      const uint32_t tm_d = tmem_base + p.tm_a + (uint32_t)(g & 1) * p.tm_ab_stride;
      for (int q = 0; q < spt; ++q) {
        FB_T(2, mbar_wait(&x_full[slot], xpar));
        const long long i0 = kDbg ? clock64() : 0;
        if (p.x_cpasync) fence_proxy_async_smem();   // cp.async wrote the chunk through the generic proxy
        tc_fence_after();
        if (elect_one()) {
          fb_issue(tm_d, x_hi, x_hi, xring_lo + (uint32_t)slot * stage_lo, wa_lo + (uint32_t)(q * cps) * wa_blk_lo, tab_a, n_a,
                   idesc_a, q == 0 ? 0u : 1u);
          umma_commit(&x_empty[slot]);
          if (q == spt - 1) umma_commit(&a_full[g & 1]);
        }
        __syncwarp();
        if (kDbg) dbg_acc[18] += (uint32_t)(clock64() - i0);
        if (++slot == S) {
          slot = 0;
          xpar ^= 1;
        }
      }
    }
  } else if (warp < kBEpiWarp0) {
    // ---------------------------------------------------------------- a-epilogue (warps 0-3)
    const int quarter = warp & 3;
    const uint32_t swz_mask = p.a_row_bytes == 128 ? 7u : (p.a_row_bytes == 64 ? 3u : 1u);
    const uint32_t lane_taddr = tmem_base + p.tm_a + ((uint32_t)(quarter * 32) << 16);
    const uint32_t ring_s = smem_u32(smem + p.off_ring);
    const uint32_t sb_s = smem_u32(sb_a);
    const int row = quarter * 32 + lane;
    int g = 0;
    ATileIter it;
    for (it.init(p); it.valid(p); it.next(p), ++g) {
      // is my slot a real pixel (group) of the clip, or padding?
      const int f = it.m() * 128 + row;
      const int r = (f >> p.rp_shift) + p.row_bias;  // arithmetic shift: floor for the negative halo tile
      const int xs = f & (p.RP - 1);
      const int tq = r / p.FP;
      const int t = tq - p.row_bias / p.FP, y = r - tq * p.FP;
      const uint32_t keep = (xs < p.W && y < p.H && t >= 0 && t < p.T) ? 0xFFFFFFFFu : 0u;
      const int pos = g & (kRing - 1);
      const uint32_t dst = ring_s + (uint32_t)pos * p.tile_bytes;
      const uint32_t dst2 = pos == 0 ? ring_s + (uint32_t)kRing * p.tile_bytes : 0u;
      FB_T(7, mbar_wait(&a_full[g & 1], (g >> 1) & 1));
      // ring slot g & 3 still holds a-tile g-4 until every b-tile that reads it (issued through step g-1) has completed
      if (g >= 1) FB_T(21, mbar_wait(&ring_free[(g - 1) & (kRing - 1)], ((g - 1) >> 2) & 1));
      const long long w0 = kDbg ? clock64() : 0;
      tc_fence_after();
      fb_convert_rows(lane_taddr + (uint32_t)(g & 1) * p.tm_ab_stride, p.d16, dst, dst2, p.a_row_bytes, swz_mask, row,
                      sb_s, keep);
      fence_proxy_async_smem();  // ring writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&a_done[g & 1]);
        mbar_arrive(&a_ready[g & (kRing - 1)]);
      }
      if (kDbg) {
        dbg_acc[8] += (uint32_t)(clock64() - w0);
        dbg_acc[16] += 1;
      }
    }
  } else if (warp < kCEpiWarp0) {
    // ---------------------------------------------------------------- b-epilogue (warps 4-7)
    const int quarter = warp & 3;
    const uint32_t swz_mask = p.a_row_bytes == 128 ? 7u : (p.a_row_bytes == 64 ? 3u : 1u);
    const uint32_t lane_taddr = tmem_base + p.tm_b + ((uint32_t)(quarter * 32) << 16);
    const uint32_t p_s = smem_u32(smem + p.off_p);
    const uint32_t sb_s = smem_u32(sb_b);
    const int row = quarter * 32 + lane;
    int gb = 0;
    ATileIter it;
    for (it.init(p); it.valid(p); it.next(p)) {
      if (it.j < 2) continue;
      FB_T(9, mbar_wait(&b_full[gb & 1], (gb >> 1) & 1));
      const long long w0 = kDbg ? clock64() : 0;
      tc_fence_after();
      fb_convert_rows(lane_taddr + (uint32_t)(gb & 1) * p.tm_ab_stride, p.d16, p_s + (uint32_t)(gb & 1) * p.tile_bytes, 0u,
                      p.a_row_bytes, swz_mask, row, sb_s, 0xFFFFFFFFu);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&b_done[gb & 1]);
      if (kDbg) dbg_acc[10] += (uint32_t)(clock64() - w0);
      ++gb;
    }
  } else if (warp < kProducerWarp) {
    // ---------------------------------------------------------------- c-epilogue (warps 8-15)
    // group grp (four warps = the four TMEM lane quarters) takes the c-tiles of parity grp; every warp owns 32 GEMM
    // rows = a [bw slots x bh rows] box of the output and walks the tile's column chunks through its private slabs.
    const int cw = warp - kCEpiWarp0;
    const int quarter = cw & 3, grp = cw >> 2;
    const uint32_t epi_row_bytes = p.epi_n * 2;
    const uint32_t swz_mask = epi_row_bytes == 128 ? 7u : (epi_row_bytes == 64 ? 3u : 1u);
    const uint32_t lane_taddr = tmem_base + p.tm_c + (uint32_t)(grp * p.c_bufs) * p.tm_c_stride + ((uint32_t)(quarter * 32) << 16);
    const uint32_t slab_bytes = 32 * epi_row_bytes;
    const int nb = p.epi_bufs, chunks = p.epi_chunks;
    uint8_t* my_bufs = smem + p.off_epi + (size_t)cw * nb * slab_bytes;
    uint64_t* my_ready = epi_ready + cw * kMaxEpiBufs;
    const uint32_t sb_s = smem_u32(sb_c);
    // output box of (tile m, this warp): returns false when the whole box is padding
    auto box_of = [&](int n, int m, int& xs0, int& y, int& tn) {
      const int f0 = m * 128 + quarter * 32;
      const int r = f0 >> p.rp_shift;
      xs0 = f0 & (p.RP - 1);
      const int t = r / p.FP;
      y = r - t * p.FP;
      tn = n * p.T + t;
      return y < p.H && xs0 < p.W;
    };
    // next b-tile of this group at or after the iterator's position; gbx counts all b-tiles passed
    auto seek = [&](ATileIter& it, int& gbx) {
      while (it.valid(p)) {
        if (it.j >= 2) {
          if ((gbx & 1) == grp) return true;
          ++gbx;
        }
        it.next(p);
      }
      return false;
    };
    // prefetch cursor (lane 0): next chunk of this warp whose slab has not been armed yet
    ATileIter pf;
    int pf_gb = 0, pf_chunk = 0, pf_q = 0;
    pf.init(p);
    bool pf_ok = seek(pf, pf_gb);
    auto arm_next = [&]() {
      const int bsel = pf_q % nb;
      int xs0, y, tn;
      if (box_of(pf.n, pf.m0 + pf.j - 2, xs0, y, tn)) {
        mbar_expect_tx(&my_ready[bsel], slab_bytes);
        tma_load_4d(my_bufs + bsel * slab_bytes, &map_res, &my_ready[bsel], pf_chunk * p.epi_n, xs0, y, tn);
      } else {
        mbar_arrive(&my_ready[bsel]);
      }
      ++pf_q;
      if (++pf_chunk == chunks) {
        pf_chunk = 0;
        ++pf_gb;
        pf.next(p);
        pf_ok = seek(pf, pf_gb);
      }
    };
    if (lane == 0) {
      for (int i = 0; i < nb - 1 && pf_ok; ++i) arm_next();
    }
    __syncwarp();
    int q = 0, gbx = 0, tcount = 0;
    ATileIter it;
    it.init(p);
    while (seek(it, gbx)) {
      int xs0, y, tn;
      const bool live = box_of(it.n, it.m0 + it.j - 2, xs0, y, tn);
      const int cbuf = p.c_bufs == 2 ? (tcount & 1) : 0, cphase = p.c_bufs == 2 ? (tcount >> 1) : tcount;
      FB_T(11, mbar_wait(&c_full[grp * 2 + cbuf], cphase & 1));
      if (kDbg) dbg_acc[17] += 1;
      tc_fence_after();
      for (int c = 0; c < chunks; ++c) {
        const int b = q % nb;
        uint8_t* buf = my_bufs + b * slab_bytes;
        FB_T(12, mbar_wait(&my_ready[b], (q / nb) & 1));  // slab free (+ residual landed)
        const long long w0 = kDbg ? clock64() : 0;
        const int col0 = c * p.epi_n;
        if (live)
          epi_convert_chunk<true>(lane_taddr + (uint32_t)cbuf * p.tm_c_stride + col0, p.epi_n, smem_u32(buf), epi_row_bytes, swz_mask, lane,
                                  sb_s + col0 * 8, 0.f);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (live) tma_store_4d(&map_out, buf, col0, xs0, y, tn);  // slots >= W / rows >= H are clipped by the TMA unit
          tma_store_commit();
          if (kDbg) dbg_acc[13] += (uint32_t)(clock64() - w0);
          if (pf_ok) {
            FB_T(14, tma_store_wait_read1());
            arm_next();
          }
        }
        __syncwarp();
        ++q;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&c_empty[grp * 2 + cbuf]);
      ++tcount;
      ++gbx;
      it.next(p);
    }
    if (lane == 0) tma_store_wait_all();
  }

  if (kDbg && lane == 0 && (warp == kProducerWarp || warp == kMmaWarp || warp == kMmaAWarp || warp == 0 || warp == kBEpiWarp0 || warp == kCEpiWarp0)) {
    if (warp == 0) dbg_acc[15] = (uint32_t)(clock64() - dbg_start);
    for (int i = 0; i < 24; ++i)
      if (dbg_acc[i]) atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg) + i, (unsigned long long)dbg_acc[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    __syncwarp();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}


// ---------------------------------------------------------------------------------------------- TMA rate probe
// Every CTA streams `tiles` consecutive 128-row tiles of a [n, t, h, w, c] bf16 tensor into a ring of `stages`
// 16 KiB-or-less slots and frees each slot as soon as it is full (no consumer work): cycles per CTA -> bytes per
// cycle per SM of the load path alone.  mode 0: tiled 5-D boxes [kc, RP, box_rows] as the fused block loads them
// (RP > w: out-of-bounds slots zero-filled), kt boxes per tile position (frames t-1, t, t+1 ...); mode 1: plain 2-D
// boxes [kc, 128] over the flattened pixel list (a GEMM operand tile).
struct TmaProbeParams {
  int mode, kt, kc, cchunks, RP, rp_rows, box_rows, H, T, FP, tiles_per_clip, tiles, stages, n_clips;
  uint32_t chunk_bytes, box_bytes;
  long long* clk;
};

__global__ void __launch_bounds__(64, 1)
tma_rate_probe_kernel(const __grid_constant__ CUtensorMap map5, const __grid_constant__ CUtensorMap map2,
                      const __grid_constant__ TmaProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (uint32_t)p.stages * p.chunk_bytes);
  uint64_t* empty = full + p.stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    fence_mbar_init();
  }
  __syncthreads();
  const long long t0 = clock64();
  const int total_tiles = p.tiles_per_clip * p.n_clips;
  const int m_first = (int)(((long long)blockIdx.x * total_tiles) / gridDim.x);
  if (threadIdx.x == 0) {
    int slot = 0;
    uint32_t par = 1;
    for (int i = 0; i < p.tiles; ++i) {
      const int mt = (m_first + i) % total_tiles;
      const int n = mt / p.tiles_per_clip, m = mt - n * p.tiles_per_clip;
      for (int dt = 0; dt < p.kt; ++dt) {
        for (int cc = 0; cc < p.cchunks; ++cc) {
          mbar_wait(&empty[slot], par);
          mbar_expect_tx(&full[slot], p.chunk_bytes);
          uint8_t* dst = smem + (uint32_t)slot * p.chunk_bytes;
          if (p.mode == 0) {
            const int r0 = m * p.rp_rows;
            for (int b = 0; b < p.rp_rows / p.box_rows; ++b) {
              const int r = r0 + b * p.box_rows;
              tma_load_5d(dst + b * p.box_bytes, &map5, &full[slot], cc * p.kc, 0, r % p.FP, r / p.FP + dt - (p.kt >> 1), n);
            }
          } else if (p.mode == 2) {
            const int f0 = m * 128;
            tma_load_im2col_5d(dst, &map5, &full[slot], cc * p.kc, f0 % p.RP, (f0 / p.RP) % p.FP, f0 / (p.RP * p.FP) - (p.kt >> 1), n,
                               0, 0, (uint16_t)dt);
          } else {
            tma_load_2d(dst, &map2, &full[slot], cc * p.kc, mt * 128 + dt * 4096);
          }
          if (++slot == p.stages) {
            slot = 0;
            par ^= 1;
          }
        }
      }
    }
  } else if (threadIdx.x == 32) {
    int slot = 0;
    uint32_t par = 0;
    const int loads = p.tiles * p.kt * p.cchunks;
    for (int i = 0; i < loads; ++i) {
      mbar_wait(&full[slot], par);
      mbar_arrive(&empty[slot]);
      if (++slot == p.stages) {
        slot = 0;
        par ^= 1;
      }
    }
    p.clk[blockIdx.x] = clock64() - t0;
  }
}

}  // namespace vsb

// ------------------------------------------------------------------ host side
using namespace vsb;

struct vsb_bottleneck_plan {
  vsb_bottleneck_desc desc;
  ThinPlan* thin = nullptr;  // algo 1: the warp-MMA walk kernel owns the launch (bottleneck_thin_sm100.cu)
  CUtensorMap map_x, map_wa, map_wb, map_wc, map_res, map_out;
  FusedParams params;
  size_t smem_bytes;
  unsigned grid;
};

static int fb_encode(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                     const cuuint32_t* box, CUtensorMapSwizzle swz, const char* what) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims,
                              strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(%s) failed (CUresult %d)", what, (int)r);
    return VSB_ERR_CUDA;
  }
  return VSB_OK;
}

extern "C" int vsb_bottleneck_plan_create(const vsb_bottleneck_desc* d, vsb_bottleneck_plan** out_plan) {
  VSB_CHECK_ARG(d && out_plan, "null argument");
  *out_plan = nullptr;
  VSB_CHECK_ARG(d->x && d->out && d->wa && d->wb && d->wc, "null tensor pointer");
  VSB_CHECK_ARG(d->sa && d->ba && d->sb && d->bb && d->sc && d->bc, "null scale / bias pointer");
  VSB_CHECK_ARG(d->n > 0 && d->t > 0 && d->h > 0 && d->w > 0, "non-positive extent");
  VSB_CHECK_ARG(d->kt == 1 || d->kt == 3, "kt must be 1 or 3");
  VSB_CHECK_ARG(d->algo == 0 || d->algo == 1, "algo must be 0 (tcgen05 flat raster) or 1 (warp-MMA walk)");
  if (d->algo == 1) {
    VSB_CHECK_ARG(((reinterpret_cast<uintptr_t>(d->x) | reinterpret_cast<uintptr_t>(d->out) |
                    reinterpret_cast<uintptr_t>(d->wa) | reinterpret_cast<uintptr_t>(d->wb) |
                    reinterpret_cast<uintptr_t>(d->wc)) & 15) == 0, "tensor pointers must be 16-byte aligned");
    ThinPlan* thin = nullptr;
    const int rc = thin_plan_create(d, &thin);
    if (rc != VSB_OK) return rc;
    vsb_bottleneck_plan* plan = new (std::nothrow) vsb_bottleneck_plan();
    if (!plan) {
      thin_plan_destroy(thin);
      set_error("out of host memory");
      return VSB_ERR_INVALID;
    }
    plan->desc = *d;
    plan->thin = thin;
    *out_plan = plan;
    return VSB_OK;
  }
  VSB_CHECK_ARG(d->d == 16 || d->d == 32 || d->d == 64, "bottleneck width (stored) must be 16, 32 or 64");
  VSB_CHECK_ARG(d->c % 16 == 0 && d->c >= 16 && d->c <= 256, "block width (stored) must be a multiple of 16 in [16, 256]");
  VSB_CHECK_ARG(d->x_pitch >= d->c && d->out_pitch >= d->c && d->x_pitch % 8 == 0 && d->out_pitch % 8 == 0,
                "pitches must be >= c and multiples of 8 elements");
  VSB_CHECK_ARG(((reinterpret_cast<uintptr_t>(d->x) | reinterpret_cast<uintptr_t>(d->out) |
                  reinterpret_cast<uintptr_t>(d->wa) | reinterpret_cast<uintptr_t>(d->wb) |
                  reinterpret_cast<uintptr_t>(d->wc)) & 15) == 0, "tensor pointers must be 16-byte aligned");
  VSB_CHECK_ARG(d->w < 128, "row wider than 127 slots");

  FusedParams p{};
  p.W = d->w; p.H = d->h; p.T = d->t; p.n_clips = d->n;
  int RP = 8;
  while (RP < d->w + 1) RP <<= 1;
  p.RP = RP;
  p.rp_shift = 0;
  while ((1 << p.rp_shift) < RP) ++p.rp_shift;
  p.rows_per_tile = 128 / RP;
  // c-epilogue box: 32 slots = bw slots x bh rows
  p.bw = RP >= 32 ? 32 : RP;
  p.bh = 32 / p.bw;
  // x TMA boxes: whole tiles when a frame is a whole number of tiles (FP a multiple of the tile rows is cheap),
  // else single flat rows; output boxes of bh rows must not straddle frames
  // frame pitch FP (flat rows per frame, > H): a multiple of the output box rows, a clip a whole number of tiles; the
  // x boxes take as many rows as divide both the tile and FP -- the largest box whose FP rounding costs <= 7 % rows
  auto fp_for = [&](int q) {
    int fp = ceil_div(d->h + 1, q) * q;
    while (fp % p.bh || ((long long)d->t * fp * RP) % 128) fp += q;
    return fp;
  };
  const int fp_min = fp_for(1);
  int FP = fp_min, box_rows = 1;
  const int waste_pct = getenv("VSB_FUSED_BOXWASTE") ? atoi(getenv("VSB_FUSED_BOXWASTE")) : 7;  // experiments
  for (int br = p.rows_per_tile; br > 1; br >>= 1) {
    const int fp = fp_for(br);
    if (fp * 100 <= fp_min * (100 + waste_pct)) {
      FP = fp;
      box_rows = br;
      break;
    }
  }
  VSB_CHECK_ARG(FP * RP <= (1 << 20), "frame too large");
  p.FP = FP;
  p.box_rows = box_rows;
  p.boxes_per_tile = p.rows_per_tile / box_rows;
  VSB_CHECK_ARG(p.boxes_per_tile >= 1 && p.boxes_per_tile <= kMaxBoxes, "too many TMA boxes per tile");
  p.tiles_per_clip = (int)(((long long)d->t * FP * RP) / 128);
  p.row_bias = ceil_div(p.rows_per_tile, FP) * FP;

  // channels
  p.kt = d->kt;
  p.kc = d->c % 64 == 0 ? 64 : (d->c % 32 == 0 ? 32 : 16);
  p.cchunks = d->c / p.kc;
  p.chunks_per_tile = p.kt * p.cchunks;
  p.ksteps_x = p.kc / 16;
  p.d16 = d->d;
  p.ksteps_d = d->d / 16;
  p.cout = d->c;
  p.x_row_bytes = (uint32_t)p.kc * 2;
  p.a_row_bytes = d->d * 2 < 64 ? 64u : (uint32_t)d->d * 2;  // 32-byte operand rows run the tensor core ~3x slower (measured)
  p.chunk_bytes = 128u * p.x_row_bytes;
  p.box_bytes = (uint32_t)box_rows * RP * p.x_row_bytes;
  VSB_CHECK_ARG(p.box_bytes % 256 == 0, "x box is not a whole number of swizzle atoms");

  // c epilogue
  p.epi_n = d->c % 32 == 0 ? 32 : 16;
  p.epi_chunks = d->c / p.epi_n;
  p.epi_bufs = 2;

  // TMEM
  p.tm_ab_stride = d->d < 32 ? 32 : (uint32_t)d->d;
  p.tm_c_stride = d->c < 32 ? 32 : (uint32_t)d->c;
  p.tm_a = 0;
  p.tm_b = 2 * p.tm_ab_stride;
  p.tm_c = 4 * p.tm_ab_stride;
  p.c_bufs = (4 * p.tm_ab_stride + 4 * p.tm_c_stride <= 512) ? 2 : 1;
  const uint32_t need_cols = 4 * p.tm_ab_stride + 2 * p.c_bufs * p.tm_c_stride;
  VSB_CHECK_ARG(need_cols <= 512, "accumulators need %u TMEM columns (> 512)", need_cols);
  p.tmem_cols = 32;
  while (p.tmem_cols < need_cols) p.tmem_cols <<= 1;
  p.idesc_a = umma_idesc_bf16(128, d->d);
  p.idesc_b = umma_idesc_bf16(128, d->d);
  p.idesc_c = umma_idesc_bf16(128, d->c);

  // shared memory
  auto up1k = [](uint32_t v) { return (v + 1023u) & ~1023u; };
  p.wa_block_bytes = up1k((uint32_t)d->d * p.x_row_bytes);
  p.wb_block_bytes = up1k((uint32_t)d->d * p.a_row_bytes);
  p.wc_bytes = up1k((uint32_t)d->c * p.a_row_bytes);
  p.tile_bytes = 128u * p.a_row_bytes;
  const uint32_t w_bytes = p.chunks_per_tile * p.wa_block_bytes + 9 * p.wb_block_bytes + p.wc_bytes;
  const uint32_t ring_bytes = (kRing + 1) * p.tile_bytes, p_bytes = 2 * p.tile_bytes;
  const uint32_t epi_bytes = up1k((uint32_t)kCEpiWarps * p.epi_bufs * 32u * p.epi_n * 2u);
  const uint32_t tab_bytes = up1k((uint32_t)(2 * d->d + d->c) * 8u + (uint32_t)(p.kt * (d->c / 16) + 10 * (d->d / 16)) * 8u);
  const uint32_t fixed = w_bytes + ring_bytes + p_bytes + epi_bytes + 1024 + tab_bytes + 1024;
  // x ring: slots of cps chunks (one barrier pair per slot).  Prefer whole tiles per slot (fewest waits / commits in the
  // single issuing thread) when >= 2 tiles fit, else halves ..., else single chunks.
  const long long room = 227ll * 1024 - fixed;
  int cps = 1, stages = 0;
  const int cps_max = getenv("VSB_FUSED_CPS") ? atoi(getenv("VSB_FUSED_CPS")) : p.chunks_per_tile;  // experiments
  for (int cand = cps_max < p.chunks_per_tile ? cps_max : p.chunks_per_tile; cand >= 1; --cand) {
    if (p.chunks_per_tile % cand) continue;
    const long long sb = (long long)cand * p.chunk_bytes;
    const int st = room > 0 ? (int)(room / sb) : 0;
    if (st * cand >= 2 * p.chunks_per_tile || (cand == 1 && st >= 2)) {
      cps = cand;
      stages = st;
      break;
    }
  }
  if (d->stages > 0 && d->stages < stages * cps) {  // caller's cap, in chunks
    stages = d->stages / cps;
    if (stages < 2) {
      cps = 1;
      stages = d->stages;
    }
  }
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) {
    set_error("fused bottleneck does not fit in shared memory (weights %u B, fixed %u B, x chunk %u B)", w_bytes, fixed,
              p.chunk_bytes);
    return VSB_ERR_INVALID;
  }
  p.stages = stages;
  p.cps = cps;
  p.slots_per_tile = p.chunks_per_tile / cps;
  p.stage_bytes = (uint32_t)cps * p.chunk_bytes;
  // im2col-mode x loads when the padded bounding box fits the TMA corner range (measured: tiled 5-D boxes with
  // out-of-bounds slots run at 10 - 21 B/clk/SM, one im2col load per 16 KiB chunk at the 2-D rate)
  p.x_im2col = (RP - d->w <= 15 && FP - d->h <= 15 && !getenv("VSB_FUSED_TILED_X")) ? 1 : 0;
  // cp.async x producers are an experiment kept for the record (VSB_FUSED_X=cpasync): two warps of 16-byte copies
  // sustain only ~7 B/clk/SM (too few copies in flight), against ~23 B/clk/SM of im2col TMA loads
  p.x_cpasync = getenv("VSB_FUSED_X") ? (strcmp(getenv("VSB_FUSED_X"), "cpasync") == 0) : 0;
  p.x = static_cast<const uint8_t*>(d->x);
  p.x_pitch = d->x_pitch;
  p.pf_dist = 0;
  p.off_wa = (uint32_t)stages * p.stage_bytes;
  p.off_wb = p.off_wa + p.chunks_per_tile * p.wa_block_bytes;
  p.off_wc = p.off_wb + 9 * p.wb_block_bytes;
  p.off_ring = p.off_wc + p.wc_bytes;
  p.off_p = p.off_ring + ring_bytes;
  p.off_epi = p.off_p + p_bytes;
  p.off_bar = p.off_epi + epi_bytes;
  p.off_tab = p.off_bar + 1024;
  const size_t smem_bytes = (size_t)p.off_tab + tab_bytes + 1024;

  p.sa = d->sa; p.ba = d->ba; p.sb = d->sb; p.bb = d->bb; p.sc = d->sc; p.bc = d->bc;
  p.dbg = nullptr;
  if (getenv("VSB_FUSED_DEBUG")) {  // debug only: the one place this file allocates device memory
    if (cudaMalloc(&p.dbg, 32 * sizeof(long long)) == cudaSuccess) (void)cudaMemset(p.dbg, 0, 32 * sizeof(long long));
    else p.dbg = nullptr;
  }

  // every CTA owns one contiguous range of the clip-major tile index (cut into walks at clip boundaries)
  int sms = 148;
  {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) (void)cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    (void)cudaGetLastError();
  }
  if (d->grid > 0 && d->grid < sms) sms = d->grid;
  p.total_tiles = p.tiles_per_clip * d->n;
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  p.tiles_per_cta = ceil_div(p.total_tiles, grid);
  p.walk_len = (d->walk_len > 0 && d->walk_len < p.tiles_per_clip) ? d->walk_len : p.tiles_per_clip;

  int rc = load_driver_entry_points();
  if (rc != VSB_OK) return rc;
  vsb_bottleneck_plan* plan = new (std::nothrow) vsb_bottleneck_plan();
  VSB_CHECK_ARG(plan, "out of host memory");
  plan->desc = *d;
#define FB_FAIL(rc_)   \
  do {                 \
    delete plan;       \
    return rc_;        \
  } while (0)
  {
    // x: (C, W, H, T, N), box [kc, RP, box_rows, 1, 1]; out-of-bounds slots / rows / frames are zero-filled
    cuuint64_t dims[5] = {(cuuint64_t)d->c, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->t, (cuuint64_t)d->n};
    cuuint64_t str[4] = {(cuuint64_t)d->x_pitch * 2, (cuuint64_t)d->w * d->x_pitch * 2,
                         (cuuint64_t)d->h * d->w * d->x_pitch * 2, (cuuint64_t)d->t * d->h * d->w * d->x_pitch * 2};
    cuuint32_t box[5] = {(cuuint32_t)p.kc, (cuuint32_t)RP, (cuuint32_t)box_rows, 1, 1};
    if (p.x_im2col) {
      const int lower[3] = {0, 0, -(d->kt >> 1)};
      const int upper[3] = {RP - d->w, FP - d->h, (d->kt >> 1) - (d->kt - 1)};
      const int one3[3] = {1, 1, 1};
      rc = encode_im2col_map(&plan->map_x, d->x, d->n, d->t, d->h, d->w, d->c, d->x_pitch, lower, upper, one3, p.kc, 128,
                             swizzle_for((int)p.x_row_bytes));
    } else {
      rc = fb_encode(&plan->map_x, d->x, 5, dims, str, box, swizzle_for((int)p.x_row_bytes), "block input");
    }
    if (rc != VSB_OK) FB_FAIL(rc);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)d->kt * d->c, (cuuint64_t)d->d};
    cuuint64_t str[1] = {(cuuint64_t)d->kt * d->c * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.kc, (cuuint32_t)d->d};
    rc = fb_encode(&plan->map_wa, d->wa, 2, dims, str, box, swizzle_for((int)p.x_row_bytes), "a weights");
    if (rc != VSB_OK) FB_FAIL(rc);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)9 * d->d, (cuuint64_t)d->d};
    cuuint64_t str[1] = {(cuuint64_t)9 * d->d * 2};
    cuuint32_t box[2] = {(cuuint32_t)(p.a_row_bytes / 2), (cuuint32_t)d->d};
    rc = fb_encode(&plan->map_wb, d->wb, 2, dims, str, box, swizzle_for((int)p.a_row_bytes), "b weights");
    if (rc != VSB_OK) FB_FAIL(rc);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)d->d, (cuuint64_t)d->c};
    cuuint64_t str[1] = {(cuuint64_t)d->d * 2};
    cuuint32_t box[2] = {(cuuint32_t)(p.a_row_bytes / 2), (cuuint32_t)d->c};
    rc = fb_encode(&plan->map_wc, d->wc, 2, dims, str, box, swizzle_for((int)p.a_row_bytes), "c weights");
    if (rc != VSB_OK) FB_FAIL(rc);
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->c, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->t * d->n};
    cuuint32_t box[4] = {(cuuint32_t)p.epi_n, (cuuint32_t)p.bw, (cuuint32_t)p.bh, 1};
    cuuint64_t ostr[3] = {(cuuint64_t)d->out_pitch * 2, (cuuint64_t)d->w * d->out_pitch * 2,
                          (cuuint64_t)d->h * d->w * d->out_pitch * 2};
    rc = fb_encode(&plan->map_out, d->out, 4, dims, ostr, box, swizzle_for(p.epi_n * 2), "block output");
    if (rc != VSB_OK) FB_FAIL(rc);
    cuuint64_t rstr[3] = {(cuuint64_t)d->x_pitch * 2, (cuuint64_t)d->w * d->x_pitch * 2,
                          (cuuint64_t)d->h * d->w * d->x_pitch * 2};
    rc = fb_encode(&plan->map_res, d->x, 4, dims, rstr, box, swizzle_for(p.epi_n * 2), "block residual");
    if (rc != VSB_OK) FB_FAIL(rc);
  }
  plan->params = p;
  plan->smem_bytes = smem_bytes;
  plan->grid = (unsigned)ceil_div(p.total_tiles, p.tiles_per_cta);

  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(bottleneck_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(bottleneck_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  if (attr_err != cudaSuccess) {
    set_error("cudaFuncSetAttribute(bottleneck_fused_kernel) failed: %s", cudaGetErrorString(attr_err));
    FB_FAIL(VSB_ERR_CUDA);
  }
  *out_plan = plan;
  return VSB_OK;
}

extern "C" int vsb_bottleneck_run(const vsb_bottleneck_plan* plan, void* stream) {
  VSB_CHECK_ARG(plan, "null plan");
  if (plan->thin) return thin_run(plan->thin, stream);
  if (plan->params.dbg)
    bottleneck_fused_kernel<true><<<plan->grid, kThreads, plan->smem_bytes, static_cast<cudaStream_t>(stream)>>>(
        plan->map_x, plan->map_wa, plan->map_wb, plan->map_wc, plan->map_res, plan->map_out, plan->params);
  else
    bottleneck_fused_kernel<false><<<plan->grid, kThreads, plan->smem_bytes, static_cast<cudaStream_t>(stream)>>>(
        plan->map_x, plan->map_wa, plan->map_wb, plan->map_wc, plan->map_res, plan->map_out, plan->params);
  VSB_CHECK_LAUNCH("bottleneck_fused_kernel");
  return VSB_OK;
}

extern "C" void vsb_bottleneck_plan_destroy(vsb_bottleneck_plan* plan) {
  if (plan && plan->thin) thin_plan_destroy(plan->thin);
  delete plan;
}

extern "C" int vsb_debug_bottleneck_stats(const vsb_bottleneck_plan* plan, long long* out32) {
  VSB_CHECK_ARG(plan && out32, "null argument");
  VSB_CHECK_ARG(!plan->thin, "the warp-MMA kernel keeps no role timeline");
  VSB_CHECK_ARG(plan->params.dbg, "plan was not created with VSB_FUSED_DEBUG=1");
  VSB_CHECK_CUDA(cudaDeviceSynchronize());
  VSB_CHECK_CUDA(cudaMemcpy(out32, plan->params.dbg, 32 * sizeof(long long), cudaMemcpyDeviceToHost));
  VSB_CHECK_CUDA(cudaMemset(plan->params.dbg, 0, 32 * sizeof(long long)));
  return VSB_OK;
}

extern "C" int vsb_bottleneck_plan_info(const vsb_bottleneck_plan* plan, long long* out8) {
  VSB_CHECK_ARG(plan && out8, "null argument");
  if (plan->thin) {  // {rows per strip, strips per frame, ring slots, frame steps per CTA, grid, smem bytes, a tiles * 1000 + bc tiles, 0}
    thin_plan_info(plan->thin, out8);
    return VSB_OK;
  }
  const FusedParams& p = plan->params;
  out8[0] = p.RP;
  out8[1] = p.FP;
  out8[2] = p.stages * 100 + p.cps;
  out8[3] = p.tiles_per_cta;
  out8[4] = plan->grid;
  out8[5] = (long long)plan->smem_bytes;
  out8[6] = p.tiles_per_clip;
  out8[7] = p.tmem_cols;
  return VSB_OK;
}

// TMA load-path probe (see tma_rate_probe_kernel): x bf16 [n, t, h, w, c] dense; clk int64 [grid] cycles per CTA.
extern "C" int vsb_debug_tma_rate(const void* x, int n, int t, int h, int w, int c, int mode, int kt, int box_rows,
                                  int stages, int tiles, int grid, long long* clk, void* stream) {
  VSB_CHECK_ARG(x && clk && n > 0 && t > 0 && h > 0 && w > 0 && c % 16 == 0 && tiles > 0 && grid > 0, "bad argument");
  int rc = load_driver_entry_points();
  if (rc != VSB_OK) return rc;
  TmaProbeParams p{};
  p.mode = mode; p.kt = kt; p.n_clips = n; p.H = h; p.T = t;
  p.kc = c % 64 == 0 ? 64 : (c % 32 == 0 ? 32 : 16);
  p.cchunks = c / p.kc;
  int RP = 8;
  while (RP < w + 1) RP <<= 1;
  p.RP = RP;
  p.rp_rows = 128 / RP;
  VSB_CHECK_ARG(box_rows >= 1 && p.rp_rows % box_rows == 0, "box_rows must divide the tile rows");
  p.box_rows = box_rows;
  p.FP = ceil_div(h + 1, box_rows) * box_rows;
  while (((long long)t * p.FP * RP) % 128) p.FP += box_rows;
  p.tiles_per_clip = mode != 1 ? (int)((long long)t * p.FP * RP / 128) : (int)((long long)t * h * w / 128);
  p.tiles = tiles; p.stages = stages;
  p.chunk_bytes = 128u * p.kc * 2;
  p.box_bytes = (uint32_t)box_rows * RP * p.kc * 2;
  p.clk = clk;
  CUtensorMap m5, m2;
  {
    cuuint64_t dims[5] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)t, (cuuint64_t)n};
    cuuint64_t str[4] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2, (cuuint64_t)t * h * w * c * 2};
    cuuint32_t box[5] = {(cuuint32_t)p.kc, (cuuint32_t)RP, (cuuint32_t)box_rows, 1, 1};
    if (mode == 2) {
      const int lower[3] = {0, 0, -(kt >> 1)};
      const int upper[3] = {RP - w, p.FP - h, (kt >> 1) - (kt - 1)};
      const int one3[3] = {1, 1, 1};
      VSB_CHECK_ARG(RP - w <= 15 && p.FP - h <= 15, "padded box outside the im2col corner range");
      rc = encode_im2col_map(&m5, x, n, t, h, w, c, c, lower, upper, one3, p.kc, 128, swizzle_for(p.kc * 2));
    } else {
      rc = fb_encode(&m5, x, 5, dims, str, box, swizzle_for(p.kc * 2), "probe 5d");
    }
    if (rc != VSB_OK) return rc;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)c, (cuuint64_t)n * t * h * w};
    cuuint64_t str[1] = {(cuuint64_t)c * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.kc, 128};
    rc = fb_encode(&m2, x, 2, dims, str, box, swizzle_for(p.kc * 2), "probe 2d");
    if (rc != VSB_OK) return rc;
  }
  const size_t smem = (size_t)stages * p.chunk_bytes + 2048 + 1024;
  VSB_CHECK_ARG(smem <= 227 * 1024, "ring too large");
  VSB_CHECK_CUDA(cudaFuncSetAttribute(tma_rate_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  tma_rate_probe_kernel<<<grid, 64, smem, static_cast<cudaStream_t>(stream)>>>(m5, m2, p);
  VSB_CHECK_LAUNCH("tma_rate_probe_kernel");
  return VSB_OK;
}
