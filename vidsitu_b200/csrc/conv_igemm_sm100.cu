// Conv3d (+ folded frozen BatchNorm, ReLU, residual add) as an implicit GEMM on
// the sm_100a tensor cores.
//
// Reference ops replaced: nn.Conv3d + nn.BatchNorm3d(eval) + nn.ReLU + residual
// add of SlowFast/slowfast/models/{stem_helper.py:157-178, resnet_helper.py:182-240,
// 326-358, video_model_builder.py:109-131, nonlocal_helper.py:77-90}.
//
// GEMM view:  D[M, Cout] = A[M, K] * W[Cout, K]^T,  M = N*To*Ho*Wo output pixels,
// K = (kt,kh,kw,cin).  A is never materialised: for every filter tap the TMA unit
// gathers the 128 x kchunk activation tile straight from the NTHWC tensor in
// im2col mode (zero-filling the conv padding halo) into 32/64/128B-swizzled shared
// memory; W tiles arrive through a tiled TMA map.  One elected thread issues
// tcgen05.mma (M=128, N=block_n, K=16) with the fp32 accumulator in TMEM; four
// epilogue warps read it back with tcgen05.ld and apply scale/bias/residual/ReLU
// in fp32 before the single bf16 rounding of the layer.
//
// CTA = 6 warps: 0-3 epilogue (TMEM lane quarter = warp id), 4 = TMA producer,
// 5 = TMEM allocator + MMA issuer.  One output tile per CTA; several CTAs are
// co-resident per SM (smem/TMEM permitting) so one CTA's epilogue overlaps
// another's main loop.
#include <cuda.h>  // CUtensorMap types only; entry points are fetched at run time

#include <stdlib.h>

#include <mutex>
#include <new>

#include "common.h"
#include "conv_plan.h"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace vsb {

constexpr int kBlockM = 128;
constexpr int kMaxEpiWarps = 16;
constexpr int kMaxEpiBufs = 4;  // per epilogue warp


// Persistent CTA: loops over output tiles (tile = blockIdx.x + i * gridDim.x, the n-tile index
// fastest so co-running CTAs share the activation tile in L2).  Three pipelines:
//   smem  full/empty[stages]   TMA producer  <-> MMA issuer
//   TMEM  full/empty[2]        MMA issuer    <-> epilogue (two accumulators: the epilogue of tile i
//                                                overlaps the main loop of tile i+1)
//   epi   ready[2]             staging buffers of the epilogue (residual TMA load in, TMA store out)
// Optional role timeline (kDbg, VSB_WIN_DEBUG=1 at plan time): cycles each role spends waiting, summed over CTAs
// (same slots as conv_win_sm100.cu).
#define IG_T(idx, stmt)                                   \
  do {                                                    \
    if (kDbg) {                                           \
      const long long _t0 = clock64();                    \
      stmt;                                               \
      dbg_acc[idx] += (uint32_t)(clock64() - _t0);        \
    } else {                                              \
      stmt;                                               \
    }                                                     \
  } while (0)

// KK = kchunk / 16 MMAs per channel chunk (CPS = 4 / KK chunks per pipeline stage); EW = epilogue warps:
// 8 (two CTAs per SM: <= 102 registers per thread) or 16 (one CTA per SM; four warps per TMEM lane quarter
// hide the latency chain TMEM -> registers -> shared memory of layers whose epilogue is the bottleneck).
template <int KK, bool kDbg, int EW>
__global__ void __launch_bounds__((EW + 2) * 32, EW == 8 ? 2 : 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_res,
                  const __grid_constant__ CUtensorMap map_a2, const IgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kEpiWarps = EW, kProducerWarp = EW, kMmaWarp = EW + 1;
  pdl_launch_dependents();  // the next kernel's prologue may overlap this kernel (it waits for our completion itself)

  const uint32_t row_bytes = p.kchunk * 2;
  const uint32_t a_chunk_bytes = kBlockM * row_bytes;
  const uint32_t b_chunk_bytes = p.block_n * row_bytes;
  const uint32_t stage_bytes = p.stage_bytes;
  const uint32_t epi_row_bytes = p.epi_n * 2;
  uint8_t* bres = smem + p.off_bres;   // resident weights (b_resident)
  uint8_t* epi_buf = smem + p.off_epi;  // epi_bufs buffers, each a multiple of 1024 bytes
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full = empty_bar + p.stages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* epi_ready = tmem_empty + 2;
  uint64_t* bres_bar = epi_ready + kMaxEpiWarps * kMaxEpiBufs;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres_bar + 1);
  // The grid is a multiple of n_tiles, so tile % n_tiles == blockIdx.x % n_tiles for every tile of this
  // CTA: one fixed column block -> its (scale, bias) pairs live in shared memory, and its weights can too.
  const int n_tile = blockIdx.x % p.n_tiles;
  const int m_tile0 = blockIdx.x / p.n_tiles, m_tile_step = gridDim.x / p.n_tiles;
  float2* sb_tab = reinterpret_cast<float2*>(smem + p.off_bar + 1024);
  for (int c = threadIdx.x; c < p.block_n; c += blockDim.x)
    sb_tab[c] = make_float2(p.scale ? p.scale[n_tile * p.block_n + c] : 1.f, p.bias[n_tile * p.block_n + c]);

  const int warp = threadIdx.x >> 5;  // warp-uniform
  const int lane = threadIdx.x & 31;
  uint32_t dbg_acc[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const long long dbg_start = kDbg ? clock64() : 0;

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&map_a);
    if (p.chunks1 < p.total_chunks) tma_prefetch_desc(&map_a2);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_out);
    if (p.has_residual) tma_prefetch_desc(&map_res);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kEpiWarps);  // one arrival per epilogue warp
    }
    for (int i = 0; i < kEpiWarps * kMaxEpiBufs; ++i) mbar_init(&epi_ready[i], 1);
    mbar_init(bres_bar, 1);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_kstages = (p.total_chunks + p.cps - 1) / p.cps;

  if (warp == kProducerWarp) {
    if (lane == 0) {
      // ------------------------------------------------------ TMA producer (one thread)
      // The loop body is kept free of divisions: filter-tap / channel-chunk indices advance as
      // counters, the weight K coordinate is simply chunk * kchunk.
      if (p.b_resident) {
        // weight-stationary: every chunk of the [block_n x K] weight matrix is loaded exactly once
        mbar_expect_tx(bres_bar, p.total_chunks * b_chunk_bytes);
        for (int g = 0; g < p.total_chunks; ++g)
          tma_load_2d(bres + g * b_chunk_bytes, &map_b, bres_bar, g * p.kchunk, n_tile * p.block_n);
      }
      // activations are the previous kernels' outputs (weights, scale / bias are constants).  A chained consumer
      // (tile_wait) runs next to its producer instead and waits tile by tile below.
      unsigned int* const tile_wait = p.tile_wait;
      if (!tile_wait) pdl_wait();
      const bool b_resident = p.b_resident != 0;
      const uint32_t stage_tx = b_resident ? a_chunk_bytes : a_chunk_bytes + b_chunk_bytes;
      const int cps = 4 / KK, kchunk = 16 * KK;
      const uint32_t b_off = cps * a_chunk_bytes;
      const int total_chunks = p.total_chunks, stages = p.stages, m_tiles = p.total_tiles / p.n_tiles;
      const int cin = p.cin, fkw = p.kw, fkh = p.kh, block_n = p.block_n;
      const int owo = p.wo, oho = p.ho, oto = p.to, sw = p.sw, sh = p.sh, st = p.st, lw = p.lw, lh = p.lh, lt = p.lt;
      const int chunks1 = p.chunks1, sw2 = p.sw2, sh2 = p.sh2, st2 = p.st2;
      int slot = 0;
      uint32_t parity = 1;  // first pass over the ring: slots are free
      const int clip_tiles = p.clip_tiles, clip_rows = p.clip_rows, wgt_clip_rows = p.wgt_clip_rows;
      const bool rev = p.reverse != 0;
      for (int mt_ = m_tile0; mt_ < m_tiles; mt_ += m_tile_step) {
        const int mt = rev ? m_tiles - 1 - mt_ : mt_;
        if (tile_wait) tile_counter_wait_reset(tile_wait + mt, p.tile_wait_count);  // the producer kernel finished tile mt
        int m0 = mt * kBlockM, wrow0 = 0;
        if (clip_rows) {  // per-clip weights: tile lt of clip cb (its last tile runs into the next clip: never stored)
          const int cb = mt / clip_tiles;
          m0 = cb * clip_rows + (mt - cb * clip_tiles) * kBlockM;
          wrow0 = cb * wgt_clip_rows;
        }
        const int wo = m0 % owo;
        const int r1 = m0 / owo;
        const int ho = r1 % oho;
        const int r2 = r1 / oho;
        const int to_ = r2 % oto;
        const int n0 = r2 / oto;
        const int w0 = wo * sw + lw;
        const int h0 = ho * sh + lh;
        const int d0 = to_ * st + lt;
        const int ncol = n_tile * block_n + wrow0;
        int cc = 0, kw_ = 0, kh_ = 0, kt_ = 0, kcoord = 0;
        int left = total_chunks;
        int left1 = chunks1;  // chunks still to come from the primary source
        const int w2 = wo * sw2, h2 = ho * sh2, d2 = to_ * st2;
        while (left > 0) {
          const int nch = left < cps ? left : cps;
          left -= nch;
          IG_T(0, mbar_wait(&empty_bar[slot], parity));
          mbar_expect_tx(&full_bar[slot], nch * stage_tx);
          uint8_t* a_dst = smem + (uint32_t)slot * stage_bytes;
          uint8_t* b_dst = a_dst + b_off;
          for (int c = 0; c < nch; ++c) {
            if (left1 > 0) {
              tma_load_im2col_5d(a_dst, &map_a, &full_bar[slot], cc, w0, h0, d0, n0, (uint16_t)kw_, (uint16_t)kh_,
                                 (uint16_t)kt_);
              if (--left1 == 0) cc = -kchunk;  // the second source starts at its channel 0
            } else {
              // fused shortcut projection: same 128 output pixels, strided 1x1x1 gather from the block input
              // (channels past cin2 are zero-filled by the TMA unit, their weights are zero)
              tma_load_im2col_5d(a_dst, &map_a2, &full_bar[slot], cc, w2, h2, d2, n0, 0, 0, 0);
            }
            if (!b_resident) tma_load_2d(b_dst, &map_b, &full_bar[slot], kcoord, ncol);
            a_dst += a_chunk_bytes;
            b_dst += b_chunk_bytes;
            kcoord += kchunk;
            cc += kchunk;
            if (cc == cin && left1 > 0) {
              cc = 0;
              if (++kw_ == fkw) {
                kw_ = 0;
                if (++kh_ == fkh) {
                  kh_ = 0;
                  ++kt_;
                }
              }
            }
          }
          if (++slot == stages) {
            slot = 0;
            parity ^= 1;
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // -------------------------------------------------------- MMA issuer
    // The whole warp walks the loop (uniform control flow keeps descriptors in uniform registers);
    // one elected lane issues tcgen05.mma / tcgen05.commit.  Descriptors differ only in their low
    // 32 bits (start address >> 4); a full stage is CPS*KK = 4 fully unrolled MMAs.  Every kernel
    // parameter used in the loop is copied to a local first (one constant-bank load, not one per use).
    constexpr int CPS = 4 / KK;
    const uint64_t desc_hi = umma_smem_desc(0, row_bytes) & 0xFFFFFFFF00000000ull;
    const uint32_t desc_lo_flags = (uint32_t)(umma_smem_desc(0, row_bytes) & 0xFFFFC000ull);
    const uint32_t smem_lo = (smem_u32(smem) & 0x3FFFFu) >> 4;
    const uint32_t bres_lo = (smem_u32(bres) & 0x3FFFFu) >> 4;
    const uint32_t stage_lo = stage_bytes >> 4, a_chunk_lo = a_chunk_bytes >> 4, b_chunk_lo = b_chunk_bytes >> 4;
    const uint32_t b_off_lo = (CPS * a_chunk_bytes) >> 4;
    const uint32_t idesc = p.idesc;
    const int total_chunks = p.total_chunks, stages = p.stages, m_tiles = p.total_tiles / p.n_tiles, block_n = p.block_n;
    const int full_stages = total_chunks / CPS, tail_chunks = total_chunks - full_stages * CPS;
    const bool b_resident = p.b_resident != 0;
    int slot = 0, tcount = 0;
    uint32_t parity = 0, a_slot_lo = smem_lo;
    if (b_resident) mbar_wait(bres_bar, 0);
    for (int mt = m_tile0; mt < m_tiles; mt += m_tile_step, ++tcount) {
      const int acc = tcount & 1;
      IG_T(1, mbar_wait(&tmem_empty[acc], ((tcount >> 1) & 1) ^ 1));  // epilogue drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * block_n;
      uint32_t bres_cur = bres_lo;
      for (int ks = 0; ks < full_stages; ++ks) {
        IG_T(2, mbar_wait(&full_bar[slot], parity));
        tc_fence_after();
        const long long issue_t0 = kDbg ? clock64() : 0;
        if (elect_one()) {
          const uint32_t a_lo = a_slot_lo;
          const uint32_t b_lo = b_resident ? bres_cur : a_slot_lo + b_off_lo;
#pragma unroll
          for (int c = 0; c < CPS; ++c) {
#pragma unroll
            for (int k = 0; k < KK; ++k) {
              const uint64_t adesc = desc_hi | (uint64_t)(desc_lo_flags | (a_lo + c * a_chunk_lo + 2 * k));
              const uint64_t bdesc = desc_hi | (uint64_t)(desc_lo_flags | (b_lo + c * b_chunk_lo + 2 * k));
              umma_bf16(tmem_d, adesc, bdesc, idesc, (ks | c | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[slot]);  // frees the smem slot once these MMAs retire
          if (tail_chunks == 0 && ks == full_stages - 1) umma_commit(&tmem_full[acc]);  // accumulator complete
        }
        __syncwarp();
        if (kDbg) dbg_acc[3] += (uint32_t)(clock64() - issue_t0);
        bres_cur += CPS * b_chunk_lo;
        a_slot_lo += stage_lo;
        if (++slot == stages) {
          slot = 0;
          parity ^= 1;
          a_slot_lo = smem_lo;
        }
      }
      if (tail_chunks) {  // last, partially filled stage
        IG_T(2, mbar_wait(&full_bar[slot], parity));
        tc_fence_after();
        if (elect_one()) {
          uint32_t a_lo = a_slot_lo;
          uint32_t b_lo = b_resident ? bres_cur : a_slot_lo + b_off_lo;
          for (int c = 0; c < tail_chunks; ++c) {
#pragma unroll
            for (int k = 0; k < KK; ++k) {
              const uint64_t adesc = desc_hi | (uint64_t)(desc_lo_flags | (a_lo + 2 * k));
              const uint64_t bdesc = desc_hi | (uint64_t)(desc_lo_flags | (b_lo + 2 * k));
              umma_bf16(tmem_d, adesc, bdesc, idesc, (full_stages | c | k) != 0 ? 1u : 0u);
            }
            a_lo += a_chunk_lo;
            b_lo += b_chunk_lo;
          }
          umma_commit(&empty_bar[slot]);
          umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        a_slot_lo += stage_lo;
        if (++slot == stages) {
          slot = 0;
          parity ^= 1;
          a_slot_lo = smem_lo;
        }
      }
    }
  } else {
    // ---------------------------------------------------------- epilogue (warps 0 .. EW-1)
    // Independent per-warp pipelines: warp w owns TMEM lanes / tile rows [32*(w&3), +32) and every
    // (EW/4)-th column chunk (global chunk counter mod EW/4 == w>>2).  Each warp has its own staging
    // slabs (32 rows x epi_n), its own residual TMA loads and its own TMA stores -- no block-level
    // barrier anywhere in the epilogue.
    const int quarter = warp & 3, grp = warp >> 2;
    const uint32_t swz_mask = epi_row_bytes == 128 ? 7u : (epi_row_bytes == 64 ? 3u : 1u);
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t slab_bytes = 32 * epi_row_bytes;
    const int nb = p.epi_bufs;
    uint8_t* my_bufs = epi_buf + (size_t)warp * nb * slab_bytes;
    uint64_t* my_ready = epi_ready + warp * kMaxEpiBufs;
    const int row_in_tile = quarter * 32;
    const float relu_floor = p.relu ? 0.f : -__int_as_float(0x7f800000);
    const int m_tiles = p.total_tiles / p.n_tiles;
    const int nbase = n_tile * p.block_n;
    const uint32_t sb_s = smem_u32(sb_tab);
    const int epi_n = p.epi_n, epi_chunks = p.epi_chunks;
    const bool has_res = p.has_residual != 0;
    // prefetch cursor (lane 0): next chunk of THIS warp whose staging slab has not been armed yet
    int pf_mt = m_tile0, pf_chunk = 0, pf_gq = 0, pf_b = 0;
    auto pf_skip = [&]() {  // advance the cursor to the next chunk owned by this warp's group
      while (pf_mt < m_tiles && (pf_gq & (EW / 4 - 1)) != grp) {
        ++pf_gq;
        if (++pf_chunk == epi_chunks) {
          pf_chunk = 0;
          pf_mt += m_tile_step;
        }
      }
    };
    auto arm_next = [&]() {  // hand the next slab of the ring to that chunk: start its residual load, or mark it free
      if (has_res) {
        mbar_expect_tx(&my_ready[pf_b], slab_bytes);
        tma_load_2d(my_bufs + pf_b * slab_bytes, &map_res, &my_ready[pf_b], nbase + pf_chunk * epi_n,
                    (p.reverse ? m_tiles - 1 - pf_mt : pf_mt) * kBlockM + row_in_tile);
      } else {
        mbar_arrive(&my_ready[pf_b]);
      }
      if (++pf_b == nb) pf_b = 0;
      ++pf_gq;
      if (++pf_chunk == epi_chunks) {
        pf_chunk = 0;
        pf_mt += m_tile_step;
      }
      pf_skip();
    };
    // residual loads read, and the stores overwrite, memory the previous kernel may still be using (a chained
    // consumer's first store follows its first tile counter, i.e. the producer's own wait)
    if (!p.tile_wait) pdl_wait();
    if (lane == 0) {
      pf_skip();
      for (int i = 0; i < nb - 1 && pf_mt < m_tiles; ++i) arm_next();
    }
    __syncwarp();
    int sig_mt = -1, sig_stores = 0, gq_stores = 0;  // lane 0: tile to publish next, store groups issued up to it / so far
    int b = 0;            // staging slab of this warp's current chunk (ring of nb)
    uint32_t bpar = 0;    // parity of that slab's ready barrier
    int gq = 0;           // global chunk counter of the CTA (all tiles, all chunks)
    int tcount = 0;
    for (int mt_ = m_tile0; mt_ < m_tiles; mt_ += m_tile_step, ++tcount) {
      const int mt = p.reverse ? m_tiles - 1 - mt_ : mt_;
      const int m0 = mt * kBlockM;
      const int acc = tcount & 1;
      IG_T(4, mbar_wait(&tmem_full[acc], (tcount >> 1) & 1));
      if (kDbg) dbg_acc[9] += 1;
      tc_fence_after();
      for (int c = 0; c < epi_chunks; ++c, ++gq) {
        if ((gq & (EW / 4 - 1)) != grp) continue;
        uint8_t* buf = my_bufs + b * slab_bytes;
        IG_T(5, mbar_wait(&my_ready[b], bpar));  // slab free (+ residual landed)
        const long long math_t0 = kDbg ? clock64() : 0;
        const int col0 = c * epi_n;
        {
          const uint32_t taddr = lane_taddr + acc * p.block_n + col0;
          if (has_res)
            epi_convert_chunk<true>(taddr, epi_n, smem_u32(buf), epi_row_bytes, swz_mask, lane, sb_s + col0 * 8,
                                    relu_floor);
          else if (p.out_f16)
            epi_convert_chunk<false, true>(taddr, epi_n, smem_u32(buf), epi_row_bytes, swz_mask, lane,
                                           sb_s + col0 * 8, relu_floor);
          else
            epi_convert_chunk<false>(taddr, epi_n, smem_u32(buf), epi_row_bytes, swz_mask, lane, sb_s + col0 * 8,
                                     relu_floor);
        }
        fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA store
        __syncwarp();
        if (kDbg) dbg_acc[6] += (uint32_t)(clock64() - math_t0);
        if (lane == 0) {
          if (p.clip_rows) {  // 3-D map [cout, rows of a clip, clips]: rows past the clip's end are clipped
            const int cb = mt / p.clip_tiles;
            tma_store_3d(&map_out, buf, nbase + col0, (mt - cb * p.clip_tiles) * kBlockM + row_in_tile, cb);
          } else {
            tma_store_2d(&map_out, buf, nbase + col0, m0 + row_in_tile);  // rows >= m_total are clipped by the TMA unit
          }
          tma_store_commit();
          ++gq_stores;
          // arm the slab of this warp's chunk q + nb - 1 (the one chunk q - 1 used): its store must have read it
          if (pf_mt < m_tiles) {
            IG_T(7, tma_store_wait_read1());
            arm_next();
          }
        }
        __syncwarp();  // reconverge before the next warp-aligned tcgen05.ld
        if (++b == nb) {
          b = 0;
          bpar ^= 1;
        }
      }
      // every tcgen05.ld this warp issues for the accumulator has completed: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&tmem_empty[acc]);
        if (p.tile_signal) {
          // publish this warp's rows of the PREVIOUS tile to the chained consumer kernel: its store groups are the
          // ones older than the groups of the tile just issued, so waiting for them does not stall this warp
          if (sig_mt >= 0) {
            tma_store_wait_pending(gq_stores - sig_stores);
            fence_proxy_async_all();
            red_release_gpu_add(p.tile_signal + sig_mt, 1u);
          }
          sig_mt = mt;
          sig_stores = gq_stores;
        }
      }
    }
    if (lane == 0) {
      tma_store_wait_all();  // this warp's output bytes are in global memory before the CTA exits
      if (p.tile_signal && sig_mt >= 0) {
        fence_proxy_async_all();
        red_release_gpu_add(p.tile_signal + sig_mt, 1u);
      }
    }
  }

  if (kDbg && lane == 0 && (warp == kProducerWarp || warp == kMmaWarp || warp == 0)) {
    if (warp == 0) dbg_acc[8] = (uint32_t)(clock64() - dbg_start);
    for (int i = 0; i < 11; ++i)
      if (dbg_acc[i]) atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg) + i, (unsigned long long)dbg_acc[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    __syncwarp();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// Debug probe: one im2col TMA load of `pixels` x `channels` into smem, copied out
// verbatim (un-swizzled, SWIZZLE_NONE map).  Used by tools/gpu_probe.py to pin the
// im2col traversal semantics on real hardware.
__global__ void im2col_probe_kernel(const __grid_constant__ CUtensorMap map, int c, int w, int h, int d, int n,
                                    int ow, int oh, int od, int bytes, uint8_t* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, bytes);
    tma_load_im2col_5d(smem, &map, &bar, c, w, h, d, n, (uint16_t)ow, (uint16_t)oh, (uint16_t)od);
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}

// ------------------------------------------------------------------ host side

EncodeTiledFn g_encode_tiled = nullptr;
static EncodeIm2colFn g_encode_im2col = nullptr;
static int g_driver_version = 0;

int load_driver_entry_points() {
  static std::once_flag once;
  static int status = VSB_OK;
  std::call_once(once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
      set_error("cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
      (void)cudaGetLastError();
      status = VSB_ERR_CUDA;
      return;
    }
    g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
    fn = nullptr;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
      set_error("cuTensorMapEncodeIm2col not available: %s", cudaGetErrorString(e));
      (void)cudaGetLastError();
      status = VSB_ERR_CUDA;
      return;
    }
    g_encode_im2col = reinterpret_cast<EncodeIm2colFn>(fn);
    (void)cudaDriverGetVersion(&g_driver_version);
  });
  return status;
}

CUtensorMapSwizzle swizzle_for(int row_bytes) {
  return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                             : (row_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE));
}

// (C,W,H,D,N) im2col map over a bf16 NTHWC tensor.
int encode_im2col_map(CUtensorMap* map, const void* base, int n, int t, int h, int w, int c, int pitch,
                             const int lower[3], const int upper[3], const int stride_whd[3], int chan_box,
                             int pixel_box, CUtensorMapSwizzle swz) {
  cuuint64_t dims[5] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)t, (cuuint64_t)n};
  cuuint64_t strides[4] = {(cuuint64_t)pitch * 2, (cuuint64_t)w * pitch * 2, (cuuint64_t)h * w * pitch * 2,
                           (cuuint64_t)t * h * w * pitch * 2};
  cuuint32_t estr[5] = {1, (cuuint32_t)stride_whd[0], (cuuint32_t)stride_whd[1], (cuuint32_t)stride_whd[2], 1};
  CUresult r = g_encode_im2col(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides,
                               lower, upper, (cuuint32_t)chan_box, (cuuint32_t)pixel_box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeIm2col failed (CUresult %d): dims c=%d w=%d h=%d t=%d n=%d pitch=%d lower=(%d,%d,%d) "
              "upper=(%d,%d,%d) stride=(%d,%d,%d) box=%dx%d",
              (int)r, c, w, h, t, n, pitch, lower[0], lower[1], lower[2], upper[0], upper[1], upper[2],
              stride_whd[0], stride_whd[1], stride_whd[2], chan_box, pixel_box);
    return VSB_ERR_CUDA;
  }
  // Drivers up to CUDA 13.1 mis-encode im2col maps of tensors smaller than 128 KiB
  // (same work-around as CUTLASS: clear bit 21 of the second descriptor word).
  const unsigned long long bytes = (unsigned long long)n * t * h * w * pitch * 2;
  if (g_driver_version <= 13010 && bytes < 131072ull) {
    reinterpret_cast<uint64_t*>(map)[1] &= ~(1ull << 21);
  }
  return VSB_OK;
}

static int encode_tiled_2d(CUtensorMap* map, const void* base, long long inner, long long outer,
                           long long outer_stride_bytes, int box_inner, int box_outer, CUtensorMapSwizzle swz) {
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)outer_stride_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d): dims %lld x %lld stride %lld box %d x %d", (int)r, inner,
              outer, outer_stride_bytes, box_inner, box_outer);
    return VSB_ERR_CUDA;
  }
  return VSB_OK;
}

// fp32 CUDA-core path (conv_simt.cu)
int launch_conv_simt(const vsb_conv_desc& d, int to, int ho, int wo, cudaStream_t stream);

}  // namespace vsb

using namespace vsb;

extern "C" int vsb_conv3d_plan_create(const vsb_conv_desc* d, vsb_conv_plan** out_plan) {
  VSB_CHECK_ARG(d && out_plan, "null argument");
  *out_plan = nullptr;
  VSB_CHECK_ARG(d->dtype == VSB_BF16 || d->dtype == VSB_F32, "dtype must be VSB_BF16 or VSB_F32");
  VSB_CHECK_ARG(d->in && d->wgt && d->out && d->bias, "null tensor pointer");
  VSB_CHECK_ARG(d->scale || d->dtype == VSB_BF16, "scale may be null (= 1) only on the bf16 path");
  VSB_CHECK_ARG(d->n > 0 && d->t > 0 && d->h > 0 && d->w > 0 && d->cin > 0 && d->cout > 0, "non-positive extent");
  VSB_CHECK_ARG(d->kt > 0 && d->kh > 0 && d->kw > 0 && d->st > 0 && d->sh > 0 && d->sw > 0, "bad kernel/stride");
  VSB_CHECK_ARG(d->in_pitch >= d->cin && d->out_pitch >= d->cout, "pitch smaller than channel count");
  VSB_CHECK_ARG(!d->residual || d->res_pitch >= d->cout, "residual pitch smaller than cout");
  const int to = (d->t + d->pt_lo + d->pt_hi - d->kt) / d->st + 1;
  const int ho = (d->h + d->ph_lo + d->ph_hi - d->kh) / d->sh + 1;
  const int wo = (d->w + d->pw_lo + d->pw_hi - d->kw) / d->sw + 1;
  VSB_CHECK_ARG(to > 0 && ho > 0 && wo > 0, "empty output");
  const long long m_total = (long long)d->n * to * ho * wo;
  VSB_CHECK_ARG(m_total < (1ll << 31), "too many output pixels for one launch (%lld)", m_total);

  vsb_conv_plan* plan = new (std::nothrow) vsb_conv_plan();
  VSB_CHECK_ARG(plan, "out of host memory");
  plan->desc = *d;
  plan->to = to;
  plan->ho = ho;
  plan->wo = wo;
  plan->m_total = m_total;

  plan->algo = 1;
  if (d->in2 && (d->dtype == VSB_F32 || d->algo == 2)) {
    set_error("a second source needs the bf16 im2col algorithm");
    delete plan;
    return VSB_ERR_INVALID;
  }
  if (d->dtype == VSB_F32) {
    *out_plan = plan;
    return VSB_OK;
  }
  VSB_CHECK_ARG(d->algo >= 0 && d->algo <= 2, "algo must be 0 (auto), 1 (im2col) or 2 (window)");
  if (d->wgt_clip_rows < 0 || (d->wgt_clip_rows > 0 && (d->algo == 2 || d->residual || d->in2 || d->kt * d->kh * d->kw != 1 ||
                                                        d->st * d->sh * d->sw != 1))) {
    set_error("per-clip weights need a 1x1x1 stride-1 conv on the im2col algorithm without residual / second source");
    delete plan;
    return VSB_ERR_INVALID;
  }
  if (d->in_f16 && (d->dtype != VSB_BF16 || d->algo == 2 || d->residual || d->in2 ||
                    (d->flags & (VSB_PLAN_TWO_SM | VSB_PLAN_TWO_SM_RESIDENT)))) {
    set_error("in_f16 needs the one-SM im2col algorithm of the 16-bit path without residual / second source");
    delete plan;
    return VSB_ERR_INVALID;
  }
  if (d->out_f16 && (d->algo == 2 || d->residual || d->in2)) {
    set_error("out_f16 needs the im2col algorithm without residual / second source");
    delete plan;
    return VSB_ERR_INVALID;
  }
  if (d->algo == 2) {
    const int wrc = win_plan_build(plan, d, to, ho, wo);
    if (wrc == VSB_OK) {
      plan->algo = 2;
      *out_plan = plan;
      return VSB_OK;
    }
    if (wrc > 0) set_error("conv is outside the domain of the shared-memory window algorithm (algo = 2)");
    delete plan;
    return wrc > 0 ? VSB_ERR_INVALID : wrc;
  }
  for (int k = 0; k < 8; ++k) {
    if (d->kw_c_hi[k] != 0 && (d->kw_c_lo[k] != 0 || d->kw_c_hi[k] != d->cin)) {
      set_error("partial per-tap channel ranges need the window algorithm (algo = 2)");
      delete plan;
      return VSB_ERR_INVALID;
    }
  }

#define FAIL(code, ...)     \
  do {                      \
    set_error(__VA_ARGS__); \
    delete plan;            \
    return code;            \
  } while (0)

  // ---- bf16 tensor-core plan
  if (d->cin % 16 || d->cout % 16) FAIL(VSB_ERR_INVALID, "bf16 path needs cin (%d) and cout (%d) multiples of 16", d->cin, d->cout);
  if (d->in_pitch % 8 || d->out_pitch % 8 || (d->residual && d->res_pitch % 8))
    FAIL(VSB_ERR_ALIGN, "pitches must be multiples of 8 elements (16 bytes)");
  if ((reinterpret_cast<uintptr_t>(d->in) | reinterpret_cast<uintptr_t>(d->wgt) | reinterpret_cast<uintptr_t>(d->out) |
       reinterpret_cast<uintptr_t>(d->residual) | reinterpret_cast<uintptr_t>(d->scale) |
       reinterpret_cast<uintptr_t>(d->bias)) & 15)
    FAIL(VSB_ERR_ALIGN, "tensor pointers must be 16-byte aligned");
  const int lower[3] = {-d->pw_lo, -d->ph_lo, -d->pt_lo};
  const int upper[3] = {d->pw_hi - (d->kw - 1), d->ph_hi - (d->kh - 1), d->pt_hi - (d->kt - 1)};
  for (int i = 0; i < 3; ++i) {
    if (lower[i] < -16 || lower[i] > 15 || upper[i] < -16 || upper[i] > 15)
      FAIL(VSB_ERR_INVALID, "padding/kernel outside the im2col TMA corner range [-16,15]");
  }
  if (d->kt > 31 || d->kh > 31 || d->kw > 31) FAIL(VSB_ERR_INVALID, "kernel extent above the im2col offset range");

  int kchunk = d->kchunk;
  if (!kchunk) kchunk = (d->cin % 64 == 0) ? 64 : ((d->cin % 32 == 0) ? 32 : 16);
  if ((kchunk != 16 && kchunk != 32 && kchunk != 64) || d->cin % kchunk) FAIL(VSB_ERR_INVALID, "bad kchunk %d for cin %d", kchunk, d->cin);
  // tile-granular chaining: producer and consumer share every SM (each <= half of the shared memory and of TMEM)
  const bool chained = d->tile_signal != nullptr || d->tile_wait != nullptr;
  if (chained) {
    if (d->wgt_clip_rows || d->out_f16 || (d->flags & (VSB_PLAN_TWO_SM | VSB_PLAN_TWO_SM_RESIDENT)))
      FAIL(VSB_ERR_INVALID, "tile chaining needs the plain one-SM im2col kernel");
    if (d->tile_wait && (d->kt * d->kh * d->kw != 1 || d->st * d->sh * d->sw != 1 || d->tile_wait_count <= 0 || d->in2 || d->residual))
      FAIL(VSB_ERR_INVALID, "a chained consumer is a 1x1x1 stride-1 conv without residual / second source (tile_wait_count > 0)");
    if ((reinterpret_cast<uintptr_t>(d->tile_signal) | reinterpret_cast<uintptr_t>(d->tile_wait)) & 3)
      FAIL(VSB_ERR_ALIGN, "tile counters must be 4-byte aligned");
  }
  int block_n = d->block_n;
  if (chained && block_n > 128) FAIL(VSB_ERR_INVALID, "tile chaining: block_n <= 128 (two kernels share the 512 TMEM columns)");
  if (!block_n) {
    block_n = chained ? 128 : 256;
    while (block_n > 16 && d->cout % block_n) block_n >>= 1;
    if (block_n < 64 && d->cout > 64) {
      // no power-of-two column block (e.g. the 784 keys of a res3 non-local block as output channels):
      // the widest multiple of 16 that divides cout (784 -> 112, 208 -> 208)
      for (int bn = 256; bn > block_n; bn -= 16)
        if (d->cout % bn == 0) {
          block_n = bn;
          break;
        }
    }
  }
  if (block_n < 16 || block_n > 256 || block_n % 16 || d->cout % block_n) FAIL(VSB_ERR_INVALID, "bad block_n %d for cout %d", block_n, d->cout);
  const int taps = d->kt * d->kh * d->kw;
  const int cin_chunks = d->cin / kchunk;
  // optional second source: strided 1x1x1 shortcut projection accumulated into the same tile
  int cin2_chunks = 0;
  if (d->in2) {
    if (d->residual) FAIL(VSB_ERR_INVALID, "a conv takes either a residual or a second source, not both");
    if (d->cin2 <= 0 || d->in2_pitch < d->cin2 || d->in2_pitch % 8 || d->cin2 % 8)
      FAIL(VSB_ERR_INVALID, "second source: bad channel count / pitch (cin2 %d, pitch %d)", d->cin2, d->in2_pitch);
    if (d->st2 < 1 || d->sh2 < 1 || d->sw2 < 1 || d->st2 > 8 || d->sh2 > 8 || d->sw2 > 8)
      FAIL(VSB_ERR_INVALID, "second source: bad strides");
    if ((d->t2 - 1) / d->st2 + 1 != to || (d->h2 - 1) / d->sh2 + 1 != ho || (d->w2 - 1) / d->sw2 + 1 != wo)
      FAIL(VSB_ERR_INVALID, "second source [%d,%d,%d] / strides (%d,%d,%d) does not produce the conv's output [%d,%d,%d]",
           d->t2, d->h2, d->w2, d->st2, d->sh2, d->sw2, to, ho, wo);
    if (reinterpret_cast<uintptr_t>(d->in2) & 15) FAIL(VSB_ERR_ALIGN, "second source must be 16-byte aligned");
    cin2_chunks = ceil_div(d->cin2, kchunk);
  }
  const int chunks1 = taps * cin_chunks;
  const int total_chunks = chunks1 + cin2_chunks;
  const int cps = 64 / kchunk;  // a pipeline stage always has room for 64 K-elements
  const int num_kstages = ceil_div(total_chunks, cps);
  const int bar_bytes = 1024 + 2048;  // barriers + the CTA's (scale, bias) table (block_n <= 256 float2)
  static const bool no_bres_env = getenv("VSB_NO_BRES") != nullptr;
  const bool no_bres = no_bres_env || (d->flags & (VSB_PLAN_STREAM_WEIGHTS | VSB_PLAN_TWO_SM | VSB_PLAN_TWO_SM_RESIDENT)) || d->wgt_clip_rows > 0;
  static const char* ew_env = getenv("VSB_EPI_WARPS");
  // (measured on B200: the 16-warp epilogue shape does not beat 8 warps -- these layers are HBM-bound, the
  // accumulator wait is back-pressure -- so it is opt-in: VSB_EPI_WARPS=16)
  const int epi_warps = (ew_env && atoi(ew_env) == 16) ? 16 : 8;
  if (d->epi_n != 0 && d->epi_n != 16 && d->epi_n != 32 && d->epi_n != 64) FAIL(VSB_ERR_INVALID, "epi_n must be 0, 16, 32 or 64");
  if (d->epi_bufs != 0 && (d->epi_bufs < 2 || d->epi_bufs > kMaxEpiBufs)) FAIL(VSB_ERR_INVALID, "epi_bufs must be 0 or 2..4");
  const int epi_n_env = d->epi_n, epi_bufs_env = d->epi_bufs;  // caller's tuning (0 = automatic)

  // ---- shared-memory plan for one (block_n, weight residency) choice.
  // Weight-stationary: the CTA's [block_n x K] weight block is fetched once (every CTA owns ONE column block:
  // the grid is a multiple of n_tiles) instead of once per tile.  With a residual the epilogue keeps 3
  // staging slabs per warp so 2 residual chunks are in flight (HBM latency), otherwise 2.
  struct SmemPlan { int stages, epi_n, epi_bufs, stage_bytes; size_t smem_bytes; long long bres_bytes; uint32_t tmem_cols; bool fits; };
  auto plan_smem = [&](int bn, bool resident) {
    SmemPlan sp{};
    sp.tmem_cols = 32;  // two accumulators; TMEM allocations are powers of two >= 32 columns
    while (sp.tmem_cols < (uint32_t)(2 * bn)) sp.tmem_cols <<= 1;
    const int a_stage = cps * kBlockM * kchunk * 2, b_stage = cps * bn * kchunk * 2;
    sp.bres_bytes = (long long)total_chunks * bn * kchunk * 2;
    sp.stage_bytes = ((resident ? a_stage : a_stage + b_stage) + 1023) & ~1023;
    // epilogue column chunk: 64 columns (128-byte staging rows); compute-bound layers (long K loop, no
    // residual) and the 16-warp shape take 32-column chunks so the staging slabs leave room for the ring
    sp.epi_n = bn >= 64 ? 64 : bn;
    if (bn >= 64 && ((!d->residual && num_kstages >= 8) || epi_warps == 16)) sp.epi_n = 32;
    if (bn >= 64 && epi_n_env) sp.epi_n = epi_n_env;
    while (sp.epi_n > 16 && bn % sp.epi_n) sp.epi_n >>= 1;  // column blocks like 112 or 208: 16-column chunks
    if (bn % sp.epi_n) return sp;
    const int epi_buf_bytes = epi_warps * 32 * sp.epi_n * 2;  // one 32-row slab per epilogue warp
    sp.epi_bufs = d->residual ? 3 : 2;  // slabs per epilogue warp (residual prefetch distance = epi_bufs - 1)
    if (epi_bufs_env) sp.epi_bufs = epi_bufs_env;
    for (;;) {
      const long long fixed = (resident ? ((sp.bres_bytes + 1023) & ~1023ll) : 0) + sp.epi_bufs * epi_buf_bytes + bar_bytes + 1024;
      int stages = d->stages;
      if (!stages) {
        // 512 TMEM columns or 16 epilogue warps => one CTA per SM anyway: use the whole shared memory;
        // otherwise try to leave room for two CTAs per SM and fall back to one big CTA when that starves
        // the pipeline
        long long budget = ((sp.tmem_cols == 512 || epi_warps == 16) && !chained ? 227 : 113) * 1024 - fixed;
        stages = budget > 0 ? (int)(budget / sp.stage_bytes) : 0;
        if (stages < 3 && stages < 2 * num_kstages && !chained) stages = (int)((227 * 1024 - fixed) / sp.stage_bytes);
        if (stages > 8) stages = 8;
      }
      if (stages > 16) stages = 16;
      // a caller-given depth that does not fit is shortened rather than rejected
      while (d->stages && stages > 2 && (long long)stages * sp.stage_bytes + fixed > (chained ? 113 : 227) * 1024) --stages;
      if (stages > num_kstages * 2) stages = num_kstages * 2;
      sp.stages = stages;
      sp.smem_bytes = (size_t)((stages > 0 ? stages : 0) * (long long)sp.stage_bytes + fixed);
      const size_t smem_cap = (size_t)(chained ? 113 : 227) * 1024;
      if (stages >= 3 && sp.smem_bytes <= smem_cap) { sp.fits = true; break; }
      if (stages >= 1 && sp.smem_bytes <= smem_cap && (sp.epi_bufs == 2 || stages >= 2 * num_kstages)) { sp.fits = true; break; }
      if (sp.epi_bufs > 2) {
        sp.epi_bufs = 2;  // give the shared memory back to the main-loop pipeline
        continue;
      }
      break;
    }
    return sp;
  };
  // Choice: (1) one column block covering cout with resident weights (<= 96 KB); (2) several column blocks
  // with resident weights when block_n (or, unless the caller fixed it, block_n / 2) leaves room for a
  // >= 3-stage ring and the epilogue slabs -- memory-bound 1x1x1 layers with cout > 256 otherwise spend more
  // L2 -> SM bandwidth re-fetching weights per tile than on activations; (3) weights streamed with A.
  SmemPlan sp{};
  bool b_resident = false;
  if (!no_bres) {
    int cand[2] = {block_n, (!d->block_n && block_n == 256) ? 128 : 0};
    for (int ci = 0; ci < 2 && !b_resident; ++ci) {
      const int bn = cand[ci];
      if (!bn) continue;
      const int nt = d->cout / bn;
      SmemPlan t = plan_smem(bn, true);
      const long long limit = nt == 1 ? 96 * 1024 : 128 * 1024;
      if (t.fits && t.bres_bytes <= limit && (nt == 1 || (t.stages >= 3 && t.epi_bufs == (epi_bufs_env ? epi_bufs_env : (d->residual ? 3 : 2))))) {
        // compute-bound long-K layers keep the wide block (fewer A re-reads); resident mode is for short K
        if (nt > 1 && num_kstages > 8) continue;
        sp = t;
        block_n = bn;
        b_resident = true;
      }
    }
  }
  if (!b_resident) sp = plan_smem(block_n, false);
  if (block_n % (sp.epi_n ? sp.epi_n : 1) || !sp.epi_n) FAIL(VSB_ERR_INVALID, "block_n %d is not a multiple of the epilogue chunk", block_n);
  if (!sp.fits) FAIL(VSB_ERR_INVALID, "pipeline (%d stages x %d bytes) does not fit in shared memory", sp.stages, sp.stage_bytes);
  const int n_tiles = d->cout / block_n;
  if (d->tile_wait && n_tiles != 1) FAIL(VSB_ERR_INVALID, "a chained consumer needs one column block (cout %d <= 128)", d->cout);
  if (chained && sp.tmem_cols > 256) FAIL(VSB_ERR_INVALID, "tile chaining: the plan needs %u TMEM columns (> 256)", sp.tmem_cols);
  const uint32_t tmem_cols = sp.tmem_cols;
  const long long bres_bytes = sp.bres_bytes;
  const int stage_bytes = sp.stage_bytes, stages = sp.stages, epi_n = sp.epi_n, epi_bufs = sp.epi_bufs;
  const int epi_buf_bytes = epi_warps * 32 * epi_n * 2;
  const size_t smem_bytes = sp.smem_bytes;
  const uint32_t off_bres = (uint32_t)stages * stage_bytes;
  const uint32_t off_epi = off_bres + (b_resident ? (uint32_t)((bres_bytes + 1023) & ~1023ll) : 0u);
  const uint32_t off_bar = off_epi + epi_bufs * epi_buf_bytes;

  int rc = load_driver_entry_points();
  if (rc != VSB_OK) {
    delete plan;
    return rc;
  }
  const int stride_whd[3] = {d->sw, d->sh, d->st};
  const CUtensorMapSwizzle swz = swizzle_for(kchunk * 2);
  rc = encode_im2col_map(&plan->map_a, d->in, d->n, d->t, d->h, d->w, d->cin, d->in_pitch, lower, upper, stride_whd,
                         kchunk, kBlockM, swz);
  if (rc != VSB_OK) {
    delete plan;
    return rc;
  }
  plan->map_a2 = plan->map_a;
  if (d->in2) {
    const int zero3[3] = {0, 0, 0}, stride2[3] = {d->sw2, d->sh2, d->st2};
    rc = encode_im2col_map(&plan->map_a2, d->in2, d->n, d->t2, d->h2, d->w2, d->cin2, d->in2_pitch, zero3, zero3,
                           stride2, kchunk, kBlockM, swz);
    if (rc != VSB_OK) {
      delete plan;
      return rc;
    }
  }
  const long long k_total = (long long)taps * d->cin + (long long)cin2_chunks * kchunk;
  const int clip_rows = d->wgt_clip_rows > 0 ? to * ho * wo : 0;
  // per-clip weights: clip i's [cout, K] matrix starts wgt_clip_rows rows after clip i-1's (they may overlap
  // when cout was rounded up: the extra rows are the next clip's, their products are discarded by the caller)
  const long long wgt_rows = clip_rows ? (long long)(d->n - 1) * d->wgt_clip_rows + d->cout : d->cout;
  rc = encode_tiled_2d(&plan->map_b, d->wgt, k_total, wgt_rows, k_total * 2, kchunk, block_n, swz);
  if (rc != VSB_OK) {
    delete plan;
    return rc;
  }
  // output / residual tiles move through TMA too: [m_total rows, cout cols], row pitch in bytes
  const CUtensorMapSwizzle epi_swz = swizzle_for(epi_n * 2);
  if (clip_rows) {
    cuuint64_t dims[3] = {(cuuint64_t)d->cout, (cuuint64_t)clip_rows, (cuuint64_t)d->n};
    cuuint64_t strides[2] = {(cuuint64_t)d->out_pitch * 2, (cuuint64_t)clip_rows * d->out_pitch * 2};
    cuuint32_t box[3] = {(cuuint32_t)epi_n, 32, 1}, estr[3] = {1, 1, 1};
    CUresult r = g_encode_tiled(&plan->map_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d->out, dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, epi_swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    rc = VSB_OK;
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(per-clip output) failed (CUresult %d)", (int)r);
      rc = VSB_ERR_CUDA;
    }
  } else
  rc = encode_tiled_2d(&plan->map_out, d->out, d->cout, m_total, (long long)d->out_pitch * 2, epi_n, 32, epi_swz);
  if (rc != VSB_OK) {
    delete plan;
    return rc;
  }
  if (d->residual) {
    rc = encode_tiled_2d(&plan->map_res, d->residual, d->cout, m_total, (long long)d->res_pitch * 2, epi_n, 32,
                         epi_swz);
    if (rc != VSB_OK) {
      delete plan;
      return rc;
    }
  } else {
    plan->map_res = plan->map_out;
  }
#undef FAIL

  IgemmParams& p = plan->params;
  p.m_total = (int)m_total;
  p.to = to; p.ho = ho; p.wo = wo;
  p.st = d->st; p.sh = d->sh; p.sw = d->sw;
  p.lt = lower[2]; p.lh = lower[1]; p.lw = lower[0];
  p.kh = d->kh; p.kw = d->kw;
  p.cin = d->cin; p.cin_chunks = cin_chunks; p.total_chunks = total_chunks; p.cps = cps; p.kchunk = kchunk;
  p.chunks1 = chunks1;
  p.st2 = d->in2 ? d->st2 : 1; p.sh2 = d->in2 ? d->sh2 : 1; p.sw2 = d->in2 ? d->sw2 : 1;
  p.block_n = block_n; p.n_tiles = n_tiles; p.stages = stages;
  p.epi_bufs = epi_bufs; p.b_resident = b_resident ? 1 : 0;
  p.epi_warps = epi_warps;
  p.stage_bytes = stage_bytes; p.off_bres = off_bres; p.off_epi = off_epi; p.off_bar = off_bar;
  p.clip_rows = clip_rows;
  p.clip_tiles = clip_rows ? ceil_div(clip_rows, kBlockM) : 0;
  p.wgt_clip_rows = d->wgt_clip_rows;
  p.total_tiles = clip_rows ? d->n * p.clip_tiles * p.n_tiles : (int)(ceil_div_ll(m_total, kBlockM) * p.n_tiles);
  p.epi_n = epi_n; p.epi_chunks = block_n / epi_n;
  p.idesc = d->in_f16 ? umma_idesc_f16(kBlockM, block_n) : umma_idesc_bf16(kBlockM, block_n);
  p.tmem_cols = tmem_cols;
  p.scale = d->scale; p.bias = d->bias;
  p.has_residual = d->residual != nullptr;
  p.relu = d->relu;
  p.out_f16 = d->out_f16 ? 1 : 0;
  p.reverse = (d->flags & VSB_PLAN_REVERSE) ? 1 : 0;
  p.tile_signal = d->tile_signal;
  p.tile_wait = d->tile_wait;
  p.tile_wait_count = (unsigned)d->tile_wait_count;
  p.dbg = nullptr;
  if (getenv("VSB_WIN_DEBUG")) {  // debug only: the one place the library allocates device memory
    if (cudaMalloc(&p.dbg, 16 * sizeof(long long)) == cudaSuccess) (void)cudaMemset(p.dbg, 0, 16 * sizeof(long long));
    else p.dbg = nullptr;
  }
  plan->smem_bytes = smem_bytes;
#define FAIL2(code, ...)    \
  do {                      \
    set_error(__VA_ARGS__); \
    delete plan;            \
    return code;            \
  } while (0)
  int ctas_per_sm = (int)(512 / tmem_cols);
  const int by_smem = (int)((227 * 1024) / smem_bytes);
  if (ctas_per_sm > by_smem) ctas_per_sm = by_smem;
  if (ctas_per_sm > 2) ctas_per_sm = 2;
  if (ctas_per_sm < 1 || epi_warps == 16 || (d->flags & VSB_PLAN_ONE_CTA)) ctas_per_sm = 1;
  int sms = 148;
  {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) (void)cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    (void)cudaGetLastError();
  }
  long long grid = (long long)sms * ctas_per_sm;
  if (grid > p.total_tiles) grid = p.total_tiles;
  if (d->grid_limit > 0 && grid > d->grid_limit) grid = d->grid_limit;
  grid -= grid % p.n_tiles;  // every CTA owns one column block (total_tiles is a multiple of n_tiles too)
  if (grid < p.n_tiles) FAIL2(VSB_ERR_INVALID, "grid_limit %d is below the %d column blocks", d->grid_limit, p.n_tiles);
  plan->grid = (unsigned)grid;
  plan->desc.block_n = block_n;
  plan->desc.kchunk = kchunk;
  plan->desc.stages = stages;

  // ---- two-SM variant (conv_igemm2_sm100.cu): a CTA pair per 256-pixel tile, each CTA loads half of the
  // weight rows.  Re-plans the ring for the smaller stage; everything else (maps, epilogue) is shared.
  // Automatic for the layers it was built for (measured on B200, SF50 batch 64: every res4/res5 conv with
  // streamed weights gains 12 - 20 %, s5 `b` reaches 1.39 PFLOP/s): 256-wide column blocks and K >= 512.
  static const bool no_two_sm_env = getenv("VSB_NO_TWO_SM") != nullptr;
  // (64-byte rows, i.e. the pixel-grouped stems: 2 - 4 % faster in pairs once a ring stage carries two chunks;
  // with one chunk = two MMAs per stage the issuer was the bottleneck: 0.46 vs 0.42 ms)
  const bool two_sm_auto = !no_two_sm_env && !chained && !d->in_f16 && !(d->flags & VSB_PLAN_ONE_SM) && !d->out_f16 && !d->wgt_clip_rows && block_n == 256 && (kchunk == 64 || kchunk == 32) &&
                           total_chunks >= 8 && p.total_tiles / p.n_tiles >= 16;
  if (((d->flags & (VSB_PLAN_TWO_SM | VSB_PLAN_TWO_SM_RESIDENT)) || two_sm_auto) && !d->out_f16 && !d->wgt_clip_rows && (kchunk == 64 || kchunk == 32) && !b_resident &&
      block_n % 16 == 0 && block_n >= 32 && epi_warps == 8 && p.total_tiles / p.n_tiles >= 2) {
    // weight-stationary pair (opt-in): each CTA keeps its half of the [block_n x K] block, the ring carries A only
    const long long bres_half = (long long)total_chunks * (block_n / 2) * kchunk * 2;
    const bool resident2 = (d->flags & VSB_PLAN_TWO_SM_RESIDENT) && bres_half <= 112 * 1024;
    const uint32_t stage2 = resident2 ? (uint32_t)(kBlockM * 128)
                                      : (uint32_t)((kBlockM + block_n / 2) * 128);  // 64 K-elements per stage
    const long long bres2 = resident2 ? ((bres_half + 1023) & ~1023ll) : 0;
    const long long fixed2 = bres2 + (long long)epi_bufs * epi_buf_bytes + bar_bytes + 1024;
    int stages2 = (int)((227 * 1024 - fixed2) / stage2);
    if (d->stages && stages2 > d->stages) stages2 = d->stages;
    if (stages2 > 12) stages2 = 12;
    if (stages2 > num_kstages * 2) stages2 = num_kstages * 2;
    if (stages2 >= 2) {
      rc = encode_tiled_2d(&plan->map_b, d->wgt, k_total, d->cout, k_total * 2, kchunk, block_n / 2, swz);
      if (rc != VSB_OK) {
        delete plan;
        return rc;
      }
      p.stages = stages2;
      p.stage_bytes = stage2;
      p.off_bres = (uint32_t)stages2 * stage2;
      p.off_epi = p.off_bres + (uint32_t)bres2;
      p.b_resident = resident2 ? 1 : 0;
      p.off_bar = p.off_epi + epi_bufs * epi_buf_bytes;
      p.idesc = umma_idesc_bf16(256, block_n);
      plan->smem_bytes = (size_t)stages2 * stage2 + (size_t)fixed2;
      plan->desc.stages = stages2;
      const int pm_tiles = (p.total_tiles / p.n_tiles + 1) / 2;
      long long pairs = sms / 2;
      if (pairs > (long long)pm_tiles * p.n_tiles) pairs = (long long)pm_tiles * p.n_tiles;
      pairs -= pairs % p.n_tiles;
      if (pairs >= p.n_tiles && pairs >= 1) {
        plan->grid = (unsigned)(2 * pairs);
        plan->algo = 3;
      }
    }
  }

  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
#define VSB_IG_ATTR(KK, DBG)                                                                                     \
  if (attr_err == cudaSuccess)                                                                                   \
    attr_err = cudaFuncSetAttribute(conv_igemm_kernel<KK, DBG, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                    227 * 1024);                                                                 \
  if (attr_err == cudaSuccess)                                                                                   \
    attr_err = cudaFuncSetAttribute(conv_igemm_kernel<KK, DBG, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                    227 * 1024)
    VSB_IG_ATTR(1, false); VSB_IG_ATTR(2, false); VSB_IG_ATTR(4, false);
    VSB_IG_ATTR(1, true); VSB_IG_ATTR(2, true); VSB_IG_ATTR(4, true);
#undef VSB_IG_ATTR
  });
  if (attr_err != cudaSuccess) {
    set_error("cudaFuncSetAttribute(conv_igemm_kernel) failed: %s", cudaGetErrorString(attr_err));
    delete plan;
    return VSB_ERR_CUDA;
  }
  *out_plan = plan;
  return VSB_OK;
}

extern "C" int vsb_conv3d_run(const vsb_conv_plan* plan, void* stream) {
  VSB_CHECK_ARG(plan, "null plan");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (plan->desc.dtype == VSB_F32) return launch_conv_simt(plan->desc, plan->to, plan->ho, plan->wo, s);
  if (plan->algo == 2) return win_plan_launch(plan, s);
  if (plan->algo == 3) return igemm2_launch(plan, s);
#define VSB_IG_LAUNCH(KK, DBG)                                                                              \
  do {                                                                                                      \
    if (plan->params.epi_warps == 16)                                                                       \
      (void)launch_pdl(conv_igemm_kernel<KK, DBG, 16>, plan->grid, 18 * 32, plan->smem_bytes, s, 1,         \
                       plan->map_a, plan->map_b, plan->map_out, plan->map_res, plan->map_a2, plan->params); \
    else                                                                                                    \
      (void)launch_pdl(conv_igemm_kernel<KK, DBG, 8>, plan->grid, 10 * 32, plan->smem_bytes, s, 1,          \
                       plan->map_a, plan->map_b, plan->map_out, plan->map_res, plan->map_a2, plan->params); \
  } while (0)
  const bool dbg = plan->params.dbg != nullptr;
  switch (plan->params.kchunk) {
    case 16:
      if (dbg) VSB_IG_LAUNCH(1, true); else VSB_IG_LAUNCH(1, false);
      break;
    case 32:
      if (dbg) VSB_IG_LAUNCH(2, true); else VSB_IG_LAUNCH(2, false);
      break;
    default:
      if (dbg) VSB_IG_LAUNCH(4, true); else VSB_IG_LAUNCH(4, false);
      break;
  }
#undef VSB_IG_LAUNCH
  VSB_CHECK_LAUNCH("conv_igemm_kernel");
  return VSB_OK;
}

extern "C" void vsb_conv3d_plan_destroy(vsb_conv_plan* plan) { delete plan; }

extern "C" int vsb_conv3d_plan_out_shape(const vsb_conv_plan* plan, int* to, int* ho, int* wo) {
  VSB_CHECK_ARG(plan, "null plan");
  if (to) *to = plan->to;
  if (ho) *ho = plan->ho;
  if (wo) *wo = plan->wo;
  return VSB_OK;
}

extern "C" double vsb_conv3d_plan_flops(const vsb_conv_plan* plan) {
  if (!plan) return 0.0;
  const vsb_conv_desc& d = plan->desc;
  return 2.0 * (double)plan->m_total * d.cout * ((double)d.kt * d.kh * d.kw * d.cin + (d.in2 ? d.cin2 : 0));
}

// ---- debug: im2col probe (declared in include/vidsitu_b200_debug.h)
extern "C" int vsb_debug_im2col_probe(const void* in, int n, int t, int h, int w, int c, int pitch, int lw, int lh,
                                      int lt, int uw, int uh, int ut, int sw, int sh, int st, int chan_box,
                                      int pixel_box, int cc, int cw, int ch, int cd, int cn, int ow, int oh, int od,
                                      void* out, void* stream) {
  int rc = load_driver_entry_points();
  if (rc != VSB_OK) return rc;
  CUtensorMap map;
  const int lower[3] = {lw, lh, lt}, upper[3] = {uw, uh, ut}, stride_whd[3] = {sw, sh, st};
  rc = encode_im2col_map(&map, in, n, t, h, w, c, pitch, lower, upper, stride_whd, chan_box, pixel_box,
                         CU_TENSOR_MAP_SWIZZLE_NONE);
  if (rc != VSB_OK) return rc;
  const int bytes = chan_box * pixel_box * 2;
  VSB_CHECK_ARG(bytes <= 64 * 1024, "probe box too large");
  VSB_CHECK_CUDA(cudaFuncSetAttribute(im2col_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
  im2col_probe_kernel<<<1, 128, bytes + 1024, static_cast<cudaStream_t>(stream)>>>(
      map, cc, cw, ch, cd, cn, ow, oh, od, bytes, static_cast<uint8_t*>(out));
  VSB_CHECK_LAUNCH("im2col_probe_kernel");
  return VSB_OK;
}

// ---- debug: role timeline counters of a window-algorithm plan (VSB_WIN_DEBUG=1 at plan creation)
extern "C" int vsb_debug_conv_stats(const vsb_conv_plan* plan, long long* out16) {
  VSB_CHECK_ARG(plan && out16, "null argument");
  long long* dbg = plan->algo == 2 ? plan->win.dbg : plan->params.dbg;
  VSB_CHECK_ARG(plan->desc.dtype == VSB_BF16 && dbg, "plan has no debug counters (bf16 plans created under VSB_WIN_DEBUG)");
  VSB_CHECK_CUDA(cudaDeviceSynchronize());
  VSB_CHECK_CUDA(cudaMemcpy(out16, dbg, 16 * sizeof(long long), cudaMemcpyDeviceToHost));
  VSB_CHECK_CUDA(cudaMemset(dbg, 0, 16 * sizeof(long long)));
  out16[13] = plan->grid;
  out16[14] = plan->algo == 2 ? plan->win.stages : plan->params.stages;
  out16[15] = (long long)plan->smem_bytes;
  return VSB_OK;
}

extern "C" int vsb_debug_conv_plan_info(const vsb_conv_plan* plan, long long* out8) {
  VSB_CHECK_ARG(plan && out8, "null argument");
  const bool win = plan->algo == 2;
  out8[0] = plan->algo;
  out8[1] = win ? plan->win.tsc : 0;
  out8[2] = win ? plan->win.stages : plan->params.stages;
  out8[3] = win ? plan->win.nacc : 2;
  out8[4] = plan->grid;
  out8[5] = (long long)plan->smem_bytes;
  out8[6] = win ? plan->win.block_n : plan->params.block_n;
  out8[7] = plan->smem_bytes <= 113 * 1024 ? 2 : 1;
  return VSB_OK;
}
