// Two-SM variant of the im2col implicit-GEMM conv (conv_igemm_sm100.cu): a CTA PAIR (cluster of 2,
// tcgen05 cta_group::2) computes a 256-pixel x block_n tile.
//
// Why: the compute-bound convs of res4/res5 (3x1x1 `a`, 1x3x3 `b`: N = 256, K = 2304 - 3072, weights far
// too large to stay in shared memory) stream 128 x 64 of A AND block_n x 64 of B per K stage and CTA:
// 48 KB per 512 tensor-core clocks = 94 B/clk/SM at peak, above what L2 feeds an SM (role timeline: the
// MMA issuer waits for data 35 - 45 % of the time, tensor pipe 73 %).  In a pair each CTA loads its own
// 128 pixels of A but only HALF of the weight rows (block_n / 2), the UMMA reads the other half from the
// peer's shared memory: 32 KB per CTA per stage for the same MACs per SM.
//
// Protocol (reference contract unchanged: same epilogue, same results as the one-SM kernel):
//   * both CTAs run a TMA producer; every load signals the LEADER's (cluster rank 0) full barrier
//     (cp.async.bulk.tensor ... cta_group::2, barrier address mapped to rank 0); the leader arms it with
//     the byte count of both CTAs;
//   * only the leader issues tcgen05.mma.cta_group::2 (M = 256: rows 0-127 from its smem / into its
//     TMEM, rows 128-255 the peer's); tcgen05.commit ... multicast::cluster releases the ring slot and
//     publishes the accumulator in BOTH CTAs;
//   * each CTA's 8 epilogue warps drain their own TMEM half exactly as in the one-SM kernel and arrive
//     on the leader's tmem_empty barrier (remote mbarrier.arrive for the peer).
#include <cuda.h>

#include <stdlib.h>

#include <mutex>

#include "common.h"
#include "conv_plan.h"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace vsb {

namespace {
constexpr int kBlockM = 128;
constexpr int kEW = 8;  // epilogue warps per CTA
constexpr int kMaxEpiBufs2 = 4;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p`'s counterpart in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// loads whose completion bytes land on a barrier of the pair's leader (bar_cluster = mapa(bar, 0))
__device__ __forceinline__ void tma2_load_2d(void* smem, const void* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_im2col_5d(void* smem, const void* map, uint32_t bar_cluster, int c, int w,
                                                    int h, int d, int n, uint16_t ow, uint16_t oh, uint16_t od) {
  asm volatile(
      "cp.async.bulk.tensor.5d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2], {%8, %9, %10};" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c), "r"(w), "r"(h), "r"(d), "r"(n), "h"(ow),
      "h"(oh), "h"(od)
      : "memory");
}
__device__ __forceinline__ void tmem2_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem2_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem2_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs once the MMAs issued so far retire
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
}  // namespace

// KK = kchunk / 16: 64- or 128-byte swizzled rows, KK MMAs of K = 16 per chunk, 4 / KK chunks per ring stage.
template <int KK>
__global__ void __launch_bounds__((kEW + 2) * 32, 1)
conv_igemm2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_res,
                   const __grid_constant__ CUtensorMap map_a2, const IgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kProducerWarp = kEW, kMmaWarp = kEW + 1;
  pdl_launch_dependents();  // the next kernel's prologue may overlap this kernel (ptx.cuh)
  constexpr uint32_t row_bytes = KK * 32;
  constexpr int kchunk = KK * 16;
  constexpr uint32_t a_chunk_bytes = kBlockM * row_bytes;
  const uint32_t half_n = p.block_n >> 1;
  const uint32_t b_chunk_bytes = half_n * row_bytes;  // this CTA's half of the weight rows
  const uint32_t stage_bytes = p.stage_bytes;
  const uint32_t epi_row_bytes = p.epi_n * 2;
  uint8_t* epi_buf = smem + p.off_epi;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full = empty_bar + p.stages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* epi_ready = tmem_empty + 2;
  uint64_t* bres_bar = epi_ready + kEW * kMaxEpiBufs2;  // resident weight halves of the pair have landed (leader's copy)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres_bar + 1);
  uint8_t* bres = smem + p.off_bres;                    // this CTA's half of the resident weight block

  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int n_tile = pair % p.n_tiles;
  const int pm0 = pair / p.n_tiles, pm_step = npairs / p.n_tiles;
  const int m_tiles = p.total_tiles / p.n_tiles;
  const int pm_tiles = (m_tiles + 1) >> 1;  // 256-pixel tiles

  float2* sb_tab = reinterpret_cast<float2*>(smem + p.off_bar + 1024);
  for (int c = threadIdx.x; c < p.block_n; c += blockDim.x)
    sb_tab[c] = make_float2(p.scale ? p.scale[n_tile * p.block_n + c] : 1.f, p.bias[n_tile * p.block_n + c]);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&map_a);
    if (p.chunks1 < p.total_chunks) tma_prefetch_desc(&map_a2);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_out);
    if (p.has_residual) tma_prefetch_desc(&map_res);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);   // the leader's expect_tx arrival (+ the bytes of both CTAs)
      mbar_init(&empty_bar[s], 1);  // one multicast commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * kEW);  // (leader's copy is the one in use) every epilogue warp of the pair
    }
    for (int i = 0; i < kEW * kMaxEpiBufs2; ++i) mbar_init(&epi_ready[i], 1);
    mbar_init(bres_bar, 1);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    tmem2_alloc(tmem_slot, p.tmem_cols);
    tmem2_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers exist before anything is signalled across the pair
  __syncthreads();     // (CTA barrier as well: compute-sanitizer racecheck does not model barrier.cluster)
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kProducerWarp) {
    if (lane == 0) {
      // ------------------------------------------------------ TMA producer (one thread per CTA)
      constexpr int CPS = 4 / KK;  // chunks per ring stage: a stage always carries 64 K-elements (4 MMAs)
      const uint32_t chunk_tx = 2u * (a_chunk_bytes + b_chunk_bytes);  // bytes of BOTH CTAs per chunk
      const uint32_t b_off = CPS * a_chunk_bytes;
      const bool b_resident = p.b_resident != 0;
      if (b_resident) {
        // weight-stationary pair: each CTA fetches its half of the [block_n x K] block once; both halves' bytes
        // are counted on the leader's barrier (the leader issues the MMAs)
        const uint32_t bres_leader = mapa_u32(bres_bar, 0);
        if (leader) mbar_expect_tx(bres_bar, 2u * (uint32_t)p.total_chunks * b_chunk_bytes);
        for (int g = 0; g < p.total_chunks; ++g)
          tma2_load_2d(bres + g * b_chunk_bytes, &map_b, bres_leader, g * kchunk,
                       n_tile * p.block_n + (int)rank * (int)half_n);
      }
      pdl_wait();  // activations are the previous kernels' outputs
      const uint32_t chunk_tx_eff = b_resident ? 2u * a_chunk_bytes : chunk_tx;
      const int total_chunks = p.total_chunks, stages = p.stages;
      const int cin = p.cin, fkw = p.kw, fkh = p.kh, block_n = p.block_n;
      const int owo = p.wo, oho = p.ho, oto = p.to, sw = p.sw, sh = p.sh, st = p.st, lw = p.lw, lh = p.lh, lt = p.lt;
      const int chunks1 = p.chunks1, sw2 = p.sw2, sh2 = p.sh2, st2 = p.st2;
      int slot = 0;
      uint32_t parity = 1;
      for (int pm = pm0; pm < pm_tiles; pm += pm_step) {
        int mt = 2 * pm + (int)rank;
        if (mt >= m_tiles) mt = m_tiles - 1;  // odd tile count: the peer recomputes the last tile (not stored)
        if (p.reverse) mt = m_tiles - 1 - mt;
        const int m0 = mt * kBlockM;
        const int wo = m0 % owo;
        const int r1 = m0 / owo;
        const int ho = r1 % oho;
        const int r2 = r1 / oho;
        const int to_ = r2 % oto;
        const int n0 = r2 / oto;
        const int w0 = wo * sw + lw, h0 = ho * sh + lh, d0 = to_ * st + lt;
        const int w2 = wo * sw2, h2 = ho * sh2, d2 = to_ * st2;
        const int ncol = n_tile * block_n + (int)rank * (int)half_n;
        int cc = 0, kw_ = 0, kh_ = 0, kt_ = 0, kcoord = 0;
        int left1 = chunks1;
        int left = total_chunks;
        while (left > 0) {
          const int nch = left < CPS ? left : CPS;
          left -= nch;
          mbar_wait(&empty_bar[slot], parity);
          if (leader) mbar_expect_tx(&full_bar[slot], nch * chunk_tx_eff);
          const uint32_t full_leader = mapa_u32(&full_bar[slot], 0);
          uint8_t* a_dst = smem + (uint32_t)slot * stage_bytes;
          uint8_t* b_dst = a_dst + b_off;
          for (int c = 0; c < nch; ++c) {
            if (left1 > 0) {
              tma2_load_im2col_5d(a_dst, &map_a, full_leader, cc, w0, h0, d0, n0, (uint16_t)kw_, (uint16_t)kh_,
                                  (uint16_t)kt_);
              if (--left1 == 0) cc = -kchunk;
            } else {
              tma2_load_im2col_5d(a_dst, &map_a2, full_leader, cc, w2, h2, d2, n0, 0, 0, 0);
            }
            if (!b_resident) tma2_load_2d(b_dst, &map_b, full_leader, kcoord, ncol);
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
    if (leader) {
      // ------------------------------------------------------ MMA issuer (leader CTA only)
      const uint64_t desc_hi = umma_smem_desc(0, row_bytes) & 0xFFFFFFFF00000000ull;
      const uint32_t desc_lo_flags = (uint32_t)(umma_smem_desc(0, row_bytes) & 0xFFFFC000ull);
      const uint32_t smem_lo = (smem_u32(smem) & 0x3FFFFu) >> 4;
      constexpr int CPS = 4 / KK;
      const uint32_t stage_lo = stage_bytes >> 4, a_chunk_lo = a_chunk_bytes >> 4, b_chunk_lo = b_chunk_bytes >> 4;
      const uint32_t b_off_lo = (CPS * a_chunk_bytes) >> 4;
      const uint32_t idesc = p.idesc;
      const int total_chunks = p.total_chunks, stages = p.stages, block_n = p.block_n;
      int slot = 0, tcount = 0;
      uint32_t parity = 0, a_slot_lo = smem_lo;
      const bool b_resident = p.b_resident != 0;
      const uint32_t bres_lo = (smem_u32(bres) & 0x3FFFFu) >> 4;
      if (b_resident) mbar_wait(bres_bar, 0);
      for (int pm = pm0; pm < pm_tiles; pm += pm_step, ++tcount) {
        const int acc = tcount & 1;
        mbar_wait(&tmem_empty[acc], ((tcount >> 1) & 1) ^ 1);  // both CTAs drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * block_n;
        for (int g = 0; g < total_chunks; g += CPS) {
          const int nch = total_chunks - g < CPS ? total_chunks - g : CPS;
          mbar_wait(&full_bar[slot], parity);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int c = 0; c < CPS; ++c) {
              if (c < nch) {
#pragma unroll
                for (int k = 0; k < KK; ++k) {
                  const uint64_t adesc = desc_hi | (uint64_t)(desc_lo_flags | (a_slot_lo + c * a_chunk_lo + 2 * k));
                  const uint32_t b_lo = b_resident ? bres_lo + (uint32_t)(g + c) * b_chunk_lo
                                                   : a_slot_lo + b_off_lo + c * b_chunk_lo;
                  const uint64_t bdesc = desc_hi | (uint64_t)(desc_lo_flags | (b_lo + 2 * k));
                  umma2_bf16(tmem_d, adesc, bdesc, idesc, (g | c | k) != 0 ? 1u : 0u);
                }
              }
            }
            umma2_commit_both(&empty_bar[slot]);
            if (g + CPS >= total_chunks) umma2_commit_both(&tmem_full[acc]);
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
    }
  } else {
    // ---------------------------------------------------------- epilogue (warps 0 .. 7 of each CTA)
    const int quarter = warp & 3, grp = warp >> 2;
    const uint32_t swz_mask = epi_row_bytes == 128 ? 7u : (epi_row_bytes == 64 ? 3u : 1u);
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t slab_bytes = 32 * epi_row_bytes;
    const int nb = p.epi_bufs;
    uint8_t* my_bufs = epi_buf + (size_t)warp * nb * slab_bytes;
    uint64_t* my_ready = epi_ready + warp * kMaxEpiBufs2;
    const int row_in_tile = quarter * 32;
    const float relu_floor = p.relu ? 0.f : -__int_as_float(0x7f800000);
    const int nbase = n_tile * p.block_n;
    const uint32_t sb_s = smem_u32(sb_tab);
    const int epi_n = p.epi_n, epi_chunks = p.epi_chunks;
    const bool has_res = p.has_residual != 0;
    const uint32_t empty_leader[2] = {mapa_u32(&tmem_empty[0], 0), mapa_u32(&tmem_empty[1], 0)};
    // prefetch cursor (lane 0): next chunk of THIS warp whose staging slab has not been armed yet
    int pf_pm = pm0, pf_chunk = 0, pf_gq = 0, pf_b = 0;
    auto pf_skip = [&]() {
      while (pf_pm < pm_tiles && (pf_gq & 1) != grp) {
        ++pf_gq;
        if (++pf_chunk == epi_chunks) {
          pf_chunk = 0;
          pf_pm += pm_step;
        }
      }
    };
    auto arm_next = [&]() {
      const int mt = 2 * pf_pm + (int)rank;
      if (has_res && mt < m_tiles) {
        mbar_expect_tx(&my_ready[pf_b], slab_bytes);
        tma_load_2d(my_bufs + pf_b * slab_bytes, &map_res, &my_ready[pf_b], nbase + pf_chunk * epi_n,
                    (p.reverse ? m_tiles - 1 - mt : mt) * kBlockM + row_in_tile);
      } else {
        mbar_arrive(&my_ready[pf_b]);
      }
      if (++pf_b == nb) pf_b = 0;
      ++pf_gq;
      if (++pf_chunk == epi_chunks) {
        pf_chunk = 0;
        pf_pm += pm_step;
      }
      pf_skip();
    };
    pdl_wait();  // residual loads read, and the stores overwrite, memory the previous kernel may still be using
    if (lane == 0) {
      pf_skip();
      for (int i = 0; i < nb - 1 && pf_pm < pm_tiles; ++i) arm_next();
    }
    __syncwarp();
    int b = 0;
    uint32_t bpar = 0;
    int gq = 0, tcount = 0;
    for (int pm = pm0; pm < pm_tiles; pm += pm_step, ++tcount) {
      const int mt = 2 * pm + (int)rank;
      const bool valid = mt < m_tiles;
      const int m0 = (p.reverse ? m_tiles - 1 - mt : mt) * kBlockM;
      const int acc = tcount & 1;
      mbar_wait(&tmem_full[acc], (tcount >> 1) & 1);
      tc_fence_after();
      for (int c = 0; c < epi_chunks; ++c, ++gq) {
        if ((gq & 1) != grp) continue;
        uint8_t* buf = my_bufs + b * slab_bytes;
        mbar_wait(&my_ready[b], bpar);
        const int col0 = c * epi_n;
        if (valid) {
          const uint32_t taddr = lane_taddr + acc * p.block_n + col0;
          if (has_res)
            epi_convert_chunk<true>(taddr, epi_n, smem_u32(buf), epi_row_bytes, swz_mask, lane, sb_s + col0 * 8,
                                    relu_floor);
          else
            epi_convert_chunk<false>(taddr, epi_n, smem_u32(buf), epi_row_bytes, swz_mask, lane, sb_s + col0 * 8,
                                     relu_floor);
          fence_proxy_async_smem();
        }
        __syncwarp();
        if (lane == 0) {
          if (valid) {
            tma_store_2d(&map_out, buf, nbase + col0, m0 + row_in_tile);
            tma_store_commit();
          }
          if (pf_pm < pm_tiles) {
            tma_store_wait_read1();
            arm_next();
          }
        }
        __syncwarp();
        if (++b == nb) {
          b = 0;
          bpar ^= 1;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(empty_leader[acc]);
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA leaves (or frees its TMEM) while the other may still use it
  if (warp == kMmaWarp) {
    __syncwarp();
    tmem2_dealloc(tmem_base, p.tmem_cols);
  }
}

int igemm2_launch(const vsb_conv_plan* plan, cudaStream_t stream) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_igemm2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(conv_igemm2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  if (attr_err != cudaSuccess) {
    set_error("cudaFuncSetAttribute(conv_igemm2_kernel) failed: %s", cudaGetErrorString(attr_err));
    return VSB_ERR_CUDA;
  }
  cudaError_t e = plan->params.kchunk == 64
                      ? launch_pdl(conv_igemm2_kernel<4>, plan->grid, (kEW + 2) * 32, plan->smem_bytes, stream, 2, plan->map_a,
                                   plan->map_b, plan->map_out, plan->map_res, plan->map_a2, plan->params)
                      : launch_pdl(conv_igemm2_kernel<2>, plan->grid, (kEW + 2) * 32, plan->smem_bytes, stream, 2, plan->map_a,
                                   plan->map_b, plan->map_out, plan->map_res, plan->map_a2, plan->params);
  if (e != cudaSuccess) {
    set_error("launch of conv_igemm2_kernel failed: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return VSB_ERR_CUDA;
  }
  count_launch();
  return VSB_OK;
}

}  // namespace vsb
