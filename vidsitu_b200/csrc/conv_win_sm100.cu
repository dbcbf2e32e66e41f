// Conv3d (+ folded frozen BatchNorm, ReLU, residual add) with the input staged ONCE in shared
// memory: the "window" algorithm for convs with spatial taps and few channels.
//
// Reference ops replaced (same contract as conv_igemm_sm100.cu): nn.Conv3d + BatchNorm3d(eval) +
// ReLU of the stems (SlowFast/slowfast/models/stem_helper.py:157-178) and of the 1x3x3 `b` convs of
// the bottlenecks (resnet_helper.py:196-209), in their pixel-group restatement where cin < 64.
//
// Why: the im2col-mode kernel re-fetches every input byte through L2 once per filter tap (9x for a
// 3x3, 70x for the grouped 5x7x7 fast stem -- ncu: 83 % L2 throughput, tensor pipe waiting).  Here a
// tile's input rows are loaded once by tiled-mode TMA (out-of-bounds rows/columns zero-filled = conv
// padding) as rows of RP slots (slot = one GEMM row of cin channels, 32/64/128 bytes, hardware
// swizzle), so the window is one linear array of slots.  GEMM row i of the tile is output position
// (i / RP, i % RP), and filter tap (kh, kw) reads slot i + kh*RP + kw: every tap is the SAME 128-row
// UMMA operand with its descriptor start address advanced by a whole number of rows (and by a K
// slice inside the row when only part of a slot's channels carry non-zero weights).  The hardware
// applies the swizzle XOR to absolute address bits, so row-shifted descriptors are exact (pinned on
// B200 by tools/gpu_probe_umma.py).  Stride-2 convs (the stems) keep even and odd input rows in two
// sub-windows so that taps stay unit-stride.  Temporal taps (kt > 1) walk a ring of frame windows:
// consecutive output frames of a run share kt-1 of their kt windows, so each frame is fetched once.
//
// The whole weight matrix [cout x K] stays resident in shared memory.  Warp roles and the epilogue
// (TMEM double buffering, per-warp staging slabs, residual in / result out by TMA) follow
// conv_igemm_sm100.cu; output boxes are [epi_n channels x box_w slots x box_h rows], clipped by the
// TMA unit at the tensor edge (garbage GEMM rows beyond wo / ho never reach memory).
#include <cuda.h>

#include <stdlib.h>

#include <mutex>
#include <new>

#include "common.h"
#include "conv_plan.h"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace vsb {

// Optional role timeline (VSB_WIN_DEBUG=1 at plan time): cycles each role spends waiting, summed over CTAs.
//   0 producer: wait for a free window slot     1 MMA: wait for a free accumulator   2 MMA: wait for a window
//   3 MMA: issue loop                           4 epilogue warp 0: wait accumulator  5 epi: wait staging slab
//   6 epi: TMEM -> regs -> smem math            7 epi: wait for the previous store's smem read
//   8 whole CTA                                 9 tiles
#define WIN_T(idx, stmt)                                                   \
  do {                                                                     \
    if (kDbg) {                                                            \
      const long long _t0 = clock64();                                     \
      stmt;                                                                \
      dbg_acc[idx] += (uint32_t)(clock64() - _t0);                                     \
    } else {                                                               \
      stmt;                                                                \
    }                                                                      \
  } while (0)

namespace {
constexpr int kEpiWarps = 8;
constexpr int kProducerWarp = kEpiWarps;
constexpr int kMmaWarp = kEpiWarps + 1;
constexpr int kThreads = (kEpiWarps + 2) * 32;
constexpr int kMaxEpiBufs = 4;
constexpr int kMaxAcc = 16;  // TMEM accumulator ring slots (temporal-scatter mode uses all of them)
}  // namespace

template <bool kDbg>
__global__ void __launch_bounds__(kThreads, 2)
conv_win_kernel(const __grid_constant__ CUtensorMap map_in0, const __grid_constant__ CUtensorMap map_in1,
                const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_out,
                const __grid_constant__ CUtensorMap map_res, const __grid_constant__ WinParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  pdl_launch_dependents();  // the next kernel's prologue may overlap this kernel (ptx.cuh)

  const uint32_t epi_row_bytes = p.epi_n * 2;
  uint8_t* bres = smem + p.off_b;
  uint8_t* epi_buf = smem + p.off_epi;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full = empty_bar + p.stages;
  uint64_t* tmem_empty = tmem_full + kMaxAcc;
  uint64_t* epi_ready = tmem_empty + kMaxAcc;
  uint64_t* bres_bar = epi_ready + kEpiWarps * kMaxEpiBufs;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint32_t dbg_acc[13] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // (cycle counts of one CTA fit 32 bits)
  const long long dbg_start = kDbg ? clock64() : 0;

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&map_in0);
    if (p.nsub == 2) tma_prefetch_desc(&map_in1);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_out);
    if (p.has_residual) tma_prefetch_desc(&map_res);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < kMaxAcc; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], p.tsplit ? p.epi_warps >> 1 : p.epi_warps);
    }
    for (int i = 0; i < kEpiWarps * kMaxEpiBufs; ++i) mbar_init(&epi_ready[i], 1);
    mbar_init(bres_bar, 1);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  // per-K-step descriptor table of one frame: .x = low word of the A descriptor relative to the frame
  // window (flags | tap offset | K slice), .y = low word of the B descriptor relative to the frame's weights
  uint2* step_tab = reinterpret_cast<uint2*>(smem + p.off_tab);
  if (warp < 4) {
    const uint32_t a_flags = (uint32_t)(umma_smem_desc(0, p.row_bytes) & 0xFFFFC000ull);
    const uint32_t b_flags = (uint32_t)(umma_smem_desc(0, 128) & 0xFFFFC000ull);
    const uint32_t b_block_lo = p.b_block_bytes >> 4;
    for (int tap = threadIdx.x; tap < p.ntaps; tap += 128) {
      int first = 0;
      for (int k = 0; k < tap; ++k) first += p.tap_ks[k];
      const int ks = p.tap_ks[tap];
      for (int s2 = 0; s2 < ks; ++s2) {
        const int i = first + s2;
        step_tab[i] = make_uint2(a_flags | ((p.tap_aoff[tap] >> 4) + 2 * s2),
                                 b_flags | ((uint32_t)(i >> 2) * b_block_lo + 2 * (i & 3)));
      }
    }
  }
  float2* sb_tab = reinterpret_cast<float2*>(smem + p.off_tab + 2048);  // (scale, bias) per output channel
  for (int c = threadIdx.x; c < p.block_n; c += blockDim.x) sb_tab[c] = make_float2(p.scale[c], p.bias[c]);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int L = p.L, KT = p.kt, S = p.stages;

  if (warp == kProducerWarp) {
    if (lane == 0) {
      // ------------------------------------------------------ TMA producer (one thread)
      // frame kt's K range [kt*k_per_frame, +k_per_frame) starts on a block boundary in shared memory (a
      // block that runs past it just carries K columns no step refers to)
      mbar_expect_tx(bres_bar, p.b_blocks * p.b_block_bytes);
      if (p.tsc) {
        // temporal-scatter: K block j holds the weights of ALL temporal taps, kt descending = output frame
        // ascending, as kt x [block_n rows x 128 B] (one N = kt * block_n operand)
        const uint32_t tap_bytes = (uint32_t)p.block_n * 128u;
        for (int kt = 0; kt < KT; ++kt)
          for (int j = 0; j < p.b_blocks_per_frame; ++j)
            tma_load_2d(bres + j * p.b_block_bytes + (uint32_t)(KT - 1 - kt) * tap_bytes, &map_b, bres_bar,
                        kt * p.k_per_frame + j * 64, 0);
      } else
      for (int kt = 0, g = 0; kt < KT; ++kt)
        for (int j = 0; j < p.b_blocks_per_frame; ++j, ++g)
          tma_load_2d(bres + g * p.b_block_bytes, &map_b, bres_bar, kt * p.k_per_frame + j * 64, 0);
      pdl_wait();  // activations are the previous kernels' outputs (the resident weights above are constants)
      const CUtensorMap* m0 = p.sub_map[0] ? &map_in1 : &map_in0;
      const CUtensorMap* m1 = p.sub_map[1] ? &map_in1 : &map_in0;
      const int nframes = p.tsc ? p.t_in : L + KT - 1;  // temporal-scatter walks the real input frames only
      const int f_base = p.tsc ? 0 : -p.pt_lo;
      int slot = 0;
      uint32_t parity = 1;  // first pass over the ring: slots are free
      TileCursor cur;
      cur.init(blockIdx.x, gridDim.x, p.yb_count, L == 1 ? p.to : 1);
      for (; cur.run < p.total_runs; cur.next()) {
        const int t0 = L == 1 ? cur.t : 0, n = p.reverse ? p.n_clips - 1 - cur.n : cur.n;
        const int y0 = cur.yb * p.R;
        for (int fi = 0; fi < nframes; ++fi) {
          const int f = t0 + fi + f_base;  // input frame; outside [0, T) -> zero-filled window
          WIN_T(0, mbar_wait(&empty_bar[slot], parity));
          mbar_expect_tx(&full_bar[slot], p.stage_tx);
          uint8_t* dst = smem + (uint32_t)slot * p.stage_bytes;
          tma_load_5d(dst + p.sub_off[0], m0, &full_bar[slot], 0, -p.pw_lo, y0 + p.sub_hoff[0], f, n);
          if (p.nsub == 2)
            tma_load_5d(dst + p.sub_off[1], m1, &full_bar[slot], 0, -p.pw_lo, y0 + p.sub_hoff[1], f, n);
          if (++slot == S) {
            slot = 0;
            parity ^= 1;
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // -------------------------------------------------------- MMA issuer
    // Frame windows are consumed in the order they were produced; frame j of a run (j = tl + kt)
    // lives in ring slot (run_base + j) % S.  Tile tl releases frame tl after its kt = 0 taps (no
    // later tile reads it); the last kt-1 frames of a run are released after its last tile.
    const uint64_t a_hi = umma_smem_desc(0, p.row_bytes) & 0xFFFFFFFF00000000ull;
    const uint64_t b_hi = umma_smem_desc(0, 128) & 0xFFFFFFFF00000000ull;
    const uint32_t smem_lo = (smem_u32(smem) & 0x3FFFFu) >> 4;
    const uint32_t bres_lo = (smem_u32(bres) & 0x3FFFFu) >> 4;
    const uint32_t stage_lo = p.stage_bytes >> 4;
    const uint32_t b_frame_lo = (uint32_t)p.b_blocks_per_frame * (p.b_block_bytes >> 4);
    const uint32_t idesc = p.idesc;
    const int spf = p.steps_per_frame, block_n = p.block_n;
    int base_slot = 0;          // ring slot of frame 0 of the current tile
    uint32_t base_par = 0;
    int tcount = 0;
    const int amask = p.nacc - 1, ashift = p.nacc_shift;
    mbar_wait(bres_bar, 0);
    if (p.tsc) {
      // Temporal-scatter (kt > 1): input-stationary along t.  The window of input frame f is read ONCE
      // and multiplied by the weights of all temporal taps in a single N = nb * block_n MMA whose
      // column blocks are the accumulators of output frames t = f + pt_lo - kt (kt descending = t
      // ascending = consecutive slots of a ring of TMEM accumulators).  Versus the output-stationary
      // loop this divides the shared-memory A reads (the bound of N <= 96 MMAs: ~60 clk per fresh
      // 128 x 16 A slab whatever N) by kt, and the window ring only has to hide the load latency.
      const int SL = p.nacc, slmask = SL - 1;
      const int T_in = p.t_in, TO = p.to, ptl = p.pt_lo;
      const uint32_t tap_lo = ((uint32_t)block_n * 128u) >> 4;      // B row block of one temporal tap
      const uint32_t idesc0 = umma_idesc_bf16(128, 0), idesc_blk = ((uint32_t)block_n >> 3) << 17;
      int slot = 0;
      uint32_t par = 0;
      int g0 = 0;  // global tile index of output frame 0 of the run (accumulator ring position)
      for (int run = blockIdx.x; run < p.total_runs; run += gridDim.x, g0 += TO) {
        for (int f = 0; f < T_in; ++f) {
          const int t_top = f + ptl;
          const int t_lo = t_top - (KT - 1) > 0 ? t_top - (KT - 1) : 0;
          const int t_hi = t_top < TO - 1 ? t_top : TO - 1;
          const bool top_fresh = f > 0 && t_top <= TO - 1;  // block t_top is written for the first time
          // accumulators written for the first time must have been drained by the epilogue
          if (f == 0) {
            for (int t = t_lo; t <= t_hi; ++t)
              WIN_T(1, mbar_wait(&tmem_empty[(g0 + t) & slmask], (((g0 + t) >> ashift) & 1) ^ 1));
          } else if (top_fresh) {
            WIN_T(1, mbar_wait(&tmem_empty[(g0 + t_top) & slmask], (((g0 + t_top) >> ashift) & 1) ^ 1));
          }
          WIN_T(2, mbar_wait(&full_bar[slot], par));
          tc_fence_after();
          const long long issue_t0 = kDbg ? clock64() : 0;
          if (elect_one() && t_lo <= t_hi) {
            const uint32_t a_slot = smem_lo + (uint32_t)slot * stage_lo;
            const int nblk = t_hi - t_lo + 1;
            const int s0 = (g0 + t_lo) & slmask;
            const int n1 = nblk < SL - s0 ? nblk : SL - s0;  // blocks before the ring wraps
            const int n2 = nblk - n1;                        // blocks from slot 0 on
            const int rb0 = (KT - 1) - t_top + t_lo;         // B row block of output frame t_lo
            const uint32_t d1 = tmem_base + (uint32_t)s0 * block_n, d2 = tmem_base;
            const uint32_t acc_all = f == 0 ? 0u : 1u;
            // step 0: the fresh top block (if any) must overwrite, the others accumulate
            {
              const uint2 e = step_tab[0];
              const uint64_t ad = a_hi | (uint64_t)(e.x + a_slot);
              const uint32_t b0 = e.y + bres_lo + (uint32_t)rb0 * tap_lo;
              int m1 = n1, m2 = n2;
              if (top_fresh) (n2 ? m2 : m1) -= 1;
              if (m1) umma_bf16(d1, ad, b_hi | (uint64_t)b0, idesc0 + m1 * idesc_blk, acc_all);
              if (m2) umma_bf16(d2, ad, b_hi | (uint64_t)(b0 + n1 * tap_lo), idesc0 + m2 * idesc_blk, acc_all);
              if (top_fresh)
                umma_bf16(tmem_base + (uint32_t)((g0 + t_top) & slmask) * block_n, ad,
                          b_hi | (uint64_t)(b0 + (nblk - 1) * tap_lo), idesc0 + idesc_blk, 0u);
            }
            const uint32_t i1 = idesc0 + n1 * idesc_blk, i2 = idesc0 + n2 * idesc_blk;
            if (n2 == 0) {
#pragma unroll 4
              for (int i = 1; i < spf; ++i) {
                const uint2 e = step_tab[i];
                umma_bf16(d1, a_hi | (uint64_t)(e.x + a_slot),
                          b_hi | (uint64_t)(e.y + bres_lo + (uint32_t)rb0 * tap_lo), i1, 1u);
              }
            } else {
#pragma unroll 2
              for (int i = 1; i < spf; ++i) {
                const uint2 e = step_tab[i];
                const uint64_t ad = a_hi | (uint64_t)(e.x + a_slot);
                const uint32_t b0 = e.y + bres_lo + (uint32_t)rb0 * tap_lo;
                umma_bf16(d1, ad, b_hi | (uint64_t)b0, i1, 1u);
                umma_bf16(d2, ad, b_hi | (uint64_t)(b0 + n1 * tap_lo), i2, 1u);
              }
            }
            umma_commit(&empty_bar[slot]);
            // output frames whose last contributing input frame this was
            const int t_done = t_top - (KT - 1);
            if (f == T_in - 1) {
              for (int t = t_done > 0 ? t_done : 0; t < TO; ++t) umma_commit(&tmem_full[(g0 + t) & slmask]);
            } else if (t_done >= 0 && t_done < TO) {
              umma_commit(&tmem_full[(g0 + t_done) & slmask]);
            }
          }
          __syncwarp();
          if (kDbg) dbg_acc[3] += (uint32_t)(clock64() - issue_t0);
          if (++slot == S) {
            slot = 0;
            par ^= 1;
          }
        }
      }
    } else if (p.pair) {
      // Two tiles at a time (KT == 1, one frame window each): their MMAs alternate, so two independent
      // accumulation chains are in flight and every weight K slice is used twice back to back.
      int slot = 0;
      uint32_t par = 0;
      const int step = gridDim.x;
      for (int run = blockIdx.x; run < p.total_runs; run += 2 * step) {
        const bool two = run + step < p.total_runs;
        const int acc0 = tcount & amask, acc1 = (tcount + 1) & amask;
        const int slot0 = slot;
        const uint32_t par0 = par;
        if (++slot == S) { slot = 0; par ^= 1; }
        const int slot1 = slot;
        const uint32_t par1 = par;
        if (two && ++slot == S) { slot = 0; par ^= 1; }
        WIN_T(1, mbar_wait(&tmem_empty[acc0], ((tcount >> ashift) & 1) ^ 1));
        if (two) WIN_T(1, mbar_wait(&tmem_empty[acc1], (((tcount + 1) >> ashift) & 1) ^ 1));
        WIN_T(2, mbar_wait(&full_bar[slot0], par0));
        if (two) WIN_T(2, mbar_wait(&full_bar[slot1], par1));
        const long long issue_t0 = kDbg ? clock64() : 0;
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a0 = smem_lo + (uint32_t)slot0 * stage_lo, a1 = smem_lo + (uint32_t)slot1 * stage_lo;
          const uint32_t d0 = tmem_base + acc0 * block_n, d1 = tmem_base + acc1 * block_n;
          if (two) {
#pragma unroll 2
            for (int i = 0; i < spf; ++i) {
              const uint2 e = step_tab[i];
              const uint64_t bdesc = b_hi | (uint64_t)(e.y + bres_lo);
              umma_bf16(d0, a_hi | (uint64_t)(e.x + a0), bdesc, idesc, i != 0 ? 1u : 0u);
              umma_bf16(d1, a_hi | (uint64_t)(e.x + a1), bdesc, idesc, i != 0 ? 1u : 0u);
            }
          } else {
#pragma unroll 4
            for (int i = 0; i < spf; ++i) {
              const uint2 e = step_tab[i];
              umma_bf16(d0, a_hi | (uint64_t)(e.x + a0), b_hi | (uint64_t)(e.y + bres_lo), idesc, i != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[slot0]);
          umma_commit(&tmem_full[acc0]);
          if (two) {
            umma_commit(&empty_bar[slot1]);
            umma_commit(&tmem_full[acc1]);
          }
        }
        __syncwarp();
        if (kDbg) dbg_acc[3] += (uint32_t)(clock64() - issue_t0);
        tcount += two ? 2 : 1;
      }
    } else
    for (int run = blockIdx.x; run < p.total_runs; run += gridDim.x) {
      for (int tl = 0; tl < L; ++tl, ++tcount) {
        const int acc = tcount & amask;
        WIN_T(1, mbar_wait(&tmem_empty[acc], ((tcount >> ashift) & 1) ^ 1));
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * block_n;
        int slot = base_slot;
        uint32_t par = base_par;
        for (int kt = 0; kt < KT; ++kt) {
          WIN_T(2, mbar_wait(&full_bar[slot], par));
          tc_fence_after();
          const long long issue_t0 = kDbg ? clock64() : 0;
          if (elect_one()) {
            const uint32_t a_slot = smem_lo + (uint32_t)slot * stage_lo;
            const uint32_t b_frame = bres_lo + (uint32_t)kt * b_frame_lo;  // this frame's weights
#pragma unroll 4
            for (int i = 0; i < spf; ++i) {
              const uint2 e = step_tab[i];
              umma_bf16(tmem_d, a_hi | (uint64_t)(e.x + a_slot), b_hi | (uint64_t)(e.y + b_frame), idesc,
                        (kt | i) != 0 ? 1u : 0u);
            }
            if (kt == 0) umma_commit(&empty_bar[slot]);       // frame tl is done for this run
            if (kt == KT - 1) umma_commit(&tmem_full[acc]);  // accumulator complete
          }
          __syncwarp();
          if (kDbg) dbg_acc[3] += (uint32_t)(clock64() - issue_t0);
          if (++slot == S) {
            slot = 0;
            par ^= 1;
          }
        }
        if (++base_slot == S) {
          base_slot = 0;
          base_par ^= 1;
        }
      }
      // release the trailing kt-1 frames of the run (every MMA that read them was issued above)
      for (int k = 1; k < KT; ++k) {
        if (elect_one()) umma_commit(&empty_bar[base_slot]);
        __syncwarp();
        if (++base_slot == S) {
          base_slot = 0;
          base_par ^= 1;
        }
      }
    }
  } else if (warp < p.epi_warps) {
    // ---------------------------------------------------------- epilogue
    // warp w owns TMEM lanes / GEMM rows [32*(w&3), +32) = box_h image rows x box_w slots and, when
    // there are 8 epilogue warps, every second column chunk (chunk parity == w>>2).
    const int quarter = warp & 3, grp = warp >> 2;
    const int ngrp = p.epi_warps >> 2;  // 1 or 2
    const uint32_t swz_mask = epi_row_bytes == 128 ? 7u : (epi_row_bytes == 64 ? 3u : 1u);
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t slab_bytes = 32 * epi_row_bytes;
    const int nb = p.epi_bufs;
    uint8_t* my_bufs = epi_buf + (size_t)warp * nb * slab_bytes;
    uint64_t* my_ready = epi_ready + warp * kMaxEpiBufs;
    const int row0 = quarter * 32;
    const int bx0 = row0 & (p.RP - 1), by0 = row0 >> p.rp_shift;
    const int chunks = p.epi_chunks;
    const int tdim = L == 1 ? p.to : 1;
    const float relu_floor = p.relu ? 0.f : -__int_as_float(0x7f800000);
    const uint32_t sb_s = smem_u32(sb_tab);
    // prefetch cursor (lane 0): next chunk of THIS warp whose staging slab has not been armed yet
    // Work split between the two groups of four epilogue warps: column chunks of every tile (chunk parity ==
    // group), or - tile split, for narrow tiles (<= 2 chunks) whose epilogue is one latency chain per warp -
    // whole tiles (tile parity == group), so that two tiles drain concurrently.
    const bool tsplit = p.tsplit != 0;
    TileCursor pf;
    pf.init(blockIdx.x, gridDim.x, p.yb_count, tdim);
    int pf_tl = 0, pf_chunk = (ngrp == 2 && !tsplit) ? grp : 0, pf_q = 0;
    const int chunk_step = tsplit ? 1 : ngrp;  // this warp's chunks: grp, grp + ngrp, ... (all of them when tile-split)
    const int chunk0 = (ngrp == 2 && !tsplit) ? grp : 0;
    auto pf_next_tile = [&]() {
      if (++pf_tl == L) {
        pf_tl = 0;
        pf.next();
      }
    };
    if (tsplit && grp == 1) pf_next_tile();  // group 1 starts at the second tile
    auto arm_next = [&]() {
      const int bsel = pf_q % nb;
      if (p.has_residual) {
        const int tn = (p.reverse ? p.n_clips - 1 - pf.n : pf.n) * p.to + (L == 1 ? pf.t : pf_tl);
        mbar_expect_tx(&my_ready[bsel], slab_bytes);
        tma_load_4d(my_bufs + bsel * slab_bytes, &map_res, &my_ready[bsel], pf_chunk * p.epi_n, bx0,
                    pf.yb * p.R + by0, tn);
      } else {
        mbar_arrive(&my_ready[bsel]);
      }
      ++pf_q;
      pf_chunk += chunk_step;
      if (pf_chunk >= chunks) {
        pf_chunk = chunk0;
        pf_next_tile();
        if (tsplit) pf_next_tile();  // skip the other group's tile
      }
    };
    pdl_wait();  // residual loads read, and the stores overwrite, memory the previous kernel may still be using
    if (lane == 0) {
      for (int i = 0; i < nb - 1 && pf.run < p.total_runs; ++i) arm_next();
    }
    __syncwarp();
    int q = 0, tcount = 0;
    TileCursor cur;
    cur.init(blockIdx.x, gridDim.x, p.yb_count, tdim);
    for (; cur.run < p.total_runs; cur.next()) {
      for (int tl = 0; tl < L; ++tl, ++tcount) {
        const int y0 = cur.yb * p.R;
        const int tn = (p.reverse ? p.n_clips - 1 - cur.n : cur.n) * p.to + (L == 1 ? cur.t : tl);
        if (tsplit && (tcount & 1) != grp) continue;  // the other group's tile
        const int acc = tcount & (p.nacc - 1);
        WIN_T(4, mbar_wait(&tmem_full[acc], (tcount >> p.nacc_shift) & 1));
        if (kDbg) dbg_acc[9] += 1;
        tc_fence_after();
        for (int c = chunk0; c < chunks; c += chunk_step) {
          const int b = q % nb;
          uint8_t* buf = my_bufs + b * slab_bytes;
          WIN_T(5, mbar_wait(&my_ready[b], (q / nb) & 1));  // slab free (+ residual landed)
          const long long math_t0 = kDbg ? clock64() : 0;
          const int col0 = c * p.epi_n;
          const uint32_t taddr = lane_taddr + acc * p.block_n + col0;
          if (p.has_residual)
            epi_convert_chunk<true>(taddr, p.epi_n, smem_u32(buf), epi_row_bytes, swz_mask, lane, sb_s + col0 * 8,
                                    relu_floor);
          else
            epi_convert_chunk<false>(taddr, p.epi_n, smem_u32(buf), epi_row_bytes, swz_mask, lane, sb_s + col0 * 8,
                                     relu_floor);
          const long long fence_t0 = kDbg ? clock64() : 0;
          fence_proxy_async_smem();
          __syncwarp();
          if (kDbg) dbg_acc[6] += (uint32_t)(clock64() - math_t0);
          if (kDbg) dbg_acc[10] += (uint32_t)(clock64() - fence_t0);
          if (lane == 0) {
            tma_store_4d(&map_out, buf, col0, bx0, y0 + by0, tn);  // clipped at wo / ho by the TMA unit
            tma_store_commit();
            if (pf.run < p.total_runs) {
              WIN_T(7, tma_store_wait_read1());
              arm_next();
            }
          }
          __syncwarp();
          ++q;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      }
    }
    if (lane == 0) tma_store_wait_all();
  }

  if (kDbg && lane == 0 && (warp == kProducerWarp || warp == kMmaWarp || warp == 0)) {
    if (warp == 0) dbg_acc[8] = (uint32_t)(clock64() - dbg_start);
    for (int i = 0; i < 13; ++i)
      if (dbg_acc[i]) atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg) + i, (unsigned long long)dbg_acc[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    __syncwarp();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------ host side

static int encode_tiled_nd(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims,
                           const cuuint64_t* strides_bytes, const cuuint32_t* box, CUtensorMapSwizzle swz,
                           const char* what) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims,
                              strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(%s) failed (CUresult %d): rank %d dims %llu %llu %llu box %u %u %u", what, (int)r,
              rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2], box[0],
              box[1], box[2]);
    return VSB_ERR_CUDA;
  }
  return VSB_OK;
}

int win_plan_build(vsb_conv_plan* plan, const vsb_conv_desc* d, int to, int ho, int wo) {
  // ---- domain of the algorithm (anything else: the caller uses the im2col kernel)
  if (d->dtype != VSB_BF16) return 1;
  if (d->st != 1 || d->sw != 1 || (d->sh != 1 && d->sh != 2)) return 1;
  const int rb = d->cin * 2;
  if (rb != 32 && rb != 64 && rb != 128) return 1;
  if (d->cout % 16 || d->cout < 16 || d->cout > 256) return 1;
  if (d->kw > 8 || d->kh > 8 || d->kh * d->kw > kWinMaxTaps) return 1;
  if (d->in_pitch % 8 || d->out_pitch % 8 || (d->residual && d->res_pitch % 8)) return 1;
  if (wo > 128 || d->pw_lo < 0 || d->ph_lo < 0 || d->pt_lo < 0) return 1;
  // per-tap channel ranges
  int c_lo[8], c_hi[8], ksum = 0;
  for (int k = 0; k < d->kw; ++k) {
    c_lo[k] = d->kw_c_hi[k] ? d->kw_c_lo[k] : 0;
    c_hi[k] = d->kw_c_hi[k] ? d->kw_c_hi[k] : d->cin;
    if (c_lo[k] % 16 || c_hi[k] % 16 || c_lo[k] < 0 || c_hi[k] > d->cin || c_hi[k] <= c_lo[k]) {
      set_error("bad channel range [%d, %d) of tap kw=%d", c_lo[k], c_hi[k], k);
      return VSB_ERR_INVALID;
    }
    ksum += c_hi[k] - c_lo[k];
  }
  const long long k_total = (long long)d->kt * d->kh * ksum;
  // window-row pitch: power of two >= wo such that taps running past the row end land in the next
  // row's left padding (zeros), see header comment
  int RP = 8;
  while (RP < wo || wo + d->kw - 2 >= RP + d->pw_lo) RP <<= 1;
  if (RP > 128) return 1;
  const bool wraps = wo + d->kw - 2 >= RP;
  int rp_shift = 0;
  while ((1 << rp_shift) < RP) ++rp_shift;
  const int R = 128 / RP;
  // sub-windows
  WinParams& p = plan->win;
  p.nsub = d->sh;
  int rows[2] = {0, 0};
  if (d->sh == 1) {
    rows[0] = R - 1 + d->kh + (wraps ? 1 : 0);
    p.sub_map[0] = 0;
    p.sub_hoff[0] = -d->ph_lo;
  } else {
    for (int q = 0; q < 2; ++q) {
      const int cnt = (d->kh - q + 1) / 2;  // taps kh = q, q+2, ...
      if (cnt <= 0) return 1;
      rows[q] = R - 1 + cnt + (wraps ? 1 : 0);
      const int par = (q + d->ph_lo) & 1;  // parity of the sub-window's input rows
      p.sub_map[q] = par;
      p.sub_hoff[q] = (q - d->ph_lo - par) / 2;  // exact: the numerator is even
    }
  }
  if (rows[0] > 256 || rows[1] > 256) return 1;
  const uint32_t sub_bytes0 = (uint32_t)rows[0] * RP * rb, sub_bytes1 = (uint32_t)rows[1] * RP * rb;
  p.sub_off[0] = 0;
  p.sub_off[1] = (sub_bytes0 + 1023) & ~1023u;
  const uint32_t stage_bytes = (p.sub_off[1] + sub_bytes1 + 1023) & ~1023u;
  // ---- shared-memory budget
  const int block_n = d->cout;
  const int k_per_frame = d->kh * ksum;
  const int b_blocks_per_frame = (k_per_frame + 63) / 64;
  const int b_blocks = d->kt * b_blocks_per_frame;
  const long long b_bytes = (long long)b_blocks * block_n * 128;
  // temporal-scatter mode (see the MMA issuer): all temporal taps in one N = kt * block_n MMA, output
  // frames in a power-of-two ring of TMEM accumulators
  static const bool no_tsc_env = getenv("VSB_WIN_NO_TSC") != nullptr;
  const bool no_tsc = no_tsc_env || (d->flags & VSB_PLAN_NO_TSCATTER);
  int acc_slots = 512 / block_n > kMaxAcc ? kMaxAcc : 512 / block_n;
  const bool tsc = !no_tsc && d->kt > 1 && d->kt * block_n <= 256 && (block_n & (block_n - 1)) == 0 &&
                   acc_slots >= d->kt + 2 && d->pt_lo < d->kt && d->pt_hi < d->kt;
  const uint32_t b_block_bytes = (uint32_t)block_n * 128 * (tsc ? d->kt : 1);
  // Epilogue work split between the two groups of four warps.  Tiles of <= 64 columns: the groups take alternate
  // TILES, each tile one 64-column chunk (measured on B200: fast-pathway `b` convs 0.079 -> 0.063 ms, the chain
  // TMEM -> registers -> smem -> TMA store of a narrow tile is one latency-bound sequence per warp, so two tiles
  // in flight beat two half-tiles).  Wider tiles: alternate 32-column chunks of every tile.
  // (not in temporal-scatter mode: the fast stem is MMA-bound and has no shared memory left for 8 warps' slabs)
  static const bool no_tsplit_env = getenv("VSB_WIN_NO_TSPLIT") != nullptr;
  const bool want_tsplit = !no_tsplit_env && !(d->flags & VSB_PLAN_NO_TILE_SPLIT) && !tsc;
  int epi_n = block_n >= 64 ? 32 : block_n;
  if (want_tsplit && block_n == 64 && d->epi_n != 32) epi_n = 64;
  if (d->epi_n && block_n >= 2 * d->epi_n && block_n % d->epi_n == 0) epi_n = d->epi_n;  // caller's tuning
  if (block_n % epi_n) epi_n = 16;
  const int epi_chunks = block_n / epi_n;
  const bool tsplit = want_tsplit && epi_chunks == 1;
  const int epi_warps = (epi_chunks >= 2 || tsplit) ? 8 : 4;
  if (epi_warps == 8 && !tsplit && (epi_chunks & 1)) return 1;  // chunk parity split needs an even chunk count
  int epi_bufs = d->residual ? 3 : 2;
  if (d->epi_bufs >= 2 && d->epi_bufs <= kMaxEpiBufs) epi_bufs = d->epi_bufs;
  const int bar_bytes = 1024 + 2048 + 2048;  // barriers + per-step descriptor table + (scale, bias) pairs
  // the tensor-core operand read of the last stage runs up to 128 rows past the largest tap offset
  const uint32_t max_tap_rows = (uint32_t)((d->sh == 1 ? (d->kh - 1) : (d->kh - 1) / 2) * RP + d->kw - 1);
  const uint32_t overrun = (128 + max_tap_rows) * rb;  // bytes past a sub-window start that a descriptor may touch
  (void)overrun;  // always inside the dynamic allocation: the weight / staging regions follow the ring
  int stages = 0;
  size_t smem_bytes = 0;
  for (;;) {
    const long long fixed = ((b_bytes + 1023) & ~1023ll) + (long long)epi_warps * epi_bufs * 32 * epi_n * 2 +
                            bar_bytes + 1024;
    int want = tsc ? 3 : (d->kt > 1 ? d->kt + 3 : 6);
    if (d->stages > want) want = d->stages;  // caller's tuning may deepen the ring (room permitting)
    // two co-resident CTAs per SM (two independent MMA chains) when a >= 4-stage ring fits in half the SM
    long long room = 113 * 1024 - fixed;
    if (d->kt > 1 || room < 4ll * stage_bytes) room = 227 * 1024 - fixed;
    stages = room > 0 ? (int)(room / stage_bytes) : 0;
    if (stages > want) stages = want;
    if (d->stages && stages > d->stages) stages = d->stages;
    const int need = tsc ? 2 : (d->kt > 1 ? d->kt + 1 : 2);
    if (stages >= need) {
      smem_bytes = (size_t)stages * stage_bytes + (size_t)fixed;
      break;
    }
    if (epi_bufs > 2) {
      epi_bufs = 2;
      continue;
    }
    return 1;  // does not fit: weights not resident-able next to the window ring
  }
  if (stages > 16) stages = 16;

  int rc = load_driver_entry_points();
  if (rc != VSB_OK) return rc;
  // ---- tensor maps
  const CUtensorMapSwizzle a_swz = swizzle_for(rb);
  if (d->sh == 1) {
    cuuint64_t dims[5] = {(cuuint64_t)d->cin, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->t, (cuuint64_t)d->n};
    cuuint64_t str[4] = {(cuuint64_t)d->in_pitch * 2, (cuuint64_t)d->w * d->in_pitch * 2,
                         (cuuint64_t)d->h * d->w * d->in_pitch * 2, (cuuint64_t)d->t * d->h * d->w * d->in_pitch * 2};
    cuuint32_t box[5] = {(cuuint32_t)d->cin, (cuuint32_t)RP, (cuuint32_t)rows[0], 1, 1};
    rc = encode_tiled_nd(&plan->map_a, d->in, 5, dims, str, box, a_swz, "window input");
    if (rc != VSB_OK) return rc;
    plan->map_a1 = plan->map_a;
  } else {
    for (int par = 0; par < 2; ++par) {
      const int hp = (d->h - par + 1) / 2;  // rows of this parity
      int q = -1;
      for (int qq = 0; qq < 2; ++qq)
        if (p.sub_map[qq] == par) q = qq;
      CUtensorMap* m = par ? &plan->map_a1 : &plan->map_a;
      if (q < 0 || hp <= 0) {
        *m = par ? plan->map_a : plan->map_a1;
        continue;
      }
      const uint8_t* base = static_cast<const uint8_t*>(d->in) + (size_t)par * d->w * d->in_pitch * 2;
      cuuint64_t dims[5] = {(cuuint64_t)d->cin, (cuuint64_t)d->w, (cuuint64_t)hp, (cuuint64_t)d->t, (cuuint64_t)d->n};
      cuuint64_t str[4] = {(cuuint64_t)d->in_pitch * 2, (cuuint64_t)2 * d->w * d->in_pitch * 2,
                           (cuuint64_t)d->h * d->w * d->in_pitch * 2,
                           (cuuint64_t)d->t * d->h * d->w * d->in_pitch * 2};
      cuuint32_t box[5] = {(cuuint32_t)d->cin, (cuuint32_t)RP, (cuuint32_t)rows[q], 1, 1};
      rc = encode_tiled_nd(m, base, 5, dims, str, box, a_swz, "window input (row parity)");
      if (rc != VSB_OK) return rc;
    }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)d->cout};
    cuuint64_t str[1] = {(cuuint64_t)k_total * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)block_n};
    rc = encode_tiled_nd(&plan->map_b, d->wgt, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, "window weights");
    if (rc != VSB_OK) return rc;
  }
  const int box_w = RP >= 32 ? 32 : RP, box_h = RP >= 32 ? 1 : 32 / RP;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->cout, (cuuint64_t)wo, (cuuint64_t)ho, (cuuint64_t)to * d->n};
    cuuint32_t box[4] = {(cuuint32_t)epi_n, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint64_t str[3] = {(cuuint64_t)d->out_pitch * 2, (cuuint64_t)wo * d->out_pitch * 2,
                         (cuuint64_t)ho * wo * d->out_pitch * 2};
    rc = encode_tiled_nd(&plan->map_out, d->out, 4, dims, str, box, swizzle_for(epi_n * 2), "window output");
    if (rc != VSB_OK) return rc;
    if (d->residual) {
      cuuint64_t rstr[3] = {(cuuint64_t)d->res_pitch * 2, (cuuint64_t)wo * d->res_pitch * 2,
                            (cuuint64_t)ho * wo * d->res_pitch * 2};
      rc = encode_tiled_nd(&plan->map_res, d->residual, 4, dims, rstr, box, swizzle_for(epi_n * 2), "window residual");
      if (rc != VSB_OK) return rc;
    } else {
      plan->map_res = plan->map_out;
    }
  }
  // ---- parameters
  p.to = to; p.ho = ho; p.wo = wo;
  p.R = R; p.RP = RP; p.rp_shift = rp_shift;
  p.yb_count = ceil_div(ho, R);
  p.kt = d->kt;
  p.L = d->kt > 1 ? to : 1;
  p.total_runs = d->kt > 1 ? d->n * p.yb_count : d->n * to * p.yb_count;
  p.pt_lo = d->pt_lo; p.pw_lo = d->pw_lo;
  p.stage_bytes = stage_bytes;
  p.stage_tx = sub_bytes0 + sub_bytes1;
  p.stages = stages;
  p.ntaps = d->kh * d->kw;
  int tap = 0, spf = 0;
  for (int kh = 0; kh < d->kh; ++kh) {
    for (int kw = 0; kw < d->kw; ++kw, ++tap) {
      const uint32_t row = d->sh == 1 ? (uint32_t)(kh * RP + kw) : (uint32_t)((kh >> 1) * RP + kw);
      const uint32_t sub = d->sh == 1 ? 0u : p.sub_off[kh & 1];
      p.tap_aoff[tap] = sub + row * rb + (uint32_t)c_lo[kw] * 2;
      p.tap_ks[tap] = (uint8_t)((c_hi[kw] - c_lo[kw]) / 16);
      spf += p.tap_ks[tap];
    }
  }
  p.steps_per_frame = spf;
  if (spf * 8 > 2048) return 1;
  p.b_blocks = tsc ? b_blocks_per_frame : b_blocks; p.b_block_bytes = b_block_bytes;
  p.tsc = tsc ? 1 : 0;
  p.tsplit = tsplit ? 1 : 0;
  p.reverse = (d->flags & VSB_PLAN_REVERSE) ? 1 : 0;
  p.n_clips = d->n;
  p.t_in = d->t;
  p.b_blocks_per_frame = b_blocks_per_frame; p.k_per_frame = k_per_frame;
  p.block_n = block_n; p.epi_n = epi_n; p.epi_chunks = epi_chunks; p.epi_bufs = epi_bufs; p.epi_warps = epi_warps;
  p.box_w = box_w; p.box_h = box_h;
  p.row_bytes = (uint32_t)rb;
  p.off_b = (uint32_t)stages * stage_bytes;
  p.off_epi = p.off_b + (uint32_t)((b_bytes + 1023) & ~1023ll);
  p.off_bar = p.off_epi + (uint32_t)(epi_warps * epi_bufs * 32 * epi_n * 2);
  p.off_tab = p.off_bar + 1024;
  p.idesc = umma_idesc_bf16(128, block_n);
  static const bool no_pair_env = getenv("VSB_WIN_NO_PAIR") != nullptr;
  const bool no_pair = no_pair_env || (d->flags & VSB_PLAN_NO_PAIR);
  p.pair = (!no_pair && d->kt == 1 && block_n <= 128 && stages >= 4) ? 1 : 0;
  p.nacc = tsc ? acc_slots : (p.pair ? 4 : 2);
  p.nacc_shift = 0;
  while ((1 << p.nacc_shift) < p.nacc) ++p.nacc_shift;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(p.nacc * block_n)) tmem_cols <<= 1;
  p.tmem_cols = tmem_cols;
  p.scale = d->scale; p.bias = d->bias;
  p.has_residual = d->residual != nullptr;
  p.relu = d->relu;
  p.dbg = nullptr;
  if (getenv("VSB_WIN_DEBUG")) {  // debug only: the one place the library allocates device memory
    if (cudaMalloc(&p.dbg, 16 * sizeof(long long)) == cudaSuccess) (void)cudaMemset(p.dbg, 0, 16 * sizeof(long long));
    else p.dbg = nullptr;
  }
  plan->smem_bytes = smem_bytes;
  int sms = 148;
  {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) (void)cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    (void)cudaGetLastError();
  }
  static const bool one_cta_env = getenv("VSB_WIN_ONE_CTA") != nullptr;
  const bool one_cta = one_cta_env || (d->flags & VSB_PLAN_ONE_CTA);
  const int ctas_per_sm = (!one_cta && smem_bytes <= 113 * 1024 && tmem_cols <= 256) ? 2 : 1;
  plan->grid = (unsigned)(p.total_runs < sms * ctas_per_sm ? p.total_runs : sms * ctas_per_sm);
  plan->desc.block_n = block_n;
  plan->desc.kchunk = d->cin;
  plan->desc.stages = stages;

  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(conv_win_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(conv_win_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  if (attr_err != cudaSuccess) {
    set_error("cudaFuncSetAttribute(conv_win_kernel) failed: %s", cudaGetErrorString(attr_err));
    return VSB_ERR_CUDA;
  }
  return VSB_OK;
}

int win_plan_launch(const vsb_conv_plan* plan, cudaStream_t stream) {
  if (plan->win.dbg)
    (void)launch_pdl(conv_win_kernel<true>, plan->grid, kThreads, plan->smem_bytes, stream, 1, plan->map_a, plan->map_a1,
                     plan->map_b, plan->map_out, plan->map_res, plan->win);
  else
    (void)launch_pdl(conv_win_kernel<false>, plan->grid, kThreads, plan->smem_bytes, stream, 1, plan->map_a, plan->map_a1,
                     plan->map_b, plan->map_out, plan->map_res, plan->win);
  VSB_CHECK_LAUNCH("conv_win_kernel");
  return VSB_OK;
}

}  // namespace vsb
