// Conv plan shared by the two tensor-core conv kernels (conv_igemm_sm100.cu: im2col-mode
// implicit GEMM; conv_win_sm100.cu: shared-memory window + shifted-descriptor taps).
#pragma once
#include <cuda.h>

#include "common.h"

namespace vsb {

struct IgemmParams {
  int m_total, to, ho, wo;
  int st, sh, sw;
  int lt, lh, lw;  // lower corner = -leading pad
  int kh, kw;
  int cin, cin_chunks, total_chunks, cps, kchunk;
  int chunks1;            // K chunks of the primary source; chunks [chunks1, total_chunks) come from the second source
  int st2, sh2, sw2;      // pixel strides of the second source (strided 1x1x1 shortcut projection)
  int block_n, n_tiles, total_tiles, stages;
  int epi_n, epi_chunks;  // epilogue column chunk (<= 64) and chunks per tile
  int epi_bufs;           // staging buffers of the epilogue (residual prefetch distance = epi_bufs - 1)
  int epi_warps;          // 8 (two CTAs per SM) or 16 (one CTA per SM)
  int b_resident;         // all weight chunks stay in shared memory for the CTA's lifetime (n_tiles == 1)
  uint32_t stage_bytes, off_bres, off_epi, off_bar;  // shared-memory layout (bytes from the 1 KiB-aligned base)
  uint32_t idesc, tmem_cols;
  const float* scale;
  const float* bias;
  int has_residual;
  int relu;
  int reverse;     // walk the output tiles in descending order (VSB_PLAN_REVERSE)
  int out_f16;     // 16-bit outputs are IEEE half instead of bf16 (vsb_conv_desc.out_f16)
  // per-clip weights (vsb_conv_desc.wgt_clip_rows > 0): tiles never straddle clips
  int clip_rows;       // output pixels per clip (0 = one weight matrix for all clips)
  int clip_tiles;      // m-tiles per clip = ceil(clip_rows / 128)
  int wgt_clip_rows;   // weight rows between consecutive clips' matrices
  // tile-granular chaining (vsb_conv_desc.tile_signal / tile_wait): counters per 128-row output tile
  unsigned int* tile_signal;  // producer: every epilogue warp adds 1 once its rows of the tile are in global memory
  unsigned int* tile_wait;    // consumer: tile m is loaded once tile_wait[m] >= tile_wait_count (then reset to 0)
  unsigned int tile_wait_count;
  long long* dbg;  // role timeline counters (VSB_WIN_DEBUG), else null
};

// ---- shared-memory window algorithm (conv_win_sm100.cu)
constexpr int kWinMaxTaps = 64;

struct WinParams {
  int to, ho, wo;          // output extent (wo in GEMM rows = pixel groups)
  int yb_count;            // row blocks per frame = ceil(ho / R)
  int R, RP, rp_shift;     // output rows per tile, window-row pitch in slots (power of two), log2(RP); R * RP == 128
  int kt;                  // temporal taps = input frames per tile
  int L;                   // tiles per run: to when kt > 1 (consecutive frames share windows), else 1
  int total_runs;
  int pt_lo, pw_lo;
  int nsub;                // sub-windows per frame: 1 (stride 1 in H) or 2 (stride 2: rows split by parity)
  int sub_map[2];          // input tensor map used by each sub-window (row-parity view when stride 2)
  int sub_hoff[2];         // first row of the box = y0 + sub_hoff (in rows of that map)
  uint32_t sub_off[2];     // byte offset of the sub-window inside a stage
  uint32_t stage_bytes, stage_tx;
  int stages;
  int ntaps;               // taps per frame
  int steps_per_frame;     // sum of tap_ks = K / 16 per frame
  uint32_t tap_aoff[kWinMaxTaps];  // byte offset (stage-relative) of the tap's first A row / K slice
  uint8_t tap_ks[kWinMaxTaps];     // K=16 steps of the tap
  int b_blocks;            // resident weight matrix: kt x b_blocks_per_frame x [block_n rows x 128 B]
  int b_blocks_per_frame;  // ceil(steps_per_frame / 4): every frame's K range starts on a block boundary
  int k_per_frame;         // K elements of one frame (kh x sum of tap channel ranges)
  uint32_t b_block_bytes;
  int block_n, epi_n, epi_chunks, epi_bufs, epi_warps;
  int nacc, nacc_shift;    // TMEM accumulators (2, 4, or the temporal-scatter ring: up to 16)
  int reverse, n_clips;    // walk the clips in descending order (VSB_PLAN_REVERSE)
  int tsplit;              // epilogue warp groups take alternate tiles (narrow tiles) instead of alternate column chunks
  int tsc;                 // temporal-scatter mode: one N = kt * block_n MMA per K step, ring of nacc accumulators
  int t_in;                // input frames
  int pair;                // MMA issuer interleaves two tiles (independent accumulation chains, shared B reads)
  int box_w, box_h;        // per-warp output box: 32 GEMM rows = box_h image rows x box_w slots
  uint32_t row_bytes;      // A row (slot) bytes: 32 / 64 / 128
  uint32_t off_b, off_epi, off_bar, off_tab;  // off_tab: per-K-step descriptor table (uint2 x steps_per_frame)
  uint32_t idesc, tmem_cols;
  const float* scale;
  const float* bias;
  int has_residual;
  int relu;
  long long* dbg;          // role timeline counters (VSB_WIN_DEBUG), else null
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

extern EncodeTiledFn g_encode_tiled;
int load_driver_entry_points();
CUtensorMapSwizzle swizzle_for(int row_bytes);
// (C,W,H,D,N) im2col-mode map over a bf16 NTHWC tensor (conv_igemm_sm100.cu); lower / upper / stride are (w, h, d)
int encode_im2col_map(CUtensorMap* map, const void* base, int n, int t, int h, int w, int c, int pitch, const int lower[3],
                      const int upper[3], const int stride_whd[3], int chan_box, int pixel_box, CUtensorMapSwizzle swz);

// window algorithm: returns VSB_OK when the plan was built, 1 when the conv is outside the
// algorithm's domain (caller falls back to im2col), a negative vsb_status on error
int win_plan_build(struct ::vsb_conv_plan* plan, const vsb_conv_desc* d, int to, int ho, int wo);
int win_plan_launch(const struct ::vsb_conv_plan* plan, cudaStream_t stream);
// two-SM (cta_group::2) variant of the im2col kernel (conv_igemm2_sm100.cu); plan->algo == 3
int igemm2_launch(const struct ::vsb_conv_plan* plan, cudaStream_t stream);

}  // namespace vsb

struct vsb_conv_plan {
  vsb_conv_desc desc;
  int to, ho, wo;
  long long m_total;
  int algo;  // 1 = im2col implicit GEMM, 2 = shared-memory window, 3 = im2col on CTA pairs (cta_group::2)
  // bf16 tensor-core path
  CUtensorMap map_a, map_b, map_out, map_res;
  CUtensorMap map_a2;  // im2col algorithm: second source (fused shortcut projection)
  CUtensorMap map_a1;  // window algorithm, stride 2: odd-row view of the input
  vsb::IgemmParams params;
  vsb::WinParams win;
  size_t smem_bytes;
  unsigned grid;
};
