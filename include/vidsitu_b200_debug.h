/*
 * Debug-only entry points of libvidsitu_b200.so (not part of the drop-in
 * boundary; used by tools/gpu_probe.py to pin hardware semantics).
 */
#ifndef VIDSITU_B200_DEBUG_H_
#define VIDSITU_B200_DEBUG_H_
#ifdef __cplusplus
extern "C" {
#endif

/* One im2col-mode TMA load (SWIZZLE_NONE) of a bf16 [n,t,h,w,pitch] tensor with the
 * given corners / traversal strides / box, issued at base coordinate
 * (cc,cw,ch,cd,cn) with filter offsets (ow,oh,od).  `out` receives the
 * pixel_box * chan_box bf16 values exactly as they landed in shared memory. */
int vsb_debug_im2col_probe(const void* in, int n, int t, int h, int w, int c, int pitch, int lw, int lh, int lt,
                           int uw, int uh, int ut, int sw, int sh, int st, int chan_box, int pixel_box, int cc,
                           int cw, int ch, int cd, int cn, int ow, int oh, int od, void* out, void* stream);

/* One tcgen05.mma chain (M=128, N=n, K=16*ksteps) on a bf16 [a_rows, row_bytes/2] A tile and a
 * [n, row_bytes/2] B tile loaded with the matching TMA swizzle; the A/B descriptor start addresses
 * are advanced by shift_bytes / b_shift_bytes (rows * row_bytes + K-slice bytes).  base_mode 1 also
 * sets the descriptor's base_offset field to (addr >> 7) & 7.  out: fp32 [128, n]. */
int vsb_debug_umma_semantics(const void* a, int a_rows, const void* b, int n, int row_bytes, int ksteps,
                             int shift_bytes, int b_shift_bytes, int base_mode, float* out, void* stream);

/* Issue-rate probe: each of `grid` CTAs issues iters*4 MMAs (M=128, N=n, K=16) cycling over
 * a_tiles 16 KiB A tiles; clk[cta] = cycles from first issue to completion. */
int vsb_debug_umma_rate(int n, int iters, int a_tiles, int a_from_same, int grid, int smem_pad_kb, long long* clk,
                        void* stream);

/* Issue-rate probe 2: groups x per_group MMAs (M=128, N=n, K=16) with a tcgen05.commit after every group (commit_each)
 * or only at the end; A rows of row_bytes (32/64/128, matching swizzle), start shifted by shift_rows rows (+1 row per
 * MMA of a group when walk != 0).  clk[2*cta] = cycles until the last instruction was issued, clk[2*cta+1] = until
 * completion. */
int vsb_debug_umma_rate2(int n, int groups, int per_group, int commit_each, int row_bytes, int shift_rows, int walk,
                         int grid, long long* clk, void* stream);

/* Role timeline counters of a window-algorithm conv plan created with VSB_WIN_DEBUG=1 in the
 * environment (see conv_win_sm100.cu); synchronises the device, copies 16 counters and clears them. */
struct vsb_conv_plan;
int vsb_debug_conv_stats(const struct vsb_conv_plan* plan, long long* out16);

/* How a conv plan will run: out8 = {algo (1 im2col, 2 window), temporal-scatter mode (0/1), pipeline
 * stages, TMEM accumulators, grid, dynamic shared memory bytes, block_n, CTAs per SM the grid assumes}. */
int vsb_debug_conv_plan_info(const struct vsb_conv_plan* plan, long long* out8);

/* TMA load-path probe: every CTA streams `tiles` 128-row tiles of a dense bf16 [n, t, h, w, c] tensor through a ring
 * of `stages` slots that are freed as soon as they are full.  mode 0: tiled 5-D boxes [kc, RP, box_rows] (RP = power of
 * two > w, out-of-bounds slots zero-filled; kt boxes per tile = frames t-1, t, t+1) as the fused bottleneck loads x;
 * mode 1: 2-D boxes [kc, 128 pixels].  clk[cta] = cycles. */
int vsb_debug_tma_rate(const void* x, int n, int t, int h, int w, int c, int mode, int kt, int box_rows, int stages,
                       int tiles, int grid, long long* clk, void* stream);

/* Role timeline counters of a fused-bottleneck plan created with VSB_FUSED_DEBUG=1 in the environment
 * (see bottleneck_fused_sm100.cu); synchronises the device, copies 32 counters and clears them. */
struct vsb_bottleneck_plan;
int vsb_debug_bottleneck_stats(const struct vsb_bottleneck_plan* plan, long long* out32);

/* Host only (no GPU needed): the (A_c, B_c) pairs vsb_pack_frames uses in bf16 mode, i.e. coefficients with
 * bf16(fma(x, A_c, B_c)) == bf16(((x / 255) - mean_c) / std_c) for every byte value x (utils/video_utils.py:147-164
 * evaluated in fp32, then rounded once).  Returns 1 and fills a3 / b3, or 0 when no such pair exists within 3 ulps of the
 * rounded exact coefficients (vsb_pack_frames then runs the table kernel). */
int vsb_debug_pack_fma_coeffs(const float* mean3, const float* std3, float* a3, float* b3);

#ifdef __cplusplus
}
#endif
#endif
