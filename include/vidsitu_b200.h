/*
 * vidsitu_b200 -- C ABI of the B200-native SlowFast / I3D event-clip forward.
 *
 * This is the drop-in boundary for the hot path of TheShadow29/VidSitu.  The
 * reference has no FFI on this path: its "interface" is the set of torch.nn
 * calls made by SFBase.forward_encoder / head / forward_decoder
 * (vidsitu_code/mdl_sf_base.py:116-216) through the SlowFast submodule.  Each
 * entry point below replaces one family of those calls and cites it.  The
 * Python host (vidsitu_b200/sf_base.py) binds them with ctypes; INTEGRATION.md
 * shows the stub a maintainer would add to the reference.
 *
 * Conventions
 *   - every function returns 0 on success, a negative vsb_status otherwise;
 *     vsb_last_error() returns a thread-local human-readable message;
 *   - all tensor pointers are DEVICE pointers owned by the caller (torch's
 *     caching allocator in practice); nothing here allocates device memory;
 *   - work is enqueued on the caller's stream (cudaStream_t passed as void*)
 *     and never synchronises, so a whole forward can be captured in a CUDA graph;
 *   - activations are channels-last NTHWC; `*_pitch` is the element distance
 *     between consecutive pixels (>= channels), so an op can read or write a
 *     channel slice of a wider buffer (the Fast->Slow concat,
 *     SlowFast/slowfast/models/video_model_builder.py:124-131);
 *   - dtype: VSB_BF16 runs the tcgen05/TMEM tensor-core kernels, VSB_F32 runs
 *     the CUDA-core verification kernels (fp32 storage + FFMA accumulate).
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point fails with VSB_ERR_CUDA.
 */
#ifndef VIDSITU_B200_H_
#define VIDSITU_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSB_ABI_VERSION 8

typedef enum vsb_status {
  VSB_OK = 0,
  VSB_ERR_INVALID = -1, /* bad argument / unsupported shape            */
  VSB_ERR_CUDA = -2,    /* CUDA runtime or driver error                */
  VSB_ERR_ALIGN = -3    /* pointer / pitch violates a TMA requirement  */
} vsb_status;

typedef enum vsb_dtype { VSB_BF16 = 0, VSB_F32 = 1 } vsb_dtype;

int vsb_abi_version(void);
const char* vsb_last_error(void);
/* number of kernels launched by this library since load (bench "gpu_launches") */
uint64_t vsb_launch_count(void);

/* ------------------------------------------------------------------ pack
 * Replaces utils/video_utils.py:147-164 (tensor_normalize: x/255, -mean, /std)
 * + dat_loader.py:483 (permute to CTHW) + utils/video_utils.py:41-74
 * (pack_pathway_output: temporal index_select for the slow pathway).
 * frames: uint8 [n, t_in, h, w, 3]; out: [n, t_out, h, out_w, c_pad] with
 * out[.., t, y, x_off + x, c] = (frames[.., idx[t], y, x, c'] / 255 - mean[c]) / std[c],
 * c' = 2-c if reverse_channels else c; channels 3..c_pad-1 are written as 0.
 * out_w >= x_off + w lets the caller keep zero columns around each row (the stem's
 * halo, so the stem conv needs no W padding); those columns are never written.
 * idx is a HOST array of t_out frame indices (t_out <= 64).                 */
int vsb_pack_frames(const uint8_t* frames, int n, int t_in, int h, int w, const int* idx, int t_out,
                    const float* mean3, const float* std3, int reverse_channels, void* out, int c_pad, int out_w,
                    int x_off, int dtype, void* stream);

/* ------------------------------------------------------------------ conv
 * Replaces nn.Conv3d(bias=False) + eval-mode nn.BatchNorm3d (+ nn.ReLU)
 * (+ residual add) as used by stem_helper.py:157-178, resnet_helper.py:182-240,
 * 326-358 and video_model_builder.py:109-131:
 *   out[m, co] = act( scale[co] * sum_{tap,ci} in[pix(m,tap), ci] * wgt[co,tap,ci]
 *                     + bias[co] + residual[m, co] )
 * wgt is packed [cout][kt*kh*kw][cin] (K-major) in `dtype`; scale/bias are fp32
 * (frozen BN folded: scale = gamma/sqrt(var+eps), bias = beta - mean*scale, or
 * scale = 1, bias = conv bias for the Nonlocal 1x1x1 convs).                 */
typedef struct vsb_conv_desc {
  int dtype;                    /* vsb_dtype                                     */
  const void* in;               /* [n, t, h, w, in_pitch], first cin channels    */
  int n, t, h, w, cin, in_pitch;
  const void* wgt;              /* [cout][kt*kh*kw][cin]                         */
  int cout;
  int kt, kh, kw;
  int st, sh, sw;
  int pt_lo, ph_lo, pw_lo;      /* leading (lower) zero padding                  */
  int pt_hi, ph_hi, pw_hi;      /* trailing padding (== lower for every reference conv;
                                   differs only for the re-viewed stem)          */
  const float* scale;           /* [cout]                                        */
  const float* bias;            /* [cout]                                        */
  const void* residual;         /* nullable, [M, res_pitch]                      */
  int res_pitch;
  int relu;
  void* out;                    /* [M, out_pitch], M = n*to*ho*wo                */
  int out_pitch;
  /* tuning, 0 = automatic */
  int block_n;                  /* 16..256, multiple of 16, divides cout         */
  int kchunk;                   /* 16 / 32 / 64 channels per TMA box             */
  int stages;                   /* smem pipeline depth                           */
  /* algorithm of the bf16 path: 0 = automatic, 1 = im2col-mode implicit GEMM (any conv),
   * 2 = shared-memory window (input rows loaded once, every filter tap is a shifted
   * tcgen05 descriptor over them; needs st == sw == 1, sh in {1,2}, cin*2 in {32,64,128} bytes,
   * cout <= 256 and the whole weight matrix resident in shared memory).                */
  int algo;
  /* window algorithm only: input-channel range [lo, hi) (multiples of 16) that tap kw actually
   * reads -- pixel-group restated convs (block-Toeplitz weights) touch only a few pixels of
   * their outer groups.  hi == 0 means the full range.  When any range is partial, wgt holds
   * only those channels: [cout][kt][kh][sum_kw (hi - lo)].                             */
  int kw_c_lo[8];
  int kw_c_hi[8];
  /* Optional second source (bf16 im2col algorithm only): a strided 1x1x1 conv over `in2` accumulated
   * into the SAME output, i.e. the projection shortcut of a ResBlock (resnet_helper.py:326-358:
   * relu(branch1_bn(branch1(x)) + branch2(x))) computed inside the block's last conv instead of as a
   * separate launch whose result is written to and re-read from HBM as `residual`:
   *   out[m, co] = act(scale[co] * (sum_k in[..] * wgt[co, k] + sum_ci in2[pix2(m), ci] * wgt[co, K + ci]) + bias[co])
   * pix2(m) = (n, to*st2, ho*sh2, wo*sw2).  Both BatchNorms share the epilogue scale: the caller folds
   * the ratio of the two BN scales into one of the weight blocks (fp32 master weights, one bf16
   * rounding) and sums the biases.  wgt rows are [K = kt*kh*kw*cin | cin2 rounded up to kchunk].
   * in2 == NULL: no second source.                                                        */
  const void* in2;
  int t2, h2, w2, cin2, in2_pitch;
  int st2, sh2, sw2;
  /* more tuning of the bf16 epilogue, 0 = automatic: column chunk of the per-warp staging slabs
   * (32 or 64 channels) and slabs per epilogue warp (2..4; residual prefetch distance = epi_bufs - 1).
   * They only move shared memory between the main-loop ring and the epilogue; results are unchanged. */
  int epi_n;
  int epi_bufs;
  /* plan variants, OR of VSB_PLAN_* (0 = automatic); like the tuning above they never change results */
  int flags;
  /* bf16 path, im2col algorithm, no residual: write the 16-bit outputs as IEEE half instead of bf16
   * (attention scores of the non-local block: 3 more mantissa bits in front of the softmax).        */
  int out_f16;
  /* Per-clip weights (bf16, im2col algorithm, 1x1x1 stride-1 convs without residual): > 0 makes clip i use
   * the [cout][cin] weight matrix that starts wgt_clip_rows rows after clip i-1's, i.e. a batched product
   * out_i = in_i . W_i^T - the two einsums of the non-local block with phi_i / g_i^T as W_i
   * (nonlocal_helper.py:123-141).  `wgt` must hold (n-1) * wgt_clip_rows + cout rows.  0 = one matrix.   */
  int wgt_clip_rows;
  /* Tile-granular chaining of two consecutive launches of a stream (bf16, one-SM im2col kernel; ABI v6).
   * A ResBlock's last conv and the next block's first 1x1x1 conv (resnet_helper.py:352-358 followed by :225-229)
   * then run CO-RESIDENT on every SM and the second reads each 128-pixel tile of `out` from the L2 right after the
   * first wrote it, instead of re-reading the whole tensor from HBM one launch later.
   *   producer (tile_signal != NULL): every epilogue warp adds 1 to tile_signal[m] (release, gpu scope) once its
   *     rows of output tile m = rows [128 m, 128 m + 128) are complete in global memory: the counter of a finished
   *     tile is 8 * (cout / block_n).
   *   consumer (tile_wait != NULL; 1x1x1 stride 1, one column block, input = the producer's output): loads tile m
   *     once tile_wait[m] >= tile_wait_count, then resets tile_wait[m] to 0; it does NOT wait for the previous
   *     kernel as a whole (programmatic dependent launch), so it must be launched right after its producer on the
   *     same stream and must not write memory that the producer or anything before it still reads.
   * Counters are caller-owned device memory (one uint32 per 128-row tile), zero before the first use.          */
  unsigned int* tile_signal;
  unsigned int* tile_wait;
  int tile_wait_count;
  /* bf16 im2col algorithm: at most this many CTAs (0 = automatic: every SM, twice when two CTAs fit).  A chained
   * pair splits the SMs' CTA slots between producer and consumer with it.                                    */
  int grid_limit;
  /* bf16 path, one-SM im2col kernel, no residual / second source: `in` and `wgt` hold IEEE half instead of bf16
   * (ABI v7).  The score product of the non-local block: theta and phi enter the exponent of the softmax, so
   * their rounding (2^-9 relative in bf16) becomes a RELATIVE error of every attention weight; half gives 2^-12. */
  int in_f16;
} vsb_conv_desc;

#define VSB_PLAN_STREAM_WEIGHTS 1 /* im2col: never keep the weight block resident in shared memory      */
#define VSB_PLAN_ONE_CTA 2        /* one CTA per SM even when two would fit                              */
#define VSB_PLAN_NO_PAIR 4        /* window: the MMA issuer does not interleave two tiles                */
#define VSB_PLAN_NO_TSCATTER 8    /* window: output-stationary temporal taps instead of temporal scatter */
#define VSB_PLAN_TWO_SM 16        /* im2col: CTA pairs (tcgen05 cta_group::2), 256-pixel tiles, each CTA
                                     streams half of the weight rows; kchunk 64 layers with streamed weights
                                     (automatic for 256-wide column blocks with K >= 512)                 */
#define VSB_PLAN_NO_TILE_SPLIT 64 /* window: epilogue warp groups always split column chunks, never tiles    */
#define VSB_PLAN_REVERSE 128      /* walk the output in DESCENDING order (tiles / clips).  A kernel that starts where
                                     its producer finished finds the most recently written ~100 MB of its input still
                                     in the 126 MB L2; the engine alternates the direction along each pathway.       */
#define VSB_PLAN_TWO_SM_RESIDENT 256 /* CTA pairs with the weight block resident, half in each CTA (K*block_n <= 224 KB) */
#define VSB_PLAN_ONE_SM 32        /* im2col: never use CTA pairs                                          */

typedef struct vsb_conv_plan vsb_conv_plan;

int vsb_conv3d_plan_create(const vsb_conv_desc* desc, vsb_conv_plan** plan);
int vsb_conv3d_run(const vsb_conv_plan* plan, void* stream);
void vsb_conv3d_plan_destroy(vsb_conv_plan* plan);
/* output extent of a plan: M = n*to*ho*wo and (to, ho, wo) */
int vsb_conv3d_plan_out_shape(const vsb_conv_plan* plan, int* to, int* ho, int* wo);
/* 2*M*cout*taps*cin of the convolution as launched (padded channels included) */
double vsb_conv3d_plan_flops(const vsb_conv_plan* plan);

/* ------------------------------------------------- fused identity bottleneck block
 * Replaces, in ONE launch, the three convs of BottleneckTransform (SlowFast/slowfast/models/
 * resnet_helper.py:225-240: a = Conv3d [kt,1,1] pad [kt/2,0,0] + BN + ReLU, b = Conv3d [1,3,3] pad [0,1,1]
 * + BN + ReLU, c = Conv3d 1x1x1 + BN) and the identity ResBlock around them (resnet_helper.py:352-358:
 * relu(x + branch2(x))), for blocks WITHOUT a projection shortcut and with unit strides / dilation:
 *   out = relu(x + sc * (relu(sb * (relu(sa * (x (*) wa) + ba) (*) wb) + bb) . wc) + bc)
 * a's and b's outputs are rounded to bf16 exactly where the three-launch path rounds them, but never leave the
 * SM: x is read and out is written, 8 instead of 16 bottleneck-widths of HBM traffic per pixel.  bf16 only.
 * x / out: [n, t, h, w, pitch] with c stored channels (multiple of 16, <= 256); d = stored bottleneck width
 * (16, 32 or 64); wa [d][kt][c], wb [d][3*3][d] (tap = kh*3 + kw), wc [c][d], all bf16 K-major, zero-padded to
 * the stored widths; s? / b? = folded frozen-BatchNorm scale / bias (fp32; [d], [d], [c]).
 * A "pixel" may be a pixel GROUP: the caller can restate thin layers on groups of J pixels along W
 * (block-Toeplitz weights, [.., W, C] viewed as [.., W/J, J*C]); the kernel only sees slots.          */
typedef struct vsb_bottleneck_desc {
  const void* x;
  int n, t, h, w, c, x_pitch;
  void* out;
  int out_pitch;
  int d;
  int kt;                       /* temporal taps of conv a: 1 or 3                 */
  const void* wa;
  const void* wb;
  const void* wc;
  const float* sa; const float* ba;
  const float* sb; const float* bb;
  const float* sc; const float* bc;
  /* tuning, 0 = automatic: x ring depth, tiles per walk (algo 1: rows per strip), CTAs */
  int stages, walk_len, grid;
  /* ABI v8.  0 = tcgen05 flat-raster kernel (bottleneck_fused_sm100.cu).  1 = warp-level MMA walk kernel for THIN
   * blocks (bottleneck_thin_sm100.cu): d = 8 or 16 UNGROUPED stored channels, c = 4 d, x dense (x_pitch == c); a CTA
   * walks a strip of rows through time, every frame of the strip is fetched once by one bulk copy, the three convs
   * run on mma.sync fragments (the Fast pathway's res2 / res3, where a tcgen05.mma would be N <= 64 wide and
   * issue-bound).  Same tensors, weight layouts and rounding points as algo 0. */
  int algo;
  /* algo 1 only.  0 = identity block (x has c channels).  8 = block 0 of a stage WITH a 1x1x1 projection shortcut
   * (resnet_helper.py:296-310, 352-358: relu(BN_1(W1 x) + BN_c(c(..)))) and unit strides: x has cin = d = 8 channels,
   * wa is [d][kt][cin], wc is [c][d + cin] = [Wc * r_c | W1 * r_1] with the two BatchNorm scales folded as ratios to
   * the common per-channel scale sc, bc = the sum of the two biases; out = relu(sc * ([b | x] . wc) + bc). */
  int cin;
} vsb_bottleneck_desc;

typedef struct vsb_bottleneck_plan vsb_bottleneck_plan;
int vsb_bottleneck_plan_create(const vsb_bottleneck_desc* desc, vsb_bottleneck_plan** plan);
int vsb_bottleneck_run(const vsb_bottleneck_plan* plan, void* stream);
void vsb_bottleneck_plan_destroy(vsb_bottleneck_plan* plan);
/* out8 = {slots per flat row, flat rows per frame, x ring slots * 100 + chunks per slot, tiles per CTA, grid, dynamic shared memory
 * bytes, tiles per clip, TMEM columns}; algo 1: {rows per strip, strips per frame, ring slots, frame steps per CTA, grid, dynamic
 * shared memory bytes, conv-a tiles * 1000 + conv-b/c tiles per frame step, 0} */
int vsb_bottleneck_plan_info(const vsb_bottleneck_plan* plan, long long* out8);

/* ------------------------------------------------- fused [kt,7,7] stem: conv + BN + ReLU + max-pool (ABI v7)
 * Replaces, in ONE launch, ResNetBasicStem.forward for the 64-channel stems (SlowFast/slowfast/models/
 * stem_helper.py:157-178: Conv3d kernel [kt,7,7] stride [1,2,2] pad [kt/2,3,3] + frozen BatchNorm3d + ReLU +
 * MaxPool3d kernel [1,3,3] stride [1,2,2] pad [0,1,1]) with 3 input and 64 output channels: kt = 1 - the Slow
 * pathway of SlowFast, Slow-only and C2D; kt = 5 - I3D.  bf16.  The conv output never reaches HBM.
 *   in     [frames, h, w_buf, 4] packed frames as vsb_pack_frames writes them with x_off = 3 (image pixel x at
 *          buffer pixel x + 3; pixels outside the image and channel 3 are zero); h, w multiples of 32,
 *          w_buf >= w + 8;  frames = clips * T
 *   wgt    kt = 1: [64][7][8][4] bf16: wgt[co][kh][kw][c] = W[co][c][0][kh][kw] for kw < 7, c < 3, else 0;
 *          kt = 5: [5 * 64][7][8][4], the same block per temporal tap, in the tap order 0, 2, 1, 4, 3 (the kernel
 *          multiplies taps (2,1) and (4,3) as stacked pairs)
 *   t      kt = 5 only: frames per clip (temporal taps do not cross clips): 3..8 or a multiple of 8
 *   out    [frames, h/4, w/4, out_pitch]: channels [0, 64) are written (zeroed, then max-combined), the rest of
 *          each pixel (the lateral connection's channel slice) is left untouched                       */
typedef struct vsb_stem_pool_desc {
  const void* in;
  int frames, h, w, w_buf;
  const void* wgt;
  const float* scale;           /* [64] folded BatchNorm */
  const float* bias;            /* [64]                  */
  void* out;
  int out_pitch;
  int kt;                       /* temporal taps: 1 or 5 */
  int t;                        /* frames per clip (kt = 5) */
} vsb_stem_pool_desc;
typedef struct vsb_stem_pool_plan vsb_stem_pool_plan;
int vsb_stem_pool_plan_create(const vsb_stem_pool_desc* desc, vsb_stem_pool_plan** plan);
int vsb_stem_pool_run(const vsb_stem_pool_plan* plan, void* stream);   /* 2 launches: zero-fill + fused kernel */
void vsb_stem_pool_plan_destroy(vsb_stem_pool_plan* plan);
int vsb_stem_pool_plan_desc(const vsb_stem_pool_plan* plan, vsb_stem_pool_desc* desc);

/* --------------------------------------------------------------- max-pool
 * Replaces nn.MaxPool3d (stem_helper.py:169-171, video_model_builder.py:235-241,
 * nonlocal_helper.py:98-103).  Implicit -inf padding as in PyTorch.  Channels
 * [c, c_out) of the output are written as zero (channel padding).           */
int vsb_maxpool3d(const void* in, int n, int t, int h, int w, int c, int in_pitch, void* out, int out_pitch,
                  int c_out, int kt, int kh, int kw, int st, int sh, int sw, int pt, int ph, int pw, int dtype,
                  void* stream);

/* ------------------------------------------------- global average pool
 * Replaces nn.AdaptiveAvgPool3d((1,1,1)) + torch.cat of ResNetBasicHead_Trimmed
 * (vidsitu_code/mdl_sf_base.py:95-113): feats[i, feat_off + ch] = mean over the
 * thw positions of clip i.  feats is fp32 [n, feat_pitch].                   */
int vsb_global_avgpool(const void* in, int n, int thw, int c, int in_pitch, float* feats, int feat_pitch,
                       int feat_off, int dtype, void* stream);

/* ------------------------------------------------------------------ linear
 * Replaces nn.Linear (+ nn.ReLU) of proj_head (vidsitu_code/mdl_sf_base.py:161-167):
 * y[i, o] = act(b[o] + sum_k x[i,k] * w[o,k]); all fp32, w row-major [dout, din]. */
int vsb_linear(const float* x, int n, int din, const float* w, const float* b, float* y, int dout, int relu,
               void* stream);

/* --------------------------------------------------------- non-local attention
 * Replaces the two einsums + softmax of Nonlocal.forward
 * (SlowFast/slowfast/models/nonlocal_helper.py:123-141).  Per clip i:
 *   A = theta_i [tq, c] . phi_i [tk, c]^T ; softmax: A = softmax(A * c^-0.5) over tk;
 *   dot_product: A = A / tk ;  out_i [tq, c] = A . g_i [tk, c].
 * theta/phi/g/out are channels-last with their own pitches.                  */
int vsb_nonlocal_attention(const void* theta, int theta_pitch, const void* phi, int phi_pitch, const void* g,
                           int g_pitch, void* out, int out_pitch, int n, int tq, int tk, int c, int softmax,
                           int dtype, void* stream);

/* Tensor-core route of the same block (bf16): the caller runs the two einsums as per-clip 1x1x1 convs
 * (vsb_conv3d_*: scores_i = theta_i . phi_i^T with phi_i as the [tk, c] weight matrix and the softmax
 * scale in `scale`; out_i = P_i . g_i with g_i^T as the [c, tk] weight matrix) and these two kernels in
 * between.  vsb_score_rows: in place on bf16 scores [rows, pitch]: columns [0, valid) <- softmax over them
 * (softmax != 0; nonlocal_helper.py:128-131) or unchanged (dot_product, already divided by tk), columns
 * [valid, width) <- 0 (the zero K padding of the second product).  vsb_transpose_pad: bf16
 * out[i][ch][k] = in[i][k][ch] for k < rows, 0 for rows <= k < out_pitch.                              */
int vsb_score_rows(void* scores, long long rows, int valid, int width, int pitch, int softmax, void* stream);
int vsb_transpose_pad(const void* in, int in_pitch, void* out, int n, int rows, int cols, int out_pitch,
                      void* stream);

/* ------------------------------------------------------- verb softmax + top-k
 * Replaces F.softmax(mdl_out, -1) + sort(descending=True)[:topk_save] of
 * EvalB.forward_one_batch (vidsitu_code/evl_vsitu.py:39-47): per row of the
 * fp32 [n, pitch] logits (first v entries), idx[row, r] = index of the r-th largest
 * probability (ties: lower index first, as a stable descending sort) and
 * prob[row, r] = exp(x - max) / sum(exp(x - max)) of that entry, r < k <= 16.   */
int vsb_softmax_topk(const float* logits, int n, int v, int pitch, int k, int* idx, float* prob, void* stream);

/* ------------------------------------------------------------- layout helper
 * NTHWC (pitch) -> NCTHW fp32 contiguous, for callers that want the reference's
 * forward_features() tensors (mdl_sf_base.py:21-34) materialised.           */
int vsb_nthwc_to_ncthw_f32(const void* in, int n, int thw, int c, int in_pitch, float* out, int dtype,
                           void* stream);

/* NCTHW fp32 (the reference's already-normalised clip tensors,
 * vidsitu_code/mdl_sf_base.py:169-180) -> NTHWC with the c <= 3 planes packed into
 * c_pad == 4 channels (channel 3 = 0), bf16 or fp32; rows of w pixels are written at
 * pixel offset x_off of output rows of out_w pixels (see vsb_pack_frames).    */
int vsb_ncthw_f32_to_nthwc(const float* in, int n, int c, long long thw, int w, void* out, int c_pad, int out_w,
                           int x_off, int dtype, void* stream);

/* ------------------------------------------------------------- frame ingest (ABI v7)
 * Replaces VsituDS.read_img (vidsitu_code/dat_loader.py:183-191):
 *     Image.open(path).convert("RGB").resize((224, 224))  ->  uint8 [224, 224, 3]
 * with the pixels produced ON THE DEVICE, bit-identical to Pillow / libjpeg(-turbo): the host only parses the headers
 * and decodes the (inherently sequential) Huffman segment into quantised coefficient blocks; dequantisation, the
 * "islow" integer IDCT, fancy chroma upsampling, YCbCr -> RGB and Pillow's fixed-point BICUBIC resampling are kernels.
 * Decoded: baseline / extended-sequential Huffman JPEGs, 8 bit, grayscale or YCbCr 4:4:4 / 4:2:2 / 4:2:0, with or
 * without restart markers (the dataset's frames are ffmpeg MJPEG 4:2:0, prep_data/dwn_yt.py:229-250).  Anything else
 * (progressive, arithmetic coding, CMYK, other sampling) is refused with VSB_ERR_INVALID: the caller decodes such a
 * file on the host.
 * A decoder owns its workspaces (pinned host staging, device planes) for frames of up to max_width x max_height; it is
 * not thread-safe - one decoder per host thread.  `out` is device memory, [out_h, out_w, 3] uint8; the call returns once
 * the host part is done, the device part is enqueued on `stream`.                                              */
typedef struct vsb_jpeg_decoder vsb_jpeg_decoder;
int vsb_jpeg_info(const uint8_t* jpeg, unsigned long long bytes, int* width, int* height, int* components, int* h_samp,
                  int* v_samp);
int vsb_jpeg_decoder_create(int max_width, int max_height, vsb_jpeg_decoder** dec);
void vsb_jpeg_decoder_destroy(vsb_jpeg_decoder* dec);
int vsb_jpeg_decode_resize(vsb_jpeg_decoder* dec, const uint8_t* jpeg, unsigned long long bytes, uint8_t* out, int out_h,
                           int out_w, void* stream);
/* Batched, FULLY on-device variant: the Huffman segments are decoded on the GPU too, one warp per frame (a batch of a
 * feature-extraction run holds hundreds of independent frames; within a frame the code stream is sequential).  The
 * host parses headers and stages the file bytes; no coefficient and no pixel crosses PCIe.  jpegs[i] / bytes[i]: the
 * i-th file in host memory; outs[i]: its device destination, uint8 [out_h, out_w, 3]; status[i] (host) receives VSB_OK
 * or VSB_ERR_INVALID (refused or corrupt file: its output is untouched and the caller decodes it on the host).  Same
 * bits as vsb_jpeg_decode_resize.  Workspaces grow on demand and are owned by the handle; the call returns after the
 * batch has completed on `stream` (the verdicts are part of the result).                                       */
typedef struct vsb_jpeg_batch vsb_jpeg_batch;
int vsb_jpeg_batch_create(vsb_jpeg_batch** batch);
void vsb_jpeg_batch_destroy(vsb_jpeg_batch* batch);
int vsb_jpeg_batch_decode_resize(vsb_jpeg_batch* batch, const uint8_t* const* jpegs, const unsigned long long* bytes, int n,
                                 uint8_t* const* outs, int out_h, int out_w, int* status, void* stream);
/* the resampling alone: device uint8 [h, w, 3] -> [out_h, out_w, 3], == PIL.Image.resize((out_w, out_h)) */
int vsb_resize_bicubic_u8(vsb_jpeg_decoder* dec, const uint8_t* in, int h, int w, uint8_t* out, int out_h, int out_w,
                          void* stream);

/* ------------------------------------------------------------- clip programs (ABI v7)
 * The model-level entry points: ONE handle for the whole forward of SFBase.forward_encoder (+ proj_head) at a
 * fixed batch size (vidsitu_code/mdl_sf_base.py:182-216; SlowFast.forward, SlowFast/slowfast/models/
 * video_model_builder.py:383-396), i.e. the create / workspace_bytes / forward triple a non-Python host binds.
 *
 *   build   vsb_program_create, then one vsb_program_add_* per launch IN PROGRAM ORDER.  Every add mirrors the
 *           per-op entry point above with (stream) replaced by (lane, name): lane 0 = the caller's stream (Slow
 *           pathway, head), lane 1 = the program's own side stream (Fast pathway and the lateral convs, which
 *           video_model_builder.py:124-131 makes the only meeting points); vsb_program_add_sync(from, to) makes
 *           lane `to` wait for everything lane `from` has been given so far.  A program that uses lane 1 must
 *           fork (sync 0 -> 1) before its first lane-1 op and join (sync 1 -> 0) after its last.
 *           Conv / bottleneck plans are borrowed: they must outlive the program.
 *           (vidsitu_b200/engine.py::ClipEngine.build_program is the planner that emits these calls.)
 *   run     vsb_program_run: the whole forward, one call.  vsb_program_capture first runs the program once
 *           eagerly, then records it into a CUDA graph that every later vsb_program_run replays.
 *   save    vsb_program_add_region names the device allocations the ops point into - CONST regions (packed
 *           weights, folded BatchNorm scale / bias: contents are saved) and SCRATCH regions (activations, inputs,
 *           outputs: size only, zero-filled at load) - and vsb_program_save writes a relocatable file.
 *   load    vsb_program_load rebuilds the program in any process: regions are laid out 1 KiB-aligned in
 *           `device_mem` (vsb_program_file_device_bytes bytes; NULL = the library allocates and owns it), constants
 *           are uploaded, plans and TMA descriptors re-created for the new addresses.  vsb_program_region
 *           returns where a named region lives ("frames": uint8 [n, t, h, w, 3] input; "feats": fp32 [n, D];
 *           "logits": fp32 [n, V] when the program has a projection head): the host copies frames in, calls
 *           vsb_program_run, copies features out.  See INTEGRATION.md for a complete C host.                 */
typedef struct vsb_program vsb_program;
#define VSB_REGION_CONST 0
#define VSB_REGION_SCRATCH 1

int vsb_program_create(vsb_program** prog);
void vsb_program_destroy(vsb_program* prog);
int vsb_program_add_region(vsb_program* prog, const char* name, void* ptr, unsigned long long bytes, int kind);
int vsb_program_region(const vsb_program* prog, const char* name, void** ptr, unsigned long long* bytes);
int vsb_program_num_ops(const vsb_program* prog);       /* launches + syncs */
int vsb_program_num_launches(const vsb_program* prog);  /* kernels per vsb_program_run */
unsigned long long vsb_program_device_bytes(const vsb_program* prog); /* sum of the registered regions, 1 KiB-aligned */

int vsb_program_add_conv(vsb_program* prog, const vsb_conv_plan* plan, int lane, const char* name);
int vsb_program_add_bottleneck(vsb_program* prog, const vsb_bottleneck_plan* plan, const vsb_bottleneck_desc* desc,
                               int lane, const char* name);
int vsb_program_add_stem_pool(vsb_program* prog, const vsb_stem_pool_plan* plan, int lane, const char* name);
int vsb_program_add_pack_frames(vsb_program* prog, const uint8_t* frames, int n, int t_in, int h, int w, const int* idx,
                                int t_out, const float* mean3, const float* std3, int reverse_channels, void* out,
                                int c_pad, int out_w, int x_off, int dtype, int lane, const char* name);
int vsb_program_add_maxpool3d(vsb_program* prog, const void* in, int n, int t, int h, int w, int c, int in_pitch,
                              void* out, int out_pitch, int c_out, int kt, int kh, int kw, int st, int sh, int sw,
                              int pt, int ph, int pw, int dtype, int lane, const char* name);
int vsb_program_add_global_avgpool(vsb_program* prog, const void* in, int n, int thw, int c, int in_pitch, float* feats,
                                   int feat_pitch, int feat_off, int dtype, int lane, const char* name);
int vsb_program_add_linear(vsb_program* prog, const float* x, int n, int din, const float* w, const float* b, float* y,
                           int dout, int relu, int lane, const char* name);
int vsb_program_add_nonlocal_attention(vsb_program* prog, const void* theta, int theta_pitch, const void* phi,
                                       int phi_pitch, const void* g, int g_pitch, void* out, int out_pitch, int n,
                                       int tq, int tk, int c, int softmax, int dtype, int lane, const char* name);
int vsb_program_add_score_rows(vsb_program* prog, void* scores, long long rows, int valid, int width, int pitch,
                               int softmax, int lane, const char* name);
int vsb_program_add_transpose_pad(vsb_program* prog, const void* in, int in_pitch, void* out, int n, int rows, int cols,
                                  int out_pitch, int lane, const char* name);
int vsb_program_add_sync(vsb_program* prog, int from_lane, int to_lane);

int vsb_program_run(vsb_program* prog, void* stream);
int vsb_program_capture(vsb_program* prog, void* stream);

int vsb_program_save(const vsb_program* prog, const char* path);
int vsb_program_file_device_bytes(const char* path, unsigned long long* bytes);
int vsb_program_load(const char* path, void* device_mem, unsigned long long device_bytes, vsb_program** prog);

#ifdef __cplusplus
}
#endif
#endif /* VIDSITU_B200_H_ */
