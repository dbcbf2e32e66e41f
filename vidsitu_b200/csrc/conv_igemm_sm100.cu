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

#include <mutex>
#include <new>

#include "common.h"
#include "ptx.cuh"

namespace vsb {

struct IgemmParams {
  int m_total, to, ho, wo;
  int st, sh, sw;
  int lt, lh, lw;  // lower corner = -leading pad
  int kh, kw;
  int cin, cin_chunks, total_chunks, cps, kchunk;
  int block_n, n_tiles, stages;
  uint32_t idesc, tmem_cols;
  const float* scale;
  const float* bias;
  const __nv_bfloat16* residual;
  long long res_pitch;
  __nv_bfloat16* out;
  long long out_pitch;
  int relu;
};

constexpr int kBlockM = 128;
constexpr int kThreads = 192;

__global__ void __launch_bounds__(kThreads)
conv_igemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const IgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const uint32_t row_bytes = p.kchunk * 2;
  const uint32_t a_chunk_bytes = kBlockM * row_bytes;
  const uint32_t b_chunk_bytes = p.block_n * row_bytes;
  const uint32_t stage_bytes = p.cps * (a_chunk_bytes + b_chunk_bytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* accum_bar = empty_bar + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;  // warp-uniform
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x % p.n_tiles;
  const int m_tile = blockIdx.x / p.n_tiles;
  const int m0 = m_tile * kBlockM;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_kstages = (p.total_chunks + p.cps - 1) / p.cps;

  if (warp == 4) {
    if (lane == 0) {
      // ------------------------------------------------------ TMA producer
      const int wo = m0 % p.wo;
      const int r1 = m0 / p.wo;
      const int ho = r1 % p.ho;
      const int r2 = r1 / p.ho;
      const int to_ = r2 % p.to;
      const int n0 = r2 / p.to;
      const int w0 = wo * p.sw + p.lw;
      const int h0 = ho * p.sh + p.lh;
      const int d0 = to_ * p.st + p.lt;
      for (int ks = 0; ks < num_kstages; ++ks) {
        const int slot = ks % p.stages;
        const uint32_t parity = ((ks / p.stages) & 1) ^ 1;
        mbar_wait(&empty_bar[slot], parity);
        const int g0 = ks * p.cps;
        const int nch = min(p.cps, p.total_chunks - g0);
        mbar_expect_tx(&full_bar[slot], nch * (a_chunk_bytes + b_chunk_bytes));
        uint8_t* a_dst = smem + (size_t)slot * stage_bytes;
        uint8_t* b_dst = a_dst + p.cps * a_chunk_bytes;
        for (int c = 0; c < nch; ++c) {
          const int g = g0 + c;
          const int tap = g / p.cin_chunks;
          const int cc = g - tap * p.cin_chunks;
          const int kw_ = tap % p.kw;
          const int r = tap / p.kw;
          const int kh_ = r % p.kh;
          const int kt_ = r / p.kh;
          tma_load_im2col_5d(a_dst + c * a_chunk_bytes, &map_a, &full_bar[slot], cc * p.kchunk, w0, h0, d0, n0,
                             (uint16_t)kw_, (uint16_t)kh_, (uint16_t)kt_);
          tma_load_2d(b_dst + c * b_chunk_bytes, &map_b, &full_bar[slot], tap * p.cin + cc * p.kchunk,
                      n_tile * p.block_n);
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      // -------------------------------------------------------- MMA issuer
      uint32_t accumulate = 0;
      const int kk = p.kchunk >> 4;
      for (int ks = 0; ks < num_kstages; ++ks) {
        const int slot = ks % p.stages;
        const uint32_t parity = (ks / p.stages) & 1;
        mbar_wait(&full_bar[slot], parity);
        tc_fence_after();
        const int nch = min(p.cps, p.total_chunks - ks * p.cps);
        const uint32_t a0 = smem_u32(smem + (size_t)slot * stage_bytes);
        const uint32_t b0 = a0 + p.cps * a_chunk_bytes;
        for (int c = 0; c < nch; ++c) {
          for (int k = 0; k < kk; ++k) {
            const uint64_t adesc = umma_smem_desc(a0 + c * a_chunk_bytes + k * 32, row_bytes);
            const uint64_t bdesc = umma_smem_desc(b0 + c * b_chunk_bytes + k * 32, row_bytes);
            umma_bf16(tmem_base, adesc, bdesc, p.idesc, accumulate);
            accumulate = 1;
          }
        }
        umma_commit(&empty_bar[slot]);  // frees the smem slot once these MMAs retire
      }
      umma_commit(accum_bar);  // accumulator complete
    }
  } else {
    // ---------------------------------------------------------- epilogue
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const long long row = (long long)m0 + warp * 32 + lane;
    const bool valid = row < p.m_total;
    const int nbase = n_tile * p.block_n;
    __nv_bfloat16* orow = p.out + row * p.out_pitch + nbase;
    const __nv_bfloat16* rrow = p.residual ? p.residual + row * p.res_pitch + nbase : nullptr;
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int j0 = 0; j0 < p.block_n; j0 += 16) {
      uint32_t v[16];
      tmem_ld16(taddr + j0, v);
      uint4 r0 = make_uint4(0, 0, 0, 0), r1 = make_uint4(0, 0, 0, 0);
      if (rrow && valid) {
        r0 = *reinterpret_cast<const uint4*>(rrow + j0);
        r1 = *reinterpret_cast<const uint4*>(rrow + j0 + 8);
      }
      float sc[16], bi[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.scale + nbase + j0) + q);
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + nbase + j0) + q);
        sc[4 * q + 0] = s4.x; sc[4 * q + 1] = s4.y; sc[4 * q + 2] = s4.z; sc[4 * q + 3] = s4.w;
        bi[4 * q + 0] = b4.x; bi[4 * q + 1] = b4.y; bi[4 * q + 2] = b4.z; bi[4 * q + 3] = b4.w;
      }
      tmem_ld_wait();
      const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
      uint32_t o[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float x0 = fmaf(__uint_as_float(v[2 * q]), sc[2 * q], bi[2 * q]) + bf16_lo(rr[q]);
        float x1 = fmaf(__uint_as_float(v[2 * q + 1]), sc[2 * q + 1], bi[2 * q + 1]) + bf16_hi(rr[q]);
        if (p.relu) {
          x0 = fmaxf(x0, 0.f);
          x1 = fmaxf(x1, 0.f);
        }
        o[q] = pack_bf16x2(x0, x1);
      }
      if (valid) {
        *reinterpret_cast<uint4*>(orow + j0) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4*>(orow + j0 + 8) = make_uint4(o[4], o[5], o[6], o[7]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn g_encode_tiled = nullptr;
static EncodeIm2colFn g_encode_im2col = nullptr;
static int g_driver_version = 0;

static int load_driver_entry_points() {
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

static CUtensorMapSwizzle swizzle_for(int row_bytes) {
  return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                             : (row_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE));
}

// (C,W,H,D,N) im2col map over a bf16 NTHWC tensor.
static int encode_im2col_map(CUtensorMap* map, const void* base, int n, int t, int h, int w, int c, int pitch,
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

struct vsb_conv_plan {
  vsb_conv_desc desc;
  int to, ho, wo;
  long long m_total;
  // bf16 tensor-core path
  CUtensorMap map_a, map_b;
  vsb::IgemmParams params;
  size_t smem_bytes;
  unsigned grid;
};

using namespace vsb;

extern "C" int vsb_conv3d_plan_create(const vsb_conv_desc* d, vsb_conv_plan** out_plan) {
  VSB_CHECK_ARG(d && out_plan, "null argument");
  *out_plan = nullptr;
  VSB_CHECK_ARG(d->dtype == VSB_BF16 || d->dtype == VSB_F32, "dtype must be VSB_BF16 or VSB_F32");
  VSB_CHECK_ARG(d->in && d->wgt && d->out && d->scale && d->bias, "null tensor pointer");
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

  if (d->dtype == VSB_F32) {
    *out_plan = plan;
    return VSB_OK;
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
  int block_n = d->block_n;
  if (!block_n) {
    block_n = 256;
    while (block_n > 16 && d->cout % block_n) block_n >>= 1;
  }
  if (block_n < 16 || block_n > 256 || block_n % 16 || d->cout % block_n) FAIL(VSB_ERR_INVALID, "bad block_n %d for cout %d", block_n, d->cout);
  const int taps = d->kt * d->kh * d->kw;
  const int cin_chunks = d->cin / kchunk;
  const int total_chunks = taps * cin_chunks;
  int cps = 64 / kchunk;
  if (cps > total_chunks) cps = total_chunks;
  const int stage_bytes = cps * (kBlockM + block_n) * kchunk * 2;
  int stages = d->stages;
  if (!stages) {
    const int small_budget = 110 * 1024;  // 2 CTAs / SM
    if (4 * stage_bytes <= small_budget) stages = 4;
    else if (3 * stage_bytes <= small_budget) stages = 3;
    else stages = 4;
  }
  const int num_kstages = ceil_div(total_chunks, cps);
  if (stages > num_kstages) stages = num_kstages;
  if (stages < 1) stages = 1;
  const size_t smem_bytes = (size_t)stages * stage_bytes + (2 * stages + 1) * 8 + 16 + 1024;
  if (smem_bytes > 227 * 1024) FAIL(VSB_ERR_INVALID, "pipeline needs %zu bytes of shared memory", smem_bytes);

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
  const long long k_total = (long long)taps * d->cin;
  rc = encode_tiled_2d(&plan->map_b, d->wgt, k_total, d->cout, k_total * 2, kchunk, block_n, swz);
  if (rc != VSB_OK) {
    delete plan;
    return rc;
  }
#undef FAIL

  IgemmParams& p = plan->params;
  p.m_total = (int)m_total;
  p.to = to; p.ho = ho; p.wo = wo;
  p.st = d->st; p.sh = d->sh; p.sw = d->sw;
  p.lt = lower[2]; p.lh = lower[1]; p.lw = lower[0];
  p.kh = d->kh; p.kw = d->kw;
  p.cin = d->cin; p.cin_chunks = cin_chunks; p.total_chunks = total_chunks; p.cps = cps; p.kchunk = kchunk;
  p.block_n = block_n; p.n_tiles = d->cout / block_n; p.stages = stages;
  p.idesc = umma_idesc_bf16(kBlockM, block_n);
  p.tmem_cols = 32;  // TMEM allocations are powers of two >= 32 columns
  while (p.tmem_cols < (uint32_t)block_n) p.tmem_cols <<= 1;
  p.scale = d->scale; p.bias = d->bias;
  p.residual = static_cast<const __nv_bfloat16*>(d->residual);
  p.res_pitch = d->res_pitch;
  p.out = static_cast<__nv_bfloat16*>(d->out);
  p.out_pitch = d->out_pitch;
  p.relu = d->relu;
  plan->smem_bytes = smem_bytes;
  plan->grid = (unsigned)(ceil_div_ll(m_total, kBlockM) * p.n_tiles);
  plan->desc.block_n = block_n;
  plan->desc.kchunk = kchunk;
  plan->desc.stages = stages;

  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
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
  conv_igemm_kernel<<<plan->grid, kThreads, plan->smem_bytes, s>>>(plan->map_a, plan->map_b, plan->params);
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
  return 2.0 * (double)plan->m_total * d.cout * d.kt * d.kh * d.kw * d.cin;
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
