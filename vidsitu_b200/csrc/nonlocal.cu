// Non-local block attention (I3D-NLN): out = normalise(theta . phi^T) . g per clip.
// Reference: SlowFast/slowfast/models/nonlocal_helper.py:123-141 (two einsums and a
// softmax over the key axis, or division by the key count for "dot_product").
// CUDA-core kernel, fp32 accumulation, the [QT x tk] score tile lives in shared
// memory so the 3136 x 784 score matrix is never written to HBM.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <float.h>

#include "common.h"

namespace vsb {

constexpr int QT = 16;  // queries per CTA

template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T>
__device__ __forceinline__ void stf(T* p, float v);
template <>
__device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename T>
__global__ void __launch_bounds__(256)
nonlocal_attention_kernel(const T* __restrict__ theta, int theta_pitch, const T* __restrict__ phi, int phi_pitch,
                          const T* __restrict__ g, int g_pitch, T* __restrict__ out, int out_pitch, int tq, int tk,
                          int c, int softmax) {
  extern __shared__ float sm[];
  float* q_s = sm;               // [QT][c]
  float* s_s = sm + QT * c;      // [QT][tk]
  const int clip = blockIdx.y;
  const int q0 = blockIdx.x * QT;
  const int nq = min(QT, tq - q0);
  const T* th = theta + ((long long)clip * tq + q0) * theta_pitch;
  const T* ph = phi + (long long)clip * tk * phi_pitch;
  const T* gg = g + (long long)clip * tk * g_pitch;

  for (int i = threadIdx.x; i < QT * c; i += blockDim.x) {
    const int q = i / c, ch = i - q * c;
    q_s[i] = q < nq ? ldf<T>(th + (long long)q * theta_pitch + ch) : 0.f;
  }
  __syncthreads();

  // scores: one thread per key, QT dot products of length c
  const float scale = softmax ? rsqrtf((float)c) : 1.0f / (float)tk;
  for (int k = threadIdx.x; k < tk; k += blockDim.x) {
    float acc[QT];
#pragma unroll
    for (int q = 0; q < QT; ++q) acc[q] = 0.f;
    const T* kr = ph + (long long)k * phi_pitch;
    for (int ch = 0; ch < c; ++ch) {
      const float kv = ldf<T>(kr + ch);
#pragma unroll
      for (int q = 0; q < QT; ++q) acc[q] = fmaf(q_s[q * c + ch], kv, acc[q]);
    }
#pragma unroll
    for (int q = 0; q < QT; ++q) s_s[q * tk + k] = acc[q] * scale;
  }
  __syncthreads();

  if (softmax) {
    // one warp per query row
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int q = warp; q < QT; q += 8) {
      float* row = s_s + q * tk;
      float mx = -FLT_MAX;
      for (int k = lane; k < tk; k += 32) mx = fmaxf(mx, row[k]);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      float sum = 0.f;
      for (int k = lane; k < tk; k += 32) {
        const float e = expf(row[k] - mx);
        row[k] = e;
        sum += e;
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
      const float inv = 1.0f / sum;
      for (int k = lane; k < tk; k += 32) row[k] *= inv;
    }
    __syncthreads();
  }

  // out[q][ch] = sum_k P[q][k] * g[k][ch]; one thread per channel
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float acc[QT];
#pragma unroll
    for (int q = 0; q < QT; ++q) acc[q] = 0.f;
    for (int k = 0; k < tk; ++k) {
      const float gv = ldf<T>(gg + (long long)k * g_pitch + ch);
#pragma unroll
      for (int q = 0; q < QT; ++q) acc[q] = fmaf(s_s[q * tk + k], gv, acc[q]);
    }
    for (int q = 0; q < nq; ++q)
      stf<T>(out + ((long long)clip * tq + q0 + q) * out_pitch + ch, acc[q]);
  }
}

// ---- tensor-core route (bf16): the two einsums run as per-clip 1x1x1 "convs" on the tcgen05 kernels
// (scores = theta_i . phi_i^T with phi_i as the weight matrix, out = P_i . g_i with g_i^T as the weight
// matrix); the two kernels below are the glue: row normalisation of the score matrix and the g transpose.

// One warp per score row, in place: the scores arrive as IEEE half (vsb_conv_desc.out_f16) and leave as bf16
// probabilities (the A operand of the second GEMM).  cols [0, valid) <- softmax (fp32) or the value itself,
// cols [valid, width) <- 0 (the K padding of the second GEMM; the scores there were computed against rows
// that are not this clip's keys).
__global__ void __launch_bounds__(256)
score_rows_kernel(void* __restrict__ s, long long rows, int valid, int width, int pitch, int softmax) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  __half2* in2 = reinterpret_cast<__half2*>(s) + row * (pitch >> 1);  // pitch, valid, width are even
  __nv_bfloat162* out2 = reinterpret_cast<__nv_bfloat162*>(in2);
  const int nv2 = valid >> 1, nw2 = width >> 1;
  float mx = 0.f, inv = 1.f;
  if (softmax) {
    mx = -FLT_MAX;
    for (int i = lane; i < nv2; i += 32) {
      const float2 v = __half22float2(in2[i]);
      mx = fmaxf(mx, fmaxf(v.x, v.y));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int i = lane; i < nv2; i += 32) {
      const float2 v = __half22float2(in2[i]);
      sum += __expf(v.x - mx) + __expf(v.y - mx);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    inv = 1.f / sum;
  }
  for (int i = lane; i < nv2; i += 32) {  // element i is read and written by the same lane
    const float2 v = __half22float2(in2[i]);
    out2[i] = softmax ? __floats2bfloat162_rn(__expf(v.x - mx) * inv, __expf(v.y - mx) * inv)
                      : __floats2bfloat162_rn(v.x, v.y);
  }
  const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
  for (int i = nv2 + lane; i < nw2; i += 32) out2[i] = z;
}

// out[clip][ch][k] = in[clip][k][ch] (k < rows), 0 for rows <= k < out_pitch: g_i^T as a K-major weight matrix.
__global__ void __launch_bounds__(256)
transpose_pad_kernel(const __nv_bfloat16* __restrict__ in, int in_pitch, __nv_bfloat16* __restrict__ out, int rows,
                     int cols, int out_pitch) {
  __shared__ __nv_bfloat16 tile[32][33];
  const int clip = blockIdx.z;
  const int k0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const __nv_bfloat16* src = in + (long long)clip * rows * in_pitch;
  __nv_bfloat16* dst = out + (long long)clip * cols * out_pitch;
  for (int j = ty; j < 32; j += 8) {
    const int k = k0 + j, c = c0 + tx;
    tile[j][tx] = (k < rows && c < cols) ? src[(long long)k * in_pitch + c] : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, k = k0 + tx;
    if (c < cols && k < out_pitch) dst[(long long)c * out_pitch + k] = tile[tx][j];
  }
}

}  // namespace vsb

using namespace vsb;

extern "C" int vsb_score_rows(void* scores, long long rows, int valid, int width, int pitch, int softmax,
                              void* stream) {
  VSB_CHECK_ARG(scores, "null argument");
  VSB_CHECK_ARG(rows > 0 && valid > 0 && valid <= width && width <= pitch, "bad extent");
  VSB_CHECK_ARG(!(valid & 1) && !(width & 1) && !(pitch & 1), "valid, width and pitch must be even");
  VSB_CHECK_ARG((rows + 7) / 8 < (1ll << 31), "too many rows");
  score_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(scores, rows, valid, width,
                                                                                            pitch, softmax);
  VSB_CHECK_LAUNCH("score_rows_kernel");
  return VSB_OK;
}

extern "C" int vsb_transpose_pad(const void* in, int in_pitch, void* out, int n, int rows, int cols, int out_pitch,
                                 void* stream) {
  VSB_CHECK_ARG(in && out, "null argument");
  VSB_CHECK_ARG(n > 0 && n <= 65535 && rows > 0 && cols > 0 && in_pitch >= cols && out_pitch >= rows, "bad extent");
  dim3 grid(ceil_div(out_pitch, 32), ceil_div(cols, 32), n);
  transpose_pad_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), in_pitch, static_cast<__nv_bfloat16*>(out), rows, cols, out_pitch);
  VSB_CHECK_LAUNCH("transpose_pad_kernel");
  return VSB_OK;
}

extern "C" int vsb_nonlocal_attention(const void* theta, int theta_pitch, const void* phi, int phi_pitch,
                                      const void* g, int g_pitch, void* out, int out_pitch, int n, int tq, int tk,
                                      int c, int softmax, int dtype, void* stream) {
  VSB_CHECK_ARG(theta && phi && g && out, "null argument");
  VSB_CHECK_ARG(dtype == VSB_BF16 || dtype == VSB_F32, "bad dtype");
  VSB_CHECK_ARG(n > 0 && n <= 65535 && tq > 0 && tk > 0 && c > 0, "bad extent");
  VSB_CHECK_ARG(theta_pitch >= c && phi_pitch >= c && g_pitch >= c && out_pitch >= c, "pitch below channel count");
  const size_t smem = (size_t)QT * (c + tk) * sizeof(float);
  VSB_CHECK_ARG(smem <= 200 * 1024, "score tile (%zu bytes) exceeds shared memory", smem);
  dim3 grid(ceil_div(tq, QT), n);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == VSB_BF16) {
    VSB_CHECK_CUDA(cudaFuncSetAttribute(nonlocal_attention_kernel<__nv_bfloat16>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    nonlocal_attention_kernel<__nv_bfloat16><<<grid, 256, smem, s>>>(
        static_cast<const __nv_bfloat16*>(theta), theta_pitch, static_cast<const __nv_bfloat16*>(phi), phi_pitch,
        static_cast<const __nv_bfloat16*>(g), g_pitch, static_cast<__nv_bfloat16*>(out), out_pitch, tq, tk, c,
        softmax);
  } else {
    VSB_CHECK_CUDA(cudaFuncSetAttribute(nonlocal_attention_kernel<float>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    nonlocal_attention_kernel<float><<<grid, 256, smem, s>>>(
        static_cast<const float*>(theta), theta_pitch, static_cast<const float*>(phi), phi_pitch,
        static_cast<const float*>(g), g_pitch, static_cast<float*>(out), out_pitch, tq, tk, c, softmax);
  }
  VSB_CHECK_LAUNCH("nonlocal_attention_kernel");
  return VSB_OK;
}
