// fp32 verification path of the conv op: the same contract as the tensor-core
// kernel (include/vidsitu_b200.h, vsb_conv_desc) computed on the CUDA cores with
// fp32 storage and FFMA accumulation.  It exists so that "fp32 runs" of the
// network (bit-exact top-5 verb sets against the reference's fp32 forward,
// vidsitu_code/evl_vsitu.py:39-75) do not go through bf16/tensor-core rounding.
// Plain smem-tiled direct convolution: 64 output pixels x 64 output channels per
// CTA, 4x4 micro-tile per thread, K walked as (tap, 16-channel slab).
#include "common.h"

namespace vsb {

struct SimtParams {
  const float* in;
  const float* wgt;
  const float* scale;
  const float* bias;
  const float* residual;
  float* out;
  int n, t, h, w, cin, in_pitch;
  int cout, kt, kh, kw, st, sh, sw, pt, ph, pw;
  int to, ho, wo;
  long long m_total;
  int res_pitch, out_pitch, relu;
};

constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256) conv_simt_kernel(const SimtParams p) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;

  // loader mapping: 4 threads per row, 4 consecutive channels each
  const int lrow = tid >> 2;
  const int lc = (tid & 3) * 4;
  const long long lm = m0 + lrow;
  const bool lvalid = lm < p.m_total;
  int wo = 0, ho = 0, to_ = 0, nn = 0;
  if (lvalid) {
    long long r = lm;
    wo = (int)(r % p.wo); r /= p.wo;
    ho = (int)(r % p.ho); r /= p.ho;
    to_ = (int)(r % p.to); r /= p.to;
    nn = (int)r;
  }
  const int lco = n0 + lrow;  // weight row loaded by this thread
  const int taps = p.kt * p.kh * p.kw;

  // compute mapping: 16 x 16 threads, 4 x 4 outputs each
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4] = {};

  for (int tap = 0; tap < taps; ++tap) {
    const int kw_ = tap % p.kw;
    const int kh_ = (tap / p.kw) % p.kh;
    const int kt_ = tap / (p.kw * p.kh);
    const int iw = wo * p.sw - p.pw + kw_;
    const int ih = ho * p.sh - p.ph + kh_;
    const int it = to_ * p.st - p.pt + kt_;
    const bool inb = lvalid && iw >= 0 && iw < p.w && ih >= 0 && ih < p.h && it >= 0 && it < p.t;
    const float* arow = p.in + ((((long long)nn * p.t + it) * p.h + ih) * p.w + iw) * p.in_pitch;
    const float* brow = p.wgt + ((long long)lco * taps + tap) * p.cin;
    for (int c0 = 0; c0 < p.cin; c0 += TK) {
      float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = make_float4(0.f, 0.f, 0.f, 0.f);
      const int c = c0 + lc;
      if (inb && c < p.cin) av = *reinterpret_cast<const float4*>(arow + c);
      if (lco < p.cout && c < p.cin) bv = *reinterpret_cast<const float4*>(brow + c);
      __syncthreads();
      As[lc + 0][lrow] = av.x; As[lc + 1][lrow] = av.y; As[lc + 2][lrow] = av.z; As[lc + 3][lrow] = av.w;
      Bs[lc + 0][lrow] = bv.x; Bs[lc + 1][lrow] = bv.y; Bs[lc + 2][lrow] = bv.z; Bs[lc + 3][lrow] = bv.w;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < TK; ++k) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= p.m_total) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co >= p.cout) continue;
      float x = fmaf(acc[i][j], p.scale[co], p.bias[co]);
      if (p.residual) x += p.residual[m * p.res_pitch + co];
      if (p.relu) x = fmaxf(x, 0.f);
      p.out[m * p.out_pitch + co] = x;
    }
  }
}

int launch_conv_simt(const vsb_conv_desc& d, int to, int ho, int wo, cudaStream_t stream) {
  VSB_CHECK_ARG(d.cin % 4 == 0 && d.in_pitch % 4 == 0, "fp32 path needs cin and in_pitch multiples of 4");
  SimtParams p;
  p.in = static_cast<const float*>(d.in);
  p.wgt = static_cast<const float*>(d.wgt);
  p.scale = d.scale; p.bias = d.bias;
  p.residual = static_cast<const float*>(d.residual);
  p.out = static_cast<float*>(d.out);
  p.n = d.n; p.t = d.t; p.h = d.h; p.w = d.w; p.cin = d.cin; p.in_pitch = d.in_pitch;
  p.cout = d.cout; p.kt = d.kt; p.kh = d.kh; p.kw = d.kw; p.st = d.st; p.sh = d.sh; p.sw = d.sw;
  p.pt = d.pt_lo; p.ph = d.ph_lo; p.pw = d.pw_lo;
  p.to = to; p.ho = ho; p.wo = wo;
  p.m_total = (long long)d.n * to * ho * wo;
  p.res_pitch = d.res_pitch; p.out_pitch = d.out_pitch; p.relu = d.relu;
  dim3 grid((unsigned)ceil_div_ll(p.m_total, TM), (unsigned)ceil_div(d.cout, TN));
  conv_simt_kernel<<<grid, 256, 0, stream>>>(p);
  VSB_CHECK_LAUNCH("conv_simt_kernel");
  return VSB_OK;
}

}  // namespace vsb
