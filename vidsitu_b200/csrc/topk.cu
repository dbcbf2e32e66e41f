// Verb prediction tail of EvalB.forward_one_batch (vidsitu_code/evl_vsitu.py:39-75):
//   probs = softmax(mdl_out, -1); probs.sort(descending=True); keep the first topk_save = 5 ids and scores.
// One CTA per event clip: block-wide max and sum(exp) reductions for the softmax, then k rounds of a
// block-wide arg-max over the candidates that come after the previous pick in (logit descending, index
// ascending) order -- the order of a stable descending sort.  exp() is monotone, so ranking the logits ranks the
// probabilities; only the k selected probabilities are ever formed.  V ~ 1.5 k floats per row: the row stays in
// L1/registers, the kernel is launch-latency sized; it exists so that validation needs a D2H of 5 ids + 5 scores
// per clip instead of the [N, V] logits.
#include "common.h"

namespace vsb {

constexpr int TOPK_THREADS = 256;
constexpr int TOPK_MAX_K = 16;

struct Cand {
  float v;
  int i;
};

__device__ __forceinline__ bool cand_before(const Cand& a, const Cand& b) {  // a ranks ahead of b
  return a.v > b.v || (a.v == b.v && a.i < b.i);
}

__device__ __forceinline__ Cand block_best(Cand c, Cand* sh) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    Cand o;
    o.v = __shfl_xor_sync(0xffffffffu, c.v, off);
    o.i = __shfl_xor_sync(0xffffffffu, c.i, off);
    if (cand_before(o, c)) c = o;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // sh is reused between rounds
  if (lane == 0) sh[warp] = c;
  __syncthreads();
  Cand r = sh[0];
#pragma unroll
  for (int w = 1; w < TOPK_THREADS / 32; ++w)
    if (cand_before(sh[w], r)) r = sh[w];
  return r;
}

__device__ __forceinline__ float block_sum(float v, float* sh) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int w = 0; w < TOPK_THREADS / 32; ++w) r += sh[w];  // fixed order: deterministic
  return r;
}

__global__ void __launch_bounds__(TOPK_THREADS)
softmax_topk_kernel(const float* __restrict__ logits, int v, int pitch, int k, int* __restrict__ out_idx,
                    float* __restrict__ out_prob) {
  __shared__ Cand sh_c[TOPK_THREADS / 32];
  __shared__ float sh_f[TOPK_THREADS / 32];
  const float* row = logits + (long long)blockIdx.x * pitch;
  const float ninf = -__int_as_float(0x7f800000);
  Cand best{ninf, 0x7fffffff};
  for (int i = threadIdx.x; i < v; i += TOPK_THREADS) {
    const Cand c{row[i], i};
    if (cand_before(c, best)) best = c;
  }
  Cand pick = block_best(best, sh_c);
  const float vmax = pick.v;
  float s = 0.f;
  for (int i = threadIdx.x; i < v; i += TOPK_THREADS) s += expf(row[i] - vmax);
  const float denom = block_sum(s, sh_f);
  for (int r = 0;; ++r) {
    if (threadIdx.x == 0) {
      out_idx[(long long)blockIdx.x * k + r] = pick.i;
      out_prob[(long long)blockIdx.x * k + r] = expf(pick.v - vmax) / denom;
    }
    if (r + 1 == k) break;
    Cand nb{ninf, 0x7fffffff};
    for (int i = threadIdx.x; i < v; i += TOPK_THREADS) {
      const Cand c{row[i], i};
      if (cand_before(pick, c) && cand_before(c, nb)) nb = c;
    }
    pick = block_best(nb, sh_c);
  }
}

}  // namespace vsb

using namespace vsb;

extern "C" int vsb_softmax_topk(const float* logits, int n, int v, int pitch, int k, int* idx, float* prob,
                                void* stream) {
  VSB_CHECK_ARG(logits && idx && prob, "null argument");
  VSB_CHECK_ARG(n > 0 && v > 0 && pitch >= v, "bad extent");
  VSB_CHECK_ARG(k >= 1 && k <= TOPK_MAX_K && k <= v, "k must be in [1, min(16, v)]");
  softmax_topk_kernel<<<n, TOPK_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(logits, v, pitch, k, idx, prob);
  VSB_CHECK_LAUNCH("softmax_topk_kernel");
  return VSB_OK;
}
