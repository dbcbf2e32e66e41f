// Debug probes that pin tcgen05.mma behaviour on real hardware (declared in
// include/vidsitu_b200_debug.h; driven by tools/gpu_probe_umma.py):
//   * semantics of shared-memory descriptors whose start address is shifted by a
//     whole number of rows / a K slice inside a swizzled tile (what the window conv
//     kernel relies on to express filter taps without re-loading the input);
//   * issue rate of M=128 MMAs as a function of N (shared-memory operand bound).
#include <cuda.h>

#include "common.h"
#include "ptx.cuh"

namespace vsb {

struct ProbeParams {
  int n;            // MMA N (16..256)
  int ksteps;       // K = 16 * ksteps
  int a_rows;       // rows of A resident in smem (>= 128 + shift)
  int shift_bytes;  // added to the A start address (rows * row_bytes + K-slice bytes)
  int b_shift_bytes;
  int row_bytes;    // 32 / 64 / 128
  int base_mode;    // 0: base_offset = 0; 1: base_offset = (addr >> 7) & 7
  int k_stride_bytes;  // start-address advance per K step (32)
};

__device__ __forceinline__ uint64_t probe_desc(uint32_t addr, uint32_t row_bytes, int base_mode) {
  uint64_t d = umma_smem_desc(addr, row_bytes);
  if (base_mode == 1) d |= (uint64_t)((addr >> 7) & 7u) << 49;
  return d;
}

__global__ void __launch_bounds__(128) umma_semantics_kernel(const __grid_constant__ CUtensorMap map_a,
                                                             const __grid_constant__ CUtensorMap map_b,
                                                             const ProbeParams p, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, done;
  __shared__ uint32_t tmem_slot;
  const uint32_t a_bytes = (uint32_t)p.a_rows * p.row_bytes;
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + ((a_bytes + 1023) & ~1023u);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_init(&done, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, a_bytes + (uint32_t)p.n * p.row_bytes);
    for (int r = 0; r < p.a_rows; r += 128) tma_load_2d(a_s + (size_t)r * p.row_bytes, &map_a, &bar, 0, r);
    tma_load_2d(b_s, &map_b, &bar, 0, 0);
    mbar_wait(&bar, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, p.n);
    for (int k = 0; k < p.ksteps; ++k) {
      const uint64_t ad = probe_desc(smem_u32(a_s) + p.shift_bytes + k * p.k_stride_bytes, p.row_bytes, p.base_mode);
      const uint64_t bd = probe_desc(smem_u32(b_s) + p.b_shift_bytes + k * p.k_stride_bytes, p.row_bytes, p.base_mode);
      umma_bf16(tmem, ad, bd, idesc, k ? 1u : 0u);
    }
    umma_commit(&done);
  }
  __syncthreads();
  mbar_wait(&done, 0);
  tc_fence_after();
  for (int j0 = 0; j0 < p.n; j0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + j0, v);
    tmem_ld_wait();
    for (int e = 0; e < 16; ++e) out[(size_t)threadIdx.x * p.n + j0 + e] = __uint_as_float(v[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// Issue-rate probe: every CTA issues `iters` x 4 MMAs (M=128, N=n, K=16) whose A operand cycles
// through `a_tiles` different 16 KiB tiles, and reports the cycles from first issue to completion.
__global__ void __launch_bounds__(128) umma_rate_kernel(int n, int iters, int a_tiles, int a_from_same, long long* clk) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  // operands: whatever is in shared memory (zero-fill so no NaN slow paths)
  const uint32_t total = (uint32_t)a_tiles * 16384u + 256u * 128u;
  for (uint32_t i = threadIdx.x * 16; i < total; i += blockDim.x * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&done, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, n);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + (uint32_t)a_tiles * 16384u;
    long long t0 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const uint32_t a = a0 + (a_from_same ? 0u : (uint32_t)(it % a_tiles) * 16384u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem, umma_smem_desc(a + 32 * k, 128), umma_smem_desc(b0 + 32 * k, 128), idesc, 1u);
      }
      umma_commit(&done);
    }
    __syncwarp();
    mbar_wait(&done, 0);
    if (threadIdx.x == 0) clk[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}


// Issue-rate probe 2: `groups` groups of `per_group` MMAs (M=128, N=n, K=16), a tcgen05.commit after every group
// (commit_each != 0) or only at the end; the A operand has `row_bytes` rows and its start address is shifted by
// `shift_rows` rows (+1 row per MMA inside a group when walk != 0, like the taps of a 3x3 window).  Reports the
// cycles of the issuing thread until its last instruction has been issued (clk[2*cta]) and until completion
// (clk[2*cta+1]).
__global__ void __launch_bounds__(128) umma_rate2_kernel(int n, int groups, int per_group, int commit_each, int row_bytes,
                                                         int shift_rows, int walk, long long* clk) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t done, sink;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t total = 3u * 16384u + 256u * 128u;
  for (uint32_t i = threadIdx.x * 16; i < total; i += blockDim.x * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&done, 1);
    mbar_init(&sink, 1u << 20);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, n);
    const uint32_t a0 = smem_u32(smem) + (uint32_t)shift_rows * row_bytes, b0 = smem_u32(smem) + 3u * 16384u;
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int g = 0; g < groups; ++g) {
        for (int k = 0; k < per_group; ++k)
          umma_bf16(tmem + (g & 1) * n, umma_smem_desc(a0 + (walk ? (uint32_t)k * row_bytes : 0u), row_bytes),
                    umma_smem_desc(b0, row_bytes), idesc, k != 0 ? 1u : 0u);
        if (commit_each) umma_commit(&sink);
      }
      umma_commit(&done);
      t1 = clock64();
    }
    __syncwarp();
    mbar_wait(&done, 0);
    const long long t2 = clock64();
    if (t0) {
      clk[2 * blockIdx.x] = t1 - t0;
      clk[2 * blockIdx.x + 1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int probe_encode_2d(CUtensorMap* map, const void* base, int cols, int rows, int box_rows, int row_bytes) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    set_error("cuTensorMapEncodeTiled not available");
    return VSB_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle swz = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                                  : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = reinterpret_cast<EncodeTiledFn>(fn)(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base),
                                                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("probe cuTensorMapEncodeTiled failed (%d)", (int)r);
    return VSB_ERR_CUDA;
  }
  return VSB_OK;
}

}  // namespace vsb

using namespace vsb;

// a: bf16 [a_rows, row_bytes/2] row-major; b: bf16 [n, row_bytes/2]; out: fp32 [128, n]
extern "C" int vsb_debug_umma_semantics(const void* a, int a_rows, const void* b, int n, int row_bytes, int ksteps,
                                        int shift_bytes, int b_shift_bytes, int base_mode, float* out, void* stream) {
  VSB_CHECK_ARG(a && b && out, "null pointer");
  VSB_CHECK_ARG(row_bytes == 32 || row_bytes == 64 || row_bytes == 128, "row_bytes must be 32/64/128");
  VSB_CHECK_ARG(n >= 16 && n <= 256 && n % 16 == 0, "bad n");
  VSB_CHECK_ARG(a_rows >= 128 && a_rows % 128 == 0 && a_rows <= 512, "a_rows must be 128..512, multiple of 128");
  CUtensorMap ma, mb;
  int rc = probe_encode_2d(&ma, a, row_bytes / 2, a_rows, 128, row_bytes);
  if (rc != VSB_OK) return rc;
  rc = probe_encode_2d(&mb, b, row_bytes / 2, n, n, row_bytes);
  if (rc != VSB_OK) return rc;
  ProbeParams p;
  p.n = n; p.ksteps = ksteps; p.a_rows = a_rows; p.shift_bytes = shift_bytes; p.b_shift_bytes = b_shift_bytes;
  p.row_bytes = row_bytes; p.base_mode = base_mode; p.k_stride_bytes = 32;
  const size_t smem = (size_t)a_rows * row_bytes + (size_t)n * row_bytes + 4096;
  VSB_CHECK_CUDA(cudaFuncSetAttribute(umma_semantics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  umma_semantics_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(ma, mb, p, out);
  VSB_CHECK_LAUNCH("umma_semantics_kernel");
  return VSB_OK;
}

// clk: int64 [grid] cycles for iters*4 MMAs per CTA
extern "C" int vsb_debug_umma_rate(int n, int iters, int a_tiles, int a_from_same, int grid, int smem_pad_kb,
                                   long long* clk, void* stream) {
  VSB_CHECK_ARG(clk && n >= 16 && n <= 256 && n % 16 == 0 && a_tiles >= 1 && a_tiles <= 8, "bad argument");
  const size_t smem = (size_t)a_tiles * 16384 + 256 * 128 + 2048 + (size_t)smem_pad_kb * 1024;
  VSB_CHECK_ARG(smem <= 224 * 1024, "too much shared memory");
  VSB_CHECK_CUDA(cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  umma_rate_kernel<<<grid, 128, smem, static_cast<cudaStream_t>(stream)>>>(n, iters, a_tiles, a_from_same, clk);
  VSB_CHECK_LAUNCH("umma_rate_kernel");
  return VSB_OK;
}

// clk: int64 [2 * grid]: (issue cycles, completion cycles) per CTA
extern "C" int vsb_debug_umma_rate2(int n, int groups, int per_group, int commit_each, int row_bytes, int shift_rows,
                                    int walk, int grid, long long* clk, void* stream) {
  VSB_CHECK_ARG(clk && n >= 16 && n <= 128 && n % 16 == 0 && groups > 0 && per_group > 0 && per_group <= 64, "bad argument");
  VSB_CHECK_ARG(row_bytes == 32 || row_bytes == 64 || row_bytes == 128, "row_bytes must be 32/64/128");
  VSB_CHECK_ARG(shift_rows >= 0 && shift_rows <= 128, "bad shift");
  const size_t smem = 3 * 16384 + 256 * 128 + 2048;
  VSB_CHECK_CUDA(cudaFuncSetAttribute(umma_rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  umma_rate2_kernel<<<grid, 128, smem, static_cast<cudaStream_t>(stream)>>>(n, groups, per_group, commit_each, row_bytes,
                                                                            shift_rows, walk, clk);
  VSB_CHECK_LAUNCH("umma_rate2_kernel");
  return VSB_OK;
}
