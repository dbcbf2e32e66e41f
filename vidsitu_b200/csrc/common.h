// Host-side helpers shared by the translation units of libvidsitu_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vidsitu_b200.h"

namespace vsb {

// thread-local last-error buffer (vsb_last_error)
void set_error(const char* fmt, ...);
// count one kernel launch (vsb_launch_count)
void count_launch(int n = 1);

#define VSB_CHECK_ARG(cond, ...)      \
  do {                                \
    if (!(cond)) {                    \
      ::vsb::set_error(__VA_ARGS__);  \
      return VSB_ERR_INVALID;         \
    }                                 \
  } while (0)

#define VSB_CHECK_CUDA(expr)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ::vsb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return VSB_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

// Launch-error check that does not synchronise (safe during graph capture).
#define VSB_CHECK_LAUNCH(name)                                                          \
  do {                                                                                  \
    cudaError_t _e = cudaPeekAtLastError();                                             \
    if (_e != cudaSuccess) {                                                            \
      ::vsb::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));        \
      (void)cudaGetLastError();                                                         \
      return VSB_ERR_CUDA;                                                              \
    }                                                                                   \
    ::vsb::count_launch();                                                              \
  } while (0)

// Programmatic dependent launch between consecutive kernels of a stream (VSB_PDL=0 turns it off): the kernels
// call pdl_wait() before they touch activations (ptx.cuh).
bool pdl_enabled();

// cudaLaunchKernelEx with the programmatic-stream-serialization attribute (and an optional cluster size).
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t stream,
                                     unsigned cluster, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(block, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  unsigned na = 0;
  if (cluster > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cluster;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

}  // namespace vsb
