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

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

}  // namespace vsb
