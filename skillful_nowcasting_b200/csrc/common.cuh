// Shared helpers for libdgmr_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/dgmr_b200.h"

namespace dgmr {

void set_error(const char* fmt, ...);

#define DGMR_CHECK_LAUNCH(name)                                                        \
  do {                                                                                 \
    cudaError_t e__ = cudaGetLastError();                                              \
    if (e__ != cudaSuccess) {                                                          \
      dgmr::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));         \
      return 2;                                                                        \
    }                                                                                  \
  } while (0)

#define DGMR_CUDA(call)                                                                \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      dgmr::set_error("%s failed: %s", #call, cudaGetErrorString(e__));                \
      return 2;                                                                        \
    }                                                                                  \
  } while (0)

#define DGMR_REQUIRE(cond, ...)                                                        \
  do {                                                                                 \
    if (!(cond)) {                                                                     \
      dgmr::set_error(__VA_ARGS__);                                                    \
      return 1;                                                                        \
    }                                                                                  \
  } while (0)

static inline cudaStream_t S(dgmr_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// number of SMs (cached)
int sm_count();

// grid size for grid-stride element-wise kernels: enough CTAs for ~8 waves max, multiple of SM count
static inline int ew_grid(int64_t n, int threads, int per_thread = 4) {
  int64_t want = ceil_div(n, (int64_t)threads * per_thread);
  int64_t cap = (int64_t)sm_count() * 16;
  if (want < 1) want = 1;
  if (want > cap) want = cap;
  return (int)want;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace dgmr
