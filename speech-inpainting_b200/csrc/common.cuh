// Shared helpers for the sm_100a kernels of the Speech-Inpainting hot path.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/speech_inpainting_b200.h"

namespace sib {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define SIB_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      sib::set_error(__VA_ARGS__);             \
      return SIB_ERR_INVALID;                  \
    }                                          \
  } while (0)

// Checks the launch (not the execution: calls are asynchronous).
#define SIB_CHECK_LAUNCH(name)                                              \
  do {                                                                      \
    cudaError_t e_ = cudaGetLastError();                                    \
    if (e_ != cudaSuccess) {                                                \
      sib::set_error("%s: CUDA error %s", name, cudaGetErrorString(e_));    \
      return SIB_ERR_CUDA;                                                  \
    }                                                                       \
    sib::count_launch();                                                    \
  } while (0)

__device__ __forceinline__ float gelu_erf(float x) {
  // exact GELU (ACT2FN["gelu"] == F.gelu)
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  switch (act) {
    case SIB_ACT_GELU: return gelu_erf(v);
    case SIB_ACT_LRELU: return v > 0.f ? v : v * slope;
    case SIB_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace sib
