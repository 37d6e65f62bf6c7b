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

// ---- programmatic dependent launch (PDL): a kernel launched through launch_pdl() may start while its predecessor in
// the stream is still draining; its prologue (barrier init, TMEM allocation, tensor-map prefetch) overlaps that tail.
// It must execute pdl_wait() before touching global memory (no-op when launched without the attribute);
// pdl_launch_dependents() lets the NEXT kernel's CTAs be scheduled as this grid's CTAs retire.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// ---- tile-level dataflow flags (sib_flow, see the header): counters in global memory, one per 128-row block.
// flow_wait: the calling thread spins (acquire loads at gpu scope) until the counter has reached `target`, then fences
// the async proxy so that TMA loads issued afterwards are ordered behind the acquire.  A producer that never arrives is
// a bug of the launch chain, not a runtime condition: after 4 s the kernel traps (a loud launch failure, no hung GPU).
// The spin loop lives out of line: inlined into the tcgen05 kernels it cost the short GEMMs 2-8 us per launch through
// register allocation and scheduling of the surrounding loops even when it never ran (same-box microbenchmark of library
// builds with and without the code, scripts/flow_microbench.py).
static __device__ __noinline__ void flow_spin(const int32_t* ctr, int32_t target) {
  int32_t v;
  uint64_t t_start, t_now;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
  do {
    __nanosleep(64);
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_now));
    if (t_now - t_start > 4000000000ull) __trap();
  } while (v < target);
}
template <bool ASYNC_PROXY_READER = true>
__device__ __forceinline__ void flow_wait(const int32_t* ctr, int32_t target) {
  int32_t v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
  if (v < target) flow_spin(ctr, target);
  // TMA loads issued after this point read through the async proxy: order them behind the (generic-proxy) acquire
  if (ASYNC_PROXY_READER) asm volatile("fence.proxy.async;" ::: "memory");
}
// flow_signal: the calling thread's earlier writes (and, through a preceding CTA / warp barrier, its peers') are
// published before the counter moves.  Writes made through the async proxy (TMA stores) must have COMPLETED first
// (cp.async.bulk.wait_group, not .read); their completion carries an implicit generic-async proxy fence.
__device__ __forceinline__ void flow_signal(int32_t* ctr, int32_t amount) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(ctr), "r"(amount) : "memory");
}
// the same for rows written by TMA stores of the calling thread: all of its bulk store groups except the `newer` most
// recent ones must have completed (out of line for the same reason as flow_spin)
static __device__ __noinline__ void flow_signal_stores(int32_t* ctr, int32_t amount, int newer) {
  if (newer <= 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  else if (newer == 1) asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");
  else if (newer == 2) asm volatile("cp.async.bulk.wait_group 2;" ::: "memory");
  else asm volatile("cp.async.bulk.wait_group 3;" ::: "memory");
  flow_signal(ctr, amount);
}

bool pdl_enabled();   // api.cu: false when SIB_NO_PDL is set or after sib_set_pdl(0)
void set_pdl(int on);

// cluster_x > 1 launches thread-block clusters of that many CTAs along x (CTA pairs for cta_group::2 kernels)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                      unsigned cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.numAttrs = 1;
  if (cluster_x > 1) {
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = cluster_x;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
  }
  cfg.attrs = attr;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  return launch_pdl_cluster(kernel, grid, block, smem, stream, 1u, static_cast<Args&&>(args)...);
}

}  // namespace sib
