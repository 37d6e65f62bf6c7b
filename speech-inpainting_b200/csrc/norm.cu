// Bandwidth-bound normalisation kernels of the HuBERT front: processor z-norm (a2), conv0 fused with
// GroupNorm/GELU (a3), LayerNorm (a4, a8, a10).  Warp-shuffle reductions, vectorised row access.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------ LayerNorm
// One warp per row; the row lives in registers (C <= 32*MAXV) so x is read exactly once.
__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <int MAXV, typename TX, typename TR, typename TY>
__global__ void __launch_bounds__(256) layernorm_kernel(const TX* __restrict__ x, const TR* __restrict__ res,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        TY* __restrict__ y, int64_t rows, int C, float eps,
                                                        int post_act) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const TX* xr = x + row * C;
  const TR* rr = res ? res + row * C : nullptr;
  float v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    float t = 0.f;
    if (c < C) {
      t = ldf(xr + c);
      if (rr) t += ldf(rr + c);
    }
    v[i] = t;
    s += t;
  }
  const float mean = sib::warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    const float dlt = (c < C) ? v[i] - mean : 0.f;
    q += dlt * dlt;
  }
  const float rstd = rsqrtf(sib::warp_sum(q) / C + eps);
  TY* yr = y + row * C;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < C) {
      float o = (v[i] - mean) * rstd * gamma[c] + beta[c];
      if (post_act == SIB_ACT_GELU) o = sib::gelu_erf(o);
      stf(yr + c, o);
    }
  }
}

// bf16 rows with C % 8 == 0: one warp per row, 16-byte loads / stores (lane owns chunks lane, lane+32, ... of 8 channels),
// row in registers, two-pass variance in fp32.  3 vector loads per tensor per lane at C = 768 instead of 24 scalar ones.
template <int NV, bool COHERENT>
__device__ __forceinline__ void layernorm_bf16_row(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ res,
                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                   __nv_bfloat16* __restrict__ y, int64_t row, int C, float eps, int post_act,
                                                   int lane);
// FLOW (sib_layernorm_flow_bf16, see sib_flow in the header): instead of waiting for the previous grid, a row's warp polls
// the counter of its 128-row block, reads the row with L2-coherent loads (it may have been written while this grid was
// already running) and adds C / 32 to the outgoing counter once the row is stored.
template <int NV, bool FLOW = false>
__global__ void __launch_bounds__(256) layernorm_bf16_vec_kernel(const __nv_bfloat16* __restrict__ x,
                                                                 const __nv_bfloat16* __restrict__ res,
                                                                 const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta,
                                                                 __nv_bfloat16* __restrict__ y, int64_t rows, int C,
                                                                 float eps, int post_act, const sib_flow flow) {
  if (!(FLOW && flow.wait)) sib::pdl_wait();   // PDL-launched: may start while the producing GEMM is still draining
  sib::pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  if (FLOW) {
    // warp-granular: a row's warp polls and publishes on its own (no CTA barrier in this single-wave, latency-bound kernel)
    const int rb = (int)(row >> 7);
    if (flow.wait) {
      if (lane == 0) sib::flow_wait<false>(flow.wait + rb, rb == (int)((rows - 1) >> 7) ? flow.wait_target_last : flow.wait_target);
      __syncwarp();
    }
    layernorm_bf16_row<NV, true>(x, res, gamma, beta, y, row, C, eps, post_act, lane);
    if (flow.signal) {
      __syncwarp();            // the lanes' stores are ordered before lane 0's release
      if (lane == 0) sib::flow_signal(flow.signal + rb, C >> 5);
    }
    return;
  }
  layernorm_bf16_row<NV, false>(x, res, gamma, beta, y, row, C, eps, post_act, lane);
}

template <int NV, bool COHERENT>
__device__ __forceinline__ void layernorm_bf16_row(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ res,
                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                   __nv_bfloat16* __restrict__ y, int64_t row, int C, float eps, int post_act,
                                                   int lane) {
  const int nchunks = C >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * C);
  const uint4* rr = res ? reinterpret_cast<const uint4*>(res + row * C) : nullptr;
  float v[NV][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int ch = lane + i * 32;
    if (ch < nchunks) {
      const uint4 a = COHERENT ? __ldcg(xr + ch) : __ldg(xr + ch);
      const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 t = __bfloat1622float2(a2[u]);
        v[i][2 * u] = t.x; v[i][2 * u + 1] = t.y;
      }
      if (rr) {
        const uint4 b = COHERENT ? __ldcg(rr + ch) : __ldg(rr + ch);
        const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float2 t = __bfloat1622float2(b2[u]);
          v[i][2 * u] += t.x; v[i][2 * u + 1] += t.y;
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[i][u];
    } else {
#pragma unroll
      for (int u = 0; u < 8; ++u) v[i][u] = 0.f;
    }
  }
  const float mean = sib::warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (lane + i * 32 < nchunks) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float dlt = v[i][u] - mean;
        q += dlt * dlt;
      }
    }
  }
  const float rstd = rsqrtf(sib::warp_sum(q) / C + eps);
  uint4* yr = reinterpret_cast<uint4*>(y + row * C);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int ch = lane + i * 32;
    if (ch < nchunks) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * ch), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * ch + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * ch), b1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * ch + 1);
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      uint4 o;
      __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float e0 = (v[i][2 * u] - mean) * rstd * g[2 * u] + bb[2 * u];
        float e1 = (v[i][2 * u + 1] - mean) * rstd * g[2 * u + 1] + bb[2 * u + 1];
        if (post_act == SIB_ACT_GELU) { e0 = sib::gelu_erf(e0); e1 = sib::gelu_erf(e1); }
        o2[u] = __floats2bfloat162_rn(e0, e1);
      }
      yr[ch] = o;
    }
  }
}

// ------------------------------------------------------------------------------------------ conv0
// HuBERT conv0: Conv1d(1 -> C, k, stride) over the raw waveform (HF:154-175).  10 MACs per output, so
// it is recomputed in both GroupNorm passes instead of materialising [B,C,N/5] twice.
// CTA = (64-frame tile, batch); 256 threads; thread owns channels tid, tid+256, ...
constexpr int C0_TILE = 64;
constexpr int C0_MAXK = 16;

template <int MODE, typename TY>
__global__ void __launch_bounds__(256) conv0_kernel(const float* __restrict__ wave, int n, int64_t wave_bs,
                                                    const float* __restrict__ w, const float* __restrict__ bias,
                                                    int C, int K, int S, int T0, float* __restrict__ partial,
                                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                                    TY* __restrict__ y) {
  extern __shared__ float xs[];  // (C0_TILE-1)*S + K samples
  const int b = blockIdx.y, tile = blockIdx.x;
  const int t_begin = tile * C0_TILE;
  const int nt = min(C0_TILE, T0 - t_begin);
  const int span = (nt - 1) * S + K;
  const float* wb = wave + (int64_t)b * wave_bs + (int64_t)t_begin * S;
  for (int i = threadIdx.x; i < span; i += blockDim.x) xs[i] = wb[i];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float wr[C0_MAXK];
#pragma unroll
    for (int j = 0; j < C0_MAXK; ++j) wr[j] = j < K ? w[c * K + j] : 0.f;
    const float bs = bias ? bias[c] : 0.f;
    float s = 0.f, q = 0.f;
    float m = 0.f, r = 0.f, ga = 0.f, be = 0.f;
    if (MODE == 1) {
      m = mean[b * C + c]; r = rstd[b * C + c]; ga = gamma[c]; be = beta[c];
    }
    for (int t = 0; t < nt; ++t) {
      float acc = bs;
#pragma unroll
      for (int j = 0; j < C0_MAXK; ++j)
        if (j < K) acc = fmaf(wr[j], xs[t * S + j], acc);
      if (MODE == 0) {
        s += acc; q += acc * acc;
      } else if (MODE == 1) {
        stf(y + ((int64_t)b * T0 + t_begin + t) * C + c, sib::gelu_erf((acc - m) * r * ga + be));
      } else {
        stf(y + ((int64_t)b * T0 + t_begin + t) * C + c, acc);
      }
    }
    if (MODE == 0) {
      float* pp = partial + (((int64_t)b * gridDim.x + tile) * C + c) * 2;
      pp[0] = s; pp[1] = q;
    }
  }
}

__global__ void gn_finalize_kernel(const float* __restrict__ partial, int n_tiles, int C, int T0, float eps,
                                   float* __restrict__ mean, float* __restrict__ rstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int t = 0; t < n_tiles; ++t) {
    const float* pp = partial + (((int64_t)b * n_tiles + t) * C + c) * 2;
    s += pp[0]; q += pp[1];
  }
  const double m = s / T0;
  double var = q / T0 - m * m;
  if (var < 0.0) var = 0.0;
  mean[b * C + c] = (float)m;
  rstd[b * C + c] = (float)(1.0 / sqrt(var + (double)eps));
}

// ------------------------------------------------------------------------------------------ conv0: closed-form GN stats
// GroupNorm(C, C) statistics of conv0's output WITHOUT evaluating conv0: y[c,t] = sum_j w[c,j] x[S t + j] is linear
// in the waveform, so with  s1[j] = sum_t x[S t + j]  and  R[j][j'] = sum_t x[S t + j] x[S t + j']  (K + K(K+1)/2
// numbers per utterance, one pass over the samples):
//     sum_t y[c,t]   = w_c . s1 + T0 b_c         sum_t (y[c,t] - b_c)^2 = w_c^T R w_c
// One CTA per utterance; per-thread fp32 partial sums over ~T0/256 frames, then a fixed-order (deterministic)
// reduction and the per-channel quadratic forms in double.  Replaces a full [B, T0, C] evaluation pass.
template <int K, int S>
__global__ void __launch_bounds__(256) conv0_gn_stats_kernel(const float* __restrict__ wave, int64_t wave_bs,
                                                             const float* __restrict__ w, const float* __restrict__ bias,
                                                             int C, int T0, float eps, float* __restrict__ mean,
                                                             float* __restrict__ rstd) {
  constexpr int NR = K * (K + 1) / 2;
  constexpr int NV = K + NR;
  __shared__ float part[16][NV];
  __shared__ double tot[NV];
  const float* xb = wave + (int64_t)blockIdx.x * wave_bs;
  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.f;
  for (int t = threadIdx.x; t < T0; t += blockDim.x) {
    float x[K];
#pragma unroll
    for (int j = 0; j < K; ++j) x[j] = xb[(int64_t)t * S + j];
    int idx = K;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      acc[j] += x[j];
#pragma unroll
      for (int j2 = j; j2 < K; ++j2) acc[idx++] = fmaf(x[j], x[j2], acc[idx]);
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float v = sib::warp_sum(acc[i]);
    if (lane == 0) part[wid][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double v = 0.0;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) v += (double)part[wv][threadIdx.x];
    tot[threadIdx.x] = v;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double wr[K];
#pragma unroll
    for (int j = 0; j < K; ++j) wr[j] = (double)w[c * K + j];
    double lin = 0.0, quad = 0.0;
    int idx = K;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      lin += wr[j] * tot[j];
#pragma unroll
      for (int j2 = j; j2 < K; ++j2) quad += (j2 == j ? 1.0 : 2.0) * wr[j] * wr[j2] * tot[idx++];
    }
    const double m0 = lin / T0;                  // mean of the bias-free response
    double var = quad / T0 - m0 * m0;
    if (var < 0.0) var = 0.0;
    mean[(int64_t)blockIdx.x * C + c] = (float)(m0 + (bias ? (double)bias[c] : 0.0));
    rstd[(int64_t)blockIdx.x * C + c] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

// conv0 + GroupNorm + GELU apply pass for K = 10, S = 5 (HuBERT): the normalisation is folded into the taps
// (w' = w rstd gamma, b' = beta + (bias - mean) rstd gamma), a thread owns two adjacent channels and walks the tile four
// frames at a time so that 25 broadcast shared-memory loads feed 80 FMAs; outputs leave as packed pairs (coalesced
// 4 / 8 bytes x 256 threads per frame).  bf16 output uses the 2-MUFU erf (|err| < 1.5e-7), fp32 output erff.
__device__ __forceinline__ float gelu_erf_as(float x) {
  // z = |x| / sqrt 2 never materialises: 0.3275911 z = 0.23164190 |x| and z^2 log2(e) = 0.72134752 x^2
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.2316419f, fabsf(x), 1.0f)));   // argument in [1, inf)
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * -0.72134752044448170f));
  const float erfc_abs = poly * t * e;
  return fmaf(-0.5f, fabsf(x) * erfc_abs, fmaxf(x, 0.f));   // x Phi(x) = relu(x) - 0.5 |x| erfc(|x| / sqrt 2): no select
}
__device__ __forceinline__ void st2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void st2(__nv_bfloat16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}
__device__ __forceinline__ float gelu_for(float v, float*) { return sib::gelu_erf(v); }
__device__ __forceinline__ float gelu_for(float v, __nv_bfloat16*) { return gelu_erf_as(v); }

template <typename TY>
__global__ void __launch_bounds__(256) conv0_gn_apply_k10s5_kernel(const float* __restrict__ wave, int64_t wave_bs,
                                                                    const float* __restrict__ w,
                                                                    const float* __restrict__ bias, int C, int T0,
                                                                    const float* __restrict__ mean,
                                                                    const float* __restrict__ rstd,
                                                                    const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta, TY* __restrict__ y) {
  constexpr int K = 10, S = 5, TILE = 64;
  __shared__ float xs[(TILE - 1) * S + K + 3];
  const int b = blockIdx.y, t_begin = blockIdx.x * TILE;
  const int nt = min(TILE, T0 - t_begin);
  const int span = (nt - 1) * S + K;
  const float* wb = wave + (int64_t)b * wave_bs + (int64_t)t_begin * S;
  for (int i = threadIdx.x; i < (TILE - 1) * S + K + 3; i += blockDim.x) xs[i] = i < span ? wb[i] : 0.f;
  __syncthreads();
  for (int c = 2 * threadIdx.x; c < C; c += 2 * blockDim.x) {
    float w0[K], w1[K];
    const float g0 = rstd[b * C + c] * gamma[c], g1 = rstd[b * C + c + 1] * gamma[c + 1];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      w0[j] = w[c * K + j] * g0;
      w1[j] = w[(c + 1) * K + j] * g1;
    }
    const float b0 = fmaf((bias ? bias[c] : 0.f) - mean[b * C + c], g0, beta[c]);
    const float b1 = fmaf((bias ? bias[c + 1] : 0.f) - mean[b * C + c + 1], g1, beta[c + 1]);
    TY* yb = y + ((int64_t)b * T0 + t_begin) * C + c;
    for (int t = 0; t < nt; t += 4) {
      float x[3 * S + K];
#pragma unroll
      for (int i = 0; i < 3 * S + K; ++i) x[i] = xs[t * S + i];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float a0 = b0, a1 = b1;
#pragma unroll
        for (int j = 0; j < K; ++j) {
          a0 = fmaf(w0[j], x[u * S + j], a0);
          a1 = fmaf(w1[j], x[u * S + j], a1);
        }
        if (t + u < nt) st2(yb + (int64_t)(t + u) * C, gelu_for(a0, (TY*)nullptr), gelu_for(a1, (TY*)nullptr));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ z-norm
// One CTA per utterance: mean, then centred sum of squares (two passes; the second hits L2), then apply.
__global__ void __launch_bounds__(1024) znorm_kernel(const float* __restrict__ x, float* __restrict__ y, int n,
                                                     const int32_t* __restrict__ lengths, float eps) {
  __shared__ float red[32];
  __shared__ float bc;
  const int b = blockIdx.x;
  const float* xb = x + (int64_t)b * n;
  float* yb = y + (int64_t)b * n;
  const int len = lengths ? min(max(lengths[b], 0), n) : n;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  auto block_sum = [&](float v) {
    v = sib::warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    if (wid == 0) {
      float t = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
      t = sib::warp_sum(t);
      if (lane == 0) bc = t;
    }
    __syncthreads();
    return bc;
  };
  float s = 0.f;
  for (int i = threadIdx.x; i < len; i += blockDim.x) s += xb[i];
  const float mean = block_sum(s) / (float)max(len, 1);
  float q = 0.f;
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    const float dlt = xb[i] - mean;
    q += dlt * dlt;
  }
  const float rstd = rsqrtf(block_sum(q) / (float)max(len, 1) + eps);
  for (int i = threadIdx.x; i < n; i += blockDim.x) yb[i] = i < len ? (xb[i] - mean) * rstd : 0.f;
}

// Cluster variant: an utterance is shared by a thread-block cluster of ZN_CL CTAs (one CTA per utterance leaves a 32-utterance
// batch on 32 of 148 SMs).  Every thread keeps its PER samples in registers, so the waveform is read from HBM exactly once;
// the two reductions (sum, centred sum of squares) meet through distributed shared memory in a fixed order.
constexpr int ZN_CL = 8;
template <int PER>
__global__ void __cluster_dims__(ZN_CL, 1, 1) __launch_bounds__(1024)
znorm_cluster_kernel(const float* __restrict__ x, float* __restrict__ y, int n, const int32_t* __restrict__ lengths, float eps) {
  __shared__ float red[32];
  __shared__ float part[2];
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int b = blockIdx.x / ZN_CL;
  const float* xb = x + (int64_t)b * n;
  float* yb = y + (int64_t)b * n;
  const int len = lengths ? min(max(lengths[b], 0), n) : n;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  auto cluster_sync = [] {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  };
  // sum over the cluster: block reduction -> part[slot] of every CTA -> all CTAs add the ZN_CL partials in rank order
  auto cluster_sum = [&](float v, int slot) {
    v = sib::warp_sum(v);
    if (lane == 0) red[wid] = v;
    __syncthreads();
    if (wid == 0) {
      float t = red[lane];               // 1024 threads = 32 warps
      t = sib::warp_sum(t);
      if (lane == 0) part[slot] = t;
    }
    cluster_sync();
    float tot = 0.f;
    const uint32_t local = (uint32_t)__cvta_generic_to_shared(&part[slot]);
#pragma unroll
    for (uint32_t r = 0; r < ZN_CL; ++r) {
      uint32_t remote;
      float pv;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(pv) : "r"(remote) : "memory");
      tot += pv;
    }
    return tot;
  };
  float v[PER];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int i = (j * ZN_CL + (int)rank) * 1024 + (int)threadIdx.x;   // interleaved slices: coalesced per CTA
    v[j] = i < len ? xb[i] : 0.f;
    s += v[j];
  }
  const float mean = cluster_sum(s, 0) / (float)max(len, 1);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int i = (j * ZN_CL + (int)rank) * 1024 + (int)threadIdx.x;
    const float dlt = v[j] - mean;
    q += i < len ? dlt * dlt : 0.f;
  }
  const float rstd = rsqrtf(cluster_sum(q, 1) / (float)max(len, 1) + eps);
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int i = (j * ZN_CL + (int)rank) * 1024 + (int)threadIdx.x;
    if (i < n) yb[i] = i < len ? (v[j] - mean) * rstd : 0.f;
  }
  cluster_sync();   // no CTA leaves while a peer may still read its partials
}

__global__ void zero_ranges_kernel(float* __restrict__ wave, int n, const int32_t* __restrict__ lo,
                                   const int32_t* __restrict__ hi, float add_eps) {
  const int b = blockIdx.y;
  // numpy slice semantics: negative indices count from the end, then clamp to [0, n]
  int l = lo[b], h = hi[b];
  if (l < 0) l += n;
  if (h < 0) h += n;
  l = min(max(l, 0), n);
  h = min(max(h, 0), n);
  float* wb = wave + (int64_t)b * n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (i >= l && i < h)
      wb[i] = 0.f;
    else if (add_eps != 0.f)
      wb[i] = wb[i] + add_eps;
  }
}

// Feature front-end of I_ea (predict.py:99-103): zero [lo, hi) of the 22.05 kHz wave, then librosa.util.normalize
// (x / max|x|, rows whose peak is below the smallest normal float are left as they are) times `scale` (0.95).
// One CTA per utterance; the masked samples never count towards the peak and come out as exact zeros.
__global__ void __launch_bounds__(1024) mask_peak_normalize_kernel(const float* __restrict__ x, float* __restrict__ y, int n,
                                                                   const int32_t* __restrict__ lo,
                                                                   const int32_t* __restrict__ hi, float scale) {
  __shared__ float red[32];
  __shared__ float bc;
  const int b = blockIdx.x;
  const float* xb = x + (int64_t)b * n;
  float* yb = y + (int64_t)b * n;
  int l = 0, h = 0;
  if (lo && hi) {   // numpy slice semantics, as sib_zero_ranges_f32
    l = lo[b]; h = hi[b];
    if (l < 0) l += n;
    if (h < 0) h += n;
    l = min(max(l, 0), n);
    h = min(max(h, 0), n);
  }
  float m = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    if (i < l || i >= h) m = fmaxf(m, fabsf(xb[i]));
  m = sib::warp_max(m);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) red[wid] = m;
  __syncthreads();
  if (wid == 0) {
    float t = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.f;
    t = sib::warp_max(t);
    if (lane == 0) bc = t;
  }
  __syncthreads();
  const float peak = bc >= 1.17549435e-38f ? bc : 1.0f;   // np.finfo(np.float32).tiny
  for (int i = threadIdx.x; i < n; i += blockDim.x) yb[i] = (i >= l && i < h) ? 0.f : (xb[i] / peak) * scale;
}

template <typename TY>
__global__ void zero_padded_frames_kernel(TY* __restrict__ h, const int32_t* __restrict__ key_len, int T, int C) {
  const int b = blockIdx.y;
  const int kl = key_len[b];
  const int64_t total = (int64_t)(T - kl) * C;
  if (total <= 0) return;
  TY* hb = h + ((int64_t)b * T + kl) * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    stf(hb + i, 0.f);
}

}  // namespace

namespace {
template <typename TX, typename TR, typename TY>
int launch_ln(const void* x, const void* residual, const float* gamma, const float* beta, void* y, int64_t rows, int c,
              float eps, int post_act, cudaStream_t s) {
  const int wpb = 8;
  const unsigned grid = (unsigned)((rows + wpb - 1) / wpb);
  const TX* xx = static_cast<const TX*>(x);
  const TR* rr = static_cast<const TR*>(residual);
  TY* yy = static_cast<TY*>(y);
  if (c <= 256)
    layernorm_kernel<8><<<grid, wpb * 32, 0, s>>>(xx, rr, gamma, beta, yy, rows, c, eps, post_act);
  else if (c <= 512)
    layernorm_kernel<16><<<grid, wpb * 32, 0, s>>>(xx, rr, gamma, beta, yy, rows, c, eps, post_act);
  else if (c <= 1024)
    layernorm_kernel<32><<<grid, wpb * 32, 0, s>>>(xx, rr, gamma, beta, yy, rows, c, eps, post_act);
  else
    layernorm_kernel<128><<<grid, wpb * 32, 0, s>>>(xx, rr, gamma, beta, yy, rows, c, eps, post_act);
  return 0;
}
}  // namespace

extern "C" int sib_layernorm(const void* x, int x_dtype, const void* residual, int r_dtype, const float* gamma,
                             const float* beta, void* y, int y_dtype, int64_t rows, int c, float eps, int post_act,
                             sib_stream_t stream) {
  SIB_REQUIRE(x && gamma && beta && y && rows > 0 && c > 0, "sib_layernorm: bad argument");
  SIB_REQUIRE(c <= 4096, "sib_layernorm: c=%d > 4096 unsupported", c);
  SIB_REQUIRE(post_act == SIB_ACT_NONE || post_act == SIB_ACT_GELU, "sib_layernorm: unsupported post_act");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int key = (x_dtype << 2) | ((residual ? r_dtype : x_dtype) << 1) | y_dtype;
  auto al16 = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
  if (key == 7 && c % 8 == 0 && c <= 2048 && al16(x) && al16(y) && al16(gamma) && al16(beta) && (!residual || al16(residual))) {
    const int warps = 8;
    const unsigned grid = (unsigned)((rows + warps - 1) / warps);
    const __nv_bfloat16* xb = (const __nv_bfloat16*)x;
    const __nv_bfloat16* rb = (const __nv_bfloat16*)residual;
    __nv_bfloat16* yb = (__nv_bfloat16*)y;
    const int nv = (c / 8 + 31) / 32;
    cudaError_t le;
    const sib_flow nf = {nullptr, nullptr, 0, 0};
    if (nv <= 2) le = sib::launch_pdl(layernorm_bf16_vec_kernel<2>, dim3(grid), dim3(warps * 32), 0, s, xb, rb, gamma, beta, yb, rows, c, eps, post_act, nf);
    else if (nv <= 4) le = sib::launch_pdl(layernorm_bf16_vec_kernel<4>, dim3(grid), dim3(warps * 32), 0, s, xb, rb, gamma, beta, yb, rows, c, eps, post_act, nf);
    else le = sib::launch_pdl(layernorm_bf16_vec_kernel<8>, dim3(grid), dim3(warps * 32), 0, s, xb, rb, gamma, beta, yb, rows, c, eps, post_act, nf);
    if (le != cudaSuccess) {
      sib::set_error("sib_layernorm: launch failed: %s", cudaGetErrorString(le));
      return SIB_ERR_CUDA;
    }
    SIB_CHECK_LAUNCH("sib_layernorm");
    return SIB_OK;
  }
  switch (key) {
    case 0: launch_ln<float, float, float>(x, residual, gamma, beta, y, rows, c, eps, post_act, s); break;
    case 1: launch_ln<float, float, __nv_bfloat16>(x, residual, gamma, beta, y, rows, c, eps, post_act, s); break;
    case 6: launch_ln<__nv_bfloat16, __nv_bfloat16, float>(x, residual, gamma, beta, y, rows, c, eps, post_act, s); break;
    case 7: launch_ln<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16>(x, residual, gamma, beta, y, rows, c, eps, post_act, s); break;
    default: SIB_REQUIRE(false, "sib_layernorm: unsupported dtype combination x=%d r=%d y=%d", x_dtype, r_dtype, y_dtype);
  }
  SIB_CHECK_LAUNCH("sib_layernorm");
  return SIB_OK;
}

extern "C" int sib_layernorm_flow_bf16(const void* x, const void* residual, const float* gamma, const float* beta, void* y,
                                       int64_t rows, int c, float eps, const sib_flow* flow, sib_stream_t stream) {
  SIB_REQUIRE(x && gamma && beta && y && rows > 0 && flow, "sib_layernorm_flow_bf16: bad argument");
  auto al16 = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
  SIB_REQUIRE(c % 32 == 0 && c <= 2048 && al16(x) && al16(y) && al16(gamma) && al16(beta) && (!residual || al16(residual)),
              "sib_layernorm_flow_bf16: c=%d must be a multiple of 32, <= 2048, with 16-byte aligned operands", c);
  SIB_REQUIRE(!flow->wait || (flow->wait_target > 0 && flow->wait_target_last > 0), "sib_layernorm_flow_bf16: wait targets");
  const int warps = 8;    // 128 % warps == 0: a CTA's rows never straddle a 128-row block
  const unsigned grid = (unsigned)((rows + warps - 1) / warps);
  const __nv_bfloat16* xb = (const __nv_bfloat16*)x;
  const __nv_bfloat16* rb = (const __nv_bfloat16*)residual;
  __nv_bfloat16* yb = (__nv_bfloat16*)y;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int nv = (c / 8 + 31) / 32;
  const int act = SIB_ACT_NONE;
  cudaError_t le;
  if (nv <= 2) le = sib::launch_pdl(layernorm_bf16_vec_kernel<2, true>, dim3(grid), dim3(warps * 32), 0, s, xb, rb, gamma, beta, yb, rows, c, eps, act, *flow);
  else if (nv <= 4) le = sib::launch_pdl(layernorm_bf16_vec_kernel<4, true>, dim3(grid), dim3(warps * 32), 0, s, xb, rb, gamma, beta, yb, rows, c, eps, act, *flow);
  else le = sib::launch_pdl(layernorm_bf16_vec_kernel<8, true>, dim3(grid), dim3(warps * 32), 0, s, xb, rb, gamma, beta, yb, rows, c, eps, act, *flow);
  if (le != cudaSuccess) {
    sib::set_error("sib_layernorm_flow_bf16: launch failed: %s", cudaGetErrorString(le));
    return SIB_ERR_CUDA;
  }
  SIB_CHECK_LAUNCH("sib_layernorm_flow_bf16");
  return SIB_OK;
}

extern "C" int sib_layernorm_f32(const float* x, const float* residual, const float* gamma, const float* beta,
                                 float* y, int64_t rows, int c, float eps, int post_act, sib_stream_t stream) {
  return sib_layernorm(x, SIB_F32, residual, SIB_F32, gamma, beta, y, SIB_F32, rows, c, eps, post_act, stream);
}

extern "C" int sib_conv0_num_tiles(int t0) { return (t0 + C0_TILE - 1) / C0_TILE; }

extern "C" int sib_conv0(int mode, const float* wave, int batch, int n_samples, int64_t wave_batch_stride,
                         const float* w, const float* bias, int c, int k, int stride, int t0, float* partial,
                         const float* mean, const float* rstd, const float* gamma, const float* beta, void* y,
                         int y_dtype, sib_stream_t stream) {
  SIB_REQUIRE(wave && w && batch > 0 && c > 0 && t0 > 0, "sib_conv0: bad argument");
  SIB_REQUIRE(k > 0 && k <= C0_MAXK && stride > 0, "sib_conv0_f32: k=%d stride=%d unsupported", k, stride);
  SIB_REQUIRE((int64_t)(t0 - 1) * stride + k <= n_samples, "sib_conv0_f32: t0=%d does not fit n_samples=%d", t0, n_samples);
  SIB_REQUIRE(batch <= 65535, "sib_conv0_f32: batch too large");
  dim3 grid(sib_conv0_num_tiles(t0), batch);
  const size_t smem = ((size_t)(C0_TILE - 1) * stride + k) * sizeof(float);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (mode == 0) {
    SIB_REQUIRE(partial, "sib_conv0_f32: mode 0 needs partial");
    conv0_kernel<0, float><<<grid, 256, smem, s>>>(wave, n_samples, wave_batch_stride, w, bias, c, k, stride, t0, partial,
                                                   nullptr, nullptr, nullptr, nullptr, nullptr);
  } else if (mode == 1) {
    SIB_REQUIRE(mean && rstd && gamma && beta && y, "sib_conv0_f32: mode 1 needs mean/rstd/gamma/beta/y");
    if (k == 10 && stride == 5 && c % 2 == 0) {
      if (y_dtype == SIB_BF16)
        conv0_gn_apply_k10s5_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(wave, wave_batch_stride, w, bias, c, t0, mean, rstd,
                                                                        gamma, beta, (__nv_bfloat16*)y);
      else
        conv0_gn_apply_k10s5_kernel<float><<<grid, 256, 0, s>>>(wave, wave_batch_stride, w, bias, c, t0, mean, rstd, gamma,
                                                                beta, (float*)y);
    } else if (y_dtype == SIB_BF16)
      conv0_kernel<1, __nv_bfloat16><<<grid, 256, smem, s>>>(wave, n_samples, wave_batch_stride, w, bias, c, k, stride, t0,
                                                             nullptr, mean, rstd, gamma, beta, (__nv_bfloat16*)y);
    else
      conv0_kernel<1, float><<<grid, 256, smem, s>>>(wave, n_samples, wave_batch_stride, w, bias, c, k, stride, t0,
                                                     nullptr, mean, rstd, gamma, beta, (float*)y);
  } else if (mode == 2) {
    SIB_REQUIRE(y, "sib_conv0_f32: mode 2 needs y");
    if (y_dtype == SIB_BF16)
      conv0_kernel<2, __nv_bfloat16><<<grid, 256, smem, s>>>(wave, n_samples, wave_batch_stride, w, bias, c, k, stride, t0,
                                                             nullptr, nullptr, nullptr, nullptr, nullptr, (__nv_bfloat16*)y);
    else
      conv0_kernel<2, float><<<grid, 256, smem, s>>>(wave, n_samples, wave_batch_stride, w, bias, c, k, stride, t0,
                                                     nullptr, nullptr, nullptr, nullptr, nullptr, (float*)y);
  } else {
    SIB_REQUIRE(false, "sib_conv0_f32: unknown mode %d", mode);
  }
  SIB_CHECK_LAUNCH("sib_conv0");
  return SIB_OK;
}

extern "C" int sib_conv0_gn_stats_f32(const float* wave, int batch, int n_samples, int64_t wave_batch_stride,
                                      const float* w, const float* bias, int c, int k, int stride, int t0, float eps,
                                      float* mean, float* rstd, sib_stream_t stream) {
  SIB_REQUIRE(wave && w && mean && rstd && batch > 0 && c > 0 && t0 > 0, "sib_conv0_gn_stats_f32: bad argument");
  SIB_REQUIRE(k == 10 && stride == 5, "sib_conv0_gn_stats_f32: k=%d stride=%d unsupported (HuBERT conv0 is k=10, stride=5); "
                                      "use sib_conv0 mode 0 + sib_gn_finalize_f32", k, stride);
  SIB_REQUIRE((int64_t)(t0 - 1) * stride + k <= n_samples, "sib_conv0_gn_stats_f32: t0=%d does not fit n_samples=%d", t0,
              n_samples);
  conv0_gn_stats_kernel<10, 5><<<batch, 256, 0, static_cast<cudaStream_t>(stream)>>>(wave, wave_batch_stride, w, bias, c, t0,
                                                                                    eps, mean, rstd);
  SIB_CHECK_LAUNCH("sib_conv0_gn_stats_f32");
  return SIB_OK;
}

extern "C" int sib_conv0_f32(int mode, const float* wave, int batch, int n_samples, int64_t wave_batch_stride,
                             const float* w, const float* bias, int c, int k, int stride, int t0, float* partial,
                             const float* mean, const float* rstd, const float* gamma, const float* beta, float* y,
                             sib_stream_t stream) {
  return sib_conv0(mode, wave, batch, n_samples, wave_batch_stride, w, bias, c, k, stride, t0, partial, mean, rstd, gamma,
                   beta, y, SIB_F32, stream);
}

extern "C" int sib_gn_finalize_f32(const float* partial, int batch, int n_tiles, int c, int t0, float eps,
                                   float* mean, float* rstd, sib_stream_t stream) {
  SIB_REQUIRE(partial && mean && rstd && batch > 0 && n_tiles > 0 && c > 0 && t0 > 0, "sib_gn_finalize_f32: bad argument");
  dim3 grid(sib::ceil_div(c, 128), batch);
  gn_finalize_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(partial, n_tiles, c, t0, eps, mean, rstd);
  SIB_CHECK_LAUNCH("sib_gn_finalize_f32");
  return SIB_OK;
}

extern "C" int sib_znorm_f32(const float* x, float* y, int batch, int n, const int32_t* lengths, float eps,
                             sib_stream_t stream) {
  SIB_REQUIRE(x && y && batch > 0 && n > 0, "sib_znorm_f32: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t per_cluster_pass = (int64_t)ZN_CL * 1024;
  if ((int64_t)batch * ZN_CL <= 0x7fffffff && n <= 32 * per_cluster_pass) {
    // registers hold the whole slice: 8 / 16 / 32 samples per thread cover 4 / 8 / 16 s at 16 kHz
    if (n <= 8 * per_cluster_pass) znorm_cluster_kernel<8><<<batch * ZN_CL, 1024, 0, s>>>(x, y, n, lengths, eps);
    else if (n <= 16 * per_cluster_pass) znorm_cluster_kernel<16><<<batch * ZN_CL, 1024, 0, s>>>(x, y, n, lengths, eps);
    else znorm_cluster_kernel<32><<<batch * ZN_CL, 1024, 0, s>>>(x, y, n, lengths, eps);
  } else {
    znorm_kernel<<<batch, 1024, 0, s>>>(x, y, n, lengths, eps);
  }
  SIB_CHECK_LAUNCH("sib_znorm_f32");
  return SIB_OK;
}

extern "C" int sib_zero_ranges_f32(float* wave, int batch, int n, const int32_t* lo, const int32_t* hi,
                                   float add_eps, sib_stream_t stream) {
  SIB_REQUIRE(wave && lo && hi && batch > 0 && n > 0 && batch <= 65535, "sib_zero_ranges_f32: bad argument");
  dim3 grid(min(sib::ceil_div(n, 256), 64), batch);
  zero_ranges_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(wave, n, lo, hi, add_eps);
  SIB_CHECK_LAUNCH("sib_zero_ranges_f32");
  return SIB_OK;
}

extern "C" int sib_mask_peak_normalize_f32(const float* x, float* y, int batch, int n, const int32_t* lo, const int32_t* hi,
                                           float scale, sib_stream_t stream) {
  SIB_REQUIRE(x && y && batch > 0 && n > 0, "sib_mask_peak_normalize_f32: bad argument");
  SIB_REQUIRE((lo == nullptr) == (hi == nullptr), "sib_mask_peak_normalize_f32: lo and hi must both be given or both be null");
  mask_peak_normalize_kernel<<<batch, 1024, 0, static_cast<cudaStream_t>(stream)>>>(x, y, n, lo, hi, scale);
  SIB_CHECK_LAUNCH("sib_mask_peak_normalize_f32");
  return SIB_OK;
}

extern "C" int sib_zero_padded_frames(void* h, int dtype, const int32_t* key_len, int batch, int t, int c,
                                      sib_stream_t stream) {
  SIB_REQUIRE(h && key_len && batch > 0 && t > 0 && c > 0 && batch <= 65535, "sib_zero_padded_frames: bad argument");
  dim3 grid(32, batch);
  if (dtype == SIB_BF16)
    zero_padded_frames_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>((__nv_bfloat16*)h, key_len, t, c);
  else
    zero_padded_frames_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>((float*)h, key_len, t, c);
  SIB_CHECK_LAUNCH("sib_zero_padded_frames");
  return SIB_OK;
}

extern "C" int sib_zero_padded_frames_f32(float* h, const int32_t* key_len, int batch, int t, int c,
                                          sib_stream_t stream) {
  return sib_zero_padded_frames(h, SIB_F32, key_len, batch, t, c, stream);
}
