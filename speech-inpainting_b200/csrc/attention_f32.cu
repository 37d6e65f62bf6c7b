// fp32 multi-head self-attention with online softmax (HF:234-259 eager_attention_forward,
// HF:296-345).  T <= ~500 frames, head_dim 64, additive key-padding mask expressed as key_len[b].
// CTA = 64 queries of one (batch, head); K/V streamed in 64-key tiles through shared memory.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int D = 64;    // head dim
constexpr int TQ = 64;   // queries per CTA
constexpr int TK = 64;   // keys per tile
constexpr int LD = 68;   // padded leading dim (floats), keeps float4 alignment

struct Smem {
  float Qt[D][LD];   // Q^T  [d][i], pre-scaled by d^-1/2
  float Kt[D][LD];   // K^T  [d][j]
  float Vs[TK][LD];  // V    [j][d]
  float Ss[TQ][LD];  // scores / probabilities [i][j]
  float alpha[TQ];   // per-row rescale of the running output
  float linv[TQ];
};

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
  const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
  uint2 r;
  *reinterpret_cast<__nv_bfloat162*>(&r.x) = __floats2bfloat162_rn(v.x, v.y);
  *reinterpret_cast<__nv_bfloat162*>(&r.y) = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = r;
}

template <typename TE>
__global__ void __launch_bounds__(256) attention_kernel(const TE* __restrict__ qkv,
                                                        const int32_t* __restrict__ key_len,
                                                        TE* __restrict__ out, int T, int heads) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int H = heads * D;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int q0 = blockIdx.x * TQ;
  const int h = blockIdx.y, b = blockIdx.z;
  const int kl = key_len ? min(key_len[b], T) : T;
  const float scale = rsqrtf((float)D);
  const TE* base = qkv + (int64_t)b * T * 3 * H + h * D;

  // load Q^T (scaled)
  {
    const int i = tid & 63, dq = tid >> 6;
    const bool ok = q0 + i < T;
    const TE* src = base + (int64_t)(q0 + i) * 3 * H + dq * 16;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      float4 v = ok ? ld4(src + m * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      const int dd = dq * 16 + m * 4;
      sm.Qt[dd + 0][i] = v.x * scale; sm.Qt[dd + 1][i] = v.y * scale;
      sm.Qt[dd + 2][i] = v.z * scale; sm.Qt[dd + 3][i] = v.w * scale;
    }
  }
  float o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  // softmax role: row = tid/4, quarter = tid%4 (16 columns each)
  const int sr = tid >> 2, sq = tid & 3;
  float m_run = -INFINITY, l_run = 0.f;

  for (int k0 = 0; k0 < kl; k0 += TK) {
    __syncthreads();  // previous tile fully consumed (also orders the Q^T stores on the first pass)
    {
      const int j = tid & 63, dq = tid >> 6;
      const bool ok = k0 + j < kl;
      const TE* ksrc = base + (int64_t)(k0 + j) * 3 * H + H + dq * 16;
      const TE* vsrc = ksrc + H;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 kv = ok ? ld4(ksrc + m * 4) : z;
        const float4 vv = ok ? ld4(vsrc + m * 4) : z;
        const int dd = dq * 16 + m * 4;
        sm.Kt[dd + 0][j] = kv.x; sm.Kt[dd + 1][j] = kv.y; sm.Kt[dd + 2][j] = kv.z; sm.Kt[dd + 3][j] = kv.w;
        *reinterpret_cast<float4*>(&sm.Vs[j][dd]) = vv;
      }
    }
    __syncthreads();
    // S = (Q * scale) K^T
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
    for (int dd = 0; dd < D; ++dd) {
      const float4 qa = *reinterpret_cast<const float4*>(&sm.Qt[dd][ty * 4]);
      const float4 kb = *reinterpret_cast<const float4*>(&sm.Kt[dd][tx * 4]);
      const float qv[4] = {qa.x, qa.y, qa.z, qa.w};
      const float kv[4] = {kb.x, kb.y, kb.z, kb.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(qv[i], kv[j], s[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 v;
      v.x = (k0 + tx * 4 + 0 < kl) ? s[i][0] : -INFINITY;
      v.y = (k0 + tx * 4 + 1 < kl) ? s[i][1] : -INFINITY;
      v.z = (k0 + tx * 4 + 2 < kl) ? s[i][2] : -INFINITY;
      v.w = (k0 + tx * 4 + 3 < kl) ? s[i][3] : -INFINITY;
      *reinterpret_cast<float4*>(&sm.Ss[ty * 4 + i][tx * 4]) = v;
    }
    __syncthreads();
    // online softmax over this tile: 4 threads per row
    {
      float vals[16];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        vals[c] = sm.Ss[sr][sq * 16 + c];
        mx = fmaxf(mx, vals[c]);
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_new = fmaxf(m_run, mx);  // finite: every tile has >= 1 valid key
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const float pexp = expf(vals[c] - m_new);
        sm.Ss[sr][sq * 16 + c] = pexp;
        sum += pexp;
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float a = expf(m_run - m_new);  // exp(-inf) = 0 on the first tile
      l_run = l_run * a + sum;
      m_run = m_new;
      if (sq == 0) sm.alpha[sr] = a;
    }
    __syncthreads();
    // O = alpha * O + P V
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float a = sm.alpha[ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] *= a;
    }
#pragma unroll 8
    for (int j = 0; j < TK; ++j) {
      const float4 vv = *reinterpret_cast<const float4*>(&sm.Vs[j][tx * 4]);
      const float vb[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float pv = sm.Ss[ty * 4 + i][j];
#pragma unroll
        for (int c = 0; c < 4; ++c) o[i][c] = fmaf(pv, vb[c], o[i][c]);
      }
    }
  }
  if (sq == 0) sm.linv[sr] = 1.f / l_run;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = q0 + ty * 4 + i;
    if (t >= T) continue;
    const float li = sm.linv[ty * 4 + i];
    float4 v = make_float4(o[i][0] * li, o[i][1] * li, o[i][2] * li, o[i][3] * li);
    st4(out + ((int64_t)b * T + t) * H + h * D + tx * 4, v);
  }
}

}  // namespace

// tcgen05 arm (attention_tc.cu)
int sib_attention_bf16_tc(const void* qkv, const int32_t* key_len, void* out, int batch, int t, int heads,
                          cudaStream_t stream);

extern "C" int sib_attention(const void* qkv, int dtype, const int32_t* key_len, void* out, int batch, int t, int heads,
                             int head_dim, sib_stream_t stream) {
  SIB_REQUIRE(qkv && out && batch > 0 && t > 0 && heads > 0, "sib_attention: bad argument");
  SIB_REQUIRE(head_dim == D, "sib_attention: head_dim=%d unsupported (64 only)", head_dim);
  SIB_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              "sib_attention: pointers must be 16B aligned");
  SIB_REQUIRE(batch <= 65535 && heads <= 65535, "sib_attention: grid too large");
  static const bool force_simt = getenv("SIB_ATTN_SIMT") != nullptr;  // A/B switch for profiling
  if (dtype == SIB_BF16 && !force_simt)
    return sib_attention_bf16_tc(qkv, key_len, out, batch, t, heads, static_cast<cudaStream_t>(stream));
  static_assert(sizeof(Smem) <= 100 * 1024, "attention smem");
  const void* fn = dtype == SIB_BF16 ? (const void*)attention_kernel<__nv_bfloat16> : (const void*)attention_kernel<float>;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
  if (e != cudaSuccess) {
    sib::set_error("sib_attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return SIB_ERR_CUDA;
  }
  dim3 grid(sib::ceil_div(t, TQ), heads, batch);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == SIB_BF16)
    attention_kernel<__nv_bfloat16><<<grid, 256, sizeof(Smem), s>>>((const __nv_bfloat16*)qkv, key_len,
                                                                    (__nv_bfloat16*)out, t, heads);
  else
    attention_kernel<float><<<grid, 256, sizeof(Smem), s>>>((const float*)qkv, key_len, (float*)out, t, heads);
  SIB_CHECK_LAUNCH("sib_attention");
  return SIB_OK;
}

extern "C" int sib_attention_f32(const float* qkv, const int32_t* key_len, float* out, int batch, int t, int heads,
                                 int head_dim, sib_stream_t stream) {
  return sib_attention(qkv, SIB_F32, key_len, out, batch, t, heads, head_dim, stream);
}
