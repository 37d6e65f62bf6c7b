// Glue kernels between the two models: ragged mask-frame gather,
// cosine / L2 codebook assignment, centroid paste, extend_mel, embedding concat, int16 pack, casts.
// All are HBM/latency-bound index work: coalesced rows, one warp per row where a reduction is needed.
#include "common.cuh"

namespace {

__global__ void gather_frames_kernel(const float* __restrict__ src, int T, int Dm, const int32_t* __restrict__ pos,
                                     const int32_t* __restrict__ len, const int32_t* __restrict__ off,
                                     float* __restrict__ out) {
  const int b = blockIdx.y;
  const int L = len[b];
  const int64_t total = (int64_t)L * Dm;
  const float* s = src + ((int64_t)b * T + pos[b]) * Dm;
  float* o = out + (int64_t)off[b] * Dm;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    o[i] = s[i];
}

// One CTA (8 warps) per query row; warp w scores centroids w, w+8, ... with the lanes striding the feature axis, so every
// (row, centroid) score is the same sequence of fp32 operations whatever the launch shape.  MODE 0: argmax cosine
// similarity (torch.cosine_similarity, eps 1e-8); MODE 1: argmin squared L2 distance.  Ties resolve to the lowest index
// (torch.argmax / np.argmin) - within a warp by visiting k in increasing order, across warps in the final reduction.
template <int MODE>
__global__ void __launch_bounds__(256) assign_kernel(const float* __restrict__ v, const float* __restrict__ cb, int M,
                                                     int K, int Dm, int64_t* __restrict__ labels) {
  __shared__ float s_best[8];
  __shared__ int s_k[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int m = blockIdx.x;
  if (m >= M) return;
  const float* vr = v + (int64_t)m * Dm;
  float vn = 0.f;
  for (int c = lane; c < Dm; c += 32) vn += vr[c] * vr[c];
  vn = sqrtf(sib::warp_sum(vn));
  float best = MODE == 0 ? -INFINITY : INFINITY;
  int best_k = 0x7fffffff;
  for (int k = w; k < K; k += 8) {
    const float* cr = cb + (int64_t)k * Dm;
    float dot = 0.f, cn = 0.f;
    for (int c = lane; c < Dm; c += 32) {
      const float a = vr[c], bb = cr[c];
      if (MODE == 0) {
        dot = fmaf(a, bb, dot);
        cn = fmaf(bb, bb, cn);
      } else {
        const float dlt = a - bb;
        dot = fmaf(dlt, dlt, dot);
      }
    }
    dot = sib::warp_sum(dot);
    float score;
    if (MODE == 0) {
      cn = sqrtf(sib::warp_sum(cn));
      score = dot / (fmaxf(vn, 1e-8f) * fmaxf(cn, 1e-8f));
      if (score > best) { best = score; best_k = k; }
    } else {
      score = dot;
      if (score < best) { best = score; best_k = k; }
    }
  }
  if (lane == 0) { s_best[w] = best; s_k[w] = best_k; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float b = s_best[0];
    int bk = s_k[0];
    for (int i = 1; i < 8; ++i) {
      const float sc = s_best[i];
      const int sk = s_k[i];
      const bool better = MODE == 0 ? (sc > b) : (sc < b);
      if (better || (sc == b && sk < bk)) { b = sc; bk = sk; }
    }
    labels[m] = bk == 0x7fffffff ? 0 : bk;
  }
}

__global__ void paste_centroids_kernel(float* __restrict__ mel, int Dm, int T, const float* __restrict__ cc,
                                       const float* __restrict__ center, const int64_t* __restrict__ labels,
                                       const int32_t* __restrict__ pos, const int32_t* __restrict__ len,
                                       const int32_t* __restrict__ off, int K) {
  const int b = blockIdx.y;
  const int L = len[b], p0 = pos[b];
  float* mb = mel + (int64_t)b * Dm * T;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L * Dm; i += gridDim.x * blockDim.x) {
    const int c = i / L, f = i - c * L;
    if (p0 + f < 0 || p0 + f >= T) continue;
    const int64_t lab = labels[off[b] + f];
    if (lab < 0 || lab >= K) {   // as nn.Embedding / tensor indexing: an out-of-range index is a device-side assert
      printf("sib_paste_centroids_f32: label %lld outside the %d-entry codebook\n", (long long)lab, K);
      __trap();
    }
    mb[(int64_t)c * T + p0 + f] = cc[lab * Dm + c] + center[c];
  }
}

// F.interpolate(bilinear, scale_factor=(1, 441/256), align_corners=False) along time:
// src = max(0, (dst + 0.5) * (256/441) - 0.5); i0 = floor(src); i1 = min(i0+1, T-1).
__device__ __forceinline__ void sto(float* p, float v) { *p = v; }
__device__ __forceinline__ void sto(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename TY>
__global__ void extend_mel_kernel(const float* __restrict__ in, TY* __restrict__ out, int Dm, int T, int Tm,
                                  int frame_major) {
  const int b = blockIdx.y;
  const float scale = (float)(1.0 / (441.0 / 256.0));
  const float* ib = in + (int64_t)b * Dm * T;
  TY* ob = out + (int64_t)b * Dm * Tm;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Dm * Tm; i += gridDim.x * blockDim.x) {
    int c, t;
    if (frame_major) { t = i / Dm; c = i - t * Dm; } else { c = i / Tm; t = i - c * Tm; }
    float src = ((float)t + 0.5f) * scale - 0.5f;
    src = src < 0.f ? 0.f : src;
    int i0 = (int)src;
    i0 = min(i0, T - 1);
    const int i1 = min(i0 + 1, T - 1);
    const float w1 = src - (float)i0, w0 = 1.f - w1;
    const float* row = ib + (int64_t)c * T;
    sto(ob + i, w0 * row[i0] + w1 * row[i1]);
  }
}

// 32x32 smem-tiled transpose: in [B][R][Cc] -> out [B][Cc][R]
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int Cc) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const float* ib = in + (int64_t)b * R * Cc;
  float* ob = out + (int64_t)b * R * Cc;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < Cc) ? ib[(int64_t)r * Cc + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < Cc) ob[(int64_t)c * R + r] = tile[threadIdx.x][i];
  }
}

__global__ void embed_concat_kernel(const int64_t* __restrict__ code, const int64_t* __restrict__ zp,
                                    const float* __restrict__ spk, const float* __restrict__ emb_c,
                                    const float* __restrict__ emb_p, float* __restrict__ out, int T, int Tp, int E,
                                    int Es, int n_codes, int n_bins) {
  const int b = blockIdx.y, t = blockIdx.x;
  const int W = 2 * E + Es;
  const int rep = T / Tp;  // _upsample repeats each pitch step T // Tp times (model.py:104)
  const int64_t ci = code[(int64_t)b * T + t];
  const int64_t pi = zp[(int64_t)b * Tp + min(t / rep, Tp - 1)];
  if (ci < 0 || ci >= n_codes || pi < 0 || pi >= n_bins) {   // nn.Embedding raises a device assert in the same case
    if (threadIdx.x == 0)
      printf("sib_embed_concat_f32: code %lld / pitch bin %lld outside the %d / %d-row embedding tables\n", (long long)ci,
             (long long)pi, n_codes, n_bins);
    __trap();
  }
  float* o = out + ((int64_t)b * T + t) * W;
  for (int i = threadIdx.x; i < W; i += blockDim.x) {
    float v;
    if (i < E) v = emb_c[ci * E + i];
    else if (i < 2 * E) v = emb_p[pi * E + (i - E)];
    else v = spk[(int64_t)b * Es + (i - 2 * E)];
    o[i] = v;
  }
}

__global__ void pack_int16_kernel(const float* __restrict__ y, int16_t* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    // numpy float32 -> int16 astype: C conversion through a wider int, truncation toward zero
    out[i] = (int16_t)(int32_t)(y[i] * 32768.0f);
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __bfloat162float(in[i]);
}

inline unsigned ew_grid(int64_t n) {
  int64_t g = (n + 255) / 256;
  return (unsigned)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

}  // namespace

extern "C" int sib_gather_frames_f32(const float* src, int batch, int t, int d, const int32_t* pos,
                                     const int32_t* len, const int32_t* off, float* out, sib_stream_t stream) {
  SIB_REQUIRE(src && pos && len && off && out && batch > 0 && batch <= 65535 && t > 0 && d > 0,
              "sib_gather_frames_f32: bad argument");
  gather_frames_kernel<<<dim3(8, batch), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, t, d, pos, len, off, out);
  SIB_CHECK_LAUNCH("sib_gather_frames_f32");
  return SIB_OK;
}

extern "C" int sib_cos_argmax_f32(const float* v, const float* cc, int m, int k, int d, int64_t* labels,
                                  sib_stream_t stream) {
  SIB_REQUIRE(v && cc && labels && m > 0 && k > 0 && d > 0, "sib_cos_argmax_f32: bad argument");
  assign_kernel<0><<<m, 256, 0, static_cast<cudaStream_t>(stream)>>>(v, cc, m, k, d, labels);
  SIB_CHECK_LAUNCH("sib_cos_argmax_f32");
  return SIB_OK;
}

extern "C" int sib_l2_argmin_f32(const float* f, const float* mu, int m, int k, int d, int64_t* labels,
                                 sib_stream_t stream) {
  SIB_REQUIRE(f && mu && labels && m > 0 && k > 0 && d > 0, "sib_l2_argmin_f32: bad argument");
  assign_kernel<1><<<m, 256, 0, static_cast<cudaStream_t>(stream)>>>(f, mu, m, k, d, labels);
  SIB_CHECK_LAUNCH("sib_l2_argmin_f32");
  return SIB_OK;
}

extern "C" int sib_paste_centroids_f32(float* mel, int batch, int d, int t, const float* cc, const float* center,
                                       const int64_t* labels, const int32_t* pos, const int32_t* len,
                                       const int32_t* off, int k, sib_stream_t stream) {
  SIB_REQUIRE(mel && cc && center && labels && pos && len && off && batch > 0 && batch <= 65535 && d > 0 && t > 0 && k > 0,
              "sib_paste_centroids_f32: bad argument");
  paste_centroids_kernel<<<dim3(8, batch), 256, 0, static_cast<cudaStream_t>(stream)>>>(mel, d, t, cc, center, labels,
                                                                                        pos, len, off, k);
  SIB_CHECK_LAUNCH("sib_paste_centroids_f32");
  return SIB_OK;
}

extern "C" int sib_extend_mel(const float* in, void* out, int out_dtype, int batch, int d, int t, int tm, int frame_major,
                              sib_stream_t stream) {
  SIB_REQUIRE(in && out && batch > 0 && batch <= 65535 && d > 0 && t > 0 && tm > 0, "sib_extend_mel: bad argument");
  dim3 grid(sib::ceil_div((int64_t)d * tm, 256), batch);
  if (out_dtype == SIB_BF16)
    extend_mel_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, (__nv_bfloat16*)out, d, t, tm, frame_major);
  else
    extend_mel_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, (float*)out, d, t, tm, frame_major);
  SIB_CHECK_LAUNCH("sib_extend_mel");
  return SIB_OK;
}

extern "C" int sib_extend_mel_f32(const float* in, float* out, int batch, int d, int t, int tm, int frame_major,
                                  sib_stream_t stream) {
  return sib_extend_mel(in, out, SIB_F32, batch, d, t, tm, frame_major, stream);
}

extern "C" int sib_transpose_f32(const float* in, float* out, int batch, int rows, int cols, sib_stream_t stream) {
  SIB_REQUIRE(in && out && batch > 0 && batch <= 65535 && rows > 0 && cols > 0, "sib_transpose_f32: bad argument");
  SIB_REQUIRE(sib::ceil_div(rows, 32) <= 65535, "sib_transpose_f32: rows too large");
  dim3 grid(sib::ceil_div(cols, 32), sib::ceil_div(rows, 32), batch);
  transpose_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(in, out, rows, cols);
  SIB_CHECK_LAUNCH("sib_transpose_f32");
  return SIB_OK;
}

extern "C" int sib_embed_concat_f32(const int64_t* code, const int64_t* zp, const float* spk, const float* emb_c,
                                    const float* emb_p, float* out, int batch, int t, int t_p, int e, int e_spk,
                                    int n_codes, int n_bins, sib_stream_t stream) {
  SIB_REQUIRE(code && zp && spk && emb_c && emb_p && out && batch > 0 && batch <= 65535 && t > 0 && t_p > 0 && e > 0 &&
                  n_codes > 0 && n_bins > 0,
              "sib_embed_concat_f32: bad argument");
  SIB_REQUIRE(t >= t_p && (t - t_p * (t / t_p)) / (t / t_p) == 0,
              "sib_embed_concat_f32: misaligned condition lengths t=%d t_p=%d (model.py:110-114)", t, t_p);
  embed_concat_kernel<<<dim3(t, batch), 128, 0, static_cast<cudaStream_t>(stream)>>>(code, zp, spk, emb_c, emb_p, out,
                                                                                     t, t_p, e, e_spk, n_codes, n_bins);
  SIB_CHECK_LAUNCH("sib_embed_concat_f32");
  return SIB_OK;
}

extern "C" int sib_pack_int16_f32(const float* y, int16_t* out, int64_t n, sib_stream_t stream) {
  SIB_REQUIRE(y && out && n > 0, "sib_pack_int16_f32: bad argument");
  pack_int16_kernel<<<ew_grid(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(y, out, n);
  SIB_CHECK_LAUNCH("sib_pack_int16_f32");
  return SIB_OK;
}

extern "C" int sib_cast_f32_to_bf16(const float* in, void* out, int64_t n, sib_stream_t stream) {
  SIB_REQUIRE(in && out && n > 0, "sib_cast_f32_to_bf16: bad argument");
  cast_f32_bf16_kernel<<<ew_grid(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, (__nv_bfloat16*)out, n);
  SIB_CHECK_LAUNCH("sib_cast_f32_to_bf16");
  return SIB_OK;
}

// zeroes the dataflow counters (sib_flow) at the head of a launch chain; a memset node orders itself after everything
// queued before it, so a previous replay of the same plan has drained its counters by then
extern "C" int sib_fill_zero(void* p, int64_t bytes, sib_stream_t stream) {
  SIB_REQUIRE(p && bytes > 0, "sib_fill_zero: bad argument");
  cudaError_t e = cudaMemsetAsync(p, 0, (size_t)bytes, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) {
    sib::set_error("sib_fill_zero: %s", cudaGetErrorString(e));
    return SIB_ERR_CUDA;
  }
  return SIB_OK;
}

extern "C" int sib_cast_bf16_to_f32(const void* in, float* out, int64_t n, sib_stream_t stream) {
  SIB_REQUIRE(in && out && n > 0, "sib_cast_bf16_to_f32: bad argument");
  cast_bf16_f32_kernel<<<ew_grid(n), 256, 0, static_cast<cudaStream_t>(stream)>>>((const __nv_bfloat16*)in, out, n);
  SIB_CHECK_LAUNCH("sib_cast_bf16_to_f32");
  return SIB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Row-wise linear layer with a narrow output, y[m, :] = x[m, :] @ w + b (w [K][N], N <= 128): the head of CustomModel
// (I_ea/model.py:75-78,88: Linear(H, 80)) on the sum(L) gathered mask frames.  M is a few hundred rows, so the tiled
// GEMM kernels would run on a handful of CTAs; here one CTA owns one row, KSPLIT groups of N threads stride the
// reduction axis (coalesced w rows, L2-resident) and the partial sums meet in shared memory in a fixed order.
namespace {
constexpr int SK_MAXN = 128, SK_THREADS = 512;
__global__ void __launch_bounds__(SK_THREADS) linear_skinny_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                  const float* __restrict__ bias, float* __restrict__ y,
                                                                  int K, int N) {
  extern __shared__ float sk_smem[];
  float* xs = sk_smem;                 // [K]
  float* part = sk_smem + K;           // [ksplit][N]
  const int m = blockIdx.x;
  const float* xr = x + (int64_t)m * K;
  for (int k = threadIdx.x; k < K; k += SK_THREADS) xs[k] = xr[k];
  __syncthreads();
  const int ksplit = SK_THREADS / N;   // >= 4 for N <= 128
  const int grp = threadIdx.x / N, n = threadIdx.x - grp * N;
  if (grp < ksplit) {
    float a0 = 0.f, a1 = 0.f;
    int k = grp;
    for (; k + ksplit < K; k += 2 * ksplit) {
      a0 = fmaf(xs[k], __ldg(w + (int64_t)k * N + n), a0);
      a1 = fmaf(xs[k + ksplit], __ldg(w + (int64_t)(k + ksplit) * N + n), a1);
    }
    if (k < K) a0 = fmaf(xs[k], __ldg(w + (int64_t)k * N + n), a0);
    part[grp * N + n] = a0 + a1;
  }
  __syncthreads();
  if (threadIdx.x < N) {
    float acc = bias ? bias[threadIdx.x] : 0.f;
    for (int g = 0; g < ksplit; ++g) acc += part[g * N + threadIdx.x];
    y[(int64_t)m * N + threadIdx.x] = acc;
  }
}
}  // namespace

extern "C" int sib_linear_skinny_f32(const float* x, const float* w, const float* bias, float* y, int m, int k, int n,
                                     sib_stream_t stream) {
  SIB_REQUIRE(x && w && y && m > 0 && k > 0 && n > 0, "sib_linear_skinny_f32: bad argument");
  SIB_REQUIRE(n <= SK_MAXN, "sib_linear_skinny_f32: n=%d > %d (use sib_conv1d_f32 for wide outputs)", n, SK_MAXN);
  const size_t smem = ((size_t)k + (size_t)(SK_THREADS / n) * n) * sizeof(float);
  SIB_REQUIRE(smem <= 48 * 1024, "sib_linear_skinny_f32: k=%d too large", k);
  linear_skinny_kernel<<<m, SK_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(x, w, bias, y, k, n);
  SIB_CHECK_LAUNCH("sib_linear_skinny_f32");
  return SIB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// torch weight_norm folded on the device: w = v * (g / ||v||), the norm taken over every dimension but `dim`
// (remove_weight_norm(), I_ea/hifi_gan/models.py:125-132 -> dim 0, g [C0,1,1]; HF pos-conv HF:59-78 -> dim 2, g [1,1,k]).
// One launch per tensor at load / pack time instead of the pow / sum / sqrt / div / mul chain of eager ops.
namespace {
// dim 0: v [rows][inner], one CTA per row
__global__ void __launch_bounds__(256) weight_norm_rows_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                                               float* __restrict__ w, int inner) {
  __shared__ float part[8];
  const int r = blockIdx.x;
  const float* vr = v + (int64_t)r * inner;
  float s = 0.f;
  for (int i = threadIdx.x; i < inner; i += 256) s = fmaf(vr[i], vr[i], s);
  s = sib::warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += part[i];
  const float scale = g[r] / sqrtf(tot);
  for (int i = threadIdx.x; i < inner; i += 256) w[(int64_t)r * inner + i] = vr[i] * scale;
}
// last dim: v [outer][k]; CTA = 32 consecutive k (lanes, coalesced) x 8 warps striding the outer index
__global__ void __launch_bounds__(256) weight_norm_last_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                                               float* __restrict__ w, int outer, int k) {
  __shared__ float part[8][33];
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (j < k)
    for (int o = wi; o < outer; o += 8) {
      const float x = v[(int64_t)o * k + j];
      s = fmaf(x, x, s);
    }
  part[wi][lane] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += part[i][lane];
  if (j < k) {
    const float scale = g[j] / sqrtf(tot);
    for (int o = wi; o < outer; o += 8) w[(int64_t)o * k + j] = v[(int64_t)o * k + j] * scale;
  }
}
}  // namespace

extern "C" int sib_weight_norm_fold_f32(const float* v, const float* g, float* w, int outer, int inner, int norm_dim_last,
                                        sib_stream_t stream) {
  SIB_REQUIRE(v && g && w && outer > 0 && inner > 0, "sib_weight_norm_fold_f32: bad argument");
  if (norm_dim_last)
    weight_norm_last_kernel<<<sib::ceil_div(inner, 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(v, g, w, outer, inner);
  else
    weight_norm_rows_kernel<<<outer, 256, 0, static_cast<cudaStream_t>(stream)>>>(v, g, w, inner);
  SIB_CHECK_LAUNCH("sib_weight_norm_fold_f32");
  return SIB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// k-means assignment at scale (a17, I_da/scripts/inpainting.py:204-205: `kmeans_model.predict(feats)`, T x H features
// against K x H centres).  argmin_k ||f - mu_k||^2 = argmax_k (f . mu_k - 0.5 ||mu_k||^2): the dot products are one
// fp32 GEMM on the tiled SIMT kernel (sib_conv1d_f32 with bias = -0.5 ||mu_k||^2: the exact form sklearn itself
// evaluates for float32 features), the two kernels below are its bookends.  The one-CTA-per-row kernel above walks
// K x H per row without any reuse: 3.1 ms for 12 736 x 768 vs 500 centres, 21 % of an I_da step before this.
namespace {
// out[r] = scale * sum_c x[r, c]^2, one warp per row
__global__ void __launch_bounds__(256) row_sqnorm_kernel(const float* __restrict__ x, int rows, int d, float scale,
                                                         float* __restrict__ out) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* xr = x + (int64_t)r * d;
  float s = 0.f;
  for (int c = lane; c < d; c += 32) s = fmaf(xr[c], xr[c], s);
  s = sib::warp_sum(s);
  if (lane == 0) out[r] = scale * s;
}
// labels[r] = argmax_k s[r, k], ties -> lowest k; one warp per row
__global__ void __launch_bounds__(256) row_argmax_kernel(const float* __restrict__ s, int rows, int k,
                                                         int64_t* __restrict__ labels) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* sr = s + (int64_t)r * k;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int c = lane; c < k; c += 32) {
    const float v = sr[c];
    if (v > best) { best = v; bi = c; }           // increasing c per lane: first maximum wins
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0) labels[r] = bi == 0x7fffffff ? 0 : bi;
}
}  // namespace

extern "C" int sib_row_sqnorm_f32(const float* x, int rows, int d, float scale, float* out, sib_stream_t stream) {
  SIB_REQUIRE(x && out && rows > 0 && d > 0, "sib_row_sqnorm_f32: bad argument");
  row_sqnorm_kernel<<<sib::ceil_div(rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, rows, d, scale, out);
  SIB_CHECK_LAUNCH("sib_row_sqnorm_f32");
  return SIB_OK;
}

extern "C" int sib_row_argmax_f32(const float* s, int rows, int k, int64_t* labels, sib_stream_t stream) {
  SIB_REQUIRE(s && labels && rows > 0 && k > 0, "sib_row_argmax_f32: bad argument");
  row_argmax_kernel<<<sib::ceil_div(rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(s, rows, k, labels);
  SIB_CHECK_LAUNCH("sib_row_argmax_f32");
  return SIB_OK;
}
