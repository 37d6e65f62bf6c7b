// fp32 frame-major implicit-GEMM convolution (SIMT FFMA path).
//
// This is the full-precision arm of the hot path: every Conv1d / ConvTranspose1d / Linear of
// HuBERT (HF:106-175, 83-92, 228-229, 320-343, 363-367) and HiFi-GAN (I_ea/hifi_gan/models.py:
// 36-43, 108-121) runs through one kernel whose rows are time steps and whose reduction runs over
// (tap, input channel).  No im2col buffer is materialised: the A tile for tap j is simply the
// activation rows shifted by tap_offset[j], zero-filled outside [0, t_in).  The bf16 tcgen05 arm
// (conv_tc.cu) uses the same decomposition with TMA doing the shifted loads.
#include "common.cuh"

namespace {

constexpr int BM = 128;  // time rows per CTA
constexpr int BN = 64;   // output channels per CTA
constexpr int BK = 16;   // input channels per pipeline step
constexpr int NT = 256;
constexpr int AS = BM + 4;
constexpr int BS = BN + 4;

struct ConvArgs {
  sib_conv_desc d;
  const float* x;
  const float* w;
  const float* bias;
  const float* res;
  float* y;
  int cin_g, cout_g;
};

template <bool VEC>
__global__ void __launch_bounds__(NT) conv1d_f32_kernel(const __grid_constant__ ConvArgs p) {
  __shared__ __align__(16) float As[2][BK][AS];
  __shared__ __align__(16) float Bs[2][BK][BS];

  const sib_conv_desc& d = p.d;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int t0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int b = blockIdx.z / d.groups, g = blockIdx.z % d.groups;
  const int cin_g = p.cin_g, cout_g = p.cout_g;

  const float* xb = p.x + (int64_t)b * d.x_batch_stride + (int64_t)g * cin_g;
  const float* wg = p.w + (int64_t)g * d.n_taps * cin_g * cout_g;

  // A-load role: one time row, 8 consecutive channels
  const int a_r = tid & (BM - 1), a_kh = tid >> 7;
  const bool a_row_ok = (t0 + a_r) < d.t_out;
  const int64_t a_tbase = (int64_t)(t0 + a_r) * d.stride;
  // B-load role: one k row, 4 consecutive output channels
  const int b_k = tid >> 4, b_n = (tid & 15) * 4;

  const int kchunks = (cin_g + BK - 1) / BK;
  const int iters = d.n_taps * kchunks;

  float a_reg[8], b_reg[4];
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  auto load_tile = [&](int it) {
    const int j = it / kchunks, kc = (it - j * kchunks) * BK;
    // ---- A
    const int64_t tin = a_tbase + d.tap_offset[j];
    const bool ok = a_row_ok && tin >= 0 && tin < d.t_in;
    const int c = kc + a_kh * 8;
    const float* src = xb + tin * d.x_row_stride + c;
    if (VEC) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok && c + h * 4 < cin_g) v = __ldg(reinterpret_cast<const float4*>(src + h * 4));
        a_reg[h * 4 + 0] = v.x; a_reg[h * 4 + 1] = v.y; a_reg[h * 4 + 2] = v.z; a_reg[h * 4 + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) a_reg[i] = (ok && c + i < cin_g) ? __ldg(src + i) : 0.f;
    }
    if (d.pre_act == SIB_ACT_LRELU) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a_reg[i] = a_reg[i] > 0.f ? a_reg[i] : a_reg[i] * d.pre_slope;
    }
    // ---- B
    const int kk = kc + b_k;
    const float* wsrc = wg + ((int64_t)j * cin_g + kk) * cout_g + n0 + b_n;
    if (VEC) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kk < cin_g && n0 + b_n < cout_g) v = __ldg(reinterpret_cast<const float4*>(wsrc));
      b_reg[0] = v.x; b_reg[1] = v.y; b_reg[2] = v.z; b_reg[3] = v.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) b_reg[i] = (kk < cin_g && n0 + b_n + i < cout_g) ? __ldg(wsrc + i) : 0.f;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) As[buf][a_kh * 8 + i][a_r] = a_reg[i];
    *reinterpret_cast<float4*>(&Bs[buf][b_k][b_n]) = make_float4(b_reg[0], b_reg[1], b_reg[2], b_reg[3]);
  };

  load_tile(0);
  store_tile(0);
  __syncthreads();

  for (int it = 0; it < iters; ++it) {
    const int cur = it & 1;
    if (it + 1 < iters) load_tile(it + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 8 + 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (it + 1 < iters) store_tile(cur ^ 1);
    __syncthreads();
  }

  // ---- epilogue
  const int n = n0 + tx * 4;
  if (n >= cout_g) return;
  const int nc = g * cout_g + n;
  float bias4[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n + j < cout_g) bias4[j] = __ldg(p.bias + nc + j);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int t = t0 + ty * 8 + i;
    if (t >= d.t_out) break;
    float* yp = p.y + (int64_t)b * d.y_batch_stride + (int64_t)t * d.y_row_stride + nc;
    const float* rp = p.res ? p.res + (int64_t)b * d.r_batch_stride + (int64_t)t * d.r_row_stride + nc : nullptr;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bias4[j];
    const bool res_pre = rp && !d.res_after_act, res_post = rp && d.res_after_act;
    if (VEC && n + 3 < cout_g) {
      if (res_pre) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(rp));
        v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
      }
      if (d.accumulate) {
        const float4 o = *reinterpret_cast<const float4*>(yp);
        v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = sib::apply_act(v[j] * d.out_scale, d.post_act, d.post_slope);
      if (res_post) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(rp));
        v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
      }
      *reinterpret_cast<float4*>(yp) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (n + j >= cout_g) break;
        float u = v[j];
        if (res_pre) u += __ldg(rp + j);
        if (d.accumulate) u += yp[j];
        u = sib::apply_act(u * d.out_scale, d.post_act, d.post_slope);
        if (res_post) u += __ldg(rp + j);
        yp[j] = u;
      }
    }
  }
}

// One output channel: each thread owns one time step; the k*C weights sit in shared memory.
__device__ __forceinline__ float ldx(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ldx(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename TX>
__global__ void __launch_bounds__(256) conv1d_cout1_kernel(const TX* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ bias, float* __restrict__ y,
                                                           int T, int C, int K, int pad, float pre_slope,
                                                           int post_act) {
  extern __shared__ float ws[];
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const TX* xb = x + (int64_t)b * T * C;
  float acc = bias ? bias[0] : 0.f;
  for (int j = 0; j < K; ++j) {
    const int ti = t + j - pad;
    if (ti < 0 || ti >= T) continue;
    const TX* xr = xb + (int64_t)ti * C;
    const float* wr = ws + j * C;
    if (sizeof(TX) == 4 && (C & 3) == 0) {
      for (int c = 0; c < C; c += 4) {
        float4 v = __ldg(reinterpret_cast<const float4*>(xr + c));
        v.x = v.x > 0.f ? v.x : v.x * pre_slope; v.y = v.y > 0.f ? v.y : v.y * pre_slope;
        v.z = v.z > 0.f ? v.z : v.z * pre_slope; v.w = v.w > 0.f ? v.w : v.w * pre_slope;
        acc = fmaf(v.x, wr[c], acc); acc = fmaf(v.y, wr[c + 1], acc);
        acc = fmaf(v.z, wr[c + 2], acc); acc = fmaf(v.w, wr[c + 3], acc);
      }
    } else if (sizeof(TX) == 2 && (C & 7) == 0) {
      for (int c = 0; c < C; c += 8) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(xr + c));
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float2 f = __bfloat1622float2(h2[i]);
          f.x = f.x > 0.f ? f.x : f.x * pre_slope; f.y = f.y > 0.f ? f.y : f.y * pre_slope;
          acc = fmaf(f.x, wr[c + 2 * i], acc); acc = fmaf(f.y, wr[c + 2 * i + 1], acc);
        }
      }
    } else {
      for (int c = 0; c < C; ++c) {
        float v = ldx(xr + c);
        v = v > 0.f ? v : v * pre_slope;
        acc = fmaf(v, wr[c], acc);
      }
    }
  }
  y[(int64_t)b * T + t] = sib::apply_act(acc, post_act, 0.f);
}

// bf16 input, C % 8 == 0, K <= 8: the k-tap dot product is split per INPUT row.  Thread r loads row (t0 - pad + r) once
// (16-byte loads), forms its K partial products <x_row, w_j> against broadcast float4 weight loads and parks them in
// shared memory; output t then sums K neighbouring partials.  Every activation byte is read once per CTA instead of
// K times and the per-output shared-memory traffic drops from K*C scalar weight loads to K*C/4 vector ones.
constexpr int CO1_ROWS = 2, CO1_THREADS = 256, CO1_TILE = CO1_ROWS * CO1_THREADS, CO1_MAXK = 8;
__global__ void __launch_bounds__(CO1_THREADS) conv1d_cout1_bf16_rows_kernel(const __nv_bfloat16* __restrict__ x,
                                                                            const float* __restrict__ w,
                                                                            const float* __restrict__ bias,
                                                                            float* __restrict__ y, int T, int C, int K, int pad,
                                                                            float pre_slope, int post_act) {
  extern __shared__ float sm[];
  float* ws = sm;                                  // [K][C]
  float* part = sm + K * C;                        // [K][CO1_TILE + CO1_MAXK]
  constexpr int PW = CO1_TILE + CO1_MAXK;
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int b = blockIdx.y;
  // a CTA produces CO1_TILE - (K - 1) outputs from exactly CO1_TILE input rows.  r2: a thread owns CO1_ROWS rows (r, r + 256,
  // ...) and uses every broadcast weight load for all of them - the kernel was bound by the load / store unit (ncu: LSU
  // wavefronts 88 %, 56 weight loads of 16 bytes per row), not by HBM: 81 -> 70.6 us with two rows, 73.6 with four
  const int tile_out = CO1_TILE - (K - 1);
  const int t0 = blockIdx.x * tile_out;
  const __nv_bfloat16* xb = x + (int64_t)b * T * C;
  {
    const int r = threadIdx.x;
    float p[CO1_ROWS][CO1_MAXK];
    const uint4* xr[CO1_ROWS];
    bool in[CO1_ROWS];
#pragma unroll
    for (int q = 0; q < CO1_ROWS; ++q) {
      const int ti = t0 - pad + r + q * CO1_THREADS;
      in[q] = ti >= 0 && ti < T;
      xr[q] = reinterpret_cast<const uint4*>(xb + (int64_t)(in[q] ? ti : 0) * C);
#pragma unroll
      for (int j = 0; j < CO1_MAXK; ++j) p[q][j] = 0.f;
    }
    for (int c8 = 0; c8 < (C >> 3); ++c8) {
      float f[CO1_ROWS][8];
#pragma unroll
      for (int q = 0; q < CO1_ROWS; ++q) {
        const uint4 raw = in[q] ? __ldg(xr[q] + c8) : make_uint4(0u, 0u, 0u, 0u);
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 v = __bfloat1622float2(h2[i]);
          f[q][2 * i] = v.x > 0.f ? v.x : v.x * pre_slope;
          f[q][2 * i + 1] = v.y > 0.f ? v.y : v.y * pre_slope;
        }
      }
#pragma unroll
      for (int j = 0; j < CO1_MAXK; ++j) {
        if (j < K) {
          const float4 wa = *reinterpret_cast<const float4*>(ws + j * C + c8 * 8);
          const float4 wb = *reinterpret_cast<const float4*>(ws + j * C + c8 * 8 + 4);
#pragma unroll
          for (int q = 0; q < CO1_ROWS; ++q) {
            float a = p[q][j];
            a = fmaf(f[q][0], wa.x, a); a = fmaf(f[q][1], wa.y, a); a = fmaf(f[q][2], wa.z, a); a = fmaf(f[q][3], wa.w, a);
            a = fmaf(f[q][4], wb.x, a); a = fmaf(f[q][5], wb.y, a); a = fmaf(f[q][6], wb.z, a); a = fmaf(f[q][7], wb.w, a);
            p[q][j] = a;
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < CO1_ROWS; ++q)
#pragma unroll
      for (int j = 0; j < CO1_MAXK; ++j)
        if (j < K) part[j * PW + r + q * CO1_THREADS] = p[q][j];
  }
  __syncthreads();
  for (int o = threadIdx.x; o < tile_out; o += CO1_THREADS) {
    const int t = t0 + o;
    if (t >= T) break;
    float acc = bias ? bias[0] : 0.f;
#pragma unroll
    for (int j = 0; j < CO1_MAXK; ++j)
      if (j < K) acc += part[j * PW + o + j];   // input row t + j - pad sits at local index o + j
    y[(int64_t)b * T + t] = sib::apply_act(acc, post_act, 0.f);
  }
}

}  // namespace

extern "C" int sib_conv1d_f32(const sib_conv_desc* d, const float* x, const float* w, const float* bias,
                              const float* residual, float* y, sib_stream_t stream) {
  SIB_REQUIRE(d && x && w && y, "sib_conv1d_f32: null argument");
  SIB_REQUIRE(d->batch > 0 && d->t_in > 0 && d->t_out > 0 && d->c_in > 0 && d->c_out > 0, "sib_conv1d_f32: empty shape");
  SIB_REQUIRE(d->groups > 0 && d->c_in % d->groups == 0 && d->c_out % d->groups == 0,
              "sib_conv1d_f32: groups=%d must divide c_in=%d and c_out=%d", d->groups, d->c_in, d->c_out);
  SIB_REQUIRE(d->n_taps > 0 && d->n_taps <= SIB_MAX_TAPS, "sib_conv1d_f32: n_taps=%d out of range", d->n_taps);
  SIB_REQUIRE(d->stride > 0, "sib_conv1d_f32: stride must be positive");
  SIB_REQUIRE(d->pre_act == SIB_ACT_NONE || d->pre_act == SIB_ACT_LRELU, "sib_conv1d_f32: unsupported pre_act");
  SIB_REQUIRE((int64_t)d->batch * d->groups <= 65535, "sib_conv1d_f32: batch*groups too large for grid.z");
  ConvArgs a;
  a.d = *d;
  a.x = x; a.w = w; a.bias = bias; a.res = residual; a.y = y;
  a.cin_g = d->c_in / d->groups;
  a.cout_g = d->c_out / d->groups;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec = al16(x) && al16(w) && al16(y) && (!residual || al16(residual)) && a.cin_g % 4 == 0 &&
                   a.cout_g % 4 == 0 && d->x_row_stride % 4 == 0 && d->x_batch_stride % 4 == 0 &&
                   d->y_row_stride % 4 == 0 && d->y_batch_stride % 4 == 0 &&
                   (!residual || (d->r_row_stride % 4 == 0 && d->r_batch_stride % 4 == 0));
  dim3 grid(sib::ceil_div(d->t_out, BM), sib::ceil_div(a.cout_g, BN), d->batch * d->groups);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (vec)
    conv1d_f32_kernel<true><<<grid, NT, 0, s>>>(a);
  else
    conv1d_f32_kernel<false><<<grid, NT, 0, s>>>(a);
  SIB_CHECK_LAUNCH("sib_conv1d_f32");
  return SIB_OK;
}

extern "C" int sib_conv1d_cout1(const void* x, int x_dtype, const float* w, const float* bias, float* y, int batch, int t,
                                int c, int k, int pad, float pre_slope, int post_act, sib_stream_t stream) {
  SIB_REQUIRE(x && w && y && batch > 0 && t > 0 && c > 0 && k > 0, "sib_conv1d_cout1: bad argument");
  SIB_REQUIRE((size_t)k * c * sizeof(float) <= 48 * 1024, "sib_conv1d_cout1: k*c too large");
  SIB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "sib_conv1d_cout1: x must be 16B aligned");
  dim3 grid(sib::ceil_div(t, 256), batch);
  const size_t smem = (size_t)k * c * sizeof(float);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (x_dtype == SIB_BF16 && c % 8 == 0 && k <= CO1_MAXK) {
    const size_t smem2 = ((size_t)k * c + (size_t)k * (CO1_TILE + CO1_MAXK)) * sizeof(float);
    conv1d_cout1_bf16_rows_kernel<<<dim3(sib::ceil_div(t, CO1_TILE - (k - 1)), batch), CO1_THREADS, smem2, s>>>(
        (const __nv_bfloat16*)x, w, bias, y, t, c, k, pad, pre_slope, post_act);
  } else if (x_dtype == SIB_BF16)
    conv1d_cout1_kernel<<<grid, 256, smem, s>>>((const __nv_bfloat16*)x, w, bias, y, t, c, k, pad, pre_slope, post_act);
  else
    conv1d_cout1_kernel<<<grid, 256, smem, s>>>((const float*)x, w, bias, y, t, c, k, pad, pre_slope, post_act);
  SIB_CHECK_LAUNCH("sib_conv1d_cout1");
  return SIB_OK;
}

extern "C" int sib_conv1d_cout1_f32(const float* x, const float* w, const float* bias, float* y, int batch, int t,
                                    int c, int k, int pad, float pre_slope, int post_act, sib_stream_t stream) {
  return sib_conv1d_cout1(x, SIB_F32, w, bias, y, batch, t, c, k, pad, pre_slope, post_act, stream);
}
