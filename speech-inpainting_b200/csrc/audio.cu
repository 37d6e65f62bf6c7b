// Audio-format kernels either side of the hot path:
//   * rational poly-phase resampler (16 kHz <-> 22.05 kHz = 320 : 441), int16 PCM or fp32 in, fp32 out - replaces the
//     `librosa.load(path, sr=...)` pair of I_ea/predict.py:79-80 and resampy.resample of I_da/scripts/preprocess.py:43-45;
//   * SI-SDR (I_ea/metrics.py:127-141) and the mean absolute difference behind mel-L1 (I_ea/hifi_gan/train.py:224-227),
//     reduced per utterance on the device with double accumulation in a fixed order (bit-reproducible whatever the batch).
// All three are streaming kernels: every input sample is read from HBM once, staged in shared memory where it is reused.
#include "common.cuh"

namespace {

constexpr int RS_R = 8;        // outputs per thread and residue: one filter-tap load feeds RS_R FMAs
constexpr int RS_MAX_THREADS = 512;  // block = residues rounded up to whole warps (320 -> 320, 441 -> 448), capped here
constexpr int RED_CHUNKS = 32; // partial sums per utterance (fixed => deterministic finalisation)

__device__ __forceinline__ float load_sample(const float* p, int64_t i) { return p[i]; }
__device__ __forceinline__ float load_sample(const int16_t* p, int64_t i) { return (float)p[i] * (1.0f / 32768.0f); }

// y[b, n] = sum_j filt[j][n % up] * x[b, (n*down)/up + first + j], x == 0 outside [0, len_in[b]); y == 0 for n >= len_out[b].
// A CTA owns up*RS_R consecutive outputs starting at a multiple of `up`, so its first input index is exact
// (blockIdx.x * RS_R * down) and thread r (= output residue) reuses each of its `taps` weights for RS_R outputs that lie
// `down` input samples apart.  The input span (RS_R*down + taps samples) is staged once in shared memory.
template <typename TIn>
__global__ void __launch_bounds__(RS_MAX_THREADS) resample_kernel(const TIn* __restrict__ x, int n_in, int64_t x_batch_stride,
                                                              const int32_t* __restrict__ len_in,
                                                              const float* __restrict__ filt, int up, int down, int taps,
                                                              int first, float* __restrict__ y, int n_out,
                                                              int64_t y_batch_stride,
                                                              const int32_t* __restrict__ len_out) {
  extern __shared__ float s_x[];
  const int b = blockIdx.y;
  const int li = len_in ? min(len_in[b], n_in) : n_in;
  const int64_t n0 = (int64_t)blockIdx.x * up * RS_R;
  const int64_t q0 = (int64_t)blockIdx.x * RS_R * down + first;   // input index of s_x[0]
  const int span = RS_R * down + taps;
  const TIn* xb = x + (int64_t)b * x_batch_stride;
  for (int i = threadIdx.x; i < span; i += blockDim.x) {
    const int64_t q = q0 + i;
    s_x[i] = (q >= 0 && q < li) ? load_sample(xb, q) : 0.f;
  }
  __syncthreads();
  float* yb = y + (int64_t)b * y_batch_stride;
  const int lo = len_out ? min(len_out[b], n_out) : n_out;   // rows of a padded batch: zeros past the utterance's own end
  for (int r = threadIdx.x; r < up; r += blockDim.x) {
    const int qr = (int)(((int64_t)r * down) / up);
    const float* f = filt + r;
    const float* s = s_x + qr;
    float acc[RS_R];
#pragma unroll
    for (int m = 0; m < RS_R; ++m) acc[m] = 0.f;
    for (int j = 0; j < taps; ++j) {
      const float w = __ldg(f + (int64_t)j * up);
#pragma unroll
      for (int m = 0; m < RS_R; ++m) acc[m] = fmaf(w, s[m * down + j], acc[m]);
    }
#pragma unroll
    for (int m = 0; m < RS_R; ++m) {
      const int64_t n = n0 + r + (int64_t)m * up;
      if (n < n_out) yb[n] = n < lo ? acc[m] : 0.f;
    }
  }
}

__device__ __forceinline__ double block_sum(double v, double* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_red[w] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_red[i];   // same order in every thread
  return t;
}

// pass 0: partial (sum r^2, sum r*e); pass 1: a = (eps + <r,e>) / (<r,r> + eps) from the pass-0 partials, then partial
// (sum (a r)^2, sum (e - a r)^2).  ws[b][chunk][4] doubles.
template <int PASS>
__global__ void __launch_bounds__(256) si_sdr_partial_kernel(const float* __restrict__ est, const float* __restrict__ ref,
                                                             int n, const int32_t* __restrict__ lengths, double eps,
                                                             double* __restrict__ ws) {
  __shared__ double s_red[8];
  const int b = blockIdx.y, ch = blockIdx.x;
  const int len = lengths ? min(lengths[b], n) : n;
  double* wb = ws + (int64_t)b * RED_CHUNKS * 4;
  double a = 0.0;
  if (PASS == 1) {
    double rss = 0.0, rse = 0.0;
    for (int i = 0; i < RED_CHUNKS; ++i) { rss += wb[i * 4 + 0]; rse += wb[i * 4 + 1]; }
    a = (eps + rse) / (rss + eps);
  }
  const int per = (len + RED_CHUNKS - 1) / RED_CHUNKS;
  const int lo = ch * per, hi = min(len, lo + per);
  const float* e = est + (int64_t)b * n;
  const float* r = ref + (int64_t)b * n;
  double s0 = 0.0, s1 = 0.0;
  for (int i = lo + threadIdx.x; i < hi; i += 256) {
    const double rv = r[i], ev = e[i];
    if (PASS == 0) {
      s0 += rv * rv;
      s1 += rv * ev;
    } else {
      const double t = a * rv, d = ev - t;
      s0 += t * t;
      s1 += d * d;
    }
  }
  s0 = block_sum(s0, s_red);
  s1 = block_sum(s1, s_red);
  if (threadIdx.x == 0) {
    wb[ch * 4 + 2 * PASS + 0] = s0;
    wb[ch * 4 + 2 * PASS + 1] = s1;
  }
}

__global__ void si_sdr_final_kernel(const double* __restrict__ ws, int batch, double eps, float* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const double* wb = ws + (int64_t)b * RED_CHUNKS * 4;
  double sss = 0.0, snn = 0.0;
  for (int i = 0; i < RED_CHUNKS; ++i) { sss += wb[i * 4 + 2]; snn += wb[i * 4 + 3]; }
  out[b] = (float)(10.0 * log10((eps + sss) / (eps + snn)));
}

__global__ void __launch_bounds__(256) abs_diff_partial_kernel(const float* __restrict__ a, const float* __restrict__ bsrc,
                                                               int64_t n, double* __restrict__ ws) {
  __shared__ double s_red[8];
  const int b = blockIdx.y, ch = blockIdx.x;
  const int64_t per = (n + RED_CHUNKS - 1) / RED_CHUNKS;
  const int64_t lo = ch * per, hi = min(n, lo + per);
  const float* pa = a + (int64_t)b * n;
  const float* pb = bsrc + (int64_t)b * n;
  double s = 0.0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += 256) s += (double)fabsf(pa[i] - pb[i]);
  s = block_sum(s, s_red);
  if (threadIdx.x == 0) ws[(int64_t)b * RED_CHUNKS + ch] = s;
}

__global__ void abs_diff_final_kernel(const double* __restrict__ ws, int batch, int64_t n, float* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  double s = 0.0;
  for (int i = 0; i < RED_CHUNKS; ++i) s += ws[(int64_t)b * RED_CHUNKS + i];
  out[b] = (float)(s / (double)n);
}

}  // namespace

extern "C" size_t sib_resample_smem_bytes(int down, int taps) { return ((size_t)RS_R * down + taps) * sizeof(float); }

extern "C" int sib_resample(const void* x, int x_dtype, int batch, int n_in, int64_t x_batch_stride,
                            const int32_t* len_in, const float* filt, int up, int down, int taps, int first, float* y,
                            int n_out, int64_t y_batch_stride, const int32_t* len_out, sib_stream_t stream) {
  SIB_REQUIRE(x && filt && y && batch > 0 && batch <= 65535 && n_in > 0 && n_out > 0, "sib_resample: bad argument");
  SIB_REQUIRE(up > 0 && down > 0 && taps > 0, "sib_resample: up=%d down=%d taps=%d must be positive", up, down, taps);
  SIB_REQUIRE(x_dtype == SIB_F32 || x_dtype == SIB_I16, "sib_resample: input must be SIB_F32 or SIB_I16");
  const size_t smem = sib_resample_smem_bytes(down, taps);
  SIB_REQUIRE(smem <= 200 * 1024, "sib_resample: ratio %d:%d with %d taps needs %zu B of shared memory", up, down, taps, smem);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  dim3 grid(sib::ceil_div(n_out, (int64_t)up * RS_R), batch);
  const int RS_THREADS = up >= RS_MAX_THREADS ? RS_MAX_THREADS : ((up + 31) / 32) * 32;
  if (x_dtype == SIB_I16) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(resample_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    resample_kernel<int16_t><<<grid, RS_THREADS, smem, s>>>((const int16_t*)x, n_in, x_batch_stride, len_in, filt, up, down,
                                                            taps, first, y, n_out, y_batch_stride, len_out);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(resample_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    resample_kernel<float><<<grid, RS_THREADS, smem, s>>>((const float*)x, n_in, x_batch_stride, len_in, filt, up, down, taps,
                                                          first, y, n_out, y_batch_stride, len_out);
  }
  SIB_CHECK_LAUNCH("sib_resample");
  return SIB_OK;
}

extern "C" size_t sib_si_sdr_workspace_bytes(int batch) { return (size_t)batch * RED_CHUNKS * 4 * sizeof(double); }

extern "C" int sib_si_sdr_f32(const float* est, const float* ref, int batch, int n, const int32_t* lengths, float eps,
                              float* out, void* workspace, sib_stream_t stream) {
  SIB_REQUIRE(est && ref && out && workspace && batch > 0 && batch <= 65535 && n > 0, "sib_si_sdr_f32: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  double* ws = static_cast<double*>(workspace);
  si_sdr_partial_kernel<0><<<dim3(RED_CHUNKS, batch), 256, 0, s>>>(est, ref, n, lengths, (double)eps, ws);
  SIB_CHECK_LAUNCH("sib_si_sdr_f32");
  si_sdr_partial_kernel<1><<<dim3(RED_CHUNKS, batch), 256, 0, s>>>(est, ref, n, lengths, (double)eps, ws);
  SIB_CHECK_LAUNCH("sib_si_sdr_f32");
  si_sdr_final_kernel<<<sib::ceil_div(batch, 128), 128, 0, s>>>(ws, batch, (double)eps, out);
  SIB_CHECK_LAUNCH("sib_si_sdr_f32");
  return SIB_OK;
}

extern "C" size_t sib_abs_diff_workspace_bytes(int batch) { return (size_t)batch * RED_CHUNKS * sizeof(double); }

extern "C" int sib_abs_diff_mean_f32(const float* a, const float* b, int batch, int64_t n, float* out, void* workspace,
                                     sib_stream_t stream) {
  SIB_REQUIRE(a && b && out && workspace && batch > 0 && batch <= 65535 && n > 0, "sib_abs_diff_mean_f32: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  double* ws = static_cast<double*>(workspace);
  abs_diff_partial_kernel<<<dim3(RED_CHUNKS, batch), 256, 0, s>>>(a, b, n, ws);
  SIB_CHECK_LAUNCH("sib_abs_diff_mean_f32");
  abs_diff_final_kernel<<<sib::ceil_div(batch, 128), 128, 0, s>>>(ws, batch, n, out);
  SIB_CHECK_LAUNCH("sib_abs_diff_mean_f32");
  return SIB_OK;
}
