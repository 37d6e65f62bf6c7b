// Log-mel spectrogram: reflect pad -> hann-1024 STFT -> sqrt(re^2+im^2+1e-9) ->
// mel basis (80x513) -> log(clamp(., 1e-5)).   I_ea/hifi_gan/meldataset.py:49-79 (hop 256, pad 384,
// the mel-L1 metric) and I_ea/dataset/mel_dump.py:40-98 (hop 441, pad 312, the HiFi-GAN features).
//
// Shared-memory staged: a CTA stages the samples of FR consecutive frames once (each sample is read
// from HBM once per CTA instead of n_fft/hop = 4 times), runs a 1024-point radix-2 FFT per frame in
// shared memory and writes only the 80 log-mel values per frame.  Algorithmic traffic:
// 4 B/sample in + 4*80 B/frame out (~5.25 B/sample at hop 256).
#include "common.cuh"

namespace {

constexpr int NFFT = 1024;
constexpr int NBIN = NFFT / 2 + 1;
constexpr int FR = 4;     // frames per CTA
constexpr int NTH = 256;

__global__ void __launch_bounds__(NTH) mel_kernel(const float* __restrict__ wave, int n, int hop, int pad,
                                                  const float* __restrict__ basis, int n_mels,
                                                  float* __restrict__ out, int frames) {
  extern __shared__ float sm[];
  float* re = sm;                  // [NFFT]
  float* im = re + NFFT;           // [NFFT]
  float* twc = im + NFFT;          // [NFFT/2] cos(2 pi k / NFFT)
  float* tws = twc + NFFT / 2;     // [NFFT/2] -sin(2 pi k / NFFT)
  float* win = tws + NFFT / 2;     // [NFFT]
  float* mag = win + NFFT;         // [NBIN + pad]
  float* stage = mag + 520;        // [(FR-1)*hop + NFFT]

  const int b = blockIdx.y;
  const int f0 = blockIdx.x * FR;
  const int nf = min(FR, frames - f0);
  const int tid = threadIdx.x;
  const float* wb = wave + (int64_t)b * n;

  for (int k = tid; k < NFFT / 2; k += NTH) {
    float s, c;
    sincospif(2.0f * (float)k / (float)NFFT, &s, &c);
    twc[k] = c;
    tws[k] = -s;
  }
  for (int i = tid; i < NFFT; i += NTH) {
    // torch.hann_window(1024) (periodic): 0.5 - 0.5 cos(2 pi i / N) = sin^2(pi i / N)
    const float s = sinpif((float)i / (float)NFFT);
    win[i] = s * s;
  }
  const int span = (nf - 1) * hop + NFFT;
  const int start = f0 * hop - pad;  // index into the un-padded signal
  for (int i = tid; i < span; i += NTH) {
    int j = start + i;
    if (j < 0) j = -j;                      // reflect (no edge repeat), F.pad(mode='reflect')
    if (j >= n) j = 2 * (n - 1) - j;
    stage[i] = wb[j];
  }
  __syncthreads();

  for (int f = 0; f < nf; ++f) {
    // bit-reversed load with window
    for (int i = tid; i < NFFT; i += NTH) {
      const int r = __brev((unsigned)i) >> 22;  // 10-bit reversal
      re[r] = stage[f * hop + i] * win[i];
      im[r] = 0.f;
    }
    __syncthreads();
#pragma unroll 1
    for (int s = 1; s <= 10; ++s) {
      const int half = 1 << (s - 1);
      for (int bf = tid; bf < NFFT / 2; bf += NTH) {
        const int grp = bf >> (s - 1), k = bf & (half - 1);
        const int i0 = (grp << s) + k, i1 = i0 + half;
        const int tw = k << (10 - s);
        const float c = twc[tw], sn = tws[tw];
        const float xr = re[i1], xi = im[i1];
        const float tr = xr * c - xi * sn, ti = xr * sn + xi * c;
        const float ur = re[i0], ui = im[i0];
        re[i0] = ur + tr; im[i0] = ui + ti;
        re[i1] = ur - tr; im[i1] = ui - ti;
      }
      __syncthreads();
    }
    for (int k = tid; k < NBIN; k += NTH) mag[k] = sqrtf(re[k] * re[k] + im[k] * im[k] + 1e-9f);
    __syncthreads();
    // mel projection: one warp per mel row, lanes stride over bins
    const int lane = tid & 31, wid = tid >> 5;
    for (int m = wid; m < n_mels; m += NTH / 32) {
      const float* br = basis + (int64_t)m * NBIN;
      float acc = 0.f;
      for (int k = lane; k < NBIN; k += 32) acc = fmaf(__ldg(br + k), mag[k], acc);
      acc = sib::warp_sum(acc);
      if (lane == 0) out[((int64_t)b * n_mels + m) * frames + f0 + f] = logf(fmaxf(acc, 1e-5f));
    }
    __syncthreads();
  }
}

}  // namespace

extern "C" int sib_mel_spectrogram_f32(const float* wave, int batch, int n, int hop, int pad, const float* mel_basis,
                                       int n_mels, float* out, int frames, sib_stream_t stream) {
  SIB_REQUIRE(wave && mel_basis && out && batch > 0 && batch <= 65535 && n > 0 && hop > 0 && pad >= 0 && n_mels > 0,
              "sib_mel_spectrogram_f32: bad argument");
  SIB_REQUIRE(pad < n, "sib_mel_spectrogram_f32: reflect pad %d must be < n=%d", pad, n);
  SIB_REQUIRE(n + 2 * pad >= NFFT, "sib_mel_spectrogram_f32: signal too short for n_fft=1024");
  const int expect = 1 + (n + 2 * pad - NFFT) / hop;
  SIB_REQUIRE(frames == expect, "sib_mel_spectrogram_f32: frames=%d but shape implies %d", frames, expect);
  SIB_REQUIRE(hop <= 1024, "sib_mel_spectrogram_f32: hop=%d > 1024 unsupported", hop);
  const size_t smem = (size_t)(NFFT * 2 + NFFT + NFFT + 520 + (FR - 1) * hop + NFFT) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    sib::set_error("sib_mel_spectrogram_f32: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return SIB_ERR_CUDA;
  }
  dim3 grid(sib::ceil_div(frames, FR), batch);
  mel_kernel<<<grid, NTH, smem, static_cast<cudaStream_t>(stream)>>>(wave, n, hop, pad, mel_basis, n_mels, out, frames);
  SIB_CHECK_LAUNCH("sib_mel_spectrogram_f32");
  return SIB_OK;
}
