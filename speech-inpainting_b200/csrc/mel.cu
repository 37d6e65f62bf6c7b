// Log-mel spectrogram: reflect pad -> hann-1024 STFT -> sqrt(re^2+im^2+1e-9) ->
// mel basis (80x513) -> log(clamp(., 1e-5)).   I_ea/hifi_gan/meldataset.py:49-79 (hop 256, pad 384,
// the mel-L1 metric) and I_ea/dataset/mel_dump.py:40-98 (hop 441, pad 312, the HiFi-GAN features).
//
// A CTA (8 warps) stages the samples of 32 consecutive frames once in shared memory (reflect pad folded into the load;
// each sample is read from HBM ~1.2x instead of n_fft/hop = 4x) and every warp transforms TWO frames per pass:
//   * the two real windowed frames ride as the real / imaginary part of ONE 1024-point complex FFT;
//   * the FFT is the 32 x 32 four-step form held in registers: lane n2 runs a 32-point FFT over n1 of x[32 n1 + n2]
//     (fully unrolled radix-2, compile-time twiddles), multiplies by W_1024^(n2 k1), the warp transposes through its
//     private 8 KB of shared memory, lane k1 runs the second 32-point FFT: no block-level barrier, 2 x 80 butterflies
//     per lane;
//   * the two spectra are separated by conjugate symmetry, magnitudes go back to the warp's scratch, and the mel
//     projection reads only each filter's non-zero bins (slaney triangles: ~1100 of the 41k basis entries);
//   * log-mels are parked in shared memory and leave as 128-byte rows of 32 consecutive frames.
// Algorithmic traffic: 4 B/sample in + 4*80 B/frame out (~5.25 B/sample at hop 256); the kernel is bound by fp32 FFT
// arithmetic (~25 kFLOP per frame), not by HBM - see DESIGN.md for the measured figures.
#include "common.cuh"

namespace {

constexpr int NFFT = 1024;
constexpr int NBIN = NFFT / 2 + 1;
constexpr int FRB = 32;       // frames per CTA
constexpr int NWARP = 8;
constexpr int NTH = NWARP * 32;
constexpr int SCR = 33 * 32 * 2;   // per-warp scratch floats: padded 32 x 32 complex transpose / X[1024] / magnitudes
constexpr int PACKED = 2048;       // packed non-zero filter weights (slaney triangles: <= 2 per bin + ends, ~1100 for 80 mels)
constexpr int MAXMEL = 128;

__device__ __forceinline__ constexpr int brev5(int i) {
  return ((i & 1) << 4) | ((i & 2) << 2) | (i & 4) | ((i & 8) >> 2) | ((i & 16) >> 4);
}

// forward 32-point DFT in registers, radix-2 decimation in frequency: output bin brev5(i) ends up in slot i
__device__ __forceinline__ void fft32(float (&re)[32], float (&im)[32]) {
  constexpr float C[16] = {1.000000000f, 0.980785280f, 0.923879533f, 0.831469612f, 0.707106781f, 0.555570233f,
                           0.382683432f, 0.195090322f, 0.000000000f, -0.195090322f, -0.382683432f, -0.555570233f,
                           -0.707106781f, -0.831469612f, -0.923879533f, -0.980785280f};
  constexpr float S[16] = {0.000000000f, 0.195090322f, 0.382683432f, 0.555570233f, 0.707106781f, 0.831469612f,
                           0.923879533f, 0.980785280f, 1.000000000f, 0.980785280f, 0.923879533f, 0.831469612f,
                           0.707106781f, 0.555570233f, 0.382683432f, 0.195090322f};
#pragma unroll
  for (int len = 32; len >= 2; len >>= 1) {
    const int half = len >> 1, tstep = 32 / len;
#pragma unroll
    for (int blk = 0; blk < 32; blk += len) {
#pragma unroll
      for (int j = 0; j < half; ++j) {
        const int i0 = blk + j, i1 = i0 + half;
        const float ar = re[i0], ai = im[i0], br = re[i1], bi = im[i1];
        re[i0] = ar + br;
        im[i0] = ai + bi;
        const float dr = ar - br, di = ai - bi;
        const int t = j * tstep;                 // (dr + i di) * exp(-2 pi i t / 32) = (dr + i di)(c - i s)
        if (t == 0) {
          re[i1] = dr; im[i1] = di;
        } else if (t == 8) {
          re[i1] = di; im[i1] = -dr;
        } else {
          re[i1] = dr * C[t] + di * S[t];
          im[i1] = di * C[t] - dr * S[t];
        }
      }
    }
  }
}

__global__ void __launch_bounds__(NTH) mel_kernel(const float* __restrict__ wave, int n, int hop, int pad,
                                                  const float* __restrict__ basis, const int32_t* __restrict__ row_range,
                                                  int n_mels, float* __restrict__ out, int frames) {
  extern __shared__ float sm[];
  float* win = sm;                         // [1024] periodic hann
  float* twr = win + NFFT;                 // [32][32] cos(2 pi n2 k1 / 1024), indexed [k1][n2]
  float* twi = twr + NFFT;                 // [32][32] -sin
  float* outs = twi + NFFT;                // [n_mels][FRB] log-mels of this CTA
  float* scr = outs + n_mels * FRB;        // [NWARP][SCR]
  float* packed = scr + NWARP * SCR;       // [PACKED] non-zero weights, row after row
  int* roff = reinterpret_cast<int*>(packed + PACKED);   // [MAXMEL + 1] start of row m in `packed`
  int* rlo = roff + MAXMEL + 1;            // [MAXMEL] first non-zero bin of row m
  float* stage = reinterpret_cast<float*>(rlo + MAXMEL);   // [(FRB-1)*hop + 1024] samples

  const int b = blockIdx.y;
  const int f0 = blockIdx.x * FRB;
  const int nf = min(FRB, frames - f0);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float* wb = wave + (int64_t)b * n;

  for (int i = tid; i < NFFT; i += NTH) {
    const float s = sinpif((float)i / (float)NFFT);   // torch.hann_window(1024): 0.5 - 0.5 cos(2 pi i / N) = sin^2(pi i / N)
    win[i] = s * s;
    float sn, cs;
    sincospif(2.0f * (float)((i >> 5) * (i & 31)) / (float)NFFT, &sn, &cs);   // i = k1 * 32 + n2
    twr[i] = cs;
    twi[i] = -sn;
  }
  // pack the non-zero filter weights (a lane-per-filter dot product over global rows touches 32 cache lines per
  // load; from shared memory it is a plain gather)
  if (tid == 0) {
    int acc = 0;
    for (int m = 0; m < n_mels; ++m) {
      roff[m] = acc;
      rlo[m] = row_range[2 * m];
      acc += row_range[2 * m + 1] - row_range[2 * m];
    }
    roff[n_mels] = acc;
  }
  __syncthreads();
  const bool use_packed = roff[n_mels] <= PACKED;
  if (use_packed) {
    for (int m = wid; m < n_mels; m += NWARP) {
      const int len = roff[m + 1] - roff[m];
      for (int i = lane; i < len; i += 32) packed[roff[m] + i] = basis[(int64_t)m * NBIN + rlo[m] + i];
    }
  }
  const int span = (nf - 1) * hop + NFFT;
  const int start = f0 * hop - pad;        // index into the un-padded signal
  for (int i = tid; i < span; i += NTH) {
    int j = start + i;
    if (j < 0) j = -j;                      // reflect (no edge repeat), F.pad(mode='reflect')
    if (j >= n) j = 2 * (n - 1) - j;
    stage[i] = wb[j];
  }
  __syncthreads();

  float* my = scr + wid * SCR;
  for (int pr = wid; 2 * pr < nf; pr += NWARP) {
    const int fa = 2 * pr, fb = fa + 1;
    const bool has_b = fb < nf;
    const float* xa = stage + fa * hop;
    const float* xb = stage + (has_b ? fb : fa) * hop;
    float re[32], im[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) {          // n = 32 r + lane
      const float w = win[32 * r + lane];
      re[r] = xa[32 * r + lane] * w;
      im[r] = has_b ? xb[32 * r + lane] * w : 0.f;
    }
    fft32(re, im);                          // slot i: k1 = brev5(i), this lane's n2 = lane
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int k1 = brev5(i);
      const float c = twr[k1 * 32 + lane], s = twi[k1 * 32 + lane];
      const float yr = re[i] * c - im[i] * s, yi = re[i] * s + im[i] * c;
      my[k1 * 33 + lane] = yr;              // transpose buffer T[k1][n2], row pitch 33
      my[33 * 32 + k1 * 33 + lane] = yi;
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 32; ++r) {          // lane = k1, r = n2
      re[r] = my[lane * 33 + r];
      im[r] = my[33 * 32 + lane * 33 + r];
    }
    __syncwarp();
    fft32(re, im);                          // slot i: k2 = brev5(i) -> bin k = lane + 32 k2
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      my[lane + 32 * brev5(i)] = re[i];     // X[k], re at [0, 1024), im at [1024, 2048)
      my[NFFT + lane + 32 * brev5(i)] = im[i];
    }
    __syncwarp();
    // split the two real spectra: A[k] = (Z[k] + conj Z[N-k]) / 2, B[k] = (Z[k] - conj Z[N-k]) / 2i
    float ma[17], mb[17];
#pragma unroll
    for (int j = 0; j < 17; ++j) {
      const int k = lane + 32 * j;
      ma[j] = mb[j] = 0.f;
      if (k <= NFFT / 2) {
        const int kn = (NFFT - k) & (NFFT - 1);
        const float zr = my[k], zi = my[NFFT + k], wr = my[kn], wi = my[NFFT + kn];
        const float ar = 0.5f * (zr + wr), ai = 0.5f * (zi - wi);
        const float br = 0.5f * (zi + wi), bi = -0.5f * (zr - wr);
        ma[j] = sqrtf(ar * ar + ai * ai + 1e-9f);
        mb[j] = sqrtf(br * br + bi * bi + 1e-9f);
      }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 17; ++j) {
      const int k = lane + 32 * j;
      if (k <= NFFT / 2) {
        my[k] = ma[j];
        my[NFFT + k] = mb[j];
      }
    }
    __syncwarp();
    // mel projection over each filter's non-zero bins
    for (int m = lane; m < n_mels; m += 32) {
      const int lo = rlo[m], len = roff[m + 1] - roff[m];
      const float* wrow = use_packed ? packed + roff[m] : basis + (int64_t)m * NBIN + lo;
      float accA = 0.f, accB = 0.f;
      for (int i = 0; i < len; ++i) {
        const float w = wrow[i];
        accA = fmaf(w, my[lo + i], accA);
        accB = fmaf(w, my[NFFT + lo + i], accB);
      }
      outs[m * FRB + fa] = logf(fmaxf(accA, 1e-5f));
      if (has_b) outs[m * FRB + fb] = logf(fmaxf(accB, 1e-5f));
    }
    __syncwarp();
  }
  __syncthreads();
  for (int i = tid; i < n_mels * FRB; i += NTH) {
    const int m = i / FRB, f = i - m * FRB;
    if (f < nf) out[((int64_t)b * n_mels + m) * frames + f0 + f] = outs[i];
  }
}

// [lo, hi) of the non-zero bins of every mel filter (filters are contiguous triangles; zero rows give lo = hi = 0)
__global__ void mel_row_range_kernel(const float* __restrict__ basis, int n_mels, int32_t* __restrict__ row_range) {
  const int m = blockIdx.x;
  const int lane = threadIdx.x;
  int lo = NBIN, hi = 0;
  for (int k = lane; k < NBIN; k += 32) {
    if (basis[(int64_t)m * NBIN + k] != 0.f) {
      lo = min(lo, k);
      hi = max(hi, k + 1);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) {
    row_range[2 * m] = hi > 0 ? lo : 0;
    row_range[2 * m + 1] = hi;
  }
}

}  // namespace

extern "C" size_t sib_mel_workspace_bytes(int n_mels) { return (size_t)n_mels * 2 * sizeof(int32_t); }

extern "C" int sib_mel_spectrogram_f32(const float* wave, int batch, int n, int hop, int pad, const float* mel_basis,
                                       int n_mels, float* out, int frames, void* workspace, sib_stream_t stream) {
  SIB_REQUIRE(wave && mel_basis && out && workspace && batch > 0 && batch <= 65535 && n > 0 && hop > 0 && pad >= 0 && n_mels > 0,
              "sib_mel_spectrogram_f32: bad argument");
  SIB_REQUIRE(pad < n, "sib_mel_spectrogram_f32: reflect pad %d must be < n=%d", pad, n);
  SIB_REQUIRE(n + 2 * pad >= NFFT, "sib_mel_spectrogram_f32: signal too short for n_fft=1024");
  const int expect = 1 + (n + 2 * pad - NFFT) / hop;
  SIB_REQUIRE(frames == expect, "sib_mel_spectrogram_f32: frames=%d but shape implies %d", frames, expect);
  SIB_REQUIRE(hop <= 1024, "sib_mel_spectrogram_f32: hop=%d > 1024 unsupported", hop);
  SIB_REQUIRE(n_mels <= 128, "sib_mel_spectrogram_f32: n_mels=%d > 128 unsupported", n_mels);
  const size_t smem = (size_t)(3 * NFFT + n_mels * FRB + NWARP * SCR + PACKED + (2 * MAXMEL + 1) + (FRB - 1) * hop + NFFT) *
                      sizeof(float);
  SIB_REQUIRE(smem <= 227 * 1024, "sib_mel_spectrogram_f32: hop=%d needs %zu bytes of shared memory", hop, smem);
  cudaError_t e = cudaFuncSetAttribute(mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    sib::set_error("sib_mel_spectrogram_f32: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return SIB_ERR_CUDA;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int32_t* rr = static_cast<int32_t*>(workspace);
  mel_row_range_kernel<<<n_mels, 32, 0, s>>>(mel_basis, n_mels, rr);
  dim3 grid(sib::ceil_div(frames, FRB), batch);
  mel_kernel<<<grid, NTH, smem, s>>>(wave, n, hop, pad, mel_basis, rr, n_mels, out, frames);
  SIB_CHECK_LAUNCH("sib_mel_spectrogram_f32");
  sib::count_launch();
  return SIB_OK;
}
