// bf16 implicit-GEMM convolution / linear layer on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
// Same contract as sib_conv1d_f32 (rows = time steps, reduction over (tap, input channel)); the A
// operand of K-block (tap-block tb, channel-chunk cc) is fetched by ONE TMA box straight from the
// frame-major activation tensor - no im2col buffer:
//   * the tap shift is a row coordinate of a 3-D map (channel, row, batch): zero padding = TMA OOB fill, batch
//     boundaries are respected because the batch index is its own coordinate;
//   * c_in/groups % 64 == 0 (HuBERT linears/convs, HiFi-GAN C >= 64): K-block = 1 tap x 64 channels, 128-byte
//     swizzle; stride-s convs (HuBERT conv1-6) read the same tensor viewed as [B, T/s, s*C];
//   * c_in/groups in {32, 16, 48, 80, ...} (HiFi-GAN C = 32/16, pos-conv 48/group, conv_pre 80): K-block =
//     TB taps x CC channels (CC = 32 or 16, CC*TB = 64): TB sub-tiles of 128 x CC with 64- / 32-byte swizzle,
//     one TMA box and CC/16 MMAs each, so a pipeline stage always carries 64 K-elements.
// Pipeline: warp 0 = TMA producer, warp 1 = TMEM owner + single-thread tcgen05.mma issuer, warps 2-5 =
// epilogue (tcgen05.ld -> bias / residual / accumulate / activation -> bf16 stores).  Accumulator
// 128 x BN fp32 lives in TMEM.  Several CTAs per SM overlap one tile's epilogue with another's mainloop.
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int BM = 128;      // time rows per CTA == UMMA M == TMEM lanes
constexpr int BK = 64;       // bf16 elements per K-block == one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;

struct TcArgs {
  // epilogue
  const float* bias;
  const __nv_bfloat16* res;
  __nv_bfloat16* y;
  __nv_bfloat16* y2;
  int64_t y_bs, r_bs;
  int y_rs, r_rs;
  int t_out, cout_g, groups;
  int post_act, accumulate, res_after_act;
  float post_slope, out_scale, act2_slope;
  // mainloop
  int n_chunks, n_tapblocks, tb, cc, n_taps, cin_g;
  int a_sub_bytes, b_sub_bytes;  // bytes of one (tap) sub-tile of A / B inside a stage
  uint32_t desc_hi;              // SBO / version / swizzle bits of the smem matrix descriptor (bits 32..63)
  int tap_row[SIB_MAX_TAPS];     // row coordinate delta per tap
  int tap_ch[SIB_MAX_TAPS];      // channel coordinate delta per tap (stride-s view)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// K-major swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm_100):
//   bits [0,14) start >> 4 | [16,30) LBO >> 4 (=1, unused for swizzled K-major) | [32,46) SBO >> 4 (8 rows of the
//   swizzle width) | [46,48) version = 1 | [61,64) layout: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B.
__host__ __device__ constexpr uint32_t make_desc_hi(int row_bytes) {
  return (uint32_t)((8 * row_bytes) >> 4) | (1u << 14) | ((row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u)) << 29);
}
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t desc_hi) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)desc_hi << 32);
}

// kind::f16 instruction descriptor: D=f32 (bit 4), A=B=bf16 (bits 7, 10), both K-major, N>>3 at 17, M>>4 at 24.
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int BN>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 128 ? 3 : 4);
  static constexpr int BAR_BYTES = 256;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // + alignment slack
};

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS)
conv1d_bf16_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const __grid_constant__ TcArgs p) {
  using L = SmemLayout<BN>;
  constexpr int STAGES = L::STAGES;
  constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * L::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int b = blockIdx.z / p.groups, g = blockIdx.z % p.groups;
  const int iters = p.n_chunks * p.n_tapblocks;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        const int cc = it / p.n_tapblocks, tb = it - cc * p.n_tapblocks;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* a_dst = smem + stage * L::STAGE_BYTES;
        uint8_t* b_dst = a_dst + L::A_BYTES;
        const int nsub = min(p.tb, p.n_taps - tb * p.tb);
        mbar_expect_tx(&full_bar[stage], (uint32_t)nsub * (uint32_t)(p.a_sub_bytes + p.b_sub_bytes));
        for (int sidx = 0; sidx < nsub; ++sidx) {
          const int j = tb * p.tb + sidx;
          tma_load_3d(a_dst + sidx * p.a_sub_bytes, &map_a, &full_bar[stage], g * p.cin_g + cc * p.cc + p.tap_ch[j],
                      t0 + p.tap_row[j], b);
          tma_load_2d(b_dst + sidx * p.b_sub_bytes, &map_b, &full_bar[stage], (cc * p.n_taps + j) * p.cc,
                      g * p.cout_g + n0);
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&full_bar[stage], phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_addr = smem_u32(smem + stage * L::STAGE_BYTES);
        const uint32_t b_addr = a_addr + L::A_BYTES;
        const int tb = it % p.n_tapblocks;
        const int nsub = min(p.tb, p.n_taps - tb * p.tb);
        const int ksteps = p.cc / UMMA_K;
        for (int sidx = 0; sidx < nsub; ++sidx) {
          const uint64_t adesc = make_smem_desc(a_addr + sidx * p.a_sub_bytes, p.desc_hi);
          const uint64_t bdesc = make_smem_desc(b_addr + sidx * p.b_sub_bytes, p.desc_hi);
          for (int k = 0; k < ksteps; ++k) {
            // advancing K by 16 bf16 = 32 bytes inside the swizzle atom: +2 in the (addr >> 4) field
            umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (it | sidx | k) ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // ===================== epilogue: 4 warps, warp q owns TMEM lanes [32q, 32q+32) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int t = t0 + row;
    mbar_wait(tmem_full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const bool row_ok = t < p.t_out;
    const int64_t y_off = (int64_t)b * p.y_bs + (int64_t)t * p.y_rs + (int64_t)g * p.cout_g;
    const int64_t r_off = (int64_t)b * p.r_bs + (int64_t)t * p.r_rs + (int64_t)g * p.cout_g;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);  // warp-collective: no divergence here
      if (!row_ok) continue;
#pragma unroll
      for (int c8 = 0; c8 < 32; c8 += 8) {
        const int n = n0 + c0 + c8;
        if (n >= p.cout_g) break;
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[c8 + i]);
        if (p.bias) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (n + i < p.cout_g) f[i] += __ldg(p.bias + g * p.cout_g + n + i);
        }
        const bool full8 = (n + 8 <= p.cout_g);
        float r[8];
        bool have_r = false;
        if (p.res) {
          have_r = true;
          if (full8) {
            const uint4 rv = *reinterpret_cast<const uint4*>(p.res + r_off + n);
            const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 ff = __bfloat1622float2(r2[i]);
              r[2 * i] = ff.x;
              r[2 * i + 1] = ff.y;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = (n + i < p.cout_g) ? __bfloat162float(p.res[r_off + n + i]) : 0.f;
          }
          if (!p.res_after_act) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] += r[i];
          }
        }
        if (p.accumulate) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (n + i < p.cout_g) f[i] += __bfloat162float(p.y[y_off + n + i]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          f[i] = sib::apply_act(f[i] * p.out_scale, p.post_act, p.post_slope);
          if (have_r && p.res_after_act) f[i] += r[i];
        }
        if (full8) {
          uint4 o;
          __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
          for (int i = 0; i < 4; ++i) o2[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
          *reinterpret_cast<uint4*>(p.y + y_off + n) = o;
          if (p.y2) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float a0 = f[2 * i] > 0.f ? f[2 * i] : f[2 * i] * p.act2_slope;
              const float a1 = f[2 * i + 1] > 0.f ? f[2 * i + 1] : f[2 * i + 1] * p.act2_slope;
              o2[i] = __floats2bfloat162_rn(a0, a1);
            }
            *reinterpret_cast<uint4*>(p.y2 + y_off + n) = o;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (n + i >= p.cout_g) break;
            p.y[y_off + n + i] = __float2bfloat16_rn(f[i]);
            if (p.y2) p.y2[y_off + n + i] = __float2bfloat16_rn(f[i] > 0.f ? f[i] : f[i] * p.act2_slope);
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      ptr = nullptr;
    return reinterpret_cast<EncodeTiledFn>(ptr);
  }();
  return fn;
}

int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, CUtensorMapSwizzle swz, const char* what) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    sib::set_error("sib_conv1d_bf16: cuTensorMapEncodeTiled unavailable");
    return SIB_ERR_CUDA;
  }
  cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims,
                   strides_bytes + 1, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    sib::set_error("sib_conv1d_bf16: cuTensorMapEncodeTiled(%s) failed with CUresult %d (dims %llu,%llu,%llu stride1 %llu)",
                   what, (int)r, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                   (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 1 ? strides_bytes[1] : 0));
    return SIB_ERR_CUDA;
  }
  return SIB_OK;
}

template <int BN>
int launch_tc(const CUtensorMap& ma, const CUtensorMap& mb, const TcArgs& a, dim3 grid, cudaStream_t s) {
  using L = SmemLayout<BN>;
  static bool attr_set = false;  // benign race: idempotent
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv1d_bf16_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) {
      sib::set_error("sib_conv1d_bf16: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return SIB_ERR_CUDA;
    }
    attr_set = true;
  }
  conv1d_bf16_tc_kernel<BN><<<grid, NUM_THREADS, L::TOTAL, s>>>(ma, mb, a);
  return SIB_OK;
}

}  // namespace

extern "C" int sib_conv1d_bf16_kblock(int c_in_per_group, int* cc, int* tb) {
  if (c_in_per_group % 64 == 0) { *cc = 64; *tb = 1; return SIB_OK; }
  if (c_in_per_group % 32 == 0) { *cc = 32; *tb = 2; return SIB_OK; }
  if (c_in_per_group % 16 == 0) { *cc = 16; *tb = 4; return SIB_OK; }
  sib::set_error("sib_conv1d_bf16: c_in/groups=%d must be a multiple of 16", c_in_per_group);
  return SIB_ERR_UNSUPPORTED;
}

extern "C" int sib_conv1d_bf16(const sib_conv_desc* d, const void* x, const void* w, const float* bias,
                               const void* residual, void* y, void* y_act, sib_stream_t stream) {
  SIB_REQUIRE(d && x && w && y, "sib_conv1d_bf16: null argument");
  SIB_REQUIRE(d->batch > 0 && d->t_in > 0 && d->t_out > 0 && d->c_in > 0 && d->c_out > 0, "sib_conv1d_bf16: empty shape");
  SIB_REQUIRE(d->groups > 0 && d->c_in % d->groups == 0 && d->c_out % d->groups == 0,
              "sib_conv1d_bf16: groups=%d must divide c_in=%d and c_out=%d", d->groups, d->c_in, d->c_out);
  SIB_REQUIRE(d->n_taps > 0 && d->n_taps <= SIB_MAX_TAPS, "sib_conv1d_bf16: n_taps=%d out of range", d->n_taps);
  SIB_REQUIRE(d->pre_act == SIB_ACT_NONE, "sib_conv1d_bf16: pre-activation is not available on the TMA path; "
                                          "have the producer write the activated tensor (y_act)");
  SIB_REQUIRE((int64_t)d->batch * d->groups <= 65535, "sib_conv1d_bf16: batch*groups too large for grid.z");
  const int cin_g = d->c_in / d->groups, cout_g = d->c_out / d->groups;
  int cc, tb;
  if (int rc = sib_conv1d_bf16_kblock(cin_g, &cc, &tb)) return rc;
  SIB_REQUIRE(cout_g % 8 == 0, "sib_conv1d_bf16: c_out/groups=%d must be a multiple of 8", cout_g);
  SIB_REQUIRE(d->x_row_stride % 8 == 0 && d->x_batch_stride % 8 == 0 && d->y_row_stride % 8 == 0 &&
                  d->y_batch_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(y) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0,
              "sib_conv1d_bf16: x / y / w must be 16-byte aligned with strides that are multiples of 8 elements");
  SIB_REQUIRE(!residual || ((reinterpret_cast<uintptr_t>(residual) & 15) == 0 && d->r_row_stride % 8 == 0 &&
                            d->r_batch_stride % 8 == 0),
              "sib_conv1d_bf16: residual must be 16-byte aligned with strides that are multiples of 8 elements");
  SIB_REQUIRE(!y_act || (reinterpret_cast<uintptr_t>(y_act) & 15) == 0, "sib_conv1d_bf16: y_act must be 16-byte aligned");

  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.bias = bias;
  a.res = static_cast<const __nv_bfloat16*>(residual);
  a.y = static_cast<__nv_bfloat16*>(y);
  a.y2 = static_cast<__nv_bfloat16*>(y_act);
  a.y_bs = d->y_batch_stride; a.r_bs = d->r_batch_stride; a.y_rs = d->y_row_stride; a.r_rs = d->r_row_stride;
  a.t_out = d->t_out; a.cout_g = cout_g; a.groups = d->groups;
  a.post_act = d->post_act; a.accumulate = d->accumulate; a.res_after_act = d->res_after_act;
  a.post_slope = d->post_slope; a.out_scale = d->out_scale; a.act2_slope = d->act2_slope;
  a.cc = cc; a.tb = tb; a.cin_g = cin_g;
  a.n_chunks = cin_g / cc;
  a.n_tapblocks = (d->n_taps + tb - 1) / tb;
  a.n_taps = d->n_taps;
  const int row_bytes = cc * 2;  // one K-row of a sub-tile == the swizzle width (128 / 64 / 32 bytes)
  a.desc_hi = make_desc_hi(row_bytes);
  const CUtensorMapSwizzle swz = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  const int bn = cout_g >= 128 ? 128 : (cout_g > 32 ? 64 : 32);
  a.a_sub_bytes = BM * row_bytes;
  a.b_sub_bytes = bn * row_bytes;

  const int s = d->stride;
  SIB_REQUIRE(s >= 1, "sib_conv1d_bf16: stride must be positive");
  if (s > 1) {
    SIB_REQUIRE(d->groups == 1 && d->x_row_stride == d->c_in && tb == 1,
                "sib_conv1d_bf16: stride>1 needs groups=1, dense rows and c_in %% 64 == 0");
    int max_off = 0;
    for (int j = 0; j < d->n_taps; ++j) {
      SIB_REQUIRE(d->tap_offset[j] >= 0, "sib_conv1d_bf16: stride>1 supports valid (un-padded) convolutions only");
      if (d->tap_offset[j] > max_off) max_off = d->tap_offset[j];
    }
    SIB_REQUIRE((int64_t)(d->t_out - 1) * s + max_off <= d->t_in - 1,
                "sib_conv1d_bf16: stride>1 output would read past t_in (valid convolution required)");
  }
  for (int j = 0; j < d->n_taps; ++j) {
    const int off = d->tap_offset[j];
    const int qd = (off >= 0) ? off / s : -((-off + s - 1) / s);  // floor division
    a.tap_row[j] = qd;
    a.tap_ch[j] = (off - qd * s) * d->c_in;
  }
  CUtensorMap map_a, map_b;
  {
    // A viewed as [batch][ceil(t_in/s)][s*c_in]; for s>1 the last (partial) row may extend past t_in: the caller
    // guarantees those bytes are readable (see header) - they only feed masked outputs.
    const cuuint64_t dims[3] = {(cuuint64_t)d->c_in * s, (cuuint64_t)((d->t_in + s - 1) / s), (cuuint64_t)d->batch};
    const cuuint64_t strides[3] = {2, (cuuint64_t)d->x_row_stride * s * 2, (cuuint64_t)d->x_batch_stride * 2};
    const cuuint32_t box[3] = {(cuuint32_t)cc, BM, 1};
    if (int rc = encode_map(&map_a, x, 3, dims, strides, box, swz, "A")) return rc;
  }
  {
    const cuuint64_t ktot = (cuuint64_t)a.n_chunks * d->n_taps * cc;
    const cuuint64_t dims[2] = {ktot, (cuuint64_t)d->c_out};
    const cuuint64_t strides[2] = {2, ktot * 2};
    const cuuint32_t box[2] = {(cuuint32_t)cc, (cuuint32_t)bn};
    if (int rc = encode_map(&map_b, w, 2, dims, strides, box, swz, "B")) return rc;
  }
  dim3 grid(sib::ceil_div(d->t_out, BM), sib::ceil_div(cout_g, bn), d->batch * d->groups);
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  int rc;
  if (bn == 128) rc = launch_tc<128>(map_a, map_b, a, grid, cs);
  else if (bn == 64) rc = launch_tc<64>(map_a, map_b, a, grid, cs);
  else rc = launch_tc<32>(map_a, map_b, a, grid, cs);
  if (rc) return rc;
  SIB_CHECK_LAUNCH("sib_conv1d_bf16");
  return SIB_OK;
}
