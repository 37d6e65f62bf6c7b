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
// Persistent kernel, one CTA per SM looping over output tiles (128 time rows x bn channels):
//   warp 0      TMA producer: shared-memory ring of STAGES K-blocks, runs ahead across tile boundaries
//   warp 1      TMEM owner + single-thread tcgen05.mma issuer; TWO accumulators (2 x bn TMEM columns) so the
//               epilogue of tile i overlaps the mainloop of tile i+1
//   warps 2-5   epilogue, warp q owns TMEM lanes / tile rows [32q, 32q+32): residual and the accumulate input are
//               prefetched by TMA into a swizzled staging slab, tcgen05.ld -> bias / residual / accumulate /
//               activation -> bf16 into the slab -> TMA store (coalesced, rows >= t_out clipped by the TMA unit).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int BM = 128;      // time rows per CTA == UMMA M == TMEM lanes
constexpr int BK = 64;       // bf16 elements per K-block == one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;                 // two per TMEM lane quarter: latency hiding in the epilogue
constexpr int EPI_WARP0 = 4;                     // warps 0-3: A producer, MMA issuer, B producer, spare
constexpr int NUM_THREADS = 32 * (EPI_WARP0 + NUM_EPI_WARPS);

struct TcArgs {
  // epilogue
  const float* bias;
  int has_res, has_y2;
  int t_out, cout_g, groups, batch;
  int post_act, accumulate, res_after_act;
  float post_slope, out_scale, act2_slope;
  int bn;            // tile width in output channels (multiple of 16, <= 128) == UMMA N
  int cw, cw_shift;  // staging box width in channels (64 / 32 / 16) and its log2
  int tiles_m, tiles_n, total_tiles;
  int acc_stride;    // TMEM columns between the two accumulators
  int stages, stage_bytes;        // mode 0: combined A+B ring
  // mode 1 ("halo"): the A ring holds one (128 + span) x cc halo tile per channel chunk; every tap is a row-shifted
  // view of it.  B (weights) either streams through its own ring (tg taps per stage) or stays resident.
  int mode, rows_h, tap_step, off0;
  int a_stages, a_stage_bytes, b_stages, b_stage_bytes, b_tap_bytes, tg, b_resident, b_region_bytes;
  int ring_bytes;                 // bytes of all operand rings (slabs start here)
  int slab_depth;                 // epilogue staging ring depth D: residual prefetched D-1 tiles ahead
  // mainloop
  int n_chunks, n_tapblocks, tb, cc, n_taps, cin_g;
  int a_sub_bytes, b_sub_bytes;  // bytes of one (tap) sub-tile of A / B inside a stage
  uint32_t desc_hi;              // SBO / version / swizzle bits of the smem matrix descriptor (bits 32..63)
  uint32_t idesc;
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void unpack8(const uint4& raw, float (&f)[8]) {
  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h2[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 o;
  __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int i = 0; i < 4; ++i) o2[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return o;
}


constexpr int A_STAGE_BYTES = BM * BK * 2;  // mode 0: 16 KB of A per pipeline stage regardless of the sub-tile split
constexpr int MAX_A_STAGES = 8, MAX_B_STAGES = 16, MAX_SLAB_DEPTH = 4;

template <int POST_ACT>
__device__ __forceinline__ float act_t(float v, float slope) {
  if (POST_ACT == SIB_ACT_GELU) return sib::gelu_erf(v);
  if (POST_ACT == SIB_ACT_LRELU) return v > 0.f ? v : v * slope;
  if (POST_ACT == SIB_ACT_TANH) return tanhf(v);
  return v;
}

template <int POST_ACT>
__global__ void __launch_bounds__(NUM_THREADS, 2)
conv1d_bf16_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_y2,
                      const __grid_constant__ CUtensorMap map_r, const __grid_constant__ TcArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int slab_bytes = 32 * p.bn * 2;                       // one epilogue warp-pair's 32 rows x bn bf16
  uint8_t* b_region = smem + p.a_stages * p.a_stage_bytes;    // mode 1 only
  uint8_t* stage_y = smem + p.ring_bytes;                     // D x 4 slabs: output (and accumulate input)
  uint8_t* stage_r = stage_y + p.slab_depth * 4 * slab_bytes; // D x 4 slabs: residual input / second output
  uint64_t* a_full = reinterpret_cast<uint64_t*>(stage_r + p.slab_depth * 4 * slab_bytes);
  uint64_t* a_empty = a_full + MAX_A_STAGES;
  uint64_t* b_full = a_empty + MAX_A_STAGES;
  uint64_t* b_empty = b_full + MAX_B_STAGES;
  uint64_t* tmem_full_bar = b_empty + MAX_B_STAGES;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
  uint64_t* res_bar = tmem_empty_bar + 2;            // [4 pairs][MAX_SLAB_DEPTH]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_bar + 4 * MAX_SLAB_DEPTH);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y) : "memory");
    for (int s = 0; s < MAX_A_STAGES; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < MAX_B_STAGES; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], NUM_EPI_WARPS);  // one arrive per epilogue warp
    }
    for (int s = 0; s < 4 * MAX_SLAB_DEPTH; ++s) mbar_init(&res_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t tmem_cols = 2 * p.acc_stride;  // power of two >= 32 (host guarantees)
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  // tile -> (n fastest, m, batch*group): CTAs that run concurrently share the same A rows in L2
  auto decode = [&](int tile, int& t0, int& n0, int& b, int& g) {
    const int nt = tile % p.tiles_n;
    const int rest = tile / p.tiles_n;
    const int mt = rest % p.tiles_m;
    const int z = rest / p.tiles_m;
    t0 = mt * BM;
    n0 = nt * p.bn;
    b = z / p.groups;
    g = z - b * p.groups;
  };
  const int row_bytes_k = p.cc * 2;        // bytes of one K-row of an operand sub-tile (= its swizzle width)
  const int ksteps = p.cc / UMMA_K;

  if (warp == 0) {
    // ===================== A producer (mode 0: A and B of every K-block) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int t0, n0, b, g;
        decode(tile, t0, n0, b, g);
        if (p.mode == 1) {
          for (int cc = 0; cc < p.n_chunks; ++cc) {
            mbar_wait(&a_empty[stage], phase ^ 1);
            mbar_expect_tx(&a_full[stage], (uint32_t)(p.rows_h * row_bytes_k));
            tma_load_3d(smem + stage * p.a_stage_bytes, &map_a, &a_full[stage], g * p.cin_g + cc * p.cc, t0 + p.off0, b);
            if (++stage == p.a_stages) { stage = 0; phase ^= 1; }
          }
        } else {
          const int iters = p.n_chunks * p.n_tapblocks;
          for (int it = 0; it < iters; ++it) {
            const int cc = it / p.n_tapblocks, tb = it - cc * p.n_tapblocks;
            mbar_wait(&a_empty[stage], phase ^ 1);
            uint8_t* a_dst = smem + stage * p.stage_bytes;
            uint8_t* b_dst = a_dst + A_STAGE_BYTES;
            const int nsub = min(p.tb, p.n_taps - tb * p.tb);
            mbar_expect_tx(&a_full[stage], (uint32_t)nsub * (uint32_t)(p.a_sub_bytes + p.b_sub_bytes));
            for (int sidx = 0; sidx < nsub; ++sidx) {
              const int j = tb * p.tb + sidx;
              tma_load_3d(a_dst + sidx * p.a_sub_bytes, &map_a, &a_full[stage], g * p.cin_g + cc * p.cc + p.tap_ch[j],
                          t0 + p.tap_row[j], b);
              tma_load_3d(b_dst + sidx * p.b_sub_bytes, &map_b, &a_full[stage], 0, n0,
                          (g * p.n_chunks + cc) * p.n_taps + j);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== B producer (mode 1) =====================
    if (lane == 0 && p.mode == 1) {
      if (p.b_resident) {
        // every tile of this launch uses the same weights: load them once (groups == 1, tiles_n == 1)
        const int slabs = p.n_chunks * p.n_taps;
        mbar_expect_tx(&b_full[0], (uint32_t)(((slabs + p.tg - 1) / p.tg) * p.b_stage_bytes));
        for (int s0 = 0; s0 < slabs; s0 += p.tg)
          tma_load_3d(b_region + (s0 / p.tg) * p.b_stage_bytes, &map_b, &b_full[0], 0, 0, s0);
      } else {
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
          int t0, n0, b, g;
          decode(tile, t0, n0, b, g);
          for (int cc = 0; cc < p.n_chunks; ++cc) {
            for (int j0 = 0; j0 < p.n_taps; j0 += p.tg) {
              mbar_wait(&b_empty[stage], phase ^ 1);
              mbar_expect_tx(&b_full[stage], (uint32_t)p.b_stage_bytes);
              tma_load_3d(b_region + stage * p.b_stage_bytes, &map_b, &b_full[stage], 0, n0,
                          (g * p.n_chunks + cc) * p.n_taps + j0);
              if (++stage == p.b_stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      int stage = 0, bstage = 0;
      uint32_t phase = 0, bphase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      if (p.mode == 1 && p.b_resident) {
        mbar_wait(&b_full[0], 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * p.acc_stride);
        if (p.mode == 1) {
          // descriptor arithmetic is strength-reduced: this single thread paces every MMA of the SM
          const uint32_t b_base = smem_u32(b_region);
          const uint64_t a_tap_inc = (uint64_t)((p.tap_step * row_bytes_k) >> 4);   // next tap = shifted rows
          const uint64_t b_tap_inc = (uint64_t)(p.b_tap_bytes >> 4);
          uint64_t bdesc_res = make_smem_desc(b_base, p.desc_hi);                    // resident weights: walk the slabs
          uint32_t accum = 0;
          for (int cc = 0; cc < p.n_chunks; ++cc) {
            mbar_wait(&a_full[stage], phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint64_t adesc = make_smem_desc(smem_u32(smem + stage * p.a_stage_bytes), p.desc_hi);
            if (p.b_resident) {
              for (int j = 0; j < p.n_taps; ++j) {
                for (int k = 0; k < ksteps; ++k) {
                  umma_bf16(tmem_d, adesc + 2 * k, bdesc_res + 2 * k, p.idesc, accum);
                  accum = 1;
                }
                adesc += a_tap_inc;
                bdesc_res += b_tap_inc;
              }
            } else {
              int j = 0;
              while (j < p.n_taps) {
                mbar_wait(&b_full[bstage], bphase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint64_t bdesc = make_smem_desc(b_base + (uint32_t)(bstage * p.b_stage_bytes), p.desc_hi);
                const int jend = min(j + p.tg, p.n_taps);
                for (; j < jend; ++j) {
                  for (int k = 0; k < ksteps; ++k) {
                    umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, p.idesc, accum);
                    accum = 1;
                  }
                  adesc += a_tap_inc;
                  bdesc += b_tap_inc;
                }
                umma_commit(&b_empty[bstage]);
                if (++bstage == p.b_stages) { bstage = 0; bphase ^= 1; }
              }
            }
            umma_commit(&a_empty[stage]);
            if (++stage == p.a_stages) { stage = 0; phase ^= 1; }
          }
        } else {
          const int iters = p.n_chunks * p.n_tapblocks;
          for (int it = 0; it < iters; ++it) {
            mbar_wait(&a_full[stage], phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_addr = smem_u32(smem + stage * p.stage_bytes);
            const uint32_t b_addr = a_addr + A_STAGE_BYTES;
            const int tb = it % p.n_tapblocks;
            const int nsub = min(p.tb, p.n_taps - tb * p.tb);
            for (int sidx = 0; sidx < nsub; ++sidx) {
              const uint64_t adesc = make_smem_desc(a_addr + sidx * p.a_sub_bytes, p.desc_hi);
              const uint64_t bdesc = make_smem_desc(b_addr + sidx * p.b_sub_bytes, p.desc_hi);
              for (int k = 0; k < ksteps; ++k) {
                // advancing K by 16 bf16 = 32 bytes inside the swizzle atom: +2 in the (addr >> 4) field
                umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, p.idesc, (it | sidx | k) ? 1u : 0u);
              }
            }
            umma_commit(&a_empty[stage]);  // frees the smem slot when these MMAs retire
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
        umma_commit(&tmem_full_bar[acc]);  // accumulator complete
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue: warps (q, q+4) share TMEM lanes / tile rows [32q, 32q+32) ============
    // and split the 16-column chunks between them (even / odd); warp `half == 0` drives the TMA traffic.
    const int q = warp & 3;
    const int half = (warp - EPI_WARP0) >> 2;
    const int D = p.slab_depth;
    const int row_bytes = p.cw * 2;                  // 128 / 64 / 32: also the TMA swizzle width of the slab boxes
    const int box_bytes = 32 * row_bytes;
    const int chunks_per_row = row_bytes >> 4;       // 16-byte chunks per box row
    const int swz_shift = row_bytes == 128 ? 0 : (row_bytes == 64 ? 1 : 2);
    const uint32_t swz = ((uint32_t)lane >> swz_shift) & (uint32_t)(chunks_per_row - 1);
    const int nboxes = p.bn >> p.cw_shift;
    const bool prefetch = p.has_res || p.accumulate;
    const uint32_t pre_bytes = (uint32_t)((p.has_res ? 1 : 0) + (p.accumulate ? 1 : 0)) * (uint32_t)slab_bytes;
    const uint32_t pair_bar = 1 + q;                 // named barrier of the warp pair (64 threads)
    const bool leader = half == 0 && lane == 0;
    uint64_t* my_res_bar = res_bar + q * MAX_SLAB_DEPTH;
    // residual / accumulate-input prefetch of one tile into slab slot `slot` (leader lane only)
    auto issue_prefetch = [&](int tile, int slot) {
      int t0, n0, b, g;
      decode(tile, t0, n0, b, g);
      uint8_t* sy = stage_y + (slot * 4 + q) * slab_bytes;
      uint8_t* sr = stage_r + (slot * 4 + q) * slab_bytes;
      mbar_expect_tx(&my_res_bar[slot], pre_bytes);
      for (int bx = 0; bx < nboxes; ++bx) {
        if (p.has_res) tma_load_3d(sr + bx * box_bytes, &map_r, &my_res_bar[slot], g * p.cout_g + n0 + bx * p.cw, t0 + q * 32, b);
        if (p.accumulate) tma_load_3d(sy + bx * box_bytes, &map_y, &my_res_bar[slot], g * p.cout_g + n0 + bx * p.cw, t0 + q * 32, b);
      }
    };
    if (prefetch && leader) {
      const int ahead = D > 1 ? D - 1 : 1;
      int tile = blockIdx.x;
      for (int k = 0; k < ahead && tile < p.total_tiles; ++k, tile += gridDim.x) issue_prefetch(tile, k % D);
    }
    int acc = 0, slot = 0;
    uint32_t acc_phase = 0, res_phase_bits = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int t0, n0, b, g;
      decode(tile, t0, n0, b, g);
      const int r0 = t0 + q * 32;
      const int ch0 = g * p.cout_g + n0;
      uint8_t* slab_y = stage_y + (slot * 4 + q) * slab_bytes;
      uint8_t* slab_r = stage_r + (slot * 4 + q) * slab_bytes;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (prefetch) {
        mbar_wait(&my_res_bar[slot], (res_phase_bits >> slot) & 1u);
        res_phase_bits ^= 1u << slot;
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.acc_stride);
#pragma unroll 1
      for (int c0 = half * 16; c0 < p.bn; c0 += 32) {
        uint32_t v[16];
        tmem_ld16(taddr + (uint32_t)c0, v);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int col = c0 + 8 * h;
          const int bx = col >> p.cw_shift;
          const uint32_t chunk = (uint32_t)((col - (bx << p.cw_shift)) >> 3);
          const uint32_t off = (uint32_t)(bx * box_bytes + lane * row_bytes) + ((chunk ^ swz) << 4);
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[8 * h + i]);
          if (p.bias) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + ch0 + col));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + ch0 + col + 4));
            f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
            f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
          }
          float r[8];
          if (p.has_res) {
            unpack8(*reinterpret_cast<const uint4*>(slab_r + off), r);
            if (!p.res_after_act) {
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] += r[i];
            }
          }
          if (p.accumulate) {
            float o[8];
            unpack8(*reinterpret_cast<const uint4*>(slab_y + off), o);
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] += o[i];
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = act_t<POST_ACT>(f[i] * p.out_scale, p.post_slope);
          if (p.has_res && p.res_after_act) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] += r[i];
          }
          *reinterpret_cast<uint4*>(slab_y + off) = pack8(f);
          if (p.has_y2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = f[i] > 0.f ? f[i] : f[i] * p.act2_slope;
            *reinterpret_cast<uint4*>(slab_r + off) = pack8(f);
          }
        }
      }
      // accumulator drained: hand it back to the MMA warp
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
      // slab -> global through the async proxy, once both warps of the pair have written their columns
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      if (leader) {
        for (int bx = 0; bx < nboxes; ++bx) {
          tma_store_3d(&map_y, slab_y + bx * box_bytes, ch0 + bx * p.cw, r0, b);
          if (p.has_y2) tma_store_3d(&map_y2, slab_r + bx * box_bytes, ch0 + bx * p.cw, r0, b);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        // the slot written D-1 tiles from now is the one the PREVIOUS tile used (or this one when D == 1):
        // wait until its stores have left shared memory, then refill it with the residual of tile i + D - 1
        if (D > 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (prefetch) {
          const int ahead = D > 1 ? D - 1 : 1;
          const long long nt = (long long)tile + (long long)ahead * gridDim.x;
          if (nt < p.total_tiles) issue_prefetch((int)nt, D > 1 ? (slot + D - 1) % D : 0);
        }
      }
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      if (++slot == D) slot = 0;
    }
    if (half == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
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
  const int cin_g = d->c_in / d->groups, cout_g = d->c_out / d->groups;
  int cc, tb;
  if (int rc = sib_conv1d_bf16_kblock(cin_g, &cc, &tb)) return rc;
  SIB_REQUIRE(cout_g % 16 == 0, "sib_conv1d_bf16: c_out/groups=%d must be a multiple of 16", cout_g);
  auto al16 = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
  SIB_REQUIRE(d->x_row_stride % 8 == 0 && d->x_batch_stride % 8 == 0 && d->y_row_stride % 8 == 0 &&
                  d->y_batch_stride % 8 == 0 && al16(x) && al16(y) && al16(w),
              "sib_conv1d_bf16: x / y / w must be 16-byte aligned with strides that are multiples of 8 elements");
  SIB_REQUIRE(!residual || (al16(residual) && d->r_row_stride % 8 == 0 && d->r_batch_stride % 8 == 0),
              "sib_conv1d_bf16: residual must be 16-byte aligned with strides that are multiples of 8 elements");
  SIB_REQUIRE(!y_act || al16(y_act), "sib_conv1d_bf16: y_act must be 16-byte aligned (it shares y's strides)");
  SIB_REQUIRE(!bias || al16(bias), "sib_conv1d_bf16: bias must be 16-byte aligned");

  // tile width: the largest multiple of 16 that is <= 128 and divides c_out/groups (== UMMA N)
  int bn = 0;
  for (int cand = 128; cand >= 16; cand -= 16)
    if (cout_g % cand == 0) { bn = cand; break; }
  const int cw = bn % 64 == 0 ? 64 : (bn % 32 == 0 ? 32 : 16);

  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.bias = bias;
  a.has_res = residual != nullptr;
  a.has_y2 = y_act != nullptr;
  a.t_out = d->t_out; a.cout_g = cout_g; a.groups = d->groups; a.batch = d->batch;
  a.post_act = d->post_act; a.accumulate = d->accumulate; a.res_after_act = d->res_after_act;
  a.post_slope = d->post_slope; a.out_scale = d->out_scale; a.act2_slope = d->act2_slope;
  a.bn = bn; a.cw = cw; a.cw_shift = cw == 64 ? 6 : (cw == 32 ? 5 : 4);
  a.tiles_m = sib::ceil_div(d->t_out, BM);
  a.tiles_n = cout_g / bn;
  const int64_t total = (int64_t)a.tiles_m * a.tiles_n * d->batch * d->groups;
  SIB_REQUIRE(total < (1ll << 31), "sib_conv1d_bf16: too many tiles");
  a.total_tiles = (int)total;
  a.acc_stride = bn <= 16 ? 16 : (bn <= 32 ? 32 : (bn <= 64 ? 64 : 128));
  a.cc = cc; a.tb = tb; a.cin_g = cin_g;
  a.n_chunks = cin_g / cc;
  a.n_tapblocks = (d->n_taps + tb - 1) / tb;
  a.n_taps = d->n_taps;
  const int row_bytes = cc * 2;  // one K-row of a sub-tile == the swizzle width (128 / 64 / 32 bytes)
  a.desc_hi = make_desc_hi(row_bytes);
  a.idesc = make_idesc(bn);
  auto swz_of = [](int rb) {
    return rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  };
  a.a_sub_bytes = BM * row_bytes;
  a.b_sub_bytes = bn * row_bytes;
  a.b_tap_bytes = bn * row_bytes;
  // epilogue staging ring: deep enough to hide the residual TMA latency behind D-1 tiles of work; layers whose
  // tiles carry a long K loop and no residual keep a single slab and give the memory to the operand rings
  const bool pre = a.has_res || a.accumulate;
  const int kblocks = a.n_chunks * d->n_taps * cc / BK;
  a.slab_depth = pre ? (bn <= 64 ? 3 : 2) : (bn <= 64 ? 2 : (kblocks >= 8 ? 1 : 2));
  const int slab_total = a.slab_depth * 2 * 4 * 32 * bn * 2;  // D x (y + r) staging, 4 epilogue warp pairs
  // narrow layers (bn <= 32: 8 KB output tiles) are bound by the per-tile latency chain of one CTA, not by any
  // throughput: run two persistent CTAs per SM on half the shared memory each
  const int ctas_per_sm = bn <= 32 ? 2 : 1;
  const int smem_budget = ctas_per_sm == 2 ? 112 * 1024 : 227 * 1024;
  const int avail = smem_budget - 2048 - slab_total - 1024;

  // ---- mode selection: "halo" mode whenever all taps are row shifts of one (128 + span)-row tile
  a.mode = 0;
  a.tg = 1;
  int step = 1;
  bool uniform = d->stride == 1;
  if (uniform && d->n_taps > 1) {
    step = d->tap_offset[1] - d->tap_offset[0];
    uniform = step >= 1;
    for (int j = 2; uniform && j < d->n_taps; ++j) uniform = (d->tap_offset[j] - d->tap_offset[j - 1]) == step;
  }
  const int rows_h = BM + (d->n_taps - 1) * step;
  static const bool force_mode0 = getenv("SIB_TC_FORCE_PER_TAP") != nullptr;  // A/B switch for profiling
  // (plain GEMMs, n_taps == 1, have nothing to share between taps: they keep the combined A+B ring of mode 0)
  if (uniform && d->n_taps > 1 && rows_h <= 256 && !force_mode0) {
    a.rows_h = rows_h; a.tap_step = step; a.off0 = d->tap_offset[0];
    a.a_stage_bytes = (rows_h * row_bytes + 1023) / 1024 * 1024;
    int tg = 16384 / a.b_tap_bytes;
    if (tg < 1) tg = 1;
    if (tg > d->n_taps) tg = d->n_taps;
    const int slabs = a.n_chunks * d->n_taps;
    // resident weights: split the slabs evenly over as few TMA boxes as possible (box depth <= 16 KB worth of taps)
    const int rloads = (slabs + tg - 1) / tg;
    const int rtg = (slabs + rloads - 1) / rloads;
    const int resident_bytes = rloads * rtg * a.b_tap_bytes;
    if (d->groups == 1 && a.tiles_n == 1 && resident_bytes <= 120 * 1024 && resident_bytes + 2 * a.a_stage_bytes <= avail) {
      a.mode = 1; a.b_resident = 1; a.tg = rtg; a.b_stage_bytes = rtg * a.b_tap_bytes; a.b_region_bytes = resident_bytes;
      a.b_stages = 1;
      a.a_stages = (avail - resident_bytes) / a.a_stage_bytes;
      if (a.a_stages > MAX_A_STAGES) a.a_stages = MAX_A_STAGES;
    } else {
      a.b_stage_bytes = tg * a.b_tap_bytes;
      // split the operand budget about evenly between the A ring and the B ring
      int as = (avail / 2) / a.a_stage_bytes;
      if (as > a.n_chunks * 3) as = a.n_chunks * 3;   // no point in buffering more than ~3 tiles of A
      if (as > MAX_A_STAGES) as = MAX_A_STAGES;
      if (as < 2) as = 2;
      for (; as >= 2 && a.mode == 0; --as) {
        int bs = (avail - as * a.a_stage_bytes) / a.b_stage_bytes;
        if (bs > MAX_B_STAGES) bs = MAX_B_STAGES;
        if (bs >= 2) {
          a.mode = 1; a.b_resident = 0; a.tg = tg; a.a_stages = as; a.b_stages = bs; a.b_region_bytes = bs * a.b_stage_bytes;
        }
      }
    }
  }
  if (a.mode == 1) {
    a.ring_bytes = a.a_stages * a.a_stage_bytes + a.b_region_bytes;
  } else {
    a.stage_bytes = A_STAGE_BYTES + bn * BK * 2;
    a.stages = avail / a.stage_bytes;
    if (a.stages > MAX_A_STAGES) a.stages = MAX_A_STAGES;
    SIB_REQUIRE(a.stages >= 2, "sib_conv1d_bf16: shared memory budget too small");
    a.ring_bytes = a.stages * a.stage_bytes;
    a.a_stages = 0;
  }
  const int smem_bytes = a.ring_bytes + slab_total + 1024 /*barriers*/ + 1024 /*alignment slack*/;

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
  CUtensorMap map_a, map_b, map_y, map_y2, map_r;
  {
    // A viewed as [batch][ceil(t_in/s)][s*c_in]; for s>1 the last (partial) row may extend past t_in: the caller
    // guarantees those bytes are readable (see header) - they only feed masked outputs.
    const cuuint64_t dims[3] = {(cuuint64_t)d->c_in * s, (cuuint64_t)((d->t_in + s - 1) / s), (cuuint64_t)d->batch};
    const cuuint64_t strides[3] = {2, (cuuint64_t)d->x_row_stride * s * 2, (cuuint64_t)d->x_batch_stride * 2};
    const cuuint32_t box[3] = {(cuuint32_t)cc, (cuuint32_t)(a.mode == 1 ? a.rows_h : BM), 1};
    if (int rc = encode_map(&map_a, x, 3, dims, strides, box, swz_of(row_bytes), "A")) return rc;
  }
  {
    // weights: [groups][c_in/g / cc][n_taps][c_out/g][cc]; one (chunk, tap) slab = the K-major B operand of one MMA group
    const cuuint64_t dims[3] = {(cuuint64_t)cc, (cuuint64_t)cout_g, (cuuint64_t)d->groups * a.n_chunks * d->n_taps};
    const cuuint64_t strides[3] = {2, (cuuint64_t)cc * 2, (cuuint64_t)cout_g * cc * 2};
    const cuuint32_t box[3] = {(cuuint32_t)cc, (cuuint32_t)bn, (cuuint32_t)a.tg};
    if (int rc = encode_map(&map_b, w, 3, dims, strides, box, swz_of(row_bytes), "B")) return rc;
  }
  {
    // output / residual slabs: boxes of 32 rows x cw channels, swizzle width = cw*2 bytes
    const cuuint64_t dims[3] = {(cuuint64_t)d->c_out, (cuuint64_t)d->t_out, (cuuint64_t)d->batch};
    const cuuint64_t ys[3] = {2, (cuuint64_t)d->y_row_stride * 2, (cuuint64_t)d->y_batch_stride * 2};
    const cuuint32_t box[3] = {(cuuint32_t)cw, 32, 1};
    if (int rc = encode_map(&map_y, y, 3, dims, ys, box, swz_of(cw * 2), "Y")) return rc;
    if (int rc = encode_map(&map_y2, y_act ? y_act : y, 3, dims, ys, box, swz_of(cw * 2), "Y2")) return rc;
    const cuuint64_t rs[3] = {2, (cuuint64_t)(residual ? d->r_row_stride : d->y_row_stride) * 2,
                              (cuuint64_t)(residual ? d->r_batch_stride : d->y_batch_stride) * 2};
    if (int rc = encode_map(&map_r, residual ? residual : y, 3, dims, rs, box, swz_of(cw * 2), "R")) return rc;
  }
  static int sm_count[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && sm_count[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    const void* fns[4] = {(const void*)conv1d_bf16_tc_kernel<SIB_ACT_NONE>, (const void*)conv1d_bf16_tc_kernel<SIB_ACT_GELU>,
                          (const void*)conv1d_bf16_tc_kernel<SIB_ACT_LRELU>, (const void*)conv1d_bf16_tc_kernel<SIB_ACT_TANH>};
    for (const void* fn : fns) {
      cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) {
        sib::set_error("sib_conv1d_bf16: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return SIB_ERR_CUDA;
      }
    }
    sm_count[dev] = n > 0 ? n : 148;
  }
  const int nsm = (dev >= 0 && dev < 64) ? sm_count[dev] : 148;
  const int grid = a.total_tiles < nsm * ctas_per_sm ? a.total_tiles : nsm * ctas_per_sm;
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  switch (d->post_act) {
    case SIB_ACT_NONE:
      conv1d_bf16_tc_kernel<SIB_ACT_NONE><<<grid, NUM_THREADS, smem_bytes, cs>>>(map_a, map_b, map_y, map_y2, map_r, a);
      break;
    case SIB_ACT_GELU:
      conv1d_bf16_tc_kernel<SIB_ACT_GELU><<<grid, NUM_THREADS, smem_bytes, cs>>>(map_a, map_b, map_y, map_y2, map_r, a);
      break;
    case SIB_ACT_LRELU:
      conv1d_bf16_tc_kernel<SIB_ACT_LRELU><<<grid, NUM_THREADS, smem_bytes, cs>>>(map_a, map_b, map_y, map_y2, map_r, a);
      break;
    case SIB_ACT_TANH:
      conv1d_bf16_tc_kernel<SIB_ACT_TANH><<<grid, NUM_THREADS, smem_bytes, cs>>>(map_a, map_b, map_y, map_y2, map_r, a);
      break;
    default:
      SIB_REQUIRE(false, "sib_conv1d_bf16: unknown post_act %d", d->post_act);
  }
  SIB_CHECK_LAUNCH("sib_conv1d_bf16");
  return SIB_OK;
}
