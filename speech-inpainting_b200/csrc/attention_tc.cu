// bf16 multi-head self-attention on the 5th-gen tensor cores (HF:234-259 eager_attention_forward, HF:296-345):
//   out = softmax(q k^T * d^-1/2 + key-padding mask) v,   head_dim 64, T up to a few hundred frames.
// One CTA per (128-query tile, head, utterance); keys stream in blocks of 128 with an online softmax:
//   warp 4     TMA producer: Q tile once, then (K, V) blocks through a 2-stage ring - all boxes come straight out
//              of the packed [B, T, 3H] projection output through ONE 3-D tensor map (channel, frame, utterance);
//              frames past T are TMA zero fill, so no tile ever reads another utterance
//   warp 5     single-thread tcgen05.mma issuer:  S = Q K^T  (M 128 x N keys x K 64, both operands K-major) into
//              TMEM columns [0, 128);  O_blk = P V  (M 128 x N 64 x K keys; V is read in place as an MN-major B
//              operand, no transpose pass) into TMEM columns [128, 192)
//   warps 0-3  softmax: thread r owns query row r == TMEM lane r.  Two tcgen05.ld passes over S (row max, then
//              exp2 / row sum), P goes to shared memory as bf16 in the 128-byte-swizzled K-major layout the P V
//              MMA reads; the running output lives in registers and is rescaled per block (fp32 throughout).
// Two CTAs fit per SM (112 KB of shared memory, 256 TMEM columns each), so one CTA's softmax overlaps the
// other's MMAs and loads.
#include "tc_common.cuh"

namespace {

using namespace sib_tc;

constexpr int HD = 64;     // head dim == one 128-byte swizzle row
constexpr int QT = 128;    // queries per CTA == UMMA M == TMEM lanes
constexpr int KB = 128;    // keys per block
constexpr int KV_STAGES = 2;
constexpr int TILE_BYTES = 128 * HD * 2;          // 16 KB: Q tile, one K block, one V block, one 64-key P chunk
constexpr int SM_Q = 0;
constexpr int SM_K = SM_Q + TILE_BYTES;
constexpr int SM_V = SM_K + KV_STAGES * TILE_BYTES;
constexpr int SM_P = SM_V + KV_STAGES * TILE_BYTES;
constexpr int SM_BAR = SM_P + 2 * TILE_BYTES;     // 112 KB of tiles
constexpr int SMEM_BYTES = SM_BAR + 128;
constexpr int NUM_THREADS = 192;
constexpr uint32_t TMEM_COLS = 256;               // S: [0,128)  O_blk: [128,192)
constexpr uint32_t TMEM_O = 128;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Persistent: gridDim.x CTAs (two per SM) walk the (query tile, head, utterance) work items with stride gridDim.x.  Barriers,
// the TMEM allocation and the tensor-map prefetch are paid once per CTA, and the producer warp runs ahead across items: the
// next item's Q is fetched as soon as the last S MMAs of the current one have retired (q_empty), its first K / V block as
// soon as a ring stage is free - both under the current item's softmax.  All barrier parities follow running counters.
//
// FLOW (sib_attention_flow_bf16, see sib_flow in the header): the kernel boundaries on both sides become per-128-row-block
// counters over the FLAT rows b * T + t.  An item reads Q / K / V of utterance b once every row block that utterance touches
// has been completed by the QKV projection (wait); after its output rows are stored it adds 2 per row (head_dim / 32) to the
// counters of the one or two row blocks they lie in (signal), so a block is complete at rows x heads x 2 = rows x H / 32 -
// the same accounting as a LayerNorm producer, which is what the out-projection's gate expects.
template <bool FLOW>
__global__ void __launch_bounds__(NUM_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, __nv_bfloat16* __restrict__ out,
                    const int32_t* __restrict__ key_len, int T, int heads, int tiles_q, int total_items, const sib_flow flow,
                    int total_rows) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* p_empty = bars + 7;
  uint64_t* o_full = bars + 8;
  uint64_t* o_empty = bars + 9;
  uint64_t* q_empty = bars + 10;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform role dispatch
  const int H = heads * HD;
  // item -> (query tile fastest, head, utterance): neighbouring CTAs share K / V of one (head, utterance) in L2
  auto decode = [&](int item, int& q0, int& h, int& b) {
    const int qt = item % tiles_q;
    const int rest = item / tiles_q;
    q0 = qt * QT;
    h = rest % heads;
    b = rest / heads;
  };
  auto blocks_of = [&](int b) {
    const int kl = key_len ? max(1, min(key_len[b], T)) : T;
    return kl;
  };

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) asm volatile("trap;");  // the swizzled tiles need 1024-byte alignment
    prefetch_tensormap(&map_qkv);
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(p_empty, 1);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 128);
    mbar_init(q_empty, 1);
    fence_barrier_init();
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (!(FLOW && flow.wait)) sib::pdl_wait();   // PDL: the prologue above overlapped the previous kernel's tail
  sib::pdl_launch_dependents();

  if (warp == 4) {
    // ===================== TMA producer (whole warp in the loop, one elected lane issues) =====================
    const uint32_t issuer = elect_one_sync();
    uint32_t g = 0, it = 0;                     // running K/V block counter, running item counter
    int gated_b = -1;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      int q0, h, b;
      decode(item, q0, h, b);
      const int nblk = (blocks_of(b) + KB - 1) / KB;
      if (FLOW && flow.wait && b != gated_b) {
        // every row block utterance b touches has all 3H columns of the projection (every lane polls: any may be the issuer)
        const int last_rb = (total_rows - 1) >> 7;
        for (int rb = (b * T) >> 7; rb <= (b * T + T - 1) >> 7; ++rb)
          sib::flow_wait(flow.wait + rb, rb == last_rb ? flow.wait_target_last : flow.wait_target);
        gated_b = b;
      }
      mbar_wait(q_empty, (it & 1) ^ 1);         // the S MMAs of the previous item have read Q
      if (issuer) {
        mbar_expect_tx(q_full, TILE_BYTES);
        tma_load_3d(smem + SM_Q, &map_qkv, q_full, h * HD, q0, b);
      }
      for (int j = 0; j < nblk; ++j, ++g) {
        const int s = (int)(g & 1);
        mbar_wait(&kv_empty[s], ((g >> 1) & 1) ^ 1);
        if (issuer) {
          mbar_expect_tx(&kv_full[s], 2 * TILE_BYTES);
          tma_load_3d(smem + SM_K + s * TILE_BYTES, &map_qkv, &kv_full[s], H + h * HD, j * KB, b);
          tma_load_3d(smem + SM_V + s * TILE_BYTES, &map_qkv, &kv_full[s], 2 * H + h * HD, j * KB, b);
        }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer: warp-uniform loop (descriptors in uniform registers), one lane issues ======
    {
      const uint32_t issuer = elect_one_sync();
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      constexpr uint32_t DESC_HI = make_desc_hi(128);
      const uint64_t qdesc = make_smem_desc(smem_u32(smem + SM_Q), DESC_HI);
      const uint64_t pdesc = make_smem_desc(smem_u32(smem + SM_P), DESC_HI);
      constexpr uint32_t IDESC_O = make_idesc_bf16(QT, HD, /*b_mn_major=*/1);
      uint32_t g = 0, it = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      int q0, h, b;
      decode(item, q0, h, b);
      const int kl = blocks_of(b);
      const int nblk = (kl + KB - 1) / KB;
      mbar_wait(q_full, it & 1);
      for (int j = 0; j < nblk; ++j, ++g) {
        const int s = (int)(g & 1);
        const int nk16 = (min(KB, kl - j * KB) + 15) & ~15;
        mbar_wait(&kv_full[s], (g >> 1) & 1);
        tc_fence_after();
        // S = Q K^T.  (S of block j-1 has been consumed: the P V MMAs of j-1 were issued after p_full(j-1).)
        const uint64_t kdesc = make_smem_desc(smem_u32(smem + SM_K + s * TILE_BYTES), DESC_HI);
        const uint32_t idesc_s = make_idesc_bf16(QT, nk16);
        if (issuer) {
#pragma unroll
          for (int ks = 0; ks < HD / 16; ++ks) umma_bf16(tmem_u, qdesc + 2 * ks, kdesc + 2 * ks, idesc_s, ks > 0);
          umma_commit(s_full);
          if (j == nblk - 1) umma_commit(q_empty);   // last use of Q by this item: the producer may fetch the next one
        }
        // O_blk = P V
        mbar_wait(p_full, g & 1);
        mbar_wait(o_empty, (g & 1) ^ 1);
        tc_fence_after();
        const uint64_t vdesc = make_smem_desc(smem_u32(smem + SM_V + s * TILE_BYTES), DESC_HI);
        if (issuer) {
          for (int ks = 0; ks < nk16 / 16; ++ks) {
            // A: 16 keys = 32 bytes inside the swizzle row of P chunk ks/4;  B: 16 key rows of V = 2048 bytes further
            const uint64_t a = pdesc + (uint64_t)((ks >> 2) * (TILE_BYTES >> 4) + (ks & 3) * 2);
            const uint64_t bd = vdesc + (uint64_t)(ks * (16 * 128 >> 4));
            umma_bf16(tmem_u + TMEM_O, a, bd, IDESC_O, ks > 0);
          }
          umma_commit(o_full);
          umma_commit(&kv_empty[s]);
          umma_commit(p_empty);
        }
      }
      }
    }
  } else {
    // ===================== softmax + output: thread r = query row r = TMEM lane r =====================
    const int r = threadIdx.x;
    const uint32_t t_s = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t t_o = t_s + TMEM_O;
    const uint32_t prow = smem_u32(smem + SM_P + r * 128);
    const uint32_t swz = (uint32_t)(r & 7);
    const float c = 0.125f * 1.4426950408889634f;  // d^-1/2 * log2(e)
    uint32_t g = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
    int q0, h, b;
    decode(item, q0, h, b);
    const int kl = blocks_of(b);
    const int nblk = (kl + KB - 1) / KB;
    float o_acc[HD];
#pragma unroll
    for (int i = 0; i < HD; ++i) o_acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nblk; ++j, ++g) {
      const int nk = min(KB, kl - j * KB);
      const int nk16 = (nk + 15) & ~15;
      mbar_wait(s_full, g & 1);
      tc_fence_after();
      // only the last 16-column group of a block can hold keys past the utterance's end: the full groups run without the
      // per-element compare / select (two of the ~7 instructions per score; the softmax warps are 60 % ALU-busy, ncu)
      const int nfull = nk & ~15;
      float mx = -INFINITY;
      int c1 = 0;
      for (; c1 + 32 <= nfull; c1 += 32) {            // two TMEM loads in flight per wait
        uint32_t va[16], vb[16];
        tmem_ld16_nowait(t_s + (uint32_t)c1, va);
        tmem_ld16_nowait(t_s + (uint32_t)(c1 + 16), vb);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(va[i]), __uint_as_float(vb[i])));
      }
      for (; c1 < nfull; c1 += 16) {
        uint32_t v[16];
        tmem_ld16(t_s + (uint32_t)c1, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      if (nfull < nk16) {
        uint32_t v[16];
        tmem_ld16(t_s + (uint32_t)nfull, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) mx = fmaxf(mx, (nfull + i < nk) ? __uint_as_float(v[i]) : -INFINITY);
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = fast_exp2((m_run - m_new) * c);   // exp2(-inf) = 0 on the first block
      const float mc = m_new * c;
      mbar_wait(p_empty, (g & 1) ^ 1);                      // the P V MMAs of the previous block have read P
      float sum = 0.f;
      auto emit_p = [&](const float (&p)[16], int c0) {
        const uint32_t chunk = prow + (uint32_t)((c0 >> 6) * TILE_BYTES);
        const uint32_t u = (uint32_t)((c0 & 63) >> 3);
        float lo[8], hi[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { lo[i] = p[i]; hi[i] = p[8 + i]; }
        sts128(chunk + ((u ^ swz) << 4), pack8(lo));
        sts128(chunk + (((u + 1) ^ swz) << 4), pack8(hi));
      };
      for (int c0 = 0; c0 < nfull; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(t_s + (uint32_t)c0, v);
        float p[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          p[i] = fast_exp2(fmaf(__uint_as_float(v[i]), c, -mc));
          sum += p[i];
        }
        emit_p(p, c0);
      }
      if (nfull < nk16) {
        uint32_t v[16];
        tmem_ld16(t_s + (uint32_t)nfull, v);
        float p[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          p[i] = (nfull + i < nk) ? fast_exp2(fmaf(__uint_as_float(v[i]), c, -mc)) : 0.f;
          sum += p[i];
        }
        emit_p(p, nfull);
      }
      l_run = l_run * alpha + sum;
      m_run = m_new;
      tc_fence_before();
      fence_async_smem();   // generic-proxy stores of P -> visible to the tensor core's async-proxy reads
      mbar_arrive(p_full);
      mbar_wait(o_full, g & 1);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < HD; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(t_o + (uint32_t)c0, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) o_acc[c0 + i] = fmaf(o_acc[c0 + i], alpha, __uint_as_float(v[i]));
      }
      tc_fence_before();
      mbar_arrive(o_empty);
    }
    if (q0 + r < T) {
      const float inv = 1.f / l_run;
      uint4* dst = reinterpret_cast<uint4*>(out + ((int64_t)b * T + q0 + r) * H + h * HD);
#pragma unroll
      for (int u = 0; u < HD / 8; ++u) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = o_acc[8 * u + i] * inv;
        dst[u] = pack8(f);
      }
    }
    if (FLOW && flow.signal) {
      asm volatile("bar.sync 1, 128;" ::: "memory");     // the four softmax warps: this item's rows are stored
      if (r == 0) {
        const int row0 = b * T + q0, row1 = b * T + min(q0 + QT, T);   // flat rows [row0, row1)
        const int rb0 = row0 >> 7, rb1 = (row1 - 1) >> 7;
        const int split = min(row1, (rb0 + 1) << 7);
        sib::flow_signal(flow.signal + rb0, 2 * (split - row0));
        if (rb1 != rb0) sib::flow_signal(flow.signal + rb1, 2 * (row1 - split));
      }
    }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace

// bf16 arm of sib_attention (see attention_f32.cu for the dispatcher and the fp32 SIMT arm).
static int attention_bf16_tc_impl(const void* qkv, const int32_t* key_len, void* out, int batch, int t, int heads,
                                  cudaStream_t stream, const sib_flow* flow) {
  const int H = heads * HD;
  CUtensorMap map;
  const cuuint64_t dims[3] = {(cuuint64_t)3 * H, (cuuint64_t)t, (cuuint64_t)batch};
  const cuuint64_t strides[3] = {2, (cuuint64_t)3 * H * 2, (cuuint64_t)t * 3 * H * 2};
  const cuuint32_t box[3] = {HD, 128, 1};
  if (int rc = encode_map(&map, qkv, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, "sib_attention", "qkv")) return rc;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute((const void*)attention_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute((const void*)attention_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) {
      sib::set_error("sib_attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return SIB_ERR_CUDA;
    }
    attr_set[dev] = true;
  }
  const int tiles_q = sib::ceil_div(t, QT);
  const int64_t items = (int64_t)tiles_q * heads * batch;
  if (items >= (1ll << 31)) {
    sib::set_error("sib_attention: too many work items");
    return SIB_ERR_INVALID;
  }
  const int slots = 2 * sib_tc::sm_count_of_current_device();        // two persistent CTAs per SM
  dim3 grid((unsigned)(items < slots ? items : slots));
  const sib_flow nf = {nullptr, nullptr, 0, 0};
  const cudaError_t le =
      flow ? sib::launch_pdl(attention_tc_kernel<true>, grid, dim3(NUM_THREADS), (size_t)SMEM_BYTES, stream, map,
                             (__nv_bfloat16*)out, key_len, t, heads, tiles_q, (int)items, *flow, batch * t)
           : sib::launch_pdl(attention_tc_kernel<false>, grid, dim3(NUM_THREADS), (size_t)SMEM_BYTES, stream, map,
                             (__nv_bfloat16*)out, key_len, t, heads, tiles_q, (int)items, nf, batch * t);
  if (le != cudaSuccess) {
    sib::set_error("sib_attention: launch failed: %s", cudaGetErrorString(le));
    return SIB_ERR_CUDA;
  }
  SIB_CHECK_LAUNCH("sib_attention");
  return SIB_OK;
}

int sib_attention_bf16_tc(const void* qkv, const int32_t* key_len, void* out, int batch, int t, int heads,
                          cudaStream_t stream) {
  return attention_bf16_tc_impl(qkv, key_len, out, batch, t, heads, stream, nullptr);
}

extern "C" int sib_attention_flow_bf16(const void* qkv, const int32_t* key_len, void* out, int batch, int t, int heads,
                                       int head_dim, const sib_flow* flow, sib_stream_t stream) {
  SIB_REQUIRE(qkv && out && flow && batch > 0 && t > 0 && heads > 0, "sib_attention_flow_bf16: bad argument");
  SIB_REQUIRE(head_dim == HD, "sib_attention_flow_bf16: head_dim must be %d", HD);
  SIB_REQUIRE((int64_t)batch * t < (1ll << 31), "sib_attention_flow_bf16: too many rows");
  SIB_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              "sib_attention_flow_bf16: qkv / out must be 16-byte aligned");
  SIB_REQUIRE(!flow->wait || (flow->wait_target > 0 && flow->wait_target_last > 0), "sib_attention_flow_bf16: wait targets");
  return attention_bf16_tc_impl(qkv, key_len, out, batch, t, heads, static_cast<cudaStream_t>(stream), flow);
}
