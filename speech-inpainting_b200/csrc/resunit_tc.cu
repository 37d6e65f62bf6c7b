// Fused HiFi-GAN ResBlock1 unit on the tensor cores (I_ea/hifi_gan/models.py:36-43, I_da/src/models.py ResBlock1):
//     xt = leaky_relu(x, 0.1); xt = conv1_{k, dilation d}(xt); xt = leaky_relu(xt, 0.1); xt = conv2_{k, 1}(xt); x = xt + x
// for the narrow, HBM-bound stages (C = 64 / 32 / 16 channels, frame-major bf16 [B, T, C]).  Unfused, one unit moves six
// activation tensors through HBM (conv1: read lrelu(x), write t1; conv2: read t1, read x, write y, write lrelu(y));
// here it reads x once (plus the dilated halo) and writes y once:
//   * TMA brings a (128 + 2 p1)-row tile of raw x into shared memory (zero fill outside [0, T) == the convs' padding);
//     two warps apply the leaky-relu IN PLACE (generic proxy -> fence.proxy.async) - the TMA-fed MMA cannot
//     transform its A operand, but it can be handed a transformed tile;
//   * conv1 = k taps x (C/16) tcgen05.mma, every tap a row-shifted descriptor on that tile, weights resident;
//   * epilogue 1: TMEM -> + b1 -> leaky-relu -> rows outside [0, T) zeroed (conv2 pads t1 with zeros, not with
//     conv1 of padding) -> bf16 -> shared memory in the swizzled K-major layout of an A operand (never touches HBM);
//   * conv2 = k taps on that intermediate tile; epilogue 2: + b2 + x (+ the running MRF sum) x scale -> TMA store.
//     R = 128 - (k - 1) output rows per tile (the intermediate needs conv2's halo); the last quarter stores through a
//     shorter box so that rows >= R are never written.
// Warp roles (512 threads): 0 TMA producer, 1 MMA issuer (conv1 runs one tile ahead of conv2 so the tensor pipe works
// while epilogue 1 converts), 2-3 + 12-15 activation, 4-7 epilogue 1, 8-11 epilogue 2 (one warp per TMEM lane quarter).
// Both accumulators are double-buffered in TMEM (4 C columns).  C = 32 / 16 fit two CTAs per SM.
#include <stdlib.h>

#include "tc_common.cuh"

namespace {

using namespace sib_tc;

constexpr int NUM_THREADS = 512;
constexpr int ACT_WARPS = 6;                        // warps 2-3 and 12-15
// wide (C = 128) builds: 18 warps - epilogue 1 gets EIGHT warps (4-7: columns 0-63, 12-15: columns 64-127), because with the
// single intermediate buffer that fits, conv2(i) -> epilogue 1(i+1) -> conv2(i+1) is a serial chain and its length is
// what paces the kernel; the activation warps are 2-3 only (bf16x2 arithmetic: two warps keep up with the 36 KB tiles)
constexpr int NUM_THREADS_WIDE = 640;              // 20 warps: producer, MMA, 2 activation, 8 epilogue 1, 8 epilogue 2 (<= 102 registers)
constexpr int ACT_WARPS_WIDE = 2;

struct RuArgs {
  const float* b1;
  const float* b2;
  int T, C, k, dil, batch;
  int R, p1, p2, xr, tail_rows;
  int tiles_m, total_tiles;
  int row_bytes, ksteps, tap_bytes;
  int x_stage_bytes, t1_bytes, w_bytes, w_tx_bytes, w_tg, w_loads;   // per-conv weight region: w_loads boxes of w_tg taps
  int stage_box_bytes;                                   // one staging tile: 128 rows x C bf16
  int need_b;                                            // second staging box per slot (accumulate input / y_act output)
  int nxs;                                               // x-tile ring depth (2..4): short tiles need the loads further ahead
  int slots;                                             // staging slots (2, or 1 when shared memory is tight)
  int t1_bufs;                                           // intermediate tile ring (1..3)
  int la, na1;                                           // conv1 runs `la` tiles ahead of conv2; acc1 ring = la + 1 buffers
  int c_cta;                                             // weight rows (output channels) this CTA holds per tap: C, or C/2 in pair mode
  float slope_in, slope_mid, out_scale, act2_slope;
  int accumulate, has_y2;
  int early_w;                                           // resident weights fetched before the PDL dependency wait
  int e2w;                                               // epilogue-2 warps of the C <= 64 builds: 4, or 8 (warps 12-15 join: two
                                                         // column halves per TMEM lane quarter; the activation stage keeps warps 2-3)
  // C = 128 ("wide", CTA pairs only): a row is TWO 64-channel chunks, every operand tile is two 128-byte-swizzled slabs
  int nch;                                               // channel chunks per row: 1, or 2 when C = 128
  int x_chunk_bytes, t1_chunk_bytes;                     // one chunk slab of an x stage / of the intermediate tile
  const void* xg;                                        // wide: epilogue 2 reads the residual rows straight from global x
  long long xg_batch_stride;                             //       (elements)
  int xg_row_stride;
  uint32_t desc_hi, idesc;
  uint32_t tmem_cols;
};

// 16-byte shared-memory load of data that is constant for the kernel's lifetime (the bias vectors): NOT volatile and no
// memory clobber, so the compiler may hoist and batch it freely
__device__ __forceinline__ float4 lds_const_f4(uint32_t saddr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
// read-only 16-byte global load that does not allocate in L1 (streamed once per tile row)
__device__ __forceinline__ uint4 ldg_stream(const uint4* ptr) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr));
  return v;
}

template <bool PAIR>
__device__ __forceinline__ void commit_u(uint64_t* bar) {
  if (PAIR) umma_commit_pair(bar, (uint16_t)3);
  else umma_commit(bar);
}

// PAIR: two CTAs (cluster of 2, cta_group::2) run two independent tiles in lock-step and split the OUTPUT CHANNELS of both
// weight sets between their shared memories (C/2 rows per tap each), which is what lets C = 64, k = 11 (176 KB of weights)
// stay resident.  The even CTA issues M = 256 MMAs for both; every hand-off towards the MMA warp (activated x tile,
// intermediate tile, drained accumulators) is a remote arrive on the even CTA's barrier, every hand-off from it a
// multicast commit.
// WIDE (C = 128, pairs only, k <= 3): both weight sets split over the pair are 96 KB per CTA, which leaves room for two x
// stages, ONE intermediate tile and 16 KB of output staging - so epilogue 2 takes its residual rows straight from global
// memory (the tile's x rows were fetched by TMA moments ago: L2 hits) and drains the tile in four 32-channel groups
// through two small staging boxes per warp.  A separate instantiation: the C <= 64 builds keep their code.
template <bool PAIR, bool WIDE = false>
__global__ void __launch_bounds__(WIDE ? NUM_THREADS_WIDE : NUM_THREADS, PAIR ? 1 : 2)
resunit_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1,
                  const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_res,
                  const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_yt,
                  const __grid_constant__ CUtensorMap map_y2, const __grid_constant__ CUtensorMap map_y2t,
                  const __grid_constant__ RuArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sm_x = smem;                                   // [nxs] raw -> activated x tiles
  uint8_t* sm_t1 = sm_x + p.nxs * p.x_stage_bytes;        // conv1 output tile (A operand of conv2)
  uint8_t* sm_w1 = sm_t1 + p.t1_bufs * p.t1_bytes;
  uint8_t* sm_w2 = sm_w1 + p.w_bytes;
  uint8_t* sm_sa = sm_w2 + p.w_bytes;                     // [2] staging A: residual in -> y out
  uint8_t* sm_sb = sm_sa + p.slots * p.stage_box_bytes;   // [slots] staging B: accumulate in -> y_act out
  // wide: both bias vectors live in shared memory (the residual rows that epilogue 2 streams through L1 evict the bias
  // lines: ncu showed every bias add waiting on an L2 round trip)
  float* sm_bias = reinterpret_cast<float*>(sm_sb + (p.need_b ? p.slots * p.stage_box_bytes : 0));   // [2][128] when WIDE
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sm_bias) + (WIDE ? 1024 : 0));
  uint64_t* x_full = bars;            // [4]
  uint64_t* x_empty = bars + 4;       // [4]
  uint64_t* act_done = bars + 8;      // [4]
  uint64_t* acc1_full = bars + 12;    // [4]
  uint64_t* acc1_empty = bars + 16;   // [4]
  uint64_t* acc2_full = bars + 20;    // [2]
  uint64_t* acc2_empty = bars + 22;   // [2]
  uint64_t* t1_full = bars + 24;      // [4]
  uint64_t* t1_empty = bars + 28;     // [4]
  uint64_t* w_full = bars + 32;
  uint64_t* res_bar = bars + 33;      // [4 quarters][3 slots]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 45);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&map_x);
    prefetch_tensormap(&map_w1);
    prefetch_tensormap(&map_w2);
    prefetch_tensormap(&map_res);
    prefetch_tensormap(&map_y);
    for (int s = 0; s < 4; ++s) {
      mbar_init(&x_full[s], 1);
      mbar_init(&x_empty[s], 1);
      mbar_init(&act_done[s], (PAIR ? 2 : 1) * (WIDE ? ACT_WARPS_WIDE : (p.e2w == 8 ? 2 : ACT_WARPS)));
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&acc1_full[s], 1);
      mbar_init(&acc1_empty[s], WIDE ? 16 : (PAIR ? 8 : 4));     // one arrive per epilogue-1 warp (of both CTAs)
      mbar_init(&t1_full[s], WIDE ? 16 : (PAIR ? 8 : 4));
      mbar_init(&t1_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc2_full[s], 1);
      mbar_init(&acc2_empty[s], WIDE ? 16 : (PAIR ? 2 : 1) * p.e2w);
    }
    mbar_init(w_full, 1);
    for (int s = 0; s < 12; ++s) mbar_init(&res_bar[s], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                   "r"(p.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                   "r"(p.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  if (WIDE && threadIdx.x >= 128 && threadIdx.x < 384) {       // parameters, not activations: no dependency wait needed
    const int i = (int)threadIdx.x - 128;
    sm_bias[i] = i < 128 ? p.b1[i] : p.b2[i - 128];
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  // arrive on a barrier the MMA warp waits on: in pair mode that barrier lives in the even CTA
  auto arrive_mma = [&](uint64_t* bar) {
    if (PAIR) mbar_arrive_cluster(bar, 0);
    else mbar_arrive(bar);
  };
  // PDL: the prologue above overlapped the previous kernel's tail.  Only the x producer and epilogue 2 touch activations
  // in global memory and wait for the previous grid; the resident weights are fetched before that wait.
  sib::pdl_launch_dependents();

  // work list: tile i of this CTA = first + i * step.  Pair mode: cluster c takes tile pairs c, c + #clusters, ...; CTA r of
  // the pair runs tile 2 * pair + r (a missing odd tile is a masked duplicate of the last one: loads clamp, stores skip)
  const int first = PAIR ? 2 * (int)(blockIdx.x >> 1) + (int)cta_rank : (int)blockIdx.x;
  const int step = PAIR ? 2 * (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int units = PAIR ? (p.total_tiles + 1) / 2 : p.total_tiles;           // tiles or tile pairs
  const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int ustep = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int n_my = units > unit0 ? (units - 1 - unit0) / ustep + 1 : 0;
  // tile i of this CTA = blockIdx.x + i * gridDim.x = (utterance b, row tile mt); every role walks the same sequence with
  // add-and-wrap counters (two integer divisions per kernel instead of one ~40-instruction division per tile and warp:
  // on the short k = 3 tiles the SM is instruction-issue bound, ncu 2.7 of 4 IPC)
  struct TileCursor {
    int b, mt, step_b, step_m, tiles_m, R, batch;
    __device__ __forceinline__ void next() {
      mt += step_m;
      b += step_b;
      if (mt >= tiles_m) { mt -= tiles_m; ++b; }
    }
    __device__ __forceinline__ bool valid() const { return b < batch; }
    __device__ __forceinline__ int bb() const { return b < batch ? b : batch - 1; }          // clamped for loads
    __device__ __forceinline__ int t0() const { return (b < batch ? mt : tiles_m - 1) * R; }
  };
  auto cursor0 = [&]() {
    TileCursor c;
    c.tiles_m = p.tiles_m; c.R = p.R; c.batch = p.batch;
    c.b = first / p.tiles_m; c.mt = first - c.b * p.tiles_m;
    c.step_b = step / p.tiles_m; c.step_m = step - c.step_b * p.tiles_m;
    return c;
  };

  if (warp == 0) {
    // ===================== TMA producer: weights once, then one raw x tile per output tile =====================
    const uint32_t issuer = elect_one_sync();
    if (!p.early_w) sib::pdl_wait();
    if (issuer) {
      // pair: each CTA loads its half of the output channels; both halves complete on the even CTA's barrier
      if (!PAIR || cta_rank == 0) mbar_expect_tx(w_full, (uint32_t)((PAIR ? 4 : 2) * p.w_tx_bytes));
      for (int l = 0; l < p.w_loads; ++l) {
        if (PAIR) {
          tma_load_3d_pair(sm_w1 + l * p.w_tg * p.tap_bytes, &map_w1, w_full, 0, (int)cta_rank * p.c_cta, l * p.w_tg);
          tma_load_3d_pair(sm_w2 + l * p.w_tg * p.tap_bytes, &map_w2, w_full, 0, (int)cta_rank * p.c_cta, l * p.w_tg);
        } else {
          tma_load_3d(sm_w1 + l * p.w_tg * p.tap_bytes, &map_w1, w_full, 0, 0, l * p.w_tg);
          tma_load_3d(sm_w2 + l * p.w_tg * p.tap_bytes, &map_w2, w_full, 0, 0, l * p.w_tg);
        }
      }
    }
    sib::pdl_wait();
    int s = 0;
    uint32_t ph = 0;
    TileCursor tc = cursor0();
    for (int i = 0; i < n_my; ++i, tc.next()) {
      mbar_wait(&x_empty[s], ph ^ 1);
      if (issuer) {
        mbar_expect_tx(&x_full[s], (uint32_t)((WIDE ? 2 : 1) * p.xr * p.row_bytes));
        tma_load_3d(sm_x + s * p.x_stage_bytes, &map_x, &x_full[s], 0, tc.t0() - p.p2 - p.p1, tc.bb());
        if (WIDE)
          tma_load_3d(sm_x + s * p.x_stage_bytes + p.x_chunk_bytes, &map_x, &x_full[s], 64, tc.t0() - p.p2 - p.p1, tc.bb());
      }
      if (++s == p.nxs) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1 && (!PAIR || cta_rank == 0)) {
    // ===================== MMA issuer: warp-uniform loop, one elected lane issues (pair: the even CTA) ===========
    const uint32_t issuer = elect_one_sync();
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t x_base = smem_u32(sm_x), t1_base = smem_u32(sm_t1);
    const uint32_t w1_lo = make_desc_lo(smem_u32(sm_w1)), w2_lo = make_desc_lo(smem_u32(sm_w2));
    const uint32_t a1_inc = (uint32_t)((p.dil * p.row_bytes) >> 4);
    const uint32_t a2_inc = (uint32_t)(p.row_bytes >> 4);
    const uint32_t w_inc = (uint32_t)(p.tap_bytes >> 4);
    mbar_wait(w_full, 0);
    tc_fence_after();
    int xs = 0, a1 = 0, tb = 0;
    uint32_t xph = 0, a1ph = 0, tbph = 0;
    auto conv1 = [&]() {                       // next tile in order: x ring slot xs, accumulator a1
      mbar_wait(&act_done[xs], xph);
      // wide: epilogue 1 signals "intermediate tile full" and "accumulator drained" at the same instant, and this warp has
      // always seen t1_full of the tile that last used this accumulator (conv2 of that tile precedes this point in program
      // order) - so the second barrier, and the release fence its remote arrive costs every epilogue-1 warp, are dropped
      if (!WIDE) mbar_wait(&acc1_empty[a1], a1ph ^ 1);
      tc_fence_after();
      umma_taps_ks<PAIR>(p.ksteps, issuer, tmem_u + (uint32_t)(a1 * p.C), make_desc_lo(x_base + (uint32_t)(xs * p.x_stage_bytes)), w1_lo,
                   a1_inc, w_inc, p.k, p.desc_hi, p.idesc, 0u);
      if (WIDE)   // second 64-channel chunk of the reduction: its own x slab and weight slabs [k, 2k)
        umma_taps_ks<PAIR>(p.ksteps, issuer, tmem_u + (uint32_t)(a1 * p.C),
                           make_desc_lo(x_base + (uint32_t)(xs * p.x_stage_bytes + p.x_chunk_bytes)), w1_lo + (uint32_t)p.k * w_inc,
                           a1_inc, w_inc, p.k, p.desc_hi, p.idesc, 1u);
      if (issuer) {
        commit_u<PAIR>(&acc1_full[a1]);
        if (!WIDE) commit_u<PAIR>(&x_empty[xs]);   // wide: the slot goes on to hold the intermediate tile (freed by conv2)
      }
      if (++xs == p.nxs) { xs = 0; xph ^= 1; }
      if (++a1 == p.na1) { a1 = 0; a1ph ^= 1; }
    };
    // wide: the intermediate tile of a tile is written IN PLACE over its x slot (conv1 has consumed it by then), so it is as
    // deeply buffered as the x ring and conv2(i) -> epilogue 1(i+1) -> conv2(i+1) is no longer a serial chain through one
    // shared buffer; conv2 frees the slot for the producer.
    auto conv2 = [&](int i) {
      const int a = i & 1;
      mbar_wait(&t1_full[tb], tbph);
      mbar_wait(&acc2_empty[a], (uint32_t)(((i >> 1) & 1) ^ 1));
      tc_fence_after();
      const uint32_t t1_tile = WIDE ? x_base + (uint32_t)(tb * p.x_stage_bytes) : t1_base + (uint32_t)(tb * p.t1_bytes);
      umma_taps_ks<PAIR>(p.ksteps, issuer, tmem_u + (uint32_t)((p.na1 + a) * p.C), make_desc_lo(t1_tile), w2_lo,
                   a2_inc, w_inc, p.k, p.desc_hi, p.idesc, 0u);
      if (WIDE)
        umma_taps_ks<PAIR>(p.ksteps, issuer, tmem_u + (uint32_t)((p.na1 + a) * p.C),
                           make_desc_lo(t1_tile + (uint32_t)p.x_chunk_bytes), w2_lo + (uint32_t)p.k * w_inc,
                           a2_inc, w_inc, p.k, p.desc_hi, p.idesc, 1u);
      if (issuer) {
        commit_u<PAIR>(&acc2_full[a]);
        commit_u<PAIR>(WIDE ? &x_empty[tb] : &t1_empty[tb]);
      }
      if (++tb == (WIDE ? p.nxs : p.t1_bufs)) { tb = 0; tbph ^= 1; }
    };
    // conv1 runs `la` tiles ahead: the commit -> epilogue 1 -> intermediate tile -> conv2 chain of one tile (~2-3k clocks
    // of barrier hand-offs) is then spread over la tiles instead of pacing every tile
    for (int i = 0; i < p.la && i < n_my; ++i) conv1();
    for (int i = 0; i < n_my; ++i) {
      if (i + p.la < n_my) conv1();
      conv2(i);
    }
  } else if (warp == 2 || warp == 3 || (!WIDE && p.e2w == 4 && warp >= 12)) {
    // ===================== activation: leaky-relu in place on the freshly landed x tile =====================
    // (six warps: with two, this stage paced the whole kernel on the short k = 3 tiles - the MMA warp sat on act_done)
    // bf16x2 arithmetic as in sib_conv1d_bf16: slope * x = x * hi + x * lo with hi + lo = slope to ~2^-17 (no slope bias
    // from rounding 0.1 to bf16), then max(x, slope * x): 12 instead of 28 ALU instructions per 16-byte chunk
    const int AW = WIDE ? ACT_WARPS_WIDE : (p.e2w == 8 ? 2 : ACT_WARPS);
    const int tid = (warp < 4 ? warp - 2 : warp - 10) * 32 + lane;
    const int n16 = ((WIDE ? p.x_chunk_bytes : 0) + p.xr * p.row_bytes) >> 4;   // wide: chunk 0's slab (with its pad) + chunk 1
    const __nv_bfloat16 s_hi = __float2bfloat16_rn(p.slope_in);
    const __nv_bfloat16 s_lo = __float2bfloat16_rn(p.slope_in - __bfloat162float(s_hi));
    const __nv_bfloat162 hi2 = __halves2bfloat162(s_hi, s_hi), lo2 = __halves2bfloat162(s_lo, s_lo);
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < n_my; ++i) {
      mbar_wait(&x_full[s], ph);
      const uint32_t tile = smem_u32(sm_x + s * p.x_stage_bytes);
      int e = tid;
      for (; e + 3 * AW * 32 < n16; e += 4 * AW * 32) {      // four chunks in flight per thread
        uint4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = lds128(tile + (uint32_t)(e + j * AW * 32) * 16u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v[j]);
#pragma unroll
          for (int u = 0; u < 4; ++u) h[u] = __hmax2(h[u], __hfma2(h[u], lo2, __hmul2(h[u], hi2)));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) sts128(tile + (uint32_t)(e + j * AW * 32) * 16u, v[j]);
      }
      for (; e < n16; e += AW * 32) {
        uint4 v = lds128(tile + (uint32_t)e * 16u);
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
        for (int u = 0; u < 4; ++u) h[u] = __hmax2(h[u], __hfma2(h[u], lo2, __hmul2(h[u], hi2)));
        sts128(tile + (uint32_t)e * 16u, v);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) arrive_mma(&act_done[s]);
      if (++s == p.nxs) { s = 0; ph ^= 1; }
    }
  } else if ((warp >= 4 && warp < 8) || (WIDE && warp >= 12 && warp < 16)) {
    // ===================== epilogue 1: conv1 accumulator -> lrelu -> bf16 A tile of conv2 =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;                               // tile row == TMEM lane == t1-local row
    const int chunks_per_row = p.row_bytes >> 4;
    const int swz_shift = p.row_bytes == 128 ? 0 : (p.row_bytes == 64 ? 1 : 2);
    const uint32_t swz = ((uint32_t)r >> swz_shift) & (uint32_t)(chunks_per_row - 1);
    int a = 0, tb = 0;
    uint32_t aph = 0, tbph = 0;
    TileCursor tc = cursor0();
    for (int i = 0; i < n_my; ++i, tc.next()) {
      const int t0 = tc.t0();
      const uint32_t row_ptr = smem_u32((WIDE ? sm_x + tb * p.x_stage_bytes : sm_t1 + tb * p.t1_bytes) + r * p.row_bytes);
      mbar_wait(&acc1_full[a], aph);
      // conv2 of the tile that last used this intermediate buffer has read it (wide: the tile's own x slot, which conv1
      // - complete, or acc1_full would not have fired - was the last reader of)
      if (!WIDE) mbar_wait(&t1_empty[tb], tbph ^ 1);
      tc_fence_after();
      const int tg = t0 - p.p2 + r;                            // global frame of this t1 row
      const bool inside = tg >= 0 && tg < p.T;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * p.C);
      auto emit16 = [&](const uint32_t (&v)[16], int c0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int col = c0 + 8 * h;
          const float4 ba = WIDE ? lds_const_f4(smem_u32(sm_bias + col)) : __ldg(reinterpret_cast<const float4*>(p.b1 + col));
          const float4 bb = WIDE ? lds_const_f4(smem_u32(sm_bias + col + 4)) : __ldg(reinterpret_cast<const float4*>(p.b1 + col + 4));
          float f[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            f[u] += __uint_as_float(v[8 * h + u]);
            f[u] = fmaxf(f[u], f[u] * p.slope_mid);          // leaky-relu, 0 < slope < 1
          }
          // (wide: columns 64.. live in the second chunk slab of the intermediate tile)
          const uint32_t cofs = WIDE ? (uint32_t)(col >> 6) * (uint32_t)p.x_chunk_bytes + (((((uint32_t)col & 63u) >> 3) ^ swz) << 4)
                                     : ((((uint32_t)col >> 3) ^ swz) << 4);
          sts128(row_ptr + cofs, inside ? pack8(f) : make_uint4(0u, 0u, 0u, 0u));
        }
      };
      const int c_lo = WIDE ? (warp >= 12 ? 64 : 0) : 0, c_hi = WIDE ? c_lo + 64 : p.C;   // wide: this warp's column half
      for (int c0 = c_lo; c0 < c_hi; c0 += 32) {               // two TMEM loads in flight per wait
        uint32_t va[16], vb[16];
        tmem_ld16_nowait(taddr + (uint32_t)c0, va);
        if (c0 + 16 < c_hi) tmem_ld16_nowait(taddr + (uint32_t)(c0 + 16), vb);
        tmem_ld_wait();
        emit16(va, c0);
        if (c0 + 16 < c_hi) emit16(vb, c0 + 16);
      }
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        arrive_mma(&t1_full[tb]);
        if (!WIDE) arrive_mma(&acc1_empty[a]);
      }
      if (++a == p.na1) { a = 0; aph ^= 1; }
      if (++tb == (WIDE ? p.nxs : p.t1_bufs)) { tb = 0; tbph ^= 1; }
    }
  } else if (WIDE && ((warp >= 8 && warp < 12) || warp >= 16)) {
    // ===================== epilogue 2, C = 128: eight warps, two 32-channel groups each ==========
    // (warps 8-11: channels 0-63, warps 16-19: channels 64-127 of the same lane quarter.)  Residual rows by plain
    // read-only loads from global x - this lane's 128-byte half row, issued before the accumulator wait and bypassing L1 -,
    // output through one 2 KB staging box per warp (32 rows x 32 channels, 64-byte swizzle) and TMA stores
    const int q = warp & 3;
    const int half = warp >= 16 ? 1 : 0;
    const uint32_t swz = ((uint32_t)lane >> 1) & 3u;           // 64-byte swizzle: chunk ^= (row >> 1) & 3
    const int rows_q = q < 3 ? 32 : p.tail_rows;
    sib::pdl_wait();
    uint8_t* my_box = sm_sa + (q * 2 + half) * 2048;           // [32 rows][64 B]
    const uint32_t box = smem_u32(my_box + lane * 64);
    const __nv_bfloat16* xg = reinterpret_cast<const __nv_bfloat16*>(p.xg);
    TileCursor tc = cursor0();
    for (int i = 0; i < n_my; ++i, tc.next()) {
      const int t0 = tc.t0(), b = tc.bb();
      const bool valid = tc.valid();
      const int a = i & 1;
      const int row = t0 + q * 32 + lane;
      const bool row_ok = row < p.T;
      const uint4* xrow = reinterpret_cast<const uint4*>(xg + (long long)b * p.xg_batch_stride +
                                                         (long long)(row_ok ? row : 0) * p.xg_row_stride) + half * 8;
      uint4 xq[2][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) xq[j >> 2][j & 3] = row_ok ? ldg_stream(xrow + j) : make_uint4(0u, 0u, 0u, 0u);
      mbar_wait(&acc2_full[a], (uint32_t)((i >> 1) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((p.na1 + a) * p.C + half * 64);
#pragma unroll
      for (int gg = 0; gg < 2; ++gg) {
        uint32_t va[16], vb[16];
        tmem_ld16_nowait(taddr + (uint32_t)(gg * 32), va);
        tmem_ld16_nowait(taddr + (uint32_t)(gg * 32 + 16), vb);
        // the box was handed to the TMA one group ago: its read-out must have finished before it is rewritten
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        tmem_ld_wait();
        if (gg == 1) {                                         // this warp's half of the accumulator is drained
          tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_mma(&acc2_empty[a]);
        } else {
          __syncwarp();
        }
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int col = half * 64 + gg * 32 + 8 * h;
          const float4 ba = lds_const_f4(smem_u32(sm_bias + 128 + col));
          const float4 bb = lds_const_f4(smem_u32(sm_bias + 128 + col + 4));
          float f[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
          if (p.out_scale == 1.f) {                            // y = bf16(conv + bias) + x in packed bf16x2 (see epilogue 2 below)
#pragma unroll
            for (int u = 0; u < 8; ++u) f[u] += __uint_as_float(h < 2 ? va[8 * h + u] : vb[8 * (h - 2) + u]);
            uint4 o = pack8(f);
            __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
            const __nv_bfloat162* x2 = reinterpret_cast<const __nv_bfloat162*>(&xq[gg][h]);
#pragma unroll
            for (int u = 0; u < 4; ++u) o2[u] = __hadd2(o2[u], x2[u]);
            sts128(box + ((((uint32_t)h) ^ swz) << 4), o);
          } else {
            float xf[8];
            unpack8(xq[gg][h], xf);
#pragma unroll
            for (int u = 0; u < 8; ++u) f[u] = (f[u] + __uint_as_float(h < 2 ? va[8 * h + u] : vb[8 * (h - 2) + u]) + xf[u]) * p.out_scale;
            sts128(box + ((((uint32_t)h) ^ swz) << 4), pack8(f));
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (rows_q > 0 && valid) tma_store_3d(q < 3 ? &map_y : &map_yt, my_box, half * 64 + gg * 32, t0 + q * 32, b);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    tc_fence_before();
  } else if (warp >= 8 && (warp < 12 || (!WIDE && p.e2w == 8))) {
    // ===================== epilogue 2: conv2 accumulator + bias + x (+ running sum) -> y (and lrelu(y)) ==========
    // e2w == 8 (r2): warps (8 + q, 12 + q) share TMEM lane quarter q and take the two column halves of the tile's rows - this
    // stage paced the k = 3 units (7 % of its samples in waits against 40-57 % for the others); warp 8 + q drives the TMA
    // traffic of the quarter's staging box, a 64-thread named barrier orders the partner's writes before the store
    const int q = warp & 3;
    const int half = warp >= 12 ? 1 : 0;
    const bool two = p.e2w == 8;
    const int c_begin = two ? half * (p.C >> 1) : 0, c_end = two ? c_begin + (p.C >> 1) : p.C;
    const bool leader = half == 0 && lane == 0;
    const int chunks_per_row = p.row_bytes >> 4;
    const int swz_shift = p.row_bytes == 128 ? 0 : (p.row_bytes == 64 ? 1 : 2);
    const uint32_t swz = ((uint32_t)lane >> swz_shift) & (uint32_t)(chunks_per_row - 1);
    const int box_bytes = 32 * p.row_bytes;
    const uint32_t pre_bytes = (uint32_t)(1 + (p.accumulate ? 1 : 0)) * (uint32_t)box_bytes;
    const int rows_q = q < 3 ? 32 : p.tail_rows;               // rows of this quarter that belong to the tile (R = 96 + tail)
    uint64_t* my_res = res_bar + q * 3;
    sib::pdl_wait();                                           // residual / running-sum reads and the stores below
    auto prefetch = [&](const TileCursor& c, int slot) {       // lane 0 only
      const int t0 = c.t0(), b = c.bb();
      mbar_expect_tx(&my_res[slot], pre_bytes);
      tma_load_3d(sm_sa + slot * p.stage_box_bytes + q * box_bytes, &map_res, &my_res[slot], 0, t0 + q * 32, b);
      if (p.accumulate)
        tma_load_3d(sm_sb + slot * p.stage_box_bytes + q * box_bytes, &map_y, &my_res[slot], 0, t0 + q * 32, b);
    };
    TileCursor tc = cursor0(), tn = cursor0();                 // this tile / the next one (prefetch target)
    if (leader && n_my > 0) prefetch(tn, 0);
    tn.next();
    int slot = 0;
    uint32_t res_phase_bits = 0;
    for (int i = 0; i < n_my; ++i, tc.next(), tn.next()) {
      const int t0 = tc.t0(), b = tc.bb();
      const bool valid = tc.valid();
      const int a = i & 1;
      const int next_slot = slot + 1 == p.slots ? 0 : slot + 1;
      if (leader && p.slots >= 2) {
        // the slot of tile i+1 was last stored from slots-1 tiles ago: with three slots the store of the previous tile
        // may still be draining while the next residual is already being fetched
        if (p.slots == 2) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        if (i + 1 < n_my) prefetch(tn, next_slot);
      }
      __syncwarp();
      const uint32_t box_a = smem_u32(sm_sa + slot * p.stage_box_bytes + q * box_bytes + lane * p.row_bytes);
      const uint32_t box_b = smem_u32(sm_sb + slot * p.stage_box_bytes + q * box_bytes + lane * p.row_bytes);
      mbar_wait(&acc2_full[a], (uint32_t)((i >> 1) & 1));
      mbar_wait(&my_res[slot], (res_phase_bits >> slot) & 1u);
      res_phase_bits ^= 1u << slot;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((p.na1 + a) * p.C);
      // 16 columns per call: both residual (and running-sum) chunks are fetched up front - the shared-memory accesses are
      // volatile asm and keep their program order, so loads issued inside the per-chunk loop would serialise every chunk's
      // LDS latency behind the previous chunk's stores (ncu: this warp group was 93 % busy and paced the k = 3 units)
      const bool simple = !p.accumulate && !p.has_y2 && p.out_scale == 1.f;   // 13 of the 18 units of a V1 forward
      auto emit16 = [&](const uint32_t (&v)[16], int c0) {
        const uint32_t off0 = (((uint32_t)c0 >> 3) ^ swz) << 4, off1 = ((((uint32_t)c0 >> 3) + 1) ^ swz) << 4;
        const uint4 x0 = lds128(box_a + off0), x1 = lds128(box_a + off1);
        if (simple) {
          // y = bf16(conv + bias) + x in packed bf16x2: four HADD2 per 8 columns instead of eight unpacks and eight
          // adds (this warp group paces the k = 3 / 7 units).  One extra rounding of the branch value, which is no larger
          // than half an ulp of the result it is added into.
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int col = c0 + 8 * h;
            const float4 ba = __ldg(reinterpret_cast<const float4*>(p.b2 + col));
            const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b2 + col + 4));
            float f[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
            for (int u = 0; u < 8; ++u) f[u] += __uint_as_float(v[8 * h + u]);
            uint4 o = pack8(f);
            const uint4 xr = h ? x1 : x0;
            __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
            const __nv_bfloat162* x2 = reinterpret_cast<const __nv_bfloat162*>(&xr);
#pragma unroll
            for (int u = 0; u < 4; ++u) o2[u] = __hadd2(o2[u], x2[u]);
            sts128(box_a + (h ? off1 : off0), o);
          }
          return;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int col = c0 + 8 * h;
          const float4 ba = __ldg(reinterpret_cast<const float4*>(p.b2 + col));
          const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b2 + col + 4));
          float f[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
          float xr[8];
          unpack8(h ? x1 : x0, xr);
#pragma unroll
          for (int u = 0; u < 8; ++u) f[u] += __uint_as_float(v[8 * h + u]) + xr[u];
          if (p.accumulate) {                                    // (2 of the 9 units of a stage: loaded in place)
            float o[8];
            unpack8(lds128(box_b + (h ? off1 : off0)), o);
#pragma unroll
            for (int u = 0; u < 8; ++u) f[u] += o[u];
          }
          if (p.out_scale != 1.f) {                              // only the last unit of a stage carries the MRF 1/3
#pragma unroll
            for (int u = 0; u < 8; ++u) f[u] *= p.out_scale;
          }
          sts128(box_a + (h ? off1 : off0), pack8(f));
          if (p.has_y2) {
#pragma unroll
            for (int u = 0; u < 8; ++u) f[u] = fmaxf(f[u], f[u] * p.act2_slope);   // leaky-relu, 0 < slope <= 1
            sts128(box_b + (h ? off1 : off0), pack8(f));
          }
        }
      };
      for (int c0 = c_begin; c0 < c_end; c0 += 32) {           // two TMEM loads in flight per wait
        uint32_t va[16], vb[16];
        tmem_ld16_nowait(taddr + (uint32_t)c0, va);
        if (c0 + 16 < c_end) tmem_ld16_nowait(taddr + (uint32_t)(c0 + 16), vb);
        tmem_ld_wait();
        emit16(va, c0);
        if (c0 + 16 < c_end) emit16(vb, c0 + 16);
      }
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) arrive_mma(&acc2_empty[a]);
      if (two) asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");   // both column halves of the box are written
      if (leader) {
        if (rows_q > 0 && valid) {
          const uint8_t* src_a = sm_sa + slot * p.stage_box_bytes + q * box_bytes;
          const uint8_t* src_b = sm_sb + slot * p.stage_box_bytes + q * box_bytes;
          tma_store_3d(q < 3 ? &map_y : &map_yt, src_a, 0, t0 + q * 32, b);
          if (p.has_y2) tma_store_3d(q < 3 ? &map_y2 : &map_y2t, src_b, 0, t0 + q * 32, b);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (p.slots == 1) {
          // single staging slot: refill it as soon as this tile's stores have been read out
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          if (i + 1 < n_my) prefetch(tn, 0);
        }
      }
      slot = next_slot;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    tc_fence_before();
  }
  if (PAIR) cluster_sync_all();
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

struct RuPlan {
  RuArgs a;
  int smem_bytes, ctas_per_sm, pair;
};

int plan_resunit_mode(int c, int k, int dil, int accumulate, int has_y2, int pair, RuPlan* out);

// single CTAs when the weights fit one SM's shared memory, CTA pairs (output channels split) when only half of them do
int plan_resunit(int c, int k, int dil, int accumulate, int has_y2, RuPlan* out) {
  static const int force_pair = getenv("SIB_RU_PAIR") ? atoi(getenv("SIB_RU_PAIR")) : -1;   // 0 never, 1 whenever legal
  if (force_pair == 1 && c >= 32 && plan_resunit_mode(c, k, dil, accumulate, has_y2, 1, out) == SIB_OK) return SIB_OK;
  static const bool wide_on = !(getenv("SIB_RU_WIDE") && atoi(getenv("SIB_RU_WIDE")) == 0);   // A/B switch for the C = 128 units
  if (c == 128) return wide_on ? plan_resunit_mode(c, k, dil, accumulate, has_y2, 1, out) : SIB_ERR_UNSUPPORTED;
  if (plan_resunit_mode(c, k, dil, accumulate, has_y2, 0, out) == SIB_OK) return SIB_OK;
  if (force_pair != 0 && c == 64) return plan_resunit_mode(c, k, dil, accumulate, has_y2, 1, out);
  return SIB_ERR_UNSUPPORTED;
}

int plan_resunit_mode(int c, int k, int dil, int accumulate, int has_y2, int pair, RuPlan* out) {
  const bool wide = c == 128;
  if (!(c == 128 || c == 64 || c == 32 || c == 16)) {
    sib::set_error("sib_resunit_bf16: c=%d unsupported (16 / 32 / 64 / 128; wider stages use sib_conv1d_bf16)", c);
    return SIB_ERR_UNSUPPORTED;
  }
  if (wide && (!pair || k > 3 || accumulate || has_y2)) {
    sib::set_error("sib_resunit_bf16: c=128 runs as CTA pairs with k <= 3 and neither a running sum nor a second output "
                   "(k=%d accumulate=%d y_act=%d): use sib_conv1d_bf16 for this unit", k, accumulate, has_y2);
    return SIB_ERR_UNSUPPORTED;
  }
  if (k < 1 || (k & 1) == 0 || k > 15 || dil < 1) {
    sib::set_error("sib_resunit_bf16: k=%d dilation=%d unsupported (odd k <= 15)", k, dil);
    return SIB_ERR_UNSUPPORTED;
  }
  RuArgs& a = out->a;
  memset(&a, 0, sizeof(a));
  a.C = c; a.k = k; a.dil = dil;
  a.p2 = (k - 1) / 2;
  a.p1 = (k - 1) * dil / 2;
  a.R = 128 - 2 * a.p2;
  a.xr = 128 + 2 * a.p1;
  a.tail_rows = a.R - 96;
  if (a.xr > 256 || a.tail_rows < 1) {
    sib::set_error("sib_resunit_bf16: k=%d dilation=%d needs a %d-row halo tile (max 256)", k, dil, a.xr);
    return SIB_ERR_UNSUPPORTED;
  }
  a.nch = wide ? 2 : 1;
  const int cc = c / a.nch;                          // channels per chunk = K-row of an operand slab
  a.row_bytes = cc * 2;
  a.ksteps = cc / 16;
  a.c_cta = pair ? c / 2 : c;
  a.tap_bytes = a.c_cta * a.row_bytes;               // one (tap, chunk) weight slab of this CTA
  out->pair = pair;
  a.x_chunk_bytes = (a.xr * a.row_bytes + 1023) / 1024 * 1024;
  a.x_stage_bytes = a.nch * a.x_chunk_bytes;
  a.t1_chunk_bytes = ((128 + k - 1) * a.row_bytes + 1023) / 1024 * 1024;
  a.t1_bytes = a.nch * a.t1_chunk_bytes;
  // weights: as few TMA boxes as possible (<= 32 KB each), no padding slabs when one box takes them all
  const int slabs = k * a.nch;                       // [chunk][tap] slabs per conv
  a.w_loads = (slabs * a.tap_bytes + 32767) / 32768;
  a.w_tg = (slabs + a.w_loads - 1) / a.w_loads;
  a.w_tx_bytes = a.w_loads * a.w_tg * a.tap_bytes;   // full boxes: slabs past the last are TMA zero fill but still counted
  a.w_bytes = (a.w_tx_bytes + 1023) / 1024 * 1024;
  a.stage_box_bytes = wide ? 16384 : 128 * a.row_bytes;   // wide: 4 warps x 2 slots x (32 rows x 64 B)
  a.need_b = (accumulate || has_y2) ? 1 : 0;
  a.accumulate = accumulate; a.has_y2 = has_y2;
  a.desc_hi = make_desc_hi(a.row_bytes);
  a.idesc = make_idesc_bf16(pair ? 256 : 128, c);
  // ring depths: prefer (4 x-tiles, 2 staging slots) inside the two-CTAs-per-SM budget, then the same inside one SM,
  // then shrink (3, 2 x-tiles; finally a single staging slot) until the resident weights fit
  const int fixed = 2 * a.w_bytes + 512 + 1024 + (wide ? 1024 : 0);   // (+ both bias vectors in shared memory when wide)
  auto need = [&](int nxs, int slots, int t1b) {
    return fixed + t1b * a.t1_bytes + nxs * a.x_stage_bytes + slots * a.stage_box_bytes * (1 + a.need_b);
  };
  const int two_cta = 115 * 1024 - 1024, one_cta = 227 * 1024;
  a.nxs = 0;
  // (three staging slots are supported by the kernel but measured 5-8 % slower than two: not offered)
  const int tries[9][3] = {{4, 2, 3}, {3, 2, 3}, {4, 2, 2}, {3, 2, 2}, {2, 2, 2}, {2, 2, 1}, {3, 1, 1}, {2, 1, 2}, {2, 1, 1}};
  if (wide) {                                        // no separate intermediate buffer: it is written over the tile's x slot
    if (need(3, 1, 0) <= one_cta) { a.nxs = 3; a.slots = 1; a.t1_bufs = 0; }
    else if (need(2, 1, 0) <= one_cta) { a.nxs = 2; a.slots = 1; a.t1_bufs = 0; }
  } else
  for (int pass = pair ? 1 : 0; pass < 2 && a.nxs == 0; ++pass)
    for (const auto& tr : tries)
      if (need(tr[0], tr[1], tr[2]) <= (pass == 0 ? two_cta : one_cta) && (pass == 1 || tr[1] >= 2)) {
        a.nxs = tr[0]; a.slots = tr[1]; a.t1_bufs = tr[2];
        break;
      }
  if (a.nxs == 0) {
    sib::set_error("sib_resunit_bf16: c=%d k=%d dilation=%d needs %d bytes of shared memory (weights must stay resident)",
                   c, k, dil, need(2, 1, 1));
    return SIB_ERR_UNSUPPORTED;
  }
  out->smem_bytes = need(a.nxs, a.slots, a.t1_bufs);
  // lookahead of conv1 over conv2: bounded by the intermediate ring and the x ring; acc1 ring = la + 1 TMEM buffers
  static const int force_la = getenv("SIB_RU_LA") ? atoi(getenv("SIB_RU_LA")) : 0;
  a.la = a.t1_bufs >= 3 ? 2 : 1;
  if (force_la >= 1 && force_la <= 3) a.la = force_la;
  if (a.la > a.nxs - 1) a.la = a.nxs - 1 > 0 ? a.nxs - 1 : 1;
  a.na1 = a.la + 1;
  {
    uint32_t cols = (uint32_t)((a.na1 + 2) * c), pw = 32;
    while (pw < cols) pw <<= 1;
    a.tmem_cols = pw;
  }
  static const int force_e2w = getenv("SIB_RU_EPI2_WARPS") ? atoi(getenv("SIB_RU_EPI2_WARPS")) : 0;   // A/B switch: 4 or 8
  // measured per configuration (standalone, same box, 4 -> 8 warps): C = 64 k = 3 0.121 -> 0.106 ms, C = 32 k = 11 0.226 -> 0.217;
  // C = 32 k = 3 / 7 and C = 64 k = 7 within +-3 %; the C = 64 k = 11 pair units 0.240 -> 0.273 (their 23 KB tiles need the six
  // activation warps) - so eight only where it pays
  const bool e2w8 = !wide && ((c == 64 && !pair && k <= 3) || (c == 32 && k >= 11));
  a.e2w = force_e2w == 4 ? 4 : (force_e2w == 8 && !wide && c >= 32 ? 8 : (e2w8 ? 8 : 4));
  out->ctas_per_sm = (!pair && out->smem_bytes <= 115 * 1024 - 1024) ? 2 : 1;
  return SIB_OK;
}

}  // namespace

extern "C" int sib_resunit_bf16_supported(int c, int k, int dilation, int accumulate, int has_y_act) {
  RuPlan pl;
  return plan_resunit(c, k, dilation, accumulate, has_y_act, &pl) == SIB_OK ? 1 : 0;
}

extern "C" int sib_resunit_bf16(const sib_resunit_desc* d, const void* x, const void* w1, const float* b1, const void* w2,
                                const float* b2, void* y, void* y_act, sib_stream_t stream) {
  SIB_REQUIRE(d && x && w1 && b1 && w2 && b2 && y, "sib_resunit_bf16: null argument");
  SIB_REQUIRE(d->batch > 0 && d->t > 0, "sib_resunit_bf16: empty shape");
  SIB_REQUIRE(x != y && x != y_act, "sib_resunit_bf16: in-place operation is not supported (tiles read x halos)");
  RuPlan pl;
  if (int rc = plan_resunit(d->c, d->k, d->dilation, d->accumulate, y_act != nullptr, &pl)) return rc;
  RuArgs& a = pl.a;
  auto al16 = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
  SIB_REQUIRE(al16(x) && al16(y) && al16(w1) && al16(w2) && al16(b1) && al16(b2) && (!y_act || al16(y_act)),
              "sib_resunit_bf16: pointers must be 16-byte aligned");
  SIB_REQUIRE(d->x_row_stride % 8 == 0 && d->x_batch_stride % 8 == 0 && d->y_row_stride % 8 == 0 && d->y_batch_stride % 8 == 0,
              "sib_resunit_bf16: strides must be multiples of 8 elements");
  SIB_REQUIRE(d->slope_in > 0.f && d->slope_in <= 1.f && d->slope_mid > 0.f && d->slope_mid <= 1.f,
              "sib_resunit_bf16: leaky-relu slopes must be in (0, 1] (max(x, slope x) form)");
  a.b1 = b1; a.b2 = b2;
  a.T = d->t; a.batch = d->batch;
  a.xg = x; a.xg_batch_stride = d->x_batch_stride; a.xg_row_stride = d->x_row_stride;
  static const bool early_w = !(getenv("SIB_PDL_EARLY_W") && atoi(getenv("SIB_PDL_EARLY_W")) == 0);
  a.early_w = early_w ? 1 : 0;
  a.slope_in = d->slope_in; a.slope_mid = d->slope_mid; a.out_scale = d->out_scale; a.act2_slope = d->act2_slope;
  a.tiles_m = sib::ceil_div(d->t, a.R);
  const int64_t total = (int64_t)a.tiles_m * d->batch;
  SIB_REQUIRE(total < (1ll << 31), "sib_resunit_bf16: too many tiles");
  a.total_tiles = (int)total;
  const CUtensorMapSwizzle swz = a.row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                                    : (a.row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  const char* who = "sib_resunit_bf16";
  CUtensorMap map_x, map_w1, map_w2, map_res, map_y, map_yt, map_y2, map_y2t;
  {
    const cuuint64_t dims[3] = {(cuuint64_t)d->c, (cuuint64_t)d->t, (cuuint64_t)d->batch};
    const cuuint64_t xs[3] = {2, (cuuint64_t)d->x_row_stride * 2, (cuuint64_t)d->x_batch_stride * 2};
    const cuuint64_t ys[3] = {2, (cuuint64_t)d->y_row_stride * 2, (cuuint64_t)d->y_batch_stride * 2};
    const cuuint32_t ccx = (cuuint32_t)(d->c / a.nch);          // channels per operand slab
    const cuuint32_t cy = a.nch == 2 ? 32u : (cuuint32_t)d->c;  // wide: output boxes of 32 channels (64-byte swizzle)
    const CUtensorMapSwizzle swz_y = a.nch == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : swz;
    const cuuint32_t box_x[3] = {ccx, (cuuint32_t)a.xr, 1};
    const cuuint32_t box_r[3] = {ccx, 32, 1};
    const cuuint32_t box_q[3] = {cy, 32, 1};
    const cuuint32_t box_t[3] = {cy, (cuuint32_t)a.tail_rows, 1};
    if (int rc = encode_map(&map_x, x, 3, dims, xs, box_x, swz, who, "x")) return rc;
    if (int rc = encode_map(&map_res, x, 3, dims, xs, box_r, swz, who, "x (residual)")) return rc;
    if (int rc = encode_map(&map_y, y, 3, dims, ys, box_q, swz_y, who, "y")) return rc;
    if (int rc = encode_map(&map_yt, y, 3, dims, ys, box_t, swz_y, who, "y (tail)")) return rc;
    if (int rc = encode_map(&map_y2, y_act ? y_act : y, 3, dims, ys, box_q, swz_y, who, "y_act")) return rc;
    if (int rc = encode_map(&map_y2t, y_act ? y_act : y, 3, dims, ys, box_t, swz_y, who, "y_act (tail)")) return rc;
  }
  {
    // weights in the sib_conv1d_bf16 layout [1][chunks][k][c_out][cc]: one K-major slab per (chunk, tap)
    const cuuint64_t ccw = (cuuint64_t)(d->c / a.nch);
    const cuuint64_t dims[3] = {ccw, (cuuint64_t)d->c, (cuuint64_t)d->k * a.nch};
    const cuuint64_t ws[3] = {2, ccw * 2, (cuuint64_t)d->c * ccw * 2};
    const cuuint32_t box[3] = {(cuuint32_t)ccw, (cuuint32_t)a.c_cta, (cuuint32_t)a.w_tg};
    if (int rc = encode_map(&map_w1, w1, 3, dims, ws, box, swz, who, "w1")) return rc;
    if (int rc = encode_map(&map_w2, w2, 3, dims, ws, box, swz, who, "w2")) return rc;
  }
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    const void* fns[3] = {(const void*)resunit_tc_kernel<false>, (const void*)resunit_tc_kernel<true>,
                          (const void*)resunit_tc_kernel<true, true>};
    for (const void* fn : fns) {
      cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) {
        sib::set_error("sib_resunit_bf16: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return SIB_ERR_CUDA;
      }
    }
    attr_set[dev] = true;
  }
  int grid;
  if (pl.pair) {
    const int clusters = sm_count_of_current_device() / 2, units = (a.total_tiles + 1) / 2;
    grid = 2 * (units < clusters ? units : clusters);
  } else {
    const int slots = sm_count_of_current_device() * pl.ctas_per_sm;
    grid = a.total_tiles < slots ? a.total_tiles : slots;
  }
  const RuArgs args = a;
  static const bool verbose = getenv("SIB_TC_VERBOSE") != nullptr;
  if (verbose)
    fprintf(stderr, "[sib_resunit_bf16] B%d T%d C%d k%d d%d acc=%d y2=%d: pair=%d R=%d xr=%d nxs=%d slots=%d t1=%d la=%d smem=%d ctas/sm=%d tiles=%d\n",
            d->batch, d->t, d->c, d->k, d->dilation, a.accumulate, a.has_y2, pl.pair, a.R, a.xr, a.nxs, a.slots, a.t1_bufs, a.la, pl.smem_bytes,
            pl.ctas_per_sm, a.total_tiles);
  const cudaError_t le =
      a.nch == 2 ? sib::launch_pdl_cluster(resunit_tc_kernel<true, true>, dim3(grid), dim3(NUM_THREADS_WIDE), (size_t)pl.smem_bytes,
                                           static_cast<cudaStream_t>(stream), 2u, map_x, map_w1, map_w2, map_res, map_y, map_yt,
                                           map_y2, map_y2t, args)
      : pl.pair ? sib::launch_pdl_cluster(resunit_tc_kernel<true>, dim3(grid), dim3(NUM_THREADS), (size_t)pl.smem_bytes,
                                        static_cast<cudaStream_t>(stream), 2u, map_x, map_w1, map_w2, map_res, map_y, map_yt,
                                        map_y2, map_y2t, args)
              : sib::launch_pdl(resunit_tc_kernel<false>, dim3(grid), dim3(NUM_THREADS), (size_t)pl.smem_bytes,
                                static_cast<cudaStream_t>(stream), map_x, map_w1, map_w2, map_res, map_y, map_yt, map_y2,
                                map_y2t, args);
  if (le != cudaSuccess) {
    sib::set_error("sib_resunit_bf16: launch failed: %s", cudaGetErrorString(le));
    return SIB_ERR_CUDA;
  }
  SIB_CHECK_LAUNCH("sib_resunit_bf16");
  return SIB_OK;
}
