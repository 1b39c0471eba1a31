// K3f: layer-fused MLP trunk forward  x -> ELU(L1, 512) -> ELU(L2, 256) -> ELU(L3, 128) [-> scalar head]
// for up to four network instances per launch.  One CTA owns one 128-row tile of one network
// and keeps every activation ON CHIP:
//
//   * the input tile and the weight tiles are staged by TMA (128-byte swizzle, mbarrier ring);
//   * layer 1 runs as four 128-column quarters into two ping-pong TMEM regions P0/P1;
//   * eight epilogue warps convert a finished quarter IN PLACE (tcgen05.ld -> +bias -> ELU ->
//     RN to TF32 -> tcgen05.st); the converted quarter is then the A operand of layer 2
//     straight from tensor memory (tcgen05.mma with A in TMEM), accumulated into Y (256 columns)
//     while the next quarter of layer 1 is already running on the tensor core;
//   * Y is converted in place the same way and feeds layer 3 (accumulator Z reuses P0);
//   * the last epilogue applies bias + ELU and, for the twin-Q critic, the scalar head as a
//     fused fp32 dot product; h1/h2/h3 are written to HBM (TMA stores through swizzled staging
//     chunks) only for the networks whose backward pass needs them.
//
// TMEM columns: P0 [0,128) P1 [128,256) Y [256,512); Z = P0.  Weights stream from L2 once per
// tile; activations never leave the SM unless stored for backward.
// Replaces: the nn.Linear + nn.ELU launches of pql/models/mlp.py:15-24 for actor and critics
// (pql/algo/pql_v_learner.py:81-107, pql/algo/pql_p_learner.py:55-56).
#include "tcgen05_utils.cuh"

namespace pqlb {

constexpr int kFH1 = 512, kFH2 = 256, kFH3 = 128;
constexpr int kFEpiWarps = 16;             // four per TMEM lane quarter: one 32-column chunk each per 128-column region
constexpr int kFThreads = 64 + 32 * kFEpiWarps;
constexpr int kFRingBytes = 5 * 16384;        // weight ring; the rest of shared memory is X + 16 store-staging chunks
constexpr int kFTileBytes = 128 * 128;        // every weight tile: 128 rows x 32 k (W2 goes as two row halves)
constexpr int kFXKb = 4;                      // input width <= 128
constexpr int kFXBytes = kFXKb * 128 * 128;
constexpr int kFChunk = 32 * 128;
constexpr int kFSmem = 1024 + kFXBytes + kFRingBytes + kFEpiWarps * kFChunk + (kFH1 + kFH2 + 2 * kFH3) * 4;

struct alignas(64) MlpGroupDev {
  CUtensorMap tmX, tmW1, tmW2, tmW3, tmH1, tmH2, tmH3, tmW4;
  const float* b1; const float* b2; const float* b3; const float* head_w; const float* head_b;
  float* q;
  // policy head (tanh): a fourth contraction [128 x 128] . W4^T -> [128 x act_n], act_n <= 16
  const float* act_b; const float* act_noise; float* act_out; float* act_out2;
  long long act_ldo, act_ldo2, act_ldnoise;
  float noise_std, noise_bound;
  int act_n;
  int st1, st2, st3;
};
struct alignas(64) MlpDev {
  MlpGroupDev g[PQLB_MAX_GROUPS];
  int M, k_in, kb1, pad;
  unsigned long long* dbg;     // optional timeline of CTA (0,0): clock64 stamps (PQLB_MLP_DEBUG)
};

__host__ __device__ constexpr uint32_t idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
}

// PAIR = false: one CTA per 128-row tile.  PAIR = true: two CTAs of a (2,1,1) cluster (row tiles
// 2i, 2i + 1 of one network) run every contraction as one M = 256 cta_group::2 MMA issued by the
// leader; each CTA streams only HALF of every weight tile (64 of its 128 rows), which halves the
// L2 -> SM traffic that bounds the one-CTA version (~40 B/clk per SM, measured).
template <bool PAIR>
__global__ void __launch_bounds__(kFThreads, 1)
mlp_fwd_kernel(const __grid_constant__ MlpDev P) {
  // Ring stage = 16 KB either way.  One CTA per tile: every weight tile is 128 rows x 32 k.  Pair:
  // layer 2 runs as N = 256 MMAs (the A operand is read from tensor memory once per k-step instead
  // of once per 128-column half: TMEM delivers only 64 B/clk to tcgen05.ld and A-operand reads
  // together), so a W2 tile is 256 rows x 32 k and each CTA holds 128 of them (16 KB); W1 / W3
  // tiles are 128 rows, 64 per CTA (8 KB of the stage).
  constexpr int kFStageBytes = kFTileBytes;
  constexpr int kFStages = kFRingBytes / kFStageBytes;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t x_full, full_bar[kFStages], empty_bar[kFStages];
  __shared__ __align__(8) uint64_t p_full[2], p_conv[2], y_full, y_conv[2], z_full, z_conv, a_full;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_q[4][128];

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const MlpGroupDev& G = P.g[blockIdx.y];
  const int m0 = blockIdx.x * 128;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t xs = base;
  const uint32_t ring = xs + kFXBytes;
  const uint32_t stage_buf = ring + kFStages * kFStageBytes;
  float* s_vec = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)) + kFXBytes + kFRingBytes + kFEpiWarps * kFChunk);
  float* s_b1 = s_vec; float* s_b2 = s_b1 + kFH1; float* s_b3 = s_b2 + kFH2; float* s_w4 = s_b3 + kFH3;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&x_full), 1);
    // "converted" barriers live in the leader and collect the epilogue warps of BOTH CTAs of a pair
    constexpr uint32_t kConv = PAIR ? 2 * kFEpiWarps : kFEpiWarps;
    for (int s = 0; s < kFStages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&p_full[b]), 1); mbar_init(smem_u32(&p_conv[b]), kConv); mbar_init(smem_u32(&y_conv[b]), kConv); }
    mbar_init(smem_u32(&y_full), 1); mbar_init(smem_u32(&z_full), 1);
    mbar_init(smem_u32(&z_conv), kFEpiWarps); mbar_init(smem_u32(&a_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int j = threadIdx.x; j < kFH1; j += kFThreads) s_b1[j] = G.b1[j];
  for (int j = threadIdx.x; j < kFH2; j += kFThreads) s_b2[j] = G.b2[j];
  for (int j = threadIdx.x; j < kFH3; j += kFThreads) { s_b3[j] = G.b3[j]; s_w4[j] = G.q ? G.head_w[j] : 0.f; }
  __syncthreads();
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  if (PAIR) cluster_sync_all();             // both CTAs' barriers are initialised before any remote arrive / TMA completion
  if (warp == 1) {                          // one warp of EACH CTA of a pair takes part in the cta_group::2 allocation
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();             // the leader issues MMAs into the peer's tensor memory: both allocations are done
  tcgen05_fence_after();
  const uint32_t tmem = uniform_u32(tmem_slot);
  unsigned long long* dbg = (blockIdx.x == 0 && blockIdx.y == 0) ? P.dbg : nullptr;
  int dbg_i = 0;
#define PQLB_STAMP(base) do { if (dbg && lane == 0) dbg[(base) + dbg_i] = clock64(); ++dbg_i; } while (0)
  const uint32_t tY = tmem + 256u;
  const uint32_t tZ = tmem;

  // Weight-tile schedule shared by producer and MMA issuer, one hex digit per phase (phase 0 is
  // the lowest digit).  kind 0 = layer-1 quarter q, 1 = layer-2 K-chunk c, 2 = layer 3; the loops
  // over it are fully unrolled, so every branch below folds to straight-line code.
  constexpr unsigned long long kKinds = 0x211010100ull, kArgs = 0x032312010ull;
  // M = 256 for a pair (field M >> 4 at bit 24), N = 128 either way
  constexpr uint32_t idesc = PAIR ? (idesc_tf32(128) & ~(0x1Fu << 24)) | ((256u >> 4) << 24) : idesc_tf32(128);
  constexpr uint32_t idesc_n256 = (idesc_tf32(256) & ~(0x1Fu << 24)) | ((256u >> 4) << 24);       // pair, layer 2
  const uint64_t desc0 = make_smem_desc(0, 16, 1024, kLayoutSw128);
  // barrier of the LEADER at the same offset (identity for the leader itself / without pairs)
  auto leader = [&](uint32_t local_bar) -> uint32_t { return PAIR ? mapa_shared(local_bar, 0) : local_bar; };
  auto mma_ss = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t acc) { if (PAIR) umma_tf32_pair(d, a, b, idesc, acc); else umma_tf32(d, a, b, idesc, acc); };
  auto mma_ts = [&](uint32_t d, uint32_t a, uint64_t b, uint32_t acc) { if (PAIR) umma_tf32_ts_pair(d, a, b, idesc, acc); else umma_tf32_ts(d, a, b, idesc, acc); };
  auto commit = [&](uint32_t bar) { if (PAIR) umma_commit_pair(bar); else umma_commit(bar); };

  // Warps 0 and 1 run with all 32 lanes and issue from one elected lane (see elect_one()).
  if (warp == 0) {
    // ===================== TMA producer =====================
    const uint32_t xb = leader(smem_u32(&x_full));
    if (elect_one()) {
      // the leader's barrier counts the X tiles of both CTAs and both halves of every weight tile
      if (rank == 0) mbar_expect_tx(smem_u32(&x_full), (uint32_t)P.kb1 * 16384u * (PAIR ? 2u : 1u));
      for (int kb = 0; kb < P.kb1; ++kb) {
        if (PAIR) tma_load_2d_pair(xs + kb * 16384, &G.tmX, kb * 32, m0, xb);
        else tma_load_2d(xs + kb * 16384, &G.tmX, kb * 32, m0, xb);
      }
    }
    __syncwarp();
    int stage = 0; uint32_t phase = 0;
    const int r0 = PAIR ? (int)rank * 64 : 0;       // this CTA's rows of every 128-row weight tile (W1, W3)
#pragma unroll
    for (int ph = 0; ph < 9; ++ph) {
      const int kind = (int)((kKinds >> (4 * ph)) & 15), arg = (int)((kArgs >> (4 * ph)) & 15);
      const int n_tiles = kind == 0 ? P.kb1 : (kind == 1 && PAIR ? 4 : 8);
#pragma unroll 1
      for (int t = 0; t < n_tiles; ++t) {
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
        const uint32_t bar = leader(smem_u32(&full_bar[stage]));
        const uint32_t dst = ring + stage * kFStageBytes;
        if (elect_one()) {
          // the leader's barrier counts what BOTH CTAs of a pair load
          if (rank == 0) mbar_expect_tx(smem_u32(&full_bar[stage]), (uint32_t)(PAIR && kind == 1 ? 2 * kFTileBytes : kFTileBytes));
          const CUtensorMap* map; int c0, c1;
          if (kind == 0) { map = &G.tmW1; c0 = t * 32; c1 = arg * 128 + r0; }
          else if (kind == 1) {
            map = &G.tmW2;
            if (PAIR) { c0 = arg * 128 + t * 32; c1 = (int)rank * 128; }          // 128 of the 256 rows of W2
            else { c0 = arg * 128 + (t >> 1) * 32; c1 = (t & 1) * 128; }
          }
          else { map = &G.tmW3; c0 = t * 32; c1 = r0; }
          if (PAIR) tma_load_2d_pair(dst, map, c0, c1, bar);
          else tma_load_2d(dst, map, c0, c1, bar);
        }
        __syncwarp();
        if (++stage == kFStages) { stage = 0; phase ^= 1u; }
      }
    }
    if (!PAIR && G.act_n > 0) {                 // policy head weights: 16 rows x 128 k as four 2 KB boxes in one stage
      mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
      const uint32_t bar = smem_u32(&full_bar[stage]);
      const uint32_t dst = ring + stage * kFStageBytes;
      if (elect_one()) {
        mbar_expect_tx(bar, 4u * 2048u);
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(dst + kb * 2048, &G.tmW4, kb * 32, 0, bar);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (the leader CTA of a pair only) =====================
    if (rank == 0) {
    auto free_stage = [&](uint32_t bar) { commit(bar); };
    PQLB_STAMP(0);
    mbar_wait(smem_u32(&x_full), 0);
    tcgen05_fence_after();
    PQLB_STAMP(0);
    int stage = 0; uint32_t phase = 0;
#pragma unroll
    for (int ph = 0; ph < 9; ++ph) {
      const int kind = (int)((kKinds >> (4 * ph)) & 15), arg = (int)((kArgs >> (4 * ph)) & 15);
      if (kind == 0) {
        const uint32_t tP = tmem + (uint32_t)((arg & 1) * 128);
#pragma unroll 1
        for (int kb = 0; kb < P.kb1; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tcgen05_fence_after();
          const int krem = P.k_in - kb * 32;
          const uint64_t adesc = desc0 | (uint64_t)(((xs + kb * 16384) >> 4) & 0x3FFF);
          const uint64_t bdesc = desc0 | (uint64_t)(((ring + stage * kFStageBytes) >> 4) & 0x3FFF);
          if (elect_one()) {
            if (krem >= 32) {
#pragma unroll
              for (int k = 0; k < 4; ++k) mma_ss(tP, adesc + 2u * k, bdesc + 2u * k, (uint32_t)(kb | k) != 0u);
            } else {
              const int ksteps = (krem + 7) / 8;
              for (int k = 0; k < ksteps; ++k) mma_ss(tP, adesc + 2u * k, bdesc + 2u * k, (uint32_t)(kb | k) != 0u);
            }
            free_stage(smem_u32(&empty_bar[stage]));
          }
          __syncwarp();
          if (++stage == kFStages) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) commit(smem_u32(&p_full[arg & 1]));
        __syncwarp();
        PQLB_STAMP(0);
      } else if (kind == 1) {
        const int c = arg, b = c & 1;
        const uint32_t tP = tmem + (uint32_t)(b * 128);
        mbar_wait(smem_u32(&p_conv[b]), (uint32_t)(c >> 1) & 1u);      // quarter c converted in place
        tcgen05_fence_after();
        PQLB_STAMP(0);
        if (PAIR) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {       // k-block kk of the chunk, all 256 rows of W2 in one N = 256 MMA
            mbar_wait(smem_u32(&full_bar[stage]), phase);
            tcgen05_fence_after();
            const uint64_t bdesc = desc0 | (uint64_t)(((ring + stage * kFStageBytes) >> 4) & 0x3FFF);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_tf32_ts_pair(tY, tP + (uint32_t)(kk * 32 + k * 8), bdesc + 2u * k, idesc_n256, (uint32_t)((c | kk | k) != 0));
              free_stage(smem_u32(&empty_bar[stage]));
            }
            __syncwarp();
            if (++stage == kFStages) { stage = 0; phase ^= 1u; }
          }
        } else {
#pragma unroll
          for (int t = 0; t < 8; ++t) {          // (k-block t>>1 of the chunk) x (row half t&1 of W2)
            mbar_wait(smem_u32(&full_bar[stage]), phase);
            tcgen05_fence_after();
            const uint64_t bdesc = desc0 | (uint64_t)(((ring + stage * kFStageBytes) >> 4) & 0x3FFF);
            const int kk = t >> 1;
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                mma_ts(tY + (uint32_t)((t & 1) * 128), tP + (uint32_t)(kk * 32 + k * 8), bdesc + 2u * k,
                       (uint32_t)((c | kk | k) != 0));
              free_stage(smem_u32(&empty_bar[stage]));
            }
            __syncwarp();
            if (++stage == kFStages) { stage = 0; phase ^= 1u; }
          }
        }
        if (c == 3) { if (elect_one()) commit(smem_u32(&y_full)); __syncwarp(); }
        PQLB_STAMP(0);
      } else {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          if ((t & 3) == 0) { mbar_wait(smem_u32(&y_conv[t >> 2]), 0); tcgen05_fence_after(); }
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tcgen05_fence_after();
          const uint64_t bdesc = desc0 | (uint64_t)(((ring + stage * kFStageBytes) >> 4) & 0x3FFF);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma_ts(tZ, tY + (uint32_t)(t * 32 + k * 8), bdesc + 2u * k, (uint32_t)((t | k) != 0));
            free_stage(smem_u32(&empty_bar[stage]));
          }
          __syncwarp();
          if (++stage == kFStages) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) commit(smem_u32(&z_full));
        __syncwarp();
        PQLB_STAMP(0);
      }
    }
    if (!PAIR && G.act_n > 0) {
      // policy head: h3 (converted in place in Z) . W4^T into 16 columns of P1, N = 16
      mbar_wait(smem_u32(&z_conv), 0);
      tcgen05_fence_after();
      mbar_wait(smem_u32(&full_bar[stage]), phase);
      tcgen05_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          const uint64_t bdesc = desc0 | (uint64_t)(((ring + stage * kFStageBytes + kb * 2048) >> 4) & 0x3FFF);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_tf32_ts(tmem + 128u, tZ + (uint32_t)(kb * 32 + k * 8), bdesc + 2u * k, idesc_tf32(16), (uint32_t)((kb | k) != 0));
        }
        umma_commit(smem_u32(&a_full));
      }
      __syncwarp();
    }
    }   // rank == 0
  } else {
    // ===================== conversion / epilogue warps =====================
    // Warp e owns TMEM lanes [32 (warp % 4), +32) (hardware rule) and the 32-column chunk e >> 2
    // of every 128-column region.
    const int e = warp - 2;
    const int quarter = warp & 3;
    const int chunk = e >> 2;
    const int col = chunk * 32;
    const int row0 = m0 + quarter * 32;
    const int row = row0 + lane;
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    const uint32_t my_stage = stage_buf + e * kFChunk;
    const uint32_t swz = (uint32_t)(lane & 7) << 4;
    const uint32_t row_off = (uint32_t)lane * 128u;
    bool pending = false;

    // hands this warp's 32x32 chunk (already TF32-rounded) to a TMA store through its staging buffer
    auto store_chunk = [&](const float* v, const CUtensorMap* omap, int n_col) {
      if (pending) { if (elect_one()) bulk_wait_read<0>(); __syncwarp(); }
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4)
        sts128(my_stage + row_off + (((uint32_t)j4 << 4) ^ swz), v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
      fence_proxy_async();
      __syncwarp();
      if (elect_one()) { tma_store_3d(omap, my_stage, n_col, row0, 0); bulk_commit(); }
      pending = true;
    };
    // converts this warp's chunk of a 128-column TMEM region in place (+bias, ELU, RN to TF32), tells
    // the MMA issuer (``done``, a barrier of the leader) and only then walks the HBM store path, so
    // that the next contraction never waits for staging / TMA stores
    auto convert = [&](uint32_t region, const float* bias, int n_base, const CUtensorMap* omap, bool store, uint32_t done) {
      const uint32_t taddr = region + lane_sel + (uint32_t)col;
      float v[32];
      tmem_ld32(taddr, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = rn_tf32(elu_fast(v[j] + bias[n_base + col + j]));
      tmem_st32(taddr, v);
      tmem_wait_st();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) { if (PAIR) mbar_arrive_cluster(leader(done)); else mbar_arrive(done); }
      if (store) store_chunk(v, omap, n_base + col);
    };

    if (e != 0) dbg = nullptr;
    PQLB_STAMP(32);
    for (int q = 0; q < 4; ++q) {
      const int b = q & 1;
      mbar_wait(smem_u32(&p_full[b]), (uint32_t)(q >> 1) & 1u);
      tcgen05_fence_after();
      PQLB_STAMP(32);
      convert(tmem + (uint32_t)(b * 128), s_b1, q * 128, &G.tmH1, G.st1 != 0, smem_u32(&p_conv[b]));
      PQLB_STAMP(32);
    }
    mbar_wait(smem_u32(&y_full), 0);
    tcgen05_fence_after();
    PQLB_STAMP(32);
    for (int hh = 0; hh < 2; ++hh) {
      convert(tY + (uint32_t)(hh * 128), s_b2, hh * 128, &G.tmH2, G.st2 != 0, smem_u32(&y_conv[hh]));
      PQLB_STAMP(32);
    }
    mbar_wait(smem_u32(&z_full), 0);
    tcgen05_fence_after();
    PQLB_STAMP(32);
    {
      float qacc = 0.f;
      float v[32];
      tmem_ld32(tZ + lane_sel + (uint32_t)col, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float h = elu_fast(v[j] + s_b3[col + j]);
        qacc = fmaf(h, s_w4[col + j], qacc);
        v[j] = rn_tf32(h);
      }
      if (!PAIR && G.act_n > 0) {                // h3 back into Z: the A operand of the policy-head contraction
        tmem_st32(tZ + lane_sel + (uint32_t)col, v);
        tmem_wait_st();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&z_conv));
      }
      if (G.st3) store_chunk(v, &G.tmH3, col);
      if (!PAIR && G.act_n > 0 && chunk == 0) {
        // one warp per lane quarter finishes the head: + bias, tanh, (+ clipped noise, clamp), TF32 rounding
        mbar_wait(smem_u32(&a_full), 0);
        tcgen05_fence_after();
        float a[16];
        tmem_ld16(tmem + 128u + lane_sel, a);
        if (row < P.M) {
          float o[16], o2[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float t = 0.f, r = 0.f;
            if (j < G.act_n) {
              t = tanhf(a[j] + G.act_b[j]);
              r = t;
              if (G.act_noise) {                   // noise.py:19-27 on N(0,1) draws scaled by std (see pqlb_gemm_tf32)
                const float z = fminf(fmaxf(G.act_noise[(long long)row * G.act_ldnoise + j] * G.noise_std, -G.noise_bound), G.noise_bound);
                r = fminf(fmaxf(t + z, -1.f), 1.f);
              }
              t = r;                               // act_out2: the same value before the operand rounding
              r = rn_tf32(r);
            }
            o[j] = r; o2[j] = t;
          }
          float* dst = G.act_out + (long long)row * G.act_ldo;
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            if (j < G.act_n) *reinterpret_cast<float4*>(dst + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
          if (G.act_out2) {
            float* dst2 = G.act_out2 + (long long)row * G.act_ldo2;
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              if (j < G.act_n) *reinterpret_cast<float4*>(dst2 + j) = make_float4(o2[j], o2[j + 1], o2[j + 2], o2[j + 3]);
          }
        }
      }
      if (G.q) {
        // the four warps of a lane quarter each hold the dot product over their 32 columns;
        // summed in chunk order (fixed, so q is reproducible)
        s_q[chunk][quarter * 32 + lane] = qacc;
        named_bar_sync(1 + quarter, 128);
        if (chunk == 0 && row < P.M) {
          const int r = quarter * 32 + lane;
          G.q[row] = (((s_q[0][r] + s_q[1][r]) + s_q[2][r]) + s_q[3][r]) + G.head_b[0];
        }
      }
    }
    if (pending) { if (elect_one()) bulk_wait_read<0>(); __syncwarp(); }
    PQLB_STAMP(32);
  }

  tcgen05_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();             // nobody frees tensor memory or leaves while the peer may still use / signal it
  if (warp == 1) {
    tcgen05_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

}  // namespace pqlb

using namespace pqlb;

static unsigned long long* g_mlp_debug = nullptr;
static int g_mlp_cluster = 0;
constexpr int kFDefaultCluster = 1;
/* 1 = one CTA per row tile, 2 = CTA pairs (cta_group::2), 0 = default. */
extern "C" void pqlb_mlp_forward_cluster(int c) { g_mlp_cluster = (c == 1 || c == 2) ? c : 0; }
/* Debug: device buffer of 64 uint64 receiving a clock64 timeline of CTA (0,0) (NULL = off). */
extern "C" void pqlb_mlp_forward_debug(unsigned long long* buf) { g_mlp_debug = buf; }

extern "C" int pqlb_mlp_forward_init(void) {
  cudaError_t e = cudaFuncSetAttribute(mlp_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFSmem);
  return e == cudaSuccess ? PQLB_OK : (int)e;
}

extern "C" int pqlb_mlp_forward(const pqlb_mlp_desc* d, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(d && d->M > 0 && d->k_in > 0 && d->n_groups >= 1 && d->n_groups <= PQLB_MAX_GROUPS);
  if (d->k_in > kFXKb * 32) return PQLB_E_UNSUPPORTED;      // wider inputs take the per-layer path
  { int rc = pqlb_init(); if (rc != PQLB_OK) return rc; }
  static MlpDev P;
  P.M = d->M; P.k_in = d->k_in; P.kb1 = (d->k_in + 31) / 32;
  P.dbg = g_mlp_debug;
  const int tiles_m = (d->M + 127) / 128;
  const int cluster = g_mlp_cluster > 0 ? g_mlp_cluster : kFDefaultCluster;
  for (int i = 0; i < d->n_groups; ++i) {
    const pqlb_mlp_group& s = d->g[i];
    MlpGroupDev& G = P.g[i];
    PQLB_CHECK_ARG(s.x && s.w1 && s.w2 && s.w3 && s.b1 && s.b2 && s.b3);
    PQLB_CHECK_ARG(!s.q || (s.head_w && s.head_b));
    int rc;
    if ((rc = make_map(&G.tmX, s.x, (uint64_t)d->k_in, (uint64_t)d->M, s.ldx, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B)) != PQLB_OK) return rc;
    const uint32_t wrows = 128u / (uint32_t)cluster;         // rows of a weight tile each CTA fetches (a pair: half each)
    if ((rc = make_map(&G.tmW1, s.w1, (uint64_t)d->k_in, kFH1, s.ldw1, 32, wrows, CU_TENSOR_MAP_SWIZZLE_128B)) != PQLB_OK) return rc;
    if ((rc = make_map(&G.tmW2, s.w2, kFH1, kFH2, kFH1, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B)) != PQLB_OK) return rc;   // pair: half of a 256-row tile
    if ((rc = make_map(&G.tmW3, s.w3, kFH2, kFH3, kFH2, 32, wrows, CU_TENSOR_MAP_SWIZZLE_128B)) != PQLB_OK) return rc;
    G.st1 = s.h1 != nullptr; G.st2 = s.h2 != nullptr; G.st3 = s.h3 != nullptr;
    if (G.st1 && !make_tile_map(&G.tmH1, s.h1, kFH1, (uint64_t)d->M, kFH1, 1, 0)) return PQLB_E_ALIGN;
    if (G.st2 && !make_tile_map(&G.tmH2, s.h2, kFH2, (uint64_t)d->M, kFH2, 1, 0)) return PQLB_E_ALIGN;
    if (G.st3 && !make_tile_map(&G.tmH3, s.h3, kFH3, (uint64_t)d->M, kFH3, 1, 0)) return PQLB_E_ALIGN;
    if (!G.st1) G.tmH1 = G.tmX;
    if (!G.st2) G.tmH2 = G.tmX;
    if (!G.st3) G.tmH3 = G.tmX;
    G.b1 = s.b1; G.b2 = s.b2; G.b3 = s.b3; G.head_w = s.head_w; G.head_b = s.head_b; G.q = s.q;
    G.act_n = 0; G.tmW4 = G.tmX;
    if (s.act_w) {
      // policy head: act_n a multiple of 4 up to 16, 16-byte aligned output rows (else: pqlb_gemm_tf32)
      PQLB_CHECK_ARG(!s.q && s.act_b && s.act_out && s.act_n > 0);
      if (s.act_n > 16 || s.act_n % 4 || s.act_ldo % 4 || !aligned16(s.act_out) ||
          (s.act_out2 && (s.act_ldo2 % 4 || !aligned16(s.act_out2)))) return PQLB_E_UNSUPPORTED;
      if (cluster != 1) return PQLB_E_UNSUPPORTED;
      if ((rc = make_map(&G.tmW4, s.act_w, kFH3, (uint64_t)s.act_n, kFH3, 32, 16, CU_TENSOR_MAP_SWIZZLE_128B)) != PQLB_OK) return rc;
      G.act_n = s.act_n; G.act_b = s.act_b; G.act_noise = s.act_noise; G.act_out = s.act_out; G.act_out2 = s.act_out2;
      G.act_ldo = s.act_ldo; G.act_ldo2 = s.act_ldo2; G.act_ldnoise = s.act_ldnoise;
      G.noise_std = s.noise_std; G.noise_bound = s.noise_bound;
    }
  }
  // a ghost CTA rounds an odd number of row tiles up to whole pairs: its X tile is zero-filled and
  // its stores are clipped by TMA
  const unsigned gx = (unsigned)((tiles_m + cluster - 1) / cluster * cluster);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(gx, (unsigned)d->n_groups, 1);
  cfg.blockDim = dim3(kFThreads, 1, 1);
  cfg.dynamicSmemBytes = kFSmem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t le = cluster == 2 ? cudaLaunchKernelEx(&cfg, mlp_fwd_kernel<true>, P) : cudaLaunchKernelEx(&cfg, mlp_fwd_kernel<false>, P);
  PQLB_COUNT_LAUNCH(1);
  return le == cudaSuccess ? PQLB_OK : (int)le;
}
