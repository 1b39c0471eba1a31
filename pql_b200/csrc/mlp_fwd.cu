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
constexpr int kFStages = 5;                   // weight ring (16 KB tiles); the rest of shared memory is X + 16 store-staging chunks
constexpr int kFStageBytes = 128 * 128;       // every weight tile: 128 rows x 32 k (W2 goes as two row halves)
constexpr int kFXKb = 4;                      // input width <= 128
constexpr int kFXBytes = kFXKb * 128 * 128;
constexpr int kFChunk = 32 * 128;
constexpr int kFSmem = 1024 + kFXBytes + kFStages * kFStageBytes + kFEpiWarps * kFChunk + (kFH1 + kFH2 + 2 * kFH3) * 4;

struct alignas(64) MlpGroupDev {
  CUtensorMap tmX, tmW1, tmW2, tmW3, tmH1, tmH2, tmH3;
  const float* b1; const float* b2; const float* b3; const float* head_w; const float* head_b;
  float* q;
  int st1, st2, st3, pad;
};
struct alignas(64) MlpDev {
  MlpGroupDev g[PQLB_MAX_GROUPS];
  int M, k_in, kb1, cluster;   // cluster = CTAs (consecutive row tiles of one network) sharing every weight tile
  unsigned long long* dbg;     // optional timeline of CTA (0,0): clock64 stamps (PQLB_MLP_DEBUG)
};

__host__ __device__ constexpr uint32_t idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
}

__global__ void __launch_bounds__(kFThreads, 1)
mlp_fwd_kernel(const __grid_constant__ MlpDev P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t x_full, full_bar[kFStages], empty_bar[kFStages];
  __shared__ __align__(8) uint64_t p_full[2], p_conv[2], y_full, y_conv[2], z_full;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_q[4][128];

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const MlpGroupDev& G = P.g[blockIdx.y];
  const int m0 = blockIdx.x * 128;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t xs = base;
  const uint32_t ring = xs + kFXBytes;
  const uint32_t stage_buf = ring + kFStages * kFStageBytes;
  float* s_vec = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)) + kFXBytes + kFStages * kFStageBytes + kFEpiWarps * kFChunk);
  float* s_b1 = s_vec; float* s_b2 = s_b1 + kFH1; float* s_b3 = s_b2 + kFH2; float* s_w4 = s_b3 + kFH3;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&x_full), 1);
    for (int s = 0; s < kFStages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), (uint32_t)P.cluster); }
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&p_full[b]), 1); mbar_init(smem_u32(&p_conv[b]), kFEpiWarps); mbar_init(smem_u32(&y_conv[b]), kFEpiWarps); }
    mbar_init(smem_u32(&y_full), 1); mbar_init(smem_u32(&z_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int j = threadIdx.x; j < kFH1; j += kFThreads) s_b1[j] = G.b1[j];
  for (int j = threadIdx.x; j < kFH2; j += kFThreads) s_b2[j] = G.b2[j];
  for (int j = threadIdx.x; j < kFH3; j += kFThreads) { s_b3[j] = G.b3[j]; s_w4[j] = G.q ? G.head_w[j] : 0.f; }
  tcgen05_fence_before();
  __syncthreads();
  const int C = P.cluster;
  const uint32_t rank = C > 1 ? cluster_ctarank() : 0u;
  const uint16_t mc_mask = (uint16_t)((1u << C) - 1u);
  if (C > 1) cluster_sync_all();            // every CTA's barriers are initialised before any remote arrive / multicast
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
  constexpr uint32_t idesc = idesc_tf32(128);
  const uint64_t desc0 = make_smem_desc(0, 16, 1024, kLayoutSw128);

  // Warps 0 and 1 run with all 32 lanes and issue from one elected lane (see elect_one()).
  if (warp == 0) {
    // ===================== TMA producer =====================
    const uint32_t xb = smem_u32(&x_full);
    if (elect_one()) {
      mbar_expect_tx(xb, (uint32_t)P.kb1 * 16384u);
      for (int kb = 0; kb < P.kb1; ++kb) tma_load_2d(xs + kb * 16384, &G.tmX, kb * 32, m0, xb);
    }
    __syncwarp();
    int stage = 0; uint32_t phase = 0;
    const int slice_rows = 128 / C;
#pragma unroll
    for (int ph = 0; ph < 9; ++ph) {
      const int kind = (int)((kKinds >> (4 * ph)) & 15), arg = (int)((kArgs >> (4 * ph)) & 15);
      const int n_tiles = kind == 0 ? P.kb1 : 8;
#pragma unroll 1
      for (int t = 0; t < n_tiles; ++t) {
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
        const uint32_t bar = smem_u32(&full_bar[stage]);
        if (elect_one()) {
          mbar_expect_tx(bar, 16384u);         // the whole tile: one slice from every CTA of the cluster
          if (C == 1) {
            const uint32_t dst = ring + stage * kFStageBytes;
            if (kind == 0) tma_load_2d(dst, &G.tmW1, t * 32, arg * 128, bar);
            else if (kind == 1) tma_load_2d(dst, &G.tmW2, arg * 128 + (t >> 1) * 32, (t & 1) * 128, bar);
            else tma_load_2d(dst, &G.tmW3, t * 32, 0, bar);
          } else {
            // this CTA fetches rows [rank * 128 / C, +128 / C) of the tile and multicasts them
            const int r0 = (int)rank * slice_rows;
            const uint32_t dst = ring + stage * kFStageBytes + rank * (uint32_t)(slice_rows * 128);
            if (kind == 0) tma_load_2d_mc(dst, &G.tmW1, t * 32, arg * 128 + r0, bar, mc_mask);
            else if (kind == 1) tma_load_2d_mc(dst, &G.tmW2, arg * 128 + (t >> 1) * 32, (t & 1) * 128 + r0, bar, mc_mask);
            else tma_load_2d_mc(dst, &G.tmW3, t * 32, r0, bar, mc_mask);
          }
        }
        __syncwarp();
        if (++stage == kFStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    auto free_stage = [&](uint32_t bar) { if (C == 1) umma_commit(bar); else umma_commit_mc(bar, mc_mask); };
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
              for (int k = 0; k < 4; ++k) umma_tf32(tP, adesc + 2u * k, bdesc + 2u * k, idesc, (uint32_t)(kb | k) != 0u);
            } else {
              const int ksteps = (krem + 7) / 8;
              for (int k = 0; k < ksteps; ++k) umma_tf32(tP, adesc + 2u * k, bdesc + 2u * k, idesc, (uint32_t)(kb | k) != 0u);
            }
            free_stage(smem_u32(&empty_bar[stage]));
          }
          __syncwarp();
          if (++stage == kFStages) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) umma_commit(smem_u32(&p_full[arg & 1]));
        __syncwarp();
        PQLB_STAMP(0);
      } else if (kind == 1) {
        const int c = arg, b = c & 1;
        const uint32_t tP = tmem + (uint32_t)(b * 128);
        mbar_wait(smem_u32(&p_conv[b]), (uint32_t)(c >> 1) & 1u);      // quarter c converted in place
        tcgen05_fence_after();
        PQLB_STAMP(0);
#pragma unroll
        for (int t = 0; t < 8; ++t) {          // (k-block t>>1 of the chunk) x (row half t&1 of W2)
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tcgen05_fence_after();
          const uint64_t bdesc = desc0 | (uint64_t)(((ring + stage * kFStageBytes) >> 4) & 0x3FFF);
          const int kk = t >> 1;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_tf32_ts(tY + (uint32_t)((t & 1) * 128), tP + (uint32_t)(kk * 32 + k * 8), bdesc + 2u * k,
                           idesc, (uint32_t)((c | kk | k) != 0));
            free_stage(smem_u32(&empty_bar[stage]));
          }
          __syncwarp();
          if (++stage == kFStages) { stage = 0; phase ^= 1u; }
        }
        if (c == 3) { if (elect_one()) umma_commit(smem_u32(&y_full)); __syncwarp(); }
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
              umma_tf32_ts(tZ, tY + (uint32_t)(t * 32 + k * 8), bdesc + 2u * k, idesc, (uint32_t)((t | k) != 0));
            free_stage(smem_u32(&empty_bar[stage]));
          }
          __syncwarp();
          if (++stage == kFStages) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) umma_commit(smem_u32(&z_full));
        __syncwarp();
        PQLB_STAMP(0);
      }
    }
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
    // converts this warp's chunk of a 128-column TMEM region in place: +bias, ELU, RN to TF32
    auto convert = [&](uint32_t region, const float* bias, int n_base, const CUtensorMap* omap, bool store) {
      const uint32_t taddr = region + lane_sel + (uint32_t)col;
      float v[32];
      tmem_ld32(taddr, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = rn_tf32(elu_fast(v[j] + bias[n_base + col + j]));
      tmem_st32(taddr, v);
      if (store) store_chunk(v, omap, n_base + col);
      tmem_wait_st();
      tcgen05_fence_before();
      __syncwarp();
    };

    if (e != 0) dbg = nullptr;
    PQLB_STAMP(32);
    for (int q = 0; q < 4; ++q) {
      const int b = q & 1;
      mbar_wait(smem_u32(&p_full[b]), (uint32_t)(q >> 1) & 1u);
      tcgen05_fence_after();
      PQLB_STAMP(32);
      convert(tmem + (uint32_t)(b * 128), s_b1, q * 128, &G.tmH1, G.st1 != 0);
      if (lane == 0) mbar_arrive(smem_u32(&p_conv[b]));
      PQLB_STAMP(32);
    }
    mbar_wait(smem_u32(&y_full), 0);
    tcgen05_fence_after();
    PQLB_STAMP(32);
    for (int hh = 0; hh < 2; ++hh) {
      convert(tY + (uint32_t)(hh * 128), s_b2, hh * 128, &G.tmH2, G.st2 != 0);
      if (lane == 0) mbar_arrive(smem_u32(&y_conv[hh]));
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
      if (G.st3) store_chunk(v, &G.tmH3, col);
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
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
  if (C > 1) cluster_sync_all();            // nobody leaves while a peer may still signal its barriers
}

}  // namespace pqlb

using namespace pqlb;

static unsigned long long* g_mlp_debug = nullptr;
static int g_mlp_cluster = 0;
/* Experiments: force the cluster size (1, 2 or 4) of pqlb_mlp_forward; 0 = automatic. */
extern "C" void pqlb_mlp_forward_cluster(int c) { g_mlp_cluster = (c == 1 || c == 2 || c == 4) ? c : 0; }
/* Debug: device buffer of 64 uint64 receiving a clock64 timeline of CTA (0,0) (NULL = off). */
extern "C" void pqlb_mlp_forward_debug(unsigned long long* buf) { g_mlp_debug = buf; }

extern "C" int pqlb_mlp_forward_init(void) {
  cudaError_t e = cudaFuncSetAttribute(mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFSmem);
  return e == cudaSuccess ? PQLB_OK : (int)e;
}

extern "C" int pqlb_mlp_forward(const pqlb_mlp_desc* d, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(d && d->M > 0 && d->k_in > 0 && d->n_groups >= 1 && d->n_groups <= PQLB_MAX_GROUPS);
  if (d->k_in > kFXKb * 32) return PQLB_E_UNSUPPORTED;      // wider inputs take the per-layer path
  { int rc = pqlb_init(); if (rc != PQLB_OK) return rc; }
  static MlpDev P;
  P.M = d->M; P.k_in = d->k_in; P.kb1 = (d->k_in + 31) / 32;
  const int tiles_m = (d->M + 127) / 128;
  // Measured on B200 (M = 8192, 1 and 4 networks): multicast clusters of 2 / 4 row tiles are 0-7 % SLOWER
  // than independent CTAs - the kernel is bound by the ~40 B/clk each SM can take in from L2, not by
  // L2 read traffic - so the default is no cluster; the path stays for the cta_group::2 follow-up.
  P.cluster = g_mlp_cluster > 0 ? g_mlp_cluster : 1;
  P.dbg = g_mlp_debug;
  for (int i = 0; i < d->n_groups; ++i) {
    const pqlb_mlp_group& s = d->g[i];
    MlpGroupDev& G = P.g[i];
    PQLB_CHECK_ARG(s.x && s.w1 && s.w2 && s.w3 && s.b1 && s.b2 && s.b3);
    PQLB_CHECK_ARG(!s.q || (s.head_w && s.head_b));
    int rc;
    if ((rc = make_map(&G.tmX, s.x, (uint64_t)d->k_in, (uint64_t)d->M, s.ldx, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B)) != PQLB_OK) return rc;
    const uint32_t wrows = 128u / (uint32_t)P.cluster;      // rows of a weight tile each CTA of the cluster fetches
    if ((rc = make_map(&G.tmW1, s.w1, (uint64_t)d->k_in, kFH1, s.ldw1, 32, wrows, CU_TENSOR_MAP_SWIZZLE_128B)) != PQLB_OK) return rc;
    if ((rc = make_map(&G.tmW2, s.w2, kFH1, kFH2, kFH1, 32, wrows, CU_TENSOR_MAP_SWIZZLE_128B)) != PQLB_OK) return rc;
    if ((rc = make_map(&G.tmW3, s.w3, kFH2, kFH3, kFH2, 32, wrows, CU_TENSOR_MAP_SWIZZLE_128B)) != PQLB_OK) return rc;
    G.st1 = s.h1 != nullptr; G.st2 = s.h2 != nullptr; G.st3 = s.h3 != nullptr;
    if (G.st1 && !make_tile_map(&G.tmH1, s.h1, kFH1, (uint64_t)d->M, kFH1, 1, 0)) return PQLB_E_ALIGN;
    if (G.st2 && !make_tile_map(&G.tmH2, s.h2, kFH2, (uint64_t)d->M, kFH2, 1, 0)) return PQLB_E_ALIGN;
    if (G.st3 && !make_tile_map(&G.tmH3, s.h3, kFH3, (uint64_t)d->M, kFH3, 1, 0)) return PQLB_E_ALIGN;
    if (!G.st1) G.tmH1 = G.tmX;
    if (!G.st2) G.tmH2 = G.tmX;
    if (!G.st3) G.tmH3 = G.tmX;
    G.b1 = s.b1; G.b2 = s.b2; G.b3 = s.b3; G.head_w = s.head_w; G.head_b = s.head_b; G.q = s.q;
  }
  // ghost CTAs round the row tiles up to whole clusters: their X tile is zero-filled and their
  // stores are clipped by TMA, so they only take part in the weight multicast
  const unsigned gx = (unsigned)((tiles_m + P.cluster - 1) / P.cluster * P.cluster);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(gx, (unsigned)d->n_groups, 1);
  cfg.blockDim = dim3(kFThreads, 1, 1);
  cfg.dynamicSmemBytes = kFSmem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)P.cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, mlp_fwd_kernel, P);
  PQLB_COUNT_LAUNCH(1);
  return le == cudaSuccess ? PQLB_OK : (int)le;
}
