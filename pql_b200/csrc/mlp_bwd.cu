// K3b: layer-fused dgrad chain of the MLP trunk
//     dz2 = (dz3 . W3) * elu'(h2)        [M,128] -> [M,256]
//     dz1 = (dz2 . W2) * elu'(h1)        [M,256] -> [M,512]
// for up to four network instances per launch, plus the per-128-row partial column sums of dz2 and
// dz1 (the bias gradients).  One CTA owns one 128-row tile of one network:
//
//   * dz3's tile is staged once by TMA (K-major, 128-byte swizzle); the weights stream through a
//     ring of 16 KB tiles read as [K][N] (MN-major, 32-byte swizzle atoms: no transposed copy);
//   * dz2 accumulates in TMEM (256 columns), is multiplied by elu'(h2) and rounded IN PLACE, and
//     is then the A operand of the second contraction straight from tensor memory;
//   * dz1 is produced as four 128-column quarters into two ping-pong TMEM regions, so the
//     epilogue of quarter q (x elu'(h1), rounding, TMA store, column sums) overlaps the MMAs of
//     quarter q + 1;
//   * sixteen epilogue warps (four per TMEM lane quarter, one 32-column chunk each per 128-column
//     region): the elu' operand chunk arrives by TMA in the warp's staging buffer, the result
//     goes back into the same buffer and leaves with a TMA store; the column sums are read out of
//     the staged chunk (one column per lane), combined over the four lane quarters in a fixed
//     order and written as part[tile][col].
//
// Compared with the two per-layer launches this reads dz3 / dz2 once instead of once per output
// column tile, never reads dz2 back from memory, and removes the separate bias-gradient pass over
// dz2 and dz1.  Replaces: autograd backward of nn.Linear + nn.ELU, pql/models/mlp.py:15-24, as
// exercised by loss.backward() in pql/algo/pql_v_learner.py:125 and pql_p_learner.py:60.
#include "tcgen05_utils.cuh"

namespace pqlb {

constexpr int kBH1 = 512, kBH2 = 256, kBH3 = 128;
constexpr int kBEpiWarps = 16;
constexpr int kBThreads = 64 + 32 * kBEpiWarps;
constexpr int kBStages = 5;
constexpr int kBStageBytes = 16384;           // one weight tile: 32 k x 128 n, four MN-major boxes
constexpr int kBABytes = 4 * 16384;           // dz3 tile: 128 rows x 128 k
constexpr int kBChunk = 4096;
constexpr int kBSmem = 1024 + kBABytes + kBStages * kBStageBytes + kBEpiWarps * kBChunk;

struct alignas(64) BwdGroupDev {
  CUtensorMap tmDz3, tmW3, tmW2, tmH2, tmH1, tmDz2, tmDz1;
  float* part2; float* part1;
};
struct alignas(64) BwdDev {
  BwdGroupDev g[PQLB_MAX_GROUPS];
  int M, pad0, pad1, pad2;
};

// instruction descriptor: fp32 accumulate, TF32 operands, B MN-major (bit 16), N = 128, M = 128
__host__ __device__ constexpr uint32_t idesc_tf32_bmn(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((unsigned)(n >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
}

__global__ void __launch_bounds__(kBThreads, 1)
mlp_bwd_kernel(const __grid_constant__ BwdDev P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full, full_bar[kBStages], empty_bar[kBStages];
  __shared__ __align__(8) uint64_t t2_full, t2_conv[2], q_full[2], q_free[2];
  __shared__ __align__(8) uint64_t aux_bar[kBEpiWarps], aux_free[kBEpiWarps];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_col[2][4][4][32];          // [round parity][chunk][lane quarter][column]

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const BwdGroupDev& G = P.g[blockIdx.y];
  const int m0 = blockIdx.x * 128;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t as = base;
  const uint32_t ring = as + kBABytes;
  const uint32_t stage_buf = ring + kBStages * kBStageBytes;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&a_full), 1);
    for (int s = 0; s < kBStages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    mbar_init(smem_u32(&t2_full), 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&t2_conv[b]), kBEpiWarps); mbar_init(smem_u32(&q_full[b]), 1); mbar_init(smem_u32(&q_free[b]), kBEpiWarps);
    }
    for (int w = 0; w < kBEpiWarps; ++w) { mbar_init(smem_u32(&aux_bar[w]), 1); mbar_init(smem_u32(&aux_free[w]), 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = uniform_u32(tmem_slot);
  const uint32_t tT2 = tmem;                    // dz2 accumulator / A operand, 256 columns
  constexpr uint32_t idesc = idesc_tf32_bmn(128);
  const uint64_t adesc0 = make_smem_desc(0, 16, 1024, kLayoutSw128);
  const uint64_t bdesc0 = make_smem_desc(0, 4096, 512, kLayoutSw128Base32);

  // Weight-tile order (shared by producer and MMA issuer): 8 tiles of W3 (k-block kb, column half
  // hf), then for every dz1 quarter q its 8 k-blocks of W2.
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      const uint32_t ab = smem_u32(&a_full);
      mbar_expect_tx(ab, (uint32_t)kBABytes);
      for (int kb = 0; kb < 4; ++kb) tma_load_2d(as + kb * 16384, &G.tmDz3, kb * 32, m0, ab);
    }
    __syncwarp();
    if (blockIdx.x < 2) {
      // L2 prefetch of this network's W3 / W2 (640 KB, cold after the forward's activation traffic): see tma_prefetch_2d
      if (elect_one()) {
        for (int t = 0; t < 40; ++t) {
          const CUtensorMap* map; int n0, k0;
          if (t < 8) { map = &G.tmW3; k0 = (t >> 1) * 32; n0 = (t & 1) * 128; }
          else { map = &G.tmW2; const int u = t - 8; k0 = (u & 7) * 32; n0 = (u >> 3) * 128; }
          for (int cb = 0; cb < 4; ++cb) tma_prefetch_2d(map, n0 + cb * 32, k0);
        }
      }
      __syncwarp();
    }
    int stage = 0; uint32_t phase = 0;
#pragma unroll 1
    for (int t = 0; t < 40; ++t) {
      mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
      const uint32_t bar = smem_u32(&full_bar[stage]);
      const uint32_t dst = ring + stage * kBStageBytes;
      if (elect_one()) {
        mbar_expect_tx(bar, (uint32_t)kBStageBytes);
        const CUtensorMap* map; int n0, k0;
        if (t < 8) { map = &G.tmW3; k0 = (t >> 1) * 32; n0 = (t & 1) * 128; }
        else { map = &G.tmW2; const int u = t - 8; k0 = (u & 7) * 32; n0 = (u >> 3) * 128; }
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) tma_load_2d(dst + cb * 4096, map, n0 + cb * 32, k0, bar);
      }
      __syncwarp();
      if (++stage == kBStages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    mbar_wait(smem_u32(&a_full), 0);
    tcgen05_fence_after();
    int stage = 0; uint32_t phase = 0;
    // ---- dz2 = dz3 . W3: A from shared memory, two 128-column halves of the accumulator
    // (tile loops not unrolled: 160 unrolled MMAs were 25 KB of straight-line code that a CTA walks once and that is
    // never in the instruction cache during a training step)
#pragma unroll 1
    for (int t = 0; t < 8; ++t) {
      const int kb = t >> 1, hf = t & 1;
      mbar_wait(smem_u32(&full_bar[stage]), phase);
      tcgen05_fence_after();
      const uint64_t adesc = adesc0 | (uint64_t)(((as + kb * 16384) >> 4) & 0x3FFF);
      const uint64_t bdesc = bdesc0 | (uint64_t)(((ring + stage * kBStageBytes) >> 4) & 0x3FFF);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_tf32(tT2 + (uint32_t)(hf * 128), adesc + 2u * k, bdesc + 64u * k, idesc, (uint32_t)((kb | k) != 0));
        umma_commit(smem_u32(&empty_bar[stage]));
      }
      __syncwarp();
      if (++stage == kBStages) { stage = 0; phase ^= 1u; }
    }
    if (elect_one()) umma_commit(smem_u32(&t2_full));
    __syncwarp();
    // ---- dz1 quarter q = dz2 . W2[:, 128 q .. 128 q + 128): A from tensor memory
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {
      const int b = q & 1;
      const uint32_t tQ = tmem + 256u + (uint32_t)(b * 128);
      if (q >= 2) { mbar_wait(smem_u32(&q_free[b]), 0); tcgen05_fence_after(); }   // epilogue of quarter q - 2 has drained Q[b]
#pragma unroll 1
      for (int kb = 0; kb < 8; ++kb) {
        if (q == 0 && (kb & 3) == 0) { mbar_wait(smem_u32(&t2_conv[kb >> 2]), 0); tcgen05_fence_after(); }
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        tcgen05_fence_after();
        const uint64_t bdesc = bdesc0 | (uint64_t)(((ring + stage * kBStageBytes) >> 4) & 0x3FFF);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_tf32_ts(tQ, tT2 + (uint32_t)(kb * 32 + k * 8), bdesc + 64u * k, idesc, (uint32_t)((kb | k) != 0));
          umma_commit(smem_u32(&empty_bar[stage]));
        }
        __syncwarp();
        if (++stage == kBStages) { stage = 0; phase ^= 1u; }
      }
      if (elect_one()) umma_commit(smem_u32(&q_full[b]));
      __syncwarp();
    }
  } else {
    // ===================== epilogue warps =====================
    const int e = warp - 2;
    const int quarter = warp & 3;                 // TMEM lanes [32 quarter, +32)
    const int chunk = e >> 2;                     // 32-column chunk of every 128-column region
    const int col = chunk * 32;
    const int row0 = m0 + quarter * 32;
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    const uint32_t my_stage = stage_buf + e * kBChunk;
    const uint32_t my_aux_bar = smem_u32(&aux_bar[e]);
    const uint32_t my_aux_free = smem_u32(&aux_free[e]);
    const uint32_t swz = (uint32_t)(lane & 7) << 4;
    const uint32_t row_off = (uint32_t)lane * 128u;
    uint32_t aux_phase = 0;

    // round r = 0, 1: columns [128 r, +128) of dz2 (elu' operand h2); round 2 + q: quarter q of dz1 (h1)
    auto fetch_aux = [&](int round) {
      if (elect_one()) {
        mbar_expect_tx(my_aux_bar, kBChunk);
        if (round < 2) tma_load_3d(my_stage, &G.tmH2, round * 128 + col, row0, 0, my_aux_bar);
        else tma_load_3d(my_stage, &G.tmH1, (round - 2) * 128 + col, row0, 0, my_aux_bar);
      }
      __syncwarp();
    };
    fetch_aux(0);

#pragma unroll 1
    for (int round = 0; round < 6; ++round) {
      uint32_t taddr;
      if (round < 2) {
        if (round == 0) { mbar_wait(smem_u32(&t2_full), 0); tcgen05_fence_after(); }
        taddr = tT2 + lane_sel + (uint32_t)(round * 128 + col);
      } else {
        const int q = round - 2, b = q & 1;
        mbar_wait(smem_u32(&q_full[b]), (uint32_t)(q >> 1) & 1u);
        tcgen05_fence_after();
        taddr = tmem + 256u + (uint32_t)(b * 128) + lane_sel + (uint32_t)col;
      }
      float v[32];
      tmem_ld32(taddr, v);
      if (round >= 2) {                           // Q[b] may be overwritten by quarter q + 2
        tcgen05_fence_before();
        if (lane == 0) mbar_arrive(smem_u32(&q_free[(round - 2) & 1]));
      }
      // elu' operand chunk
      mbar_wait(my_aux_bar, aux_phase);
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 t = lds128(my_stage + row_off + (((uint32_t)j4 << 4) ^ swz));
        const float h[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) v[4 * j4 + k] = rn_tf32(v[4 * j4 + k] * (h[k] > 0.f ? 1.f : h[k] + 1.f));
      }
      if (round < 2) {                            // in place: dz2 becomes the A operand of the second contraction
        tmem_st32(taddr, v);
      }
      // result into the staging buffer.  Each lane overwrites only the row it has just read, in
      // program order, so no cross-lane hazard with the loads above.
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4)
        sts128(my_stage + row_off + (((uint32_t)j4 << 4) ^ swz), v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
      fence_proxy_async();
      __syncwarp();
      const int n_col = (round < 2 ? round * 128 : (round - 2) * 128) + col;
      if (elect_one()) {
        tma_store_3d(round < 2 ? &G.tmDz2 : &G.tmDz1, my_stage, n_col, row0, 0);
        bulk_commit();
      }
      __syncwarp();
      if (round < 2) {
        tmem_wait_st();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&t2_conv[round]));
      }
      // column sums of the staged chunk: lane j adds column j over the 32 rows in row order
      float* part = round < 2 ? G.part2 : G.part1;
      if (part) {
        float cs = 0.f;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          float x;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x)
                       : "r"(my_stage + (uint32_t)r * 128u + ((((uint32_t)lane >> 2) ^ (uint32_t)(r & 7)) << 4) + ((uint32_t)lane & 3u) * 4u));
          cs += x;
        }
        s_col[round & 1][chunk][quarter][lane] = cs;
        named_bar_sync(1 + chunk, 128);           // the four lane quarters of this chunk
        if (quarter == 0) {
          const float tot = ((s_col[round & 1][chunk][0][lane] + s_col[round & 1][chunk][1][lane]) +
                             s_col[round & 1][chunk][2][lane]) + s_col[round & 1][chunk][3][lane];
          const int n_cols = round < 2 ? kBH2 : kBH1;
          part[(long long)blockIdx.x * n_cols + n_col + lane] = tot;
        }
      }
      // the staging buffer is refilled through the async proxy: the store must have read it and
      // every lane's column-sum loads must have completed (released through aux_free)
      aux_phase ^= 1u;
      if (round + 1 < 6) {
        if (elect_one()) bulk_wait_read<0>();
        mbar_arrive(my_aux_free);
        mbar_wait(my_aux_free, aux_phase ^ 1u);
        fetch_aux(round + 1);
      }
    }
    if (elect_one()) bulk_wait_read<0>();
    __syncwarp();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

}  // namespace pqlb

using namespace pqlb;

extern "C" int pqlb_init(void);

extern "C" int pqlb_mlp_backward_init(void) {
  cudaError_t e = cudaFuncSetAttribute(mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBSmem);
  return e == cudaSuccess ? PQLB_OK : (int)e;
}

extern "C" int pqlb_mlp_backward(const pqlb_mlp_bwd_desc* d, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(d && d->M > 0 && d->n_groups >= 1 && d->n_groups <= PQLB_MAX_GROUPS);
  { int rc = pqlb_init(); if (rc != PQLB_OK) return rc; }
  static BwdDev P;
  P.M = d->M;
  const uint64_t M = (uint64_t)d->M;
  for (int i = 0; i < d->n_groups; ++i) {
    const pqlb_mlp_bwd_group& s = d->g[i];
    BwdGroupDev& G = P.g[i];
    PQLB_CHECK_ARG(s.dz3 && s.w3 && s.w2 && s.h2 && s.h1 && s.dz2 && s.dz1);
    int rc;
    if ((rc = make_map(&G.tmDz3, s.dz3, kBH3, M, kBH3, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B)) != PQLB_OK) return rc;
    // weights [out = K][in = N] row-major read as [K][N]: boxes of 32 n x 32 k, 32-byte swizzle atoms
    if ((rc = make_map(&G.tmW3, s.w3, kBH2, kBH3, kBH2, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) != PQLB_OK) return rc;
    if ((rc = make_map(&G.tmW2, s.w2, kBH1, kBH2, kBH1, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) != PQLB_OK) return rc;
    if (!make_tile_map(&G.tmH2, s.h2, kBH2, M, kBH2, 1, 0)) return PQLB_E_ALIGN;
    if (!make_tile_map(&G.tmH1, s.h1, kBH1, M, kBH1, 1, 0)) return PQLB_E_ALIGN;
    if (!make_tile_map(&G.tmDz2, s.dz2, kBH2, M, kBH2, 1, 0)) return PQLB_E_ALIGN;
    if (!make_tile_map(&G.tmDz1, s.dz1, kBH1, M, kBH1, 1, 0)) return PQLB_E_ALIGN;
    G.part2 = s.bias_part2; G.part1 = s.bias_part1;
  }
  dim3 grid((unsigned)((d->M + 127) / 128), (unsigned)d->n_groups);
  mlp_bwd_kernel<<<grid, kBThreads, kBSmem, (cudaStream_t)stream>>>(P);
  PQLB_LAUNCH_RET();
}
