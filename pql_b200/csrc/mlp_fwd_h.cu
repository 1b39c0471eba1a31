// K3f': layer-fused MLP trunk forward with SPLIT-FP16 operands
//     x -> ELU(L1, 512) -> ELU(L2, 256) -> ELU(L3, 128) [-> scalar Q head | tanh policy head]
// for up to PQLB_MAX_FWD_GROUPS network instances per launch; same dataflow as mlp_fwd.cu (one CTA
// per 128-row tile of one network, activations never leave tensor memory between layers), other
// number format.  Every GEMM operand v is represented as hi + lo with hi = fp16(v) and
// lo = fp16(v - hi) (22 significand bits together) and a product a.w is evaluated as
//       a_hi.w_hi + a_lo.w_hi + a_hi.w_lo        (terms = 3; the lo.lo term is 2^-22 and dropped)
// by three tcgen05.mma kind::f16 into the same fp32 accumulator, or as a_hi.w_hi alone (terms = 1,
// the accuracy of one TF32 MMA: fp16 and TF32 both keep 11 significand bits).  Why: with one TF32
// MMA per product the critics' Q values carry ~5e-4 relative error, which the TD error (Q - y) and the
// DPG gradient amplify past BASELINE.json's 1e-3 (tools/precision_study.py: bias gradients up to
// 6e-3, P-learner weight gradients 2e-3); three fp16 MMAs cost 1.5x the tensor-pipe time of one TF32
// MMA (K = 16 per instruction instead of 8), move the same bytes through shared and tensor memory
// (two 2-byte halves instead of one 4-byte word) and bring every gradient tensor back under 1e-3.
//   * weights arrive as fp16 hi / lo copies SCALED by 2^8 (kept by the optimiser kernel,
//     pqlb_split_f16): typical |w| ~ 0.05 puts w_lo ~ 1e-5 below fp16's normal range, the scale
//     keeps both halves normal; the epilogue multiplies the accumulator by 2^-8 (exact);
//   * the input tile arrives by TMA as fp32 and the conversion warps split it IN PLACE in shared
//     memory: a 128-byte row of 32 floats becomes [32 hi halves | 32 lo halves] in the same 128-byte
//     swizzle, i.e. K-steps 0,1 of the block are the hi parts and K-steps 2,3 the lo parts;
//   * a finished 128-column fp32 accumulator region is converted in place (tcgen05.ld -> *2^-8 + bias
//     -> ELU -> split -> pack) into 64 columns of packed hi pairs + 64 columns of packed lo pairs,
//     which are the A operand (in tensor memory) of the next layer's MMAs;
//   * h1 / h2 / h3 go to HBM (TF32-rounded fp32, what the backward kernels read) only where asked.
// TMEM columns: P0 [0,128) P1 [128,256) Y [256,512); Z = P0; policy head output in P1.
// Replaces: the nn.Linear + nn.ELU launches of pql/models/mlp.py:15-24 for actor and critics
// (pql/algo/pql_v_learner.py:81-107, pql/algo/pql_p_learner.py:55-56).
#include "mlp_fwd_h.cuh"

namespace pqlb {

template <int T, bool WIDE>
__device__ __forceinline__ void mlp_fwd_h_body(const MlpHDev& P, const MlpHGroupDev& G, uint8_t* smem_raw) {
  // WIDE: inputs of 129..256 columns (eight 32-float blocks resident), a shorter ring, short staging chunks (mlp_fwd_h.cuh)
  constexpr int kStages = WIDE ? kHStagesWide : kHStages;
  constexpr int kXBytes = (WIDE ? kHXKbWide : kHXKb) * 128 * 128;
  constexpr int kChunk = WIDE ? kHChunkWide : kHChunk;
  __shared__ __align__(8) uint64_t x_full, x_conv, full_bar[kStages], empty_bar[kStages];
  __shared__ __align__(8) uint64_t p_full[2], p_conv[2], y_full, y_conv[2], z_full, z_conv, a_full;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_q[4][128];
  // T = 0: the number of terms is a run-time property of the group (ONE copy of the code for policy and critic tiles:
  // the kernel is far larger than the instruction cache, and a policy tile then runs code the critics' tiles have
  // just pulled in instead of its own cold instantiation); T = 1 / 3: compile-time specialisations
  const bool t3 = T == 3 || (T == 0 && G.terms == 3);
  const int NP = t3 ? 2 : 1;                    // weight tiles per (rows, k) block: hi [, lo]

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t xs = base;
  const uint32_t ring = xs + kXBytes;
  const uint32_t stage_buf = ring + kStages * kHTileBytes;
  float* s_vec = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)) + kXBytes + kStages * kHTileBytes + kHEpiWarps * kChunk);
  float* s_b1 = s_vec; float* s_b2 = s_b1 + kHH1; float* s_b3 = s_b2 + kHH2; float* s_w4 = s_b3 + kHH3;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&x_full), 1); mbar_init(smem_u32(&x_conv), kHEpiWarps);
    for (int s = 0; s < kStages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&p_full[b]), 1); mbar_init(smem_u32(&p_conv[b]), kHEpiWarps); mbar_init(smem_u32(&y_conv[b]), kHEpiWarps); }
    mbar_init(smem_u32(&y_full), 1); mbar_init(smem_u32(&z_full), 1);
    mbar_init(smem_u32(&z_conv), kHEpiWarps); mbar_init(smem_u32(&a_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();                               // barriers initialised
  // The TMA producer needs neither the bias vectors nor the tensor-memory address: it starts streaming the
  // input tile and the first weight tiles right away, while the other warps load the biases (global-memory
  // latency) and allocate tensor memory, and meets nobody until the final __syncthreads.
  if (warp != 0) {
    const int t = (int)threadIdx.x - 32;
    for (int j = t; j < kHH1; j += kHThreads - 32) s_b1[j] = G.b1[j];
    for (int j = t; j < kHH2; j += kHThreads - 32) s_b2[j] = G.b2[j];
    for (int j = t; j < kHH3; j += kHThreads - 32) { s_b3[j] = G.b3[j]; s_w4[j] = G.q ? G.head_w[j] : 0.f; }
    if (warp == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    named_bar_sync(7, kHThreads - 32);
    tcgen05_fence_after();
  }
  const uint32_t tmem = warp != 0 ? uniform_u32(tmem_slot) : 0u;
  const uint32_t tY = tmem + 256u;
  const uint32_t tZ = tmem;
  unsigned long long* dbg = (blockIdx.x == 0 && blockIdx.y == 0) ? P.dbg : nullptr;
  int dbg_i = 0;
#define PQLB_HSTAMP(base) do { if (dbg && lane == 0) dbg[(base) + dbg_i] = clock64(); ++dbg_i; } while (0)

  // Weight-tile schedule shared by producer and MMA issuer, one hex digit per phase (phase 0 is the
  // lowest digit): kind 0 = layer-1 quarter q, 1 = layer-2 K-chunk c, 2 = layer 3.
  constexpr unsigned long long kKinds = 0x211010100ull, kArgs = 0x032312010ull;
  constexpr uint32_t idesc = idesc_f16(128);
  const uint64_t desc0 = make_smem_desc(0, 16, 1024, kLayoutSw128);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (G.wait) {
      // this tile's input is produced by another group of the same launch (target policy -> target
      // critics): its CTAs have lower block indices, are therefore resident or finished, and never
      // wait themselves, so spinning here cannot deadlock
      if (lane == 0) {
        const volatile unsigned* f = P.tile_sync + 2 + blockIdx.x;
        const unsigned epoch = 1u + *reinterpret_cast<const volatile unsigned*>(P.tile_sync + 1);
        const long long t0 = clock64();
        while (*f != epoch) { if (clock64() - t0 > 4000000000LL) __trap(); }
        __threadfence();
      }
      __syncwarp();
      asm volatile("fence.proxy.async.global;" ::: "memory");      // the producer's generic-proxy stores, read by TMA below
    }
    if (elect_one()) {
      mbar_expect_tx(smem_u32(&x_full), (uint32_t)G.kb1 * 16384u);
      for (int kb = 0; kb < G.kb1; ++kb) tma_load_2d(xs + kb * 16384, &G.tmX, kb * 32, m0, smem_u32(&x_full));
    }
    __syncwarp();
    if (blockIdx.x < kHPrefetchCtas) {
      // a few CTAs of every group pull the group's whole weight stream into L2 right away (see tma_prefetch_2d):
      // the ring then streams from L2 instead of paying HBM latency tile by tile
      if (elect_one()) {
        for (int q = 0; q < 4; ++q)
          for (int kw = 0; kw < G.kw1; ++kw)
            for (int part = 0; part < NP; ++part) tma_prefetch_2d(&G.tmW1[part], kw * 32, q * 128);
        for (int c = 0; c < 4; ++c)
          for (int t = 0; t < 4; ++t)
            for (int part = 0; part < NP; ++part) tma_prefetch_2d(&G.tmW2[part], c * 64 + (t >> 1) * 32, (t & 1) * 128);
        for (int t = 0; t < 4; ++t)
          for (int part = 0; part < NP; ++part) tma_prefetch_2d(&G.tmW3[part], t * 32, 0);
        if (G.head_rows > 0)
          for (int part = 0; part < NP; ++part)
            for (int kb = 0; kb < 2; ++kb) tma_prefetch_2d(&G.tmW4[part], kb * 32, 0);
      }
      __syncwarp();
    }
    int stage = 0; uint32_t phase = 0;
    auto load_tile = [&](const CUtensorMap* map, int c0, int c1) {
      mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(smem_u32(&full_bar[stage]), (uint32_t)kHTileBytes);
        tma_load_2d(ring + stage * kHTileBytes, map, c0, c1, smem_u32(&full_bar[stage]));
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1u; }
    };
#pragma unroll
    for (int ph = 0; ph < 9; ++ph) {
      const int kind = (int)((kKinds >> (4 * ph)) & 15), arg = (int)((kArgs >> (4 * ph)) & 15);
      if (kind == 0) {
#pragma unroll 1
        for (int kw = 0; kw < G.kw1; ++kw)
#pragma unroll
          for (int part = 0; part < NP; ++part) load_tile(&G.tmW1[part], kw * 32, arg * 128);
      } else if (kind == 1) {
#pragma unroll 1
        for (int t = 0; t < 4; ++t)             // (k half t >> 1 of the chunk) x (row half t & 1 of W2)
#pragma unroll
          for (int part = 0; part < NP; ++part) load_tile(&G.tmW2[part], arg * 64 + (t >> 1) * 32, (t & 1) * 128);
      } else {
#pragma unroll 1
        for (int t = 0; t < 4; ++t)             // k block t of W3 (64 halves each)
#pragma unroll
          for (int part = 0; part < NP; ++part) load_tile(&G.tmW3[part], t * 32, 0);
      }
    }
    if (G.head_rows > 0) {
      // head weights (policy: 16 rows, softmax: 64 rows) x 128 k: two boxes of head_rows x 64 halves, one stage per part
      const uint32_t box_bytes = (uint32_t)G.head_rows * 128u;
#pragma unroll
      for (int part = 0; part < NP; ++part) {
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
        const uint32_t bar = smem_u32(&full_bar[stage]);
        const uint32_t dst = ring + stage * kHTileBytes;
        if (elect_one()) {
          mbar_expect_tx(bar, 2u * box_bytes);
          for (int kb = 0; kb < 2; ++kb) tma_load_2d(dst + kb * box_bytes, &G.tmW4[part], kb * 32, 0, bar);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    PQLB_HSTAMP(0);
    mbar_wait(smem_u32(&x_conv), 0);             // the input tile has been split into fp16 hi | lo in place
    tcgen05_fence_after();
    PQLB_HSTAMP(0);
    int stage = 0; uint32_t phase = 0;
    auto wait_tile = [&]() -> uint64_t {
      mbar_wait(smem_u32(&full_bar[stage]), phase);
      tcgen05_fence_after();
      return desc0 | (uint64_t)(((ring + stage * kHTileBytes) >> 4) & 0x3FFF);
    };
    auto release_tile = [&]() {                 // called by the elected lane after its MMAs
      umma_commit(smem_u32(&empty_bar[stage]));
    };
    auto next_stage = [&]() { __syncwarp(); if (++stage == kStages) { stage = 0; phase ^= 1u; } };
#pragma unroll
    for (int ph = 0; ph < 9; ++ph) {
      const int kind = (int)((kKinds >> (4 * ph)) & 15), arg = (int)((kArgs >> (4 * ph)) & 15);
      if (kind == 0) {
        const uint32_t tP = tmem + (uint32_t)((arg & 1) * 128);
#pragma unroll 1
        for (int kw = 0; kw < G.kw1; ++kw) {
#pragma unroll
          for (int part = 0; part < NP; ++part) {
            const uint64_t bdesc = wait_tile();
            if (elect_one()) {
#pragma unroll
              for (int s = 0; s < 4; ++s) {
                const int k16 = kw * 4 + s;                  // 16-wide k step of the input
                if (k16 < G.ksteps1) {
                  // input block k16 >> 1 (32 floats -> [32 hi | 32 lo] halves), hi step k16 & 1, lo step 2 + (k16 & 1)
                  const uint64_t a_hi = (desc0 | (uint64_t)(((xs + (k16 >> 1) * 16384) >> 4) & 0x3FFF)) + 2u * (uint32_t)(k16 & 1);
                  if (part == 0) {
                    umma_f16(tP, a_hi, bdesc + 2u * s, idesc, (uint32_t)(k16 != 0));
                    if (t3) umma_f16(tP, a_hi + 4u, bdesc + 2u * s, idesc, 1u);
                  } else {
                    umma_f16(tP, a_hi, bdesc + 2u * s, idesc, 1u);
                  }
                }
              }
              release_tile();
            }
            next_stage();
          }
        }
        if (elect_one()) umma_commit(smem_u32(&p_full[arg & 1]));
        __syncwarp();
        PQLB_HSTAMP(0);
      } else if (kind == 1) {
        const int c = arg, b = c & 1;
        const uint32_t tP = tmem + (uint32_t)(b * 128);
        mbar_wait(smem_u32(&p_conv[b]), (uint32_t)(c >> 1) & 1u);      // quarter c converted in place
        tcgen05_fence_after();
        PQLB_HSTAMP(0);
        // (t not unrolled: the issuer walks its code once per tile, and in a training step that code is never in the
        // instruction cache - 64 unrolled MMAs per phase were 9 KB of straight-line SASS fetched from L2 each time)
#pragma unroll 1
        for (int t = 0; t < 4; ++t) {
#pragma unroll
          for (int part = 0; part < NP; ++part) {
            const uint64_t bdesc = wait_tile();
            if (elect_one()) {
              const uint32_t d = tY + (uint32_t)((t & 1) * 128);
#pragma unroll
              for (int s = 0; s < 4; ++s) {
                const uint32_t a_hi = tP + (uint32_t)((t >> 1) * 32 + s * 8);
                if (part == 0) {
                  umma_f16_ts(d, a_hi, bdesc + 2u * s, idesc, (uint32_t)((c | (t >> 1) | s) != 0));
                  if (t3) umma_f16_ts(d, a_hi + 64u, bdesc + 2u * s, idesc, 1u);
                } else {
                  umma_f16_ts(d, a_hi, bdesc + 2u * s, idesc, 1u);
                }
              }
              release_tile();
            }
            next_stage();
          }
        }
        if (c == 3) { if (elect_one()) umma_commit(smem_u32(&y_full)); __syncwarp(); }
        PQLB_HSTAMP(0);
      } else {
#pragma unroll 1
        for (int t = 0; t < 4; ++t) {            // k block t of layer 3: Y half t >> 1, 64-k half t & 1
          if ((t & 1) == 0) { mbar_wait(smem_u32(&y_conv[t >> 1]), 0); tcgen05_fence_after(); }
#pragma unroll
          for (int part = 0; part < NP; ++part) {
            const uint64_t bdesc = wait_tile();
            if (elect_one()) {
#pragma unroll
              for (int s = 0; s < 4; ++s) {
                const uint32_t a_hi = tY + (uint32_t)((t >> 1) * 128 + (t & 1) * 32 + s * 8);
                if (part == 0) {
                  umma_f16_ts(tZ, a_hi, bdesc + 2u * s, idesc, (uint32_t)((t | s) != 0));
                  if (t3) umma_f16_ts(tZ, a_hi + 64u, bdesc + 2u * s, idesc, 1u);
                } else {
                  umma_f16_ts(tZ, a_hi, bdesc + 2u * s, idesc, 1u);
                }
              }
              release_tile();
            }
            next_stage();
          }
        }
        if (elect_one()) umma_commit(smem_u32(&z_full));
        __syncwarp();
        PQLB_HSTAMP(0);
      }
    }
    if (G.head_rows > 0) {
      // head: h3 (split in place in Z) . W4^T into head_rows columns of P1 (policy N = 16, softmax N = 64)
      mbar_wait(smem_u32(&z_conv), 0);
      tcgen05_fence_after();
      const uint32_t idesc_head = G.head_rows == 16 ? idesc_f16(16) : (WIDE && G.head_rows == 32 ? idesc_f16(32) : idesc_f16(64));
      const uint32_t box_bytes = (uint32_t)G.head_rows * 128u;
#pragma unroll
      for (int part = 0; part < NP; ++part) {
        const uint64_t bdesc0 = wait_tile();
        if (elect_one()) {
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t bdesc = bdesc0 + (uint64_t)((kb * box_bytes) >> 4);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
              const uint32_t a_hi = tZ + (uint32_t)(kb * 32 + s * 8);
              if (part == 0) {
                umma_f16_ts(tmem + 128u, a_hi, bdesc + 2u * s, idesc_head, (uint32_t)((kb | s) != 0));
                if (t3) umma_f16_ts(tmem + 128u, a_hi + 64u, bdesc + 2u * s, idesc_head, 1u);
              } else {
                umma_f16_ts(tmem + 128u, a_hi, bdesc + 2u * s, idesc_head, 1u);
              }
            }
          }
          release_tile();
        }
        next_stage();
      }
      if (elect_one()) umma_commit(smem_u32(&a_full));
      __syncwarp();
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
    const uint32_t my_stage = stage_buf + e * kChunk;
    const uint32_t swz = (uint32_t)(lane & 7) << 4;
    const uint32_t row_off = (uint32_t)(WIDE ? (lane & (kHRowsWide - 1)) : lane) * 128u;
    bool pending = false;

    // ---- input tile: fp32 -> [hi | lo] fp16, in place (thread = one 128-byte row of one 32-float block)
    if (e != 0) dbg = nullptr;
    PQLB_HSTAMP(32);
    mbar_wait(smem_u32(&x_full), 0);
    PQLB_HSTAMP(32);
    {
      const int t = (int)threadIdx.x - 64;
      const int r = t & 127;
      // 512 threads = four blocks of 128 rows per pass; a wide tile (up to eight blocks) takes two passes
#pragma unroll 1
      for (int kb = t >> 7; kb < G.kb1; kb += (WIDE ? 4 : 64)) {
        const uint32_t rowaddr = xs + (uint32_t)kb * 16384u + (uint32_t)r * 128u;
        const uint32_t sw = (uint32_t)(r & 7) << 4;
        float f[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = lds128(rowaddr + (((uint32_t)c << 4) ^ sw));
          f[4 * c] = v.x; f[4 * c + 1] = v.y; f[4 * c + 2] = v.z; f[4 * c + 3] = v.w;
        }
        uint32_t hp[16], lp[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          hp[j] = pack_hi(f[2 * j], f[2 * j + 1]);
          lp[j] = 0u;
        }
        if (t3) {
#pragma unroll
          for (int j = 0; j < 16; ++j) lp[j] = pack_lo(f[2 * j], f[2 * j + 1], hp[j]);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + (((uint32_t)c << 4) ^ sw)),
                       "r"(hp[4 * c]), "r"(hp[4 * c + 1]), "r"(hp[4 * c + 2]), "r"(hp[4 * c + 3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + (((uint32_t)(c + 4) << 4) ^ sw)),
                       "r"(lp[4 * c]), "r"(lp[4 * c + 1]), "r"(lp[4 * c + 2]), "r"(lp[4 * c + 3]) : "memory");
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&x_conv));
    }
    PQLB_HSTAMP(32);

    // hands this warp's 32x32 chunk (already TF32-rounded) to a TMA store through its staging buffer
    auto store_chunk = [&](const float* v, const CUtensorMap* omap, int n_col) {
      if constexpr (WIDE) {
        // short staging chunk: the warp's 32 rows leave in pieces of kHRowsWide rows (tile maps with 32 x kHRowsWide boxes)
#pragma unroll 1
        for (int half = 0; half < 32 / kHRowsWide; ++half) {
          if (pending) { if (elect_one()) bulk_wait_read<0>(); __syncwarp(); }
          if (lane / kHRowsWide == half) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
              sts128(my_stage + row_off + (((uint32_t)j4 << 4) ^ swz), v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
          }
          fence_proxy_async();
          __syncwarp();
          if (elect_one()) { tma_store_3d(omap, my_stage, n_col, row0 + kHRowsWide * half, 0); bulk_commit(); }
          __syncwarp();
          pending = true;
        }
      } else {
        if (pending) { if (elect_one()) bulk_wait_read<0>(); __syncwarp(); }
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
          sts128(my_stage + row_off + (((uint32_t)j4 << 4) ^ swz), v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
        fence_proxy_async();
        __syncwarp();
        if (elect_one()) { tma_store_3d(omap, my_stage, n_col, row0, 0); bulk_commit(); }
        pending = true;
      }
    };
    // Converts this warp's chunk of a 128-column accumulator region: v * 2^-8 + bias, ELU, split; the
    // region becomes [64 columns of packed hi | 64 columns of packed lo].  The four warps of a lane
    // quarter read disjoint column chunks but write into each other's (chunk c's hi pairs land in
    // columns [16c, 16c + 16)), hence the named barrier between the loads and the stores.
    auto convert = [&](uint32_t region, const float* bias, int n_base, const CUtensorMap* omap, bool store, uint32_t done) {
      float v[32];
      tmem_ld32(region + lane_sel + (uint32_t)col, v);
      uint32_t hp[16], lp[16];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = elu_fast(fmaf(v[j], kWInv, bias[n_base + col + j]));
#pragma unroll
      for (int j = 0; j < 16; ++j) hp[j] = pack_hi_pos(v[2 * j], v[2 * j + 1]);
      if (t3) {
#pragma unroll
        for (int j = 0; j < 16; ++j) lp[j] = pack_lo(v[2 * j], v[2 * j + 1], hp[j]);
      }
      named_bar_sync(1 + quarter, 128);
      tmem_st16(region + lane_sel + (uint32_t)(chunk * 16), hp);
      if (t3) tmem_st16(region + lane_sel + (uint32_t)(64 + chunk * 16), lp);
      tmem_wait_st();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(done);
      if (store) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = rn_tf32(v[j]);
        store_chunk(v, omap, n_base + col);
      }
    };

    // (not unrolled: the kernel's code is far larger than the instruction cache, and a tile walks it once - every copy
    // of the conversion body is another ~7 KB fetched from L2 behind the epilogue warps; ncu: no_instruction stalls)
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {
      const int b = q & 1;
      mbar_wait(smem_u32(&p_full[b]), (uint32_t)(q >> 1) & 1u);
      tcgen05_fence_after();
      PQLB_HSTAMP(32);
      convert(tmem + (uint32_t)(b * 128), s_b1, q * 128, &G.tmH1, G.st1 != 0, smem_u32(&p_conv[b]));
      PQLB_HSTAMP(32);
    }
    mbar_wait(smem_u32(&y_full), 0);
    tcgen05_fence_after();
    PQLB_HSTAMP(32);
#pragma unroll 1
    for (int hh = 0; hh < 2; ++hh) {
      convert(tY + (uint32_t)(hh * 128), s_b2, hh * 128, &G.tmH2, G.st2 != 0, smem_u32(&y_conv[hh]));
      PQLB_HSTAMP(32);
    }
    mbar_wait(smem_u32(&z_full), 0);
    tcgen05_fence_after();
    PQLB_HSTAMP(32);
    {
      float qacc = 0.f;
      float v[32];
      tmem_ld32(tZ + lane_sel + (uint32_t)col, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = elu_fast(fmaf(v[j], kWInv, s_b3[col + j]));
        qacc = fmaf(v[j], s_w4[col + j], qacc);
      }
      if (G.head_rows > 0) {                     // h3 back into Z as packed halves: the A operand of the head contraction
        uint32_t hp[16], lp[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) hp[j] = pack_hi_pos(v[2 * j], v[2 * j + 1]);
        if (t3) {
#pragma unroll
          for (int j = 0; j < 16; ++j) lp[j] = pack_lo(v[2 * j], v[2 * j + 1], hp[j]);
        }
        named_bar_sync(1 + quarter, 128);
        tmem_st16(tZ + lane_sel + (uint32_t)(chunk * 16), hp);
        if (t3) tmem_st16(tZ + lane_sel + (uint32_t)(64 + chunk * 16), lp);
        tmem_wait_st();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&z_conv));
      }
      if (G.st3) {
        float r[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = rn_tf32(v[j]);
        store_chunk(r, &G.tmH3, col);
      }
      if (G.act_n > 0 && chunk == 0) {
        // one warp per lane quarter finishes the head: * 2^-8 + bias, tanh, (+ clipped noise, clamp).  Bias and noise
        // are fetched BEFORE the wait for the head contraction (cold global loads at the very end of a tile otherwise).
        // The wide-input kernel also takes heads of 17..32 actions (a second pass over columns 16..31) and output rows
        // that are not 16-byte aligned (scalar stores): ShadowHand's 20 actions behind 211 observations.
        const int n_pass = WIDE ? (G.act_n + 15) >> 4 : 1;
#pragma unroll 1
        for (int pass = 0; pass < n_pass; ++pass) {
          const int c0 = pass * 16;
          float hb[16], nz[16];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), n4 = b4;
            if (c0 + j < G.act_n) {
              b4 = *reinterpret_cast<const float4*>(G.act_b + c0 + j);            // arena biases are 128-byte aligned and padded
              if (G.act_noise && row < P.M) n4 = *reinterpret_cast<const float4*>(G.act_noise + (long long)row * G.act_ldnoise + c0 + j);
            }
            hb[j] = b4.x; hb[j + 1] = b4.y; hb[j + 2] = b4.z; hb[j + 3] = b4.w;
            nz[j] = n4.x; nz[j + 1] = n4.y; nz[j + 2] = n4.z; nz[j + 3] = n4.w;
          }
          mbar_wait(smem_u32(&a_full), 0);
          tcgen05_fence_after();
          float a[16];
          tmem_ld16(tmem + 128u + (uint32_t)c0 + lane_sel, a);
          if (row < P.M) {
            float o[16], o2[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float t = 0.f, r = 0.f;
              if (c0 + j < G.act_n) {
                t = tanh_fast(fmaf(a[j], kWInv, hb[j]));
                r = t;
                if (G.act_noise) {                   // noise.py:19-27 on N(0,1) draws scaled by std
                  const float z = fminf(fmaxf(nz[j] * G.noise_std, -G.noise_bound), G.noise_bound);
                  r = fminf(fmaxf(t + z, -1.f), 1.f);
                }
                t = r;                               // act_out2: the value itself; act_out: its TF32 operand rounding
                r = rn_tf32(r);
              }
              o[j] = r; o2[j] = t;
            }
            if (G.act_out) {
              float* dst = G.act_out + (long long)row * G.act_ldo + c0;
              if (!WIDE || G.act_vec) {
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                  if (c0 + j < G.act_n) *reinterpret_cast<float4*>(dst + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (c0 + j < G.act_n) dst[j] = o[j];
              }
            }
            if (G.act_out2) {
              float* dst2 = G.act_out2 + (long long)row * G.act_ldo2 + c0;
              if (!WIDE || G.act_vec) {
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                  if (c0 + j < G.act_n) *reinterpret_cast<float4*>(dst2 + j) = make_float4(o2[j], o2[j + 1], o2[j + 2], o2[j + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (c0 + j < G.act_n) dst2[j] = o2[j];
              }
            }
          }
        }
        if (G.publish) {
          // the four head warps of this tile have written their rows: publish the tile to the CTAs
          // that wait for it (a later group of this launch)
          __threadfence();
          named_bar_sync(5, 128);
          if (quarter == 0 && lane == 0) {
            __threadfence();
            *reinterpret_cast<volatile unsigned*>(P.tile_sync + 2 + blockIdx.x) =
                1u + *reinterpret_cast<const volatile unsigned*>(P.tile_sync + 1);
          }
        }
      }
      if (G.sm_n > 0 && chunk == 0) {
        // C51 head (pql/models/mlp.py:261-263): all sm_n <= 64 logits of a row live in this thread: softmax
        // in registers, same arithmetic as the EPI_BIAS_SOFTMAX epilogue of gemm_tf32.cu
        mbar_wait(smem_u32(&a_full), 0);
        tcgen05_fence_after();
        float l[64];
        tmem_ld32(tmem + 128u + lane_sel, l);
        tmem_ld32(tmem + 128u + lane_sel + 32u, l + 32);
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 64; ++j) if (j < G.sm_n) { l[j] = fmaf(l[j], kWInv, G.sm_b[j]); mx = fmaxf(mx, l[j]); }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 64; ++j) if (j < G.sm_n) { l[j] = __expf(l[j] - mx); sum += l[j]; }
        const float inv = 1.f / sum;
        if (row < P.M) {
          float* dst = G.sm_out + (long long)row * G.sm_ldp;
#pragma unroll
          for (int j = 0; j < 64; j += 4) {
            float4 o;
            o.x = j < G.sm_n ? l[j] * inv : 0.f; o.y = j + 1 < G.sm_n ? l[j + 1] * inv : 0.f;
            o.z = j + 2 < G.sm_n ? l[j + 2] * inv : 0.f; o.w = j + 3 < G.sm_n ? l[j + 3] * inv : 0.f;
            *reinterpret_cast<float4*>(dst + j) = o;
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
    PQLB_HSTAMP(32);
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
  if (P.tile_sync && threadIdx.x == 0) {
    // exit ticket: the last CTA of the launch closes the epoch (every CTA has read [1] long before)
    const unsigned total = gridDim.x * gridDim.y;
    const unsigned epoch = *reinterpret_cast<const volatile unsigned*>(P.tile_sync + 1);
    __threadfence();
    if (atomicAdd(P.tile_sync, 1u) == total - 1u) {
      *reinterpret_cast<volatile unsigned*>(P.tile_sync) = 0u;
      __threadfence();
      *reinterpret_cast<volatile unsigned*>(P.tile_sync + 1) = epoch + 1u;
    }
  }
}

__global__ void __launch_bounds__(kHThreads, 1)
mlp_fwd_h_kernel(const __grid_constant__ MlpHDev P) {
  extern __shared__ uint8_t smem_raw[];
  const MlpHGroupDev& G = P.g[blockIdx.y];
  mlp_fwd_h_body<0, false>(P, G, smem_raw);
}

// the same tile program for inputs of 129..256 columns (ShadowHand critics: obs 211 + act 20 = 231)
__global__ void __launch_bounds__(kHThreads, 1)
mlp_fwd_hw_kernel(const __grid_constant__ MlpHDev P) {
  extern __shared__ uint8_t smem_raw[];
  const MlpHGroupDev& G = P.g[blockIdx.y];
  mlp_fwd_h_body<0, true>(P, G, smem_raw);
}

// hi[i] = fp16(scale * src[i]), lo[i] = fp16(scale * src[i] - hi[i]) (lo optional): the fp16 operand
// copies of a parameter arena, element i of the arena <-> half i of each copy.
__global__ void __launch_bounds__(256)
split_f16_kernel(const float* __restrict__ src, __half* __restrict__ hi, __half* __restrict__ lo, int64_t n, float scale) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 2;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n; i += stride) {
    const float a = src[i] * scale, b = i + 1 < n ? src[i + 1] * scale : 0.f;
    const uint32_t h = pack_hi(a, b);
    if (i + 1 < n) {
      *reinterpret_cast<uint32_t*>(hi + i) = h;
      if (lo) *reinterpret_cast<uint32_t*>(lo + i) = pack_lo(a, b, h);
    } else {
      const __half2 hh = *reinterpret_cast<const __half2*>(&h);
      hi[i] = __low2half(hh);
      if (lo) lo[i] = __float2half_rn(a - __half2float(__low2half(hh)));
    }
  }
}

}  // namespace pqlb

using namespace pqlb;

static unsigned long long* g_fwd_h_debug = nullptr;
/* Debug: device buffer of 64 uint64 receiving a clock64 timeline of CTA (0,0) of the CTA-per-tile kernel (NULL = off). */
extern "C" void pqlb_mlp_forward_h_debug(unsigned long long* buf) { g_fwd_h_debug = buf; }
static int g_fwd_h_mode = 0;
/* Tuning / tests: 1 = one CTA per (network, tile), 2 = persistent CTAs pipelining consecutive tiles, 0 = default (1:
 * measured 3718 vs 3653 critic updates/s on the bench, profiles/r2_fwd_schedules.txt). */
extern "C" void pqlb_mlp_forward_h_mode(int mode) { g_fwd_h_mode = (mode == 1 || mode == 2) ? mode : 0; }

extern "C" int pqlb_mlp_forward_h_init(void) {
  cudaError_t e = cudaFuncSetAttribute(mlp_fwd_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_fwd_hw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHSmemWide);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_fwd_hp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHPSmem);
  return e == cudaSuccess ? PQLB_OK : (int)e;
}

extern "C" float pqlb_f16_weight_scale(void) { return kWScale; }

extern "C" int pqlb_split_f16(const float* src, void* hi, void* lo, int64_t n, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(src && hi && n > 0);
  PQLB_CHECK_ALIGN((reinterpret_cast<uintptr_t>(hi) & 3) == 0 && (!lo || (reinterpret_cast<uintptr_t>(lo) & 3) == 0));
  split_f16_kernel<<<grid_for(n, 256, 2), 256, 0, (cudaStream_t)stream>>>(src, reinterpret_cast<__half*>(hi), reinterpret_cast<__half*>(lo), n, kWScale);
  PQLB_LAUNCH_RET();
}

// fp16 matrix [rows][k halves] (row stride ld halves) as a map of 32-bit words: the tiles are plain
// byte copies, so a pair of halves travels as one word
static int make_half_map(CUtensorMap* map, const void* base, int k, int rows, int64_t ld, uint32_t box_rows) {
  if (ld % 8 != 0) return PQLB_E_ALIGN;          // 16-byte row stride
  return make_map(map, reinterpret_cast<const float*>(base), (uint64_t)((k + 1) / 2), (uint64_t)rows, ld / 2, 32, box_rows,
                  CU_TENSOR_MAP_SWIZZLE_128B);
}

extern "C" int pqlb_mlp_forward_h(const pqlb_mlp_h_desc* d, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(d && d->M > 0 && d->k_in > 0 && d->n_groups >= 1 && d->n_groups <= PQLB_MAX_FWD_GROUPS);
  if (d->k_in > kHXKbWide * 32) return PQLB_E_UNSUPPORTED;      // wider inputs take the per-layer path
  { int rc = pqlb_init(); if (rc != PQLB_OK) return rc; }
  static MlpHDev P;
  P.M = d->M;
  P.tile_sync = d->tile_sync;
  P.tiles_m = (d->M + 127) / 128; P.n_groups = d->n_groups;
  P.dbg = g_fwd_h_debug;
  const int tiles_m = (d->M + 127) / 128;
  // any group wider than 128 columns: the whole launch runs the wide-input kernel (its tile maps have short boxes)
  bool wide = false;
  for (int i = 0; i < d->n_groups; ++i) wide = wide || (d->g[i].k_in > 0 ? d->g[i].k_in : d->k_in) > kHXKb * 32;
  const uint32_t box_rows = wide ? (uint32_t)kHRowsWide : 32u;
  for (int i = 0; i < d->n_groups; ++i) {
    const pqlb_mlp_h_group& s = d->g[i];
    MlpHGroupDev& G = P.g[i];
    PQLB_CHECK_ARG(s.terms == 1 || s.terms == 3);
    const int k_in = s.k_in > 0 ? s.k_in : d->k_in;          // groups of one launch may differ in input width
    if (k_in > kHXKbWide * 32) return PQLB_E_UNSUPPORTED;
    G.kb1 = (k_in + 31) / 32; G.kw1 = (k_in + 63) / 64; G.ksteps1 = (k_in + 15) / 16;
    PQLB_CHECK_ARG(s.x && s.w1h && s.w2h && s.w3h && s.b1 && s.b2 && s.b3);
    PQLB_CHECK_ARG(s.terms == 1 || (s.w1l && s.w2l && s.w3l));
    PQLB_CHECK_ARG(!s.q || (s.head_w && s.head_b));
    int rc;
    if ((rc = make_map(&G.tmX, s.x, (uint64_t)k_in, (uint64_t)d->M, s.ldx, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B)) != PQLB_OK) return rc;
    const void* w1[2] = {s.w1h, s.w1l}; const void* w2[2] = {s.w2h, s.w2l}; const void* w3[2] = {s.w3h, s.w3l};
    const void* w4[2] = {s.act_wh, s.act_wl};
    for (int p = 0; p < 2; ++p) {
      const bool have = p == 0 || s.terms == 3;
      if ((rc = make_half_map(&G.tmW1[p], have ? w1[p] : w1[0], k_in, kHH1, s.ldw1, 128)) != PQLB_OK) return rc;
      if ((rc = make_half_map(&G.tmW2[p], have ? w2[p] : w2[0], kHH1, kHH2, kHH1, 128)) != PQLB_OK) return rc;
      if ((rc = make_half_map(&G.tmW3[p], have ? w3[p] : w3[0], kHH2, kHH3, kHH2, 128)) != PQLB_OK) return rc;
    }
    G.st1 = s.h1 != nullptr; G.st2 = s.h2 != nullptr; G.st3 = s.h3 != nullptr;
    if (G.st1 && !make_tile_map(&G.tmH1, s.h1, kHH1, (uint64_t)d->M, kHH1, 1, 0, box_rows)) return PQLB_E_ALIGN;
    if (G.st2 && !make_tile_map(&G.tmH2, s.h2, kHH2, (uint64_t)d->M, kHH2, 1, 0, box_rows)) return PQLB_E_ALIGN;
    if (G.st3 && !make_tile_map(&G.tmH3, s.h3, kHH3, (uint64_t)d->M, kHH3, 1, 0, box_rows)) return PQLB_E_ALIGN;
    if (!G.st1) G.tmH1 = G.tmX;
    if (!G.st2) G.tmH2 = G.tmX;
    if (!G.st3) G.tmH3 = G.tmX;
    G.b1 = s.b1; G.b2 = s.b2; G.b3 = s.b3; G.head_w = s.head_w; G.head_b = s.head_b; G.q = s.q;
    G.terms = s.terms;
    G.act_n = 0; G.act_vec = 1; G.tmW4[0] = G.tmW4[1] = G.tmX;
    G.act_b = nullptr; G.act_noise = nullptr; G.act_out = nullptr; G.act_out2 = nullptr;
    G.sm_n = 0; G.sm_b = nullptr; G.sm_out = nullptr; G.sm_ldp = 0; G.head_rows = 0;
    if (s.sm_wh) {
      // C51 softmax head: sm_n atoms <= 64, probability rows of 64 floats (16-byte aligned)
      PQLB_CHECK_ARG(!s.q && !s.act_wh && s.sm_b && s.sm_out && s.sm_n > 0 && (s.terms == 1 || s.sm_wl));
      if (s.sm_n > 64 || s.sm_ldp < 64 || s.sm_ldp % 4 || !aligned16(s.sm_out)) return PQLB_E_UNSUPPORTED;
      const void* w5[2] = {s.sm_wh, s.sm_wl};
      for (int p = 0; p < 2; ++p) {
        const bool have = p == 0 || s.terms == 3;
        if ((rc = make_half_map(&G.tmW4[p], have ? w5[p] : w5[0], kHH3, s.sm_n, kHH3, 64)) != PQLB_OK) return rc;
      }
      G.sm_n = s.sm_n; G.sm_b = s.sm_b; G.sm_out = s.sm_out; G.sm_ldp = s.sm_ldp; G.head_rows = 64;
    }
    if (s.act_wh) {
      // policy head: act_n a multiple of 4 up to 16, 16-byte aligned output rows; the wide-input kernel also takes
      // up to 32 actions and output rows of any alignment (scalar stores)
      PQLB_CHECK_ARG(!s.q && s.act_b && (s.act_out || s.act_out2) && s.act_n > 0 && (s.terms == 1 || s.act_wl));
      const bool vec = !(s.act_out && (s.act_ldo % 4 || !aligned16(s.act_out))) && !(s.act_out2 && (s.act_ldo2 % 4 || !aligned16(s.act_out2)));
      if (s.act_n > (wide ? 32 : 16) || s.act_n % 4 || (!vec && !wide) || !aligned16(s.act_b) ||
          (s.act_noise && (s.act_ldnoise % 4 || !aligned16(s.act_noise)))) return PQLB_E_UNSUPPORTED;
      const uint32_t head_rows = s.act_n > 16 ? 32u : 16u;
      for (int p = 0; p < 2; ++p) {
        const bool have = p == 0 || s.terms == 3;
        if ((rc = make_half_map(&G.tmW4[p], have ? w4[p] : w4[0], kHH3, s.act_n, kHH3, head_rows)) != PQLB_OK) return rc;
      }
      G.head_rows = (int)head_rows; G.act_vec = vec ? 1 : 0;
      G.act_n = s.act_n; G.act_b = s.act_b; G.act_noise = s.act_noise; G.act_out = s.act_out; G.act_out2 = s.act_out2;
      G.act_ldo = s.act_ldo; G.act_ldo2 = s.act_ldo2; G.act_ldnoise = s.act_ldnoise;
      G.noise_std = s.noise_std; G.noise_bound = s.noise_bound;
    }
    // tile dependencies inside one launch: a group may only wait for a group with a LOWER index (its
    // CTAs are dispatched first), and only a policy group publishes
    PQLB_CHECK_ARG(!s.publish || (s.act_wh && d->tile_sync));
    PQLB_CHECK_ARG(!s.wait || (i > 0 && d->tile_sync));
    G.publish = s.publish != 0; G.wait = s.wait != 0;
  }
  const bool persistent = g_fwd_h_mode == 2 && !wide;      // the persistent schedule keeps the narrow geometry
  const int n_items = tiles_m * d->n_groups;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = persistent ? dim3((unsigned)(n_items < kNumSMs ? n_items : kNumSMs), 1, 1) : dim3((unsigned)tiles_m, (unsigned)d->n_groups, 1);
  cfg.blockDim = dim3(kHThreads, 1, 1);
  cfg.dynamicSmemBytes = persistent ? kHPSmem : (wide ? kHSmemWide : kHSmem);
  cfg.stream = (cudaStream_t)stream;
  cudaError_t le = persistent ? cudaLaunchKernelEx(&cfg, mlp_fwd_hp_kernel, P)
                   : wide     ? cudaLaunchKernelEx(&cfg, mlp_fwd_hw_kernel, P)
                              : cudaLaunchKernelEx(&cfg, mlp_fwd_h_kernel, P);
  PQLB_COUNT_LAUNCH(1);
  return le == cudaSuccess ? PQLB_OK : (int)le;
}
