// K3f'' : the split-fp16 fused forward (see mlp_fwd_h.cu for the number format and the per-tile
// dataflow) as a PERSISTENT kernel: one CTA per SM walks the work items (network instance, 128-row
// tile) w = blockIdx.x, blockIdx.x + gridDim.x, ... and software-pipelines consecutive items.
//
// Why: a single tile is a chain of dependent steps - TMA in, split the input, layer-1 quarter, convert,
// layer-2 chunk, ..., convert Y, layer 3, convert Z, head - measured at 24 us per tile of which the
// tensor pipe is busy 10.5 us (tools/fwd_bench.py); tensor memory (512 columns) holds one tile's
// accumulators, so two tiles cannot simply run side by side.  But the chain's ends can overlap:
//   * the input buffer is free once layer 1's last quarter has been issued, so the next item's input
//     is loaded and split while layers 2 / 3 of the current item run;
//   * after layer 2's last chunk both ping-pong regions P0 / P1 are free and layer 3 needs only one of
//     them (Z): the next item's first quarter is issued BEFORE the current item's layer 3 and fills the
//     tensor pipe while Y is being converted; its second quarter follows the Z conversion;
//   * the per-CTA set-up (barrier init, TMEM allocation, launch) is paid once per SM, not per tile.
// Issue order shared by the TMA producer and the MMA issuer (ring order = consumption order):
//   X(0) L1q0(0) | for item i:  L1q1 L2c0 L1q2 L2c1 L1q3 L2c2 L2c3  X(i+1) L1q0(i+1)  L3  head
// Ping-pong regions alternate per item (s = i & 1): quarters q0, q2 -> P[s], q1, q3 -> P[1-s], Z -> P[s];
// the head's accumulator lives in Y's first columns (Y is dead after layer 3).
// Every mbarrier is used a known number of times per item, each role keeps its own use counters.
#include "mlp_fwd_h.cuh"

namespace pqlb {

namespace {

struct Phases {           // completed uses of each barrier as seen by one role (parity = count & 1)
  uint32_t x_full = 0, x_conv = 0, x_free = 0, p0 = 0, p1 = 0, y_full = 0, y_conv = 0,
           z_full = 0, z_conv = 0, a_full = 0, a_read = 0;
  // parity of the next completion of the ping-pong barrier of region `buf`, then count it (scalars, not an
  // array: a dynamically indexed array would live in local memory)
  __device__ __forceinline__ uint32_t next_p(int buf) {
    const uint32_t par = (buf ? p1 : p0) & 1u;
    if (buf) ++p1; else ++p0;
    return par;
  }
};

}  // namespace

__global__ void __launch_bounds__(kHThreads, 1)
mlp_fwd_hp_kernel(const __grid_constant__ MlpHDev P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t x_full, x_conv, x_free, full_bar[kHStages], empty_bar[kHStages];
  __shared__ __align__(8) uint64_t p_full[2], p_conv[2], y_full, y_conv[2], z_full, z_conv, a_full, a_read;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_q[4][128];

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int n_items = P.tiles_m * P.n_groups;
  const int first = (int)blockIdx.x, step = (int)gridDim.x;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t xs = base;
  const uint32_t ring = xs + kHXBytes;
  const uint32_t stage_buf = ring + kHStages * kHTileBytes;
  float* s_vec = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)) + kHXBytes + kHStages * kHTileBytes + kHEpiWarps * kHChunk);
  constexpr int kVec = kHH1 + kHH2 + 2 * kHH3;        // b1 | b2 | b3 | w4 of one item; two sets

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&x_full), 1); mbar_init(smem_u32(&x_conv), kHEpiWarps); mbar_init(smem_u32(&x_free), 1);
    for (int s = 0; s < kHStages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&p_full[b]), 1); mbar_init(smem_u32(&p_conv[b]), kHEpiWarps); mbar_init(smem_u32(&y_conv[b]), kHEpiWarps); }
    mbar_init(smem_u32(&y_full), 1); mbar_init(smem_u32(&z_full), 1);
    mbar_init(smem_u32(&z_conv), kHEpiWarps); mbar_init(smem_u32(&a_full), 1); mbar_init(smem_u32(&a_read), 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = uniform_u32(tmem_slot);
  const uint32_t tY = tmem + 256u;
  const uint32_t tHead = tY;                       // head accumulator: Y's first columns, dead after layer 3
  constexpr uint32_t idesc = idesc_f16(128);
  const uint64_t desc0 = make_smem_desc(0, 16, 1024, kLayoutSw128);

  if (first < n_items) {
  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0; uint32_t phase = 0;
    Phases ph;
    auto load_tile = [&](const CUtensorMap* map, int c0, int c1) {
      mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(smem_u32(&full_bar[stage]), (uint32_t)kHTileBytes);
        tma_load_2d(ring + stage * kHTileBytes, map, c0, c1, smem_u32(&full_bar[stage]));
      }
      __syncwarp();
      if (++stage == kHStages) { stage = 0; phase ^= 1u; }
    };
    auto load_x = [&](int w, bool have_prev) {
      const MlpHGroupDev& G = P.g[w / P.tiles_m];
      const int tile = w % P.tiles_m;
      if (G.wait) {
        // this item's input rows are produced by the publishing group of the same launch: its items have
        // lower indices, so every CTA reaches them first, and they never wait themselves
        if (lane == 0) {
          const volatile unsigned* f = P.tile_sync + 2 + tile;
          const unsigned epoch = 1u + *reinterpret_cast<const volatile unsigned*>(P.tile_sync + 1);
          const long long t0 = clock64();
          while (*f != epoch) { if (clock64() - t0 > 4000000000LL) __trap(); }
          __threadfence();
        }
        __syncwarp();
        asm volatile("fence.proxy.async.global;" ::: "memory");      // the publisher's generic-proxy stores, read by TMA below
      }
      if (have_prev) { mbar_wait(smem_u32(&x_free), ph.x_free & 1u); ++ph.x_free; }     // layer 1 of the previous item has read X
      if (elect_one()) {
        mbar_expect_tx(smem_u32(&x_full), (uint32_t)G.kb1 * 16384u);
        for (int kb = 0; kb < G.kb1; ++kb) tma_load_2d(xs + kb * 16384, &G.tmX, kb * 32, tile * 128, smem_u32(&x_full));
      }
      __syncwarp();
    };
    auto load_l1 = [&](const MlpHGroupDev& G, int q) {
      const int np = G.terms == 3 ? 2 : 1;
      for (int kw = 0; kw < G.kw1; ++kw)
        for (int part = 0; part < np; ++part) load_tile(&G.tmW1[part], kw * 32, q * 128);
    };
    auto load_l2 = [&](const MlpHGroupDev& G, int c) {
      const int np = G.terms == 3 ? 2 : 1;
      for (int t = 0; t < 4; ++t)
        for (int part = 0; part < np; ++part) load_tile(&G.tmW2[part], c * 64 + (t >> 1) * 32, (t & 1) * 128);
    };
    load_x(first, false);
    load_l1(P.g[first / P.tiles_m], 0);
    for (int w = first; w < n_items; w += step) {
      const MlpHGroupDev& G = P.g[w / P.tiles_m];
      const int np = G.terms == 3 ? 2 : 1;
      load_l1(G, 1); load_l2(G, 0); load_l1(G, 2); load_l2(G, 1); load_l1(G, 3); load_l2(G, 2); load_l2(G, 3);
      // the next item's input and first quarter go in before this item's layer 3 - unless that item waits
      // for a published tile: spinning here would hold back the loads THIS item still needs (and with them
      // the tile another CTA may be waiting for), so a waiting item starts only after the current one
      const bool has_next = w + step < n_items;
      const bool early = has_next && !P.g[(w + step) / P.tiles_m].wait;
      if (early) { load_x(w + step, true); load_l1(P.g[(w + step) / P.tiles_m], 0); }
      for (int t = 0; t < 4; ++t)
        for (int part = 0; part < np; ++part) load_tile(&G.tmW3[part], t * 32, 0);
      if (G.head_rows > 0) {
        const uint32_t box_bytes = (uint32_t)G.head_rows * 128u;
        for (int part = 0; part < np; ++part) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
          const uint32_t bar = smem_u32(&full_bar[stage]);
          const uint32_t dst = ring + stage * kHTileBytes;
          if (elect_one()) {
            mbar_expect_tx(bar, 2u * box_bytes);
            for (int kb = 0; kb < 2; ++kb) tma_load_2d(dst + kb * box_bytes, &G.tmW4[part], kb * 32, 0, bar);
          }
          __syncwarp();
          if (++stage == kHStages) { stage = 0; phase ^= 1u; }
        }
      }
      if (has_next && !early) { load_x(w + step, true); load_l1(P.g[(w + step) / P.tiles_m], 0); }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int stage = 0; uint32_t phase = 0;
    Phases ph;
    auto wait_tile = [&]() -> uint64_t {
      mbar_wait(smem_u32(&full_bar[stage]), phase);
      tcgen05_fence_after();
      return desc0 | (uint64_t)(((ring + stage * kHTileBytes) >> 4) & 0x3FFF);
    };
    auto next_stage = [&]() { __syncwarp(); if (++stage == kHStages) { stage = 0; phase ^= 1u; } };
    // layer-1 quarter q of an item into region P[buf]
    auto mma_l1 = [&](const MlpHGroupDev& G, int buf) {
      const uint32_t tP = tmem + (uint32_t)(buf * 128);
      const int np = G.terms == 3 ? 2 : 1;
      for (int kw = 0; kw < G.kw1; ++kw) {
        for (int part = 0; part < np; ++part) {
          const uint64_t bdesc = wait_tile();
          if (elect_one()) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
              const int k16 = kw * 4 + s;
              if (k16 < G.ksteps1) {
                const uint64_t a_hi = (desc0 | (uint64_t)(((xs + (k16 >> 1) * 16384) >> 4) & 0x3FFF)) + 2u * (uint32_t)(k16 & 1);
                if (part == 0) {
                  umma_f16(tP, a_hi, bdesc + 2u * s, idesc, (uint32_t)(k16 != 0));
                  if (np == 2) umma_f16(tP, a_hi + 4u, bdesc + 2u * s, idesc, 1u);
                } else {
                  umma_f16(tP, a_hi, bdesc + 2u * s, idesc, 1u);
                }
              }
            }
            umma_commit(smem_u32(&empty_bar[stage]));
          }
          next_stage();
        }
      }
      if (elect_one()) umma_commit(smem_u32(&p_full[buf]));
      __syncwarp();
    };
    // layer-2 chunk c: A = converted quarter in P[buf]
    auto mma_l2 = [&](const MlpHGroupDev& G, int c, int buf) {
      const uint32_t tP = tmem + (uint32_t)(buf * 128);
      const int np = G.terms == 3 ? 2 : 1;
      mbar_wait(smem_u32(&p_conv[buf]), ph.next_p(buf));
      tcgen05_fence_after();
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        for (int part = 0; part < np; ++part) {
          const uint64_t bdesc = wait_tile();
          if (elect_one()) {
            const uint32_t d = tY + (uint32_t)((t & 1) * 128);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
              const uint32_t a_hi = tP + (uint32_t)((t >> 1) * 32 + s * 8);
              if (part == 0) {
                umma_f16_ts(d, a_hi, bdesc + 2u * s, idesc, (uint32_t)((c | (t >> 1) | s) != 0));
                if (np == 2) umma_f16_ts(d, a_hi + 64u, bdesc + 2u * s, idesc, 1u);
              } else {
                umma_f16_ts(d, a_hi, bdesc + 2u * s, idesc, 1u);
              }
            }
            umma_commit(smem_u32(&empty_bar[stage]));
          }
          next_stage();
        }
      }
    };

    mbar_wait(smem_u32(&x_conv), ph.x_conv & 1u); ++ph.x_conv;
    tcgen05_fence_after();
    mma_l1(P.g[first / P.tiles_m], 0);
    int it = 0;
    bool prev_head = false;
    for (int w = first; w < n_items; w += step, ++it) {
      const MlpHGroupDev& G = P.g[w / P.tiles_m];
      const int np = G.terms == 3 ? 2 : 1;
      const int s = it & 1;
      // q1 -> P[1-s], the previous item's Z region: its conversion (and the head MMAs, in order) are done
      if (it > 0) { mbar_wait(smem_u32(&z_conv), ph.z_conv & 1u); ++ph.z_conv; tcgen05_fence_after(); }
      mma_l1(G, 1 - s);
      // layer 2 writes Y: the previous item's head accumulator (in Y) has been read
      if (prev_head) { mbar_wait(smem_u32(&a_read), ph.a_read & 1u); ++ph.a_read; tcgen05_fence_after(); }
      mma_l2(G, 0, s);
      mma_l1(G, s);
      mma_l2(G, 1, 1 - s);
      mma_l1(G, 1 - s);
      if (elect_one()) umma_commit(smem_u32(&x_free));        // layer 1 has read the input tile
      __syncwarp();
      mma_l2(G, 2, s);
      mma_l2(G, 3, 1 - s);
      if (elect_one()) umma_commit(smem_u32(&y_full));
      __syncwarp();
      const bool has_next = w + step < n_items;
      const bool early = has_next && !P.g[(w + step) / P.tiles_m].wait;      // same rule as the producer
      if (early) {
        // the next item's first quarter into P[1-s] (free since chunk 3): runs while Y is converted
        mbar_wait(smem_u32(&x_conv), ph.x_conv & 1u); ++ph.x_conv;
        tcgen05_fence_after();
        mma_l1(P.g[(w + step) / P.tiles_m], 1 - s);
      }
      // layer 3: A = converted Y, D = Z = P[s]
      const uint32_t tZ = tmem + (uint32_t)(s * 128);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if ((t & 1) == 0) { mbar_wait(smem_u32(&y_conv[t >> 1]), ph.y_conv & 1u); tcgen05_fence_after(); }
        for (int part = 0; part < np; ++part) {
          const uint64_t bdesc = wait_tile();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t a_hi = tY + (uint32_t)((t >> 1) * 128 + (t & 1) * 32 + k * 8);
              if (part == 0) {
                umma_f16_ts(tZ, a_hi, bdesc + 2u * k, idesc, (uint32_t)((t | k) != 0));
                if (np == 2) umma_f16_ts(tZ, a_hi + 64u, bdesc + 2u * k, idesc, 1u);
              } else {
                umma_f16_ts(tZ, a_hi, bdesc + 2u * k, idesc, 1u);
              }
            }
            umma_commit(smem_u32(&empty_bar[stage]));
          }
          next_stage();
        }
      }
      ++ph.y_conv;
      if (elect_one()) umma_commit(smem_u32(&z_full));
      __syncwarp();
      prev_head = G.head_rows > 0;
      if (prev_head) {
        // head: h3 (split in place in Z) . W4^T into the first head_rows columns of Y
        mbar_wait(smem_u32(&z_conv), ph.z_conv & 1u);          // not consumed: the next item's q1 waits for the same completion
        tcgen05_fence_after();
        const uint32_t idesc_head = G.head_rows == 16 ? idesc_f16(16) : idesc_f16(64);
        const uint32_t box_bytes = (uint32_t)G.head_rows * 128u;
        for (int part = 0; part < np; ++part) {
          const uint64_t bdesc0 = wait_tile();
          if (elect_one()) {
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
              const uint64_t bdesc = bdesc0 + (uint64_t)((kb * box_bytes) >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t a_hi = tZ + (uint32_t)(kb * 32 + k * 8);
                if (part == 0) {
                  umma_f16_ts(tHead, a_hi, bdesc + 2u * k, idesc_head, (uint32_t)((kb | k) != 0));
                  if (np == 2) umma_f16_ts(tHead, a_hi + 64u, bdesc + 2u * k, idesc_head, 1u);
                } else {
                  umma_f16_ts(tHead, a_hi, bdesc + 2u * k, idesc_head, 1u);
                }
              }
            }
            umma_commit(smem_u32(&empty_bar[stage]));
          }
          next_stage();
        }
        if (elect_one()) umma_commit(smem_u32(&a_full));
        __syncwarp();
      }
      if (has_next && !early) {
        mbar_wait(smem_u32(&x_conv), ph.x_conv & 1u); ++ph.x_conv;
        tcgen05_fence_after();
        mma_l1(P.g[(w + step) / P.tiles_m], 1 - s);
      }
    }
  } else {
    // ===================== conversion / epilogue warps =====================
    const int e = warp - 2;
    const int quarter = warp & 3;
    const int chunk = e >> 2;
    const int col = chunk * 32;
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    const uint32_t my_stage = stage_buf + e * kHChunk;
    const uint32_t swz = (uint32_t)(lane & 7) << 4;
    const uint32_t row_off = (uint32_t)lane * 128u;
    const int et = (int)threadIdx.x - 64;            // 0 .. 511
    bool pending = false;
    Phases ph;

    auto load_vec = [&](const MlpHGroupDev& G, int set) {
      float* v = s_vec + set * kVec;
      named_bar_sync(6, 32 * kHEpiWarps);          // every warp is done with the item that used this set two items ago
      for (int j = et; j < kHH1; j += 32 * kHEpiWarps) v[j] = G.b1[j];
      for (int j = et; j < kHH2; j += 32 * kHEpiWarps) v[kHH1 + j] = G.b2[j];
      for (int j = et; j < kHH3; j += 32 * kHEpiWarps) {
        v[kHH1 + kHH2 + j] = G.b3[j];
        v[kHH1 + kHH2 + kHH3 + j] = G.q ? G.head_w[j] : 0.f;
      }
      named_bar_sync(6, 32 * kHEpiWarps);
    };
    // input tile: fp32 -> [hi | lo] fp16, in place (thread = one 128-byte row of one 32-float block)
    auto convert_x = [&](const MlpHGroupDev& G) {
      mbar_wait(smem_u32(&x_full), ph.x_full & 1u); ++ph.x_full;
      const int r = et & 127, kb = et >> 7;
      if (kb < G.kb1) {
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
          lp[j] = pack_lo(f[2 * j], f[2 * j + 1], hp[j]);
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
    };
    auto store_chunk = [&](const float* v, const CUtensorMap* omap, int n_col, int row0) {
      if (pending) { if (elect_one()) bulk_wait_read<0>(); __syncwarp(); }
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4)
        sts128(my_stage + row_off + (((uint32_t)j4 << 4) ^ swz), v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
      fence_proxy_async();
      __syncwarp();
      if (elect_one()) { tma_store_3d(omap, my_stage, n_col, row0, 0); bulk_commit(); }
      pending = true;
    };
    // 128-column accumulator region -> [64 columns of packed hi | 64 columns of packed lo] in place
    auto convert = [&](uint32_t region, const float* bias, int n_base, const CUtensorMap* omap, bool store, uint32_t done,
                       bool lo, int row0) {
      float v[32];
      tmem_ld32(region + lane_sel + (uint32_t)col, v);
      uint32_t hp[16], lp[16];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = elu_fast(fmaf(v[j], kWInv, bias[n_base + col + j]));
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        hp[j] = pack_hi(v[2 * j], v[2 * j + 1]);
        lp[j] = pack_lo(v[2 * j], v[2 * j + 1], hp[j]);
      }
      named_bar_sync(1 + quarter, 128);            // the four warps of this lane quarter have read their columns
      tmem_st16(region + lane_sel + (uint32_t)(chunk * 16), hp);
      if (lo) tmem_st16(region + lane_sel + (uint32_t)(64 + chunk * 16), lp);
      tmem_wait_st();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(done);
      if (store) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = rn_tf32(v[j]);
        store_chunk(v, omap, n_base + col, row0);
      }
    };
    auto conv_quarter = [&](const MlpHGroupDev& G, const float* vec, int q, int buf, int row0) {
      mbar_wait(smem_u32(&p_full[buf]), ph.next_p(buf));
      tcgen05_fence_after();
      convert(tmem + (uint32_t)(buf * 128), vec, q * 128, &G.tmH1, G.st1 != 0, smem_u32(&p_conv[buf]), G.terms == 3, row0);
    };

    {
      const MlpHGroupDev& G0 = P.g[first / P.tiles_m];
      load_vec(G0, 0);
      convert_x(G0);
      conv_quarter(G0, s_vec, 0, 0, (first % P.tiles_m) * 128 + quarter * 32);
    }
    int it = 0;
    for (int w = first; w < n_items; w += step, ++it) {
      const MlpHGroupDev& G = P.g[w / P.tiles_m];
      const int s = it & 1;
      const float* vec = s_vec + (it & 1) * kVec;
      const int m0 = (w % P.tiles_m) * 128;
      const int row0 = m0 + quarter * 32;
      const int row = row0 + lane;
      const bool lo = G.terms == 3;
      const bool has_next = w + step < n_items;
      const MlpHGroupDev& Gn = P.g[has_next ? (w + step) / P.tiles_m : 0];
      const bool early = has_next && !Gn.wait;               // same rule as the producer
      const int row0n = ((w + step) % P.tiles_m) * 128 + quarter * 32;
      if (has_next) load_vec(Gn, (it + 1) & 1);
      conv_quarter(G, vec, 1, 1 - s, row0);
      conv_quarter(G, vec, 2, s, row0);
      conv_quarter(G, vec, 3, 1 - s, row0);
      if (early) convert_x(Gn);
      mbar_wait(smem_u32(&y_full), ph.y_full & 1u); ++ph.y_full;
      tcgen05_fence_after();
      for (int hh = 0; hh < 2; ++hh)
        convert(tY + (uint32_t)(hh * 128), vec + kHH1, hh * 128, &G.tmH2, G.st2 != 0, smem_u32(&y_conv[hh]), lo, row0);
      if (early) conv_quarter(Gn, s_vec + ((it + 1) & 1) * kVec, 0, 1 - s, row0n);
      mbar_wait(smem_u32(&z_full), ph.z_full & 1u); ++ph.z_full;
      tcgen05_fence_after();
      {
        const uint32_t tZ = tmem + (uint32_t)(s * 128);
        const float* b3 = vec + kHH1 + kHH2;
        const float* w4 = b3 + kHH3;
        float qacc = 0.f;
        float v[32];
        tmem_ld32(tZ + lane_sel + (uint32_t)col, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = elu_fast(fmaf(v[j], kWInv, b3[col + j]));
          qacc = fmaf(v[j], w4[col + j], qacc);
        }
        if (G.head_rows > 0) {                     // h3 back into Z as packed halves: the A operand of the head contraction
          uint32_t hp[16], lp[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            hp[j] = pack_hi(v[2 * j], v[2 * j + 1]);
            lp[j] = pack_lo(v[2 * j], v[2 * j + 1], hp[j]);
          }
          named_bar_sync(1 + quarter, 128);
          tmem_st16(tZ + lane_sel + (uint32_t)(chunk * 16), hp);
          if (lo) tmem_st16(tZ + lane_sel + (uint32_t)(64 + chunk * 16), lp);
          tmem_wait_st();
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&z_conv));        // Z has been read (and rewritten for the head): the region may be reused
        if (G.st3) {
          float r[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = rn_tf32(v[j]);
          store_chunk(r, &G.tmH3, col, row0);
        }
        if (G.head_rows > 0 && chunk == 0) {
          mbar_wait(smem_u32(&a_full), ph.a_full & 1u);
          tcgen05_fence_after();
        }
        if (G.head_rows > 0) ++ph.a_full;
        if (G.act_n > 0 && chunk == 0) {
          // one warp per lane quarter finishes the policy head: * 2^-8 + bias, tanh, (+ clipped noise, clamp)
          float a[16];
          tmem_ld16(tHead + lane_sel, a);
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&a_read));
          if (row < P.M) {
            float o[16], o2[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float t = 0.f, r = 0.f;
              if (j < G.act_n) {
                t = tanhf(fmaf(a[j], kWInv, G.act_b[j]));
                r = t;
                if (G.act_noise) {                   // noise.py:19-27 on N(0,1) draws scaled by std
                  const float z = fminf(fmaxf(G.act_noise[(long long)row * G.act_ldnoise + j] * G.noise_std, -G.noise_bound), G.noise_bound);
                  r = fminf(fmaxf(t + z, -1.f), 1.f);
                }
                t = r;
                r = rn_tf32(r);
              }
              o[j] = r; o2[j] = t;
            }
            if (G.act_out) {
              float* dst = G.act_out + (long long)row * G.act_ldo;
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                if (j < G.act_n) *reinterpret_cast<float4*>(dst + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
            }
            if (G.act_out2) {
              float* dst2 = G.act_out2 + (long long)row * G.act_ldo2;
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                if (j < G.act_n) *reinterpret_cast<float4*>(dst2 + j) = make_float4(o2[j], o2[j + 1], o2[j + 2], o2[j + 3]);
            }
          }
          if (G.publish) {
            __threadfence();
            named_bar_sync(5, 128);
            if (quarter == 0 && lane == 0) {
              __threadfence();
              *reinterpret_cast<volatile unsigned*>(P.tile_sync + 2 + (w % P.tiles_m)) =
                  1u + *reinterpret_cast<const volatile unsigned*>(P.tile_sync + 1);
            }
          }
        }
        if (G.sm_n > 0 && chunk == 0) {
          // C51 head: all sm_n <= 64 logits of a row live in this thread (same arithmetic as gemm_tf32.cu's softmax epilogue)
          float l[64];
          tmem_ld32(tHead + lane_sel, l);
          tmem_ld32(tHead + lane_sel + 32u, l + 32);
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&a_read));
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
          s_q[chunk][quarter * 32 + lane] = qacc;
          named_bar_sync(1 + quarter, 128);
          if (chunk == 0 && row < P.M) {
            const int r = quarter * 32 + lane;
            G.q[row] = (((s_q[0][r] + s_q[1][r]) + s_q[2][r]) + s_q[3][r]) + G.head_b[0];
          }
          named_bar_sync(1 + quarter, 128);          // s_q is rewritten by the next item
        }
      }
      if (has_next && !early) { convert_x(Gn); conv_quarter(Gn, s_vec + ((it + 1) & 1) * kVec, 0, 1 - s, row0n); }
    }
    if (pending) { if (elect_one()) bulk_wait_read<0>(); __syncwarp(); }
  }
  }   // first < n_items

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
  if (P.tile_sync && threadIdx.x == 0) {
    // exit ticket: the last CTA of the launch closes the epoch (every CTA has read [1] long before)
    const unsigned total = gridDim.x;
    const unsigned epoch = *reinterpret_cast<const volatile unsigned*>(P.tile_sync + 1);
    __threadfence();
    if (atomicAdd(P.tile_sync, 1u) == total - 1u) {
      *reinterpret_cast<volatile unsigned*>(P.tile_sync) = 0u;
      __threadfence();
      *reinterpret_cast<volatile unsigned*>(P.tile_sync + 1) = epoch + 1u;
    }
  }
}

}  // namespace pqlb
