// K3 building block: grouped GEMM  D[M,N] = epilogue(A . B)  on the 5th-gen tensor cores.
//
//   * operands fp32 in HBM/L2, already RN-rounded to TF32 by their producers; tiles are staged
//     into shared memory by TMA (cp.async.bulk.tensor, 128-byte swizzle) through a
//     full/empty mbarrier ring;
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=tile_n, K=8) with
//     shared-memory descriptors; the fp32 accumulator lives in TMEM;
//   * eight epilogue warps (two per TMEM lane quarter, each owning half of the tile's columns)
//     read the accumulator back with tcgen05.ld (32 lanes x 32 columns per chunk), apply the
//     fused epilogue (bias / ELU / tanh+noise / softmax / ELU' / tanh' / scalar Q head /
//     split-K partial store) and hand every 32x32 chunk to a TMA store through a swizzled
//     shared-memory staging buffer (the operand ring, idle by then), so that global writes are
//     full 128-byte lines; the ELU'/tanh' operand tile is prefetched by TMA the same way;
//   * shared memory is sized for two CTAs per SM, so one tile's epilogue overlaps the next
//     tile's TMA + MMA main loop.
//
// Both operands may be K-major (memory [rows][K], forward + the A side of dgrad) or MN-major
// (memory [K][rows], the weight side of dgrad and both sides of wgrad), so no transposed copy
// of activations or weights is ever materialised.
//
// Replaces: nn.Linear / nn.ELU / tanh / softmax forward and autograd backward launches of
// pql/models/mlp.py:15-24,177-179,197-199,261-263 (cuBLAS fp32 sgemm + ATen elementwise).
#include "tcgen05_utils.cuh"

namespace pqlb {

constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = 64 + 32 * kEpiWarps;   // warp 0: TMA, warp 1: TMEM alloc + MMA, warps 2-9: epilogue
constexpr int kTileM = 128;
constexpr int kTileK = 32;              // fp32 words per k-block = one 128-byte swizzle row
constexpr int kATileBytes = kTileM * 128;
constexpr int kMaxStages = 8;
constexpr int kRingBudget = 96 * 1024;  // operand ring; leaves room for two CTAs per SM
constexpr int kChunkBytes = 32 * 128;   // one warp's 32-row x 32-column staging chunk
constexpr int kSmemMax = 200 * 1024;

struct alignas(64) GroupDev {
  CUtensorMap tmA, tmB, tmA2, tmB2, tmOut, tmAux;
  const float* bias; const float* aux; const float* head_w; const float* head_b;
  float* q; float* out; float* out2;
  long long ldaux, ldo, ldo2, split_stride;
  int out_tma, aux_tma;
};

struct alignas(64) GemmDev {
  GroupDev g[PQLB_MAX_GROUPS];
  int M, N, K, K2;
  int a_mn, b_mn, epi, tile_n;
  int splits, kb1, kb_total, kb_per_split;
  int stages, stage_bytes, b_tile_bytes, tmem_cols;
  int col_lo, col_hi, ring_bytes, aux_bytes;
  int cluster;
  float noise_bound, noise_std;
  unsigned idesc;
};

// ------------------------------------------------------------------------------ the kernel
template <int EPI> struct EpiTraits {
  static constexpr bool kBias = EPI == PQLB_EPI_BIAS || EPI == PQLB_EPI_BIAS_ELU || EPI == PQLB_EPI_BIAS_ELU_HEAD ||
                                EPI == PQLB_EPI_BIAS_TANH || EPI == PQLB_EPI_BIAS_TANH_NOISE || EPI == PQLB_EPI_BIAS_SOFTMAX;
  static constexpr bool kAux = EPI == PQLB_EPI_MUL_ELUGRAD || EPI == PQLB_EPI_MUL_TANHGRAD || EPI == PQLB_EPI_BIAS_TANH_NOISE;
};

template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 2)
gemm_tf32_kernel(const __grid_constant__ GemmDev P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ __align__(8) uint64_t aux_bar[kEpiWarps];
  __shared__ __align__(8) uint64_t aux_free[kEpiWarps];       // all 32 lanes have read their aux chunk out of shared memory
  __shared__ uint32_t tmem_slot;
  __shared__ float s_q[kTileM];

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const int grp = blockIdx.z / P.splits;
  const int split = blockIdx.z - grp * P.splits;
  const GroupDev& G = P.g[grp];
  const int m0 = blockIdx.x * kTileM;
  const int n0 = blockIdx.y * P.tile_n;
  const int kb_begin = split * P.kb_per_split;
  const int kb_end = min(kb_begin + P.kb_per_split, P.kb_total);

  const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B wants 1024-B alignment
  const uint32_t aux_stage = tiles + P.ring_bytes;                 // kEpiWarps chunks (aux_bytes = 0 without an epilogue operand)
  float* s_vec = reinterpret_cast<float*>(smem_raw + (tiles - smem_u32(smem_raw)) + P.ring_bytes + P.aux_bytes);
  float* s_bias = s_vec;                 // [tile_n]
  float* s_head = s_vec + P.tile_n;      // [tile_n]

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    mbar_init(smem_u32(&accum_bar), 1);
    for (int w = 0; w < kEpiWarps; ++w) { mbar_init(smem_u32(&aux_bar[w]), 1); mbar_init(smem_u32(&aux_free[w]), 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"((uint32_t)P.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (EpiTraits<EPI>::kBias || EPI == PQLB_EPI_BIAS_ELU_HEAD) {
    for (int j = threadIdx.x; j < P.tile_n; j += kGemmThreads) {
      const int n = n0 + j;
      s_bias[j] = (EpiTraits<EPI>::kBias && n < P.N) ? G.bias[n] : 0.f;
      if (EPI == PQLB_EPI_BIAS_ELU_HEAD) s_head[j] = n < P.N ? G.head_w[n] : 0.f;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_d = uniform_u32(tmem_slot);

  // Warps 0 and 1 run their loops with all 32 lanes and issue from one elected lane, so every
  // TMA / MMA operand stays in uniform registers (see elect_one()).
  if (warp == 0) {
    // ===================== TMA producer =====================
    const uint32_t tx_bytes = kATileBytes + P.b_tile_bytes;
    int stage = 0; uint32_t phase = 0;
    for (int kb = kb_begin; kb < kb_end; ++kb) {
      mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
      const uint32_t bar = smem_u32(&full_bar[stage]);
      const bool second = kb >= P.kb1;
      const int k0 = (second ? kb - P.kb1 : kb) * kTileK;
      const CUtensorMap* mapA = second ? &G.tmA2 : &G.tmA;
      const CUtensorMap* mapB = second ? &G.tmB2 : &G.tmB;
      const uint32_t sa = tiles + stage * P.stage_bytes;
      const uint32_t sb = sa + kATileBytes;
      if (elect_one()) {
        mbar_expect_tx(bar, tx_bytes);
        if (!P.a_mn) tma_load_2d(sa, mapA, k0, m0, bar);
        else for (int cb = 0; cb < kTileM / 32; ++cb) tma_load_2d(sa + cb * 4096, mapA, m0 + cb * 32, k0, bar);
        if (!P.b_mn) tma_load_2d(sb, mapB, k0, n0, bar);
        else for (int cb = 0; cb < (P.tile_n + 31) / 32; ++cb) tma_load_2d(sb + cb * 4096, mapB, n0 + cb * 32, k0, bar);
      }
      __syncwarp();
      if (++stage == P.stages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int stage = 0; uint32_t phase = 0; uint32_t accumulate = 0;
    // K-major: 8-row groups 1024 B apart (SBO), k-step = 32 B inside the swizzle row.
    // MN-major: 32-element column blocks 4096 B apart (LBO), 4-k-row swizzle atoms 512 B
    // apart (SBO); one MMA (K = 8) consumes two atoms = 1024 B per k-step.
    const uint64_t adesc0 = P.a_mn ? make_smem_desc(0, 4096, 512, kLayoutSw128Base32) : make_smem_desc(0, 16, 1024, kLayoutSw128);
    const uint64_t bdesc0 = P.b_mn ? make_smem_desc(0, 4096, 512, kLayoutSw128Base32) : make_smem_desc(0, 16, 1024, kLayoutSw128);
    const uint32_t a_step = P.a_mn ? (1024u >> 4) : (32u >> 4);
    const uint32_t b_step = P.b_mn ? (1024u >> 4) : (32u >> 4);
    const uint32_t idesc = P.idesc;
    for (int kb = kb_begin; kb < kb_end; ++kb) {
      mbar_wait(smem_u32(&full_bar[stage]), phase);
      tcgen05_fence_after();
      const bool second = kb >= P.kb1;
      const int kseg = second ? P.K2 : P.K;
      const int krem = kseg - (second ? kb - P.kb1 : kb) * kTileK;
      const uint32_t sa = tiles + stage * P.stage_bytes;
      const uint32_t sb = sa + kATileBytes;
      const uint64_t adesc = adesc0 | (uint64_t)((sa >> 4) & 0x3FFF);
      const uint64_t bdesc = bdesc0 | (uint64_t)((sb >> 4) & 0x3FFF);
      if (elect_one()) {
        if (krem >= kTileK) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_tf32(tmem_d, adesc + (uint64_t)(a_step * k), bdesc + (uint64_t)(b_step * k), idesc, k ? 1u : accumulate);
          }
        } else {
          const int ksteps = (krem + 7) / 8;
          for (int k = 0; k < ksteps; ++k)
            umma_tf32(tmem_d, adesc + (uint64_t)(a_step * k), bdesc + (uint64_t)(b_step * k), idesc, k ? 1u : accumulate);
        }
        umma_commit(smem_u32(&empty_bar[stage]));          // frees the smem slot when the MMAs retire
      }
      __syncwarp();
      accumulate = 1;
      if (++stage == P.stages) { stage = 0; phase ^= 1u; }
    }
    if (elect_one()) umma_commit(smem_u32(&accum_bar));       // accumulator complete
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..9) =====================
    if (EPI == PQLB_EPI_STORE && P.cluster > 1) {
      // split-K partials reduced inside the cluster (below, with every warp at the cluster barriers):
      // here only wait until the accumulator is complete, i.e. until this CTA's operand ring is idle
      mbar_wait(smem_u32(&accum_bar), 0);
      tcgen05_fence_after();
    } else {
    const int e = warp - 2;
    const int quarter = warp & 3;                           // TMEM lanes [32q, 32q+32) belong to warp%4 == q
    const int half = e >> 2;                                // which half of the tile's column chunks
    const int row0 = m0 + quarter * 32;
    const int row = row0 + lane;
    const bool row_ok = row < P.M;
    const int chunkw = P.tile_n >= 32 ? 32 : 16;
    const int n_chunks = max(0, (min(P.tile_n, P.col_hi - n0) + 31) / 32);   // chunks that reach below col_hi
    int c_begin, c_end;
    if (EPI == PQLB_EPI_BIAS_SOFTMAX) { c_begin = 0; c_end = half == 0 ? n_chunks : 0; }
    else { const int per = (n_chunks + 1) >> 1; c_begin = half * per; c_end = min(n_chunks, c_begin + per); }
    const bool out_tma = G.out_tma && chunkw == 32;
    const bool aux_tma = EpiTraits<EPI>::kAux && G.aux_tma && chunkw == 32;
    const uint32_t my_aux = aux_stage + e * kChunkBytes;
    const uint32_t my_aux_bar = smem_u32(&aux_bar[e]);
    const uint32_t my_aux_free = smem_u32(&aux_free[e]);
    const uint32_t my_out = tiles + e * 2 * kChunkBytes;    // two staging chunks in the (idle) operand ring
    const uint32_t swz = (uint32_t)(lane & 7) << 4;
    const uint32_t row_off = (uint32_t)lane * 128u;
    uint32_t aux_phase = 0;

    if (aux_tma && c_begin < c_end) {                        // prefetch the first ELU'/tanh' chunk under the main loop
      if (elect_one()) {
        mbar_expect_tx(my_aux_bar, kChunkBytes);
        tma_load_3d(my_aux, &G.tmAux, n0 + c_begin * 32, row0, 0, my_aux_bar);
      }
      __syncwarp();
    }
    mbar_wait(smem_u32(&accum_bar), 0);
    tcgen05_fence_after();
    const uint32_t taddr = tmem_d + ((uint32_t)(quarter * 32) << 16);

    float qacc = 0.f;
    float sm_v[EPI == PQLB_EPI_BIAS_SOFTMAX ? 64 : 1];
    if (EPI == PQLB_EPI_BIAS_SOFTMAX && half == 0) {
      // all N (<= 64) logits of a row live in this thread: softmax in registers
      tmem_ld32(taddr, sm_v);
      if (P.tile_n > 32) tmem_ld32(taddr + 32, sm_v + 32);
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 64; ++j) if (j < P.N) { sm_v[j] += s_bias[j]; mx = fmaxf(mx, sm_v[j]); }
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 64; ++j) if (j < P.N) { sm_v[j] = __expf(sm_v[j] - mx); sum += sm_v[j]; }
      const float inv = 1.f / sum;
#pragma unroll
      for (int j = 0; j < 64; ++j) sm_v[j] = j < P.N ? sm_v[j] * inv : 0.f;
    }

    int n_stores = 0;
    for (int c = c_begin; c < c_end; ++c) {
      const int c0 = c * 32;
      const int nb = n0 + c0;
      if (nb + 32 <= P.col_lo) continue;                       // chunk left of the stored column window
      float v[32];
      if (EPI == PQLB_EPI_BIAS_SOFTMAX) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = sm_v[(c & 1) * 32 + j];
      } else if (chunkw == 32) tmem_ld32(taddr + c0, v);
      else tmem_ld16(taddr + c0, v);

      // ---- second epilogue operand (h for ELU', a for tanh', noise): one 32x32 chunk
      float a[EpiTraits<EPI>::kAux ? 32 : 1];
      if (EpiTraits<EPI>::kAux) {
        if (aux_tma) {
          mbar_wait(my_aux_bar, aux_phase); aux_phase ^= 1u;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 t = lds128(my_aux + row_off + (((uint32_t)j4 << 4) ^ swz));
            a[4 * j4] = t.x; a[4 * j4 + 1] = t.y; a[4 * j4 + 2] = t.z; a[4 * j4 + 3] = t.w;
          }
          // The refill below overwrites the buffer through the async proxy, and nothing orders it
          // after the loads above (they may still sit in the LSU queue when a second CTA shares the
          // SM; __syncwarp compiles to nothing in converged code): every lane releases its reads
          // on aux_free and the refill is issued only once that phase completes.
          mbar_arrive(my_aux_free);
          if (c + 1 < c_end) {                               // next chunk lands while this one is computed
            mbar_wait(my_aux_free, aux_phase ^ 1u);
            if (elect_one()) {
              mbar_expect_tx(my_aux_bar, kChunkBytes);
              tma_load_3d(my_aux, &G.tmAux, nb + 32, row0, 0, my_aux_bar);
            }
            __syncwarp();
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = nb + j;
            a[j] = (row_ok && j < chunkw && n < P.N && n >= P.col_lo && n < P.col_hi) ? G.aux[(long long)row * G.ldaux + n] : 0.f;
          }
        }
      }
      // ---- element-wise epilogue
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = v[j];
        if (EPI == PQLB_EPI_BIAS) x += s_bias[c0 + j];
        else if (EPI == PQLB_EPI_BIAS_ELU) x = rn_tf32(elu_fast(x + s_bias[c0 + j]));
        else if (EPI == PQLB_EPI_BIAS_ELU_HEAD) { const float h = elu_fast(x + s_bias[c0 + j]); qacc = fmaf(h, s_head[c0 + j], qacc); x = rn_tf32(h); }
        else if (EPI == PQLB_EPI_BIAS_TANH) x = tanhf(x + s_bias[c0 + j]);      // rounded below (out2 keeps fp32)
        else if (EPI == PQLB_EPI_BIAS_TANH_NOISE) {
          const float t = tanhf(x + s_bias[c0 + j]);
          const float z = fminf(fmaxf(a[EpiTraits<EPI>::kAux ? j : 0] * P.noise_std, -P.noise_bound), P.noise_bound);
          x = fminf(fmaxf(t + z, -1.f), 1.f);                                   // rounded below (out2 keeps fp32)
        }
        else if (EPI == PQLB_EPI_MUL_ELUGRAD) { const float h = a[EpiTraits<EPI>::kAux ? j : 0]; x = rn_tf32(x * (h > 0.f ? 1.f : h + 1.f)); }
        else if (EPI == PQLB_EPI_MUL_TANHGRAD) { const float t = a[EpiTraits<EPI>::kAux ? j : 0]; x = rn_tf32(x * (1.f - t * t)); }
        v[j] = x;
      }
      if (EPI == PQLB_EPI_BIAS_TANH || EPI == PQLB_EPI_BIAS_TANH_NOISE) {
        if (G.out2 && row_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) { const int n = nb + j; if (j < chunkw && n < P.N) G.out2[(long long)row * G.ldo2 + n] = v[j]; }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = rn_tf32(v[j]);
      }
      // ---- store the chunk
      if (G.out == nullptr) continue;
      if (out_tma) {
        const uint32_t buf = my_out + (uint32_t)(n_stores & 1) * kChunkBytes;
        if (n_stores >= 2) { if (elect_one()) bulk_wait_read<1>(); __syncwarp(); }   // staging buffer free again
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
          sts128(buf + row_off + (((uint32_t)j4 << 4) ^ swz), v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
        fence_proxy_async();
        __syncwarp();
        if (elect_one()) { tma_store_3d(&G.tmOut, buf, nb - P.col_lo, row0, split); bulk_commit(); }
        ++n_stores;
      } else if (row_ok) {
        float* orow = G.out + (long long)row * G.ldo + (EPI == PQLB_EPI_STORE ? (long long)split * G.split_stride : 0ll);
        const int lo = max(nb, P.col_lo), hi = min(min(nb + chunkw, P.N), P.col_hi);
        float* dst = orow + (nb - P.col_lo);
        if (lo == nb && hi == nb + chunkw && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (j >= chunkw) break;
            *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = nb + j;
            if (j < chunkw && n >= lo && n < hi) dst[j] = v[j];
          }
        }
      }
    }
    if (EPI == PQLB_EPI_BIAS_ELU_HEAD) {
      // the two warps of a lane quarter each hold the dot product over their half of the columns
      if (half == 1) s_q[quarter * 32 + lane] = qacc;
      named_bar_sync(1 + quarter, 64);
      if (half == 0 && row_ok) G.q[row] = (qacc + s_q[quarter * 32 + lane]) + G.head_b[0];
    }
    if (n_stores > 0) { if (elect_one()) bulk_wait_read<0>(); __syncwarp(); }   // staging memory must outlive the bulk reads
    }
  }

  if (EPI == PQLB_EPI_STORE && P.cluster > 1) {
    // ---- split-K reduction through distributed shared memory.  The tile's accumulator is 4 lane
    // quarters x tile_n / 32 chunks of 32 x 32 words; chunk j belongs to the CTA of cluster rank j % C.
    // Every CTA sends each of its chunks to the owner's receive area (the idle operand ring:
    // slot [j / C][sender]), the owner adds the C versions in sender (= split) order and stores one
    // partial per cluster: C times fewer bytes leave the SMs, and the reduction kernel reads C times
    // fewer.  Fixed order => bit-reproducible.
    const int C = P.cluster;
    const uint32_t rank = cluster_ctarank();
    const int NC = P.tile_n >> 5;
    const int n_chunks = max(0, (min(P.tile_n, P.col_hi - n0) + 31) / 32);
    const uint32_t swz = (uint32_t)(lane & 7) << 4;
    const uint32_t row_off = (uint32_t)lane * 128u;
    cluster_arrive_release();          // #0: my ring is idle (epilogue warps arrive after accum_bar) ...
    cluster_wait_acquire();            //     ... and so is everybody else's
    if (warp >= 2) {
      const int e = warp - 2, quarter = warp & 3, half = e >> 2;
      const int per = (n_chunks + 1) >> 1;
      const int c_begin = half * per, c_end = min(n_chunks, c_begin + per);
      const uint32_t taddr = tmem_d + ((uint32_t)(quarter * 32) << 16);
      for (int c = c_begin; c < c_end; ++c) {
        float v[32];
        tmem_ld32(taddr + c * 32, v);
        const int j = quarter * NC + c;
        const uint32_t slot = tiles + (uint32_t)((j / C) * C + (int)rank) * kChunkBytes;
        const uint32_t dst = mapa_shared(slot, (uint32_t)(j % C)) + row_off;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
          sts128_cluster(dst + (((uint32_t)j4 << 4) ^ swz), v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
      }
    }
    cluster_arrive_release();          // #1: my chunks are delivered ...
    cluster_wait_acquire();            //     ... and all chunks I own have arrived
    if (warp >= 2) {
      const int e = warp - 2;
      const int owned = (4 * NC + C - 1 - (int)rank) / C;        // chunks j = rank + C * t < 4 * NC
      for (int t = e; t < owned; t += kEpiWarps) {
        const int j = (int)rank + C * t;
        const int quarter = j / NC, c = j - quarter * NC;
        if (c >= n_chunks) continue;
        float acc[32];
        const uint32_t base = tiles + (uint32_t)(t * C) * kChunkBytes + row_off;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 x = lds128(base + (((uint32_t)j4 << 4) ^ swz));
          acc[4 * j4] = x.x; acc[4 * j4 + 1] = x.y; acc[4 * j4 + 2] = x.z; acc[4 * j4 + 3] = x.w;
        }
        for (int s2 = 1; s2 < C; ++s2) {
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 x = lds128(base + (uint32_t)s2 * kChunkBytes + (((uint32_t)j4 << 4) ^ swz));
            acc[4 * j4] += x.x; acc[4 * j4 + 1] += x.y; acc[4 * j4 + 2] += x.z; acc[4 * j4 + 3] += x.w;
          }
        }
        const int row = m0 + quarter * 32 + lane;
        const int nb = n0 + c * 32;
        if (row < P.M) {
          float* dst = G.out + (long long)(split / C) * G.split_stride + (long long)row * G.ldo + nb;
          const int hi = min(nb + 32, P.col_hi);
          if (hi == nb + 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
            for (int q = 0; q < 32; q += 4) *reinterpret_cast<float4*>(dst + q) = make_float4(acc[q], acc[q + 1], acc[q + 2], acc[q + 3]);
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) if (nb + q < hi) dst[q] = acc[q];
          }
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)P.tmem_cols) : "memory");
  }
}

typedef void (*GemmKernel)(const GemmDev);
static GemmKernel kernel_for(int epi) {
  switch (epi) {
    case PQLB_EPI_STORE: return gemm_tf32_kernel<PQLB_EPI_STORE>;
    case PQLB_EPI_BIAS: return gemm_tf32_kernel<PQLB_EPI_BIAS>;
    case PQLB_EPI_BIAS_ELU: return gemm_tf32_kernel<PQLB_EPI_BIAS_ELU>;
    case PQLB_EPI_BIAS_ELU_HEAD: return gemm_tf32_kernel<PQLB_EPI_BIAS_ELU_HEAD>;
    case PQLB_EPI_BIAS_TANH: return gemm_tf32_kernel<PQLB_EPI_BIAS_TANH>;
    case PQLB_EPI_BIAS_TANH_NOISE: return gemm_tf32_kernel<PQLB_EPI_BIAS_TANH_NOISE>;
    case PQLB_EPI_BIAS_SOFTMAX: return gemm_tf32_kernel<PQLB_EPI_BIAS_SOFTMAX>;
    case PQLB_EPI_MUL_ELUGRAD: return gemm_tf32_kernel<PQLB_EPI_MUL_ELUGRAD>;
    case PQLB_EPI_MUL_TANHGRAD: return gemm_tf32_kernel<PQLB_EPI_MUL_TANHGRAD>;
    default: return nullptr;
  }
}

// ------------------------------------------------------------------------------ host side
PFN_cuTensorMapEncodeTiled get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
  }
  return fn;
}

// 2-D fp32 tensor map with 128-byte swizzle; dim0 is the contiguous dimension.
int make_map(CUtensorMap* map, const float* base, uint64_t dim0, uint64_t dim1, int64_t ld_words,
                    uint32_t box0, uint32_t box1, CUtensorMapSwizzle swizzle) {
  PFN_cuTensorMapEncodeTiled enc = get_encode_fn();
  if (!enc) return PQLB_E_DRIVER;
  if (!aligned16(base) || (ld_words % 4) != 0 || ld_words < (int64_t)dim0) return PQLB_E_ALIGN;
  cuuint64_t gdim[2] = {dim0, dim1};
  cuuint64_t gstr[1] = {(cuuint64_t)ld_words * 4};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? PQLB_OK : PQLB_E_DRIVER;
}

// Output / epilogue-operand map: [splits][rows][cols] fp32, 32x32 boxes, 128-byte swizzle (matches
// the staging chunks the epilogue warps write / read).  Returns false when TMA cannot address it.
bool make_tile_map(CUtensorMap* map, const float* base, uint64_t cols, uint64_t rows, int64_t ld_words,
                          uint64_t splits, int64_t split_stride_words, uint32_t box_rows) {
  PFN_cuTensorMapEncodeTiled enc = get_encode_fn();
  if (!enc || !base || !aligned16(base) || (ld_words % 4) != 0 || ld_words < (int64_t)cols) return false;
  if (splits > 1 && ((split_stride_words % 4) != 0 || split_stride_words <= 0)) return false;
  cuuint64_t gdim[3] = {cols, rows, splits};
  cuuint64_t gstr[2] = {(cuuint64_t)ld_words * 4, (cuuint64_t)(splits > 1 ? split_stride_words : ld_words * (int64_t)rows) * 4};
  cuuint32_t box[3] = {32, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int make_operand_map(CUtensorMap* map, const float* base, int64_t ld, int major, int rows, int k,
                            int tile_rows) {
  if (major == PQLB_K_MAJOR)   // memory [rows][k]
    return make_map(map, base, (uint64_t)k, (uint64_t)rows, ld, kTileK, (uint32_t)tile_rows, CU_TENSOR_MAP_SWIZZLE_128B);
  // memory [k][rows]: boxes of 32 rows(MN) x 32 k, 32-byte swizzle atoms (see make_smem_desc)
  return make_map(map, base, (uint64_t)rows, (uint64_t)k, ld, 32, kTileK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
}

}  // namespace pqlb

using namespace pqlb;

extern "C" int pqlb_mlp_forward_init(void);
extern "C" int pqlb_mlp_forward_h_init(void);
extern "C" int pqlb_mlp_backward_init(void);
extern "C" int pqlb_wgrad_init(void);

// One-time, non-stream setup (opt-in shared memory size, driver entry point) so that nothing but
// kernel launches happens while a caller is capturing a CUDA graph.  Per device.
extern "C" int pqlb_init(void) {
  static bool done[64] = {false};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev < 64 && done[dev]) return PQLB_OK;
  // static (barriers) + dynamic shared memory must stay <= 227 KB
  for (int epi = PQLB_EPI_STORE; epi <= PQLB_EPI_MUL_TANHGRAD; ++epi) {
    e = cudaFuncSetAttribute(kernel_for(epi), cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
    if (e != cudaSuccess) return (int)e;
  }
  if (!get_encode_fn()) return PQLB_E_DRIVER;
  { int rc = pqlb_mlp_forward_init(); if (rc != PQLB_OK) return rc; }
  { int rc = pqlb_mlp_forward_h_init(); if (rc != PQLB_OK) return rc; }
  { int rc = pqlb_mlp_backward_init(); if (rc != PQLB_OK) return rc; }
  { int rc = pqlb_wgrad_init(); if (rc != PQLB_OK) return rc; }
  if (dev < 64) done[dev] = true;
  return PQLB_OK;
}

extern "C" int pqlb_gemm_tf32(const pqlb_gemm_desc* d, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(d != nullptr);
  PQLB_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0 && d->K2 >= 0);
  PQLB_CHECK_ARG(d->n_groups >= 1 && d->n_groups <= PQLB_MAX_GROUPS && d->splits >= 1);
  PQLB_CHECK_ARG(d->epilogue >= PQLB_EPI_STORE && d->epilogue <= PQLB_EPI_MUL_TANHGRAD);
  const int tn = d->tile_n;
  PQLB_CHECK_ARG(tn == 16 || tn == 32 || tn == 64 || tn == 128 || tn == 256);
  PQLB_CHECK_ARG(d->splits == 1 || d->epilogue == PQLB_EPI_STORE);
  if (d->epilogue == PQLB_EPI_BIAS_ELU_HEAD) PQLB_CHECK_SHAPE(d->N <= tn);
  if (d->epilogue == PQLB_EPI_BIAS_SOFTMAX) PQLB_CHECK_SHAPE(d->N <= tn && tn <= 64 && tn >= 32);

  static GemmDev P;   // host staging copy (single-threaded callers, SURVEY §8b threading)
  P.M = d->M; P.N = d->N; P.K = d->K; P.K2 = d->K2;
  P.a_mn = d->a_major == PQLB_MN_MAJOR; P.b_mn = d->b_major == PQLB_MN_MAJOR;
  P.epi = d->epilogue; P.tile_n = tn; P.splits = d->splits;
  P.cluster = d->cluster > 1 ? d->cluster : 1;
  if (P.cluster > 1) {
    PQLB_CHECK_ARG(d->epilogue == PQLB_EPI_STORE && (P.cluster == 2 || P.cluster == 4 || P.cluster == 8));
    PQLB_CHECK_SHAPE(d->splits % P.cluster == 0 && tn >= 32 && d->col_lo == 0);
  }
  P.kb1 = (d->K + kTileK - 1) / kTileK;
  P.kb_total = P.kb1 + (d->K2 + kTileK - 1) / kTileK;
  PQLB_CHECK_SHAPE(P.kb_total % d->splits == 0);
  P.kb_per_split = P.kb_total / d->splits;
  P.b_tile_bytes = P.b_mn ? ((tn + 31) / 32) * 4096 : tn * 128;
  P.stage_bytes = kATileBytes + ((P.b_tile_bytes + 1023) / 1024) * 1024;
  const bool wants_aux = d->epilogue == PQLB_EPI_MUL_ELUGRAD || d->epilogue == PQLB_EPI_MUL_TANHGRAD ||
                         d->epilogue == PQLB_EPI_BIAS_TANH_NOISE;
  // Split-K weight gradients with 256-wide tiles stream 48 KB per k-block: two stages (what fits
  // next to a second CTA) cannot keep the L2 -> SM pipe full, so these launches take the whole SM
  // (one CTA, four stages).
  const bool big_ring = d->epilogue == PQLB_EPI_STORE && d->splits > 1 && tn == 256;
  int stages = ((big_ring ? 2 * kRingBudget : kRingBudget) - (wants_aux ? kEpiWarps * kChunkBytes : 0)) / P.stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages > P.kb_per_split) stages = P.kb_per_split < 2 ? 2 : P.kb_per_split;
  if (stages < 2) stages = 2;
  P.stages = stages;
  // the operand ring doubles as the epilogue's staging area: two 4 KB chunks per epilogue warp
  P.ring_bytes = stages * P.stage_bytes;
  if (P.ring_bytes < 2 * kEpiWarps * kChunkBytes) P.ring_bytes = 2 * kEpiWarps * kChunkBytes;
  P.aux_bytes = wants_aux ? kEpiWarps * kChunkBytes : 0;
  P.tmem_cols = tn < 32 ? 32 : tn;
  P.col_lo = d->col_lo; P.col_hi = d->col_hi > 0 ? d->col_hi : d->N;
  PQLB_CHECK_SHAPE(P.col_lo >= 0 && P.col_lo < P.col_hi && P.col_hi <= d->N);
  P.noise_bound = d->noise_bound; P.noise_std = d->noise_std;
  // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6), a/b=TF32 [7,10)/[10,13),
  // a_major [15], b_major [16], N>>3 [17,23), M>>4 [24,29)
  P.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)P.a_mn << 15) | ((unsigned)P.b_mn << 16) |
            ((unsigned)(tn >> 3) << 17) | ((unsigned)(kTileM >> 4) << 24);

  for (int i = 0; i < d->n_groups; ++i) {
    const pqlb_gemm_group& s = d->g[i];
    GroupDev& G = P.g[i];
    PQLB_CHECK_ARG(s.a && s.b);
    int rc;
    if ((rc = make_operand_map(&G.tmA, s.a, s.lda, d->a_major, d->M, d->K, kTileM)) != PQLB_OK) return rc;
    if ((rc = make_operand_map(&G.tmB, s.b, s.ldb, d->b_major, d->N, d->K, tn)) != PQLB_OK) return rc;
    if (d->K2 > 0) {
      PQLB_CHECK_ARG(s.a2 && s.b2);
      if ((rc = make_operand_map(&G.tmA2, s.a2, s.lda2, d->a_major, d->M, d->K2, kTileM)) != PQLB_OK) return rc;
      if ((rc = make_operand_map(&G.tmB2, s.b2, s.ldb2, d->b_major, d->N, d->K2, tn)) != PQLB_OK) return rc;
    } else { G.tmA2 = G.tmA; G.tmB2 = G.tmB; }
    const int e = d->epilogue;
    if (e != PQLB_EPI_STORE && e != PQLB_EPI_MUL_ELUGRAD && e != PQLB_EPI_MUL_TANHGRAD) PQLB_CHECK_ARG(s.bias);
    if (e == PQLB_EPI_MUL_ELUGRAD || e == PQLB_EPI_MUL_TANHGRAD || e == PQLB_EPI_BIAS_TANH_NOISE) PQLB_CHECK_ARG(s.aux);
    if (e == PQLB_EPI_BIAS_ELU_HEAD) PQLB_CHECK_ARG(s.head_w && s.head_b && s.q); else PQLB_CHECK_ARG(s.out);
    G.bias = s.bias; G.aux = s.aux; G.head_w = s.head_w; G.head_b = s.head_b; G.q = s.q;
    G.out = s.out; G.out2 = s.out2; G.ldaux = s.ldaux; G.ldo = s.ldo; G.ldo2 = s.ldo2;
    G.split_stride = s.split_stride;
    // TMA paths of the epilogue (fall back to per-thread accesses when the layout is not addressable)
    const int n_store = P.col_hi - P.col_lo;
    G.out_tma = s.out && P.col_lo == 0 && make_tile_map(&G.tmOut, s.out, (uint64_t)n_store, (uint64_t)d->M, s.ldo,
                                       (uint64_t)d->splits, s.split_stride) ? 1 : 0;
    G.aux_tma = wants_aux && P.col_lo == 0 && P.col_hi == d->N &&
                make_tile_map(&G.tmAux, s.aux, (uint64_t)d->N, (uint64_t)d->M, s.ldaux, 1, 0) ? 1 : 0;
    if (!G.out_tma) G.tmOut = G.tmA;
    if (!G.aux_tma) G.tmAux = G.tmA;
  }

  const int smem_bytes = 1024 + P.ring_bytes + P.aux_bytes + 2 * tn * (int)sizeof(float);
  PQLB_CHECK_SHAPE(smem_bytes <= kSmemMax);
  { int rc = pqlb_init(); if (rc != PQLB_OK) return rc; }
  dim3 grid((unsigned)((d->M + kTileM - 1) / kTileM), (unsigned)((d->N + tn - 1) / tn), (unsigned)(d->n_groups * d->splits));
  if (P.cluster > 1) {
    PQLB_CHECK_SHAPE(4 * (tn / 32) * kChunkBytes <= P.ring_bytes);        // the receive area is the operand ring
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(kGemmThreads); cfg.dynamicSmemBytes = (size_t)smem_bytes; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = (unsigned)P.cluster;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, kernel_for(d->epilogue), P);
    PQLB_COUNT_LAUNCH(1);
    if (le != cudaSuccess) { (void)cudaGetLastError(); return (int)le; }
    return PQLB_OK;
  }
  kernel_for(d->epilogue)<<<grid, kGemmThreads, smem_bytes, (cudaStream_t)stream>>>(P);
  PQLB_LAUNCH_RET();
}
