// K3w: every weight gradient of one update in ONE launch.
//
//   dW_l[n_out, n_in] = dz_l^T . h_{l-1}      (contraction over the batch, K = B rows)
//
// for all layers of all networks of the update (twin-Q critic: 2 nets x 3 layers, actor: 4 layers; the
// C51 head adds one).  The outputs are tiny (434 K words for both critics) and K is huge (8192), so the
// work is split along K; a problem's (tile, split) pairs are the CTAs of the launch - one CTA per SM,
// one wave - and each CTA leaves one fp32 partial tile that pqlb_grad_reduce sums in split order
// (bit-reproducible).  The split counts are chosen by the caller so that every CTA streams about the
// same number of operand bytes: these GEMMs are bound by the L2 -> SM ingest (48 B/clk/SM measured,
// profiles/r1_tma_ingest_microbench.txt; a 128 x 256 TF32 tile wants 87 B/clk), not by the tensor pipe.
// Round 1 launched one grouped GEMM per layer with power-of-two splits sized to fill the SMs on its own
// (64 / 16 / 32 splits, 64 MB of partials per critic update, K loops of 4-16 blocks behind a prologue
// and an epilogue of the same length); here a critic update is 148 CTAs with 6-10 splits, 15.7 MB of
// partials and K loops of 26-43 blocks.
//
// Both operands are MN-major ([K][rows] in memory: the activations and back-propagated errors as the
// forward / backward kernels left them), staged by TMA in 32 x 32 boxes (128-byte rows, 32-byte swizzle
// atoms) through a full/empty mbarrier ring that takes the whole SM's shared memory; one elected thread
// issues tcgen05.mma.kind::tf32 (M = 128, N = tile_n, K = 8), the accumulator lives in tensor memory;
// eight epilogue warps move it through swizzled staging chunks (the idle operand ring) to TMA stores.
// Replaces: the weight-gradient half of loss.backward() (pql/algo/pql_v_learner.py:125,
// pql/algo/pql_p_learner.py:60: autograd's addmm backward, one cuBLAS sgemm per nn.Linear).
#include "tcgen05_utils.cuh"

namespace pqlb {

constexpr int kWgEpiWarps = 8;
constexpr int kWgThreads = 64 + 32 * kWgEpiWarps;   // warp 0: TMA, warp 1: TMEM alloc + MMA, warps 2-9: epilogue
constexpr int kWgRing = 192 * 1024;                  // operand ring (one CTA per SM)
constexpr int kWgMaxStages = 8;
constexpr int kWgChunk = 32 * 128;                   // one warp's 32-row x 32-column staging chunk

struct alignas(64) WgProblemDev {
  CUtensorMap tmA, tmB, tmOut;
  int M, N, tile_n, splits;
  int m_tiles, n_tiles, item_begin, stages;
  int stage_bytes, tmem_cols;
  unsigned idesc;
  int pad;
};

struct alignas(64) WgDev {
  WgProblemDev p[PQLB_MAX_WGRAD];
  int n_problems, kb_total, n_items, pad;
};

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_multi_kernel(const __grid_constant__ WgDev P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kWgMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kWgMaxStages];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_slot;

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  int pi = 0;
#pragma unroll
  for (int i = 1; i < PQLB_MAX_WGRAD; ++i)
    if (i < P.n_problems && (int)blockIdx.x >= P.p[i].item_begin) pi = i;
  const WgProblemDev& Q = P.p[pi];
  const int local = (int)blockIdx.x - Q.item_begin;
  const int split = local % Q.splits;
  const int tile = local / Q.splits;
  const int m0 = (tile % Q.m_tiles) * 128;
  const int n0 = (tile / Q.m_tiles) * Q.tile_n;
  // uneven splits: split s owns k-blocks [s * kb / S, (s + 1) * kb / S)
  const int kb_begin = (int)(((long long)split * P.kb_total) / Q.splits);
  const int kb_end = (int)(((long long)(split + 1) * P.kb_total) / Q.splits);
  const int stages = Q.stages;
  const int b_boxes = Q.tile_n >> 5;

  const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    mbar_init(smem_u32(&accum_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"((uint32_t)Q.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_d = uniform_u32(tmem_slot);

  if (warp == 0) {
    // ===================== TMA producer =====================
    const uint32_t tx_bytes = 16384u + (uint32_t)b_boxes * 4096u;
    int stage = 0; uint32_t phase = 0;
    for (int kb = kb_begin; kb < kb_end; ++kb) {
      mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
      const uint32_t bar = smem_u32(&full_bar[stage]);
      const uint32_t sa = tiles + stage * Q.stage_bytes;
      const uint32_t sb = sa + 16384u;
      const int k0 = kb * 32;
      if (elect_one()) {
        mbar_expect_tx(bar, tx_bytes);
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) tma_load_2d(sa + cb * 4096, &Q.tmA, m0 + cb * 32, k0, bar);
        for (int cb = 0; cb < b_boxes; ++cb) tma_load_2d(sb + cb * 4096, &Q.tmB, n0 + cb * 32, k0, bar);
      }
      __syncwarp();
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // MN-major operands: 32-element column blocks 4096 B apart (LBO), 4-k-row swizzle atoms 512 B apart
    // (SBO); one MMA (K = 8) consumes two atoms = 1024 B
    int stage = 0; uint32_t phase = 0; uint32_t accumulate = 0;
    const uint64_t desc0 = make_smem_desc(0, 4096, 512, kLayoutSw128Base32);
    const uint32_t idesc = Q.idesc;
    for (int kb = kb_begin; kb < kb_end; ++kb) {
      mbar_wait(smem_u32(&full_bar[stage]), phase);
      tcgen05_fence_after();
      const uint32_t sa = tiles + stage * Q.stage_bytes;
      const uint64_t adesc = desc0 | (uint64_t)((sa >> 4) & 0x3FFF);
      const uint64_t bdesc = desc0 | (uint64_t)(((sa + 16384u) >> 4) & 0x3FFF);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_tf32(tmem_d, adesc + (uint64_t)(64u * k), bdesc + (uint64_t)(64u * k), idesc, k ? 1u : accumulate);
        umma_commit(smem_u32(&empty_bar[stage]));
      }
      __syncwarp();
      accumulate = 1;
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
    if (elect_one()) umma_commit(smem_u32(&accum_bar));
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..9): the partial tile, 32 x 32 chunk by chunk =====================
    const int e = warp - 2;
    const int quarter = warp & 3;                 // TMEM lanes [32q, 32q + 32) belong to warps with warp % 4 == q
    const int half = e >> 2;
    const int row0 = m0 + quarter * 32;
    const int n_chunks = (min(Q.tile_n, Q.N - n0) + 31) >> 5;
    const int per = (n_chunks + 1) >> 1;
    const int c_begin = half * per, c_end = min(n_chunks, c_begin + per);
    const uint32_t my_out = tiles + (uint32_t)e * 2u * kWgChunk;       // two staging chunks in the (idle) operand ring
    const uint32_t swz = (uint32_t)(lane & 7) << 4;
    const uint32_t row_off = (uint32_t)lane * 128u;
    mbar_wait(smem_u32(&accum_bar), 0);
    tcgen05_fence_after();
    const uint32_t taddr = tmem_d + ((uint32_t)(quarter * 32) << 16);
    int n_stores = 0;
    if (row0 < Q.M) {
      for (int c = c_begin; c < c_end; ++c) {
        float v[32];
        tmem_ld32(taddr + c * 32, v);
        const uint32_t buf = my_out + (uint32_t)(n_stores & 1) * kWgChunk;
        if (n_stores >= 2) { if (elect_one()) bulk_wait_read<1>(); __syncwarp(); }
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
          sts128(buf + row_off + (((uint32_t)j4 << 4) ^ swz), v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
        fence_proxy_async();
        __syncwarp();
        if (elect_one()) { tma_store_3d(&Q.tmOut, buf, n0 + c * 32, row0, split); bulk_commit(); }
        ++n_stores;
      }
    }
    if (n_stores > 0) { if (elect_one()) bulk_wait_read<0>(); __syncwarp(); }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)Q.tmem_cols) : "memory");
  }
}

}  // namespace pqlb

using namespace pqlb;

extern "C" int pqlb_wgrad_init(void) {
  cudaError_t e = cudaFuncSetAttribute(wgrad_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgRing + 1024);
  return e == cudaSuccess ? PQLB_OK : (int)e;
}

extern "C" int pqlb_wgrad_multi(const pqlb_wgrad_desc* d, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(d != nullptr && d->K > 0 && d->n_problems >= 1 && d->n_problems <= PQLB_MAX_WGRAD);
  { int rc = pqlb_init(); if (rc != PQLB_OK) return rc; }
  static WgDev P;     // host staging copy (single-threaded callers, like pqlb_gemm_tf32)
  P.n_problems = d->n_problems;
  P.kb_total = (d->K + 31) / 32;
  int items = 0;
  for (int i = 0; i < d->n_problems; ++i) {
    const pqlb_wgrad_problem& s = d->p[i];
    WgProblemDev& Q = P.p[i];
    PQLB_CHECK_ARG(s.dz && s.h && s.part && s.M > 0 && s.N > 0 && s.splits >= 1);
    const int tn = s.tile_n;
    PQLB_CHECK_ARG(tn == 32 || tn == 64 || tn == 128 || tn == 256);
    PQLB_CHECK_SHAPE(s.splits <= P.kb_total && s.lddz >= s.M && s.ldh >= s.N && s.ldo >= s.N);
    PQLB_CHECK_SHAPE(s.splits == 1 || s.split_stride >= (int64_t)s.M * s.ldo);
    Q.M = s.M; Q.N = s.N; Q.tile_n = tn; Q.splits = s.splits;
    Q.m_tiles = (s.M + 127) / 128; Q.n_tiles = (s.N + tn - 1) / tn;
    Q.item_begin = items;
    items += Q.m_tiles * Q.n_tiles * s.splits;
    Q.stage_bytes = 16384 + (tn >> 5) * 4096;
    int stages = kWgRing / Q.stage_bytes;
    if (stages > kWgMaxStages) stages = kWgMaxStages;
    Q.stages = stages;
    Q.tmem_cols = tn;
    // instruction descriptor: c = F32 [4,6), a / b = TF32 [7,10) / [10,13), both MN-major [15] [16], N >> 3 [17,23), M >> 4 [24,29)
    Q.idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((unsigned)(tn >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
    int rc;
    // memory [K][rows]: boxes of 32 rows x 32 k, 32-byte swizzle atoms (what kLayoutSw128Base32 reads)
    if ((rc = make_map(&Q.tmA, s.dz, (uint64_t)s.M, (uint64_t)d->K, s.lddz, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) != PQLB_OK) return rc;
    if ((rc = make_map(&Q.tmB, s.h, (uint64_t)s.N, (uint64_t)d->K, s.ldh, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) != PQLB_OK) return rc;
    if (!make_tile_map(&Q.tmOut, s.part, (uint64_t)s.N, (uint64_t)s.M, s.ldo, (uint64_t)s.splits, s.split_stride)) return PQLB_E_ALIGN;
  }
  P.n_items = items;
  // 16 staging chunks of the epilogue live in the ring
  static_assert(2 * kWgEpiWarps * kWgChunk <= kWgRing, "staging");
  wgrad_multi_kernel<<<(unsigned)items, kWgThreads, kWgRing + 1024, (cudaStream_t)stream>>>(P);
  PQLB_LAUNCH_RET();
}
