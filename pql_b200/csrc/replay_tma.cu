// K1 (TMA path): ring insert as a 2-D tile mover.  The trajectory columns obs[n][O], next_obs[n][O],
// action[n][A] and the ring [capacity][rec_ld] are described by tensor maps; a tile of kRows
// transitions is TMA-loaded into shared memory (three boxes) and TMA-stored into the record
// columns of its slots (three boxes + one 32-byte box holding reward, done and the zero padding of
// their sector), through a kStages-deep mbarrier ring.  No thread touches the payload, so the
// bytes in flight per SM are bounded by shared memory (2 CTAs x 4 stages x 25.6 KB), not by
// registers: the LDG/STG version (replay.cu) sat at 0.72-0.78 of the copy roofline with 64 KB in
// flight per SM.  Replaces pql/replay/simple_replay.py:40-83 (five slice assignments).
//
// Constraints (otherwise pqlb_ring_insert uses the LDG kernel): O % 4 == 0, A % 4 == 0 (box rows
// are multiples of 16 bytes), O, A <= 256, all pointers 16-byte aligned, and the insert must not
// lap itself (n <= capacity).
#include "tcgen05_utils.cuh"

namespace pqlb {

#ifndef PQLB_INS_ROWS
#define PQLB_INS_ROWS 32
#endif
#ifndef PQLB_INS_STAGES
#define PQLB_INS_STAGES 4
#endif
constexpr int kInsRows = PQLB_INS_ROWS;          // transitions per tile
constexpr int kInsStages = PQLB_INS_STAGES;
constexpr int kInsThreads = 64;       // warp 0: loads, warp 1: reward/done tile + stores

struct alignas(64) InsertPart {
  CUtensorMap tmObs, tmNext, tmAct;                 // sources, rows [row0, row0 + rows)
  CUtensorMap tmRingO, tmRingA, tmRingT;            // ring rows [slot0, slot0 + rows), boxes {O,R} {A,R} {8,R}
  const float* rew; const float* done;              // already offset to row0
  long long rows;
};
struct alignas(64) InsertParams {
  InsertPart part[2];                               // head (slots next_p..) and wrapped tail (slots 0..)
  int n_parts, O, A, off_next, off_act, off_rew;
  int tiles0, tiles_total;                          // tiles of part 0, of both parts
};

__global__ void __launch_bounds__(kInsThreads)
ring_insert_tma_kernel(const __grid_constant__ InsertParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kInsStages], empty_bar[kInsStages];
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t obs_bytes = (uint32_t)(kInsRows * P.O * 4), act_bytes = (uint32_t)(kInsRows * P.A * 4);
  const uint32_t tail_bytes = kInsRows * 32u;
  const uint32_t stage_bytes = 2u * obs_bytes + act_bytes + tail_bytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kInsStages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto part_of = [&](int tile, int& row0) -> const InsertPart& {
    if (tile < P.tiles0) { row0 = tile * kInsRows; return P.part[0]; }
    row0 = (tile - P.tiles0) * kInsRows; return P.part[1];
  };

  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < P.tiles_total; tile += gridDim.x) {
      int row0; const InsertPart& Q = part_of(tile, row0);
      mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
      const uint32_t bar = smem_u32(&full_bar[stage]);
      const uint32_t dst = base + stage * stage_bytes;
      if (elect_one()) {
        mbar_expect_tx(bar, 2u * obs_bytes + act_bytes);       // out-of-range rows of the last tile are zero-filled, still counted
        tma_load_2d(dst, &Q.tmObs, 0, row0, bar);
        tma_load_2d(dst + obs_bytes, &Q.tmNext, 0, row0, bar);
        tma_load_2d(dst + 2u * obs_bytes, &Q.tmAct, 0, row0, bar);
      }
      __syncwarp();
      if (++stage == kInsStages) { stage = 0; phase ^= 1u; }
    }
  } else {
    int stage = 0; uint32_t phase = 0; int issued = 0;
    for (int tile = blockIdx.x; tile < P.tiles_total; tile += gridDim.x) {
      int row0; const InsertPart& Q = part_of(tile, row0);
      const uint32_t src = base + stage * stage_bytes;
      // reward / done / padding sector of row0 + lane: done is stored as the reference's bool column
      constexpr int kPerLane = (kInsRows + 31) / 32;
      float r[kPerLane], d[kPerLane];
#pragma unroll
      for (int u = 0; u < kPerLane; ++u) {
        const long long row = (long long)row0 + lane + 32 * u;
        r[u] = 0.f; d[u] = 0.f;
        if (lane + 32 * u < kInsRows && row < Q.rows) { r[u] = __ldcs(Q.rew + row); d[u] = __ldcs(Q.done + row) != 0.f ? 1.f : 0.f; }
      }
      // the store that last read this stage's buffers must be done before they are rewritten: the
      // producer waits for that too (empty_bar), so by the time full_bar flips the tail tile is free
      mbar_wait(smem_u32(&full_bar[stage]), phase);
#pragma unroll
      for (int u = 0; u < kPerLane; ++u) {
        if (lane + 32 * u < kInsRows) {
          const uint32_t t = src + 2u * obs_bytes + act_bytes + (uint32_t)(lane + 32 * u) * 32u;
          sts128(t, r[u], d[u], 0.f, 0.f); sts128(t + 16u, 0.f, 0.f, 0.f, 0.f);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (elect_one()) {
        tma_store_2d(&Q.tmRingO, src, 0, row0);
        tma_store_2d(&Q.tmRingO, src + obs_bytes, P.off_next, row0);
        tma_store_2d(&Q.tmRingA, src + 2u * obs_bytes, P.off_act, row0);
        tma_store_2d(&Q.tmRingT, src + 2u * obs_bytes + act_bytes, P.off_rew, row0);
        bulk_commit();
      }
      __syncwarp();
      ++issued;
      // free the OLDEST outstanding stage once its stores have read shared memory: at most
      // kInsStages - 1 groups stay in flight
      if (issued >= kInsStages - 1) {
        if (elect_one()) { bulk_wait_read<kInsStages - 2>(); mbar_arrive(smem_u32(&empty_bar[(stage + kInsStages - (kInsStages - 2)) % kInsStages])); }
        __syncwarp();
      }
      if (++stage == kInsStages) { stage = 0; phase ^= 1u; }
    }
    if (elect_one()) bulk_wait_read<0>();                       // shared memory must outlive the last stores
    __syncwarp();
  }
}

}  // namespace pqlb

using namespace pqlb;

static int g_ins_smem = 0;

// Returns PQLB_E_UNSUPPORTED when the shape / alignment needs the LDG kernel instead.
int pqlb_ring_insert_tma(float* ring, int64_t capacity, int obs_dim, int act_dim, const float* obs,
                         const float* action, const float* reward, const float* next_obs,
                         const float* done, int64_t n, int64_t next_p, cudaStream_t stream) {
  if (obs_dim % 4 || act_dim % 4 || obs_dim > 256 || act_dim > 256 || n > capacity) return PQLB_E_UNSUPPORTED;
  if (!aligned16(ring) || !aligned16(obs) || !aligned16(action) || !aligned16(next_obs)) return PQLB_E_UNSUPPORTED;
  if (!get_encode_fn()) return PQLB_E_UNSUPPORTED;
  const RecGeom g = rec_geom(obs_dim, act_dim);
  if ((g.off_rew % 8) != 0) return PQLB_E_UNSUPPORTED;        // the reward/done sector must start a 32-byte sector
  static InsertParams P;
  int64_t head = n, tail = 0;
  if (next_p + n > capacity) { head = capacity - next_p; tail = next_p + n - capacity; }
  const int64_t rows_of[2] = {head, tail};
  const int64_t slot_of[2] = {next_p, 0};
  const int64_t row_of[2] = {0, head};
  P.n_parts = 0; P.tiles0 = 0; P.tiles_total = 0;
  for (int k = 0; k < 2; ++k) {
    if (rows_of[k] <= 0) continue;
    InsertPart& Q = P.part[P.n_parts];
    const uint64_t rows = (uint64_t)rows_of[k];
    const int64_t r0 = row_of[k];
    float* dst = ring + slot_of[k] * (int64_t)g.rec_ld;
    int rc = 0;
    rc |= make_map(&Q.tmObs, obs + r0 * obs_dim, (uint64_t)obs_dim, rows, obs_dim, (uint32_t)obs_dim, kInsRows, CU_TENSOR_MAP_SWIZZLE_NONE);
    rc |= make_map(&Q.tmNext, next_obs + r0 * obs_dim, (uint64_t)obs_dim, rows, obs_dim, (uint32_t)obs_dim, kInsRows, CU_TENSOR_MAP_SWIZZLE_NONE);
    rc |= make_map(&Q.tmAct, action + r0 * act_dim, (uint64_t)act_dim, rows, act_dim, (uint32_t)act_dim, kInsRows, CU_TENSOR_MAP_SWIZZLE_NONE);
    rc |= make_map(&Q.tmRingO, dst, (uint64_t)g.rec_ld, rows, g.rec_ld, (uint32_t)obs_dim, kInsRows, CU_TENSOR_MAP_SWIZZLE_NONE);
    rc |= make_map(&Q.tmRingA, dst, (uint64_t)g.rec_ld, rows, g.rec_ld, (uint32_t)act_dim, kInsRows, CU_TENSOR_MAP_SWIZZLE_NONE);
    rc |= make_map(&Q.tmRingT, dst, (uint64_t)g.rec_ld, rows, g.rec_ld, 8u, kInsRows, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc != PQLB_OK) return PQLB_E_UNSUPPORTED;
    Q.rew = reward + r0; Q.done = done + r0; Q.rows = (long long)rows;
    const int tiles = (int)((rows + kInsRows - 1) / kInsRows);
    if (P.n_parts == 0) P.tiles0 = tiles;
    P.tiles_total += tiles;
    ++P.n_parts;
  }
  if (P.n_parts == 1) P.part[1] = P.part[0];
  P.O = obs_dim; P.A = act_dim; P.off_next = g.off_next; P.off_act = g.off_act; P.off_rew = g.off_rew;
  const int stage_bytes = kInsRows * (2 * obs_dim + act_dim + 8) * 4;
  const int smem = 128 + kInsStages * stage_bytes;
  if (smem > 110 * 1024) return PQLB_E_UNSUPPORTED;          // two CTAs per SM
  if (smem > g_ins_smem) {
    if (cudaFuncSetAttribute(ring_insert_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return PQLB_E_UNSUPPORTED;
    g_ins_smem = smem;
  }
  int blocks = 2 * kNumSMs;
  if (blocks > P.tiles_total) blocks = P.tiles_total;
  ring_insert_tma_kernel<<<blocks, kInsThreads, smem, stream>>>(P);
  PQLB_LAUNCH_RET();
}
