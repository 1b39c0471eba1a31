// Shared by the two split-fp16 forward kernels (mlp_fwd_h.cu: one CTA per tile; mlp_fwd_hp.cu:
// persistent CTAs that overlap the tail of one tile with the head of the next): device-side
// descriptors, tile geometry and the kind::f16 / tensor-memory PTX wrappers.
#pragma once
#include "f16split.cuh"
#include "tcgen05_utils.cuh"

namespace pqlb {

constexpr int kHH1 = 512, kHH2 = 256, kHH3 = 128;
constexpr int kHEpiWarps = 16;
constexpr int kHThreads = 64 + 32 * kHEpiWarps;
constexpr int kHStages = 5;
constexpr int kHTileBytes = 128 * 128;        // every weight tile: 128 rows x 64 halves
constexpr int kHXKb = 4;                      // input width <= 128 floats: four 32-float blocks
constexpr int kHXBytes = kHXKb * 128 * 128;
constexpr int kHChunk = 32 * 128;
constexpr int kHPrefetchCtas = 2;             // row tiles per group that issue the L2 prefetch of the group's weights
constexpr int kHSmem = 1024 + kHXBytes + kHStages * kHTileBytes + kHEpiWarps * kHChunk + (kHH1 + kHH2 + 2 * kHH3) * 4;
// Wide inputs (129..256 columns: the ShadowHand critics, obs 211 + act 20): the input tile is eight 32-float
// blocks = 128 KB, paid for with a four-stage weight ring and 8-row staging chunks (a warp's 32 x 32 chunk leaves in
// four TMA stores).  Measured on the ShadowHand loop: 4 stages / 8 rows 92 us per four-critic launch, 3 stages /
// 16 rows 97 us (the ring depth is what the L2 -> SM weight stream needs to stay in flight).
#ifndef PQLB_WIDE_STAGES
#define PQLB_WIDE_STAGES 4
#endif
#ifndef PQLB_WIDE_ROWS
#define PQLB_WIDE_ROWS 8
#endif
constexpr int kHXKbWide = 8;
constexpr int kHStagesWide = PQLB_WIDE_STAGES;
constexpr int kHRowsWide = PQLB_WIDE_ROWS;         // rows of a staging chunk: a warp's 32 rows leave in 32 / kHRowsWide TMA stores
constexpr int kHChunkWide = kHRowsWide * 128;
static_assert(kHRowsWide == 8 || kHRowsWide == 16 || kHRowsWide == 32, "staging chunks are whole 8-row swizzle groups");
constexpr int kHSmemWide = 1024 + kHXKbWide * 128 * 128 + kHStagesWide * kHTileBytes + kHEpiWarps * kHChunkWide + (kHH1 + kHH2 + 2 * kHH3) * 4;
static_assert(kHSmemWide <= 227 * 1024, "wide-input tile does not fit shared memory");

struct alignas(64) MlpHGroupDev {
  CUtensorMap tmX, tmW1[2], tmW2[2], tmW3[2], tmW4[2], tmH1, tmH2, tmH3;      // [0] hi, [1] lo
  const float* b1; const float* b2; const float* b3; const float* head_w; const float* head_b;
  float* q;
  const float* act_b; const float* act_noise; float* act_out; float* act_out2;
  long long act_ldo, act_ldo2, act_ldnoise;
  float noise_std, noise_bound;
  int act_n;
  const float* sm_b; float* sm_out; long long sm_ldp; int sm_n;      // fused softmax head (C51): sm_n atoms <= 64
  int head_rows;                    // rows of the head weight tile / N of the head MMA: 16 or 32 (policy), 64 (softmax), 0 = no head
  int act_vec;                      // policy head output rows are 16-byte aligned (float4 stores); 0: scalar stores (wide-input kernel only)
  int st1, st2, st3;
  int terms;
  int kb1, kw1, ksteps1;            // this group's input width in 32-float blocks / 64-half weight blocks / 16-wide k steps
  int publish, wait;                // tile dependencies inside the launch: this group publishes / waits for tile_sync flags
};
struct alignas(64) MlpHDev {
  MlpHGroupDev g[PQLB_MAX_FWD_GROUPS];
  int M;
  // Optional [2 + row tiles] words, zero-initialised once by the caller: [0] exit ticket, [1] number of
  // launches completed so far, [2 + t] = (launches completed + 1) once the publishing group has written
  // the rows of tile t in THIS launch.  The last CTA to leave advances [1], so no reset between launches.
  unsigned* tile_sync;
  int tiles_m, n_groups;            // persistent kernel: work item w -> group w / tiles_m, row tile w % tiles_m
  unsigned long long* dbg;          // optional clock64 timeline of CTA 0 (pqlb_mlp_forward_h_debug): [0,32) MMA issuer, [32,64) epilogue warp 0
};

__host__ __device__ constexpr uint32_t idesc_f16(int n) {
  // D fp32 (bits 4-5 = 1), A / B fp16 (format 0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
  return (1u << 4) | ((unsigned)(n >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand in tensor memory: lane = row, one 32-bit column = two consecutive k (low half first)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
constexpr int kHPSmem = kHSmem + (kHH1 + kHH2 + 2 * kHH3) * 4;      // persistent kernel: two sets of bias / head vectors

__global__ void mlp_fwd_hp_kernel(const __grid_constant__ MlpHDev P);

}  // namespace pqlb
