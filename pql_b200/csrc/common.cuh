// Shared device/host helpers for libpqlb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pqlb200.h"

#define PQLB_CHECK_ARG(cond) do { if (!(cond)) return PQLB_E_ARG; } while (0)
#define PQLB_CHECK_SHAPE(cond) do { if (!(cond)) return PQLB_E_SHAPE; } while (0)
#define PQLB_CHECK_ALIGN(cond) do { if (!(cond)) return PQLB_E_ALIGN; } while (0)
// Every kernel launch of the library is counted (pqlb_launch_count): bench.py reports the number.
#define PQLB_COUNT_LAUNCH(n) (pqlb::g_launches += (n))
#define PQLB_LAUNCH_RET() do { cudaError_t e__ = cudaGetLastError(); PQLB_COUNT_LAUNCH(1); return e__ == cudaSuccess ? PQLB_OK : (int)e__; } while (0)

namespace pqlb {

constexpr int kNumSMs = 148;  // B200
extern unsigned long long g_launches;
extern int g_rec_stride_mode;       // 0: record stride = next power of two (default); 1: next multiple of 32 words (128 B)

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Record geometry (see pqlb200.h).
struct RecGeom {
  int O, A, obs_pad, act_pad, rec_ld;
  int off_obs, off_next, off_act, off_rew, off_done;
};
inline RecGeom rec_geom(int O, int A) {
  RecGeom g;
  g.O = O; g.A = A;
  g.obs_pad = round_up(O, 4);
  g.act_pad = round_up(A, 4);
  g.off_obs = 0;
  g.off_next = g.obs_pad;
  g.off_act = 2 * g.obs_pad;
  g.off_rew = 2 * g.obs_pad + g.act_pad;
  g.off_done = g.off_rew + 1;
  // Record stride: the next power of two (words).  Measured on B200 (tools/gather_bench.cu,
  // profiles/r1_gather_record_size.txt): gathering random 800-byte records out of an 800 MB table
  // runs at 3.9-4.0 TB/s, 896-byte (128-B aligned) records at 4.2 TB/s, 1024-byte records at
  // 5.8 TB/s - DRAM efficiency collapses when a record straddles a power-of-two block - so the
  // ring trades 28 % more memory (1.02 GB for 1 M AllegroHand slots) for ~1.4x sample bandwidth.
  int ld = 8;
  while (ld < g.off_done + 1) ld <<= 1;
  // measurement switch (pqlb_record_stride_mode): the densest 128-byte-aligned stride instead - 896 B for AllegroHand
  if (g_rec_stride_mode == 1) ld = round_up(g.off_done + 1, 32);
  g.rec_ld = ld;
  return g;
}

// Round-to-nearest (ties away) fp32 -> tf32, kept in an fp32 container.  tcgen05 kind::tf32
// ignores the low 13 mantissa bits (truncation); rounding the operands when they are produced
// turns that systematic shrink into an unbiased 2^-11 rounding (numerics).
__device__ __forceinline__ float rn_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__device__ __forceinline__ float elu1(float v) { return v > 0.f ? v : expm1f(v); }

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

inline int grid_for(int64_t items, int threads, int items_per_thread, int max_blocks_per_sm = 16) {
  int64_t blocks = (items + (int64_t)threads * items_per_thread - 1) / ((int64_t)threads * items_per_thread);
  int64_t cap = (int64_t)kNumSMs * max_blocks_per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace pqlb
