// What limits the L2 -> shared-memory stream of the TF32 weight tiles: the bytes, or the number of
// 128-byte rows a tensor-map box is made of?  Each CTA streams the same 4 MB region (L2 resident)
// through an 8-stage ring of 16 KB slots: (a) 2-D tensor-map boxes of 128 rows x 128 B (128-byte
// swizzle), exactly as the kernels load a weight tile; (b) 1-D bulk copies of 16 KB contiguous.
#include <cstdio>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include "../pql_b200/csrc/tcgen05_utils.cuh"
using namespace pqlb;
constexpr int kStages = 8, kTile = 16384;
__global__ void __launch_bounds__(64, 1) stream_k(const __grid_constant__ CUtensorMap map, const float* base, int mode, int n_tiles, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages];
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  if (threadIdx.x == 0) { for (int s = 0; s < kStages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  long long t0 = clock64();
  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int t = 0; t < n_tiles; ++t) {
      mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
      if (elect_one()) {
        const uint32_t bar = smem_u32(&full_bar[stage]);
        mbar_expect_tx(bar, kTile);
        const int tt = (t + blockIdx.x * 7) % 256;           // 256 tiles = 4 MB
        if (mode == 0) tma_load_2d(ring + stage * kTile, &map, (tt & 15) * 32, (tt >> 4) * 128, bar);
        else bulk_load_1d(ring + stage * kTile, base + (long long)tt * (kTile / 4), kTile, bar);
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1u; }
    }
  } else {
    int stage = 0; uint32_t phase = 0;
    for (int t = 0; t < n_tiles; ++t) {
      mbar_wait(smem_u32(&full_bar[stage]), phase);
      if (elect_one()) mbar_arrive(smem_u32(&empty_bar[stage]));
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1u; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}
int main() {
  float* buf; cudaMalloc(&buf, 4 << 20); cudaMemset(buf, 0, 4 << 20);
  long long* out; cudaMalloc(&out, 148 * 8);
  CUtensorMap map;
  if (make_map(&map, buf, 512, 2048, 512, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B) != 0) { printf("map failed\n"); return 1; }
  cudaFuncSetAttribute(stream_k, cudaFuncAttributeMaxDynamicSharedMemorySize, kStages * kTile + 1024);
  for (int ctas : {1, 16, 64, 148}) for (int mode = 0; mode < 2; ++mode) {
    const int n_tiles = 512;
    long long h[148];
    for (int rep = 0; rep < 2; ++rep) { stream_k<<<ctas, 64, kStages * kTile + 1024>>>(map, buf, mode, n_tiles, out); cudaDeviceSynchronize(); }
    cudaMemcpy(h, out, ctas * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < ctas; ++i) avg += h[i]; avg /= ctas;
    printf("%3d CTAs, %s: %8.0f cycles for %d tiles -> %.1f B/clk/SM  (%s)\n", ctas, mode ? "1-D bulk 16 KB      " : "2-D box 128 x 128 B ", avg, n_tiles, n_tiles * (double)kTile / avg, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
