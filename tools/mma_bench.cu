// Microbenchmark: issue rate of tcgen05.mma kind::tf32 (M=128) as a function of N, operand source
// (SS: A from smem, TS: A from TMEM) and the number of independent accumulators interleaved.
#include <cstdio>
#include <cuda_runtime.h>
#include "../pql_b200/csrc/tcgen05_utils.cuh"
using namespace pqlb;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

template <bool uniform>
__global__ void __launch_bounds__(128, 1) bench(int n, int ts, int chains, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem_raw)[i] = 0.f;
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tmem = uniform ? __shfl_sync(0xffffffffu, slot, 0) : slot;
  if (uniform ? (threadIdx.x < 32) : (threadIdx.x == 0)) {
    const bool bf = ts == 2;
    const uint32_t fmt = bf ? 1u : 2u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
    const uint64_t adesc = make_smem_desc(base, 16, 1024, kLayoutSw128);
    const uint64_t bdesc = make_smem_desc(base + 16384, 16, 1024, kLayoutSw128);
    const int cw = 512 / 4;     // accumulator regions of 128 columns (n <= 128) ; for n=256 use 2 regions
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int c = 0; c < chains; ++c) {
        const uint32_t d = tmem + (uint32_t)(c * (n > 128 ? 256 : cw));
        if (!uniform || elect_one()) {
          if (bf) umma_bf16(d, adesc + 2u * (it & 3), bdesc + 2u * (it & 3), idesc, it > 0);
          else if (ts) umma_tf32_ts(d, tmem + (uint32_t)(((c + 1) % 4) * cw) + (it & 3) * 8, bdesc + 2u * (it & 3), idesc, it > 0);
          else umma_tf32(d, adesc + 2u * (it & 3), bdesc + 2u * (it & 3), idesc, it > 0);
        }
        if (uniform) __syncwarp();
      }
    }
    long long t1 = clock64();
    if (!uniform || elect_one()) umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tcgen05_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tcgen05_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory"); }
}

int main() {
  long long* out; cudaMalloc(&out, 16);
  cudaFuncSetAttribute(bench<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(bench<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 64;
  for (int uniform = 0; uniform < 2; ++uniform)
  for (int ts = 0; ts < 3; ++ts)
    for (int n : {128, 256})
      for (int chains : {1, 2, 3, 4}) {
        if (n == 256 && chains > 2) continue;
        if (ts == 1 && n == 256 && chains > 1) continue;
        long long h[2];
        for (int rep = 0; rep < 2; ++rep) { if (uniform) bench<true><<<1, 128, 64 * 1024>>>(n, ts, chains, iters, out); else bench<false><<<1, 128, 64 * 1024>>>(n, ts, chains, iters, out); cudaDeviceSynchronize(); }
        cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaGetLastError();
        printf("%s %s N=%3d chains=%d: issue %6lld cyc, complete %6lld cyc -> %.1f cyc/MMA (ideal %d)  %s\n", uniform ? "warp-uniform" : "one-thread  ", ts == 2 ? "BF16 SS" : (ts ? "TF32 TS" : "TF32 SS"), n, chains,
               h[0], h[1], (double)h[1] / (iters * chains), n / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
