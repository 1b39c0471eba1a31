// Microbenchmark 2: tcgen05.mma kind::tf32 dispatch rate with everything compile-time / uniform.
#include <cstdio>
#include <cuda_runtime.h>
#include "../pql_b200/csrc/tcgen05_utils.cuh"
using namespace pqlb;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <int N, int CHAINS, int ITERS>
__global__ void __launch_bounds__(128, 1) bench(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem_raw)[i] = 0.f;
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  if (threadIdx.x < 32) {
    const uint32_t base = smem_u32(smem_raw);
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
    const uint64_t adesc = make_smem_desc(base, 16, 1024, kLayoutSw128);
    const uint64_t bdesc = make_smem_desc(base + 16384, 16, 1024, kLayoutSw128);
    long long t0 = clock64();
    if (elect_one()) {
#pragma unroll 1
      for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
          for (int c = 0; c < CHAINS; ++c)
            umma_tf32((uint32_t)(c * (512 / CHAINS)), adesc + 2u * k, bdesc + 2u * k, idesc, 1u);
        }
      }
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tcgen05_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tcgen05_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(0u), "r"(512u) : "memory"); }
}

template <int N, int CHAINS>
void run(long long* out) {
  constexpr int ITERS = 32;
  cudaFuncSetAttribute(bench<N, CHAINS, ITERS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  long long h[2];
  for (int rep = 0; rep < 2; ++rep) { bench<N, CHAINS, ITERS><<<1, 128, 64 * 1024>>>(out); cudaDeviceSynchronize(); }
  cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  printf("N=%3d chains=%d: issue %6lld complete %6lld -> %.1f cyc/MMA (ideal %d) %s\n", N, CHAINS, h[0], h[1],
         (double)h[1] / (ITERS * 4 * CHAINS), N / 2, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  long long* out; cudaMalloc(&out, 16);
  run<64, 1>(out); run<64, 2>(out); run<64, 4>(out); run<64, 8>(out);
  run<128, 1>(out); run<128, 2>(out); run<128, 4>(out);
  run<256, 1>(out); run<256, 2>(out);
  return 0;
}
