// Measurement aid, not on the path: tcgen05.mma issue throughput with every SM busy, operands
// resident in shared memory (SS), accumulators in tensor memory - the denominator bench.py uses for
// the tensor-pipe roofline of the dense kernels ("measure tf32").  One CTA per SM, one
// warp issuing `iters` groups of four back-to-back MMAs (M = 128, N = 128 or 256, K = 32 bytes per
// MMA) into two alternating accumulators; nothing is loaded or stored, so the only limiter is the
// tensor pipe itself.  FLOPs = grid * iters * 4 * 2 * 128 * N * K, K = 8 (kind::tf32) / 16 (kind::f16).
#include "tcgen05_utils.cuh"

namespace pqlb {

__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

template <int KIND>
__global__ void __launch_bounds__(64, 1) mma_peak_kernel(int n, int iters) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t done;
  __shared__ uint32_t tmem_slot;
  const int warp = uniform_warp_idx();
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // A: 128 rows x 128 bytes, B: 256 rows x 128 bytes (128-byte swizzled K-major tiles of ones)
  uint32_t* words = reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)));
  const uint32_t one = KIND == 0 ? 0x3F800000u : 0x3C003C00u;
  for (int i = threadIdx.x; i < (128 + 256) * 32; i += blockDim.x) words[i] = one;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&done), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  fence_proxy_async();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = uniform_u32(tmem_slot);
  if (warp == 0) {
    const uint64_t desc0 = make_smem_desc(0, 16, 1024, kLayoutSw128);
    const uint64_t adesc = desc0 | (uint64_t)((base >> 4) & 0x3FFF);
    const uint64_t bdesc = desc0 | (uint64_t)(((base + 128 * 128) >> 4) & 0x3FFF);
    // instruction descriptor: D fp32, A/B tf32 (format 2) or f16 (format 0), K-major, N >> 3, M >> 4
    const uint32_t fmt = KIND == 0 ? 2u : 0u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tmem + (uint32_t)((it & 1) * 256);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (KIND == 0) umma_tf32(d, adesc + 2u * k, bdesc + 2u * k, idesc, (uint32_t)(it > 1 || k > 0));
          else umma_f16_ss(d, adesc + 2u * k, bdesc + 2u * k, idesc, (uint32_t)(it > 1 || k > 0));
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(smem_u32(&done));
    __syncwarp();
    mbar_wait(smem_u32(&done), 0);
    tcgen05_fence_after();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

}  // namespace pqlb

using namespace pqlb;

extern "C" int pqlb_mma_peak(int kind, int n, int iters, pqlb_stream_t stream) {
  PQLB_CHECK_ARG((kind == 0 || kind == 1) && (n == 128 || n == 256) && iters > 0);
  constexpr int smem = 1024 + (128 + 256) * 128;
  static bool init = false;
  if (!init) {
    // a large dynamic allocation keeps it at one CTA per SM (tensor memory is allocated whole)
    cudaError_t e = cudaFuncSetAttribute(mma_peak_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mma_peak_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    init = true;
  }
  (void)smem;
  if (kind == 0) mma_peak_kernel<0><<<kNumSMs, 64, 200 * 1024, (cudaStream_t)stream>>>(n, iters);
  else mma_peak_kernel<1><<<kNumSMs, 64, 200 * 1024, (cudaStream_t)stream>>>(n, iters);
  PQLB_LAUNCH_RET();
}
