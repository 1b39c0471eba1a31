// How fast can one SM push activation tiles out to L2 / HBM, and does the path matter?
// The forward / backward epilogues hand every 32-row x 32-column fp32 chunk (4 KB, rows 2 KB apart in
// memory, as in h1 [B][512]) to a TMA store from a swizzled staging buffer.  The step's launch lists show
// those kernels paying ~18 B/clk/SM for their stores; this measures, for 16 warps per CTA doing nothing
// else, (0) exactly that (one staging buffer per warp, wait_group.read 0 before reuse), (1) two staging
// buffers per warp, (2) st.global.v4 with four full 128-byte lines per warp instruction (chunk read back
// row-wise from shared memory), (3) st.global.v4 straight from registers, one row per lane (32 half-filled
// sectors per instruction), (4) 1-D bulk stores of 4 KB contiguous.  Every CTA writes its own region.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -o tools/store_bench tools/store_bench.cu pql_b200/csrc/gemm_tf32.cu ... (see main)
#include <cstdio>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include "../pql_b200/csrc/tcgen05_utils.cuh"
using namespace pqlb;

constexpr int kWarps = 16, kChunk = 4096, kCols = 512, kRowsPerCta = 1024;   // region per CTA: 1024 rows x 512 cols = 2 MB

static PFN_cuTensorMapEncodeTiled enc_fn() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  return reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
}

__device__ __forceinline__ void bulk_store_1d(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(32 * kWarps, 1) store_k(const __grid_constant__ CUtensorMap map, float* base, int mode, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t stage = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t my = stage + warp * 2 * kChunk;
  const uint32_t swz = (uint32_t)(lane & 7) << 4, row_off = (uint32_t)lane * 128u;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = (float)(threadIdx.x + j);
  float* region = base + (size_t)blockIdx.x * kRowsPerCta * kCols;
  __syncthreads();
  const long long t0 = clock64();
  // chunk c of this warp: rows [32 (c / 16 * ... )], simple walk over the region: 32 row groups x 16 column chunks
  for (int it = 0; it < iters; ++it) {
    const int c = (it * kWarps + warp) % (32 * 16);
    const int row0 = (c / 16) * 32, col0 = (c % 16) * 32;
    if (mode <= 1 || mode == 4) {
      const uint32_t buf = my + (uint32_t)((mode == 1) ? (it & 1) * kChunk : 0);
      if (it >= (mode == 1 ? 2 : 1)) { if (elect_one()) { if (mode == 1) bulk_wait_read<1>(); else bulk_wait_read<0>(); } __syncwarp(); }
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) sts128(buf + row_off + (((uint32_t)j4 << 4) ^ swz), v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
      fence_proxy_async();
      __syncwarp();
      if (elect_one()) {
        if (mode == 4) bulk_store_1d(region + (size_t)c * 1024, buf, kChunk);
        else tma_store_3d(&map, buf, col0, blockIdx.x * kRowsPerCta + row0, 0);
        bulk_commit();
      }
      __syncwarp();
    } else if (mode == 2) {
      // through shared memory, then 4 full lines per instruction: lane -> (row 4 j + lane / 8, 16-byte column lane % 8)
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) sts128(my + row_off + (((uint32_t)j4 << 4) ^ swz), v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = 4 * j + (lane >> 3), c16 = lane & 7;
        const float4 x = lds128(my + (uint32_t)r * 128u + (((uint32_t)c16 << 4) ^ ((uint32_t)(r & 7) << 4)));
        *reinterpret_cast<float4*>(region + (size_t)(row0 + r) * kCols + col0 + c16 * 4) = x;
      }
      __syncwarp();
    } else {
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4)
        *reinterpret_cast<float4*>(region + (size_t)(row0 + lane) * kCols + col0 + j4 * 4) = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
    }
  }
  if (mode <= 1 || mode == 4) { if (elect_one()) bulk_wait_read<0>(); __syncwarp(); }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

int main() {
  const size_t bytes = (size_t)148 * kRowsPerCta * kCols * 4;      // 310 MB: also exercises the write-back to HBM
  float* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
  long long* out; cudaMalloc(&out, 148 * 8);
  CUtensorMap map;
  cuuint64_t gdim[3] = {(cuuint64_t)kCols, (cuuint64_t)148 * kRowsPerCta, 1};
  cuuint64_t gstr[2] = {(cuuint64_t)kCols * 4, (cuuint64_t)kCols * 4 * 148 * kRowsPerCta};
  cuuint32_t box[3] = {32, 32, 1}, estr[3] = {1, 1, 1};
  if (enc_fn()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("map failed\n"); return 1; }
  const int smem = kWarps * 2 * kChunk + 1024;
  cudaFuncSetAttribute(store_k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[5] = {"TMA 32x32 box, 1 buffer/warp ", "TMA 32x32 box, 2 buffers/warp", "st.global.v4, 4 lines/instr  ", "st.global.v4, row per lane   ", "bulk 1-D 4 KB contiguous     "};
  for (int ctas : {1, 64, 148}) for (int mode = 0; mode < 5; ++mode) {
    const int iters = 256;             // per warp: 256 chunks = 1 MB; per CTA 16 MB (8 passes over its 2 MB region)
    long long h[148];
    for (int rep = 0; rep < 2; ++rep) { store_k<<<ctas, 32 * kWarps, smem>>>(map, buf, mode, iters, out); cudaDeviceSynchronize(); }
    cudaMemcpy(h, out, ctas * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < ctas; ++i) avg += h[i]; avg /= ctas;
    printf("%3d CTAs, %s: %9.0f cycles for %d KB -> %5.1f B/clk/SM  (%s)\n", ctas, names[mode], avg, iters * kWarps * 4, iters * kWarps * (double)kChunk / avg, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
