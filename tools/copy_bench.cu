// Ceiling check for the replay kernels: what does a plain float4 copy of the same size reach?
#include <cstdio>
#include <cuda_runtime.h>
template <int U, int HINT>
__global__ void __launch_bounds__(256) copy_k(const float4* __restrict__ src, float4* __restrict__ dst, long n) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) if (i + u * stride < n) v[u] = HINT ? __ldcs(src + i + u * stride) : src[i + u * stride];
#pragma unroll
    for (int u = 0; u < U; ++u) if (i + u * stride < n) { if (HINT) __stcs(dst + i + u * stride, v[u]); else dst[i + u * stride] = v[u]; }
  }
}
template <int U, int HINT> void run(const float4* s, float4* d, long n, int blocks, const char* name) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) copy_k<U, HINT><<<blocks, 256>>>(s, d, n);
  cudaEventRecord(e0);
  for (int i = 0; i < 50; ++i) copy_k<U, HINT><<<blocks, 256>>>(s, d, n);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 50;
  printf("%-28s blocks %5d: %7.2f us  %7.1f GB/s (%s)\n", name, blocks, ms * 1e3, 2.0 * n * 16 / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  for (long mb : {95L, 400L, 2048L}) {
    long n = mb * 1000000 / 16;
    float4 *s, *d; cudaMalloc(&s, n * 16); cudaMalloc(&d, n * 16); cudaMemset(s, 1, n * 16);
    printf("copy of %ld MB (read) + %ld MB (write)\n", mb, mb);
    for (int bl : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
      run<4, 0>(s, d, n, bl, "U4 plain"); run<8, 0>(s, d, n, bl, "U8 plain"); run<4, 1>(s, d, n, bl, "U4 ldcs/stcs"); run<8, 1>(s, d, n, bl, "U8 ldcs/stcs");
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); for (int i = 0; i < 20; ++i) cudaMemcpyAsync(d, s, n * 16, cudaMemcpyDeviceToDevice); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 20;
    printf("cudaMemcpyAsync D2D: %7.2f us %7.1f GB/s\n", ms * 1e3, 2.0 * n * 16 / (ms * 1e-3) / 1e9);
    cudaFree(s); cudaFree(d);
  }
  return 0;
}
