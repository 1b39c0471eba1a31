// Ceiling check for K2: an idealised gather of random 800-byte rows out of an 800 MB table into a
// contiguous output (every lane moves 16 B; no field logic), and the same with 1024-byte rows.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
template <int ROW_F4, int U>
__global__ void __launch_bounds__(256) gather_k(const float4* __restrict__ tab, const long* __restrict__ idx, float4* __restrict__ out, long n_rows) {
  const long total = n_rows * ROW_F4, stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride * U) {
    float4 v[U]; const float4* src[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { long it = i + u * stride; src[u] = nullptr; if (it < total) { long r = it / ROW_F4; int s = (int)(it - r * ROW_F4); src[u] = tab + __ldg(idx + r) * ROW_F4 + s; } }
#pragma unroll
    for (int u = 0; u < U; ++u) if (src[u]) v[u] = __ldcs(src[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) if (src[u]) out[i + u * stride] = v[u];
  }
}
template <int ROW_F4> void run(long cap, long n, int blocks) {
  float4 *tab, *out; long* idx; long* h = (long*)malloc(n * 8);
  cudaMalloc(&tab, cap * ROW_F4 * 16); cudaMalloc(&out, n * ROW_F4 * 16); cudaMalloc(&idx, n * 8);
  srand(1); for (long i = 0; i < n; ++i) h[i] = (((long)rand() << 16) ^ rand()) % cap;
  cudaMemcpy(idx, h, n * 8, cudaMemcpyHostToDevice); cudaMemset(tab, 0, cap * ROW_F4 * 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) gather_k<ROW_F4, 8><<<blocks, 256>>>(tab, idx, out, n);
  cudaEventRecord(e0); for (int i = 0; i < 30; ++i) gather_k<ROW_F4, 8><<<blocks, 256>>>(tab, idx, out, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 30;
  printf("row %4d B, %7ld rows, blocks %5d: %8.2f us  %7.1f GB/s (read+write) %s\n", ROW_F4 * 16, n, blocks, ms * 1e3, 2.0 * n * ROW_F4 * 16 / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(tab); cudaFree(out); cudaFree(idx); free(h);
}
int main() {
  for (long n : {65536L, 262144L}) for (int bl : {148 * 8}) { run<50>(1000000, n, bl); run<52>(1000000, n, bl); run<56>(1000000, n, bl); run<64>(1000000, n, bl); run<112>(1000000, n, bl); run<128>(1000000, n, bl); }
  return 0;
}
