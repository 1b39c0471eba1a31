// Device-side replication of the two random draws on the learner path, so that they can be fused
// into the gather kernel instead of being two torch launches per update:
//   torch.randint(cur_capacity, (B,))            pql/replay/simple_replay.py:87, pql_p_learner.py:49
//   torch.normal(zeros, full(std)) = normal_()*std pql/utils/noise.py:20-21
// "Same seeds" parity means: for a given (seed, offset) of a torch CUDA generator the values are
// the ones ATen produces.  ATen's launch policy (ATen/native/cuda/DistributionTemplates.h:50-90,
// SURVEY App. D): T = 256 * min(SMs * maxThreadsPerSM / 256, ceil(numel / 256)) threads, thread t
// runs curand_init(seed, t, offset) and one curand4 / curand_normal4 per loop trip; element
// li < 4T takes component li / T of thread li % T.  randint (range < 2^28): value = u32 % range
// (ATen/core/TransformationHelper.h:42-44); normal_: curand_normal4, * 1.0f + 0.0f.  Each draw
// advances the generator offset by 4 while numel <= 4T.  The Philox / Box-Muller code is cuRAND's
// own device header, the one ATen compiles, so the bits agree by construction.
#pragma once
#include <curand_kernel.h>

#include "common.cuh"

namespace pqlb {

struct RngArgs {
  const long long* state;    // [seed, base_offset, offset_increment_per_update]; NULL = caller-provided draws
  const long long* counter;  // completed updates (device-resident): offset = base + increment * counter
  const long long* range;    // cur_capacity (device-resident)
  float* noise;              // optional N(0,1) draws [noise_numel], taken at offset + 4
  long long noise_numel;
  int threads_idx, threads_noise;   // T of the two draws
};

// T of ATen's launch policy for a tensor of numel elements on the current device; 0 if numel needs
// more than one loop trip (numel > 4T: not supported by the fused path).
inline int aten_rng_threads(long long numel) {
  int dev = 0, sms = 0, tpsm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&tpsm, cudaDevAttrMaxThreadsPerMultiProcessor, dev);
  long long grid = (numel + 255) / 256;
  const long long cap = (long long)sms * (tpsm / 256);
  if (grid > cap) grid = cap;
  const long long T = 256 * grid;
  if (numel > 4 * T || T > 0x7fffffffLL) return 0;
  return (int)T;
}

// thread / component of element li under ATen's policy: (li % T, li / T).  The fused draws have numel <= 4T < 2^32:
// 32-bit division (the 64-bit one is a ~90-instruction subroutine, and there were three of them per index drawn)
__device__ __forceinline__ void aten_thread_of(long long li, int T, unsigned* t, int* c) {
  if (li < (long long)T) { *t = (unsigned)li; *c = 0; }
  else if ((li >> 32) == 0) { const unsigned q = (unsigned)li / (unsigned)T; *c = (int)q; *t = (unsigned)li - q * (unsigned)T; }
  else { *c = (int)(li / T); *t = (unsigned)(li % T); }
}

static __device__ __noinline__ unsigned torch_rand_u32(unsigned long long seed, unsigned long long offset, int T, long long li) {
  unsigned t; int c;
  aten_thread_of(li, T, &t, &c);
  curandStatePhilox4_32_10_t st;
  curand_init(seed, (unsigned long long)t, offset, &st);
  const uint4 r = curand4(&st);
  return c == 0 ? r.x : c == 1 ? r.y : c == 2 ? r.z : r.w;
}

// u32 % range as ATen computes it (uint64 arithmetic), through a 32-bit remainder when range fits
__device__ __forceinline__ long long mod_range(unsigned v, unsigned long long range) {
  return (range >> 32) == 0 ? (long long)(v % (unsigned)range) : (long long)((unsigned long long)v % range);
}

static __device__ __noinline__ float torch_normal_f32(unsigned long long seed, unsigned long long offset, int T, long long li) {
  unsigned t; int c;
  aten_thread_of(li, T, &t, &c);
  curandStatePhilox4_32_10_t st;
  curand_init(seed, (unsigned long long)t, offset, &st);
  const float4 r = curand_normal4(&st);
  return c == 0 ? r.x : c == 1 ? r.y : c == 2 ? r.z : r.w;
}

}  // namespace pqlb
