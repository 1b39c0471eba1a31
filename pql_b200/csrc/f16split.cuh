// Split-fp16 operand format of the precise forward (mlp_fwd_h.cu): v ~ hi + lo, hi = fp16(v),
// lo = fp16(v - hi).  Weights are stored scaled by 2^8 so that both halves stay in fp16's normal
// range (|w| ~ 0.05 -> w_lo ~ 1e-5 would be subnormal); the consumer multiplies by 2^-8 (exact).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace pqlb {

constexpr float kWScale = 256.f, kWInv = 1.f / 256.f;

// (hi, lo) halves of two floats, packed low element first; fp16 saturates at +-65504 instead of overflowing
__device__ __forceinline__ uint32_t pack_hi(float a, float b) {
  const __half2 h = __floats2half2_rn(fminf(fmaxf(a, -65504.f), 65504.f), fminf(fmaxf(b, -65504.f), 65504.f));
  return *reinterpret_cast<const uint32_t*>(&h);
}
// the same for values bounded below (ELU outputs are > -1): only the upper clamp
__device__ __forceinline__ uint32_t pack_hi_pos(float a, float b) {
  const __half2 h = __floats2half2_rn(fminf(a, 65504.f), fminf(b, 65504.f));
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_lo(float a, float b, uint32_t hi) {
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
  return *reinterpret_cast<const uint32_t*>(&l);
}
// four consecutive arena elements -> hi halves at base[i4 .. i4+4), lo halves at base[n + i4 .. ), scaled
__device__ __forceinline__ void store_split4(__half* base, int64_t n, int64_t i4, const float* p) {
  const float a = p[0] * kWScale, b = p[1] * kWScale, c = p[2] * kWScale, d = p[3] * kWScale;
  const uint32_t h0 = pack_hi(a, b), h1 = pack_hi(c, d);
  *reinterpret_cast<uint2*>(base + i4) = make_uint2(h0, h1);
  *reinterpret_cast<uint2*>(base + n + i4) = make_uint2(pack_lo(a, b, h0), pack_lo(c, d, h1));
}

}  // namespace pqlb
