// K4: deterministic gradient reduction, global-norm clip, AdamW and Polyak in one pass over a
// flat fp32 parameter arena (HBM/L2-bound elementwise work).  Reference behaviour:
// pql/algo/pql_v_learner.py:124-133 (zero_grad/backward/clip_grad_norm_/AdamW.step),
// pql/utils/torch_util.py:9-12 (soft_update); arithmetic order: torch/optim/adam.py and
// torch/nn/utils/clip_grad.py as summarised in SURVEY.md App. E.
#include <math.h>

#include "common.cuh"
#include "f16split.cuh"

namespace pqlb {

constexpr int kOptThreads = 256;
constexpr int kSegElems = 1024;   // elements per reduce segment (4 per thread)

// Fixed-shape block reduction: xor-shuffle tree inside each warp, then warp 0 adds the eight
// warp sums in index order.  Same launch shape => same association => bit-reproducible.
__device__ __forceinline__ float block_sum_256(float v, float* smem8) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) smem8[w] = v;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < kOptThreads / 32; ++i) tot += smem8[i];
  __syncthreads();
  return tot;
}

struct AdamScalars {
  float decay, one_minus_b1, b2, one_minus_b2, step_size, bc2_sqrt, eps, tau, one_minus_tau;
};
// scalar corrections in double, like torch's python-side arithmetic (torch/optim/adam.py), t 1-based
__device__ __forceinline__ AdamScalars adam_scalars(double t, float lr_f, float beta1, float beta2, float eps, float wd, float tau) {
  AdamScalars sa;
  const double b1 = beta1, b2 = beta2, lr = lr_f;
  const double bc1 = 1.0 - pow(b1, t), bc2 = 1.0 - pow(b2, t);
  sa.decay = (float)(1.0 - lr * (double)wd);
  sa.one_minus_b1 = (float)(1.0 - b1); sa.b2 = beta2; sa.one_minus_b2 = (float)(1.0 - b2);
  sa.step_size = (float)(lr / bc1); sa.bc2_sqrt = (float)sqrt(bc2); sa.eps = eps;
  sa.tau = tau; sa.one_minus_tau = (float)(1.0 - (double)tau);
  return sa;
}

// Work the LAST block of the reduction does for the rest of the update (it used to be a launch of
// its own plus a serial section at the top of every AdamW block): the loss = scale * sum of the loss
// partials (fixed order), written to out[0] and to the Tracker window ring[count % ring_len], and the
// AdamW bias corrections of step count + 1.  count (completed updates) is NOT advanced here: the
// optimiser kernel does that when it has applied the step.
struct FinishArgs {
  const float* loss_part; int n_loss; float loss_scale; float* loss_out;
  const long long* counter; float* ring; int ring_len;
  float lr, beta1, beta2, eps, weight_decay, tau;
  AdamScalars* scalars_out;
};

// One block per segment of <= 256 elements (one per thread): the n_part partial sums of an
// element are loaded sixteen at a time (independent loads in flight) and added in index order, so
// the result does not depend on the launch shape.
__global__ void __launch_bounds__(kOptThreads)
grad_reduce_kernel(const int64_t* __restrict__ seg, const float* __restrict__ ws,
                   float* __restrict__ grad, float* __restrict__ sumsq_part, bool reduce, FinishArgs fin) {
  __shared__ float red[8];
  const int64_t* s = seg + (int64_t)blockIdx.x * 5;
  const int64_t arena_off = s[0], count = s[1], ws_off = s[2], ws_stride = s[3], n_part = s[4];
  float sq = 0.f;
  for (int i = threadIdx.x; i < count; i += kOptThreads) {
    float g;
    if (reduce && n_part > 0) {
      g = 0.f;
      const float* p = ws + ws_off + i;
      int64_t k = 0;
      for (; k + 16 <= n_part; k += 16) {
        float t[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) t[u] = p[(k + u) * ws_stride];
#pragma unroll
        for (int u = 0; u < 16; ++u) g += t[u];
      }
      for (; k + 8 <= n_part; k += 8) {
        float t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = p[(k + u) * ws_stride];
#pragma unroll
        for (int u = 0; u < 8; ++u) g += t[u];
      }
      for (; k < n_part; ++k) g += p[k * ws_stride];
      grad[arena_off + i] = g;
    } else {
      g = grad[arena_off + i];
    }
    sq += g * g;
  }
  const float tot = block_sum_256(sq, red);
  if (threadIdx.x == 0) sumsq_part[blockIdx.x] = tot;

  if (fin.loss_part != nullptr && blockIdx.x == gridDim.x - 1) {
    float acc = 0.f;
    for (int i = threadIdx.x; i < fin.n_loss; i += kOptThreads) acc += fin.loss_part[i];
    const float ltot = block_sum_256(acc, red);
    if (threadIdx.x == 0) {
      const long long c = fin.counter[0];
      const float r = ltot * fin.loss_scale;
      fin.loss_out[0] = r;
      if (fin.ring && fin.ring_len > 0) fin.ring[c % fin.ring_len] = r;
      *fin.scalars_out = adam_scalars((double)(c + 1), fin.lr, fin.beta1, fin.beta2, fin.eps, fin.weight_decay, fin.tau);
    }
  }
}

struct AdamHyper {     // as passed by the caller; derived scalars are computed on the device
  float grad_scale, max_norm, lr, beta1, beta2, eps, weight_decay, tau;
};
__global__ void __launch_bounds__(kOptThreads)
adamw_polyak_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m,
                    float* __restrict__ v, float* __restrict__ target, float* __restrict__ p_tf32,
                    float* __restrict__ t_tf32, __half* __restrict__ p_h, __half* __restrict__ t_h,
                    int64_t n, const float* __restrict__ sumsq_part,
                    int n_part, AdamHyper h, int64_t step_host, const int64_t* __restrict__ step_dev,
                    float* __restrict__ grad_norm_out, const AdamScalars* __restrict__ scal_dev,
                    long long* __restrict__ counter_inc) {
  __shared__ float red[8];
  __shared__ AdamScalars sa;
  // the element loads do not depend on the norm or the bias corrections: issue them first, so that
  // they are in flight under the partial-sum reduction and thread 0's double-precision pow()
  const int64_t i4 = ((int64_t)blockIdx.x * kOptThreads + threadIdx.x) * 4;
  const int cnt = i4 >= n ? 0 : ((n - i4) >= 4 ? 4 : (int)(n - i4));
  float p[4], g[4], mm[4], vv[4], tt[4];
  if (cnt == 4) {
    *reinterpret_cast<float4*>(p) = *reinterpret_cast<const float4*>(param + i4);
    *reinterpret_cast<float4*>(g) = *reinterpret_cast<const float4*>(grad + i4);
    *reinterpret_cast<float4*>(mm) = *reinterpret_cast<const float4*>(m + i4);
    *reinterpret_cast<float4*>(vv) = *reinterpret_cast<const float4*>(v + i4);
    if (target) *reinterpret_cast<float4*>(tt) = *reinterpret_cast<const float4*>(target + i4);
  } else {
    for (int k = 0; k < cnt; ++k) { p[k] = param[i4 + k]; g[k] = grad[i4 + k]; mm[k] = m[i4 + k]; vv[k] = v[i4 + k]; if (target) tt[k] = target[i4 + k]; }
  }
  if (threadIdx.x == 0) {
    // bias corrections: precomputed by the reduction kernel's last block (scal_dev), else computed
    // here from the device counter (completed updates + 1) or the host step argument
    if (scal_dev) sa = *scal_dev;
    else sa = adam_scalars((double)(step_dev ? step_dev[0] + 1 : step_host), h.lr, h.beta1, h.beta2, h.eps, h.weight_decay, h.tau);
  }
  // global L2 norm of the (scaled) gradient from the per-segment partials, fixed order
  float part = 0.f;
  for (int i = threadIdx.x; i < n_part; i += kOptThreads) part += sumsq_part[i];
  const float total = block_sum_256(part, red);          // contains __syncthreads: sa is visible
  const AdamScalars a = sa;
  const float norm = sqrtf(total) * h.grad_scale;
  float coef = 1.f;
  if (h.max_norm >= 0.f) coef = fminf(h.max_norm / (norm + 1e-6f), 1.f);   // clip_grad.py
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (grad_norm_out) grad_norm_out[0] = norm;
    if (counter_inc) counter_inc[0] += 1;          // nobody reads the counter in this kernel when scal_dev is used
  }
  const float gmul = h.grad_scale * coef;
  if (cnt == 0) return;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k >= cnt) break;
    const float gk = g[k] * gmul;
    p[k] = p[k] * a.decay;                                          // param.mul_(1 - lr*wd)
    mm[k] = mm[k] + a.one_minus_b1 * (gk - mm[k]);                  // exp_avg.lerp_(grad, 1-b1)
    vv[k] = vv[k] * a.b2 + a.one_minus_b2 * gk * gk;                // mul_(b2).addcmul_(g,g,1-b2)
    const float denom = sqrtf(vv[k]) / a.bc2_sqrt + a.eps;
    p[k] = p[k] - a.step_size * (mm[k] / denom);                    // addcdiv_(m, denom, -step_size)
    if (target) tt[k] = p[k] * a.tau + tt[k] * a.one_minus_tau;     // soft_update
  }
  if (cnt == 4) {
    *reinterpret_cast<float4*>(param + i4) = *reinterpret_cast<float4*>(p);
    *reinterpret_cast<float4*>(m + i4) = *reinterpret_cast<float4*>(mm);
    *reinterpret_cast<float4*>(v + i4) = *reinterpret_cast<float4*>(vv);
    if (target) *reinterpret_cast<float4*>(target + i4) = *reinterpret_cast<float4*>(tt);
    if (p_tf32) *reinterpret_cast<float4*>(p_tf32 + i4) = make_float4(rn_tf32(p[0]), rn_tf32(p[1]), rn_tf32(p[2]), rn_tf32(p[3]));
    if (target && t_tf32) *reinterpret_cast<float4*>(t_tf32 + i4) = make_float4(rn_tf32(tt[0]), rn_tf32(tt[1]), rn_tf32(tt[2]), rn_tf32(tt[3]));
    if (p_h) store_split4(p_h, n, i4, p);                 // fp16 hi / lo operand copies (n % 4 == 0 when they are kept)
    if (target && t_h) store_split4(t_h, n, i4, tt);
  } else {
    for (int k = 0; k < cnt; ++k) {
      param[i4 + k] = p[k]; m[i4 + k] = mm[k]; v[i4 + k] = vv[k];
      if (target) target[i4 + k] = tt[k];
      if (p_tf32) p_tf32[i4 + k] = rn_tf32(p[k]);
      if (target && t_tf32) t_tf32[i4 + k] = rn_tf32(tt[k]);
    }
  }
}

// ---- data parallel: gradient all-reduce fused into the optimiser kernel -------------------------
// Every rank's gradient arena, a receive buffer for the reduced gradient and a small control block
// live in symmetric memory (peer-mapped over NVLink / NVSwitch).  One launch per update replaces
// ncclAllReduce + the norm kernel + AdamW:
//   A  wait until every rank's gradient is complete (flags in the peers' control blocks);
//   B  two-shot all-reduce: this rank sums slice `rank` of all peers' gradients in rank order
//      (deterministic) and stores the result, plus the per-block sums of squares of the slice, into
//      EVERY rank's receive buffer / control block;
//   C  wait until every rank has delivered its slice;
//   D  global norm from the world x grid partials (same values, same order on every rank), clip,
//      AdamW, Polyak, TF32 copies - on the full arena, redundantly on every rank, so parameters stay
//      bit-identical without a broadcast.
// All blocks of the grid must be co-resident (grid <= one block per SM); spins are bounded.
constexpr int kDpMaxWorld = 8;
constexpr int kDpFlagWords = 32;           // [0,16): phase-A flags by source rank, [16,32): phase-C flags
struct DpArgs {
  const float* grad_peers[kDpMaxWorld];
  float* red_peers[kDpMaxWorld];
  unsigned* ctl_peers[kDpMaxWorld];      // kDpFlagWords flags, then float sumsq[world][grid]
  int rank, world;
  unsigned long long* local;             // [0] epoch of the last completed exchange, [1] blocks done (monotonic)
  const float* grad_mc; float* red_mc;   // NVLS multicast addresses of the gradient arenas / receive buffers (or NULL)
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) { unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ float4 ld_sys_f4(const float* p) {
  float4 v; asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_sys_f4(float* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float ld_sys_f(const float* p) { float v; asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_sys_f(float* p, float v) { asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// bounded spin: a protocol bug or a dead peer must surface as a launch failure, not as a hung GPU.  The bound is
// wall time (globaltimer, independent of the SM clock) and settable (pqlb_dp_spin_limit): a peer that is merely
// late - checkpointing, logging, a slow env step while the other hosts run ahead - must not take every rank down.
__device__ long long g_dp_spin_ns = 120LL * 1000000000LL;
__device__ __forceinline__ void wait_flag(const unsigned* p, unsigned epoch) {
  if ((int)(ld_acquire_sys(p) - epoch) >= 0) return;
  const unsigned long long t0 = globaltimer_ns();
  const unsigned long long limit = (unsigned long long)g_dp_spin_ns;
  while ((int)(ld_acquire_sys(p) - epoch) < 0) {
    if (globaltimer_ns() - t0 > limit) __trap();
  }
}

// Phase B of the exchange: sum elements [lo, hi) (float4 units) of all ranks' gradients in rank order
// and store the result into every rank's receive buffer; returns this thread's sum of squares.  ILP
// elements per thread and trip, all peers' loads of all of them issued before the first add: an
// NVLink round trip costs ~2-3 us, serialised loads would cost world x trips of them.
// The same phase through the NVSwitch's in-fabric reduction (NVLS): ONE multimem.ld_reduce returns the sum of all
// ranks' copies of four gradient words, ONE multimem.st delivers the result to every rank's receive buffer -
// one NVLink round trip and 1/world of the transactions of the per-peer version.  The switch's summation order
// is not specified, but each word is reduced once, by its owner, and broadcast: all ranks still apply the same values.
__device__ __forceinline__ float4 multimem_ld_reduce_f4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_f4(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
template <int ILP>
__device__ __forceinline__ float dp_reduce_slice_nvls(const DpArgs& dp, int64_t lo, int64_t hi, unsigned G) {
  float sq = 0.f;
  const int64_t gstride = (int64_t)G * kOptThreads;
  for (int64_t i0 = lo + (int64_t)blockIdx.x * kOptThreads + threadIdx.x; i0 < hi; i0 += gstride * ILP) {
    float4 t[ILP];
#pragma unroll
    for (int u = 0; u < ILP; ++u)
      if (i0 + u * gstride < hi) t[u] = multimem_ld_reduce_f4(dp.grad_mc + 4 * (i0 + u * gstride));
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
      const int64_t i = i0 + u * gstride;
      if (i >= hi) break;
      sq += t[u].x * t[u].x + t[u].y * t[u].y + t[u].z * t[u].z + t[u].w * t[u].w;
      multimem_st_f4(dp.red_mc + 4 * i, t[u]);
    }
  }
  return sq;
}

template <int ILP, int WMAX>
__device__ __forceinline__ float dp_reduce_slice(const DpArgs& dp, int64_t lo, int64_t hi, unsigned G) {
  float sq = 0.f;
  const int64_t gstride = (int64_t)G * kOptThreads;
  for (int64_t i0 = lo + (int64_t)blockIdx.x * kOptThreads + threadIdx.x; i0 < hi; i0 += gstride * ILP) {
    float4 t[ILP][WMAX];
#pragma unroll
    for (int u = 0; u < ILP; ++u)
#pragma unroll
      for (int r = 0; r < WMAX; ++r)
        if (r < dp.world && i0 + u * gstride < hi) t[u][r] = ld_sys_f4(dp.grad_peers[r] + 4 * (i0 + u * gstride));
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
      const int64_t i = i0 + u * gstride;
      if (i >= hi) break;
      float4 acc = t[u][0];                          // summed in rank order: deterministic, same on every rank
#pragma unroll
      for (int r = 1; r < WMAX; ++r)
        if (r < dp.world) { acc.x += t[u][r].x; acc.y += t[u][r].y; acc.z += t[u][r].z; acc.w += t[u][r].w; }
      sq += acc.x * acc.x + acc.y * acc.y + acc.z * acc.z + acc.w * acc.w;
#pragma unroll
      for (int q = 0; q < WMAX; ++q)
        if (q < dp.world) st_sys_f4(dp.red_peers[q] + 4 * i, acc);
    }
  }
  return sq;
}

__global__ void __launch_bounds__(kOptThreads, 1)
adamw_polyak_dp_kernel(float* __restrict__ param, float* __restrict__ m, float* __restrict__ v,
                       float* __restrict__ target, float* __restrict__ p_tf32, float* __restrict__ t_tf32,
                       __half* __restrict__ p_h, __half* __restrict__ t_h, int64_t n, DpArgs dp, float max_norm, const AdamScalars* __restrict__ scal_dev,
                       long long* __restrict__ counter_inc, float* __restrict__ grad_norm_out, int exchange_only) {
  __shared__ float red[8];
  __shared__ AdamScalars sa;
  const unsigned G = gridDim.x;
  const unsigned epoch = (unsigned)(dp.local[0] + 1ull);
  // phase clocks of block 0 (ns, accumulated in local[2..5]: wait-for-gradients, reduce+deliver, wait-for-slices,
  // optimiser): how the exchange's time splits into skew between the ranks and transfer (bench.py --dp-timing)
  const bool timed = blockIdx.x == 0 && threadIdx.x == 0;
  unsigned long long tk0 = 0, tk1 = 0, tk2 = 0, tk3 = 0;
  if (timed) tk0 = globaltimer_ns();
  unsigned* my_ctl = dp.ctl_peers[dp.rank];
  if (threadIdx.x == 0 && !exchange_only) sa = *scal_dev;

  // ---- A: every rank's gradient is complete (the kernel boundary before this launch made ours visible)
  if (blockIdx.x == 0 && (int)threadIdx.x < dp.world) {
    __threadfence_system();
    st_release_sys(dp.ctl_peers[threadIdx.x] + dp.rank, epoch);
  }
  if ((int)threadIdx.x < dp.world) wait_flag(my_ctl + threadIdx.x, epoch);
  __syncthreads();
  if (timed) tk1 = globaltimer_ns();

  // ---- B: reduce my slice in rank order, deliver it (and its sums of squares) to every rank
  const int64_t n4 = n >> 2;
  const int64_t s4 = (n4 + dp.world - 1) / dp.world;
  const int64_t lo = (int64_t)dp.rank * s4, hi = lo + s4 < n4 ? lo + s4 : n4;
  // (at most 4 ranks: 4 elements per thread and trip, else 2 - the same 16 / 8 float4 in flight)
  const float sq = dp.grad_mc ? dp_reduce_slice_nvls<4>(dp, lo, hi, G)
                   : dp.world <= 4 ? dp_reduce_slice<4, 4>(dp, lo, hi, G) : dp_reduce_slice<2, kDpMaxWorld>(dp, lo, hi, G);
  const float tot = block_sum_256(sq, red);
  if (threadIdx.x == 0)
    for (int q = 0; q < dp.world; ++q)
      st_sys_f(reinterpret_cast<float*>(dp.ctl_peers[q] + kDpFlagWords) + (int64_t)dp.rank * G + blockIdx.x, tot);
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long done = atomicAdd(dp.local + 1, 1ull) + 1ull;
    if (done == (unsigned long long)epoch * G) {       // last block of this rank: the slice is out
      __threadfence_system();
      for (int q = 0; q < dp.world; ++q) st_release_sys(dp.ctl_peers[q] + 16 + dp.rank, epoch);
    }
  }

  // ---- C: every rank's slice has arrived
  if (timed) tk2 = globaltimer_ns();
  if ((int)threadIdx.x < dp.world) wait_flag(my_ctl + 16 + threadIdx.x, epoch);
  __syncthreads();
  if (timed) tk3 = globaltimer_ns();
  if (exchange_only) {
    // narrow launch (pqlb_grad_exchange_dp): the reduced gradient and the world x G sums of squares are
    // in this rank's receive buffer / control block; the full-width optimiser launch follows
    if (timed) {
      dp.local[0] = epoch;
      dp.local[2] += tk1 - tk0; dp.local[3] += tk2 - tk1; dp.local[4] += tk3 - tk2;
    }
    return;
  }

  // ---- D: global norm (world x G partials, fixed order), clip, AdamW, Polyak
  const float* sumsq_all = reinterpret_cast<const float*>(my_ctl + kDpFlagWords);
  float part = 0.f;
  for (int i = threadIdx.x; i < dp.world * (int)G; i += kOptThreads) part += ld_sys_f(sumsq_all + i);
  const float total = block_sum_256(part, red);
  const AdamScalars a = sa;
  const float grad_scale = 1.f / (float)dp.world;
  const float norm = sqrtf(total) * grad_scale;
  float coef = 1.f;
  if (max_norm >= 0.f) coef = fminf(max_norm / (norm + 1e-6f), 1.f);   // clip_grad.py
  const float gmul = grad_scale * coef;
  const float* redl = dp.red_peers[dp.rank];
  for (int64_t i = (int64_t)blockIdx.x * kOptThreads + threadIdx.x; i < n4; i += (int64_t)G * kOptThreads) {
    const int64_t i4 = 4 * i;
    float p[4], g[4], mm[4], vv[4], tt[4];
    *reinterpret_cast<float4*>(g) = ld_sys_f4(redl + i4);
    *reinterpret_cast<float4*>(p) = *reinterpret_cast<const float4*>(param + i4);
    *reinterpret_cast<float4*>(mm) = *reinterpret_cast<const float4*>(m + i4);
    *reinterpret_cast<float4*>(vv) = *reinterpret_cast<const float4*>(v + i4);
    if (target) *reinterpret_cast<float4*>(tt) = *reinterpret_cast<const float4*>(target + i4);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = g[k] * gmul;
      p[k] = p[k] * a.decay;
      mm[k] = mm[k] + a.one_minus_b1 * (gk - mm[k]);
      vv[k] = vv[k] * a.b2 + a.one_minus_b2 * gk * gk;
      const float denom = sqrtf(vv[k]) / a.bc2_sqrt + a.eps;
      p[k] = p[k] - a.step_size * (mm[k] / denom);
      if (target) tt[k] = p[k] * a.tau + tt[k] * a.one_minus_tau;
    }
    *reinterpret_cast<float4*>(param + i4) = *reinterpret_cast<float4*>(p);
    *reinterpret_cast<float4*>(m + i4) = *reinterpret_cast<float4*>(mm);
    *reinterpret_cast<float4*>(v + i4) = *reinterpret_cast<float4*>(vv);
    if (target) *reinterpret_cast<float4*>(target + i4) = *reinterpret_cast<float4*>(tt);
    if (p_tf32) *reinterpret_cast<float4*>(p_tf32 + i4) = make_float4(rn_tf32(p[0]), rn_tf32(p[1]), rn_tf32(p[2]), rn_tf32(p[3]));
    if (target && t_tf32) *reinterpret_cast<float4*>(t_tf32 + i4) = make_float4(rn_tf32(tt[0]), rn_tf32(tt[1]), rn_tf32(tt[2]), rn_tf32(tt[3]));
    if (p_h) store_split4(p_h, n, i4, p);
    if (target && t_h) store_split4(t_h, n, i4, tt);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (grad_norm_out) grad_norm_out[0] = norm;
    if (counter_inc) counter_inc[0] += 1;
    dp.local[0] = epoch;          // every block of this rank passed phase C, i.e. has read the old value
    const unsigned long long tk4 = globaltimer_ns();
    dp.local[2] += tk1 - tk0; dp.local[3] += tk2 - tk1; dp.local[4] += tk3 - tk2; dp.local[5] += tk4 - tk3;
  }
}

__global__ void __launch_bounds__(kOptThreads)
round_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = rn_tf32(src[i]);
}

// part[blk*n_cols + c] = sum over the 128 rows of block blk of dz[row, c].  A block covers
// 128 rows x 32 columns: warp w sums rows [16w, 16w+16) in ascending order, then the eight warp
// sums are added in warp order (fixed association => deterministic).
__global__ void __launch_bounds__(kOptThreads)
colsum_partial_kernel(const float* __restrict__ dz, int64_t ld, int64_t rows, int n_cols,
                      float* __restrict__ part) {
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + lane;
  const int64_t r0 = (int64_t)blockIdx.x * 128 + w * 16;
  float acc = 0.f;
  if (c < n_cols) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int64_t row = r0 + r;
      acc += row < rows ? dz[row * ld + c] : 0.f;
    }
  }
  red[w][lane] = acc;
  __syncthreads();
  if (w == 0 && c < n_cols) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += red[i][lane];
    part[(int64_t)blockIdx.x * n_cols + c] = tot;
  }
}

struct ColsumMulti { const float* dz[PQLB_MAX_COLSUM]; float* part[PQLB_MAX_COLSUM]; long long ld[PQLB_MAX_COLSUM]; int n_cols[PQLB_MAX_COLSUM]; };

// Same reduction as colsum_partial_kernel for up to PQLB_MAX_COLSUM matrices in one launch
// (blockIdx.z selects the matrix): all bias gradients of one update.
__global__ void __launch_bounds__(kOptThreads)
colsum_multi_kernel(ColsumMulti d, int64_t rows) {
  __shared__ float red[8][32];
  const int e = blockIdx.z;
  const int n_cols = d.n_cols[e];
  if ((int)blockIdx.y * 32 >= n_cols) return;
  const float* __restrict__ dz = d.dz[e];
  const long long ld = d.ld[e];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + lane;
  const int64_t r0 = (int64_t)blockIdx.x * 128 + w * 16;
  float acc = 0.f;
  if (c < n_cols) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int64_t row = r0 + r;
      acc += row < rows ? dz[row * ld + c] : 0.f;
    }
  }
  red[w][lane] = acc;
  __syncthreads();
  if (w == 0 && c < n_cols) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += red[i][lane];
    d.part[e][(int64_t)blockIdx.x * n_cols + c] = tot;
  }
}

// out[0] = scale * sum(part); optionally also ring[count % ring_len] = out[0] and ++count, where
// count = number of completed updates (the loss window of Tracker(5) and the AdamW step counter
// live on the device so that a whole update can be replayed from a CUDA graph).
__global__ void __launch_bounds__(kOptThreads)
sum_partials_kernel(const float* __restrict__ part, int n, float scale, float* __restrict__ out,
                    int64_t* __restrict__ counter, float* __restrict__ ring, int ring_len) {
  __shared__ float red[8];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += kOptThreads) acc += part[i];
  const float tot = block_sum_256(acc, red);
  if (threadIdx.x == 0) {
    const float r = tot * scale;
    out[0] = r;
    if (counter) {
      const int64_t c = counter[0];
      if (ring && ring_len > 0) ring[c % ring_len] = r;
      counter[0] = c + 1;
    }
  }
}

}  // namespace pqlb

using namespace pqlb;

extern "C" int pqlb_grad_reduce(const int64_t* seg_table, int n_seg, const float* ws, float* grad,
                                float* sumsq_part, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(seg_table && n_seg > 0 && ws && grad && sumsq_part);
  FinishArgs fin = {};
  grad_reduce_kernel<<<n_seg, kOptThreads, 0, (cudaStream_t)stream>>>(seg_table, ws, grad, sumsq_part, true, fin);
  PQLB_LAUNCH_RET();
}

extern "C" int pqlb_grad_reduce_finish(const int64_t* seg_table, int n_seg, const float* ws, float* grad,
                                       float* sumsq_part, const float* loss_part, int n_loss, float loss_scale,
                                       float* loss_out, const int64_t* counter, float* ring, int ring_len,
                                       float lr, float beta1, float beta2, float eps, float weight_decay,
                                       float tau, float* scalars_out, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(seg_table && n_seg > 0 && ws && grad && sumsq_part);
  PQLB_CHECK_ARG(loss_part && n_loss > 0 && loss_out && counter && scalars_out && ring_len >= 0);
  FinishArgs fin;
  fin.loss_part = loss_part; fin.n_loss = n_loss; fin.loss_scale = loss_scale; fin.loss_out = loss_out;
  fin.counter = reinterpret_cast<const long long*>(counter); fin.ring = ring; fin.ring_len = ring_len;
  fin.lr = lr; fin.beta1 = beta1; fin.beta2 = beta2; fin.eps = eps; fin.weight_decay = weight_decay; fin.tau = tau;
  fin.scalars_out = reinterpret_cast<AdamScalars*>(scalars_out);
  grad_reduce_kernel<<<n_seg, kOptThreads, 0, (cudaStream_t)stream>>>(seg_table, ws, grad, sumsq_part, true, fin);
  PQLB_LAUNCH_RET();
}

extern "C" int pqlb_grad_sumsq(const int64_t* seg_table, int n_seg, const float* grad,
                               float* sumsq_part, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(seg_table && n_seg > 0 && grad && sumsq_part);
  FinishArgs fin = {};
  grad_reduce_kernel<<<n_seg, kOptThreads, 0, (cudaStream_t)stream>>>(seg_table, nullptr, const_cast<float*>(grad), sumsq_part, false, fin);
  PQLB_LAUNCH_RET();
}

extern "C" int pqlb_adamw_polyak(float* param, const float* grad, float* m, float* v, float* target,
                                 float* param_tf32, float* target_tf32, void* param_h, void* target_h, int64_t n,
                                 const float* sumsq_part, int n_part, float grad_scale,
                                 float max_norm, float lr, float beta1, float beta2, float eps,
                                 float weight_decay, int64_t step, const int64_t* step_dev, float tau,
                                 float* grad_norm_out, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(param && grad && m && v && n > 0 && sumsq_part && n_part > 0 && (step_dev || step >= 1));
  PQLB_CHECK_ALIGN(aligned16(param) && aligned16(grad) && aligned16(m) && aligned16(v));
  PQLB_CHECK_ALIGN((!target || aligned16(target)) && (!param_tf32 || aligned16(param_tf32)) &&
                   (!target_tf32 || aligned16(target_tf32)));
  PQLB_CHECK_ALIGN((!param_h && !target_h) || ((n % 4) == 0 && (reinterpret_cast<uintptr_t>(param_h) & 7) == 0 && (reinterpret_cast<uintptr_t>(target_h) & 7) == 0));
  AdamHyper h;
  h.grad_scale = grad_scale; h.max_norm = max_norm; h.lr = lr; h.beta1 = beta1; h.beta2 = beta2;
  h.eps = eps; h.weight_decay = weight_decay; h.tau = tau;
  const int64_t blocks = (n + kOptThreads * 4 - 1) / (kOptThreads * 4);
  adamw_polyak_kernel<<<(unsigned)blocks, kOptThreads, 0, (cudaStream_t)stream>>>(
      param, grad, m, v, target, param_tf32, target_tf32, reinterpret_cast<__half*>(param_h), reinterpret_cast<__half*>(target_h),
      n, sumsq_part, n_part, h, step, step_dev,
      grad_norm_out, nullptr, nullptr);
  PQLB_LAUNCH_RET();
}

extern "C" int pqlb_adamw_polyak_pre(float* param, const float* grad, float* m, float* v, float* target,
                                     float* param_tf32, float* target_tf32, void* param_h, void* target_h, int64_t n,
                                     const float* sumsq_part, int n_part, float grad_scale, float max_norm,
                                     const float* scalars, int64_t* counter, float* grad_norm_out,
                                     pqlb_stream_t stream) {
  PQLB_CHECK_ARG(param && grad && m && v && n > 0 && sumsq_part && n_part > 0 && scalars && counter);
  PQLB_CHECK_ALIGN(aligned16(param) && aligned16(grad) && aligned16(m) && aligned16(v));
  PQLB_CHECK_ALIGN((!target || aligned16(target)) && (!param_tf32 || aligned16(param_tf32)) &&
                   (!target_tf32 || aligned16(target_tf32)));
  PQLB_CHECK_ALIGN((!param_h && !target_h) || ((n % 4) == 0 && (reinterpret_cast<uintptr_t>(param_h) & 7) == 0 && (reinterpret_cast<uintptr_t>(target_h) & 7) == 0));
  AdamHyper h = {};
  h.grad_scale = grad_scale; h.max_norm = max_norm;
  const int64_t blocks = (n + kOptThreads * 4 - 1) / (kOptThreads * 4);
  adamw_polyak_kernel<<<(unsigned)blocks, kOptThreads, 0, (cudaStream_t)stream>>>(
      param, grad, m, v, target, param_tf32, target_tf32, reinterpret_cast<__half*>(param_h), reinterpret_cast<__half*>(target_h),
      n, sumsq_part, n_part, h, 0, nullptr,
      grad_norm_out, reinterpret_cast<const AdamScalars*>(scalars), reinterpret_cast<long long*>(counter));
  PQLB_LAUNCH_RET();
}

extern "C" int pqlb_adamw_polyak_dp(float* param, float* m, float* v, float* target, float* param_tf32,
                                    float* target_tf32, void* param_h, void* target_h, int64_t n,
                                    const pqlb_dp_desc* dp, float max_norm,
                                    const float* scalars, int64_t* counter, float* grad_norm_out,
                                    pqlb_stream_t stream) {
  PQLB_CHECK_ARG(param && m && v && n > 0 && (n % 4) == 0 && dp && scalars && counter);
  PQLB_CHECK_ARG(dp->world >= 2 && dp->world <= kDpMaxWorld && dp->rank >= 0 && dp->rank < dp->world && dp->local);
  PQLB_CHECK_ARG(dp->grid >= 1 && dp->grid <= kNumSMs);
  PQLB_CHECK_ALIGN(aligned16(param) && aligned16(m) && aligned16(v) && (!target || aligned16(target)) &&
                   (!param_tf32 || aligned16(param_tf32)) && (!target_tf32 || aligned16(target_tf32)));
  PQLB_CHECK_ALIGN((reinterpret_cast<uintptr_t>(param_h) & 7) == 0 && (reinterpret_cast<uintptr_t>(target_h) & 7) == 0);
  DpArgs a;
  for (int r = 0; r < kDpMaxWorld; ++r) {
    const bool on = r < dp->world;
    if (on) PQLB_CHECK_ARG(dp->grad_peers[r] && dp->red_peers[r] && dp->ctl_peers[r] && aligned16(dp->grad_peers[r]) && aligned16(dp->red_peers[r]));
    a.grad_peers[r] = on ? dp->grad_peers[r] : nullptr; a.red_peers[r] = on ? dp->red_peers[r] : nullptr;
    a.ctl_peers[r] = on ? reinterpret_cast<unsigned*>(dp->ctl_peers[r]) : nullptr;
  }
  a.rank = dp->rank; a.world = dp->world; a.local = reinterpret_cast<unsigned long long*>(dp->local);
  a.grad_mc = dp->grad_mc && dp->red_mc ? dp->grad_mc : nullptr; a.red_mc = a.grad_mc ? dp->red_mc : nullptr;
  adamw_polyak_dp_kernel<<<(unsigned)dp->grid, kOptThreads, 0, (cudaStream_t)stream>>>(
      param, m, v, target, param_tf32, target_tf32, reinterpret_cast<__half*>(param_h), reinterpret_cast<__half*>(target_h),
      n, a, max_norm, reinterpret_cast<const AdamScalars*>(scalars),
      reinterpret_cast<long long*>(counter), grad_norm_out, 0);
  PQLB_LAUNCH_RET();
}

extern "C" int pqlb_dp_spin_limit(double seconds) {
  PQLB_CHECK_ARG(seconds > 0.0 && seconds < 1e6);
  const long long ns = (long long)(seconds * 1e9);
  cudaError_t e = cudaMemcpyToSymbol(g_dp_spin_ns, &ns, sizeof(ns));
  return e == cudaSuccess ? PQLB_OK : (int)e;
}

extern "C" int pqlb_grad_exchange_dp(int64_t n, const pqlb_dp_desc* dp, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(n > 0 && (n % 4) == 0 && dp);
  PQLB_CHECK_ARG(dp->world >= 2 && dp->world <= kDpMaxWorld && dp->rank >= 0 && dp->rank < dp->world && dp->local);
  PQLB_CHECK_ARG(dp->grid >= 1 && dp->grid <= kNumSMs);
  DpArgs a;
  for (int r = 0; r < kDpMaxWorld; ++r) {
    const bool on = r < dp->world;
    if (on) PQLB_CHECK_ARG(dp->grad_peers[r] && dp->red_peers[r] && dp->ctl_peers[r] && aligned16(dp->grad_peers[r]) && aligned16(dp->red_peers[r]));
    a.grad_peers[r] = on ? dp->grad_peers[r] : nullptr; a.red_peers[r] = on ? dp->red_peers[r] : nullptr;
    a.ctl_peers[r] = on ? reinterpret_cast<unsigned*>(dp->ctl_peers[r]) : nullptr;
  }
  a.rank = dp->rank; a.world = dp->world; a.local = reinterpret_cast<unsigned long long*>(dp->local);
  a.grad_mc = dp->grad_mc && dp->red_mc ? dp->grad_mc : nullptr; a.red_mc = a.grad_mc ? dp->red_mc : nullptr;
  adamw_polyak_dp_kernel<<<(unsigned)dp->grid, kOptThreads, 0, (cudaStream_t)stream>>>(
      nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, n, a, -1.f, nullptr, nullptr, nullptr, 1);
  PQLB_LAUNCH_RET();
}

extern "C" int pqlb_round_tf32(const float* src, float* dst, int64_t n, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(src && dst && n > 0);
  round_tf32_kernel<<<grid_for(n, kOptThreads, 4), kOptThreads, 0, (cudaStream_t)stream>>>(src, dst, n);
  PQLB_LAUNCH_RET();
}

extern "C" int pqlb_colsum_partial(const float* dz, int64_t ld, int64_t rows, int n_cols, float* part,
                                   pqlb_stream_t stream) {
  PQLB_CHECK_ARG(dz && rows > 0 && n_cols > 0 && part && ld >= n_cols);
  dim3 grid((unsigned)((rows + 127) / 128), (unsigned)((n_cols + 31) / 32));
  colsum_partial_kernel<<<grid, kOptThreads, 0, (cudaStream_t)stream>>>(dz, ld, rows, n_cols, part);
  PQLB_LAUNCH_RET();
}

extern "C" int pqlb_sum_partials(const float* part, int n, float scale, float* out, int64_t* counter,
                                 float* ring, int ring_len, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(part && n > 0 && out && ring_len >= 0 && (!ring || counter));
  sum_partials_kernel<<<1, kOptThreads, 0, (cudaStream_t)stream>>>(part, n, scale, out, counter, ring, ring_len);
  PQLB_LAUNCH_RET();
}

extern "C" int pqlb_colsum_partial_multi(const pqlb_colsum_desc* d, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(d && d->n >= 1 && d->n <= PQLB_MAX_COLSUM && d->rows > 0);
  ColsumMulti m; int max_cols = 0;
  for (int i = 0; i < PQLB_MAX_COLSUM; ++i) {
    const bool on = i < d->n;
    if (on) PQLB_CHECK_ARG(d->dz[i] && d->part[i] && d->n_cols[i] > 0 && d->ld[i] >= d->n_cols[i]);
    m.dz[i] = on ? d->dz[i] : nullptr; m.part[i] = on ? d->part[i] : nullptr;
    m.ld[i] = on ? d->ld[i] : 0; m.n_cols[i] = on ? d->n_cols[i] : 0;
    if (on && d->n_cols[i] > max_cols) max_cols = d->n_cols[i];
  }
  dim3 grid((unsigned)((d->rows + 127) / 128), (unsigned)((max_cols + 31) / 32), (unsigned)d->n);
  colsum_multi_kernel<<<grid, kOptThreads, 0, (cudaStream_t)stream>>>(m, d->rows);
  PQLB_LAUNCH_RET();
}
