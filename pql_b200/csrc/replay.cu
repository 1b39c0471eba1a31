// K1 / K1' / K2: replay ring insert, n-step window push, uniform-index gather (HBM-bound
// byte movers).  Reference behaviour: pql/replay/simple_replay.py:40-104,
// pql/replay/nstep_replay.py:29-92, pql/algo/pql_p_learner.py:66-83.
//
// Work decomposition: every kernel flattens its copy into "items" of VEC consecutive fp32
// words of one field of one row, so that consecutive lanes touch consecutive addresses on both
// the read and the write side, and each thread keeps UNROLL independent loads in flight.
#include "common.cuh"
#include "rng.cuh"

namespace pqlb {

constexpr int kThreads = 256;
constexpr int kUnroll = 8;          // independent 16-byte loads in flight per thread (tools/copy_bench.cu: U8 + 32 blocks/SM is the fastest copy)
constexpr int kBlocksPerSM = 32;

template <int VEC> struct VecT;
template <> struct VecT<4> { using type = float4; };
template <> struct VecT<1> { using type = float; };

struct FieldMap {          // per-row item decomposition (in VEC units)
  int ov, av, per_row;     // obs items, action items, items per row = 2*ov + av
};
template <int VEC> __host__ __device__ inline FieldMap field_map(const RecGeom& g) {
  FieldMap f; f.ov = g.O / VEC; f.av = g.A / VEC; f.per_row = 2 * f.ov + f.av; return f;
}

// Walks the flattened (row, sub-item) space with the grid stride WITHOUT a division per item: the
// first item and the stride are decomposed once per thread, every step is two adds and a carry.
// (The first version divided a 64-bit item index by per_row for every 16 bytes moved - ~100
// instructions - which made these "HBM-bound" kernels issue-bound: 0.58-0.74 of the copy roofline.)
struct ItemWalk {
  int64_t row, drow;
  int sub, dsub, per_row;
  __device__ ItemWalk(int64_t first, int64_t stride, int per_row_) : per_row(per_row_) {
    row = first / per_row; sub = (int)(first - row * per_row);
    drow = stride / per_row; dsub = (int)(stride - drow * per_row);
  }
  __device__ __forceinline__ void next() {
    sub += dsub; row += drow;
    if (sub >= per_row) { sub -= per_row; ++row; }
  }
};

// -------------------------------------------------------------------------------- K1 insert
// One warp moves kRec records at a time: lane l handles items l, l + 32, ... of each record, so
// the field decode depends on the item only (shared by the kRec records), every record pointer is
// warp-uniform and a thread keeps kRec independent 16-byte loads in flight with ~50 registers
// (the flat item-per-thread version needed 114 at 8 loads in flight and ran at 2 blocks per SM).
constexpr int kRec = 4;

// item -> (field, offset inside the field, offset inside the record), in VEC-word units.
// field: 0 obs, 1 next_obs, 2 action, 3 reward/done (+ padding up to a whole 32-byte sector)
template <int VEC> struct ItemDecode {
  int field, fo, ro;
  __device__ __forceinline__ ItemDecode(int it, const RecGeom& g, const FieldMap& f) {
    if (it < f.ov)               { field = 0; fo = it * VEC;                ro = g.off_obs + fo; }
    else if (it < 2 * f.ov)      { field = 1; fo = (it - f.ov) * VEC;       ro = g.off_next + fo; }
    else if (it < f.per_row)     { field = 2; fo = (it - 2 * f.ov) * VEC;   ro = g.off_act + fo; }
    else                         { field = 3; fo = (it - f.per_row) * VEC;  ro = g.off_rew + fo; }
  }
};

template <int VEC>
__global__ void __launch_bounds__(kThreads)
ring_insert_kernel(float* __restrict__ ring, RecGeom g,
                   const float* __restrict__ obs, const float* __restrict__ act,
                   const float* __restrict__ rew, const float* __restrict__ nobs,
                   const float* __restrict__ done, int64_t n, int64_t next_p, int64_t head,
                   int64_t tail) {
  using V = typename VecT<VEC>::type;
  const FieldMap f = field_map<VEC>(g);
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  // items of the tail: [rew, done] and zero padding up to the end of their 32-byte sector, so that
  // every sector of a record is written whole (no read-modify-write of a partial sector in L2)
  const int tail_words = ((g.off_rew + 2 + 7) & ~7) - g.off_rew;
  const int items = f.per_row + tail_words / VEC;

  auto slot_of = [&](int64_t row) -> int64_t {
    if (row < head) { int64_t s = next_p + row; return s < tail ? -1 : s; }   // superseded by the wrapped tail
    return row - head;
  };

  for (int64_t r0 = ((((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5)) * kRec; r0 < n; r0 += nwarps * kRec) {
    float* dst[kRec];
#pragma unroll
    for (int r = 0; r < kRec; ++r) {
      const int64_t s = r0 + r < n ? slot_of(r0 + r) : -1;
      dst[r] = s >= 0 ? ring + s * g.rec_ld : nullptr;
    }
    for (int it = lane; it < items; it += 32) {
      const ItemDecode<VEC> d(it, g, f);
      V val[kRec];
#pragma unroll
      for (int r = 0; r < kRec; ++r) {
        if (!dst[r]) continue;
        const int64_t row = r0 + r;
        if (d.field == 0)      val[r] = __ldcs(reinterpret_cast<const V*>(obs + row * g.O + d.fo));
        else if (d.field == 1) val[r] = __ldcs(reinterpret_cast<const V*>(nobs + row * g.O + d.fo));
        else if (d.field == 2) val[r] = __ldcs(reinterpret_cast<const V*>(act + row * g.A + d.fo));
        else {
          // done is stored as the reference's bool column (x != 0)
          float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int q = 0; q < VEC; ++q) {
            const int wq = d.fo + q;
            t[q] = wq == 0 ? rew[row] : (wq == 1 ? (done[row] != 0.f ? 1.f : 0.f) : 0.f);
          }
          val[r] = *reinterpret_cast<V*>(t);
        }
      }
#pragma unroll
      for (int r = 0; r < kRec; ++r)       // streaming stores: the ring must not evict the learner's working set from L2
        if (dst[r]) __stcs(reinterpret_cast<V*>(dst[r] + d.ro), val[r]);
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(kThreads)
obsring_insert_kernel(float* __restrict__ ring, int O, const float* __restrict__ obs, int64_t n,
                      int64_t next_p, int64_t head, int64_t tail) {
  using V = typename VecT<VEC>::type;
  const int ov = O / VEC;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  ItemWalk w((int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride, ov);
  while (w.row < n) {
    V val[kUnroll]; float* dst[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      dst[u] = nullptr;
      if (w.row < n) {
        const int64_t row = w.row;
        const int sub = w.sub;
        int64_t s;
        if (row < head) { s = next_p + row; if (s < tail) s = -1; } else s = row - head;
        val[u] = __ldcs(reinterpret_cast<const V*>(obs + row * O + sub * VEC));
        if (s >= 0) dst[u] = ring + s * O + sub * VEC;
      }
      w.next();
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (dst[u]) __stcs(reinterpret_cast<V*>(dst[u]), val[u]);
  }
}

// -------------------------------------------------------------------------------- K1' n-step
struct Gammas { float g[32]; };

// One warp per (t, e).  Window entries are addressed by global step index g (slot g % nstep);
// entries with g >= count0 are read from this call's input block instead, so every emitted
// row depends only on the window as it was BEFORE the call.  mode bit 0 = emit rows, bit 1 =
// store the last nstep steps of the block into the window.  T == 1 does both in one launch (a
// warp reads slots (c-n+1..c-1) % n and then overwrites slot c % n of its own env); for T > 1
// the host issues the emit launch and then the store launch, so no emitted row can observe a
// slot that another warp has already overwritten.
__global__ void __launch_bounds__(kThreads)
nstep_push_kernel(float* __restrict__ window, int E, int n, RecGeom g,
                  const float* __restrict__ obs, const float* __restrict__ act,
                  const float* __restrict__ rew, const float* __restrict__ nobs,
                  const float* __restrict__ done, int T, int64_t count0, Gammas gam, int t0, int mode,
                  float* __restrict__ o_obs, float* __restrict__ o_act, float* __restrict__ o_rew,
                  float* __restrict__ o_nobs, float* __restrict__ o_done) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t total = (int64_t)T * E;
  for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total; w += nwarps) {
    const int t = (int)(w / E);
    const int e = (int)(w - (int64_t)t * E);
    const int64_t c = count0 + t;
    if ((mode & 1) && t >= t0) {
      float r = 0.f, d = 0.f;
      if (lane < n) {
        const int64_t gj = c - n + 1 + lane;
        if (gj >= count0) { const int64_t i = (int64_t)e * T + (gj - count0); r = rew[i]; d = done[i]; }
        else { const float* rec = window + ((int64_t)e * n + (gj % n)) * g.rec_ld; r = rec[g.off_rew]; d = rec[g.off_done]; }
      }
      // k* = first maximum of done over the window (torch argmax), any = some entry non-zero
      float dmax = __shfl_sync(0xffffffffu, d, 0); int kstar = 0; bool any = dmax != 0.f; float dlast = dmax;
      for (int j = 1; j < n; ++j) {
        const float dj = __shfl_sync(0xffffffffu, d, j);
        if (dj > dmax) { dmax = dj; kstar = j; }
        any |= (dj != 0.f); dlast = dj;
      }
      // sum_k (r_k * gamma_k) * mask_k, left to right, every product/sum rounded (no FMA)
      float acc = 0.f;
      for (int j = 0; j < n; ++j) {
        const float rj = __shfl_sync(0xffffffffu, r, j);
        const float m = (!any || j <= kstar) ? 1.f : 0.f;
        const float term = __fmul_rn(__fmul_rn(rj, gam.g[j]), m);
        acc = (j == 0) ? term : __fadd_rn(acc, term);
      }
      const int64_t row = (int64_t)(t - t0) * E + e;
      auto src_of = [&](int j, const float* in_field, int dim, int rec_off) -> const float* {
        const int64_t gj = c - n + 1 + j;
        if (gj >= count0) return in_field + ((int64_t)e * T + (gj - count0)) * dim;
        return window + ((int64_t)e * n + (gj % n)) * g.rec_ld + rec_off;
      };
      const float* s_obs = src_of(0, obs, g.O, g.off_obs);
      const float* s_act = src_of(0, act, g.A, g.off_act);
      const float* s_nxt = src_of(any ? kstar : n - 1, nobs, g.O, g.off_next);
      for (int k = lane; k < g.O; k += 32) { o_obs[row * g.O + k] = s_obs[k]; o_nobs[row * g.O + k] = s_nxt[k]; }
      for (int k = lane; k < g.A; k += 32) o_act[row * g.A + k] = s_act[k];
      if (lane == 0) { o_rew[row] = acc; o_done[row] = any ? 1.f : dlast; }
    }
    if ((mode & 2) && t >= T - n) {   // the last n steps of the block become the stored window
      float* rec = window + ((int64_t)e * n + (c % n)) * g.rec_ld;
      const int64_t i = (int64_t)e * T + t;
      for (int k = lane; k < g.O; k += 32) { rec[g.off_obs + k] = obs[i * g.O + k]; rec[g.off_next + k] = nobs[i * g.O + k]; }
      for (int k = lane; k < g.A; k += 32) rec[g.off_act + k] = act[i * g.A + k];
      if (lane == 0) { rec[g.off_rew] = rew[i]; rec[g.off_done] = done[i]; }
    }
  }
}

// -------------------------------------------------------------------------------- K2 gather
template <int VEC>
__global__ void __launch_bounds__(kThreads)
sample_gather_kernel(const float* __restrict__ ring, RecGeom g, const int64_t* __restrict__ idx,
                     int64_t B, float* __restrict__ o_obs, float* __restrict__ o_act,
                     float* __restrict__ o_rew, float* __restrict__ o_nobs, float* __restrict__ o_done) {
  using V = typename VecT<VEC>::type;
  const FieldMap f = field_map<VEC>(g);
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int items = f.per_row + (VEC == 4 ? 1 : 2);      // + [rew, done]
  for (int64_t b0 = ((((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5)) * kRec; b0 < B; b0 += nwarps * kRec) {
    const float* rec[kRec];
#pragma unroll
    for (int r = 0; r < kRec; ++r) rec[r] = b0 + r < B ? ring + __ldg(idx + b0 + r) * g.rec_ld : nullptr;
    for (int it = lane; it < items; it += 32) {
      const ItemDecode<VEC> d(it, g, f);
      V val[kRec];
#pragma unroll
      for (int r = 0; r < kRec; ++r)
        if (rec[r]) val[r] = __ldcs(reinterpret_cast<const V*>(rec[r] + d.ro));
#pragma unroll
      for (int r = 0; r < kRec; ++r) {
        if (!rec[r]) continue;
        const int64_t b = b0 + r;
        if (d.field == 0)      *reinterpret_cast<V*>(o_obs + b * g.O + d.fo) = val[r];
        else if (d.field == 1) *reinterpret_cast<V*>(o_nobs + b * g.O + d.fo) = val[r];
        else if (d.field == 2) *reinterpret_cast<V*>(o_act + b * g.A + d.fo) = val[r];
        else {
          const float* t = reinterpret_cast<const float*>(&val[r]);   // bool column -> .float() is 0.0 / 1.0 already
          if (VEC == 4) { o_rew[b] = t[0]; o_done[b] = t[1]; }
          else if (d.fo == 0) o_rew[b] = t[0];
          else o_done[b] = t[0];
        }
      }
    }
  }
}

constexpr int kNormCols = 512;      // widest observation whose (mean, sd) table is kept in shared memory

__device__ __forceinline__ float norm_clamp(float x, float m, float v, float eps) {
  // common.py:139-145: clamp((x - mean) / sqrt(var + eps), -5, 5)
  float y = __fdiv_rn(__fsub_rn(x, m), __fsqrt_rn(__fadd_rn(v, eps)));
  return fminf(fmaxf(y, -5.f), 5.f);
}

// Fused V-learner batch: x_cur = [norm(obs)|action|0], x_tgt = [norm(next_obs)| (actor head) |0];
// TF32-rounded because both are tensor-core operands.  VEC = 4: one item = one float4 of one
// output row (requires O % 4 == 0 and A % 4 == 0, so a float4 never straddles two fields).
template <int VEC, int REC>
__global__ void __launch_bounds__(kThreads)
sample_critic_batch_kernel(const float* __restrict__ ring, RecGeom g, const int64_t* __restrict__ idx,
                           int64_t B, const float* __restrict__ mean, const float* __restrict__ var,
                           float eps, float* __restrict__ x_cur, float* __restrict__ x_tgt, int x_ld,
                           float* __restrict__ o_rew, float* __restrict__ o_done, RngArgs ra,
                           float* __restrict__ xf_cur, float* __restrict__ xf_tgt) {
  using V = typename VecT<VEC>::type;
  // fused sampler RNG (rng.cuh): the indices are torch.randint's for this generator state, drawn here
  unsigned long long seed = 0, off = 0, range = 1;
  if (ra.state) { seed = ra.state[0]; off = ra.state[1] + ra.state[2] * ra.counter[0]; range = ra.range[0]; }
  const FieldMap f = field_map<VEC>(g);
  const int lane = threadIdx.x & 31;
  // sqrt(var + eps) once per block instead of once per element (an IEEE square root is ~15 instructions, and
  // the kernel is issue-bound at learner batch sizes): same values, same rounding
  __shared__ float s_mean[kNormCols], s_sd[kNormCols];
  const bool tab = mean != nullptr && g.O <= kNormCols;
  if (tab) {
    for (int k = threadIdx.x; k < g.O; k += blockDim.x) { s_mean[k] = mean[k]; s_sd[k] = __fsqrt_rn(__fadd_rn(var[k], eps)); }
    __syncthreads();
  }
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int rd_items = VEC == 4 ? 1 : 2;
  const int pad_items = (x_ld - g.O - g.A) / VEC;           // zero padding columns of both input rows
  const int items = f.per_row + rd_items + pad_items;
  for (int64_t b0 = ((((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5)) * REC; b0 < B; b0 += nwarps * REC) {
    const float* rec[REC];
    long long mine = 0;
    if (ra.state && lane < REC && b0 + lane < B) {
      mine = mod_range(torch_rand_u32(seed, off, ra.threads_idx, b0 + lane), range);
      const_cast<int64_t*>(idx)[b0 + lane] = mine;            // kept for inspection / tests
    }
#pragma unroll
    for (int r = 0; r < REC; ++r) {
      const long long ix = ra.state ? __shfl_sync(0xffffffffu, mine, r) : (b0 + r < B ? __ldg(idx + b0 + r) : 0);
      rec[r] = b0 + r < B ? ring + ix * g.rec_ld : nullptr;
    }
    for (int it = lane; it < items; it += 32) {
      if (it >= f.per_row + rd_items) {                       // padding columns
        const int k = g.O + g.A + (it - f.per_row - rd_items) * VEC;
        float z[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < REC; ++r)
          if (rec[r]) {
            *reinterpret_cast<V*>(x_cur + (b0 + r) * x_ld + k) = *reinterpret_cast<V*>(z);
            *reinterpret_cast<V*>(x_tgt + (b0 + r) * x_ld + k) = *reinterpret_cast<V*>(z);
            if (xf_cur) *reinterpret_cast<V*>(xf_cur + (b0 + r) * x_ld + k) = *reinterpret_cast<V*>(z);
            if (xf_tgt) *reinterpret_cast<V*>(xf_tgt + (b0 + r) * x_ld + k) = *reinterpret_cast<V*>(z);
          }
        continue;
      }
      const ItemDecode<VEC> d(it, g, f);
      V val[REC];
#pragma unroll
      for (int r = 0; r < REC; ++r)
        if (rec[r]) val[r] = __ldcs(reinterpret_cast<const V*>(rec[r] + d.ro));
      float m[4] = {0.f, 0.f, 0.f, 0.f}, sd[4] = {1.f, 1.f, 1.f, 1.f};
      if (mean && d.field <= 1) {
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
          if (tab) { m[q] = s_mean[d.fo + q]; sd[q] = s_sd[d.fo + q]; }
          else { m[q] = mean[d.fo + q]; sd[q] = __fsqrt_rn(__fadd_rn(var[d.fo + q], eps)); }
        }
      }
#pragma unroll
      for (int r = 0; r < REC; ++r) {
        if (!rec[r]) continue;
        const int64_t b = b0 + r;
        float* t = reinterpret_cast<float*>(&val[r]);
        if (d.field <= 1) {
          // common.py:139-145: clamp((x - mean) / sqrt(var + eps), -5, 5), then the tensor-core operand rounding
#pragma unroll
          for (int q = 0; q < VEC; ++q) {
            float y = t[q];
            if (mean) y = fminf(fmaxf(__fdiv_rn(__fsub_rn(y, m[q]), sd[q]), -5.f), 5.f);
            t[q] = y;
          }
          // the un-rounded row feeds the split-fp16 forward (pqlb_mlp_forward_h), the TF32-rounded one the
          // weight-gradient contraction of the first layer
          float* xf = d.field == 0 ? xf_cur : xf_tgt;
          if (xf) *reinterpret_cast<V*>(xf + b * x_ld + d.fo) = val[r];
#pragma unroll
          for (int q = 0; q < VEC; ++q) t[q] = rn_tf32(t[q]);
          *reinterpret_cast<V*>((d.field == 0 ? x_cur : x_tgt) + b * x_ld + d.fo) = val[r];
        } else if (d.field == 2) {
          if (xf_cur) *reinterpret_cast<V*>(xf_cur + b * x_ld + g.O + d.fo) = val[r];
#pragma unroll
          for (int q = 0; q < VEC; ++q) t[q] = rn_tf32(t[q]);
          *reinterpret_cast<V*>(x_cur + b * x_ld + g.O + d.fo) = val[r];
        } else {
          if (VEC == 4) { o_rew[b] = t[0]; o_done[b] = t[1]; }
          else if (d.fo == 0) o_rew[b] = t[0];
          else o_done[b] = t[0];
        }
      }
    }
  }
  if (ra.state && ra.noise) {      // the update's N(0,1) draw (target-policy noise), next in the generator's stream
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ra.noise_numel; i += stride)
      ra.noise[i] = torch_normal_f32(seed, off + 4, ra.threads_noise, i);
  }
}

// P-learner batch with the fused index draw: one warp per kRec rows (lanes < kRec draw the indices),
// lanes walk the columns.  Same values as sample_obs_batch_kernel on torch.randint's indices.
template <int REC>
__global__ void __launch_bounds__(kThreads)
sample_obs_batch_rng_kernel(const float* __restrict__ obsring, int O, int64_t* __restrict__ idx,
                            int64_t B, const float* __restrict__ mean, const float* __restrict__ var,
                            float eps, float* __restrict__ x, int x_ld, int A, RngArgs ra, float* __restrict__ xf) {
  const int lane = threadIdx.x & 31;
  __shared__ float s_mean[kNormCols], s_sd[kNormCols];
  const bool tab = mean != nullptr && O <= kNormCols;
  if (tab) {
    for (int k = threadIdx.x; k < O; k += blockDim.x) { s_mean[k] = mean[k]; s_sd[k] = __fsqrt_rn(__fadd_rn(var[k], eps)); }
    __syncthreads();
  }
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const unsigned long long seed = ra.state[0], off = ra.state[1] + ra.state[2] * ra.counter[0], range = ra.range[0];
  for (int64_t b0 = ((((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5)) * REC; b0 < B; b0 += nwarps * REC) {
    long long mine = 0;
    if (lane < REC && b0 + lane < B) {
      mine = mod_range(torch_rand_u32(seed, off, ra.threads_idx, b0 + lane), range);
      idx[b0 + lane] = mine;
    }
    const float* src[REC];
#pragma unroll
    for (int r = 0; r < REC; ++r) {
      const long long ix = __shfl_sync(0xffffffffu, mine, r);
      src[r] = b0 + r < B ? obsring + ix * O : nullptr;
    }
    for (int k = lane; k < x_ld; k += 32) {
      if (k < O) {
        float v[REC];
#pragma unroll
        for (int r = 0; r < REC; ++r) v[r] = src[r] ? __ldcs(src[r] + k) : 0.f;
        float m = 0.f, sd = 1.f;
        if (tab) { m = s_mean[k]; sd = s_sd[k]; }
        else if (mean) { m = mean[k]; sd = __fsqrt_rn(__fadd_rn(var[k], eps)); }
#pragma unroll
        for (int r = 0; r < REC; ++r) {
          if (!src[r]) continue;
          float y = v[r];
          if (mean) y = fminf(fmaxf(__fdiv_rn(__fsub_rn(y, m), sd), -5.f), 5.f);     // common.py:139-145
          x[(b0 + r) * x_ld + k] = rn_tf32(y);
          if (xf) xf[(b0 + r) * x_ld + k] = y;
        }
      } else if (k >= O + A) {
#pragma unroll
        for (int r = 0; r < REC; ++r)
          if (src[r]) { x[(b0 + r) * x_ld + k] = 0.f; if (xf) xf[(b0 + r) * x_ld + k] = 0.f; }
      }
    }
  }
}

__global__ void store_i64_kernel(long long* dst, long long v) { *dst = v; }
__global__ void add_i64_kernel(long long* dst, const long long* src, long long delta) { *dst = *src + delta; }

__global__ void __launch_bounds__(kThreads)
sample_obs_batch_kernel(const float* __restrict__ obsring, int O, const int64_t* __restrict__ idx,
                        int64_t B, const float* __restrict__ mean, const float* __restrict__ var,
                        float eps, float* __restrict__ x, int x_ld, int A, float* __restrict__ xf) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (ItemWalk w((int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride, x_ld); w.row < B; w.next()) {
    const int64_t b = w.row;
    const int k = w.sub;
    if (k < O) {
      float v = obsring[__ldg(idx + b) * O + k];
      if (mean) v = norm_clamp(v, mean[k], var[k], eps);
      x[b * x_ld + k] = rn_tf32(v);
      if (xf) xf[b * x_ld + k] = v;
    } else if (k >= O + A) {
      x[b * x_ld + k] = 0.f;               // action columns [O, O+A) belong to the actor head
      if (xf) xf[b * x_ld + k] = 0.f;
    }
  }
}

// x[row] = rn_tf32([a[row, :na] | b[row, :nb] | 0...]) : torch.cat((state, action)) of
// mlp.py:198 plus the operand rounding/padding the tensor-core layers expect.
__global__ void __launch_bounds__(kThreads)
pack_x_kernel(const float* __restrict__ a, int64_t lda, int na, const float* __restrict__ b,
              int64_t ldb, int nb, float* __restrict__ x, int x_ld, int64_t rows) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (ItemWalk w((int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride, x_ld); w.row < rows; w.next()) {
    const int64_t r = w.row;
    const int k = w.sub;
    float v = 0.f;
    if (k < na) v = a[r * lda + k];
    else if (k < na + nb) v = b[r * ldb + (k - na)];
    x[r * x_ld + k] = rn_tf32(v);
  }
}

// Records per warp: four for batches that fill the machine anyway (fewer index / field decodes per byte), ONE for
// the batches the learners actually issue (8192-16384 rows): with four, a batch of 8192 is 2048 warps = 14 per SM,
// each walking its records' items in seven dependent rounds of memory latency (11.7 us, 0.2 of the HBM roofline);
// with one record per warp there are four times as many warps in flight and two rounds.
constexpr int64_t kSmallBatch = 32768;
template <int VEC, typename... Args>
static inline void launch_critic_batch(int64_t batch, cudaStream_t st, Args... args) {
  if (batch <= kSmallBatch)
    sample_critic_batch_kernel<VEC, 1><<<grid_for(batch * 32, kThreads, 1, kBlocksPerSM), kThreads, 0, st>>>(args...);
  else
    sample_critic_batch_kernel<VEC, kRec><<<grid_for(batch * 32, kThreads, kRec, kBlocksPerSM), kThreads, 0, st>>>(args...);
}

static inline void split_insert(int64_t n, int64_t next_p, int64_t capacity, int64_t* head, int64_t* tail) {
  const int64_t p = next_p + n;
  if (p > capacity) { *head = capacity - next_p; *tail = p - capacity; }
  else { *head = n; *tail = 0; }
}

}  // namespace pqlb

using namespace pqlb;

int pqlb_ring_insert_tma(float* ring, int64_t capacity, int obs_dim, int act_dim, const float* obs,
                         const float* action, const float* reward, const float* next_obs,
                         const float* done, int64_t n, int64_t next_p, cudaStream_t stream);
static bool g_insert_ldg = false;
/* Tests / measurements: 1 = force the LDG/STG insert kernel, 0 = TMA path when possible (default). */
extern "C" void pqlb_ring_insert_force_ldg(int on) { g_insert_ldg = on != 0; }

extern "C" int pqlb_obs_pad(int obs_dim) { return round_up(obs_dim, 4); }
extern "C" int pqlb_record_ld(int obs_dim, int act_dim) { return rec_geom(obs_dim, act_dim).rec_ld; }
extern "C" int pqlb_x_ld(int obs_dim, int act_dim) { return round_up(obs_dim + act_dim, 4); }

extern "C" int pqlb_ring_insert(float* ring, int64_t capacity, int obs_dim, int act_dim,
                                const float* obs, const float* action, const float* reward,
                                const float* next_obs, const float* done, int64_t n,
                                int64_t next_p, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(ring && capacity > 0 && obs_dim > 0 && act_dim > 0 && n >= 0);
  PQLB_CHECK_ARG(next_p >= 0 && next_p <= capacity);
  if (n == 0) return PQLB_OK;
  PQLB_CHECK_ARG(obs && action && reward && next_obs && done);
  int64_t head, tail; split_insert(n, next_p, capacity, &head, &tail);
  PQLB_CHECK_SHAPE(tail <= capacity);      // reference: slice-assign shape mismatch
  PQLB_CHECK_ALIGN(aligned16(ring));
  const RecGeom g = rec_geom(obs_dim, act_dim);
  cudaStream_t st = (cudaStream_t)stream;
  if (!g_insert_ldg) {         // TMA tile mover (replay_tma.cu) whenever shape and alignment allow
    const int rc = pqlb_ring_insert_tma(ring, capacity, obs_dim, act_dim, obs, action, reward, next_obs, done, n, next_p, st);
    if (rc != PQLB_E_UNSUPPORTED) return rc;
  }
  const bool vec4 = obs_dim % 4 == 0 && act_dim % 4 == 0 && aligned16(obs) && aligned16(action) && aligned16(next_obs);
  if (vec4) {
    ring_insert_kernel<4><<<grid_for(n * 32, kThreads, kRec, kBlocksPerSM), kThreads, 0, st>>>(
        ring, g, obs, action, reward, next_obs, done, n, next_p, head, tail);
  } else {
    ring_insert_kernel<1><<<grid_for(n * 32, kThreads, kRec, kBlocksPerSM), kThreads, 0, st>>>(
        ring, g, obs, action, reward, next_obs, done, n, next_p, head, tail);
  }
  PQLB_LAUNCH_RET();
}

extern "C" int pqlb_obsring_insert(float* ring, int64_t capacity, int obs_dim, const float* obs,
                                   int64_t n, int64_t next_p, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(ring && capacity > 0 && obs_dim > 0 && n >= 0 && next_p >= 0 && next_p <= capacity);
  if (n == 0) return PQLB_OK;
  PQLB_CHECK_ARG(obs);
  int64_t head, tail; split_insert(n, next_p, capacity, &head, &tail);
  PQLB_CHECK_SHAPE(tail <= capacity);
  cudaStream_t st = (cudaStream_t)stream;
  if (obs_dim % 4 == 0 && aligned16(obs) && aligned16(ring)) {
    obsring_insert_kernel<4><<<grid_for(n * (obs_dim / 4), kThreads, kUnroll, kBlocksPerSM), kThreads, 0, st>>>(
        ring, obs_dim, obs, n, next_p, head, tail);
  } else {
    obsring_insert_kernel<1><<<grid_for(n * obs_dim, kThreads, kUnroll, kBlocksPerSM), kThreads, 0, st>>>(
        ring, obs_dim, obs, n, next_p, head, tail);
  }
  PQLB_LAUNCH_RET();
}

extern "C" int pqlb_nstep_push(float* window, int num_envs, int nstep, int obs_dim, int act_dim,
                               const float* obs, const float* action, const float* reward,
                               const float* next_obs, const float* done, int T, int64_t count,
                               const float* gammas_host, float* out_obs, float* out_action,
                               float* out_reward, float* out_next_obs, float* out_done,
                               pqlb_stream_t stream) {
  PQLB_CHECK_ARG(window && num_envs > 0 && nstep >= 2 && nstep <= 32 && obs_dim > 0 && act_dim > 0);
  PQLB_CHECK_ARG(obs && action && reward && next_obs && done && gammas_host && T > 0 && count >= 0);
  int64_t t0_64 = (int64_t)nstep - 1 - count; if (t0_64 < 0) t0_64 = 0;
  const int t0 = (int)(t0_64 > T ? T : t0_64);
  if (t0 < T) PQLB_CHECK_ARG(out_obs && out_action && out_reward && out_next_obs && out_done);
  Gammas gam; for (int i = 0; i < 32; ++i) gam.g[i] = i < nstep ? gammas_host[i] : 0.f;
  const RecGeom g = rec_geom(obs_dim, act_dim);
  const int64_t warps = (int64_t)T * num_envs;
  int blocks = (int)((warps * 32 + kThreads - 1) / kThreads);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  if (T == 1) {
    nstep_push_kernel<<<blocks, kThreads, 0, (cudaStream_t)stream>>>(
        window, num_envs, nstep, g, obs, action, reward, next_obs, done, T, count, gam, t0, 3,
        out_obs, out_action, out_reward, out_next_obs, out_done);
  } else {
    if (t0 < T) {
      nstep_push_kernel<<<blocks, kThreads, 0, (cudaStream_t)stream>>>(
          window, num_envs, nstep, g, obs, action, reward, next_obs, done, T, count, gam, t0, 1,
          out_obs, out_action, out_reward, out_next_obs, out_done);
      PQLB_COUNT_LAUNCH(1);
    }
    nstep_push_kernel<<<blocks, kThreads, 0, (cudaStream_t)stream>>>(
        window, num_envs, nstep, g, obs, action, reward, next_obs, done, T, count, gam, t0, 2,
        out_obs, out_action, out_reward, out_next_obs, out_done);
  }
  PQLB_LAUNCH_RET();
}

extern "C" int pqlb_sample_gather(const float* ring, int64_t capacity, int obs_dim, int act_dim,
                                  const int64_t* idx, int64_t batch, float* out_obs,
                                  float* out_action, float* out_reward, float* out_next_obs,
                                  float* out_done, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(ring && capacity > 0 && obs_dim > 0 && act_dim > 0 && batch >= 0);
  if (batch == 0) return PQLB_OK;
  PQLB_CHECK_ARG(idx && out_obs && out_action && out_reward && out_next_obs && out_done);
  const RecGeom g = rec_geom(obs_dim, act_dim);
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec4 = obs_dim % 4 == 0 && act_dim % 4 == 0 && aligned16(ring) && aligned16(out_obs) &&
                    aligned16(out_action) && aligned16(out_next_obs);
  if (vec4) {
    sample_gather_kernel<4><<<grid_for(batch * 32, kThreads, kRec, kBlocksPerSM), kThreads, 0, st>>>(
        ring, g, idx, batch, out_obs, out_action, out_reward, out_next_obs, out_done);
  } else {
    sample_gather_kernel<1><<<grid_for(batch * 32, kThreads, kRec, kBlocksPerSM), kThreads, 0, st>>>(
        ring, g, idx, batch, out_obs, out_action, out_reward, out_next_obs, out_done);
  }
  PQLB_LAUNCH_RET();
}

extern "C" int pqlb_sample_critic_batch(const float* ring, int64_t capacity, int obs_dim, int act_dim,
                                        const int64_t* idx, int64_t batch, const float* mean,
                                        const float* var, float eps, float* x_cur, float* x_tgt,
                                        int x_ld, float* reward, float* done, float* xf_cur, float* xf_tgt,
                                        pqlb_stream_t stream) {
  PQLB_CHECK_ARG(ring && capacity > 0 && obs_dim > 0 && act_dim > 0 && batch > 0 && idx);
  PQLB_CHECK_ARG(x_cur && x_tgt && reward && done && ((mean == nullptr) == (var == nullptr)));
  const RecGeom g = rec_geom(obs_dim, act_dim);
  PQLB_CHECK_SHAPE(x_ld >= obs_dim + act_dim && x_ld % 4 == 0);
  const int per_row = obs_dim + x_ld;
  const bool vec4 = obs_dim % 4 == 0 && act_dim % 4 == 0 && aligned16(ring) && aligned16(x_cur) && aligned16(x_tgt) &&
                    (!mean || (aligned16(mean) && aligned16(var)));
  if (vec4)
    launch_critic_batch<4>(batch, (cudaStream_t)stream, ring, g, idx, batch, mean, var, eps, x_cur, x_tgt, x_ld, reward, done, RngArgs{}, xf_cur, xf_tgt);
  else
    launch_critic_batch<1>(batch, (cudaStream_t)stream, ring, g, idx, batch, mean, var, eps, x_cur, x_tgt, x_ld, reward, done, RngArgs{}, xf_cur, xf_tgt);
  PQLB_LAUNCH_RET();
}

static int make_rng_args(RngArgs* ra, const int64_t* state, const int64_t* counter, const int64_t* range,
                         int64_t batch, int64_t capacity, float* noise, int64_t noise_numel) {
  PQLB_CHECK_ARG(state && counter && range && noise_numel >= 0 && (noise_numel == 0 || noise));
  if (capacity >= (1LL << 28)) return PQLB_E_UNSUPPORTED;        // ATen switches to 64-bit draws there
  ra->state = reinterpret_cast<const long long*>(state);
  ra->counter = reinterpret_cast<const long long*>(counter);
  ra->range = reinterpret_cast<const long long*>(range);
  ra->noise = noise_numel ? noise : nullptr; ra->noise_numel = noise_numel;
  ra->threads_idx = aten_rng_threads(batch);
  ra->threads_noise = noise_numel ? aten_rng_threads(noise_numel) : 1;
  if (ra->threads_idx == 0 || ra->threads_noise == 0) return PQLB_E_UNSUPPORTED;
  return PQLB_OK;
}

extern "C" int pqlb_sample_critic_batch_rng(const float* ring, int64_t capacity, int obs_dim, int act_dim,
                                            int64_t* idx_out, int64_t batch, const float* mean,
                                            const float* var, float eps, float* x_cur, float* x_tgt,
                                            int x_ld, float* reward, float* done, const int64_t* rng_state,
                                            const int64_t* counter, const int64_t* cur_capacity,
                                            float* noise_out, int64_t noise_numel, float* xf_cur, float* xf_tgt,
                                            pqlb_stream_t stream) {
  PQLB_CHECK_ARG(ring && capacity > 0 && obs_dim > 0 && act_dim > 0 && batch > 0 && idx_out);
  PQLB_CHECK_ARG(x_cur && x_tgt && reward && done && ((mean == nullptr) == (var == nullptr)));
  const RecGeom g = rec_geom(obs_dim, act_dim);
  PQLB_CHECK_SHAPE(x_ld >= obs_dim + act_dim && x_ld % 4 == 0);
  RngArgs ra;
  const int rc = make_rng_args(&ra, rng_state, counter, cur_capacity, batch, capacity, noise_out, noise_numel);
  if (rc != PQLB_OK) return rc;
  const bool vec4 = obs_dim % 4 == 0 && act_dim % 4 == 0 && aligned16(ring) && aligned16(x_cur) && aligned16(x_tgt) &&
                    (!mean || (aligned16(mean) && aligned16(var)));
  if (vec4)
    launch_critic_batch<4>(batch, (cudaStream_t)stream, ring, g, idx_out, batch, mean, var, eps, x_cur, x_tgt, x_ld, reward, done, ra, xf_cur, xf_tgt);
  else
    launch_critic_batch<1>(batch, (cudaStream_t)stream, ring, g, idx_out, batch, mean, var, eps, x_cur, x_tgt, x_ld, reward, done, ra, xf_cur, xf_tgt);
  PQLB_LAUNCH_RET();
}

extern "C" int pqlb_sample_obs_batch_rng(const float* obsring, int64_t capacity, int obs_dim, int64_t* idx_out,
                                         int64_t batch, const float* mean, const float* var, float eps,
                                         float* x, int x_ld, int act_dim, const int64_t* rng_state,
                                         const int64_t* counter, const int64_t* cur_capacity,
                                         float* xf, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(obsring && capacity > 0 && obs_dim > 0 && batch > 0 && idx_out && x && act_dim >= 0);
  PQLB_CHECK_ARG((mean == nullptr) == (var == nullptr));
  PQLB_CHECK_SHAPE(x_ld >= obs_dim + act_dim);
  RngArgs ra;
  const int rc = make_rng_args(&ra, rng_state, counter, cur_capacity, batch, capacity, nullptr, 0);
  if (rc != PQLB_OK) return rc;
  if (batch <= kSmallBatch)
    sample_obs_batch_rng_kernel<1><<<grid_for(batch * 32, kThreads, 1, kBlocksPerSM), kThreads, 0, (cudaStream_t)stream>>>(
        obsring, obs_dim, idx_out, batch, mean, var, eps, x, x_ld, act_dim, ra, xf);
  else
    sample_obs_batch_rng_kernel<kRec><<<grid_for(batch * 32, kThreads, kRec, kBlocksPerSM), kThreads, 0, (cudaStream_t)stream>>>(
        obsring, obs_dim, idx_out, batch, mean, var, eps, x, x_ld, act_dim, ra, xf);
  PQLB_LAUNCH_RET();
}

extern "C" int pqlb_store_i64(int64_t* dst, int64_t value, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(dst);
  store_i64_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<long long*>(dst), (long long)value);
  PQLB_LAUNCH_RET();
}

extern "C" int pqlb_add_i64(int64_t* dst, const int64_t* src, int64_t delta, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(dst && src);
  add_i64_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<long long*>(dst), reinterpret_cast<const long long*>(src), (long long)delta);
  PQLB_LAUNCH_RET();
}

extern "C" int pqlb_sample_obs_batch(const float* obsring, int64_t capacity, int obs_dim,
                                     const int64_t* idx, int64_t batch, const float* mean,
                                     const float* var, float eps, float* x, int x_ld,
                                     int act_dim, float* xf, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(obsring && capacity > 0 && obs_dim > 0 && batch > 0 && idx && x && act_dim >= 0);
  PQLB_CHECK_ARG((mean == nullptr) == (var == nullptr));
  PQLB_CHECK_SHAPE(x_ld >= obs_dim + act_dim);
  sample_obs_batch_kernel<<<grid_for(batch * x_ld, kThreads, 2), kThreads, 0, (cudaStream_t)stream>>>(
      obsring, obs_dim, idx, batch, mean, var, eps, x, x_ld, act_dim, xf);
  PQLB_LAUNCH_RET();
}

extern "C" int pqlb_pack_x(const float* a, int64_t lda, int na, const float* b, int64_t ldb, int nb,
                           float* x, int x_ld, int64_t rows, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(a && na > 0 && nb >= 0 && (nb == 0 || b) && x && rows > 0);
  PQLB_CHECK_SHAPE(x_ld >= na + nb && lda >= na && (nb == 0 || ldb >= nb));
  pack_x_kernel<<<grid_for(rows * x_ld, kThreads, 2), kThreads, 0, (cudaStream_t)stream>>>(
      a, lda, na, b, ldb, nb, x, x_ld, rows);
  PQLB_LAUNCH_RET();
}
