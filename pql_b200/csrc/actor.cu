// Actor-side env-step path: what pql/algo/pql_actor.py:87-127 does around env.step()
// with ~60 small torch launches, as three kernels next to the fused policy forward (mlp_fwd.cu) and
// the n-step push (replay.cu):
//   pqlb_rms_update     RunningMeanStd.update           pql/utils/torch_util.py:77-103
//   pqlb_actor_inputs   obs_rms.normalize + operand packing + the exploration-noise draw
//                       (torch_util.py:83-85, noise.py:19-41: torch.normal(zeros, std) = N(0,1) * std)
//   pqlb_env_post       update_tracker, handle_timeout, reward scaling
//                       (pql_actor.py:129-136, pql/utils/common.py:195-202, pql_actor.py:117)
// All HBM/latency-bound elementwise or reduction work; every reduction has a fixed association.
#include "common.cuh"
#include "rng.cuh"

namespace pqlb {

constexpr int kActThreads = 256;
constexpr int kRmsRows = 128;        // rows per block of the column-statistics pass

// ---- RunningMeanStd.update ------------------------------------------------------------------
// Pass 1 (every block): per-column sum and sum of squares of a 128-row slab, accumulated in fp64
// (warp w takes rows w, w+8, ...; the eight warp sums are added in warp order).  Pass 2 (the block
// that draws the last ticket): slab partials added in slab order, batch mean / unbiased variance
// (x.mean(0), x.var(0)), then update_from_moments in fp32 with the reference's operation order
// (python scalars enter as fp32, like torch's tensor-scalar arithmetic on the CPU).
__global__ void __launch_bounds__(kActThreads)
rms_update_kernel(const float* __restrict__ x, long long rows, int cols, long long ldx,
                  float* __restrict__ mean, float* __restrict__ var, double* __restrict__ count,
                  double* __restrict__ part, unsigned* __restrict__ ticket, double* __restrict__ sums_out) {
  __shared__ double s_sum[kActThreads / 32][32], s_sq[kActThreads / 32][32];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long r0 = (long long)blockIdx.x * kRmsRows;
  const long long r1 = r0 + kRmsRows < rows ? r0 + kRmsRows : rows;
  for (int c0 = 0; c0 < cols; c0 += 32) {
    const int c = c0 + lane;
    double s = 0.0, q = 0.0;
    if (c < cols)
      for (long long r = r0 + w; r < r1; r += kActThreads / 32) {
        const double v = (double)x[r * ldx + c];
        s += v; q += v * v;
      }
    s_sum[w][lane] = s; s_sq[w][lane] = q;
    __syncthreads();
    if (w == 0 && c < cols) {
      double ts = 0.0, tq = 0.0;
#pragma unroll
      for (int i = 0; i < kActThreads / 32; ++i) { ts += s_sum[i][lane]; tq += s_sq[i][lane]; }
      part[((long long)blockIdx.x * 2 + 0) * cols + c] = ts;
      part[((long long)blockIdx.x * 2 + 1) * cols + c] = tq;
    }
    __syncthreads();
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (sums_out) {
    // data parallel: this rank's column sums / sums of squares only (slab order); they are all-reduced
    // over the ranks and applied by rms_apply_kernel, so that every rank takes the update of the
    // concatenated batch (pqlb_rms_moments / pqlb_rms_apply)
    for (int c = threadIdx.x; c < cols; c += kActThreads) {
      double s = 0.0, q = 0.0;
      for (unsigned b = 0; b < gridDim.x; ++b) {
        s += __ldcg(part + ((long long)b * 2 + 0) * cols + c);
        q += __ldcg(part + ((long long)b * 2 + 1) * cols + c);
      }
      sums_out[c] = s; sums_out[cols + c] = q;
    }
    __syncthreads();
    if (threadIdx.x == 0) *ticket = 0u;
    return;
  }
  const double n = (double)rows, cnt = count[0], tot = cnt + n;
  const float f_n = (float)n, f_cnt = (float)cnt, f_tot = (float)tot;
  for (int c = threadIdx.x; c < cols; c += kActThreads) {
    double s = 0.0, q = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) {
      s += __ldcg(part + ((long long)b * 2 + 0) * cols + c);
      q += __ldcg(part + ((long long)b * 2 + 1) * cols + c);
    }
    const float b_mean = (float)(s / n);
    const float b_var = (float)((q - s * s / n) / (n - 1.0));           // correction = 1 (torch.var default)
    const float m = mean[c], v = var[c];
    const float delta = __fsub_rn(b_mean, m);
    const float new_mean = __fadd_rn(m, __fdiv_rn(__fmul_rn(delta, f_n), f_tot));
    const float m_a = __fmul_rn(v, f_cnt), m_b = __fmul_rn(b_var, f_n);
    const float cross = __fdiv_rn(__fmul_rn(__fmul_rn(__fmul_rn(delta, delta), f_cnt), f_n), f_tot);
    const float m_2 = __fadd_rn(__fadd_rn(m_a, m_b), cross);
    mean[c] = new_mean;
    var[c] = __fdiv_rn(m_2, f_tot);
  }
  __syncthreads();
  if (threadIdx.x == 0) { count[0] = tot; *ticket = 0u; }
}

// update_from_moments (torch_util.py:91-103) from column sums / sums of squares of n rows in total.
__global__ void __launch_bounds__(kActThreads)
rms_apply_kernel(const double* __restrict__ sums, double n, int cols, float* __restrict__ mean, float* __restrict__ var,
                 double* __restrict__ count) {
  const double cnt = count[0], tot = cnt + n;
  const float f_n = (float)n, f_cnt = (float)cnt, f_tot = (float)tot;
  for (int c = threadIdx.x; c < cols; c += kActThreads) {
    const double s = sums[c], q = sums[cols + c];
    const float b_mean = (float)(s / n);
    const float b_var = (float)((q - s * s / n) / (n - 1.0));
    const float m = mean[c], v = var[c];
    const float delta = __fsub_rn(b_mean, m);
    const float new_mean = __fadd_rn(m, __fdiv_rn(__fmul_rn(delta, f_n), f_tot));
    const float m_a = __fmul_rn(v, f_cnt), m_b = __fmul_rn(b_var, f_n);
    const float cross = __fdiv_rn(__fmul_rn(__fmul_rn(__fmul_rn(delta, delta), f_cnt), f_n), f_tot);
    mean[c] = new_mean;
    var[c] = __fdiv_rn(__fadd_rn(__fadd_rn(m_a, m_b), cross), f_tot);
  }
  __syncthreads();
  if (threadIdx.x == 0) count[0] = tot;
}

// out[r, k] = (x[r, k] - mean[k]) / sqrt(var[k] + eps)  (RunningMeanStd.normalize: no clamp), or
// clamp(., -5, 5) with clamp5 (pql/utils/common.py:139-145); columns [cols, ld_out) zeroed; values
// TF32-rounded when they feed the tensor cores (round_tf32).
__device__ __forceinline__ float rms_norm(float x, float m, float v, float eps) {
  return __fdiv_rn(__fsub_rn(x, m), __fsqrt_rn(__fadd_rn(v, eps)));
}

__global__ void __launch_bounds__(kActThreads)
actor_inputs_kernel(const float* __restrict__ obs, long long rows, int O, long long ld_obs,
                    const float* __restrict__ mean, const float* __restrict__ var, float eps, int clamp5,
                    int round_tf32, float* __restrict__ x, int x_ld,
                    float* __restrict__ noise, int A, const float* __restrict__ row_std, float std,
                    unsigned long long seed, unsigned long long offset, int rng_threads) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (x)
    for (long long i = tid; i < rows * x_ld; i += stride) {
      const long long r = i / x_ld;
      const int k = (int)(i - r * x_ld);
      float v = 0.f;
      if (k < O) {
        v = obs[r * ld_obs + k];
        if (mean) v = rms_norm(v, mean[k], var[k], eps);
        if (clamp5) v = fminf(fmaxf(v, -5.f), 5.f);
        if (round_tf32) v = rn_tf32(v);
      }
      x[i] = v;
    }
  if (noise)      // torch.normal(zeros, std): out.normal_(0, 1).mul_(std).add_(0)
    for (long long i = tid; i < rows * A; i += stride) {
      const float z = torch_normal_f32(seed, offset, rng_threads, i);
      noise[i] = __fmul_rn(z, row_std ? row_std[i / A] : std);
    }
}

// ---- after env.step(): trackers, timeout handling, reward scaling ----------------------------
// One block (episode ends are rare and the pushes must keep env order, like
// Tracker.update(current_returns[env_done_indices])): thread t owns a contiguous run of envs.
// window[(pushed + k) % len] receives the k-th finished episode of this step (only the last `len`
// can survive), pushed += number of finished episodes.
constexpr int kPostThreads = 1024;
__global__ void __launch_bounds__(kPostThreads)
env_post_kernel(const float* __restrict__ reward, const float* __restrict__ done,
                const unsigned char* __restrict__ truncated, float reward_scale, int E,
                float* __restrict__ returns, float* __restrict__ lengths,
                float* __restrict__ ret_window, float* __restrict__ len_window, int win_len,
                long long* __restrict__ pushed, float* __restrict__ reward_out, float* __restrict__ done_out) {
  __shared__ int s_warp[kPostThreads / 32];
  __shared__ int s_total;
  const int per = (E + kPostThreads - 1) / kPostThreads;
  const int e0 = threadIdx.x * per, e1 = min(E, e0 + per);
  int mine = 0;
  for (int e = e0; e < e1; ++e) mine += done[e] != 0.f;
  // exclusive scan of the per-thread counts (warp shuffle scan + serial scan of the 32 warp totals)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) s_warp[w] = incl;
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int i = 0; i < kPostThreads / 32; ++i) { const int t = s_warp[i]; s_warp[i] = acc; acc += t; }
    s_total = acc;
  }
  __syncthreads();
  int k = s_warp[w] + incl - mine;
  const int total = s_total;
  const long long base = pushed[0];
  for (int e = e0; e < e1; ++e) {
    const float r = reward[e], d = done[e];
    const float ret = __fadd_rn(returns[e], r), len = __fadd_rn(lengths[e], 1.f);
    if (d != 0.f) {
      if (k >= total - win_len) {
        const long long slot = (base + k) % win_len;
        ret_window[slot] = ret; len_window[slot] = len;
      }
      ++k;
      returns[e] = 0.f; lengths[e] = 0.f;
    } else {
      returns[e] = ret; lengths[e] = len;
    }
    if (reward_out) reward_out[e] = __fmul_rn(reward_scale, r);
    if (done_out) done_out[e] = (truncated && truncated[e]) ? 0.f : d;      // dones * (~timeout)
  }
  __syncthreads();
  if (threadIdx.x == 0) pushed[0] = base + total;
}

}  // namespace pqlb

using namespace pqlb;

extern "C" int64_t pqlb_rms_workspace_bytes(int64_t rows, int cols) {
  if (rows <= 0 || cols <= 0) return 0;
  const int64_t blocks = (rows + kRmsRows - 1) / kRmsRows;
  return blocks * 2 * cols * (int64_t)sizeof(double) + 16;
}

extern "C" int pqlb_rms_update(const float* x, int64_t rows, int cols, int64_t ldx, float* mean, float* var,
                               double* count, void* workspace, int64_t workspace_bytes, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(x && rows > 0 && cols > 0 && ldx >= cols && mean && var && count && workspace);
  PQLB_CHECK_SHAPE(workspace_bytes >= pqlb_rms_workspace_bytes(rows, cols));
  PQLB_CHECK_ALIGN((reinterpret_cast<uintptr_t>(workspace) & 15) == 0 && (reinterpret_cast<uintptr_t>(count) & 7) == 0);
  const int64_t blocks = (rows + kRmsRows - 1) / kRmsRows;
  PQLB_CHECK_SHAPE(blocks <= 0x7fffffff);
  // layout: [ticket (16 bytes, zero before the first call; the kernel re-arms it)] [partials]
  unsigned* ticket = reinterpret_cast<unsigned*>(workspace);
  double* part = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + 16);
  rms_update_kernel<<<(unsigned)blocks, kActThreads, 0, (cudaStream_t)stream>>>(x, rows, cols, ldx, mean, var, count, part, ticket, nullptr);
  PQLB_LAUNCH_RET();
}

extern "C" int pqlb_rms_moments(const float* x, int64_t rows, int cols, int64_t ldx, double* sums, void* workspace,
                                int64_t workspace_bytes, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(x && rows > 0 && cols > 0 && ldx >= cols && sums && workspace);
  PQLB_CHECK_SHAPE(workspace_bytes >= pqlb_rms_workspace_bytes(rows, cols));
  PQLB_CHECK_ALIGN((reinterpret_cast<uintptr_t>(workspace) & 15) == 0 && (reinterpret_cast<uintptr_t>(sums) & 7) == 0);
  const int64_t blocks = (rows + kRmsRows - 1) / kRmsRows;
  PQLB_CHECK_SHAPE(blocks <= 0x7fffffff);
  unsigned* ticket = reinterpret_cast<unsigned*>(workspace);
  double* part = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + 16);
  rms_update_kernel<<<(unsigned)blocks, kActThreads, 0, (cudaStream_t)stream>>>(x, rows, cols, ldx, nullptr, nullptr, nullptr, part, ticket, sums);
  PQLB_LAUNCH_RET();
}

extern "C" int pqlb_rms_apply(const double* sums, int64_t total_rows, int cols, float* mean, float* var, double* count,
                              pqlb_stream_t stream) {
  PQLB_CHECK_ARG(sums && total_rows > 1 && cols > 0 && mean && var && count);
  rms_apply_kernel<<<1, kActThreads, 0, (cudaStream_t)stream>>>(sums, (double)total_rows, cols, mean, var, count);
  PQLB_LAUNCH_RET();
}

extern "C" int pqlb_actor_inputs(const float* obs, int64_t rows, int obs_dim, int64_t ld_obs, const float* mean,
                                 const float* var, float eps, int clamp5, int round_tf32, float* x, int x_ld,
                                 float* noise, int act_dim, const float* row_std, float std, int64_t seed,
                                 int64_t offset, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(rows > 0 && (x || noise) && ((mean == nullptr) == (var == nullptr)));
  if (x) { PQLB_CHECK_ARG(obs && obs_dim > 0 && ld_obs >= obs_dim); PQLB_CHECK_SHAPE(x_ld >= obs_dim); }
  int T = 1;
  if (noise) {
    PQLB_CHECK_ARG(act_dim > 0 && offset >= 0);
    T = aten_rng_threads(rows * act_dim);
    if (T == 0) return PQLB_E_UNSUPPORTED;
  }
  const int64_t items = x ? rows * x_ld : rows * act_dim;
  actor_inputs_kernel<<<grid_for(items, kActThreads, 4), kActThreads, 0, (cudaStream_t)stream>>>(
      obs, rows, obs_dim, ld_obs, mean, var, eps, clamp5, round_tf32, x, x_ld, noise, act_dim, row_std, std,
      (unsigned long long)seed, (unsigned long long)offset, T);
  PQLB_LAUNCH_RET();
}

extern "C" int pqlb_env_post(const float* reward, const float* done, const uint8_t* truncated, float reward_scale,
                             int num_envs, float* returns, float* lengths, float* ret_window, float* len_window,
                             int window_len, int64_t* pushed, float* reward_out, float* done_out,
                             pqlb_stream_t stream) {
  PQLB_CHECK_ARG(reward && done && num_envs > 0 && returns && lengths && ret_window && len_window && window_len > 0 && pushed);
  env_post_kernel<<<1, kPostThreads, 0, (cudaStream_t)stream>>>(reward, done, truncated, reward_scale, num_envs, returns,
                                                               lengths, ret_window, len_window, window_len,
                                                               reinterpret_cast<long long*>(pushed), reward_out, done_out);
  PQLB_LAUNCH_RET();
}
