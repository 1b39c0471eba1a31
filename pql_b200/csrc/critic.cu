// Loss-side kernels of the critic / actor updates: twin-Q TD target + MSE + scalar-head
// backward, DPG loss through min(Q1,Q2), and the C51 categorical projection + BCE + softmax
// backward.  All reductions use a fixed association (deterministic run to run).
// Reference behaviour: pql/algo/pql_v_learner.py:83-108, pql/algo/pql_p_learner.py:55-57,
// pql/utils/distl_util.py:4-20.
#include "common.cuh"

namespace pqlb {

constexpr int kRowsPerBlock = 64;
constexpr int kLossThreads = 256;
constexpr int kHeadN = 128;          // width of the last hidden layer (mlp.py:33-34)

// One block = 64 batch rows of ONE of the two nets (blockIdx.y).  Phase A: per-row scalars;
// phase B: dz3 rows + per-column partial sums of dq*h3 (head weight gradient) for the block.
//   mode 0 (critic): dq_i = 2 (q_i - y) / B,  loss partial = sum (q_i - y)^2
//   mode 1 (actor):  dq_i = -(1/B) [q_i is the min] (1/2 each on ties), loss partial = sum min(q1,q2)
// Partials: loss_part[2*blockIdx.x + net]; head weight/bias gradient ws_net[blockIdx.x*129 + {0..127, 128}].
__global__ void __launch_bounds__(kLossThreads)
head_loss_kernel(int mode, const float* __restrict__ q1, const float* __restrict__ q2,
                 const float* __restrict__ tq1, const float* __restrict__ tq2,
                 const float* __restrict__ reward, const float* __restrict__ done, float gamma_n,
                 long long B, const float* __restrict__ h3_1, const float* __restrict__ h3_2,
                 const float* __restrict__ w4_1, const float* __restrict__ w4_2,
                 float* __restrict__ dz3_1, float* __restrict__ dz3_2, float* __restrict__ y_out,
                 float* __restrict__ ws1, float* __restrict__ ws2, float* __restrict__ loss_part,
                 float* __restrict__ bz1, float* __restrict__ bz2) {
  __shared__ float s_dq[kRowsPerBlock];
  __shared__ float s_red[8], s_redb[8];
  __shared__ float s_col[8][kHeadN];
  __shared__ float s_colz[8][kHeadN];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int net = blockIdx.y;
  const long long r0 = (long long)blockIdx.x * kRowsPerBlock;
  const float invB = 1.f / (float)B;

  float lpart = 0.f;
  if (tid < kRowsPerBlock) {
    const long long r = r0 + tid;
    float dq = 0.f;
    if (r < B) {
      const float a = q1[r], b = q2[r];
      const float mine = net ? b : a, other = net ? a : b;
      if (mode == 0) {
        // pql_v_learner.py:104-105: y = reward + (1 - done) * gamma^n * min(tq1, tq2)
        const float y = __fadd_rn(reward[r], __fmul_rn(__fmul_rn(1.f - done[r], gamma_n), fminf(tq1[r], tq2[r])));
        if (y_out && net == 0) y_out[r] = y;
        const float e = mine - y;
        lpart = e * e;
        dq = 2.f * e * invB;
      } else {
        lpart = net == 0 ? fminf(a, b) : 0.f;
        dq = mine < other ? -invB : (mine == other ? -0.5f * invB : 0.f);
      }
    }
    s_dq[tid] = dq;
  }
  // deterministic block sum of the loss partials
  float v = lpart;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  if (tid == 0) { float t = 0.f; for (int i = 0; i < 8; ++i) t += s_red[i]; loss_part[2 * blockIdx.x + net] = t; }

  // Phase B: warp w owns rows [8w, 8w+8); lane owns columns 4*lane .. 4*lane+3
  const int c = lane * 4;
  const float* h3 = net ? h3_2 : h3_1;
  const float* w4 = net ? w4_2 : w4_1;
  float* dz3 = net ? dz3_2 : dz3_1;
  const float4 w = *reinterpret_cast<const float4*>(w4 + c);
  float4 h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long r = r0 + warp * 8 + i;
    h[i] = r < B ? *reinterpret_cast<const float4*>(h3 + r * kHeadN + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), accz = make_float4(0.f, 0.f, 0.f, 0.f);
  float accb = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int lr = warp * 8 + i;
    const long long r = r0 + lr;
    if (r >= B) break;
    const float dq = s_dq[lr];
    float4 o;
    o.x = rn_tf32(dq * w.x * (h[i].x > 0.f ? 1.f : h[i].x + 1.f));
    o.y = rn_tf32(dq * w.y * (h[i].y > 0.f ? 1.f : h[i].y + 1.f));
    o.z = rn_tf32(dq * w.z * (h[i].z > 0.f ? 1.f : h[i].z + 1.f));
    o.w = rn_tf32(dq * w.w * (h[i].w > 0.f ? 1.f : h[i].w + 1.f));
    *reinterpret_cast<float4*>(dz3 + r * kHeadN + c) = o;
    acc.x += dq * h[i].x; acc.y += dq * h[i].y; acc.z += dq * h[i].z; acc.w += dq * h[i].w;
    accz.x += o.x; accz.y += o.y; accz.z += o.z; accz.w += o.w;      // column sums of dz3: the layer-3 bias gradient
    accb += dq;
  }
  float* bz = net ? bz2 : bz1;
  if (bz) *reinterpret_cast<float4*>(&s_colz[warp][c]) = accz;
  if (mode == 0) {
    *reinterpret_cast<float4*>(&s_col[warp][c]) = acc;
    if (lane == 0) s_redb[warp] = accb;
    __syncthreads();
    float* ws = net ? ws2 : ws1;
    if (tid < kHeadN) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += s_col[i][tid];
      ws[(long long)blockIdx.x * (kHeadN + 1) + tid] = t;
    } else if (tid == kHeadN) {
      float t = 0.f;
      for (int i = 0; i < 8; ++i) t += s_redb[i];
      ws[(long long)blockIdx.x * (kHeadN + 1) + kHeadN] = t;
    }
    if (bz && tid >= kHeadN && tid < 2 * kHeadN) {       // the other half of the block sums the dz3 columns
      const int cc = tid - kHeadN;
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += s_colz[i][cc];
      bz[(long long)blockIdx.x * kHeadN + cc] = t;
    }
  }
}

// C51: one warp per batch row; lane j owns atoms j and j+32.
__global__ void __launch_bounds__(kLossThreads)
c51_td_loss_kernel(const float* __restrict__ p1, const float* __restrict__ p2,
                   const float* __restrict__ tp1, const float* __restrict__ tp2, int ld,
                   const float* __restrict__ reward, const float* __restrict__ done,
                   const float* __restrict__ z_atoms, float gamma_n, float v_min, float v_max,
                   float delta_z, int N, long long B, float* __restrict__ target_out,
                   float* __restrict__ dl1, float* __restrict__ dl2, int ld_d,
                   float* __restrict__ loss_part) {
  __shared__ float s_wl[8][2][64], s_wu[8][2][64], s_proj[8][2][64];
  __shared__ int s_l[8][64], s_u[8][64];
  __shared__ float s_red[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long r = (long long)blockIdx.x * 8 + warp;
  float bce = 0.f;
  if (r < B) {
    const float rew = reward[r];
    const float scale = __fmul_rn(1.f - done[r], gamma_n);   // (1 - done) * gamma
    for (int j = lane; j < 64; j += 32) {
      s_proj[warp][0][j] = 0.f; s_proj[warp][1][j] = 0.f;
      if (j < N) {
        // distl_util.py:8-14
        const float tz = fminf(fmaxf(__fadd_rn(rew, __fmul_rn(scale, z_atoms[j])), v_min), v_max);   // no FMA: torch rounds the product
        const float b = __fdiv_rn(tz - v_min, delta_z);
        int l = (int)floorf(b), u = (int)ceilf(b);
        if (u > 0 && l == u) l -= 1;
        if (l < N - 1 && l == u) u += 1;
        s_l[warp][j] = l; s_u[warp][j] = u;
        const float a1 = tp1[r * ld + j], a2 = tp2[r * ld + j];
        s_wl[warp][0][j] = a1 * ((float)u - b); s_wu[warp][0][j] = a1 * (b - (float)l);
        s_wl[warp][1][j] = a2 * ((float)u - b); s_wu[warp][1][j] = a2 * (b - (float)l);
      }
    }
    __syncwarp();
    {
      // index_add_ order on the CPU: for every output atom, all lower-atom terms (source ascending), then
      // all upper-atom terms (source ascending).  Segmented form over ALL lanes: lane k owns output atoms k and
      // k + 32 of both networks and walks the sources in that order, adding the ones that land on its atoms -
      // the same additions in the same order as the serial scatter (l and u are monotone in the source atom,
      // so each output atom takes a contiguous run of sources), with register accumulators instead of a
      // dependent chain of shared-memory read-modify-writes on two lanes.
      float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};          // [net][atom half]
      const int k0 = lane, k1 = lane + 32;
      for (int j = 0; j < N; ++j) {
        const int l = s_l[warp][j];
        const float w0 = s_wl[warp][0][j], w1 = s_wl[warp][1][j];
        if (l == k0) { acc[0][0] += w0; acc[1][0] += w1; }
        if (l == k1) { acc[0][1] += w0; acc[1][1] += w1; }
      }
      for (int j = 0; j < N; ++j) {
        const int u = s_u[warp][j];
        const float w0 = s_wu[warp][0][j], w1 = s_wu[warp][1][j];
        if (u == k0) { acc[0][0] += w0; acc[1][0] += w1; }
        if (u == k1) { acc[0][1] += w0; acc[1][1] += w1; }
      }
      s_proj[warp][0][k0] = acc[0][0]; s_proj[warp][1][k0] = acc[1][0];
      s_proj[warp][0][k1] = acc[0][1]; s_proj[warp][1][k1] = acc[1][1];
    }
    __syncwarp();
    const float gscale = 1.f / ((float)B * (float)N);
    float t[2], g1[2], g2[2], q1v[2], q2v[2];
    float dot1 = 0.f, dot2 = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int j = lane + 32 * k;
      t[k] = 0.f; g1[k] = 0.f; g2[k] = 0.f; q1v[k] = 0.f; q2v[k] = 0.f;
      if (j < N) {
        t[k] = fminf(s_proj[warp][0][j], s_proj[warp][1][j]);       // pql_v_learner.py:102
        if (target_out) target_out[r * N + j] = t[k];
        const float a = p1[r * ld + j], b = p2[r * ld + j];
        q1v[k] = a; q2v[k] = b;
        // F.binary_cross_entropy: -(t*max(log p,-100) + (1-t)*max(log(1-p),-100))
        bce += (t[k] - 1.f) * fmaxf(logf(1.f - a), -100.f) - t[k] * fmaxf(logf(a), -100.f);
        bce += (t[k] - 1.f) * fmaxf(logf(1.f - b), -100.f) - t[k] * fmaxf(logf(b), -100.f);
        g1[k] = gscale * (a - t[k]) / fmaxf((1.f - a) * a, 1e-12f);
        g2[k] = gscale * (b - t[k]) / fmaxf((1.f - b) * b, 1e-12f);
        dot1 += a * g1[k]; dot2 += b * g2[k];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      dot1 += __shfl_xor_sync(0xffffffffu, dot1, o);
      dot2 += __shfl_xor_sync(0xffffffffu, dot2, o);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int j = lane + 32 * k;
      if (j < ld_d) {
        // softmax backward: dlogit_j = p_j (g_j - sum_k p_k g_k); padding columns are zero
        dl1[r * ld_d + j] = j < N ? rn_tf32(q1v[k] * (g1[k] - dot1)) : 0.f;
        dl2[r * ld_d + j] = j < N ? rn_tf32(q2v[k] * (g2[k] - dot2)) : 0.f;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) bce += __shfl_xor_sync(0xffffffffu, bce, o);
  if (lane == 0) s_red[warp] = bce;
  __syncthreads();
  if (threadIdx.x == 0) { float tt = 0.f; for (int i = 0; i < 8; ++i) tt += s_red[i]; loss_part[blockIdx.x] = tt; }
}

// C51 actor loss (pql_p_learner.py:55-57 with mlp.py:255-259): q_i = sum_j p_i[j] z[j];
// loss = -mean(min(q1,q2)); gradient through softmax: dlogit_j = dq * p_j * (z_j - q).
// One warp per row, lane j owns atoms j and j+32.
__global__ void __launch_bounds__(kLossThreads)
c51_dpg_loss_kernel(const float* __restrict__ p1, const float* __restrict__ p2, int ld,
                    const float* __restrict__ z_atoms, int N, long long B,
                    float* __restrict__ dl1, float* __restrict__ dl2, int ld_d,
                    float* __restrict__ q_min_out, float* __restrict__ loss_part) {
  __shared__ float s_red[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long r = (long long)blockIdx.x * 8 + warp;
  float qm = 0.f;
  if (r < B) {
    float a[2], b[2], z[2];
    float q1 = 0.f, q2 = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int j = lane + 32 * k;
      a[k] = 0.f; b[k] = 0.f; z[k] = 0.f;
      if (j < N) { a[k] = p1[r * ld + j]; b[k] = p2[r * ld + j]; z[k] = z_atoms[j]; }
      q1 += a[k] * z[k]; q2 += b[k] * z[k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      q1 += __shfl_xor_sync(0xffffffffu, q1, o);
      q2 += __shfl_xor_sync(0xffffffffu, q2, o);
    }
    qm = fminf(q1, q2);
    if (lane == 0 && q_min_out) q_min_out[r] = qm;
    if (dl1) {
      const float invB = 1.f / (float)B;
      const float d1 = q1 < q2 ? -invB : (q1 == q2 ? -0.5f * invB : 0.f);
      const float d2 = q2 < q1 ? -invB : (q1 == q2 ? -0.5f * invB : 0.f);
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int j = lane + 32 * k;
        if (j < ld_d) {
          dl1[r * ld_d + j] = j < N ? rn_tf32(d1 * a[k] * (z[k] - q1)) : 0.f;
          dl2[r * ld_d + j] = j < N ? rn_tf32(d2 * b[k] * (z[k] - q2)) : 0.f;
        }
      }
    }
  }
  if (lane == 0) s_red[warp] = qm;       // qm is warp-uniform
  __syncthreads();
  if (threadIdx.x == 0 && loss_part) { float tt = 0.f; for (int i = 0; i < 8; ++i) tt += s_red[i]; loss_part[blockIdx.x] = tt; }
}

}  // namespace pqlb

using namespace pqlb;

extern "C" int pqlb_doubleq_td_loss(const float* q1, const float* q2, const float* tq1, const float* tq2,
                                    const float* reward, const float* done, float gamma_n, int64_t batch,
                                    const float* h3_1, const float* h3_2, const float* w4_1,
                                    const float* w4_2, float* dz3_1, float* dz3_2, float* y_out,
                                    float* ws_head1, float* ws_head2, float* loss_part,
                                    pqlb_stream_t stream) {
  PQLB_CHECK_ARG(q1 && q2 && tq1 && tq2 && reward && done && batch > 0);
  PQLB_CHECK_ARG(h3_1 && h3_2 && w4_1 && w4_2 && dz3_1 && dz3_2 && ws_head1 && ws_head2 && loss_part);
  PQLB_CHECK_ALIGN(aligned16(h3_1) && aligned16(h3_2) && aligned16(w4_1) && aligned16(w4_2) &&
                   aligned16(dz3_1) && aligned16(dz3_2));
  const dim3 blocks((unsigned)((batch + kRowsPerBlock - 1) / kRowsPerBlock), 2);
  head_loss_kernel<<<blocks, kLossThreads, 0, (cudaStream_t)stream>>>(
      0, q1, q2, tq1, tq2, reward, done, gamma_n, batch, h3_1, h3_2, w4_1, w4_2, dz3_1, dz3_2, y_out,
      ws_head1, ws_head2, loss_part, nullptr, nullptr);
  PQLB_LAUNCH_RET();
}

extern "C" int pqlb_doubleq_td_loss_b3(const float* q1, const float* q2, const float* tq1, const float* tq2,
                                       const float* reward, const float* done, float gamma_n, int64_t batch,
                                       const float* h3_1, const float* h3_2, const float* w4_1, const float* w4_2,
                                       float* dz3_1, float* dz3_2, float* y_out, float* ws_head1, float* ws_head2,
                                       float* loss_part, float* bias3_part1, float* bias3_part2,
                                       pqlb_stream_t stream) {
  PQLB_CHECK_ARG(q1 && q2 && tq1 && tq2 && reward && done && batch > 0 && bias3_part1 && bias3_part2);
  PQLB_CHECK_ARG(h3_1 && h3_2 && w4_1 && w4_2 && dz3_1 && dz3_2 && ws_head1 && ws_head2 && loss_part);
  PQLB_CHECK_ALIGN(aligned16(h3_1) && aligned16(h3_2) && aligned16(w4_1) && aligned16(w4_2) &&
                   aligned16(dz3_1) && aligned16(dz3_2));
  const dim3 blocks((unsigned)((batch + kRowsPerBlock - 1) / kRowsPerBlock), 2);
  head_loss_kernel<<<blocks, kLossThreads, 0, (cudaStream_t)stream>>>(
      0, q1, q2, tq1, tq2, reward, done, gamma_n, batch, h3_1, h3_2, w4_1, w4_2, dz3_1, dz3_2, y_out,
      ws_head1, ws_head2, loss_part, bias3_part1, bias3_part2);
  PQLB_LAUNCH_RET();
}

extern "C" int pqlb_dpg_loss(const float* q1, const float* q2, int64_t batch, const float* h3_1,
                             const float* h3_2, const float* w4_1, const float* w4_2, float* dz3_1,
                             float* dz3_2, float* loss_part, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(q1 && q2 && batch > 0 && h3_1 && h3_2 && w4_1 && w4_2 && dz3_1 && dz3_2 && loss_part);
  PQLB_CHECK_ALIGN(aligned16(h3_1) && aligned16(h3_2) && aligned16(w4_1) && aligned16(w4_2) &&
                   aligned16(dz3_1) && aligned16(dz3_2));
  const dim3 blocks((unsigned)((batch + kRowsPerBlock - 1) / kRowsPerBlock), 2);
  head_loss_kernel<<<blocks, kLossThreads, 0, (cudaStream_t)stream>>>(
      1, q1, q2, nullptr, nullptr, nullptr, nullptr, 0.f, batch, h3_1, h3_2, w4_1, w4_2, dz3_1, dz3_2,
      nullptr, nullptr, nullptr, loss_part, nullptr, nullptr);
  PQLB_LAUNCH_RET();
}

extern "C" int pqlb_c51_td_loss(const float* p1, const float* p2, const float* tp1, const float* tp2,
                                int ld, const float* reward, const float* done, const float* z_atoms,
                                float gamma_n, float v_min, float v_max, int num_atoms, int64_t batch,
                                float* target_out, float* dlogit1, float* dlogit2, int ld_d,
                                float* loss_part, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(p1 && p2 && tp1 && tp2 && reward && done && z_atoms && batch > 0);
  PQLB_CHECK_ARG(dlogit1 && dlogit2 && loss_part && num_atoms >= 2 && num_atoms <= 64);
  PQLB_CHECK_SHAPE(ld >= num_atoms && ld_d >= num_atoms && ld_d <= 64);
  const float delta_z = (float)(((double)v_max - (double)v_min) / (double)(num_atoms - 1));
  const unsigned blocks = (unsigned)((batch + 7) / 8);
  c51_td_loss_kernel<<<blocks, kLossThreads, 0, (cudaStream_t)stream>>>(
      p1, p2, tp1, tp2, ld, reward, done, z_atoms, gamma_n, v_min, v_max, delta_z, num_atoms, batch,
      target_out, dlogit1, dlogit2, ld_d, loss_part);
  PQLB_LAUNCH_RET();
}

extern "C" int pqlb_c51_dpg_loss(const float* p1, const float* p2, int ld, const float* z_atoms,
                                 int num_atoms, int64_t batch, float* dlogit1, float* dlogit2,
                                 int ld_d, float* q_min_out, float* loss_part, pqlb_stream_t stream) {
  PQLB_CHECK_ARG(p1 && p2 && z_atoms && batch > 0 && num_atoms >= 2 && num_atoms <= 64);
  PQLB_CHECK_ARG((dlogit1 == nullptr) == (dlogit2 == nullptr));
  PQLB_CHECK_SHAPE(ld >= num_atoms && (!dlogit1 || (ld_d >= num_atoms && ld_d <= 64)));
  const unsigned blocks = (unsigned)((batch + 7) / 8);
  c51_dpg_loss_kernel<<<blocks, kLossThreads, 0, (cudaStream_t)stream>>>(
      p1, p2, ld, z_atoms, num_atoms, batch, dlogit1, dlogit2, ld_d, q_min_out, loss_part);
  PQLB_LAUNCH_RET();
}
