/*
 * pqlb200.h - C ABI of libpqlb200.so: the B200 (sm_100a) learner hot path of Parallel Q-Learning.
 *
 * Every entry point takes plain device pointers and sizes, enqueues work on the given
 * cudaStream_t (passed as void*), never allocates, never synchronises and never keeps a
 * pointer after returning.  Return value: 0 = ok, <0 = argument error
 * (PQLB_E_*), >0 = cudaError_t of the failed launch.  Each function names the reference
 * code it replaces (paths relative to the reference checkout).
 *
 * Layouts
 * -------
 * Transition record (ring and n-step window), fp32 words, `rec_ld` words per record
 * (the next power of two >= the words used: 256 words = 1 KB for obs 88 / act 16, 512 words for
 * obs 211 / act 20 - random gathers of power-of-two-aligned records run ~1.4x faster than
 * 800-byte ones on B200, see common.cuh):
 *     [0, O)                     obs
 *     [obs_pad, obs_pad+O)       next_obs          obs_pad = round_up(O, 4)
 *     [2*obs_pad, 2*obs_pad+A)   action
 *     [2*obs_pad + act_pad]      reward            act_pad = round_up(A, 4)
 *     [2*obs_pad + act_pad + 1]  done   (ring: 0.0f / 1.0f == the reference's bool column;
 *                                        n-step window: the raw float the actor supplied)
 * Critic input row ("x"): torch.cat((state, action)) order - [0,O) normalised obs, [O,O+A)
 * action - with x_ld = round_up(O+A, 4) words per row; padding words are zero.
 * Parameter arena: one flat fp32 buffer per network set; every nn.Linear weight is stored
 * [out, round_up(in,4)] row-major (padding columns zero), then its bias; each tensor starts
 * at a 32-word aligned offset.
 */
#ifndef PQLB200_H
#define PQLB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PQLB_OK 0
#define PQLB_E_ARG (-1)      /* null pointer / non-positive size / bad enum           */
#define PQLB_E_SHAPE (-2)    /* sizes inconsistent with each other (torch: shape error) */
#define PQLB_E_ALIGN (-3)    /* pointer or leading dimension not aligned as required   */
#define PQLB_E_DRIVER (-4)   /* cuTensorMapEncodeTiled unavailable / failed            */
#define PQLB_E_UNSUPPORTED (-5)

typedef void* pqlb_stream_t; /* cudaStream_t */

int pqlb_version(void);
const char* pqlb_error_string(int code);
/* One-time per-device setup (shared-memory opt-in, driver entry points).  Implicit in the first
 * pqlb_gemm_tf32 call; call it explicitly before capturing a CUDA graph. */
int pqlb_init(void);
/* Number of CUDA kernels this library has launched so far in this process. */
uint64_t pqlb_launch_count(void);
/* Measurement aid (bench.py's tensor-pipe roofline denominator), not on the path:
 * `iters` groups of four tcgen05.mma (M 128, N `n` = 128 | 256, 32 bytes of K each; kind 0 = tf32,
 * 1 = f16) on every SM from resident shared-memory operands; time it with CUDA events.
 * FLOPs per launch = 148 * iters * 4 * 2 * 128 * n * (8 | 16). */
int pqlb_mma_peak(int kind, int n, int iters, pqlb_stream_t stream);
/* Geometry helpers (pure host arithmetic). */
int pqlb_obs_pad(int obs_dim);
int pqlb_record_ld(int obs_dim, int act_dim);
/* Measurement switch (call before any ring exists): 0 = the default power-of-two stride, 1 = the next multiple of
 * 32 words (128 bytes): denser records (896 B instead of 1 KB for AllegroHand), slower random gathers. */
void pqlb_record_stride_mode(int mode);
int pqlb_x_ld(int obs_dim, int act_dim);

/* ---- K1: ring scatter-insert -------------------------------------------------------------
 * Replaces ReplayBuffer.add_to_buffer, pql/replay/simple_replay.py:40-83 (the ten slice
 * copies and the dones.bool() cast).  Row i of the n input rows goes to slot
 * (next_p + i) when it fits, and the LAST (next_p+n-capacity) rows go to slots [0, ...) when
 * the insert wraps (strict p > capacity).  Pointer bookkeeping stays on the host.
 * Inputs are [n,O] [n,A] [n] [n,O] [n] fp32, contiguous.  Requires n <= capacity + (capacity-next_p). */
int pqlb_ring_insert(float* ring, int64_t capacity, int obs_dim, int act_dim,
                     const float* obs, const float* action, const float* reward,
                     const float* next_obs, const float* done,
                     int64_t n, int64_t next_p, pqlb_stream_t stream);

/* Tests / measurements: 1 = force the LDG/STG insert kernel; 0 (default) = the TMA tile mover
 * whenever obs_dim % 4 == 0, act_dim % 4 == 0, 16-byte aligned inputs and n <= capacity. */
void pqlb_ring_insert_force_ldg(int on);

/* Observation-only ring of the P-learner, pql/algo/pql_p_learner.py:66-83.  ring is [capacity, O]. */
int pqlb_obsring_insert(float* ring, int64_t capacity, int obs_dim, const float* obs,
                        int64_t n, int64_t next_p, pqlb_stream_t stream);

/* ---- K1': n-step window push ---------------------------------------------------------------
 * Replaces NStepReplay.add_to_buffer + fifo_shift + compute_nstep_return,
 * pql/replay/nstep_replay.py:29-92.  Inputs [E,T,*] fp32 contiguous (env-major), window is
 * [E, nstep, rec_ld] records addressed circularly by global step index, `count` = number of
 * steps pushed before this call.  Emits (T - max(0, nstep-1-count)) * E rows, time-major
 * (row = w*E + e).  T == 1 is one launch; T > 1 is an emit launch followed by a window-store
 * launch (stream order keeps emitted rows independent of the new window).  gammas: nstep fp32
 * on the host. */
int pqlb_nstep_push(float* window, int num_envs, int nstep, int obs_dim, int act_dim,
                    const float* obs, const float* action, const float* reward,
                    const float* next_obs, const float* done, int T, int64_t count,
                    const float* gammas_host,
                    float* out_obs, float* out_action, float* out_reward, float* out_next_obs,
                    float* out_done, pqlb_stream_t stream);

/* ---- K2: uniform random-index gather ---------------------------------------------------------
 * Replaces the five buf[indices] gathers + .float() of ReplayBuffer.sample_batch,
 * pql/replay/simple_replay.py:98-104 (indices come from the caller: torch.randint keeps the
 * reference's Philox stream).  Outputs [B,O] [B,A] [B] [B,O] [B] fp32. */
int pqlb_sample_gather(const float* ring, int64_t capacity, int obs_dim, int act_dim,
                       const int64_t* idx, int64_t batch,
                       float* out_obs, float* out_action, float* out_reward,
                       float* out_next_obs, float* out_done, pqlb_stream_t stream);

/* Fused gather for the V-learner: sample_batch + normalize (pql/utils/common.py:139-145) +
 * torch.cat((state, action)) (pql/models/mlp.py:198), values rounded to TF32 (RN) because they
 * are tensor-core operands.  x_cur = [norm(obs) | action | 0], x_tgt[:, :O] = norm(next_obs)
 * (its action columns are written later by the actor head).  mean/var may be NULL (obs_norm off). */
int pqlb_sample_critic_batch(const float* ring, int64_t capacity, int obs_dim, int act_dim,
                             const int64_t* idx, int64_t batch,
                             const float* mean, const float* var, float eps,
                             float* x_cur, float* x_tgt, int x_ld,
                             float* reward, float* done, float* xf_cur, float* xf_tgt, pqlb_stream_t stream);
/* xf_cur / xf_tgt (optional, same shape as x_cur / x_tgt): the same rows WITHOUT the TF32 rounding -
 * the input of the split-fp16 forward (pqlb_mlp_forward_h), which splits fp32 values itself. */

/* Fused gather for the P-learner: memory[indices] + normalize, pql/algo/pql_p_learner.py:49-52.
 * Writes x[:, :O] (TF32-rounded) and zeroes the padding columns [O+A, x_ld). */
int pqlb_sample_obs_batch(const float* obsring, int64_t capacity, int obs_dim,
                          const int64_t* idx, int64_t batch,
                          const float* mean, const float* var, float eps,
                          float* x, int x_ld, int act_dim, float* xf, pqlb_stream_t stream);

/* ---- Fused sampler RNG ---------------------------------------------------------------
 * The same two fused gathers with the random draws of the update made INSIDE the kernel instead of
 * by two torch launches: indices = torch.randint(cur_capacity, (batch,)) (simple_replay.py:87,
 * pql_p_learner.py:49) and, for the V-learner, noise_out[noise_numel] = the N(0,1) draw behind
 * torch.normal(zeros, full(std)) (pql/utils/noise.py:20-21), bit-identical to what ATen produces for
 * the same generator state (csrc/rng.cuh: Philox4x32-10 / Box-Muller from cuRAND's device header,
 * ATen's launch policy).  All state is device-resident so that the launch can be replayed from a
 * CUDA graph: rng_state = {seed, base_offset, offset_increment_per_update}, counter = completed
 * updates (the optimiser kernel advances it), cur_capacity = the ring's fill level (pqlb_store_i64
 * after every insert).  The draw uses offset = base_offset + increment * counter[0] for the
 * indices and + 4 for the noise; idx_out receives the indices.  PQLB_E_UNSUPPORTED: capacity >= 2^28
 * (ATen switches to 64-bit draws) or a draw larger than one pass of ATen's grid. */
int pqlb_sample_critic_batch_rng(const float* ring, int64_t capacity, int obs_dim, int act_dim,
                                 int64_t* idx_out, int64_t batch,
                                 const float* mean, const float* var, float eps,
                                 float* x_cur, float* x_tgt, int x_ld, float* reward, float* done,
                                 const int64_t* rng_state, const int64_t* counter,
                                 const int64_t* cur_capacity, float* noise_out, int64_t noise_numel,
                                 float* xf_cur, float* xf_tgt, pqlb_stream_t stream);
int pqlb_sample_obs_batch_rng(const float* obsring, int64_t capacity, int obs_dim,
                              int64_t* idx_out, int64_t batch,
                              const float* mean, const float* var, float eps,
                              float* x, int x_ld, int act_dim,
                              const int64_t* rng_state, const int64_t* counter,
                              const int64_t* cur_capacity, float* xf, pqlb_stream_t stream);
/* *dst = value, stream-ordered (device-resident copies of host-side ring state). */
int pqlb_store_i64(int64_t* dst, int64_t value, pqlb_stream_t stream);
/* *dst = *src + delta, stream-ordered (the prefetching sampler's own update index: the index of the batch
 * it draws next, kept apart from the optimiser's counter, which advances concurrently). */
int pqlb_add_i64(int64_t* dst, const int64_t* src, int64_t delta, pqlb_stream_t stream);

/* ---- Actor-side env-step path (pql/algo/pql_actor.py:87-127) ---------------------------
 * RunningMeanStd.update (pql/utils/torch_util.py:77-103) on device-resident state: batch mean and
 * unbiased variance of x[rows, cols] (column statistics accumulated in fp64, fixed order), merged
 * into mean[cols] / var[cols] / count[0] (fp64, like the reference's Python float) with the
 * reference's update_from_moments arithmetic in fp32.  workspace: pqlb_rms_workspace_bytes(rows,
 * cols) bytes, 16-byte aligned, zero-filled before the FIRST call (the kernel re-arms it). */
int64_t pqlb_rms_workspace_bytes(int64_t rows, int cols);
int pqlb_rms_update(const float* x, int64_t rows, int cols, int64_t ldx, float* mean, float* var,
                    double* count, void* workspace, int64_t workspace_bytes, pqlb_stream_t stream);
/* The same update for data-parallel actors: pqlb_rms_moments writes this rank's column sums
 * and sums of squares (2 * cols doubles, slab order) instead of applying them; the caller all-reduces
 * them over the ranks (one 2 * cols fp64 sum per env step) and pqlb_rms_apply merges the moments of all
 * total_rows rows into mean / var / count - every rank takes the reference's update on the concatenated
 * batch and the normalisers stay bit-identical. */
int pqlb_rms_moments(const float* x, int64_t rows, int cols, int64_t ldx, double* sums, void* workspace,
                     int64_t workspace_bytes, pqlb_stream_t stream);
int pqlb_rms_apply(const double* sums, int64_t total_rows, int cols, float* mean, float* var, double* count,
                   pqlb_stream_t stream);
/* Inputs of the policy forward for one env step.  x (optional): x[r, :obs_dim] = obs normalised with
 * (mean, var, eps) as RunningMeanStd.normalize does ((x - mean) / sqrt(var + eps), torch_util.py:83-85;
 * clamp5 adds pql/utils/common.py:139-145's clamp to +-5; mean == NULL: plain copy), columns up to
 * x_ld zeroed, TF32-rounded with round_tf32 (tensor-core operand).  noise (optional):
 * noise[r, j] = z * std_r, z the N(0,1) draw ATen makes for a [rows, act_dim] tensor from the
 * generator state (seed, offset) - i.e. torch.normal(zeros, std) of add_normal_noise /
 * add_mixed_normal_noise (pql/utils/noise.py:19-41); std_r = row_std[r], or std when row_std == NULL.
 * The caller advances its generator offset by 4. */
int pqlb_actor_inputs(const float* obs, int64_t rows, int obs_dim, int64_t ld_obs, const float* mean,
                      const float* var, float eps, int clamp5, int round_tf32, float* x, int x_ld,
                      float* noise, int act_dim, const float* row_std, float std, int64_t seed,
                      int64_t offset, pqlb_stream_t stream);
/* After env.step(): PQLActor.update_tracker (pql_actor.py:129-136: returns += reward, lengths += 1,
 * finished episodes pushed in env order into the Tracker windows, then reset), handle_timeout
 * (common.py:195-202: done_out = done * ~truncated; truncated may be NULL) and the reward scaling of
 * pql_actor.py:117 (reward_out = reward_scale * reward).  ret_window / len_window: Tracker deques as
 * rings of window_len floats; pushed[0] = episodes pushed so far (slot = pushed % window_len). */
int pqlb_env_post(const float* reward, const float* done, const uint8_t* truncated, float reward_scale,
                  int num_envs, float* returns, float* lengths, float* ret_window, float* len_window,
                  int window_len, int64_t* pushed, float* reward_out, float* done_out,
                  pqlb_stream_t stream);

/* ---- K3: dense layers on tcgen05 (kind::tf32, fp32 accumulate in TMEM) ---------------------- */
enum pqlb_epilogue {
  PQLB_EPI_STORE = 0,          /* out = acc                     (split-K partial, no rounding)  */
  PQLB_EPI_BIAS = 1,           /* out = acc + bias                                             */
  PQLB_EPI_BIAS_ELU = 2,       /* out = rn_tf32(elu(acc + bias))            mlp.py:15-24       */
  PQLB_EPI_BIAS_ELU_HEAD = 3,  /* + q[row] = sum_n elu(..)*head_w[n] + head_b   (scalar Q head) */
  PQLB_EPI_BIAS_TANH = 4,      /* out = rn_tf32(tanh(acc + bias)), out2 = fp32  mlp.py:177-179  */
  PQLB_EPI_BIAS_TANH_NOISE = 5,/* out = rn_tf32(clamp(tanh(.)+clamp(noise,+-nb),+-1)) noise.py:19-27 */
  PQLB_EPI_BIAS_SOFTMAX = 6,   /* out = softmax(acc + bias) over n_valid columns  mlp.py:263    */
  PQLB_EPI_MUL_ELUGRAD = 7,    /* out = rn_tf32(acc * (h > 0 ? 1 : h + 1)), h = aux             */
  PQLB_EPI_MUL_TANHGRAD = 8    /* out = rn_tf32(acc * (1 - a*a)), a = aux                       */
};
enum pqlb_major { PQLB_K_MAJOR = 0, PQLB_MN_MAJOR = 1 };

#define PQLB_MAX_GROUPS 4

/* One member of a grouped GEMM  D[M,N] = epi(A . B)  (all groups share the shape).
 * K-major operand: memory [rows][K] (row stride ld words).  MN-major operand: memory [K][rows].
 * a2/b2 (optional) continue the contraction with a second operand pair (K2 more columns). */
typedef struct {
  const float* a; int64_t lda;
  const float* b; int64_t ldb;
  const float* a2; int64_t lda2;
  const float* b2; int64_t ldb2;
  const float* bias;          /* [N] or NULL                       */
  const float* aux;  int64_t ldaux;   /* epilogue input [M, ldaux]  */
  const float* head_w;        /* [N] for BIAS_ELU_HEAD             */
  const float* head_b;        /* [1]                               */
  float* q;                   /* [M] for BIAS_ELU_HEAD             */
  float* out;  int64_t ldo;   /* [M, ldo] or NULL (HEAD may skip)  */
  float* out2; int64_t ldo2;  /* second output (BIAS_TANH) or NULL */
  int64_t split_stride;       /* EPI_STORE: words between split-K partials */
} pqlb_gemm_group;

typedef struct {
  int M, N, K, K2;            /* logical sizes; contraction K (+K2)            */
  int a_major, b_major;       /* enum pqlb_major                               */
  int epilogue;               /* enum pqlb_epilogue                            */
  int tile_n;                 /* 16, 32, 64, 128 or 256                        */
  int splits;                 /* split-K factor (EPI_STORE only), else 1       */
  int n_groups;               /* 1..PQLB_MAX_GROUPS                            */
  int col_lo, col_hi;         /* only output columns [col_lo,col_hi) are stored, at out[row*ldo + n-col_lo];
                                 col_hi == 0 means [0, N).  aux/bias are indexed with the absolute n. */
  int cluster;                /* split-K only: 0/1 = every split stores its own partial; 2, 4 or 8 = the splits of
                                 a thread-block cluster are summed (in split order) through distributed shared
                                 memory before they leave the SMs, and out receives splits / cluster partials
                                 (partial p at out + p * split_stride).  splits % cluster == 0.            */
  float noise_bound;          /* BIAS_TANH_NOISE: clamp(noise_std * aux, +-noise_bound)         */
  float noise_std;            /* aux holds N(0,1) draws (out.normal_()), scaled here like
                                 torch.normal(zeros, full(std)) does (mul_(std).add_(mean))     */
  pqlb_gemm_group g[PQLB_MAX_GROUPS];
} pqlb_gemm_desc;

/* Replaces the nn.Linear / nn.ELU / tanh / softmax launches of pql/models/mlp.py:15-24,
 * 177-179, 197-199, 261-263 and their autograd backward (cuBLAS sgemm + ATen elementwise). */
int pqlb_gemm_tf32(const pqlb_gemm_desc* desc, pqlb_stream_t stream);

/* ---- K3f: layer-fused MLP trunk forward -------------------------------------------------------
 * x[M,k_in] -> ELU(Linear 512) -> ELU(Linear 256) -> ELU(Linear 128) [-> scalar head q[M]] for up
 * to PQLB_MAX_GROUPS network instances in one launch; activations stay in tensor memory between
 * layers (tcgen05.mma with the A operand in TMEM) and h1/h2/h3 are written only where non-NULL
 * (the networks whose backward pass needs them).  Replaces the Linear+ELU launches of
 * pql/models/mlp.py:15-24 as used by pql/algo/pql_v_learner.py:81-107 and pql_p_learner.py:55-56.
 * All operands TF32-rounded fp32; w1 [512, ldw1], w2 [512-in: 256 x 512], w3 [128 x 256] row-major
 * (nn.Linear layout); k_in <= 128 (wider inputs: PQLB_E_UNSUPPORTED, use pqlb_gemm_tf32). */
typedef struct {
  const float* x; int64_t ldx;
  const float* w1; int64_t ldw1;
  const float* w2; const float* w3;
  const float* b1; const float* b2; const float* b3;
  const float* head_w; const float* head_b; float* q;     /* q == NULL: no scalar head */
  float* h1; float* h2; float* h3;                         /* [M,512] [M,256] [M,128] or NULL */
  /* Optional policy head (pql/models/mlp.py:177-179, tanh(Linear(128, act_n))) as a fourth contraction
   * in the same launch: act_w [act_n, 128] (TF32 copy), act_b [act_n], act_n a multiple of 4 <= 16.
   * act_out[row * act_ldo + j] = rn_tf32(tanh(.)) or, with act_noise (N(0,1) draws, [M, act_ldnoise]),
   * rn_tf32(clamp(tanh(.) + clamp(noise_std * noise, +-noise_bound), +-1)) (pql/utils/noise.py:19-27);
   * act_out2 (optional) keeps the same value in fp32, before the TF32 rounding.  act_w == NULL: no policy head.  Mutually exclusive with q. */
  const float* act_w; const float* act_b; const float* act_noise;
  float* act_out; float* act_out2;
  int64_t act_ldo, act_ldo2, act_ldnoise;
  float noise_std, noise_bound;
  int act_n;
} pqlb_mlp_group;
typedef struct { int M, k_in, n_groups; pqlb_mlp_group g[PQLB_MAX_GROUPS]; } pqlb_mlp_desc;
int pqlb_mlp_forward(const pqlb_mlp_desc* desc, pqlb_stream_t stream);
/* Tuning / tests: 1 = one CTA per 128-row tile, 2 = CTA pairs (cta_group::2 MMAs, M = 256, each SM
 * streams half of every weight tile), 0 = the library default. */
void pqlb_mlp_forward_cluster(int cluster);

/* ---- K3f': the same trunk with split-fp16 operands ----------------------------------------------
 * Every GEMM operand is hi + lo (two fp16 values, 22 significand bits) and a product is evaluated as
 * a_hi.w_hi + a_lo.w_hi + a_hi.w_lo by three tcgen05 kind::f16 MMAs into one fp32 accumulator
 * (terms = 3), or as a_hi.w_hi alone (terms = 1: the accuracy of one TF32 MMA at twice its rate).
 * terms = 3 is what keeps critic losses and gradients within BASELINE.json's 1e-3 of the reference's
 * fp32 SGEMM path (pql/algo/pql_v_learner.py:81-108, pql_p_learner.py:55-57); see DESIGN.md section 4.
 * x is plain fp32 (NOT rounded; the kernel splits it), w?h / w?l are the fp16 hi / lo copies of the
 * nn.Linear weights, scaled by pqlb_f16_weight_scale(), row stride ldw halves as produced by
 * pqlb_split_f16 from a parameter arena (half i of a copy <-> element i of the arena; ldw1 % 8 == 0).
 * h1 / h2 / h3 receive TF32-rounded fp32 activations (the backward kernels' operands) where non-NULL.
 * Policy head as in pqlb_mlp_group; act_out (TF32-rounded) and act_out2 (unrounded) are both optional;
 * alternatively a C51 softmax head (sm_*).
 * Tile dependencies inside one launch (optional): a policy group with `publish` marks a 128-row tile
 * in desc.tile_sync once its action rows are written, and a LATER group with `wait` loads its input
 * tile only after that mark (target policy -> target critics in ONE launch: the SMs the 64 policy tiles
 * leave idle start on the critics).  tile_sync: 2 + ceil(M / 128) uint32, zero-initialised once by the
 * caller; the kernel maintains it across launches (no reset needed).
 * Input width: k_in <= 128 per group, or - if any group is wider - k_in <= 256 for the whole launch (the
 * wide-input kernel: eight 32-column input blocks resident in shared memory, a shorter weight ring;
 * ShadowHand's critics, obs 211 + act 20, pql/models/mlp.py:186-203); wider: PQLB_E_UNSUPPORTED.
 * In a wide launch the policy head takes act_n <= 32 (a multiple of 4) and output rows of any alignment. */
#define PQLB_MAX_FWD_GROUPS 5
typedef struct {
  const float* x; int64_t ldx;
  const void* w1h; const void* w1l; int64_t ldw1;
  const void* w2h; const void* w2l; const void* w3h; const void* w3l;
  const float* b1; const float* b2; const float* b3;
  const float* head_w; const float* head_b; float* q;
  float* h1; float* h2; float* h3;
  const void* act_wh; const void* act_wl; const float* act_b; const float* act_noise;
  float* act_out; float* act_out2;
  int64_t act_ldo, act_ldo2, act_ldnoise;
  float noise_std, noise_bound;
  int act_n;
  int terms;
  int k_in;                   /* this group's input width (0: the descriptor's k_in); groups may differ */
  /* Optional C51 head (pql/models/mlp.py:261-263, softmax(Linear(128, sm_n)), sm_n <= 64) as a fourth
   * contraction in the same launch: sm_wh / sm_wl [sm_n, 128] fp16 hi / lo, sm_b [sm_n];
   * sm_out[row * sm_ldp + j] = probability of atom j, columns [sm_n, 64) zeroed (sm_ldp >= 64).
   * Mutually exclusive with q and the policy head. */
  const void* sm_wh; const void* sm_wl; const float* sm_b; float* sm_out; int64_t sm_ldp; int sm_n;
  int publish, wait;
} pqlb_mlp_h_group;
typedef struct { int M, k_in, n_groups; uint32_t* tile_sync; pqlb_mlp_h_group g[PQLB_MAX_FWD_GROUPS]; } pqlb_mlp_h_desc;
int pqlb_mlp_forward_h(const pqlb_mlp_h_desc* desc, pqlb_stream_t stream);
/* Tuning / tests: 1 = one CTA per (network, 128-row tile), 2 = persistent CTAs (one per SM) that
 * software-pipeline consecutive tiles, 0 = the library default (1).  Same results bit for bit. */
void pqlb_mlp_forward_h_mode(int mode);
/* Debug: 64 uint64 of device memory receiving clock64 stamps of CTA (0,0) of the CTA-per-tile schedule
 * ([0,32) MMA issuer, [32,64) first conversion warp; tools/fwd_timeline.py); NULL switches it off. */
void pqlb_mlp_forward_h_debug(unsigned long long* buf);
/* fp16 operand copies of a parameter arena: hi[i] = fp16(s * src[i]), lo[i] = fp16(s * src[i] - hi[i])
 * with s = pqlb_f16_weight_scale() (lo may be NULL).  The optimiser entry points below keep such
 * copies up to date themselves (param_h / target_h). */
int pqlb_split_f16(const float* src, void* hi, void* lo, int64_t n, pqlb_stream_t stream);
float pqlb_f16_weight_scale(void);

/* ---- K3b: layer-fused dgrad chain of the trunk ---------------------------------------------------
 * dz2[M,256] = (dz3[M,128] . W3) * elu'(h2);  dz1[M,512] = (dz2 . W2) * elu'(h1), both TF32-rounded,
 * for up to PQLB_MAX_GROUPS network instances in one launch; dz2 stays in tensor memory between the
 * two contractions.  w3 [128 x 256] and w2 [256 x 512] are the nn.Linear weights (TF32 copies), read
 * as [K][N] without a transpose.  bias_part2 / bias_part1 (optional) receive the per-128-row-tile
 * column sums of dz2 / dz1: part[tile * n_cols + col], tile < ceil(M / 128) - the bias gradients'
 * partials for pqlb_grad_reduce.  Replaces the autograd backward of the Linear + ELU layers 2 and 3
 * (pql/models/mlp.py:15-24 under loss.backward(), pql_v_learner.py:125, pql_p_learner.py:60). */
typedef struct {
  const float* dz3;
  const float* w3; const float* w2;
  const float* h2; const float* h1;
  float* dz2; float* dz1;
  float* bias_part2; float* bias_part1;
} pqlb_mlp_bwd_group;
typedef struct { int M, n_groups; pqlb_mlp_bwd_group g[PQLB_MAX_GROUPS]; } pqlb_mlp_bwd_desc;
int pqlb_mlp_backward(const pqlb_mlp_bwd_desc* desc, pqlb_stream_t stream);

/* ---- K3w: all weight gradients of an update in one launch ------------------------------------------
 * Problem i: part_i[s][m][n] = sum over the k-blocks of split s of dz_i[k][m] * h_i[k][n], m < M (the
 * layer's output features), n < N (its input features), k < K (the batch); dz and h are the row-major
 * [K][ld] buffers the backward / forward kernels wrote (TF32-rounded values).  The CTAs of the launch are
 * the (128 x tile_n tile, split) pairs of all problems; split s of S owns k-blocks [s kb / S, (s + 1) kb / S)
 * (kb = ceil(K / 32)), so S need not divide kb; pqlb_grad_reduce sums the S partials in split order.
 * Replaces the weight-gradient GEMMs of loss.backward() (pql/algo/pql_v_learner.py:125,
 * pql/algo/pql_p_learner.py:60) for every nn.Linear of pql/models/mlp.py:15-24 at once. */
#define PQLB_MAX_WGRAD 8
typedef struct {
  const float* dz; int64_t lddz;
  const float* h;  int64_t ldh;
  float* part; int64_t ldo;       /* partial s at part + s * split_stride, row stride ldo words */
  int64_t split_stride;
  int M, N;
  int tile_n;                     /* 32, 64, 128 or 256 */
  int splits;
} pqlb_wgrad_problem;
typedef struct { int K, n_problems; pqlb_wgrad_problem p[PQLB_MAX_WGRAD]; } pqlb_wgrad_desc;
int pqlb_wgrad_multi(const pqlb_wgrad_desc* desc, pqlb_stream_t stream);

/* dst = rn_tf32(src) elementwise (tensor-core operand copies of weights). */
int pqlb_round_tf32(const float* src, float* dst, int64_t n, pqlb_stream_t stream);

/* ---- twin-Q TD target, loss and head gradient -------------------------------------------------
 * Replaces pql/algo/pql_v_learner.py:104-108 (torch.min, TD target, 2x mse_loss) and the
 * backward of the scalar head: y = r + (1-d)*gamma_n*min(tq1,tq2); loss = mean((q1-y)^2) +
 * mean((q2-y)^2); dq_i = 2(q_i-y)/B; dz3_i = rn_tf32(dq_i * w4_i * elu'(h3_i)).
 * Per-block partials (block = 64 rows of one net, nblk = ceil(B/64)): loss_part[2*nblk], and
 * gw4/gb4 partials written to ws_i[blk*129 + {0..127, 128}] for the deterministic reduction in
 * pqlb_grad_reduce. */
int pqlb_doubleq_td_loss(const float* q1, const float* q2, const float* tq1, const float* tq2,
                         const float* reward, const float* done, float gamma_n, int64_t batch,
                         const float* h3_1, const float* h3_2, const float* w4_1, const float* w4_2,
                         float* dz3_1, float* dz3_2, float* y_out,
                         float* ws_head1, float* ws_head2, float* loss_part, pqlb_stream_t stream);

/* The same, additionally writing the column sums of dz3 per 64-row block (the bias gradient of the
 * third hidden layer): bias3_part_i[blk * 128 + c]. */
int pqlb_doubleq_td_loss_b3(const float* q1, const float* q2, const float* tq1, const float* tq2,
                            const float* reward, const float* done, float gamma_n, int64_t batch,
                            const float* h3_1, const float* h3_2, const float* w4_1, const float* w4_2,
                            float* dz3_1, float* dz3_2, float* y_out,
                            float* ws_head1, float* ws_head2, float* loss_part,
                            float* bias3_part1, float* bias3_part2, pqlb_stream_t stream);

/* DPG actor loss, pql/algo/pql_p_learner.py:55-57: loss = -mean(min(q1,q2));
 * dq_i = -(1/B) on the smaller head (split 1/2 on ties, torch.minimum backward);
 * dz3_i = rn_tf32(dq_i * w4_i * elu'(h3_i)). */
int pqlb_dpg_loss(const float* q1, const float* q2, int64_t batch,
                  const float* h3_1, const float* h3_2, const float* w4_1, const float* w4_2,
                  float* dz3_1, float* dz3_2, float* loss_part, pqlb_stream_t stream);

/* C51: projection of both target distributions + elementwise min (pql/utils/distl_util.py:4-20,
 * pql/algo/pql_v_learner.py:83-102), BCE loss against both current distributions and the
 * gradient w.r.t. the logits through softmax.  p1,p2,tp1,tp2 are [B, ld] probabilities; dlogit1,2
 * are rounded to TF32.  Accumulation order inside a row equals the reference CPU order. */
int pqlb_c51_td_loss(const float* p1, const float* p2, const float* tp1, const float* tp2, int ld,
                     const float* reward, const float* done, const float* z_atoms, float gamma_n,
                     float v_min, float v_max, int num_atoms, int64_t batch,
                     float* target_out, float* dlogit1, float* dlogit2, int ld_d,
                     float* loss_part, pqlb_stream_t stream);

/* C51 actor loss, pql/algo/pql_p_learner.py:55-57 through DistributionalDoubleQ.get_q_min
 * (pql/models/mlp.py:255-259): q_i = sum_j p_i[j] z[j], loss = -mean(min(q1,q2)); the gradient
 * goes through the softmax: dlogit_i[j] = dq_i p_i[j] (z[j] - q_i), rounded to TF32.
 * dlogit1/2 may both be NULL (inference: only q_min_out / loss_part are written). */
int pqlb_c51_dpg_loss(const float* p1, const float* p2, int ld, const float* z_atoms,
                      int num_atoms, int64_t batch, float* dlogit1, float* dlogit2, int ld_d,
                      float* q_min_out, float* loss_part, pqlb_stream_t stream);

/* Column sums of dz (bias gradients), deterministic two-stage: part[blk*N + n] over 128-row blocks. */
int pqlb_colsum_partial(const float* dz, int64_t ld, int64_t rows, int n_cols, float* part,
                        pqlb_stream_t stream);

/* The same reduction for up to PQLB_MAX_COLSUM matrices (all bias gradients of one update) in
 * one launch. */
#define PQLB_MAX_COLSUM 8
typedef struct {
  int n; int64_t rows;
  const float* dz[PQLB_MAX_COLSUM]; int64_t ld[PQLB_MAX_COLSUM]; int n_cols[PQLB_MAX_COLSUM];
  float* part[PQLB_MAX_COLSUM];
} pqlb_colsum_desc;
int pqlb_colsum_partial_multi(const pqlb_colsum_desc* desc, pqlb_stream_t stream);

/* x[r] = rn_tf32([a[r,:na] | b[r,:nb] | 0]) with row stride x_ld: torch.cat((state, action), 1)
 * of pql/models/mlp.py:198,262 plus operand rounding (module-level forward; b may be NULL). */
int pqlb_pack_x(const float* a, int64_t lda, int na, const float* b, int64_t ldb, int nb,
                float* x, int x_ld, int64_t rows, pqlb_stream_t stream);

/* ---- K4: gradient reduction, clip, AdamW, Polyak ---------------------------------------------
 * seg table (device, int64 x 5 per segment): {arena_off, count, ws_off, ws_stride, n_part}.
 * pqlb_grad_reduce: grad[arena_off+i] = sum_{s<n_part} ws[ws_off + s*ws_stride + i] (fixed order)
 * and sumsq_part[seg] = sum of squares of the segment (fixed order). */
int pqlb_grad_reduce(const int64_t* seg_table, int n_seg, const float* ws, float* grad,
                     float* sumsq_part, pqlb_stream_t stream);
/* sum of squares only (after a gradient all-reduce). */
int pqlb_grad_sumsq(const int64_t* seg_table, int n_seg, const float* grad, float* sumsq_part,
                    pqlb_stream_t stream);
/* Replaces clip_grad_norm_ + AdamW.step + soft_update (pql/algo/pql_v_learner.py:124-133, :112;
 * pql/utils/torch_util.py:9-12; torch/optim/adam.py order):
 * coef = min(1, max_norm/(sqrt(sum sumsq_part)+1e-6)) (max_norm < 0: no clipping);
 * p,m,v updated in place; target <- p*tau + target*(1-tau) when target != NULL;
 * p_tf32 / target_tf32 = rn_tf32 copies (NULL to skip); param_h / target_h = fp16 hi | lo operand copies
 * for pqlb_mlp_forward_h (2 n halves each: hi at [0, n), lo at [n, 2n), as pqlb_split_f16 writes them;
 * NULL to skip; need n % 4 == 0).  The 1-based AdamW step count is `step`,
 * or step_dev[0] + 1 when step_dev != NULL (a device-resident count of completed updates, advanced
 * by pqlb_sum_partials, so that a whole update can be replayed from a CUDA graph).
 * grad_scale multiplies the gradient first (1/world for data parallel). */
int pqlb_adamw_polyak(float* param, const float* grad, float* m, float* v, float* target,
                      float* param_tf32, float* target_tf32, void* param_h, void* target_h, int64_t n,
                      const float* sumsq_part, int n_part, float grad_scale, float max_norm,
                      float lr, float beta1, float beta2, float eps, float weight_decay,
                      int64_t step, const int64_t* step_dev, float tau, float* grad_norm_out,
                      pqlb_stream_t stream);

/* Final deterministic sum of per-block loss partials: out[0] = scale * sum(part[0..n)).  When
 * counter != NULL: ring[counter[0] % ring_len] = out[0] (the loss window of the reference's
 * Tracker(5), pql/utils/common.py:103-126, kept on the device instead of loss.item()), then
 * ++counter[0]. */
int pqlb_sum_partials(const float* part, int n, float scale, float* out, int64_t* counter,
                      float* ring, int ring_len, pqlb_stream_t stream);

/* ---- K4, fused tail (what the learners launch): two kernels instead of four ---------------------
 * pqlb_grad_reduce_finish = pqlb_grad_reduce, whose last block also does pqlb_sum_partials' job
 * (loss_out[0] = loss_scale * sum(loss_part[0..n_loss)), ring[counter[0] % ring_len] = loss_out[0])
 * and precomputes the AdamW bias corrections of step counter[0] + 1 into scalars_out (16 floats).
 * It does NOT advance the counter.  pqlb_adamw_polyak_pre = pqlb_adamw_polyak taking those scalars
 * (no per-block double-precision pow) and advancing counter[0] by one when the step is applied. */
int pqlb_grad_reduce_finish(const int64_t* seg_table, int n_seg, const float* ws, float* grad,
                            float* sumsq_part, const float* loss_part, int n_loss, float loss_scale,
                            float* loss_out, const int64_t* counter, float* ring, int ring_len,
                            float lr, float beta1, float beta2, float eps, float weight_decay,
                            float tau, float* scalars_out, pqlb_stream_t stream);
int pqlb_adamw_polyak_pre(float* param, const float* grad, float* m, float* v, float* target,
                          float* param_tf32, float* target_tf32, void* param_h, void* target_h, int64_t n,
                          const float* sumsq_part, int n_part, float grad_scale, float max_norm,
                          const float* scalars, int64_t* counter, float* grad_norm_out,
                          pqlb_stream_t stream);

/* ---- K4-DP: gradient all-reduce fused into the optimiser kernel (one process per GPU) ----------
 * grad_peers[r] / red_peers[r] / ctl_peers[r]: rank r's gradient arena, reduced-gradient receive
 * buffer (n floats each) and control block (32 words of flags + world * grid floats), all in
 * peer-mapped (symmetric) memory; local: eight int64 in local memory, zero-initialised ([0] epoch, [1] blocks done, [2..5] accumulated phase clocks in ns).  Every rank
 * launches the same grid (<= 148 blocks, all co-resident).  Replaces ncclAllReduce(grad) +
 * pqlb_grad_sumsq + pqlb_adamw_polyak_pre: slice `rank` of the gradient is summed over ranks in rank
 * order by its owner and delivered to every rank (two-shot all-reduce over NVLink), the global norm is
 * assembled from the world x grid partial sums in a fixed order, and every rank applies the same
 * clip + AdamW + Polyak step of the mean gradient (grad_scale = 1 / world) to its full arena. */
typedef struct {
  const float* grad_peers[8];
  float* red_peers[8];
  void* ctl_peers[8];
  int64_t* local;
  int rank, world, grid;
  /* Optional NVLS multicast addresses of the same gradient arenas / receive buffers (both or neither): the
   * slice reduction then runs inside the NVSwitch (multimem.ld_reduce / multimem.st), one NVLink round trip. */
  const float* grad_mc; float* red_mc;
} pqlb_dp_desc;
/* How long (wall-clock seconds, default 120) a rank's exchange kernel waits for a peer's flag before it traps: the
 * ranks of a fused exchange must reach each update within this bound of one another (current device; not stream-ordered). */
int pqlb_dp_spin_limit(double seconds);
int pqlb_adamw_polyak_dp(float* param, float* m, float* v, float* target, float* param_tf32,
                         float* target_tf32, void* param_h, void* target_h, int64_t n,
                         const pqlb_dp_desc* dp, float max_norm,
                         const float* scalars, int64_t* counter, float* grad_norm_out,
                         pqlb_stream_t stream);

/* The exchange alone (phases "wait for every rank's gradient", "reduce + deliver the own slice", "wait for the
 * other slices") as a NARROW launch of dp->grid blocks: afterwards red_peers[rank] holds the gradient summed
 * over the ranks and the floats behind the 32 flag words of ctl_peers[rank] hold world * grid partial sums of
 * squares, to be consumed by pqlb_adamw_polyak_pre (grad = red, sumsq_part = ctl + 32 words, n_part =
 * world * grid, grad_scale = 1 / world) at full width.  The spinning blocks then occupy dp->grid SMs instead of
 * sharing 64 with the optimiser pass. */
int pqlb_grad_exchange_dp(int64_t n, const pqlb_dp_desc* dp, pqlb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PQLB200_H */
