/* include/dfrl.h -- C ABI of libdfrl_b200.so, the B200 (sm_100a) implementation of the
 * on-policy bin-packing training loop of beehover/dependence_free_rl.
 *
 * The reference has no FFI: its "plugin surface" is the C++20 class hierarchy in
 * xylo/tensor.h, xylo/nn.h, xylo/rl.h, xylo/policy_gradient.h and apps/bin_packing/bin_packing.h.
 * Every entry point below names the reference interface (file:line, relative to the reference
 * root) whose work it performs on the device.  The host-side C++ mirror of those classes lives
 * in dependence_free_rl_b200/host/ and calls only this header; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types.
 *   - every function returns DFRL_OK (0) or a negative dfrl_status; the message of the last
 *     failure on the calling thread is dfrl_last_error().  The C++ mirror rethrows it as
 *     xeno::error (reference xeno/exception.h:12-23).
 *   - pointers named *_dev are device pointers obtained from dfrl_malloc(); pointers named
 *     *_host are ordinary host memory.  All work is ordered on the context's stream;
 *     functions that return data to the host synchronise that stream, others do not.
 *   - there is NO CPU fallback: every compute entry point fails with DFRL_ERR_CUDA when no
 *     sm_100 device is usable.
 *
 * Data layouts (struct-of-arrays over N environments, B bins)
 *   state   int8  [2B+2][N]  plane 2b = remaining width of bin b, 2b+1 = remaining height,
 *                            plane 2B = item width, 2B+1 = item height
 *                            (reference bp::observation, bin_packing.h:16-44)
 *   obs     fp32  [rows][4B] row = [bin.w/8, bin.h/8, item.w/8, item.h/8] per bin
 *                            (observation::to_vector, bin_packing.h:31-40)
 *   params  fp32  flat, per parametric layer [W: out x in row-major][b: out]
 *                            (matmul_layer, nn.h:56-70; model::parameters, nn.h:499-508)
 */
#ifndef DFRL_H_
#define DFRL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  DFRL_OK = 0,
  DFRL_ERR_INVALID = -1, /* bad argument / shape mismatch (reference: xeno::error, tensor.cc:41-69) */
  DFRL_ERR_CUDA = -2,    /* CUDA runtime failure or no usable sm_100 device */
  DFRL_ERR_NCCL = -3,
  DFRL_ERR_UNSUPPORTED = -4
} dfrl_status;

typedef struct dfrl_ctx dfrl_ctx;         /* one per process / GPU */
typedef struct dfrl_env dfrl_env;         /* batched bp::environment + bp::agent bookkeeping */
typedef struct dfrl_mlp dfrl_mlp;         /* xylo::model (sequential) with device parameters */
typedef struct dfrl_trainer dfrl_trainer; /* replay buffer + learner + optimizers */

/* Layer kinds, in the reference's vocabulary (nn.h). */
enum {
  DFRL_LAYER_DENSE = 0,      /* matmul_layer / full_layer, nn.h:60-110 */
  DFRL_LAYER_CONV1D_1 = 1,   /* convolution1d_1_layer, nn.h:113-194 (points = cols / in_ch) */
  DFRL_LAYER_RELU = 2,       /* relu_activation, nn.h:350-377 */
  DFRL_LAYER_SOFTMAX = 3,    /* softmax_layer, nn.h:379-422 (full Jacobian backward) */
  DFRL_LAYER_SOFTMAX_CE = 4  /* softmax_cross_entropy_layer, nn.h:424-431 (identity backward) */
};

/* Learners (policy_gradient.h). */
enum {
  DFRL_ALGO_REINFORCE = 0,    /* policy_gradient_learner, policy_gradient.h:89-148 */
  DFRL_ALGO_ACTOR_CRITIC = 1, /* actor_critic_learner, policy_gradient.h:150-287 */
  DFRL_ALGO_PPO = 2,          /* ppo_learner, policy_gradient.h:289-308 */
  DFRL_ALGO_KL_PPO = 3        /* kl_ppo_learner, policy_gradient.h:310-335 */
};

/* Optimizers (nn.h:616-698). */
enum { DFRL_OPT_SGD = 0, DFRL_OPT_MOMENTUM = 1, DFRL_OPT_ADAM = 2 };

/* Loss-gradient rules (rl.h:45-74, policy_gradient.h:16-85, nn.h:548-550). */
enum {
  DFRL_LOSS_SOFTMAX_LOG = 0, /* discrete_action::softmax_gradient_log: A (p - onehot) */
  DFRL_LOSS_CLIPPED = 1,     /* discrete_action::clipped_gradient (PPO-clip, eps = 0.2) */
  DFRL_LOSS_KL = 2           /* kl_regulated_loss: A (p - onehot) + beta (p - p_old) */
};

/* Action sources for a rollout. */
enum {
  DFRL_ACT_SAMPLE = 0, /* discrete_action::from_vector (rl.h:27-30): inverse-CDF sample */
  DFRL_ACT_ARGMAX = 1, /* from_vector_deterministic (rl.h:31): first maximum */
  DFRL_ACT_FORCED = 2  /* teacher-forced from a tape (parity runs) */
};

/* Heuristic policies (apps/bin_packing/{random,firstfit,bestfit,minwaste}_agent.cc). */
enum { DFRL_HEUR_RANDOM = 0, DFRL_HEUR_FIRSTFIT = 1, DFRL_HEUR_BESTFIT = 2, DFRL_HEUR_MINWASTE = 3 };

/* ------------------------------------------------------------------ context / memory ----- */

const char *dfrl_last_error(void);
const char *dfrl_version(void);

/* 128-byte NCCL unique id for rank 0 to broadcast (K8; nothing to replace in the reference). */
int dfrl_nccl_unique_id(void *id128_host);

/* device: CUDA ordinal. nranks == 1: nccl_id may be NULL and no communicator is made.
 * Replaces the stubbed gpu_alloc()/gpu_dealloc() seam (tensor.cc:38-39) with a real device. */
int dfrl_init(int device, int nranks, int rank, const void *nccl_id128_host, dfrl_ctx **out);
int dfrl_destroy(dfrl_ctx *ctx);
int dfrl_sync(dfrl_ctx *ctx);
/* The context's stream as a cudaStream_t (void* to keep CUDA types out of this header). */
void *dfrl_stream(dfrl_ctx *ctx);
int dfrl_device_info(dfrl_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, size_t *hbm_bytes);

/* memory_blob(size, on_device = true) (tensor.cc:78-102). */
int dfrl_malloc(dfrl_ctx *ctx, size_t bytes, void **out_dev);
int dfrl_free(dfrl_ctx *ctx, void *ptr_dev);
int dfrl_memcpy_h2d(dfrl_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);
int dfrl_memcpy_d2h(dfrl_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);
int dfrl_memcpy_d2d(dfrl_ctx *ctx, void *dst_dev, const void *src_dev, size_t bytes);
int dfrl_memset(dfrl_ctx *ctx, void *dst_dev, int byte, size_t bytes);
/* Pinned (page-locked) host memory for the e2e path's tapes and results. */
int dfrl_malloc_host(dfrl_ctx *ctx, size_t bytes, void **out_host);
int dfrl_free_host(dfrl_ctx *ctx, void *ptr_host);
/* Timing helper: device milliseconds of everything enqueued on the stream between the calls. */
int dfrl_timer_start(dfrl_ctx *ctx);
int dfrl_timer_stop(dfrl_ctx *ctx, float *ms_out);
/* Number of kernels this library has launched on this context so far. */
long long dfrl_launch_count(dfrl_ctx *ctx);
/* Per-kernel device timing with CUDA events on the context's stream (bench.py's roofline leg).
 * While enabled every launch is bracketed by an event pair (adds launch overhead: never enable
 * it inside a throughput measurement).  dfrl_profile_report synchronises and writes one line
 * per kernel name, "name launches total_ms\n", into buf (truncated to cap); returns DFRL_OK. */
int dfrl_profile_enable(dfrl_ctx *ctx, int on);
int dfrl_profile_report(dfrl_ctx *ctx, char *buf, size_t cap);

/* Self-test of the tcgen05 / TMEM building blocks the fused MLP kernels are made of: one 128-row
 * tile GEMM with bf16 hi/lo split operands in SWIZZLE_128B shared-memory panels, checked against
 * a host fp64 reference. variant 0: X.W^T (both K-major), 1: dY.W (B MN-major), 2: G^T.H (both
 * MN-major, contraction over the 128 tile rows), 3 / 4: the same as 2 / 1 with a 16-column operand
 * whose hi and lo halves share one panel (descriptor start offsets 0 / 32 B). Writes
 * max|D-ref| / max|ref|. */
int dfrl_umma_selftest(dfrl_ctx *ctx, int variant, int k, int n, float *rel_err);
/* Profiling aid: SM cycles for `count` back-to-back tcgen05.mma (bf16, K = 16, M x N, operand
 * major-ness a_mn / b_mn: 0 = K-major, 1 = MN-major, a_mn = 2: A from tensor memory) on one SM; cycles2[0] = first issue -> all
 * complete, cycles2[1] = issue loop only. */
int dfrl_umma_microbench(dfrl_ctx *ctx, int M, int N, int a_mn, int b_mn, int count, long long *cycles2);

/* ------------------------------------------------------------- K1: batched environment ---- */

typedef struct {
  int n_envs;          /* environments on this rank */
  int n_bins;          /* bp::num_bins (bin_packing.h:12), 1..64 */
  int cap_w, cap_h;    /* observation::capacity (bin_packing.h:19), default 8, 8 */
  int item_w[2];       /* shape1.w, shape2.w (bin_packing.h:73-74): 4, 1 */
  int item_h[2];       /* shape1.h, shape2.h: 2, 2 */
  float p_shape1;      /* bernoulli p (bin_packing.h:50): 0.4 */
  uint64_t seed;       /* Philox key for free-running items and sampling uniforms */
  int64_t env_offset;  /* global index of local env 0 (multi-GPU sharding by env) */
} dfrl_env_config;

void dfrl_env_config_default(dfrl_env_config *cfg);

/* environment() for N envs: every bin at capacity, first item drawn (bin_packing.h:50-52). */
int dfrl_env_create(dfrl_ctx *ctx, const dfrl_env_config *cfg, dfrl_env **out);
int dfrl_env_destroy(dfrl_env *env);
/* reset(id) for all ids (bin_packing.h:67-70); also rewinds item tapes and step counters. */
int dfrl_env_reset(dfrl_env *env);
/* Parity mode: env i draws its k-th item (k = 0 at construction, +1 per step) from
 * tape[i * len + k]; byte 1 = shape1, 0 = shape2. NULL/0 returns to Philox items. */
int dfrl_env_load_item_tape(dfrl_env *env, const uint8_t *tape_host, int len);
/* apply(action, id) + agent::step bookkeeping for every env (bin_packing.h:53-64, 94-106;
 * rl.h:333-346): bin[a] -= item; done = overflow; reward = done ? 0 : 1; on done the terminal
 * state is what `terminal_state_dev` (optional, int8 [2B+2][N], written for every env = state
 * after apply and before reset) records and the env is reset with a fresh item; otherwise the
 * next item is drawn. actions_dev: uint8 [N]. done_dev (optional): uint8 [N]. */
int dfrl_env_step(dfrl_env *env, const uint8_t *actions_dev, uint8_t *done_dev,
                  int8_t *terminal_state_dev);
/* The reference triple for ONE environment id (bin_packing.h:53-70; rl.h:163-170): apply() leaves
 * an overflowed bin negative and draws no item, reset() refills the bins and draws the next item,
 * view() copies the int8 [2B+2] state of `id` to the host (synchronising). The batched calls above
 * are the fast path; these keep environment<A,S>'s per-id signatures usable. */
int dfrl_env_apply_one(dfrl_env *env, int id, int action);
int dfrl_env_reset_one(dfrl_env *env, int id);
int dfrl_env_view_one(dfrl_env *env, int id, int8_t *state_host /* [2B+2] */);
/* view(id) for all ids (bin_packing.h:65): device pointer to the int8 [2B+2][stride] planes;
 * stride = N rounded up to 16 so that every plane starts 16-byte aligned. Device-side buffers
 * shaped like the state (terminal_state_dev) use the same stride. */
int8_t *dfrl_env_state_dev(dfrl_env *env);
int dfrl_env_state_stride(dfrl_env *env);
int dfrl_env_get_state(dfrl_env *env, int8_t *state_host /* [2B+2][N] */);
int dfrl_env_set_state(dfrl_env *env, const int8_t *state_host);
/* observation::to_vector for `rows` states (bin_packing.h:31-40). state planes have row
 * stride `stride` (= N for the live state). obs_dev: fp32 [rows][4B]. */
int dfrl_obs_encode(dfrl_ctx *ctx, const int8_t *state_dev, int rows, int stride, int n_bins,
                    int cap_w, int cap_h, float *obs_dev);
/* Heuristic policy::react for every env (firstfit/bestfit/minwaste/random agents). */
int dfrl_heuristic_react(dfrl_env *env, int kind, uint8_t *actions_dev);
/* Plays every env with a heuristic until each has finished `episodes` episodes
 * (agent::play_one_episode, rl.h:351-354). Outputs: total reward and transitions. */
int dfrl_heuristic_play(dfrl_env *env, int kind, int episodes, double *total_reward,
                        long long *env_steps);

/* -------------------------------------------------- K2/K6: layers (xylo::layer, nn.h:20-33) */

/* matmul_layer::forward (nn.h:72-79): y[rows][out] = x[rows][in] . W^T + b.
 * params_dev = [W out x in][b out]. fuse_relu != 0 additionally applies relu_activation. */
int dfrl_dense_forward(dfrl_ctx *ctx, const float *params_dev, int in, int out, const float *x_dev,
                       int rows, float *y_dev, int fuse_relu);
/* matmul_layer::backward (nn.h:81-83): dx[rows][in] = dy[rows][out] . W.
 * relu_mask_dev (optional, [rows][in]): dx is zeroed where mask <= 0 (relu_activation::backward
 * of the preceding activation, nn.h:364-376, fused). */
int dfrl_dense_backward(dfrl_ctx *ctx, const float *params_dev, int in, int out,
                        const float *dy_dev, int rows, const float *relu_mask_dev, float *dx_dev);
/* matmul_layer::gradient (nn.h:85-100): grad = [dW = dy^T . x (SUM over rows)][db = sum dy].
 * Deterministic two-stage reduction. accumulate != 0 adds into grad_dev. */
int dfrl_dense_gradient(dfrl_ctx *ctx, int in, int out, const float *x_dev, const float *dy_dev,
                        int rows, float *grad_dev, int accumulate);
/* relu_activation::forward / backward (nn.h:354-376). */
int dfrl_relu_forward(dfrl_ctx *ctx, const float *x_dev, size_t n, float *y_dev);
int dfrl_relu_backward(dfrl_ctx *ctx, const float *x_dev, const float *dy_dev, size_t n,
                       float *dx_dev);
/* softmax_layer::forward (nn.h:382-392): exp(x) / sum exp(x), NO max subtraction. */
int dfrl_softmax_forward(dfrl_ctx *ctx, const float *x_dev, int rows, int cols, float *y_dev);
/* softmax_layer::backward (nn.h:393-417): dx = (diag(s) - s s^T) dy with s = softmax(x). */
int dfrl_softmax_backward(dfrl_ctx *ctx, const float *x_dev, const float *dy_dev, int rows,
                          int cols, float *dx_dev);

/* xylo::model (nn.h:467-542). conv1d layers take `points` from the preceding width. */
int dfrl_mlp_create(dfrl_ctx *ctx, int n_layers, const int *kinds, const int *ins, const int *outs,
                    int input_cols, dfrl_mlp **out);
/* Shared-trunk model (BASELINE.json configs[2] "shared-trunk policy/value MLP"; an EXTENSION: the
 * reference's model is strictly sequential, nn.h:467-542, and ac_training.cc:9-25 builds two nets).
 * The new model's first n_shared layers ARE the first n_shared layers of `trunk` (same parameters,
 * same device memory), followed by n_layers own head layers. Every model of such a family addresses
 * ONE flat parameter vector [trunk model's layers | head 1 | head 2 ...]: param_count / get / set see
 * the whole vector, a model's flat gradient carries zeros in the other heads' slots, init_params of a
 * sharer touches its own head only. Create sharers before any trainer; destroy them before `trunk`. */
int dfrl_mlp_create_shared(dfrl_mlp *trunk, int n_shared, int n_layers, const int *kinds, const int *ins,
                           const int *outs, dfrl_mlp **out);
int dfrl_mlp_destroy(dfrl_mlp *mlp);
int dfrl_mlp_param_count(dfrl_mlp *mlp);
int dfrl_mlp_output_cols(dfrl_mlp *mlp);
float *dfrl_mlp_params_dev(dfrl_mlp *mlp);
/* model::set_parameters / parameters (nn.h:490-508); also the checkpoint format of
 * apps/bin_packing/weights.{10,20} (deep_agent.cc:21-23). */
int dfrl_mlp_set_params(dfrl_mlp *mlp, const float *params_host, int n);
int dfrl_mlp_get_params(dfrl_mlp *mlp, float *params_host, int n);
/* nn.h:12-18 initialisation (dense: N(0, 0.01); conv1d: He), Philox-driven. */
int dfrl_mlp_init_params(dfrl_mlp *mlp, uint64_t seed);
/* model::eval (nn.h:473-479). x_dev [rows][input_cols] -> y_dev [rows][output_cols]. */
int dfrl_mlp_eval(dfrl_mlp *mlp, const float *x_dev, int rows, float *y_dev);
/* optimizer::step's forward + gradient (nn.h:594-603, 510-528) for a caller-supplied loss
 * gradient dy_dev at the model output: grad_dev receives the flat gradient (SUM over rows);
 * out_dev (optional) receives the forward output. */
int dfrl_mlp_forward_gradient(dfrl_mlp *mlp, const float *x_dev, int rows, const float *dy_dev,
                              float *grad_dev, float *out_dev);

/* ------------------------------------------------------------ K3: sampling / argmax ------- */

/* discrete_distribution (tensor.cc:467-470; libstdc++ random.tcc:2657-2714): normalise in
 * double, partial sums, last forced to 1.0, lower_bound(cum, u).  u_dev: one double in [0,1)
 * per row. actions_dev: uint8 [rows]; p_sel_dev (optional): probs[row][action]. */
int dfrl_sample(dfrl_ctx *ctx, const float *probs_dev, int rows, int cols, const double *u_dev,
                uint8_t *actions_dev, float *p_sel_dev);
/* argmax, first maximum wins (tensor.cc:464-466). */
int dfrl_argmax(dfrl_ctx *ctx, const float *probs_dev, int rows, int cols, uint8_t *actions_dev);

/* ------------------------------------------------------------ K4: returns / GAE ----------- */

/* policy_gradient_learner::get_advantages (policy_gradient.h:125-147) on [L][N] step-major
 * records: per env, transitions 0..len[i]-1 split into trajectories by done flags.
 * Reversed-order quirk reproduced: G[first + (m-1-k)] = sum_{j<=k} gamma^{k-j} r_j.
 * baseline_sum/baseline_cnt (device, 2 doubles) receive sum of G[first] and #trajectories of
 * this rank; pass `baseline` (total_sum / total_count over all ranks) to the second pass. */
int dfrl_returns(dfrl_ctx *ctx, const uint8_t *done_dev, const int *len_dev, int n_envs, int max_len,
                 float gamma, float *g_dev /* [L][N] */, double *baseline_acc_dev /* [2] */);
int dfrl_subtract_baseline(dfrl_ctx *ctx, float *g_dev, const int *len_dev, int n_envs, int max_len,
                           float baseline);
/* actor_critic_learner::update_value_model targets (policy_gradient.h:196-215) and
 * calculate_advantage (220-281) on [T][N] records.
 *   v_start_dev [T][N]   V(s_t)
 *   v_end_dev   [T][N]   V(end state of step t); read only where done or t == T-1
 * targets: tgt = r + gamma * V_next, NOT masked at terminals (quirk 6).
 * advantages: V_next = 0 where done; delta = r + gamma V_next - V; A_t = delta_t +
 * gamma lambda A_{t+1} within a trajectory (quirk 7/8 handled by the caller passing values of
 * the UPDATED critic). Either output may be NULL. */
int dfrl_gae(dfrl_ctx *ctx, const uint8_t *done_dev, const float *v_start_dev,
             const float *v_end_dev, int n_envs, int T, float gamma, float lambda,
             float *targets_dev, float *adv_dev);

/* ------------------------------------------------------------ K5: loss gradients ---------- */

/* policy_loss / surrogate_loss / kl_regulated_loss rows (policy_gradient.h:16-85) on the model
 * output `probs` [rows][cols]. p_old_dev: [rows] (selected prob) for CLIPPED, [rows][cols] for
 * KL, unused for SOFTMAX_LOG. out_dev [rows][cols]. */
int dfrl_loss_grad(dfrl_ctx *ctx, int kind, const float *probs_dev, const uint8_t *actions_dev,
                   const float *adv_dev, const float *p_old_dev, float beta, int rows, int cols,
                   float *out_dev);
/* square_loss_grad (nn.h:548-550): out = v - target. */
int dfrl_square_loss_grad(dfrl_ctx *ctx, const float *v_dev, const float *target_dev, int rows,
                          float *out_dev);

/* ------------------------------------------------------------ K7: optimizers -------------- */

/* sgd_optimizer / momentum_optimizer / adam_optimizer::next_parameters (nn.h:616-698), in
 * place on params_dev. state_dev: momentum -> velocity [n]; adam -> [m n][v n]; sgd -> NULL.
 * adam_t is the reference's float step counter (starts at 1, caller increments). */
int dfrl_opt_step(dfrl_ctx *ctx, int kind, float *params_dev, const float *grad_dev,
                  float *state_dev, int n, float lr, float weight_decay, float beta1, float beta2,
                  float adam_t);

/* ------------------------------------------------------------ K8: gradient all-reduce ------ */

/* SUM of a flat fp32 buffer over all ranks, in place (no-op for nranks == 1). */
int dfrl_allreduce_sum(dfrl_ctx *ctx, float *buf_dev, size_t n);
int dfrl_allreduce_sum_f64(dfrl_ctx *ctx, double *buf_dev, size_t n);
int dfrl_barrier(dfrl_ctx *ctx);
/* The same exchange over NVLink peer memory, fused with the optimizer (fused.cu): every rank of the
 * node exports a 64-byte CUDA IPC handle of its exchange buffer, the host gathers them by any
 * transport into [nranks][64] and every rank attaches. Once attached, the fused learner kernels
 * publish their reduced gradient in the local buffer and ONE kernel per optimizer step pulls the
 * peers' buffers over NVLink, sums them in rank order (bit-identical on every rank), applies the
 * optimizer update and rebuilds the weight panels -- no NCCL call on that path. */
int dfrl_p2p_export(dfrl_ctx *ctx, void *handle64_host);
int dfrl_p2p_attach(dfrl_ctx *ctx, const void *all_handles_host /* [nranks][64] */);
int dfrl_p2p_attached(dfrl_ctx *ctx);

/* ------------------------------------------------------------ trainer (fused loop) -------- */

typedef struct {
  int algo;            /* DFRL_ALGO_* */
  int work;            /* steps per env per iteration (AC/PPO: play_steps(n), rl.h:356-360)
                          or episodes per env per iteration (REINFORCE: play_one_episode) */
  float gamma;         /* learner gamma (bin_packing.h ctor default 0.99) */
  float lambda;        /* actor_critic_learner::lambda_ (policy_gradient.h:286) 0.95 */
  int epochs;          /* ppo k (policy_gradient.h:300, 321) 4 */
  float kl_target;     /* kl_ppo_learner::d_targ_ (334) 1e-9 */
  float kl_beta0;      /* kl_ppo_learner::beta_ (333) 1 */
  int policy_opt, value_opt;     /* DFRL_OPT_* */
  float policy_lr, value_lr;
  float policy_wd, value_wd;     /* sgd weight decay (nn.h:618) */
  float adam_beta1, adam_beta2;  /* nn.h:661 */
  int action_mode;     /* DFRL_ACT_* used by rollouts */
  int fused;           /* 1: fused small-MLP kernels when the nets qualify; 0: layered kernels */
} dfrl_trainer_config;

void dfrl_trainer_config_default(dfrl_trainer_config *cfg);

/* learner ctor (policy_gradient.h:92-94, 153-157, 292-296, 313-317): borrows env, models.
 * value may be NULL for REINFORCE. */
int dfrl_trainer_create(dfrl_ctx *ctx, const dfrl_trainer_config *cfg, dfrl_env *env,
                        dfrl_mlp *policy, dfrl_mlp *value, dfrl_trainer **out);
int dfrl_trainer_destroy(dfrl_trainer *tr);
/* optimizer::set_rate (nn.h:592) after the learner exists: the learning rates (and weight decays) of the
 * next learn(). Optimizer state (momentum / Adam moments, Adam step counter), the KL beta and the
 * statistics are kept; a captured CUDA graph of the learn phase is dropped (rates are kernel arguments)
 * and re-captured by the next graph-eligible learn(). */
int dfrl_trainer_set_rates(dfrl_trainer *tr, float policy_lr, float policy_wd, float value_lr, float value_wd);

/* Rollout phase of the trainer mains (ppo_training.cc:55-61): every env plays `work` steps
 * (or episodes). Optional host tapes, all [work][N] step-major (AC/PPO only):
 *   items_host   uint8  item drawn after step t (1 = shape1); NULL = tape/Philox of the env
 *   actions_host uint8  forced actions (requires action_mode FORCED)
 *   u_host       double sampling uniforms; NULL = Philox
 * Host tapes are copied to the device inside this call (they are the e2e inputs). */
int dfrl_trainer_rollout(dfrl_trainer *tr, const uint8_t *items_host, const uint8_t *actions_host,
                         const double *u_host);
/* learner::step() then replay_buffer::forget() (ppo_training.cc:63-65). */
int dfrl_trainer_learn(dfrl_trainer *tr);
/* The phases of actor_critic_learner::learn (policy_gradient.h:159-185) one by one, for the host
 * mirror's optimize_action hook (policy_gradient.h:187-194, 297-307, 318-330 are virtual overrides of
 * it): VALUE = update_value_model (196-218), ADVANTAGE = calculate_advantage with the updated critic
 * (220-281) -- these two run together --, POLICY = the learner's own optimize_action (one step, k
 * PPO-clip steps, k KL-PPO steps). DFRL_PHASE_ALL is dfrl_trainer_learn. */
enum { DFRL_PHASE_VALUE = 1, DFRL_PHASE_ADVANTAGE = 2, DFRL_PHASE_POLICY = 4, DFRL_PHASE_ALL = 7 };
int dfrl_trainer_learn_phases(dfrl_trainer *tr, int phases);
/* `iters` x (rollout; learn) free-running, no host round trips between iterations. */
int dfrl_trainer_iterate(dfrl_trainer *tr, int iters);

/* Introspection for parity tests (all device->host, synchronising). */
enum {
  DFRL_F_REC_STATE = 0,   /* int8  [L][2B+2][N] start state of every recorded step */
  DFRL_F_REC_ACTION = 1,  /* uint8 [L][N] */
  DFRL_F_REC_DONE = 2,    /* uint8 [L][N] */
  DFRL_F_REC_PROBS = 3,   /* fp32  [L][N][B] policy output at the start state (p_old) */
  DFRL_F_REC_LEN = 4,     /* int32 [N] recorded steps per env this iteration */
  DFRL_F_ADVANTAGE = 5,   /* fp32  [L][N] */
  DFRL_F_VALUE_TARGET = 6,/* fp32  [L][N] */
  DFRL_F_POLICY_GRAD = 7, /* fp32  [P] last policy gradient (after all-reduce) */
  DFRL_F_VALUE_GRAD = 8,  /* fp32  [Pv] last value gradient */
  DFRL_F_POLICY_GRAD_LOG = 9, /* fp32 [epochs][P] every policy gradient of the last learn() */
  DFRL_F_OBS_START = 10   /* fp32  [L][N][4B] observation::to_vector of every recorded start state */
};
int dfrl_trainer_field_size(dfrl_trainer *tr, int field, size_t *bytes);
int dfrl_trainer_read(dfrl_trainer *tr, int field, void *dst_host, size_t bytes);

typedef struct {
  long long env_steps;     /* transitions made since creation (this rank) */
  long long episodes;      /* episodes finished since creation (this rank) */
  double reward_sum;       /* sum of rewards since creation (this rank) */
  double last_mean_reward; /* mean reward per step since the previous get_stats() call */
  float kl_beta;           /* current beta (KL-PPO) */
} dfrl_trainer_stats;
int dfrl_trainer_get_stats(dfrl_trainer *tr, dfrl_trainer_stats *out);
/* Asynchronous form: _begin enqueues the device->host read of the counters as they stand after the
 * work submitted so far and returns immediately; _end waits for the OLDEST begun read only (at most
 * 4 in flight). Loop: rollout(i); learn(i); stats_begin(); [submit step i + 1;] stats_end(&s_i) --
 * the device stays busy while the host prepares the next step (the reference's trainer mains print
 * the mean reward of finished iterations the same way: ppo_training.cc:67-81). */
int dfrl_trainer_stats_begin(dfrl_trainer *tr);
int dfrl_trainer_stats_end(dfrl_trainer *tr, dfrl_trainer_stats *out);
/* Profiling aid (no reference counterpart): SM-cycle stamps at the phase boundaries of CTA 0 of the
 * fused policy-step kernel (pipeline 0), 13 per row tile for its first 8 tiles, then entry / setup /
 * loop-end / exit stamps at 104..107 (n <= 112). The first call arms
 * the instrumentation and returns zeros; later calls return the stamps of the last launch. */
int dfrl_debug_policy_clocks(dfrl_trainer *tr, long long *out_host, int n);
/* The same for the fused critic-step kernel: 15 stamps per row tile, first 7 tiles (n <= 112). */
int dfrl_debug_critic_clocks(dfrl_trainer *tr, long long *out_host, int n);
/* Test hook (no reference counterpart): caps the number of persistent CTAs of the fused learner
 * kernels, so that a small problem runs many row tiles per CTA (the steady state of the two tile
 * pipelines). ctas <= 0 restores one CTA per SM. Results do not depend on the grid size beyond the
 * summation order of the per-CTA partial gradients. */
int dfrl_debug_set_fused_ctas(dfrl_trainer *tr, int ctas);
/* Test hook: 1 forces / 0 forbids the compacted V(end-state) pre-pass of the fused critic-step and
 * GAE kernels (policy_gradient.h:196-281 need V(end) only where a trajectory ends); -1: by batch size. */
int dfrl_debug_set_vend(dfrl_trainer *tr, int mode);
/* Which phases of this trainer run on the fused tcgen05 kernels (the rest runs on the layered
 * kernels): a mask of DFRL_FUSED_*. ROLLOUT = agent::play_steps (rl.h:325-360), CRITIC =
 * update_value_model + calculate_advantage (policy_gradient.h:196-281), POLICY = optimize_action
 * (187-194, 297-307), GRAPH = learn() replays as one CUDA graph. */
enum { DFRL_FUSED_ROLLOUT = 1, DFRL_FUSED_CRITIC = 2, DFRL_FUSED_POLICY = 4, DFRL_FUSED_GRAPH = 8 };
int dfrl_trainer_fused_coverage(dfrl_trainer *tr, int *mask);

/* deep_agent.cc:28-41 / the periodic eval of the trainer mains (ppo_training.cc:67-81): every
 * env of `env` plays `episodes` episodes with policy_gradient_deterministic_policy (argmax) on
 * `policy`. Returns mean reward per episode over all envs of this rank. */
int dfrl_eval_argmax(dfrl_ctx *ctx, dfrl_env *env, dfrl_mlp *policy, int episodes,
                     double *mean_reward, long long *env_steps);

#ifdef __cplusplus
}
#endif
#endif /* DFRL_H_ */
