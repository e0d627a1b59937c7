// trainer.cu -- the on-policy training loop of the reference trainer mains
// (apps/bin_packing/{pg,ac,ppo,ppo2}_training.cc) on the device: rollout -> learn -> forget.
//
// replay_buffer / trajectory / transition (xylo/rl.h:111-296) become struct-of-arrays rollout
// records, step-major:
//   rec_state  int8  [L][2B+2][stride]   start state of every recorded step
//   rec_action uint8 [L][N]              rec_done uint8 [L][N]  (reward = 1 - done)
//   rec_probs  fp32  [L][N][B]           policy output at the start state (action.distrib)
// A trajectory is a maximal run of steps of one env without a done flag; "forget()" is
// implicit: the next rollout overwrites the records and an open trajectory continues from the
// live env state (rl.h:274-291).
//
// End-state rows (policy_gradient.h:168-180): the reference appends one extra row per
// trajectory holding its end state.  Here step (t, i) owns an end row iff done[t][i] or
// t == L-1; its state is the overflowed terminal state (start state with bin[a] -= item) or
// the live env state.  End rows only matter through V(end) in the critic target / advantage
// (their own loss gradient is exactly zero: value target = own value, advantage = 0), except
// for KL-PPO where beta (p - p_old) is non-zero on them and they join the policy pass.
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "device_fns.cuh"
#include "env_dev.cuh"
#include "trainer.h"

namespace {

// One agent::step for every active env: choose the action from the policy output, record it,
// apply it, record done, reset / draw the next item (rl.h:325-349).
//   mode: DFRL_ACT_SAMPLE (u from tape or Philox), DFRL_ACT_ARGMAX, DFRL_ACT_FORCED.
//   episodic != 0 (REINFORCE / eval): env i is active while ep_done[i] < ep_target.
__global__ void act_step_kernel(env_params p, int8_t *__restrict__ state,
                                uint32_t *__restrict__ draws, uint32_t *__restrict__ steps,
                                const float *__restrict__ probs, int mode,
                                const uint8_t *__restrict__ forced, const double *__restrict__ u_tape,
                                const uint8_t *__restrict__ item_tape_step,
                                uint8_t *__restrict__ rec_action, uint8_t *__restrict__ rec_done,
                                int episodic, int ep_target, int *__restrict__ ep_done,
                                int *__restrict__ rec_len, unsigned long long *__restrict__ counters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long c_steps = 0, c_eps = 0, c_reward = 0, c_active = 0;
  if (i < p.n && (!episodic || ep_done[i] < ep_target)) {
    const float *pr = probs + (size_t)i * p.B;
    int a;
    if (mode == DFRL_ACT_FORCED) {
      a = forced[i];
    } else if (mode == DFRL_ACT_ARGMAX) {
      a = argmax_first(pr, p.B);
    } else {
      double u;
      if (u_tape)
        u = u_tape[i];
      else {
        philox4 r = philox4x32_10(p.seed, (uint64_t)(p.env_offset + i), steps[i], DFRL_STREAM_ACTION);
        u = philox_u53(r.x, r.y);
      }
      a = discrete_sample(pr, p.B, u);
    }
    a = a < p.B ? a : p.B - 1;
    if (rec_action)
      rec_action[i] = (uint8_t)a;
    uint32_t k = draws[i];
    env_params q = p;
    if (item_tape_step) {
      // per-step host tape [n]: the item drawn after this step (draw_shape1 reads tape[i*1 + 0])
      q.tape = item_tape_step;
      q.tape_len = 1;
      k = 0;
    }
    bool over = env_apply_global(q, state, i, a, k);
    draws[i] += 1;
    steps[i] += 1;
    if (rec_done)
      rec_done[i] = over;
    c_steps = 1;
    c_reward = over ? 0 : 1;
    if (over)
      c_eps = 1;
    if (episodic) {
      if (over)
        ep_done[i] += 1;
      if (rec_len)
        rec_len[i] += 1;
      c_active = (ep_done[i] < ep_target) ? 1 : 0;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    c_steps += __shfl_down_sync(0xffffffffu, c_steps, o);
    c_eps += __shfl_down_sync(0xffffffffu, c_eps, o);
    c_reward += __shfl_down_sync(0xffffffffu, c_reward, o);
    c_active += __shfl_down_sync(0xffffffffu, c_active, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (c_steps) atomicAdd(&counters[0], c_steps);
    if (c_eps) atomicAdd(&counters[1], c_eps);
    if (c_reward) atomicAdd(&counters[2], c_reward);
    if (c_active) atomicAdd(&counters[3], c_active);
  }
}

// Observations of the end rows: terminal state (done) / live state (last step) / next start
// state (otherwise; unused by the learner but finite).
__global__ void end_obs_kernel(const int8_t *__restrict__ rec_state, const int8_t *__restrict__ live,
                               const uint8_t *__restrict__ rec_action,
                               const uint8_t *__restrict__ rec_done, int n, int stride, int B, int L,
                               float cap_w, float cap_h, float4 *__restrict__ obs_end) {
  long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (long long)L * n * B)
    return;
  int b = (int)(tid % B);
  long long row = tid / B;
  int i = (int)(row % n), t = (int)(row / n);
  const int P = 2 * B + 2;
  int d = rec_done[row];
  const int8_t *src;
  if (d)
    src = rec_state + (size_t)t * P * stride;
  else if (t == L - 1)
    src = live;
  else
    src = rec_state + (size_t)(t + 1) * P * stride;
  int bw = src[(size_t)(2 * b) * stride + i], bh = src[(size_t)(2 * b + 1) * stride + i];
  int iw = src[(size_t)(2 * B) * stride + i], ih = src[(size_t)(2 * B + 1) * stride + i];
  if (d && b == rec_action[row]) {  // overflowed bin, item kept (bin_packing.h:54-61)
    bw -= iw;
    bh -= ih;
  }
  obs_end[tid] = make_float4((float)bw / cap_w, (float)bh / cap_h, (float)iw / cap_w, (float)ih / cap_h);
}

// Loss gradient at the policy output for the start rows (and, for KL, the end rows).
// rows = L*n start rows followed (KL only) by L*n end-row slots.
__global__ void policy_loss_kernel(int kind, const float *__restrict__ probs,
                                   const uint8_t *__restrict__ rec_action,
                                   const uint8_t *__restrict__ rec_done,
                                   const int *__restrict__ rec_len, const float *__restrict__ adv,
                                   const float *__restrict__ p_old, const float *__restrict__ beta_dev, int n, int L, int B,
                                   int with_end_rows, float *__restrict__ out,
                                   double *__restrict__ kl_acc) {
  long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long LN = (long long)L * n;
  long long total = with_end_rows ? 2 * LN : LN;
  const float beta = beta_dev ? *beta_dev : 0.f;  // kl_ppo_learner::beta_ lives on the device (no host round trip per step)
  double kl = 0.0, cnt = 0.0;
  if (r < total) {
    bool is_end = r >= LN;
    long long k = is_end ? r - LN : r;
    int t = (int)(k / n), i = (int)(k % n);
    const float *p = probs + (size_t)r * B;
    float *o = out + (size_t)r * B;
    int a = rec_action[k];
    bool valid = !rec_len || t < rec_len[i];
    if (is_end)
      valid = rec_done[k] || t == L - 1;
    float A = (is_end || !valid) ? 0.f : adv[k];
    if (!valid) {
      for (int c = 0; c < B; ++c)
        o[c] = 0.f;
    } else if (kind == DFRL_LOSS_CLIPPED) {
      float g = clipped_grad(p[a], p_old[(size_t)k * B + a], A);
      for (int c = 0; c < B; ++c)
        o[c] = (c == a) ? g : 0.f;
    } else {
      const float *po = p_old + (size_t)k * B;
      for (int c = 0; c < B; ++c) {
        float v = p[c] * A - (c == a ? A : 0.f);
        if (kind == DFRL_LOSS_KL) {
          v += (p[c] - po[c]) * beta;
          kl += (double)(po[c] * logf(po[c] / p[c]));  // D_KL(p_old || p), policy_gradient.h:41-45
        }
        o[c] = v;
      }
      cnt = 1.0;
    }
  }
  if (kl_acc) {
    for (int o2 = 16; o2 > 0; o2 >>= 1) {
      kl += __shfl_down_sync(0xffffffffu, kl, o2);
      cnt += __shfl_down_sync(0xffffffffu, cnt, o2);
    }
    if ((threadIdx.x & 31) == 0 && cnt > 0.0) {
      atomicAdd(&kl_acc[0], kl);
      atomicAdd(&kl_acc[1], cnt);
    }
  }
}

// Adaptive beta of kl_regulated_loss (policy_gradient.h:68-83): acc = {sum of D_KL over the rows, row count} of ALL
// ranks; halved / doubled around d_targ, clamped to [1e-25, 0.1]. One thread.
__global__ void kl_beta_update_kernel(const double *__restrict__ acc, float *__restrict__ beta, float kl_target) {
  const float d_average = (float)(acc[0] / acc[1]);
  float b = *beta;
  if (fabsf(d_average) < kl_target / 1.5f)
    b /= 2;
  else if (fabsf(d_average) > kl_target * 1.5f)
    b *= 2;
  *beta = fminf(fmaxf(b, 1e-25f), 0.1f);
}

// square_loss_grad (nn.h:548-550) for the critic on start rows; rows past rec_len get 0.
__global__ void value_loss_kernel(const float *__restrict__ v, const float *__restrict__ tgt,
                                  long long rows, float *__restrict__ out) {
  long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows)
    out[r] = v[r] - tgt[r];
}

__global__ void reinforce_adv_kernel(const float *__restrict__ g, const int *__restrict__ len, int n,
                                     int L, float baseline, float *__restrict__ adv) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= (long long)n * L)
    return;
  int t = (int)(k / n), i = (int)(k % n);
  adv[k] = (t < len[i]) ? g[k] - baseline : 0.f;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
extern "C" void dfrl_trainer_config_default(dfrl_trainer_config *c) {
  if (!c)
    return;
  memset(c, 0, sizeof(*c));
  c->algo = DFRL_ALGO_PPO;
  c->work = 4;            // ppo_training.cc:31
  c->gamma = 0.99f;       // ppo_training.cc:46
  c->lambda = 0.95f;      // policy_gradient.h:286
  c->epochs = 4;          // policy_gradient.h:300
  c->kl_target = 1e-9f;   // policy_gradient.h:334
  c->kl_beta0 = 1.0f;     // policy_gradient.h:333
  c->policy_opt = DFRL_OPT_SGD;
  c->value_opt = DFRL_OPT_SGD;
  c->policy_lr = 1e-4f;   // ppo_training.cc:17
  c->value_lr = 1e-5f;    // ppo_training.cc:26
  c->adam_beta1 = 0.9f;   // nn.h:661
  c->adam_beta2 = 0.999f;
  c->action_mode = DFRL_ACT_SAMPLE;
  c->fused = 1;
}

// kl_ppo_learner::beta_ (policy_gradient.h:333) on the device: the float behind the four statistics counters.
static float *kl_beta_dev(dfrl_trainer *t) { return reinterpret_cast<float *>(t->counters + 4); }

static size_t opt_state_floats(int kind, int n) {
  return kind == DFRL_OPT_SGD ? 0 : kind == DFRL_OPT_MOMENTUM ? (size_t)n : (size_t)2 * n;
}

static int trainer_setup(dfrl_trainer *t, dfrl_ctx *ctx, const dfrl_trainer_config *cfg, dfrl_env *env,
                         dfrl_mlp *policy, dfrl_mlp *value);

extern "C" int dfrl_trainer_create(dfrl_ctx *ctx, const dfrl_trainer_config *cfg, dfrl_env *env,
                                   dfrl_mlp *policy, dfrl_mlp *value, dfrl_trainer **out) {
  DFRL_CHECK(ctx && cfg && env && policy && out, "null argument");
  DFRL_CHECK(cfg->algo >= 0 && cfg->algo <= 3, "unknown algo %d", cfg->algo);
  DFRL_CHECK(cfg->work > 0, "work must be positive");
  DFRL_CHECK(cfg->algo == DFRL_ALGO_REINFORCE || value, "value model required");
  DFRL_CHECK(policy->input_cols == 4 * env->B, "policy input width %d != 4 * bins", policy->input_cols);
  DFRL_CHECK(policy->output_cols == env->B, "policy output width %d != bins %d (action.cardinality)",
             policy->output_cols, env->B);
  if (value) {
    DFRL_CHECK(value->input_cols == 4 * env->B, "value input width mismatch");
    DFRL_CHECK(value->output_cols == 1, "value model must output one column");
  }
  dfrl_trainer *t = new dfrl_trainer();  // value-initialised: every member null / zero
  t->ctx = ctx;
  int rc = trainer_setup(t, ctx, cfg, env, policy, value);
  if (rc != DFRL_OK) {  // a failed allocation: release what exists (dfrl_last_error keeps the reason)
    dfrl_trainer_destroy(t);
    return rc;
  }
  *out = t;
  if (cfg->fused)
    dfrl_fused_try_attach(t);  // silently stays layered when the nets do not qualify
  return DFRL_OK;
}

static int trainer_setup(dfrl_trainer *t, dfrl_ctx *ctx, const dfrl_trainer_config *cfg, dfrl_env *env,
                         dfrl_mlp *policy, dfrl_mlp *value) {
  t->cfg = *cfg;
  t->env = env;
  t->policy = policy;
  t->value = cfg->algo == DFRL_ALGO_REINFORCE ? nullptr : value;
  t->n = env->n;
  t->B = env->B;
  t->P = env->P;
  t->stride = env->stride;
  t->O = 4 * env->B;
  if (cfg->algo == DFRL_ALGO_REINFORCE) {
    int me = env_max_episode_len(env->cfg);
    DFRL_CHECK(me > 0, "episodes never end with zero-sized items");
    t->L = cfg->work * me;
  } else {
    t->L = cfg->work;
  }
  const size_t LN = (size_t)t->L * t->n;
  const int Pp = policy->n_params, Pv = t->value ? t->value->n_params : 0;
  DFRL_CUDA(cudaMalloc(&t->rec_state, (size_t)t->L * t->P * t->stride));
  DFRL_CUDA(cudaMemsetAsync(t->rec_state, 0, (size_t)t->L * t->P * t->stride, ctx->stream));
  DFRL_CUDA(cudaMalloc(&t->rec_action, LN));
  DFRL_CUDA(cudaMemsetAsync(t->rec_action, 0, LN, ctx->stream));
  DFRL_CUDA(cudaMalloc(&t->rec_done, LN));
  DFRL_CUDA(cudaMemsetAsync(t->rec_done, 0, LN, ctx->stream));
  DFRL_CUDA(cudaMalloc(&t->rec_probs, sizeof(float) * LN * t->B));
  DFRL_CUDA(cudaMemsetAsync(t->rec_probs, 0, sizeof(float) * LN * t->B, ctx->stream));
  DFRL_CUDA(cudaMalloc(&t->rec_len, sizeof(int) * t->n));
  DFRL_CUDA(cudaMalloc(&t->ep_done, sizeof(int) * t->n));
  DFRL_CUDA(cudaMalloc(&t->obs, sizeof(float) * 2 * LN * t->O));
  DFRL_CUDA(cudaMalloc(&t->v_start, sizeof(float) * LN));
  DFRL_CUDA(cudaMalloc(&t->v_end, sizeof(float) * LN));
  DFRL_CUDA(cudaMalloc(&t->targets, sizeof(float) * LN));
  DFRL_CUDA(cudaMalloc(&t->adv, sizeof(float) * LN));
  DFRL_CUDA(cudaMemsetAsync(t->adv, 0, sizeof(float) * LN, ctx->stream));
  DFRL_CUDA(cudaMemsetAsync(t->targets, 0, sizeof(float) * LN, ctx->stream));
  DFRL_CUDA(cudaMalloc(&t->dyv, sizeof(float) * LN));
  DFRL_CUDA(cudaMalloc(&t->dprobs, sizeof(float) * 2 * LN * t->B));
  int epochs = (cfg->algo == DFRL_ALGO_PPO || cfg->algo == DFRL_ALGO_KL_PPO) ? cfg->epochs : 1;
  t->epochs = epochs;
  DFRL_CUDA(cudaMalloc(&t->pgrad_log, sizeof(float) * (size_t)epochs * Pp));
  DFRL_CUDA(cudaMemsetAsync(t->pgrad_log, 0, sizeof(float) * (size_t)epochs * Pp, ctx->stream));
  DFRL_CUDA(cudaMalloc(&t->vgrad, sizeof(float) * (Pv ? Pv : 1)));
  DFRL_CUDA(cudaMemsetAsync(t->vgrad, 0, sizeof(float) * (Pv ? Pv : 1), ctx->stream));
  size_t ps = opt_state_floats(cfg->policy_opt, Pp), vs = opt_state_floats(cfg->value_opt, Pv);
  DFRL_CUDA(cudaMalloc(&t->pstate, sizeof(float) * (ps ? ps : 1)));
  DFRL_CUDA(cudaMemsetAsync(t->pstate, 0, sizeof(float) * (ps ? ps : 1), ctx->stream));
  DFRL_CUDA(cudaMalloc(&t->vstate, sizeof(float) * (vs ? vs : 1)));
  DFRL_CUDA(cudaMemsetAsync(t->vstate, 0, sizeof(float) * (vs ? vs : 1), ctx->stream));
  DFRL_CUDA(cudaMalloc(&t->counters, sizeof(unsigned long long) * 8));
  DFRL_CUDA(cudaMemsetAsync(t->counters, 0, sizeof(unsigned long long) * 8, ctx->stream));
  DFRL_CUDA(cudaMemcpyAsync(kl_beta_dev(t), &cfg->kl_beta0, sizeof(float), cudaMemcpyHostToDevice, ctx->stream));  // (synchronised below)
  DFRL_CUDA(cudaMalloc(&t->acc, sizeof(double) * 4));
  {
    const float ones[2] = {1.f, 1.f};  // nn.h:693
    DFRL_CUDA(cudaMalloc(&t->adam_t_dev, sizeof(ones)));
    DFRL_CUDA(cudaMemcpyAsync(t->adam_t_dev, ones, sizeof(ones), cudaMemcpyHostToDevice, ctx->stream));
    DFRL_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  DFRL_CUDA(cudaMalloc(&t->tape_items, LN));
  DFRL_CUDA(cudaMalloc(&t->tape_items2, LN));
  {
    cudaStream_t cs;
    DFRL_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    t->copy_stream = cs;
    for (int b = 0; b < 2; ++b) {
      cudaEvent_t e1, e2;
      DFRL_CUDA(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
      DFRL_CUDA(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
      t->tape_copied[b] = e1;
      t->tape_free[b] = e2;
    }
    t->tape_flip = 0;
  }
  DFRL_CUDA(cudaMalloc(&t->tape_actions, LN));
  DFRL_CUDA(cudaMalloc(&t->tape_u, sizeof(double) * LN));
  DFRL_CUDA(cudaMallocHost(&t->pin, 64));
  DFRL_CUDA(cudaMallocHost(&t->stats_pin, 4 * 48));
  for (int i = 0; i < 4; ++i) {
    cudaEvent_t e;
    DFRL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    t->stats_event[i] = e;
  }
  t->stats_head = t->stats_tail = 0;
  t->p_adam_t = 1.f;  // nn.h:693
  t->v_adam_t = 1.f;
  t->kl_beta = cfg->kl_beta0;
  t->last_rollout_steps = 0;
  t->last_rollout_reward = 0;
  t->fused_impl = nullptr;
  t->obs_valid = false;
  t->critic_was_fused = false;
  t->graph_exec = nullptr;
  t->graph_launches = 0;
  t->plain_iterations = 0;
  t->graph_failed = 0;
  t->graph_key = 0;
  return DFRL_OK;
}

extern "C" int dfrl_trainer_destroy(dfrl_trainer *t) {
  if (!t)
    return DFRL_OK;
  cudaStreamSynchronize(t->ctx->stream);
  if (t->graph_exec)
    cudaGraphExecDestroy((cudaGraphExec_t)t->graph_exec);
  dfrl_fused_detach(t);
  void *ptrs[] = {t->rec_state, t->rec_action, t->rec_done, t->rec_probs, t->rec_len, t->ep_done,
                  t->obs, t->v_start, t->v_end, t->targets, t->adv, t->dyv, t->dprobs, t->pgrad_log,
                  t->vgrad, t->pstate, t->vstate, t->counters, t->acc, t->adam_t_dev, t->tape_items,
                  t->tape_actions, t->tape_u};
  for (void *p : ptrs)
    if (p)
      cudaFree(p);
  if (t->tape_items2)
    cudaFree(t->tape_items2);
  if (t->copy_stream) {
    cudaStreamSynchronize((cudaStream_t)t->copy_stream);
    cudaStreamDestroy((cudaStream_t)t->copy_stream);
  }
  for (int b = 0; b < 2; ++b) {
    if (t->tape_copied[b])
      cudaEventDestroy((cudaEvent_t)t->tape_copied[b]);
    if (t->tape_free[b])
      cudaEventDestroy((cudaEvent_t)t->tape_free[b]);
  }
  if (t->pin)
    cudaFreeHost(t->pin);
  if (t->stats_pin)
    cudaFreeHost(t->stats_pin);
  for (int i = 0; i < 4; ++i)
    if (t->stats_event[i])
      cudaEventDestroy((cudaEvent_t)t->stats_event[i]);
  delete t;
  return DFRL_OK;
}

// ------------------------------------------------------------------ rollout -----------------
static int rollout_step(dfrl_trainer *t, int slot, int mode, const uint8_t *forced_step,
                        const double *u_step, const uint8_t *item_step, int episodic, int ep_target,
                        bool record) {
  dfrl_ctx *ctx = t->ctx;
  dfrl_env *e = t->env;
  const size_t n = t->n;
  float *obs_t = t->obs + (size_t)slot * n * t->O;
  float *probs_t = t->rec_probs + (size_t)slot * n * t->B;
  if (record)
    DFRL_CUDA(cudaMemcpyAsync(t->rec_state + (size_t)slot * t->P * t->stride, e->state,
                              (size_t)t->P * t->stride, cudaMemcpyDeviceToDevice, ctx->stream));
  DFRL_TRY(dfrl_obs_encode(ctx, e->state, t->n, t->stride, t->B, e->cfg.cap_w, e->cfg.cap_h, obs_t));
  DFRL_TRY(dfrl_mlp_eval(t->policy, obs_t, t->n, probs_t));
  env_params p = make_params(e);
  DFRL_LAUNCH(ctx, act_step_kernel, ceil_div(t->n, 128), 128, 0, p, e->state, e->draws, e->steps,
              probs_t, mode, forced_step, u_step, item_step,
              record ? t->rec_action + (size_t)slot * n : nullptr,
              record ? t->rec_done + (size_t)slot * n : nullptr, episodic, ep_target, t->ep_done,
              record ? t->rec_len : nullptr, t->counters);
  return DFRL_OK;
}

// The four counters, and (h5 only) the device-side KL beta in the low half of h[4].
static int read_counters(dfrl_trainer *t, unsigned long long *h, int words = 4) {
  DFRL_CUDA(cudaMemcpyAsync(t->pin, t->counters, 8 * words, cudaMemcpyDeviceToHost, t->ctx->stream));
  DFRL_CUDA(cudaStreamSynchronize(t->ctx->stream));
  memcpy(h, t->pin, 8 * words);
  return DFRL_OK;
}

static int rollout_layered(dfrl_trainer *t, const uint8_t *items_dev, const uint8_t *actions_dev,
                           const double *u_dev) {
  dfrl_ctx *ctx = t->ctx;
  const size_t n = t->n;
  int mode = t->cfg.action_mode;
  if (t->cfg.algo == DFRL_ALGO_REINFORCE) {
    // agent::play_one_episode x work for every env (pg_training.cc:51-55)
    DFRL_CUDA(cudaMemsetAsync(t->ep_done, 0, sizeof(int) * n, ctx->stream));
    DFRL_CUDA(cudaMemsetAsync(t->rec_len, 0, sizeof(int) * n, ctx->stream));
    for (int s = 0; s < t->L; ++s) {
      DFRL_CUDA(cudaMemsetAsync(t->counters + 3, 0, sizeof(unsigned long long), ctx->stream));
      DFRL_TRY(rollout_step(t, s, mode, actions_dev ? actions_dev + s * n : nullptr,
                            u_dev ? u_dev + s * n : nullptr, items_dev ? items_dev + s * n : nullptr,
                            1, t->cfg.work, true));
      if ((s & 3) == 3 || s == t->L - 1) {
        unsigned long long h[4];
        DFRL_TRY(read_counters(t, h));
        if (h[3] == 0)
          break;
      }
    }
  } else {
    for (int s = 0; s < t->L; ++s)
      DFRL_TRY(rollout_step(t, s, mode, actions_dev ? actions_dev + s * n : nullptr,
                            u_dev ? u_dev + s * n : nullptr, items_dev ? items_dev + s * n : nullptr,
                            0, 0, true));
  }
  t->obs_valid = true;
  return DFRL_OK;
}

// Start-row observations for the layered kernels when the rollout was fused (it keeps no fp32
// observations): observation::to_vector of every recorded start state.
static int ensure_obs(dfrl_trainer *t) {
  if (t->obs_valid)
    return DFRL_OK;
  dfrl_env *e = t->env;
  for (int s = 0; s < t->L; ++s)
    DFRL_TRY(dfrl_obs_encode(t->ctx, t->rec_state + (size_t)s * t->P * t->stride, t->n, t->stride, t->B,
                             e->cfg.cap_w, e->cfg.cap_h, t->obs + (size_t)s * t->n * t->O));
  t->obs_valid = true;
  return DFRL_OK;
}

extern "C" int dfrl_trainer_rollout(dfrl_trainer *t, const uint8_t *items_host,
                                    const uint8_t *actions_host, const double *u_host) {
  DFRL_CHECK(t, "null trainer");
  dfrl_ctx *ctx = t->ctx;
  const size_t LN = (size_t)t->L * t->n;
  DFRL_CHECK(t->cfg.action_mode != DFRL_ACT_FORCED || actions_host, "forced mode needs an action tape");
  const uint8_t *items_dev = nullptr, *actions_dev = nullptr;
  const double *u_dev = nullptr;
  int tape_buf = -1;
  if (items_host) {
    // copy stream: waits until the rollout that last read this buffer is done, copies, and the
    // compute stream waits for the copy only -- the copy itself runs beside the previous step
    tape_buf = t->tape_flip;
    t->tape_flip ^= 1;
    uint8_t *dst = tape_buf ? t->tape_items2 : t->tape_items;
    cudaStream_t cs = (cudaStream_t)t->copy_stream;
    DFRL_CUDA(cudaStreamWaitEvent(cs, (cudaEvent_t)t->tape_free[tape_buf], 0));
    DFRL_CUDA(cudaMemcpyAsync(dst, items_host, LN, cudaMemcpyHostToDevice, cs));
    DFRL_CUDA(cudaEventRecord((cudaEvent_t)t->tape_copied[tape_buf], cs));
    DFRL_CUDA(cudaStreamWaitEvent(ctx->stream, (cudaEvent_t)t->tape_copied[tape_buf], 0));
    items_dev = dst;
  }
  if (actions_host) {
    DFRL_CUDA(cudaMemcpyAsync(t->tape_actions, actions_host, LN, cudaMemcpyHostToDevice, ctx->stream));
    actions_dev = t->tape_actions;
  }
  if (u_host) {
    DFRL_CUDA(cudaMemcpyAsync(t->tape_u, u_host, sizeof(double) * LN, cudaMemcpyHostToDevice, ctx->stream));
    u_dev = t->tape_u;
  }
  int rc = t->fused_impl ? dfrl_fused_rollout(t, items_dev, actions_dev, u_dev) : DFRL_ERR_UNSUPPORTED;
  if (rc == DFRL_ERR_UNSUPPORTED)
    rc = rollout_layered(t, items_dev, actions_dev, u_dev);
  DFRL_TRY(rc);
  if (tape_buf >= 0)
    DFRL_CUDA(cudaEventRecord((cudaEvent_t)t->tape_free[tape_buf], ctx->stream));
  return DFRL_OK;
}

// ------------------------------------------------------------------ learn -------------------
static int apply_opt(dfrl_trainer *t, dfrl_mlp *m, int kind, float *grad, float *state, float lr,
                     float wd, float *adam_t) {
  DFRL_TRY(dfrl_allreduce_sum(t->ctx, grad, (size_t)m->n_params));  // K8: SUM over ranks
  DFRL_TRY(dfrl_opt_step(t->ctx, kind, m->params, grad, state, m->n_params, lr, wd,
                         t->cfg.adam_beta1, t->cfg.adam_beta2, *adam_t));
  if (kind == DFRL_OPT_ADAM)
    *adam_t += 1.f;  // nn.h:686
  dfrl_mlp_params_changed(m);
  return DFRL_OK;
}

// The optimizer update rides in the fused gradient-reduction kernel (single rank) or in the
// peer-memory exchange kernel (several ranks with attached peers); otherwise NCCL + opt kernel.
static bool fuse_opt(dfrl_trainer *t, dfrl_mlp *m, int kind, float *state, float lr, float wd, float adam_t,
                     dfrl_opt_spec *spec) {
  if (t->ctx->nranks != 1 && !t->ctx->p2p.attached)
    return false;
  // (the device-side counter mirrors the host's: both start at 1 and advance once per update)
  spec->t_dev = kind == DFRL_OPT_ADAM ? t->adam_t_dev + (m == t->policy ? 0 : 1) : nullptr;
  spec->kind = kind;
  spec->params = m->params;
  spec->state = state;
  spec->lr = lr;
  spec->wd = wd;
  spec->beta1 = t->cfg.adam_beta1;
  spec->beta2 = t->cfg.adam_beta2;
  spec->c1 = 1.f - powf(t->cfg.adam_beta1, adam_t);  // nn.h:683-684
  spec->c2 = 1.f - powf(t->cfg.adam_beta2, adam_t);
  return true;
}
static void after_fused_opt(dfrl_mlp *, int kind, float *adam_t) {
  if (kind == DFRL_OPT_ADAM)
    *adam_t += 1.f;  // nn.h:686 (fused.cu already marked the parameters as changed)
}

// phases: DFRL_PHASE_* mask. The whole learn() is VALUE | ADVANTAGE | POLICY; the host mirror runs
// VALUE | ADVANTAGE, hands the rows to a user's optimize_action override, or runs POLICY for the default.
static int learn_layered(dfrl_trainer *t, int phases = DFRL_PHASE_ALL) {
  dfrl_ctx *ctx = t->ctx;
  dfrl_env *e = t->env;
  const int n = t->n, L = t->L, B = t->B;
  const long long LN = (long long)L * n;
  float *obs_start = t->obs, *obs_end = t->obs + (size_t)LN * t->O;
  const dfrl_trainer_config &c = t->cfg;

  if (c.algo == DFRL_ALGO_REINFORCE) {
    DFRL_CHECK(phases == DFRL_PHASE_ALL, "REINFORCE has a single phase");
    // policy_gradient_learner::learn (policy_gradient.h:95-123)
    float *g = t->targets;
    DFRL_CUDA(cudaMemsetAsync(t->acc, 0, sizeof(double) * 2, ctx->stream));
    DFRL_TRY(dfrl_returns(ctx, t->rec_done, t->rec_len, n, L, c.gamma, g, t->acc));
    DFRL_TRY(dfrl_allreduce_sum_f64(ctx, t->acc, 2));
    double h[2];
    DFRL_CUDA(cudaMemcpyAsync(t->pin, t->acc, 16, cudaMemcpyDeviceToHost, ctx->stream));
    DFRL_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(h, t->pin, 16);
    float baseline = (float)h[0] / (float)h[1];  // total_reward / experience.size() (145)
    DFRL_LAUNCH(ctx, reinforce_adv_kernel, ceil_div(LN, 256), 256, 0, g, t->rec_len, n, L, baseline, t->adv);
    float *probs = nullptr;
    DFRL_TRY(ensure_obs(t));
    DFRL_TRY(dfrl_mlp_forward_keep(t->policy, obs_start, (int)LN, &probs));
    DFRL_LAUNCH(ctx, policy_loss_kernel, ceil_div(LN, 128), 128, 0, DFRL_LOSS_SOFTMAX_LOG, probs,
                t->rec_action, t->rec_done, t->rec_len, t->adv, t->rec_probs, (const float *)nullptr, n, L, B, 0,
                t->dprobs, (double *)nullptr);
    DFRL_TRY(dfrl_mlp_backward(t->policy, t->dprobs, t->pgrad_log));
    DFRL_TRY(apply_opt(t, t->policy, c.policy_opt, t->pgrad_log, t->pstate, c.policy_lr, c.policy_wd, &t->p_adam_t));
    return DFRL_OK;
  }

  // actor_critic_learner::learn (policy_gradient.h:159-185)
  const int va = phases & (DFRL_PHASE_VALUE | DFRL_PHASE_ADVANTAGE);
  DFRL_CHECK(va == 0 || va == (DFRL_PHASE_VALUE | DFRL_PHASE_ADVANTAGE), "VALUE and ADVANTAGE run together");
  int frc = t->critic_was_fused ? DFRL_OK : DFRL_ERR_UNSUPPORTED;
  if (va) {
  dfrl_opt_spec vspec;
  const bool vfo = fuse_opt(t, t->value, c.value_opt, t->vstate, c.value_lr, c.value_wd, t->v_adam_t, &vspec);
  frc = dfrl_fused_critic_gradient(t, t->vgrad, vfo ? &vspec : nullptr);  // tcgen05 fused critic step
  t->critic_was_fused = frc == DFRL_OK;
  if (frc == DFRL_OK) {
    if (vfo)
      after_fused_opt(t->value, c.value_opt, &t->v_adam_t);
    else
      DFRL_TRY(apply_opt(t, t->value, c.value_opt, t->vgrad, t->vstate, c.value_lr, c.value_wd, &t->v_adam_t));
    DFRL_TRY(dfrl_fused_gae(t));
  } else if (frc != DFRL_ERR_UNSUPPORTED) {
    return frc;
  } else {
    DFRL_TRY(ensure_obs(t));
    DFRL_LAUNCH(ctx, end_obs_kernel, ceil_div(LN * B, 256), 256, 0, t->rec_state, e->state,
                t->rec_action, t->rec_done, n, t->stride, B, L, (float)e->cfg.cap_w, (float)e->cfg.cap_h,
                reinterpret_cast<float4 *>(obs_end));
    // update_value_model (196-218): V on every row with the current critic, targets, one step
    DFRL_TRY(dfrl_mlp_eval(t->value, obs_end, (int)LN, t->v_end));
    float *v_now = nullptr;
    DFRL_TRY(dfrl_mlp_forward_keep(t->value, obs_start, (int)LN, &v_now));
    DFRL_TRY(dfrl_gae(ctx, t->rec_done, v_now, t->v_end, n, L, c.gamma, c.lambda, t->targets, nullptr));
    DFRL_LAUNCH(ctx, value_loss_kernel, ceil_div(LN, 256), 256, 0, v_now, t->targets, LN, t->dyv);
    DFRL_TRY(dfrl_mlp_backward(t->value, t->dyv, t->vgrad));
    DFRL_TRY(apply_opt(t, t->value, c.value_opt, t->vgrad, t->vstate, c.value_lr, c.value_wd, &t->v_adam_t));
    // calculate_advantage (220-281) with the UPDATED critic
    DFRL_TRY(dfrl_mlp_eval(t->value, obs_end, (int)LN, t->v_end));
    DFRL_TRY(dfrl_mlp_eval(t->value, obs_start, (int)LN, t->v_start));
    DFRL_TRY(dfrl_gae(ctx, t->rec_done, t->v_start, t->v_end, n, L, c.gamma, c.lambda, nullptr, t->adv));
  }
  }
  if (!(phases & DFRL_PHASE_POLICY))
    return DFRL_OK;
  // optimize_action (187-194 / 297-307 / 318-330)
  int kind = c.algo == DFRL_ALGO_ACTOR_CRITIC ? DFRL_LOSS_SOFTMAX_LOG
             : c.algo == DFRL_ALGO_PPO        ? DFRL_LOSS_CLIPPED
                                              : DFRL_LOSS_KL;
  const int with_end = kind == DFRL_LOSS_KL ? 1 : 0;
  const long long rows = with_end ? 2 * LN : LN;
  for (int ep = 0; ep < t->epochs; ++ep) {
    {
      // tcgen05 fused forward + loss + backward (fused.cu); layered kernels otherwise
      float *g = t->pgrad_log + (size_t)ep * t->policy->n_params;
      dfrl_opt_spec pspec;
      const bool pfo = fuse_opt(t, t->policy, c.policy_opt, t->pstate, c.policy_lr, c.policy_wd, t->p_adam_t, &pspec);
      int rc = dfrl_fused_policy_gradient(t, kind, g, pfo ? &pspec : nullptr);
      if (rc == DFRL_OK) {
        if (pfo)
          after_fused_opt(t->policy, c.policy_opt, &t->p_adam_t);
        else
          DFRL_TRY(apply_opt(t, t->policy, c.policy_opt, g, t->pstate, c.policy_lr, c.policy_wd, &t->p_adam_t));
        continue;
      }
      if (rc != DFRL_ERR_UNSUPPORTED)
        return rc;
    }
    DFRL_TRY(ensure_obs(t));
    if (with_end && ep == 0 && frc == DFRL_OK)  // KL rows include the end rows: encode them once
      DFRL_LAUNCH(ctx, end_obs_kernel, ceil_div(LN * B, 256), 256, 0, t->rec_state, e->state,
                  t->rec_action, t->rec_done, n, t->stride, B, L, (float)e->cfg.cap_w, (float)e->cfg.cap_h,
                  reinterpret_cast<float4 *>(obs_end));
    float *probs = nullptr;
    DFRL_TRY(dfrl_mlp_forward_keep(t->policy, t->obs, (int)rows, &probs));
    if (with_end)
      DFRL_CUDA(cudaMemsetAsync(t->acc, 0, sizeof(double) * 2, ctx->stream));
    DFRL_LAUNCH(ctx, policy_loss_kernel, ceil_div(rows, 128), 128, 0, kind, probs, t->rec_action,
                t->rec_done, (const int *)nullptr, t->adv, t->rec_probs,
                with_end ? (const float *)kl_beta_dev(t) : (const float *)nullptr, n, L, B, with_end,
                t->dprobs, with_end ? t->acc : (double *)nullptr);
    float *grad = t->pgrad_log + (size_t)ep * t->policy->n_params;
    DFRL_TRY(dfrl_mlp_backward(t->policy, t->dprobs, grad));
    if (with_end) {
      // adaptive beta (policy_gradient.h:68-83): mean KL over ALL rows of all ranks, updated on the device (the
      // next step's loss kernel reads it there: no host round trip inside the learn phase)
      DFRL_TRY(dfrl_allreduce_sum_f64(ctx, t->acc, 2));
      DFRL_LAUNCH(ctx, kl_beta_update_kernel, 1, 1, 0, (const double *)t->acc, kl_beta_dev(t), c.kl_target);
    }
    DFRL_TRY(apply_opt(t, t->policy, c.policy_opt, grad, t->pstate, c.policy_lr, c.policy_wd, &t->p_adam_t));
  }
  return DFRL_OK;
}


// The learn phase of the fully fused path is a fixed sequence of kernel launches with constant
// arguments (no host round trip, no memset / memcpy): it is captured once as a CUDA graph and
// replayed, which removes the per-launch CPU and scheduling gaps that dominate at small
// environment counts (C2: 11 launches of ~10 us each). Several ranks qualify when the gradient
// exchange runs over attached peer memory (its exchange counter lives on the device). Adam qualifies:
// its step counter t lives on the device and the kernels derive the bias corrections 1 - beta^t from
// it. Not eligible: NCCL exchanges, per-kernel profiling, learners with host decisions inside the
// phase (REINFORCE, KL-PPO).
// KL-PPO: critic step / GAE on the fused kernels, the k policy steps on the layered kernels with beta adapted on
// the device -- no host decision inside the phase either, so it replays as a graph too (one rank: its
// exchanges are NCCL calls; SGD / momentum: the layered optimizer kernel takes Adam's step count by value).
static bool kl_graph_eligible(const dfrl_trainer *t) {
  return t->cfg.algo == DFRL_ALGO_KL_PPO && dfrl_fused_covers_critic(t) && t->ctx->nranks == 1 &&
         t->cfg.policy_opt != DFRL_OPT_ADAM && t->cfg.value_opt != DFRL_OPT_ADAM && !t->graph_failed;
}
static bool graph_eligible(const dfrl_trainer *t) {
  if (t->ctx->profiling)
    return false;
  return (dfrl_fused_covers_iteration(t) && (t->ctx->nranks == 1 || t->ctx->p2p.attached)) || kl_graph_eligible(t);
}

static int learn_graphed(dfrl_trainer *t) {
  dfrl_ctx *ctx = t->ctx;
  if (!graph_eligible(t))
    return learn_layered(t);
  if (t->plain_iterations < 1) {  // the first learn runs launch by launch (shared-memory attributes)
    ++t->plain_iterations;
    return learn_layered(t);
  }
  const int key = t->fused_impl ? dfrl_fused_learn_key(t) : 0;
  if (t->graph_exec && key != t->graph_key) {  // captured under other host-side decisions: capture again
    DFRL_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaGraphExecDestroy((cudaGraphExec_t)t->graph_exec);
    t->graph_exec = nullptr;
    t->graph_launches = 0;
  }
  if (!t->graph_exec) {
    cudaGraph_t graph = nullptr;
    const long long before = ctx->launches;
    DFRL_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    int rc = learn_layered(t);
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
    t->graph_launches = ctx->launches - before;
    ctx->launches = before;  // nothing ran yet
    if (rc != DFRL_OK || e != cudaSuccess || !graph) {
      if (graph)
        cudaGraphDestroy(graph);
      cudaGetLastError();
      if (kl_graph_eligible(t)) {  // a layered kernel's workspace grew during the capture: launch by launch from now on
        t->graph_failed = 1;         // (nothing ran; the SGD / momentum layered path keeps no host-side counters:
        t->obs_valid = false;        //  only the caches that the aborted pass marked as refreshed are reset)
        dfrl_mlp_params_changed(t->policy);
        if (t->value)
          dfrl_mlp_params_changed(t->value);
        return learn_layered(t);
      }
      if (rc == DFRL_OK)
        dfrl_set_error("graph capture of the learn phase failed: %s", cudaGetErrorString(e));
      return rc != DFRL_OK ? rc : DFRL_ERR_CUDA;
    }
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
      dfrl_set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
      return DFRL_ERR_CUDA;
    }
    t->graph_exec = exec;
    t->graph_key = key;
  }
  DFRL_CUDA(cudaGraphLaunch((cudaGraphExec_t)t->graph_exec, ctx->stream));
  ctx->launches += t->graph_launches;
  // the replay changed the parameters of both nets on the device: the host-side bookkeeping that
  // apply_opt / launch_reduce do launch by launch (transposed-weight cache of the layered kernels)
  // has to be repeated here, otherwise dfrl_mlp_eval / dfrl_eval_argmax keep using stale weights
  dfrl_mlp_params_changed(t->policy);
  if (t->value)
    dfrl_mlp_params_changed(t->value);
  if (t->cfg.policy_opt == DFRL_OPT_ADAM)  // host mirrors of the device-side adam step counters
    t->p_adam_t += (float)t->epochs;
  if (t->cfg.value_opt == DFRL_OPT_ADAM)
    t->v_adam_t += 1.f;
  return DFRL_OK;
}

extern "C" int dfrl_trainer_learn(dfrl_trainer *t) {
  DFRL_CHECK(t, "null trainer");
  return learn_graphed(t);
}

// optimizer::set_rate (nn.h:592) on a live learner. The rates are arguments of the captured kernels.
extern "C" int dfrl_trainer_set_rates(dfrl_trainer *t, float policy_lr, float policy_wd, float value_lr, float value_wd) {
  DFRL_CHECK(t, "null trainer");
  dfrl_trainer_config &c = t->cfg;
  if (c.policy_lr == policy_lr && c.policy_wd == policy_wd && c.value_lr == value_lr && c.value_wd == value_wd)
    return DFRL_OK;
  c.policy_lr = policy_lr, c.policy_wd = policy_wd, c.value_lr = value_lr, c.value_wd = value_wd;
  if (t->graph_exec) {
    DFRL_CUDA(cudaStreamSynchronize(t->ctx->stream));
    cudaGraphExecDestroy((cudaGraphExec_t)t->graph_exec);
    t->graph_exec = nullptr;
    t->graph_launches = 0;
  }
  return DFRL_OK;
}

extern "C" int dfrl_trainer_learn_phases(dfrl_trainer *t, int phases) {
  DFRL_CHECK(t, "null trainer");
  DFRL_CHECK(phases > 0 && phases <= DFRL_PHASE_ALL, "bad phase mask %d", phases);
  if (phases == DFRL_PHASE_ALL)
    return learn_graphed(t);
  return learn_layered(t, phases);
}

extern "C" int dfrl_trainer_iterate(dfrl_trainer *t, int iters) {
  DFRL_CHECK(t, "null trainer");
  DFRL_CHECK(t->cfg.action_mode != DFRL_ACT_FORCED, "iterate() cannot teacher-force");
  for (int it = 0; it < iters; ++it) {
    int rc = t->fused_impl ? dfrl_fused_rollout(t, nullptr, nullptr, nullptr) : DFRL_ERR_UNSUPPORTED;
    if (rc == DFRL_ERR_UNSUPPORTED)
      rc = rollout_layered(t, nullptr, nullptr, nullptr);
    DFRL_TRY(rc);
    DFRL_TRY(learn_graphed(t));
  }
  return DFRL_OK;
}

// ------------------------------------------------------------------ introspection -----------
extern "C" int dfrl_trainer_field_size(dfrl_trainer *t, int field, size_t *bytes) {
  DFRL_CHECK(t && bytes, "null argument");
  const size_t LN = (size_t)t->L * t->n;
  switch (field) {
  case DFRL_F_REC_STATE: *bytes = (size_t)t->L * t->P * t->n; break;
  case DFRL_F_REC_ACTION:
  case DFRL_F_REC_DONE: *bytes = LN; break;
  case DFRL_F_REC_PROBS: *bytes = sizeof(float) * LN * t->B; break;
  case DFRL_F_REC_LEN: *bytes = sizeof(int) * t->n; break;
  case DFRL_F_ADVANTAGE:
  case DFRL_F_VALUE_TARGET: *bytes = sizeof(float) * LN; break;
  case DFRL_F_POLICY_GRAD: *bytes = sizeof(float) * t->policy->n_params; break;
  case DFRL_F_VALUE_GRAD: *bytes = sizeof(float) * (t->value ? t->value->n_params : 0); break;
  case DFRL_F_POLICY_GRAD_LOG: *bytes = sizeof(float) * (size_t)t->epochs * t->policy->n_params; break;
  case DFRL_F_OBS_START: *bytes = sizeof(float) * LN * t->O; break;
  default:
    dfrl_set_error("unknown field %d", field);
    return DFRL_ERR_INVALID;
  }
  return DFRL_OK;
}

extern "C" int dfrl_trainer_read(dfrl_trainer *t, int field, void *dst_host, size_t bytes) {
  DFRL_CHECK(t && dst_host, "null argument");
  size_t want = 0;
  DFRL_TRY(dfrl_trainer_field_size(t, field, &want));
  DFRL_CHECK(bytes == want, "field %d is %zu bytes, caller passed %zu", field, want, bytes);
  cudaStream_t s = t->ctx->stream;
  const void *src = nullptr;
  switch (field) {
  case DFRL_F_REC_STATE:
    DFRL_CUDA(cudaMemcpy2DAsync(dst_host, t->n, t->rec_state, t->stride, t->n, (size_t)t->L * t->P,
                                cudaMemcpyDeviceToHost, s));
    DFRL_CUDA(cudaStreamSynchronize(s));
    return DFRL_OK;
  case DFRL_F_REC_ACTION: src = t->rec_action; break;
  case DFRL_F_REC_DONE: src = t->rec_done; break;
  case DFRL_F_REC_PROBS: src = t->rec_probs; break;
  case DFRL_F_REC_LEN: src = t->rec_len; break;
  case DFRL_F_ADVANTAGE: src = t->adv; break;
  case DFRL_F_VALUE_TARGET: src = t->targets; break;
  case DFRL_F_POLICY_GRAD: src = t->pgrad_log + (size_t)(t->epochs - 1) * t->policy->n_params; break;
  case DFRL_F_VALUE_GRAD: src = t->vgrad; break;
  case DFRL_F_POLICY_GRAD_LOG: src = t->pgrad_log; break;
  case DFRL_F_OBS_START:  // observation::to_vector of every recorded start state (encoded on demand)
    DFRL_TRY(ensure_obs(t));
    src = t->obs;
    break;
  }
  if (bytes) {
    DFRL_CUDA(cudaMemcpyAsync(dst_host, src, bytes, cudaMemcpyDeviceToHost, s));
    DFRL_CUDA(cudaStreamSynchronize(s));
  }
  return DFRL_OK;
}

static void fill_stats(dfrl_trainer *t, const unsigned long long *h, dfrl_trainer_stats *out);

// Asynchronous form of dfrl_trainer_get_stats: _begin enqueues the device->host copy of the
// counters as they stand after the work submitted so far and returns at once; _end waits for the
// OLDEST begun read only. A training loop that calls _begin(step i), submits step i + 1 and then
// calls _end keeps the device busy while the host prepares the next step. At most 4 reads in flight.
extern "C" int dfrl_trainer_stats_begin(dfrl_trainer *t) {
  DFRL_CHECK(t, "null trainer");
  DFRL_CHECK(t->stats_head - t->stats_tail < 4, "4 statistics reads already in flight");
  const unsigned slot = t->stats_head & 3u;
  DFRL_CUDA(cudaMemcpyAsync((char *)t->stats_pin + 48 * slot, t->counters, 40, cudaMemcpyDeviceToHost, t->ctx->stream));
  DFRL_CUDA(cudaEventRecord((cudaEvent_t)t->stats_event[slot], t->ctx->stream));
  t->stats_head++;
  return DFRL_OK;
}
extern "C" int dfrl_trainer_stats_end(dfrl_trainer *t, dfrl_trainer_stats *out) {
  DFRL_CHECK(t && out, "null argument");
  DFRL_CHECK(t->stats_head != t->stats_tail, "no statistics read in flight");
  const unsigned slot = t->stats_tail & 3u;
  DFRL_CUDA(cudaEventSynchronize((cudaEvent_t)t->stats_event[slot]));
  unsigned long long h[5];
  memcpy(h, (char *)t->stats_pin + 48 * slot, 40);
  t->stats_tail++;
  fill_stats(t, h, out);
  return DFRL_OK;
}

extern "C" int dfrl_trainer_get_stats(dfrl_trainer *t, dfrl_trainer_stats *out) {
  DFRL_CHECK(t && out, "null argument");
  unsigned long long h[5];
  DFRL_TRY(read_counters(t, h, 5));
  fill_stats(t, h, out);
  return DFRL_OK;
}

static void fill_stats(dfrl_trainer *t, const unsigned long long *h, dfrl_trainer_stats *out) {
  out->env_steps = (long long)h[0];
  out->episodes = (long long)h[1];
  out->reward_sum = (double)h[2];
  // mean reward per step since the previous get_stats() call
  long long ds = (long long)h[0] - t->last_rollout_steps, dr = (long long)h[2] - t->last_rollout_reward;
  out->last_mean_reward = ds > 0 ? (double)dr / (double)ds : 0.0;
  t->last_rollout_steps = (long long)h[0];
  t->last_rollout_reward = (long long)h[2];
  memcpy(&t->kl_beta, &h[4], sizeof(float));  // host mirror of the device-side beta, as of this read
  out->kl_beta = t->kl_beta;
}

// deep_agent.cc:28-41: argmax policy, `episodes` episodes per env, mean reward per episode.
extern "C" int dfrl_eval_argmax(dfrl_ctx *ctx, dfrl_env *env, dfrl_mlp *policy, int episodes,
                                double *mean_reward, long long *env_steps) {
  DFRL_CHECK(ctx && env && policy && mean_reward, "null argument");
  DFRL_CHECK(episodes > 0, "episodes must be positive");
  DFRL_CHECK(policy->input_cols == 4 * env->B && policy->output_cols == env->B,
             "policy shape does not match the env");
  const int n = env->n, B = env->B;
  int max_ep = env_max_episode_len(env->cfg);
  DFRL_CHECK(max_ep > 0, "episodes never end with zero-sized items");
  float *obs, *probs;
  int *ep_done;
  unsigned long long *counters;
  DFRL_CUDA(cudaMalloc(&obs, sizeof(float) * (size_t)n * 4 * B));
  DFRL_CUDA(cudaMalloc(&probs, sizeof(float) * (size_t)n * B));
  DFRL_CUDA(cudaMalloc(&ep_done, sizeof(int) * n));
  DFRL_CUDA(cudaMalloc(&counters, 32));
  DFRL_CUDA(cudaMemsetAsync(ep_done, 0, sizeof(int) * n, ctx->stream));
  DFRL_CUDA(cudaMemsetAsync(counters, 0, 32, ctx->stream));
  env_params p = make_params(env);
  unsigned long long h[4] = {0, 0, 0, 0};
  int rc = DFRL_OK;
  for (int s = 0; s < episodes * max_ep && rc == DFRL_OK; ++s) {
    cudaMemsetAsync(counters + 3, 0, 8, ctx->stream);
    rc = dfrl_obs_encode(ctx, env->state, n, env->stride, B, env->cfg.cap_w, env->cfg.cap_h, obs);
    if (rc == DFRL_OK)
      rc = dfrl_mlp_eval(policy, obs, n, probs);
    if (rc != DFRL_OK)
      break;
    act_step_kernel<<<ceil_div(n, 128), 128, 0, ctx->stream>>>(
        p, env->state, env->draws, env->steps, probs, DFRL_ACT_ARGMAX, nullptr, nullptr, nullptr,
        nullptr, nullptr, 1, episodes, ep_done, nullptr, counters);
    ctx->launches++;
    if ((s & 7) == 7) {
      cudaMemcpyAsync(h, counters, 32, cudaMemcpyDeviceToHost, ctx->stream);
      cudaStreamSynchronize(ctx->stream);
      if (h[3] == 0)
        break;
    }
  }
  cudaMemcpyAsync(h, counters, 32, cudaMemcpyDeviceToHost, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  cudaFree(obs);
  cudaFree(probs);
  cudaFree(ep_done);
  cudaFree(counters);
  if (rc != DFRL_OK)
    return rc;
  if (e != cudaSuccess) {
    dfrl_set_error("eval_argmax: %s", cudaGetErrorString(e));
    return DFRL_ERR_CUDA;
  }
  *mean_reward = (double)h[2] / ((double)n * episodes);
  if (env_steps)
    *env_steps = (long long)h[0];
  return DFRL_OK;
}
