// trainer.h -- internal layout of dfrl_trainer, shared by trainer.cu (layered path) and
// fused.cu (fused small-MLP path).
#pragma once

#include "common.cuh"

struct dfrl_trainer {
  dfrl_ctx *ctx;
  dfrl_trainer_config cfg;
  dfrl_env *env;
  dfrl_mlp *policy, *value;
  int n, B, P, stride, O;
  int L;       // recorded steps per env per iteration (work, or work * max episode length)
  int epochs;  // policy optimizer steps per learn()
  // rollout records (see trainer.cu header)
  int8_t *rec_state;
  uint8_t *rec_action, *rec_done;
  float *rec_probs;
  int *rec_len, *ep_done;
  // learn workspace
  float *obs;  // [2][L*n][O]: start rows, then end rows (layered kernels only)
  bool obs_valid;  // start rows of `obs` hold the observations of the current records
  bool critic_was_fused;  // the last VALUE | ADVANTAGE phase ran on the fused kernels (no end-row observations encoded)
  float *v_start, *v_end, *targets, *adv, *dyv, *dprobs;
  float *pgrad_log, *vgrad;
  float *pstate, *vstate;
  float p_adam_t, v_adam_t, kl_beta;
  float *adam_t_dev;  // [2] device-side adam step counters of the fused path (policy, value)
  // counters: [0] env steps, [1] episodes, [2] reward sum, [3] active envs (episodic rollouts)
  unsigned long long *counters;
  double *acc;  // [4] device accumulators (baseline, KL)
  // device copies of the host tapes
  uint8_t *tape_items, *tape_actions;
  // item tapes go host->device on a copy stream into two alternating device buffers, so that the
  // copy for step i + 1 overlaps step i when the caller runs ahead (dfrl_trainer_stats_begin/_end)
  uint8_t *tape_items2;       // second buffer (first = tape_items)
  void *copy_stream;          // cudaStream_t
  void *tape_copied[2];       // cudaEvent_t: copy into buffer b finished
  void *tape_free[2];         // cudaEvent_t: the rollout that read buffer b has been submitted / finished
  int tape_flip;
  double *tape_u;
  void *pin;  // 64 B pinned host staging
  // asynchronous statistics reads (dfrl_trainer_stats_begin / _end): ring of pinned slots + events
  void *stats_pin;            // [DFRL_STATS_RING][32 B]
  void *stats_event[4];       // cudaEvent_t
  unsigned stats_head, stats_tail;
  long long last_rollout_steps, last_rollout_reward;
  void *fused_impl;  // non-null when the fused kernels drive this trainer
  // one free-running iteration captured as a CUDA graph (dfrl_trainer_iterate), see trainer.cu
  void *graph_exec;        // cudaGraphExec_t
  long long graph_launches;  // kernels per captured iteration
  int plain_iterations;    // iterations run launch by launch so far
  int graph_key;           // dfrl_fused_learn_key() at capture time: host decisions frozen into the graph
  int graph_failed;        // a capture attempt failed (KL-PPO's layered policy steps): no further attempts
};

// fused.cu: true when rollout, critic step, GAE and policy steps all run on the fused kernels
bool dfrl_fused_covers_iteration(const dfrl_trainer *t);

// Optimizer update fused behind the gradient reduction (single rank). params == null: none.
struct dfrl_opt_spec {
  int kind;
  float *params, *state;
  float lr, wd, beta1, beta2, c1, c2;
  float *t_dev;  // adam: device-side step counter (float, starts at 1: nn.h:693); the kernel derives the
                 // bias corrections from it and advances it, so the launch arguments stay constant
                 // (CUDA-graph replay). null: c1 / c2 above are used.
};

// fused.cu
int dfrl_fused_try_attach(dfrl_trainer *t);
void dfrl_fused_detach(dfrl_trainer *t);
int dfrl_fused_rollout(dfrl_trainer *t, const uint8_t *items_dev, const uint8_t *actions_dev,
                       const double *u_dev);
// update_value_model up to the flat gradient (writes t->targets) / calculate_advantage (t->adv).
bool dfrl_fused_covers_critic(const dfrl_trainer *t);  // rollout, critic step and GAE run on the fused kernels
int dfrl_fused_critic_gradient(dfrl_trainer *t, float *grad_dev, const dfrl_opt_spec *opt);
int dfrl_fused_gae(dfrl_trainer *t);
// The host-side decisions of the learn phase that depend on state from BEFORE the phase (bit 0: the logit table of
// the conv1d policy's table path is current, i.e. the rollout computed it with the parameters the first policy step
// sees, and that step skips its table pass). A captured learn phase is replayed only under the key it was captured with.
int dfrl_fused_learn_key(const dfrl_trainer *t);
// One policy gradient over all recorded rows (forward + loss gradient + backward), SUM over rows.
// Returns DFRL_ERR_UNSUPPORTED when the fused policy kernel does not cover this trainer.
// opt != null: the optimizer update runs inside the reduction kernel (caller bumps adam_t).
int dfrl_fused_policy_gradient(dfrl_trainer *t, int loss_kind, float *grad_dev, const dfrl_opt_spec *opt);
