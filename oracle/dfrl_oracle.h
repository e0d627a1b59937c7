/* oracle/dfrl_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's on-policy bin-packing hot path
 * (beehover/dependence_free_rl), in the batched struct-of-arrays formats of include/dfrl.h so
 * that the CUDA path and this oracle can be fed identical inputs.  Every function cites the
 * reference file:line it follows (paths relative to the reference root).
 *
 * PARITY PIN: the reference ships no tests or golden vectors, so this
 * oracle is pinned against the reference ITSELF: oracle/_ref/libdfrl_ref.so (the unmodified
 * reference compiled by oracle/Makefile) in tests/test_oracle_vs_ref.py when it is present, and
 * against the fixtures under tests/golden/ that tests/golden/make_golden.py generated from it.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 * The product (libdfrl_b200.so) never links, loads or calls it.
 */
#ifndef DFRL_ORACLE_H_
#define DFRL_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_LAYERS 16

enum { ORC_DENSE = 0, ORC_CONV1D = 1, ORC_RELU = 2, ORC_SOFTMAX = 3, ORC_SOFTMAX_CE = 4 };
enum { ORC_REINFORCE = 0, ORC_ACTOR_CRITIC = 1, ORC_PPO = 2, ORC_KL_PPO = 3 };
enum { ORC_SGD = 0, ORC_MOMENTUM = 1, ORC_ADAM = 2 };
enum { ORC_LOSS_SOFTMAX_LOG = 0, ORC_LOSS_CLIPPED = 1, ORC_LOSS_KL = 2 };
enum { ORC_HEUR_RANDOM = 0, ORC_HEUR_FIRSTFIT = 1, ORC_HEUR_BESTFIT = 2, ORC_HEUR_MINWASTE = 3 };

typedef struct {
  int n;
  int kind[ORC_MAX_LAYERS];
  int in[ORC_MAX_LAYERS];
  int out[ORC_MAX_LAYERS];
  int input_cols;
  /* Shared-trunk extension (BASELINE configs[2]; the reference's `model` is strictly sequential,
   * nn.h:467-542, so this is hand-composed from its layer forward / backward / gradient, nn.h:20-33):
   * n_params != 0 gives every parametric layer an explicit offset poff[l] into ONE flat vector of
   * n_params floats that another net addresses too -- two nets whose first layers carry the same
   * offsets share those layers. Gradients are n_params long with zeros in the other net's slots.
   * n_params == 0: the reference's sequential layout [W out x in][b out] per layer (nn.h:56-59). */
  int n_params;
  int poff[ORC_MAX_LAYERS];
} orc_net;

typedef struct {
  int n_bins, cap_w, cap_h;
  int item_w[2], item_h[2]; /* [0] = shape1 (coin toss true), [1] = shape2 */
  double p_shape1;
} orc_env_cfg;

typedef struct {
  int algo, work;
  float gamma, lambda;
  int epochs;
  float kl_target;
  int policy_opt, value_opt;
  float policy_lr, value_lr, policy_wd, value_wd, adam_beta1, adam_beta2;
} orc_train_cfg;

/* --- libstdc++ <random> restated (minstd_rand0 = std::default_random_engine) --- */
uint32_t orc_minstd_next(uint32_t *state);
uint32_t orc_minstd_seed(uint32_t seed);
double orc_canonical(uint32_t *state);
int orc_bernoulli(uint32_t *state, double p);
int orc_discrete(const float *w, int n, double u);
int orc_argmax(const float *w, int n);

/* --- environment --- */
void orc_env_cfg_default(orc_env_cfg *c);
void orc_env_reset_all(const orc_env_cfg *c, int8_t *state, int n_envs, const uint8_t *first_item);
void orc_env_step(const orc_env_cfg *c, int8_t *state, int n_envs, const uint8_t *actions,
                  const uint8_t *next_item, uint8_t *done, int8_t *terminal);
void orc_obs_encode(const int8_t *state, int rows, int stride, int n_bins, int cap_w, int cap_h,
                    float *obs);
int orc_heuristic_react(const orc_env_cfg *c, const int8_t *state, int stride, int env, int kind,
                        double u);

/* --- layers / model --- */
void orc_dense_forward(const float *params, int in, int out, const float *x, int rows, float *y);
void orc_dense_backward(const float *params, int in, int out, const float *dy, int rows, float *dx);
void orc_dense_gradient(int in, int out, const float *x, const float *dy, int rows, float *grad);
void orc_relu_forward(const float *x, size_t n, float *y);
void orc_relu_backward(const float *x, const float *dy, size_t n, float *dx);
void orc_softmax_forward(const float *x, int rows, int cols, float *y);
void orc_softmax_backward(const float *x, const float *dy, int rows, int cols, float *dx);

int orc_net_param_count(const orc_net *net);
int orc_net_output_cols(const orc_net *net);
int orc_net_layer_cols(const orc_net *net, int layer_output_index);
void orc_net_eval(const orc_net *net, const float *params, const float *x, int rows, float *y);
void orc_net_forward_gradient(const orc_net *net, const float *params, const float *x, int rows,
                              const float *dy, float *grad, float *out);

/* --- loss gradients --- */
void orc_loss_grad(int kind, const float *probs, const uint8_t *actions, const float *adv,
                   const float *p_old, float beta, int rows, int cols, float *out);
float orc_kl_next_beta(const float *probs, const float *p_old_full, int rows, int cols,
                       float d_targ, float beta);

/* --- returns / GAE --- */
void orc_returns(const uint8_t *done, const int *len, int n_envs, int max_len, float gamma,
                 float *g, double *acc2);
void orc_gae(const uint8_t *done, const float *v_start, const float *v_end, int n_envs, int T,
             float gamma, float lambda, float *targets, float *adv);

/* --- optimizers --- */
void orc_opt_step(int kind, float *params, const float *grad, float *state, int n, float lr,
                  float wd, float beta1, float beta2, float adam_t);

/* --- one learner step on [L][N] step-major records --- */
/* rec_state int8 [L][2B+2][N]; final_state int8 [2B+2][N] (env state after the rollout);
 * action/done uint8 [L][N]; len int32 [N] (REINFORCE) or NULL (= L everywhere);
 * p_old fp32 [L][N][B]. In/out: policy/value params and optimizer states (+ adam t, kl beta).
 * Optional outputs: adv [L][N], targets [L][N], grad logs (value: [Pv]; policy: [epochs][P]). */
int orc_learn(const orc_train_cfg *cfg, const orc_env_cfg *ecfg, int n_envs, int L,
              const int8_t *rec_state, const int8_t *final_state, const uint8_t *action,
              const uint8_t *done, const int *len, const float *p_old, const orc_net *pnet,
              float *pparams, float *pstate, float *p_adam_t, const orc_net *vnet, float *vparams,
              float *vstate, float *v_adam_t, float *kl_beta, float *adv_out, float *targets_out,
              float *vgrad_out, float *pgrad_log_out);

/* --- rollout with tapes (teacher forced or uniform-driven), records in dfrl layouts --- */
/* items uint8 [L][N]; forced uint8 [L][N] or NULL; u double [L][N] or NULL (one of forced/u).
 * mode 0 sample with u, 1 argmax, 2 forced. Writes rec_state/action/done/probs. */
int orc_rollout(const orc_env_cfg *ecfg, int8_t *state, int n_envs, int L, const orc_net *pnet,
                const float *pparams, int mode, const uint8_t *items, const uint8_t *forced,
                const double *u, int8_t *rec_state, uint8_t *rec_action, uint8_t *rec_done,
                float *rec_probs);

#ifdef __cplusplus
}
#endif
#endif
