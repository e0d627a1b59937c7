/* oracle/dfrl_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See dfrl_oracle.h.
 *
 * CPU restatement, in plain C, of the reference's on-policy bin-packing path.  Citations are
 * file:line relative to the reference root (beehover/dependence_free_rl).  libstdc++ pieces
 * cite /usr/include/c++/13/bits/random.tcc (the header the reference is compiled against here).
 *
 * Accumulator type: float by default (the reference accumulates in float, tensor.cc:400-438);
 * build with -DORC_ACC_DOUBLE for a double-accumulating variant used as the large-N yardstick.
 */
#include "dfrl_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef ORC_ACC_DOUBLE
typedef double acc_t;
#else
typedef float acc_t;
#endif

/* ------------------------------------------------------------------ libstdc++ <random> ---- */

/* std::default_random_engine = minstd_rand0 = LCG(16807, 0, 2^31-1) (tensor.cc:71-75). */
uint32_t orc_minstd_seed(uint32_t seed) {
  uint32_t s = seed % 2147483647u;
  return s == 0 ? 1u : s;
}
uint32_t orc_minstd_next(uint32_t *state) {
  *state = (uint32_t)(((uint64_t)*state * 16807ull) % 2147483647ull);
  return *state;
}
/* generate_canonical<double,53>(minstd_rand0): r = 2^31-2, m = 2 draws (random.tcc:3349-3381). */
double orc_canonical(uint32_t *state) {
  const long double r = 2147483646.0L - 1.0L + 1.0L;
  double sum = 0.0, tmp = 1.0;
  for (int k = 0; k < 2; ++k) {
    sum += (double)(orc_minstd_next(state) - 1u) * tmp;
    tmp = (double)((long double)tmp * r);
  }
  double ret = sum / tmp;
  if (ret >= 1.0)
    ret = nextafter(1.0, 0.0);
  return ret;
}
/* bernoulli_distribution::operator(): canonical() < p  (bin_packing.h:81). */
int orc_bernoulli(uint32_t *state, double p) { return orc_canonical(state) < p; }

/* discrete_distribution(w) then operator() with uniform u (tensor.cc:467-470;
 * random.tcc:2657-2678 normalise + partial sums + last = 1.0; 2699-2714 lower_bound). */
int orc_discrete(const float *w, int n, double u) {
  if (n < 2)
    return 0;
  double sum = 0.0;
  for (int i = 0; i < n; ++i)
    sum += (double)w[i];
  double c = 0.0;
  for (int i = 0; i < n; ++i) {
    c += (double)w[i] / sum;
    double ci = (i == n - 1) ? 1.0 : c;
    if (!(ci < u)) /* lower_bound: first element not less than u */
      return i;
  }
  return n; /* unreachable for u < 1 */
}
/* argmax: first maximal element (tensor.cc:464-466, std::ranges::max_element). */
int orc_argmax(const float *w, int n) {
  int best = 0;
  for (int i = 1; i < n; ++i)
    if (w[best] < w[i])
      best = i;
  return best;
}

/* ------------------------------------------------------------------------- environment ---- */

void orc_env_cfg_default(orc_env_cfg *c) {
  c->n_bins = 8; /* bin_packing.h:12 */
  c->cap_w = 8;  /* bin_packing.h:19 */
  c->cap_h = 8;
  c->item_w[0] = 4; /* shape1 {4,2}, bin_packing.h:73 */
  c->item_h[0] = 2;
  c->item_w[1] = 1; /* shape2 {1,2}, bin_packing.h:74 */
  c->item_h[1] = 2;
  c->p_shape1 = 0.4; /* bin_packing.h:50 */
}

static void set_item(const orc_env_cfg *c, int8_t *state, int n, int i, int shape1) {
  int B = c->n_bins;
  int k = shape1 ? 0 : 1; /* get_item(): toss ? shape1 : shape2 (bin_packing.h:76-79) */
  state[(size_t)(2 * B) * n + i] = (int8_t)c->item_w[k];
  state[(size_t)(2 * B + 1) * n + i] = (int8_t)c->item_h[k];
}
static void reset_one(const orc_env_cfg *c, int8_t *state, int n, int i, int shape1) {
  /* reset(): fresh observation(capacity) + get_item() (bin_packing.h:67-70) */
  for (int b = 0; b < c->n_bins; ++b) {
    state[(size_t)(2 * b) * n + i] = (int8_t)c->cap_w;
    state[(size_t)(2 * b + 1) * n + i] = (int8_t)c->cap_h;
  }
  set_item(c, state, n, i, shape1);
}
void orc_env_reset_all(const orc_env_cfg *c, int8_t *state, int n, const uint8_t *first_item) {
  for (int i = 0; i < n; ++i)
    reset_one(c, state, n, i, first_item[i]);
}

/* environment::apply (bin_packing.h:53-64) + agent::step tail (rl.h:336-346) +
 * agent::game_over / get_reward (bin_packing.h:94-106). terminal (optional) = state after
 * apply, before any reset (the transition's end_state). */
void orc_env_step(const orc_env_cfg *c, int8_t *state, int n, const uint8_t *actions,
                  const uint8_t *next_item, uint8_t *done, int8_t *terminal) {
  int B = c->n_bins, P = 2 * B + 2;
  for (int i = 0; i < n; ++i) {
    int a = actions[i];
    int8_t *bw = &state[(size_t)(2 * a) * n + i];
    int8_t *bh = &state[(size_t)(2 * a + 1) * n + i];
    *bw = (int8_t)(*bw - state[(size_t)(2 * B) * n + i]);
    *bh = (int8_t)(*bh - state[(size_t)(2 * B + 1) * n + i]);
    int over = (*bw < 0) || (*bh < 0);
    /* game_over scans every bin (bin_packing.h:95-101); only bin a can have gone negative
       because a negative bin ends the episode immediately. */
    if (terminal)
      for (int p = 0; p < P; ++p)
        terminal[(size_t)p * n + i] = state[(size_t)p * n + i];
    if (done)
      done[i] = (uint8_t)over;
    if (over)
      reset_one(c, state, n, i, next_item[i]); /* rl.h:341-343 */
    else
      set_item(c, state, n, i, next_item[i]); /* bin_packing.h:63 */
  }
}

/* observation::to_vector (bin_packing.h:31-40). */
void orc_obs_encode(const int8_t *state, int rows, int stride, int B, int cap_w, int cap_h,
                    float *obs) {
  for (int r = 0; r < rows; ++r) {
    float iw = (float)state[(size_t)(2 * B) * stride + r] / (float)cap_w;
    float ih = (float)state[(size_t)(2 * B + 1) * stride + r] / (float)cap_h;
    for (int b = 0; b < B; ++b) {
      float *o = obs + ((size_t)r * B + b) * 4;
      o[0] = (float)state[(size_t)(2 * b) * stride + r] / (float)cap_w;
      o[1] = (float)state[(size_t)(2 * b + 1) * stride + r] / (float)cap_h;
      o[2] = iw;
      o[3] = ih;
    }
  }
}

/* Heuristic policies: random_policy (rl.h:305-315), firstfit_agent.cc:10-28,
 * bestfit_agent.cc:10-30, minwaste_agent.cc:10-39; all end in argmax / discrete sampling. */
int orc_heuristic_react(const orc_env_cfg *c, const int8_t *state, int stride, int env, int kind,
                        double u) {
  int B = c->n_bins;
  float scores[64];
  int iw = state[(size_t)(2 * B) * stride + env], ih = state[(size_t)(2 * B + 1) * stride + env];
  if (kind == ORC_HEUR_RANDOM) {
    for (int b = 0; b < B; ++b)
      scores[b] = (float)(1.0 / B);
    return orc_discrete(scores, B, u);
  }
  for (int b = 0; b < B; ++b) {
    int bw = state[(size_t)(2 * b) * stride + env], bh = state[(size_t)(2 * b + 1) * stride + env];
    int fits = iw <= bw && ih <= bh;
    if (kind == ORC_HEUR_FIRSTFIT) {
      scores[b] = 0;
    } else if (kind == ORC_HEUR_BESTFIT) {
      scores[b] = fits ? (float)iw / bw + (float)ih / bh : -1.0f;
    } else {
      if (!fits)
        scores[b] = -1;
      else {
        float r1 = (float)(bw - iw), r2 = (float)(bh - ih);
        if ((r1 == c->cap_w / 2 && r2 == 0) || (r1 == 0 && r2 == c->cap_h / 2))
          scores[b] = 0;
        else
          scores[b] = 1;
      }
    }
  }
  if (kind == ORC_HEUR_FIRSTFIT)
    for (int b = 0; b < B; ++b) {
      int bw = state[(size_t)(2 * b) * stride + env];
      int bh = state[(size_t)(2 * b + 1) * stride + env];
      if (iw <= bw && ih <= bh) {
        scores[b] = 1;
        break;
      }
    }
  return orc_argmax(scores, B);
}

/* ------------------------------------------------------------------------------ layers ---- */

/* matmul_layer::forward (nn.h:72-79): matmul_transposed(input, a_) then += b_ per row;
 * matmul_transposed = dot of rows (tensor.cc:218-227). */
void orc_dense_forward(const float *params, int in, int out, const float *x, int rows, float *y) {
  const float *W = params, *b = params + (size_t)in * out;
  /* rows are independent: the OpenMP build (liboracle64mt.so, large-N yardstick) gives the same
   * bits for any thread count */
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (int r = 0; r < rows; ++r)
    for (int o = 0; o < out; ++o) {
      acc_t s = 0;
      for (int k = 0; k < in; ++k)
        s += (acc_t)x[(size_t)r * in + k] * W[(size_t)o * in + k];
      y[(size_t)r * out + o] = (float)s + b[o];
    }
}
/* matmul_layer::backward (nn.h:81-83): matmul(backprop, a_) = backprop . W. */
void orc_dense_backward(const float *params, int in, int out, const float *dy, int rows, float *dx) {
  const float *W = params;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < in; ++k) {
      acc_t s = 0;
      for (int o = 0; o < out; ++o)
        s += (acc_t)dy[(size_t)r * out + o] * W[(size_t)o * in + k];
      dx[(size_t)r * in + k] = (float)s;
    }
}
/* matmul_layer::gradient (nn.h:85-100): d_a = transpose(backprop) . input (SUM over rows),
 * d_b = sum of backprop rows. Every entry is accumulated over the rows in ascending order; the
 * row loop is the outer one (cache friendly at large N) and the OpenMP build splits the OUTPUT
 * index over threads, so the order of additions per entry -- and therefore every bit of the
 * result -- does not depend on the loop nest or the thread count. */
void orc_dense_gradient(int in, int out, const float *x, const float *dy, int rows, float *grad) {
  acc_t *acc = (acc_t *)calloc((size_t)(in + 1) * out, sizeof(acc_t));
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (int o = 0; o < out; ++o) {
    acc_t *a = acc + (size_t)o * in, *ab = acc + (size_t)in * out + o;
    for (int r = 0; r < rows; ++r) {
      const float d = dy[(size_t)r * out + o];
      const float *xr = x + (size_t)r * in;
      for (int k = 0; k < in; ++k)
        a[k] += (acc_t)d * xr[k];
      *ab += d;
    }
  }
  for (size_t i = 0; i < (size_t)(in + 1) * out; ++i)
    grad[i] = (float)acc[i];
  free(acc);
}
/* relu_activation::forward / backward (nn.h:354-376): mask by PRE-activation > 0. */
void orc_relu_forward(const float *x, size_t n, float *y) {
  for (size_t i = 0; i < n; ++i)
    y[i] = x[i] > 0 ? x[i] : 0;
}
void orc_relu_backward(const float *x, const float *dy, size_t n, float *dx) {
  for (size_t i = 0; i < n; ++i)
    dx[i] = x[i] > 0 ? dy[i] : 0;
}
/* softmax_layer::forward (nn.h:382-392): expf, row sum, divide -- no max subtraction. */
void orc_softmax_forward(const float *x, int rows, int cols, float *y) {
  for (int r = 0; r < rows; ++r) {
    acc_t s = 0;
    for (int c = 0; c < cols; ++c) {
      y[(size_t)r * cols + c] = expf(x[(size_t)r * cols + c]);
      s += y[(size_t)r * cols + c];
    }
    for (int c = 0; c < cols; ++c)
      y[(size_t)r * cols + c] = y[(size_t)r * cols + c] / (float)s;
  }
}
/* softmax_layer::backward (nn.h:393-417): (diag(s) - s s^T) . g per row. */
void orc_softmax_backward(const float *x, const float *dy, int rows, int cols, float *dx) {
  float *s = (float *)malloc(sizeof(float) * (size_t)cols);
  for (int r = 0; r < rows; ++r) {
    orc_softmax_forward(x + (size_t)r * cols, 1, cols, s);
    for (int j = 0; j < cols; ++j) {
      acc_t a = 0;
      for (int k = 0; k < cols; ++k) {
        float pd = (j == k ? s[j] : 0.0f) - s[j] * s[k];
        a += (acc_t)pd * dy[(size_t)r * cols + k];
      }
      dx[(size_t)r * cols + j] = (float)a;
    }
  }
  free(s);
}

/* ------------------------------------------------------------------------------- model ---- */

static int layer_params(const orc_net *net, int l) {
  int k = net->kind[l];
  return (k == ORC_DENSE || k == ORC_CONV1D) ? (net->in[l] + 1) * net->out[l] : 0;
}
/* Output width of layer l given its input width. conv1d: points = cols / in_ch (nn.h:131). */
static int layer_out_cols(const orc_net *net, int l, int in_cols) {
  switch (net->kind[l]) {
  case ORC_DENSE:
    return net->out[l];
  case ORC_CONV1D:
    return in_cols / net->in[l] * net->out[l];
  default:
    return in_cols;
  }
}
/* Offset of layer l's parameters in the flat vector: explicit (shared-trunk nets) or the running sum
 * of the reference's sequential layout. */
static size_t layer_off(const orc_net *net, int l) {
  if (net->n_params)
    return (size_t)net->poff[l];
  size_t off = 0;
  for (int j = 0; j < l; ++j)
    off += (size_t)layer_params(net, j);
  return off;
}
int orc_net_param_count(const orc_net *net) {
  if (net->n_params)
    return net->n_params;
  int p = 0;
  for (int l = 0; l < net->n; ++l)
    p += layer_params(net, l);
  return p;
}
int orc_net_layer_cols(const orc_net *net, int idx) {
  int c = net->input_cols;
  for (int l = 0; l <= idx && l < net->n; ++l)
    c = layer_out_cols(net, l, c);
  return c;
}
int orc_net_output_cols(const orc_net *net) { return orc_net_layer_cols(net, net->n - 1); }

static void layer_forward(const orc_net *net, int l, const float *p, const float *x, int rows,
                          int in_cols, float *y) {
  switch (net->kind[l]) {
  case ORC_DENSE:
    orc_dense_forward(p, net->in[l], net->out[l], x, rows, y);
    break;
  case ORC_CONV1D: /* nn.h:127-147: dense over (rows * points, in_ch) */
    orc_dense_forward(p, net->in[l], net->out[l], x, rows * (in_cols / net->in[l]), y);
    break;
  case ORC_RELU:
    orc_relu_forward(x, (size_t)rows * in_cols, y);
    break;
  default:
    orc_softmax_forward(x, rows, in_cols, y);
  }
}

/* model::eval (nn.h:473-479). */
void orc_net_eval(const orc_net *net, const float *params, const float *x, int rows, float *y) {
  int cols = net->input_cols;
  const float *cur = x;
  float *bufs[2] = {NULL, NULL};
  for (int l = 0; l < net->n; ++l) {
    int oc = layer_out_cols(net, l, cols);
    float *dst = (l == net->n - 1) ? y : (bufs[l & 1] = (float *)realloc(bufs[l & 1], sizeof(float) * (size_t)rows * oc));
    layer_forward(net, l, params + layer_off(net, l), cur, rows, cols, dst);
    cur = dst;
    cols = oc;
  }
  free(bufs[0]);
  free(bufs[1]);
}

/* optimizer::step minus the update (nn.h:594-603): model::forward (481-488), caller's loss
 * gradient dy at the output, model::gradient reverse sweep (510-528). */
void orc_net_forward_gradient(const orc_net *net, const float *params, const float *x, int rows,
                              const float *dy, float *grad, float *out) {
  int n = net->n;
  float *acts[ORC_MAX_LAYERS + 1];
  int cols[ORC_MAX_LAYERS + 1];
  size_t poff[ORC_MAX_LAYERS + 1];
  acts[0] = (float *)x;
  cols[0] = net->input_cols;
  poff[0] = layer_off(net, 0);
  for (int l = 0; l < n; ++l) {
    cols[l + 1] = layer_out_cols(net, l, cols[l]);
    acts[l + 1] = (float *)malloc(sizeof(float) * (size_t)rows * cols[l + 1]);
    layer_forward(net, l, params + poff[l], acts[l], rows, cols[l], acts[l + 1]);
    poff[l + 1] = l + 1 < n ? layer_off(net, l + 1) : 0;
  }
  if (net->n_params) /* shared-trunk nets: the other net's slots of the flat gradient are zero */
    memset(grad, 0, sizeof(float) * (size_t)net->n_params);
  if (out)
    memcpy(out, acts[n], sizeof(float) * (size_t)rows * cols[n]);
  float *back = (float *)malloc(sizeof(float) * (size_t)rows * cols[n]);
  memcpy(back, dy, sizeof(float) * (size_t)rows * cols[n]);
  for (int l = n - 1; l >= 0; --l) {
    int k = net->kind[l];
    const float *p = params + poff[l];
    int r2 = rows;
    if (k == ORC_CONV1D)
      r2 = rows * (cols[l] / net->in[l]);
    if (k == ORC_DENSE || k == ORC_CONV1D)
      orc_dense_gradient(net->in[l], net->out[l], acts[l], back, r2, grad + poff[l]);
    if (l == 0)
      break; /* first layer gets no dX (nn.h:525-527) */
    float *nb = (float *)malloc(sizeof(float) * (size_t)rows * cols[l]);
    if (k == ORC_DENSE || k == ORC_CONV1D)
      orc_dense_backward(p, net->in[l], net->out[l], back, r2, nb);
    else if (k == ORC_RELU)
      orc_relu_backward(acts[l], back, (size_t)rows * cols[l], nb);
    else if (k == ORC_SOFTMAX)
      orc_softmax_backward(acts[l], back, rows, cols[l], nb);
    else /* softmax_cross_entropy_layer::backward = identity (nn.h:428-430) */
      memcpy(nb, back, sizeof(float) * (size_t)rows * cols[l]);
    free(back);
    back = nb;
  }
  free(back);
  for (int l = 1; l <= n; ++l)
    free(acts[l]);
}

/* ----------------------------------------------------------------------- loss gradients ---- */

/* policy_loss -> softmax_gradient_log (policy_gradient.h:16-26, rl.h:45-52);
 * surrogate_loss -> clipped_gradient (policy_gradient.h:28-38, rl.h:54-74);
 * kl_regulated_loss (policy_gradient.h:47-66): softmax_gradient_log + beta (p - p_old). */
void orc_loss_grad(int kind, const float *probs, const uint8_t *actions, const float *adv,
                   const float *p_old, float beta, int rows, int cols, float *out) {
  for (int r = 0; r < rows; ++r) {
    const float *p = probs + (size_t)r * cols;
    float *o = out + (size_t)r * cols;
    int a = actions[r];
    float A = adv[r];
    if (kind == ORC_LOSS_CLIPPED) {
      const float eps = 0.2f;
      for (int c = 0; c < cols; ++c)
        o[c] = 0;
      float ratio = p[a] / p_old[r];
      float clipped = ratio;
      if (ratio > 1 + eps)
        clipped = 1 + eps;
      else if (ratio < 1 - eps)
        clipped = 1 - eps;
      float x = clipped * A, y = ratio * A;
      float g = (y < x ? y : x) * -1; /* std::min(a, b) */
      o[a] = g / p[a];
    } else {
      for (int c = 0; c < cols; ++c)
        o[c] = p[c] * A;
      o[a] -= A;
      if (kind == ORC_LOSS_KL)
        for (int c = 0; c < cols; ++c)
          o[c] += (p[c] - p_old[(size_t)r * cols + c]) * beta;
    }
  }
}
/* kl_regulated_loss tail (policy_gradient.h:68-83): mean over rows of D_KL(p_old || p)
 * (kl_divergence, 41-45), then beta halves / doubles around d_targ, clamped to [1e-25, 0.1]. */
float orc_kl_next_beta(const float *probs, const float *p_old, int rows, int cols, float d_targ,
                       float beta) {
  float d_average = 0;
  for (int r = 0; r < rows; ++r) {
    acc_t d = 0;
    for (int c = 0; c < cols; ++c) {
      float po = p_old[(size_t)r * cols + c], p = probs[(size_t)r * cols + c];
      d += po * logf(po / p);
    }
    d_average += (float)d;
  }
  d_average /= rows;
  if (fabsf(d_average) < d_targ / 1.5f)
    beta /= 2;
  else if (fabsf(d_average) > d_targ * 1.5f)
    beta *= 2;
  if (beta < 1e-25f)
    beta = 1e-25f;
  if (beta > 0.1f)
    beta = 0.1f;
  return beta;
}

/* ------------------------------------------------------------------------ returns / GAE ---- */

/* policy_gradient_learner::get_advantages (policy_gradient.h:125-147), without the final
 * baseline subtraction. Records are step-major [L][N]; env i holds len[i] transitions split
 * into trajectories by done flags.  The reference iterates a trajectory FORWARD while writing
 * BACKWARD (137-141): after consuming r_0..r_k it stores the running discounted sum at
 * position m-1-k.  acc2[0] += reward_slice[0] per trajectory, acc2[1] += 1 (142-145). */
void orc_returns(const uint8_t *done, const int *len, int n, int L, float gamma, float *g,
                 double *acc2) {
  for (int i = 0; i < n; ++i) {
    int Li = len ? len[i] : L;
    int first = 0;
    while (first < Li) {
      int last = first;
      while (last < Li - 1 && !done[(size_t)last * n + i])
        ++last;
      int m = last - first + 1;
      float reward = 0;
      for (int k = 0; k < m; ++k) {
        float r = done[(size_t)(first + k) * n + i] ? 0.0f : 1.0f;
        reward = r + gamma * reward;
        g[(size_t)(first + m - 1 - k) * n + i] = reward;
      }
      if (acc2) {
        acc2[0] += g[(size_t)first * n + i];
        acc2[1] += 1.0;
      }
      first = last + 1;
    }
  }
}

/* update_value_model targets (policy_gradient.h:196-215) and calculate_advantage (220-281) on
 * [T][N] records.  Next value of step t: end-state value where the trajectory ends there
 * (done, or the rollout's last step), else the next start-state value. */
void orc_gae(const uint8_t *done, const float *v_start, const float *v_end, int n, int T,
             float gamma, float lambda, float *targets, float *adv) {
  float *delta = (float *)malloc(sizeof(float) * (size_t)T);
  for (int i = 0; i < n; ++i) {
    for (int t = 0; t < T; ++t) {
      size_t k = (size_t)t * n + i;
      int d = done[k], ends = d || t == T - 1;
      float r = d ? 0.0f : 1.0f;
      float vn = ends ? v_end[k] : v_start[k + n];
      if (targets)
        targets[k] = r + gamma * vn; /* 207-208: NOT masked at terminals */
      float vn_adv = d ? 0.0f : vn;  /* 230-236: V[end] = 0 if frozen */
      delta[t] = r + gamma * vn_adv - v_start[k]; /* 253-255 */
    }
    if (!adv)
      continue;
    /* 262-276: A_t = sum_{i=t}^{end-1} coef_i delta_i, coef built by repeated multiplication,
       summed forward, restricted to the transition's own trajectory. */
    for (int t = 0; t < T; ++t) {
      float a = 0, coef = 1;
      for (int j = t; j < T; ++j) {
        a += delta[j] * coef;
        coef *= lambda * gamma;
        if (done[(size_t)j * n + i])
          break;
      }
      adv[(size_t)t * n + i] = a;
    }
  }
  free(delta);
}

/* --------------------------------------------------------------------------- optimizers ---- */

/* sgd (nn.h:622-625), momentum (636-650), adam (666-690: eps 1e-7 outside the sqrt, float t). */
void orc_opt_step(int kind, float *params, const float *grad, float *state, int n, float lr,
                  float wd, float beta1, float beta2, float adam_t) {
  if (kind == ORC_SGD) {
    for (int i = 0; i < n; ++i)
      params[i] = params[i] * (1 - wd) - grad[i] * lr;
  } else if (kind == ORC_MOMENTUM) {
    const float rho = 0.9f;
    for (int i = 0; i < n; ++i) {
      state[i] = state[i] * rho + grad[i];
      params[i] = params[i] - state[i] * lr;
    }
  } else {
    float *m = state, *v = state + n;
    float c1 = 1 - powf(beta1, adam_t), c2 = 1 - powf(beta2, adam_t);
    for (int i = 0; i < n; ++i) {
      m[i] = m[i] * beta1 + grad[i] * (1 - beta1);
      v[i] = v[i] * beta2 + grad[i] * grad[i] * (1 - beta2);
      float mu = m[i] / c1, vu = v[i] / c2;
      params[i] = params[i] - mu * lr / (sqrtf(vu) + 1e-7f);
    }
  }
}

/* ------------------------------------------------------------------------ learner step ---- */

static void copy_state_col(const int8_t *src, int stride, int i, int P, int8_t *dst, int dstride,
                           int j) {
  for (int p = 0; p < P; ++p)
    dst[(size_t)p * dstride + j] = src[(size_t)p * stride + i];
}

static void opt_apply(int kind, float *params, const float *grad, float *state, int n, float lr,
                      float wd, float b1, float b2, float *adam_t) {
  orc_opt_step(kind, params, grad, state, n, lr, wd, b1, b2, adam_t ? *adam_t : 1.0f);
  if (kind == ORC_ADAM && adam_t)
    *adam_t += 1; /* nn.h:686 */
}

int orc_learn(const orc_train_cfg *cfg, const orc_env_cfg *ecfg, int n, int L,
              const int8_t *rec_state, const int8_t *final_state, const uint8_t *action,
              const uint8_t *done, const int *len, const float *p_old, const orc_net *pnet,
              float *pparams, float *pstate, float *p_adam_t, const orc_net *vnet, float *vparams,
              float *vstate, float *v_adam_t, float *kl_beta, float *adv_out, float *targets_out,
              float *vgrad_out, float *pgrad_log_out) {
  int B = ecfg->n_bins, P = 2 * B + 2, O = 4 * B;
  int PP = orc_net_param_count(pnet);
  int pc = orc_net_output_cols(pnet);
  if (pc != B)
    return -1;

  if (cfg->algo == ORC_REINFORCE) {
    /* policy_gradient_learner::learn (policy_gradient.h:95-123): rows = start states of every
       transition; advantages = returns - mean over trajectories of G[first]. */
    size_t rows = 0;
    for (int i = 0; i < n; ++i)
      rows += (size_t)(len ? len[i] : L);
    int8_t *st = (int8_t *)malloc((size_t)P * rows);
    uint8_t *act = (uint8_t *)malloc(rows);
    float *A = (float *)malloc(sizeof(float) * rows);
    float *g = (float *)calloc((size_t)L * n, sizeof(float));
    double acc[2] = {0, 0};
    orc_returns(done, len, n, L, cfg->gamma, g, acc);
    float baseline = (float)acc[0] / (float)acc[1]; /* total_reward / experience.size() */
    size_t k = 0;
    for (int i = 0; i < n; ++i)
      for (int t = 0; t < (len ? len[i] : L); ++t) {
        copy_state_col(rec_state + (size_t)t * P * n, n, i, P, st, (int)rows, (int)k);
        act[k] = action[(size_t)t * n + i];
        A[k] = g[(size_t)t * n + i] - baseline;
        if (adv_out)
          adv_out[(size_t)t * n + i] = A[k];
        ++k;
      }
    float *obs = (float *)malloc(sizeof(float) * rows * O);
    orc_obs_encode(st, (int)rows, (int)rows, B, ecfg->cap_w, ecfg->cap_h, obs);
    float *probs = (float *)malloc(sizeof(float) * rows * B);
    float *dy = (float *)malloc(sizeof(float) * rows * B);
    float *grad = (float *)malloc(sizeof(float) * (size_t)PP);
    orc_net_eval(pnet, pparams, obs, (int)rows, probs);
    orc_loss_grad(ORC_LOSS_SOFTMAX_LOG, probs, act, A, NULL, 0, (int)rows, B, dy);
    orc_net_forward_gradient(pnet, pparams, obs, (int)rows, dy, grad, NULL);
    if (pgrad_log_out)
      memcpy(pgrad_log_out, grad, sizeof(float) * (size_t)PP);
    opt_apply(cfg->policy_opt, pparams, grad, pstate, PP, cfg->policy_lr, cfg->policy_wd,
              cfg->adam_beta1, cfg->adam_beta2, p_adam_t);
    free(st); free(act); free(A); free(g); free(obs); free(probs); free(dy); free(grad);
    return 0;
  }

  /* actor_critic_learner::learn (policy_gradient.h:159-185).  Rows, env-major: per env, per
     trajectory (split at done flags; the last one may be open): its start states, then ONE
     end-state row (action duplicated, 179-180).  */
  int T = L;
  int PV = orc_net_param_count(vnet);
  size_t max_rows = (size_t)2 * T * n;
  int8_t *st = (int8_t *)malloc((size_t)P * max_rows);
  uint8_t *act = (uint8_t *)malloc(max_rows);
  float *pold_sel = (float *)malloc(sizeof(float) * max_rows);
  float *pold_full = (float *)malloc(sizeof(float) * max_rows * B);
  int *row_t = (int *)malloc(sizeof(int) * max_rows);   /* t of the transition, -1 for end rows */
  int *row_env = (int *)malloc(sizeof(int) * max_rows);
  int *row_last_t = (int *)malloc(sizeof(int) * max_rows); /* end rows: t of the last transition */
  uint8_t *row_frozen = (uint8_t *)malloc(max_rows);
  int8_t *term = (int8_t *)malloc((size_t)P);
  size_t rows = 0;
  for (int i = 0; i < n; ++i) {
    for (int t = 0; t < T; ++t) {
      size_t k = (size_t)t * n + i;
      copy_state_col(rec_state + (size_t)t * P * n, n, i, P, st, (int)max_rows, (int)rows);
      act[rows] = action[k];
      pold_sel[rows] = p_old[k * B + action[k]];
      memcpy(pold_full + rows * B, p_old + k * B, sizeof(float) * B);
      row_t[rows] = t;
      row_env[rows] = i;
      row_frozen[rows] = 0;
      ++rows;
      int d = done[k];
      if (d || t == T - 1) {
        /* end-state row: transitions.back().end_state (180) */
        if (d) {
          /* terminal state = start state with bin[a] -= item, item kept (bin_packing.h:54-61) */
          for (int p = 0; p < P; ++p)
            term[p] = rec_state[((size_t)t * P + p) * n + i];
          term[2 * action[k]] = (int8_t)(term[2 * action[k]] - term[2 * B]);
          term[2 * action[k] + 1] = (int8_t)(term[2 * action[k] + 1] - term[2 * B + 1]);
          for (int p = 0; p < P; ++p)
            st[(size_t)p * max_rows + rows] = term[p];
        } else {
          copy_state_col(final_state, n, i, P, st, (int)max_rows, (int)rows);
        }
        act[rows] = action[k];
        pold_sel[rows] = pold_sel[rows - 1];
        memcpy(pold_full + rows * B, p_old + k * B, sizeof(float) * B);
        row_t[rows] = -1;
        row_env[rows] = i;
        row_last_t[rows] = t;
        row_frozen[rows] = (uint8_t)d;
        ++rows;
      }
    }
  }
  int R = (int)rows;
  float *obs = (float *)malloc(sizeof(float) * rows * O);
  orc_obs_encode(st, R, (int)max_rows, B, ecfg->cap_w, ecfg->cap_h, obs);

  /* update_value_model (196-218) */
  float *values = (float *)malloc(sizeof(float) * rows);
  float *updated = (float *)malloc(sizeof(float) * rows);
  float *dv = (float *)malloc(sizeof(float) * rows);
  float *vout = (float *)malloc(sizeof(float) * rows);
  float *vgrad = (float *)malloc(sizeof(float) * (size_t)PV);
  orc_net_eval(vnet, vparams, obs, R, values);
  for (int r = 0; r < R; ++r) {
    if (row_t[r] >= 0) {
      float rew = done[(size_t)row_t[r] * n + row_env[r]] ? 0.0f : 1.0f;
      updated[r] = rew + cfg->gamma * values[r + 1]; /* 207-208 */
      if (targets_out)
        targets_out[(size_t)row_t[r] * n + row_env[r]] = updated[r];
    } else {
      updated[r] = values[r]; /* 211 */
    }
  }
  /* optimizer::step with square_loss_grad = out - label (nn.h:548-550, 594-605) */
  orc_net_eval(vnet, vparams, obs, R, vout);
  for (int r = 0; r < R; ++r)
    dv[r] = vout[r] - updated[r];
  orc_net_forward_gradient(vnet, vparams, obs, R, dv, vgrad, NULL);
  if (vgrad_out)
    memcpy(vgrad_out, vgrad, sizeof(float) * (size_t)PV);
  opt_apply(cfg->value_opt, vparams, vgrad, vstate, PV, cfg->value_lr, cfg->value_wd,
            cfg->adam_beta1, cfg->adam_beta2, v_adam_t);

  /* calculate_advantage (220-281) with the UPDATED critic */
  float *A = (float *)malloc(sizeof(float) * rows);
  float *deltas = (float *)malloc(sizeof(float) * rows);
  orc_net_eval(vnet, vparams, obs, R, values);
  for (int r = 0; r < R; ++r)
    if (row_t[r] < 0 && row_frozen[r])
      values[r] = 0; /* 230-236 */
  for (int r = 0; r < R; ++r) {
    if (row_t[r] >= 0) {
      float rew = done[(size_t)row_t[r] * n + row_env[r]] ? 0.0f : 1.0f;
      deltas[r] = rew + cfg->gamma * values[r + 1] - values[r]; /* 253-255 */
    } else {
      deltas[r] = 0; /* 259 */
    }
  }
  for (int r = 0; r < R;) {
    int end = r;
    while (row_t[end] >= 0)
      ++end; /* end = index of this trajectory's end row */
    for (int c = r; c < end; ++c) {
      float a = 0, coef = 1; /* 266-272: forward accumulation with running coefficient */
      for (int i2 = c; i2 < end; ++i2) {
        a += deltas[i2] * coef;
        coef *= cfg->lambda * cfg->gamma;
      }
      A[c] = a;
      if (adv_out)
        adv_out[(size_t)row_t[c] * n + row_env[c]] = a;
    }
    A[end] = 0; /* 277 */
    r = end + 1;
  }

  /* optimize_action: AC one policy_loss step (187-194); PPO k surrogate steps (297-307);
     KL-PPO k kl_regulated steps (318-330). */
  int epochs = cfg->algo == ORC_ACTOR_CRITIC ? 1 : cfg->epochs;
  int kind = cfg->algo == ORC_ACTOR_CRITIC ? ORC_LOSS_SOFTMAX_LOG
             : cfg->algo == ORC_PPO        ? ORC_LOSS_CLIPPED
                                           : ORC_LOSS_KL;
  float *probs = (float *)malloc(sizeof(float) * rows * B);
  float *dy = (float *)malloc(sizeof(float) * rows * B);
  float *pgrad = (float *)malloc(sizeof(float) * (size_t)PP);
  for (int e = 0; e < epochs; ++e) {
    orc_net_eval(pnet, pparams, obs, R, probs);
    float beta = kl_beta ? *kl_beta : 0.0f;
    orc_loss_grad(kind, probs, act, A, kind == ORC_LOSS_KL ? pold_full : pold_sel, beta, R, B, dy);
    if (kind == ORC_LOSS_KL && kl_beta)
      *kl_beta = orc_kl_next_beta(probs, pold_full, R, B, cfg->kl_target, beta);
    orc_net_forward_gradient(pnet, pparams, obs, R, dy, pgrad, NULL);
    if (pgrad_log_out)
      memcpy(pgrad_log_out + (size_t)e * PP, pgrad, sizeof(float) * (size_t)PP);
    opt_apply(cfg->policy_opt, pparams, pgrad, pstate, PP, cfg->policy_lr, cfg->policy_wd,
              cfg->adam_beta1, cfg->adam_beta2, p_adam_t);
  }
  free(st); free(act); free(pold_sel); free(pold_full); free(row_t); free(row_env);
  free(row_last_t); free(row_frozen); free(term); free(obs); free(values); free(updated);
  free(dv); free(vout); free(vgrad); free(A); free(deltas); free(probs); free(dy); free(pgrad);
  return 0;
}

/* ----------------------------------------------------------------------------- rollout ---- */

/* agent::step for every env, L times (rl.h:325-349; policy_gradient_policy::react,
 * policy_gradient.h:343-350). The item tape gives the item drawn after each step. */
int orc_rollout(const orc_env_cfg *ecfg, int8_t *state, int n, int L, const orc_net *pnet,
                const float *pparams, int mode, const uint8_t *items, const uint8_t *forced,
                const double *u, int8_t *rec_state, uint8_t *rec_action, uint8_t *rec_done,
                float *rec_probs) {
  int B = ecfg->n_bins, P = 2 * B + 2, O = 4 * B;
  float *obs = (float *)malloc(sizeof(float) * (size_t)n * O);
  float *probs = (float *)malloc(sizeof(float) * (size_t)n * B);
  uint8_t *act = (uint8_t *)malloc((size_t)n);
  for (int t = 0; t < L; ++t) {
    memcpy(rec_state + (size_t)t * P * n, state, (size_t)P * n);
    orc_obs_encode(state, n, n, B, ecfg->cap_w, ecfg->cap_h, obs);
    orc_net_eval(pnet, pparams, obs, n, probs);
    for (int i = 0; i < n; ++i) {
      const float *p = probs + (size_t)i * B;
      int a;
      if (mode == 2)
        a = forced[(size_t)t * n + i];
      else if (mode == 1)
        a = orc_argmax(p, B);
      else
        a = orc_discrete(p, B, u[(size_t)t * n + i]);
      act[i] = (uint8_t)a;
    }
    if (rec_probs)
      memcpy(rec_probs + (size_t)t * n * B, probs, sizeof(float) * (size_t)n * B);
    memcpy(rec_action + (size_t)t * n, act, (size_t)n);
    orc_env_step(ecfg, state, n, act, items + (size_t)t * n, rec_done + (size_t)t * n, NULL);
  }
  free(obs); free(probs); free(act);
  return 0;
}
