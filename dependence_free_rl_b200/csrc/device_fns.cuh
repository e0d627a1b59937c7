// device_fns.cuh -- device functions shared by the standalone kernels (rl_ops.cu) and the
// fused rollout / learn kernels (fused.cu), so both paths compute the same arithmetic.
#pragma once

#include <stdint.h>

#include "../../include/dfrl.h"

// std::discrete_distribution (reference tensor.cc:467-470; libstdc++ random.tcc:2657-2714):
// normalise in double, partial sums, last cumulative forced to 1.0, lower_bound(cum, u).
__device__ __forceinline__ int discrete_sample(const float *__restrict__ w, int n, double u) {
  if (n < 2)
    return 0;
  double sum = 0.0;
  for (int i = 0; i < n; ++i)
    sum += (double)w[i];
  double c = 0.0;
  for (int i = 0; i < n; ++i) {
    c += (double)w[i] / sum;
    double ci = (i == n - 1) ? 1.0 : c;
    if (!(ci < u))
      return i;
  }
  return n - 1;
}

// argmax, first maximum wins (tensor.cc:464-466).
__device__ __forceinline__ int argmax_first(const float *__restrict__ w, int n) {
  int best = 0;
  float bv = w[0];
  for (int i = 1; i < n; ++i) {
    float v = w[i];
    if (bv < v) {
      bv = v;
      best = i;
    }
  }
  return best;
}

// discrete_action::clipped_gradient (rl.h:54-74), the one non-zero column:
// -min(clip(r, 1-eps, 1+eps) A, r A) / p[a], r = p[a] / p_old[a], eps = 0.2.
__device__ __forceinline__ float clipped_grad(float pa, float pold, float A) {
  const float eps = 0.2f;
  float ratio = pa / pold;
  float clipped = ratio;
  if (ratio > 1.f + eps)
    clipped = 1.f + eps;
  else if (ratio < 1.f - eps)
    clipped = 1.f - eps;
  float x = clipped * A, y = ratio * A;
  float g = (y < x ? y : x) * -1.f;
  return g / pa;
}

// update_value_model targets (policy_gradient.h:196-215) and calculate_advantage (220-281)
// for env i of [T][n] records. Backward recurrence A_t = delta_t + gamma lambda A_{t+1} within a
// trajectory (the reference sums the same series forward, 262-276).
__device__ __forceinline__ void gae_env(const uint8_t *__restrict__ done,
                                        const float *__restrict__ v_start,
                                        const float *__restrict__ v_end, int n, int T, int i,
                                        float gamma, float lambda, float *__restrict__ targets,
                                        float *__restrict__ adv) {
  float a_next = 0.f;
  for (int t = T - 1; t >= 0; --t) {
    size_t k = (size_t)t * n + i;
    int d = done[k];
    bool ends = d || t == T - 1;
    float r = d ? 0.f : 1.f;
    float vs = v_start[k];
    float vn = ends ? v_end[k] : v_start[k + n];
    if (targets)
      targets[k] = r + gamma * vn;  // NOT masked at terminals (quirk 6)
    if (adv) {
      float vn_adv = d ? 0.f : vn;  // V[end] = 0 if the trajectory is frozen (230-236)
      float delta = r + gamma * vn_adv - vs;
      float a = delta + (ends ? 0.f : lambda * gamma * a_next);
      adv[k] = a;
      a_next = a;
    }
  }
}

// sgd (nn.h:622-625) / momentum (636-650) / adam (666-690: eps 1e-7 outside the sqrt, bias
// corrections c1 = 1 - beta1^t, c2 = 1 - beta2^t) on values: p = parameter, g = gradient, m / v = the
// element's optimizer state (momentum: m only).
__device__ __forceinline__ void opt_update_vals(int kind, float &p, float g, float &m, float &v, float lr, float wd,
                                                float beta1, float beta2, float c1, float c2) {
  if (kind == DFRL_OPT_SGD) {
    p = p * (1.f - wd) - g * lr;
  } else if (kind == DFRL_OPT_MOMENTUM) {
    m = m * 0.9f + g;
    p = p - m * lr;
  } else {
    m = m * beta1 + g * (1.f - beta1);
    v = v * beta2 + g * g * (1.f - beta2);
    float mu = m / c1, vu = v / c2;
    p = p - mu * lr / (sqrtf(vu) + 1e-7f);
  }
}
// The same on element i of the flat vectors (state = [m (n)][v (n)]).
__device__ __forceinline__ void opt_update(int kind, float *__restrict__ params,
                                           const float *__restrict__ grad,
                                           float *__restrict__ state, int n, int i, float lr,
                                           float wd, float beta1, float beta2, float c1, float c2) {
  float p = params[i], m = 0.f, v = 0.f;
  if (kind != DFRL_OPT_SGD)
    m = state[i];
  if (kind == DFRL_OPT_ADAM)
    v = state[n + i];
  opt_update_vals(kind, p, grad[i], m, v, lr, wd, beta1, beta2, c1, c2);
  params[i] = p;
  if (kind != DFRL_OPT_SGD)
    state[i] = m;
  if (kind == DFRL_OPT_ADAM)
    state[n + i] = v;
}
