// rl_ops.cu -- K3 sampling, K4 returns/GAE, K5 loss gradients, K7 optimizers as standalone
// sm_100a kernels (the trainer's fused kernels in fused.cu inline the same device functions).
//
// All are HBM-bound streaming kernels: one thread per row / env / parameter, SoA inputs read
// coalesced, no shared memory needed (no reuse).
#include <math.h>

#include "common.cuh"
#include "device_fns.cuh"

namespace {

__global__ void sample_kernel(const float *__restrict__ probs, int rows, int cols,
                              const double *__restrict__ u, uint8_t *__restrict__ actions,
                              float *__restrict__ p_sel) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows)
    return;
  const float *p = probs + (size_t)r * cols;
  int a = discrete_sample(p, cols, u[r]);
  actions[r] = (uint8_t)a;
  if (p_sel)
    p_sel[r] = p[a < cols ? a : cols - 1];
}

__global__ void argmax_kernel(const float *__restrict__ probs, int rows, int cols,
                              uint8_t *__restrict__ actions) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows)
    return;
  actions[r] = (uint8_t)argmax_first(probs + (size_t)r * cols, cols);
}

// policy_gradient_learner::get_advantages (policy_gradient.h:125-147) without the baseline.
__global__ void returns_kernel(const uint8_t *__restrict__ done, const int *__restrict__ len, int n,
                               int L, float gamma, float *__restrict__ g, double *__restrict__ acc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double s0 = 0.0, cnt = 0.0;
  if (i < n) {
    int Li = len ? len[i] : L;
    int first = 0;
    while (first < Li) {
      int last = first;
      while (last < Li - 1 && !done[(size_t)last * n + i])
        ++last;
      int m = last - first + 1;
      float reward = 0.f;
      for (int k = 0; k < m; ++k) {
        float r = done[(size_t)(first + k) * n + i] ? 0.f : 1.f;
        reward = r + gamma * reward;                       // forward in time ...
        g[(size_t)(first + m - 1 - k) * n + i] = reward;   // ... written backward (quirk 4)
      }
      s0 += (double)g[(size_t)first * n + i];
      cnt += 1.0;
      first = last + 1;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_down_sync(0xffffffffu, s0, o);
    cnt += __shfl_down_sync(0xffffffffu, cnt, o);
  }
  if (acc && (threadIdx.x & 31) == 0 && cnt > 0.0) {
    atomicAdd(&acc[0], s0);
    atomicAdd(&acc[1], cnt);
  }
}

__global__ void subtract_baseline_kernel(float *__restrict__ g, const int *__restrict__ len, int n,
                                         int L, float baseline) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= (long long)n * L)
    return;
  int t = (int)(k / n), i = (int)(k % n);
  if (!len || t < len[i])
    g[k] -= baseline;
}

__global__ void gae_kernel(const uint8_t *__restrict__ done, const float *__restrict__ v_start,
                           const float *__restrict__ v_end, int n, int T, float gamma, float lambda,
                           float *__restrict__ targets, float *__restrict__ adv) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  gae_env(done, v_start, v_end, n, T, i, gamma, lambda, targets, adv);
}

__global__ void loss_grad_kernel(int kind, const float *__restrict__ probs,
                                 const uint8_t *__restrict__ actions, const float *__restrict__ adv,
                                 const float *__restrict__ p_old, float beta, int rows, int cols,
                                 float *__restrict__ out) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows)
    return;
  const float *p = probs + (size_t)r * cols;
  float *o = out + (size_t)r * cols;
  int a = actions[r];
  float A = adv[r];
  if (kind == DFRL_LOSS_CLIPPED) {
    float g = clipped_grad(p[a], p_old[r], A);
    for (int c = 0; c < cols; ++c)
      o[c] = (c == a) ? g : 0.f;
  } else {
    for (int c = 0; c < cols; ++c) {
      float v = p[c] * A - (c == a ? A : 0.f);
      if (kind == DFRL_LOSS_KL)
        v += (p[c] - p_old[(size_t)r * cols + c]) * beta;
      o[c] = v;
    }
  }
}

__global__ void square_loss_grad_kernel(const float *__restrict__ v, const float *__restrict__ tgt,
                                        int rows, float *__restrict__ out) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows)
    out[r] = v[r] - tgt[r];
}

__global__ void opt_kernel(int kind, float *__restrict__ params, const float *__restrict__ grad,
                           float *__restrict__ state, int n, float lr, float wd, float beta1,
                           float beta2, float c1, float c2) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  opt_update(kind, params, grad, state, n, i, lr, wd, beta1, beta2, c1, c2);
}

// K7 on large parameter vectors: HBM-bound (sgd 12 B / parameter, momentum 20, adam 28). Four
// parameters per thread as 16-byte accesses, two independent groups in flight per thread,
// grid-stride over a grid of a few CTAs per SM. n % 4 == 0 and 16-byte aligned pointers.
__global__ void __launch_bounds__(256) opt_vec4_kernel(int kind, float4 *__restrict__ params, const float4 *__restrict__ grad,
                                                       float4 *__restrict__ m4, float4 *__restrict__ v4, int n4, float lr,
                                                       float wd, float beta1, float beta2, float c1, float c2) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += 2 * stride) {
    const int j = i + stride;
    const bool two = j < n4;
    float4 p0 = params[i], g0 = __ldg(grad + i), p1, g1;
    float4 ma = make_float4(0, 0, 0, 0), va = ma, mb = ma, vb = ma;
    if (two)
      p1 = params[j], g1 = __ldg(grad + j);
    if (kind != DFRL_OPT_SGD) {
      ma = m4[i];
      if (two)
        mb = m4[j];
    }
    if (kind == DFRL_OPT_ADAM) {
      va = v4[i];
      if (two)
        vb = v4[j];
    }
    opt_update_vals(kind, p0.x, g0.x, ma.x, va.x, lr, wd, beta1, beta2, c1, c2);
    opt_update_vals(kind, p0.y, g0.y, ma.y, va.y, lr, wd, beta1, beta2, c1, c2);
    opt_update_vals(kind, p0.z, g0.z, ma.z, va.z, lr, wd, beta1, beta2, c1, c2);
    opt_update_vals(kind, p0.w, g0.w, ma.w, va.w, lr, wd, beta1, beta2, c1, c2);
    params[i] = p0;
    if (kind != DFRL_OPT_SGD)
      m4[i] = ma;
    if (kind == DFRL_OPT_ADAM)
      v4[i] = va;
    if (two) {
      opt_update_vals(kind, p1.x, g1.x, mb.x, vb.x, lr, wd, beta1, beta2, c1, c2);
      opt_update_vals(kind, p1.y, g1.y, mb.y, vb.y, lr, wd, beta1, beta2, c1, c2);
      opt_update_vals(kind, p1.z, g1.z, mb.z, vb.z, lr, wd, beta1, beta2, c1, c2);
      opt_update_vals(kind, p1.w, g1.w, mb.w, vb.w, lr, wd, beta1, beta2, c1, c2);
      params[j] = p1;
      if (kind != DFRL_OPT_SGD)
        m4[j] = mb;
      if (kind == DFRL_OPT_ADAM)
        v4[j] = vb;
    }
  }
}

// K4 on large batches: four environments per thread (16-byte value loads, 4-byte done loads);
// same recurrence as gae_env. n % 4 == 0.
__global__ void __launch_bounds__(256) gae_vec4_kernel(const uint8_t *__restrict__ done, const float *__restrict__ v_start,
                                                       const float *__restrict__ v_end, int n, int T, float gamma, float lambda,
                                                       float *__restrict__ targets, float *__restrict__ adv) {
  const int i = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (i >= n)
    return;
  float a_next[4] = {0.f, 0.f, 0.f, 0.f};
  float4 vs_next = make_float4(0, 0, 0, 0);  // v_start of step t + 1 (already loaded)
  for (int t = T - 1; t >= 0; --t) {
    const size_t k = (size_t)t * n + i;
    const uint32_t d4 = *reinterpret_cast<const uint32_t *>(done + k);
    const float4 vs4 = *reinterpret_cast<const float4 *>(v_start + k);
    const bool last = t == T - 1;
    float4 ve4 = make_float4(0, 0, 0, 0);
    if (last || d4)  // v_end is only defined / needed where a trajectory ends
      ve4 = *reinterpret_cast<const float4 *>(v_end + k);
    const float vs[4] = {vs4.x, vs4.y, vs4.z, vs4.w}, ve[4] = {ve4.x, ve4.y, ve4.z, ve4.w};
    const float vn1[4] = {vs_next.x, vs_next.y, vs_next.z, vs_next.w};
    float tg[4], ad[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int d = (d4 >> (8 * e)) & 0xff;
      const bool ends = d || last;
      const float r = d ? 0.f : 1.f;
      const float vn = ends ? ve[e] : vn1[e];
      tg[e] = r + gamma * vn;  // NOT masked at terminals (quirk 6)
      const float vn_adv = d ? 0.f : vn;
      const float delta = r + gamma * vn_adv - vs[e];
      ad[e] = delta + (ends ? 0.f : lambda * gamma * a_next[e]);
      a_next[e] = ad[e];
    }
    if (targets)
      *reinterpret_cast<float4 *>(targets + k) = make_float4(tg[0], tg[1], tg[2], tg[3]);
    if (adv)
      *reinterpret_cast<float4 *>(adv + k) = make_float4(ad[0], ad[1], ad[2], ad[3]);
    vs_next = vs4;
  }
}

}  // namespace

extern "C" int dfrl_sample(dfrl_ctx *ctx, const float *probs_dev, int rows, int cols,
                           const double *u_dev, uint8_t *actions_dev, float *p_sel_dev) {
  DFRL_CHECK(ctx && probs_dev && u_dev && actions_dev, "null argument");
  DFRL_CHECK(rows >= 0 && cols >= 1 && cols <= 256, "bad shape %d x %d", rows, cols);
  if (rows)
    DFRL_LAUNCH(ctx, sample_kernel, ceil_div(rows, 256), 256, 0, probs_dev, rows, cols, u_dev,
                actions_dev, p_sel_dev);
  return DFRL_OK;
}

extern "C" int dfrl_argmax(dfrl_ctx *ctx, const float *probs_dev, int rows, int cols,
                           uint8_t *actions_dev) {
  DFRL_CHECK(ctx && probs_dev && actions_dev, "null argument");
  DFRL_CHECK(rows >= 0 && cols >= 1 && cols <= 256, "bad shape %d x %d", rows, cols);
  if (rows)
    DFRL_LAUNCH(ctx, argmax_kernel, ceil_div(rows, 256), 256, 0, probs_dev, rows, cols, actions_dev);
  return DFRL_OK;
}

extern "C" int dfrl_returns(dfrl_ctx *ctx, const uint8_t *done_dev, const int *len_dev, int n_envs,
                            int max_len, float gamma, float *g_dev, double *baseline_acc_dev) {
  DFRL_CHECK(ctx && done_dev && g_dev, "null argument");
  DFRL_CHECK(n_envs > 0 && max_len > 0, "bad shape");
  DFRL_LAUNCH(ctx, returns_kernel, ceil_div(n_envs, 128), 128, 0, done_dev, len_dev, n_envs, max_len,
              gamma, g_dev, baseline_acc_dev);
  return DFRL_OK;
}

extern "C" int dfrl_subtract_baseline(dfrl_ctx *ctx, float *g_dev, const int *len_dev, int n_envs,
                                      int max_len, float baseline) {
  DFRL_CHECK(ctx && g_dev, "null argument");
  DFRL_CHECK(n_envs > 0 && max_len > 0, "bad shape");
  long long total = (long long)n_envs * max_len;
  DFRL_LAUNCH(ctx, subtract_baseline_kernel, ceil_div(total, 256), 256, 0, g_dev, len_dev, n_envs,
              max_len, baseline);
  return DFRL_OK;
}

extern "C" int dfrl_gae(dfrl_ctx *ctx, const uint8_t *done_dev, const float *v_start_dev,
                        const float *v_end_dev, int n_envs, int T, float gamma, float lambda,
                        float *targets_dev, float *adv_dev) {
  DFRL_CHECK(ctx && done_dev && v_start_dev && v_end_dev, "null argument");
  DFRL_CHECK(n_envs > 0 && T > 0, "bad shape");
  auto al16 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (n_envs % 4 == 0 && n_envs >= (1 << 14) && al16(v_start_dev) && al16(v_end_dev) && al16(targets_dev) && al16(adv_dev) &&
      (reinterpret_cast<uintptr_t>(done_dev) & 3) == 0)
    DFRL_LAUNCH(ctx, gae_vec4_kernel, ceil_div(n_envs / 4, 256), 256, 0, done_dev, v_start_dev, v_end_dev, n_envs, T, gamma,
                lambda, targets_dev, adv_dev);
  else
    DFRL_LAUNCH(ctx, gae_kernel, ceil_div(n_envs, 128), 128, 0, done_dev, v_start_dev, v_end_dev,
                n_envs, T, gamma, lambda, targets_dev, adv_dev);
  return DFRL_OK;
}

extern "C" int dfrl_loss_grad(dfrl_ctx *ctx, int kind, const float *probs_dev,
                              const uint8_t *actions_dev, const float *adv_dev,
                              const float *p_old_dev, float beta, int rows, int cols,
                              float *out_dev) {
  DFRL_CHECK(ctx && probs_dev && actions_dev && adv_dev && out_dev, "null argument");
  DFRL_CHECK(kind >= 0 && kind <= 2, "unknown loss kind %d", kind);
  DFRL_CHECK(kind == DFRL_LOSS_SOFTMAX_LOG || p_old_dev, "p_old required");
  DFRL_CHECK(rows >= 0 && cols >= 1, "bad shape");
  if (rows)
    DFRL_LAUNCH(ctx, loss_grad_kernel, ceil_div(rows, 256), 256, 0, kind, probs_dev, actions_dev,
                adv_dev, p_old_dev, beta, rows, cols, out_dev);
  return DFRL_OK;
}

extern "C" int dfrl_square_loss_grad(dfrl_ctx *ctx, const float *v_dev, const float *target_dev,
                                     int rows, float *out_dev) {
  DFRL_CHECK(ctx && v_dev && target_dev && out_dev, "null argument");
  if (rows > 0)
    DFRL_LAUNCH(ctx, square_loss_grad_kernel, ceil_div(rows, 256), 256, 0, v_dev, target_dev, rows,
                out_dev);
  return DFRL_OK;
}

extern "C" int dfrl_opt_step(dfrl_ctx *ctx, int kind, float *params_dev, const float *grad_dev,
                             float *state_dev, int n, float lr, float weight_decay, float beta1,
                             float beta2, float adam_t) {
  DFRL_CHECK(ctx && params_dev && grad_dev, "null argument");
  DFRL_CHECK(kind >= 0 && kind <= 2, "unknown optimizer %d", kind);
  DFRL_CHECK(kind == DFRL_OPT_SGD || state_dev, "optimizer state required");
  if (n <= 0)
    return DFRL_OK;
  // bias corrections as the reference computes them on the host: 1 - powf(beta, t) (nn.h:683-684)
  float c1 = 1.f - powf(beta1, adam_t), c2 = 1.f - powf(beta2, adam_t);
  auto al16 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (n % 4 == 0 && n >= (1 << 16) && al16(params_dev) && al16(grad_dev) && al16(state_dev)) {
    const int n4 = n / 4;
    int grid = ceil_div(n4, 2 * 256);
    if (grid > 8 * ctx->sm_count)
      grid = 8 * ctx->sm_count;
    DFRL_LAUNCH(ctx, opt_vec4_kernel, grid, 256, 0, kind, reinterpret_cast<float4 *>(params_dev),
                reinterpret_cast<const float4 *>(grad_dev), reinterpret_cast<float4 *>(state_dev),
                reinterpret_cast<float4 *>(state_dev ? state_dev + n : nullptr), n4, lr, weight_decay, beta1, beta2, c1, c2);
  } else {
    DFRL_LAUNCH(ctx, opt_kernel, ceil_div(n, 256), 256, 0, kind, params_dev, grad_dev, state_dev, n,
                lr, weight_decay, beta1, beta2, c1, c2);
  }
  return DFRL_OK;
}
