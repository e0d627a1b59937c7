// env.cu -- K1: the batched bin-packing environment (reference apps/bin_packing/bin_packing.h).
//
// N independent episodes live in struct-of-arrays planes  state int8 [2B+2][stride]:
// plane 2b / 2b+1 = remaining (w, h) of bin b, planes 2B / 2B+1 = current item (w, h).
// One launch steps every environment: environment::apply (bin_packing.h:53-64),
// agent::game_over / get_reward (94-106), and the reset-on-done of agent::step
// (xylo/rl.h:341-346) happen on the device; the next item comes from a per-env tape (parity
// runs) or from Philox4x32-10 keyed by (seed, global env id, draw index).
//
// Roofline: HBM.  Algorithmic bytes per env-step = read + write of the 2B+2 state bytes, the
// action byte and the done byte = 2(2B+2) + 2 (= 38 B at B = 8).
// The vector kernel moves 4 environments per thread with 32-bit plane accesses so a warp
// touches 128 contiguous bytes per plane.
#include "common.cuh"
#include "env_dev.cuh"

namespace {

// reset(id) for every env (bin_packing.h:67-70) + constructor draw (50-52).
__global__ void env_reset_kernel(env_params p, int8_t *__restrict__ state,
                                 uint32_t *__restrict__ draws, uint32_t *__restrict__ steps) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n)
    return;
  for (int b = 0; b < p.B; ++b) {
    state[(size_t)(2 * b) * p.stride + i] = (int8_t)p.cap_w;
    state[(size_t)(2 * b + 1) * p.stride + i] = (int8_t)p.cap_h;
  }
  int s1 = draw_shape1(p, i, 0);
  state[(size_t)(2 * p.B) * p.stride + i] = (int8_t)(s1 ? p.iw0 : p.iw1);
  state[(size_t)(2 * p.B + 1) * p.stride + i] = (int8_t)(s1 ? p.ih0 : p.ih1);
  draws[i] = 1;
  steps[i] = 0;
}

// Scalar step: any B. One env per thread, touches only the planes it needs.
__global__ void env_step_scalar_kernel(env_params p, int8_t *__restrict__ state,
                                       uint32_t *__restrict__ draws, uint32_t *__restrict__ steps,
                                       const uint8_t *__restrict__ actions,
                                       uint8_t *__restrict__ done_out,
                                       int8_t *__restrict__ terminal) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n)
    return;
  const size_t S = p.stride;
  int a = actions[i];
  a = a < p.B ? a : p.B - 1;
  int iw = state[(size_t)(2 * p.B) * S + i], ih = state[(size_t)(2 * p.B + 1) * S + i];
  int bw = state[(size_t)(2 * a) * S + i] - iw;
  int bh = state[(size_t)(2 * a + 1) * S + i] - ih;
  bool over = bw < 0 || bh < 0;
  if (terminal) {
    for (int q = 0; q < 2 * p.B + 2; ++q)
      terminal[(size_t)q * S + i] = state[(size_t)q * S + i];
    terminal[(size_t)(2 * a) * S + i] = (int8_t)bw;
    terminal[(size_t)(2 * a + 1) * S + i] = (int8_t)bh;
  }
  if (done_out)
    done_out[i] = over;
  uint32_t k = draws[i];
  int s1 = draw_shape1(p, i, k);
  draws[i] = k + 1;
  steps[i] += 1;
  if (over) {
    for (int b = 0; b < p.B; ++b) {
      state[(size_t)(2 * b) * S + i] = (int8_t)p.cap_w;
      state[(size_t)(2 * b + 1) * S + i] = (int8_t)p.cap_h;
    }
  } else {
    state[(size_t)(2 * a) * S + i] = (int8_t)bw;
    state[(size_t)(2 * a + 1) * S + i] = (int8_t)bh;
  }
  state[(size_t)(2 * p.B) * S + i] = (int8_t)(s1 ? p.iw0 : p.iw1);
  state[(size_t)(2 * p.B + 1) * S + i] = (int8_t)(s1 ? p.ih0 : p.ih1);
}

// Vector step: B compile-time, 4 envs per thread, every plane read and written as one 32-bit
// word per thread (128 B per warp per plane, fully coalesced). The four environments of a word are
// stepped together with byte-wise SIMD integer instructions (per bin: select mask = action == b,
// subtract the masked item, collect the sign bits of the selected bin) and, on the Philox path,
// share one Philox block for their next items. TERM: also write the overflowed terminal states.
template <int B, bool TERM>
__global__ void __launch_bounds__(256, B == 8 ? 3 : 1)
env_step_vec4_kernel(env_params p, int8_t *__restrict__ state, uint32_t *__restrict__ draws,
                     uint32_t *__restrict__ steps, const uint8_t *__restrict__ actions,
                     uint8_t *__restrict__ done_out, int8_t *__restrict__ terminal) {
  constexpr int P = 2 * B + 2;
  int g = blockIdx.x * blockDim.x + threadIdx.x;  // group of 4 envs
  int i0 = g * 4;
  if (i0 >= p.n)
    return;
  const size_t S = p.stride;
  uint32_t w[P];
#pragma unroll
  for (int q = 0; q < P; ++q)
    w[q] = *reinterpret_cast<const uint32_t *>(state + (size_t)q * S + i0);
  uint32_t act4 = *reinterpret_cast<const uint32_t *>(actions + i0);  // actions padded to stride
  act4 = __vminu4(act4, (uint32_t)(B - 1) * 0x01010101u);
  const bool full = i0 + 3 < p.n;
  uint4 dr = make_uint4(0, 0, 0, 0), st = make_uint4(0, 0, 0, 0);
  if (full) {
    dr = *reinterpret_cast<const uint4 *>(draws + i0);
    st = *reinterpret_cast<const uint4 *>(steps + i0);
  } else {
    uint32_t *d = &dr.x, *s = &st.x;
    for (int e = 0; e < 4 && i0 + e < p.n; ++e) {
      d[e] = draws[i0 + e];
      s[e] = steps[i0 + e];
    }
  }
  // environment::apply (bin_packing.h:53-64): bin[a] -= item; over = the selected bin went negative
  const uint32_t iw4 = w[2 * B], ih4 = w[2 * B + 1];
  uint32_t neg = 0;
#pragma unroll
  for (int b = 0; b < B; ++b) {
    const uint32_t sel = __vcmpeq4(act4, (uint32_t)b * 0x01010101u);  // 0xff in the bytes whose action is b
    const uint32_t nw = __vsub4(w[2 * b], iw4 & sel), nh = __vsub4(w[2 * b + 1], ih4 & sel);
    neg |= (nw | nh) & sel;
    w[2 * b] = nw;
    w[2 * b + 1] = nh;
  }
  const uint32_t om = __vcmplts4(neg, 0u);  // 0xff in the bytes of the environments whose episode ended
  if (TERM) {  // the terminal state keeps its item (bin_packing.h:59-61)
#pragma unroll
    for (int q = 0; q < P; ++q)
      *reinterpret_cast<uint32_t *>(terminal + (size_t)q * S + i0) = w[q];
  }
  if (done_out)
    *reinterpret_cast<uint32_t *>(done_out + i0) = om & 0x01010101u;
  // reset-on-done (rl.h:341-346, bin_packing.h:67-70)
  const uint32_t cw4 = ((uint32_t)p.cap_w & 0xffu) * 0x01010101u, ch4 = ((uint32_t)p.cap_h & 0xffu) * 0x01010101u;
#pragma unroll
  for (int b = 0; b < B; ++b) {
    w[2 * b] = (w[2 * b] & ~om) | (cw4 & om);
    w[2 * b + 1] = (w[2 * b + 1] & ~om) | (ch4 & om);
  }
  // next items (drawn after every step, also after a reset: bin_packing.h:62, 69)
  uint32_t s1m = 0;  // 0xff in the bytes that draw shape 1
  const uint64_t g0 = (uint64_t)(p.env_offset + i0);
  uint32_t *drp = &dr.x, *stp = &st.x;
  if (!p.tape && full && (g0 & 3) == 0 && dr.x == dr.y && dr.x == dr.z && dr.x == dr.w) {
    const philox4 r = item_block(p, g0 >> 2, dr.x);
    s1m = (r.x < p.thr ? 0xffu : 0u) | (r.y < p.thr ? 0xff00u : 0u) | (r.z < p.thr ? 0xff0000u : 0u) |
          (r.w < p.thr ? 0xff000000u : 0u);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (i0 + e < p.n && draw_shape1(p, i0 + e, drp[e]))
        s1m |= 0xffu << (8 * e);
  }
  const uint32_t iw0 = ((uint32_t)p.iw0 & 0xffu) * 0x01010101u, iw1 = ((uint32_t)p.iw1 & 0xffu) * 0x01010101u;
  const uint32_t ih0 = ((uint32_t)p.ih0 & 0xffu) * 0x01010101u, ih1 = ((uint32_t)p.ih1 & 0xffu) * 0x01010101u;
  w[2 * B] = (iw0 & s1m) | (iw1 & ~s1m);
  w[2 * B + 1] = (ih0 & s1m) | (ih1 & ~s1m);
#pragma unroll
  for (int q = 0; q < P; ++q)
    *reinterpret_cast<uint32_t *>(state + (size_t)q * S + i0) = w[q];
  if (full) {
    dr.x += 1, dr.y += 1, dr.z += 1, dr.w += 1;
    st.x += 1, st.y += 1, st.z += 1, st.w += 1;
    *reinterpret_cast<uint4 *>(draws + i0) = dr;
    *reinterpret_cast<uint4 *>(steps + i0) = st;
  } else {
    for (int e = 0; e < 4 && i0 + e < p.n; ++e) {
      draws[i0 + e] = drp[e] + 1;
      steps[i0 + e] = stp[e] + 1;
    }
  }
}

// observation::to_vector (bin_packing.h:31-40): one thread per (row, bin) writes one float4.
__global__ void obs_encode_kernel(const int8_t *__restrict__ state, int rows, int stride, int B,
                                  float cap_w, float cap_h, float4 *__restrict__ obs) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)rows * B)
    return;
  int r = (int)(t / B), b = (int)(t % B);
  float4 o;
  // true division, as the reference does (exact for any capacity, not only powers of two)
  o.x = (float)state[(size_t)(2 * b) * stride + r] / cap_w;
  o.y = (float)state[(size_t)(2 * b + 1) * stride + r] / cap_h;
  o.z = (float)state[(size_t)(2 * B) * stride + r] / cap_w;
  o.w = (float)state[(size_t)(2 * B + 1) * stride + r] / cap_h;
  obs[t] = o;
}

// Heuristic policies. Scores follow firstfit_agent.cc:10-28, bestfit_agent.cc:10-30,
// minwaste_agent.cc:10-39, then argmax (first max). Random: rl.h:305-315 (uniform weights
// through discrete_distribution).
__device__ __forceinline__ int heuristic_action(const env_params &p, const int8_t *state, int i,
                                                int kind, uint32_t step) {
  const size_t S = p.stride;
  const int B = p.B;
  int iw = state[(size_t)(2 * B) * S + i], ih = state[(size_t)(2 * B + 1) * S + i];
  if (kind == DFRL_HEUR_RANDOM) {
    philox4 r = philox4x32_10(p.seed, (uint64_t)(p.env_offset + i), step, DFRL_STREAM_HEUR);
    double u = philox_u53(r.x, r.y);
    // discrete_distribution over B equal float weights: normalise in double, lower_bound
    float wf = (float)(1.0 / B);
    double sum = 0.0;
    for (int b = 0; b < B; ++b)
      sum += (double)wf;
    double c = 0.0;
    for (int b = 0; b < B; ++b) {
      c += (double)wf / sum;
      double cb = (b == B - 1) ? 1.0 : c;
      if (!(cb < u))
        return b;
    }
    return B - 1;
  }
  int best = 0;
  float best_score = 0.f;
  bool found_first = false;
  for (int b = 0; b < B; ++b) {
    int bw = state[(size_t)(2 * b) * S + i], bh = state[(size_t)(2 * b + 1) * S + i];
    bool fits = iw <= bw && ih <= bh;
    float score;
    if (kind == DFRL_HEUR_FIRSTFIT) {
      score = (fits && !found_first) ? 1.f : 0.f;
      found_first = found_first || fits;
    } else if (kind == DFRL_HEUR_BESTFIT) {
      score = fits ? (float)iw / (float)bw + (float)ih / (float)bh : -1.f;
    } else {
      if (!fits)
        score = -1.f;
      else {
        float r1 = (float)(bw - iw), r2 = (float)(bh - ih);
        score = ((r1 == (float)(p.cap_w / 2) && r2 == 0.f) || (r1 == 0.f && r2 == (float)(p.cap_h / 2))) ? 0.f : 1.f;
      }
    }
    if (b == 0 || best_score < score) {
      best = b;
      best_score = score;
    }
  }
  return best;
}

__global__ void heuristic_react_kernel(env_params p, const int8_t *__restrict__ state,
                                       const uint32_t *__restrict__ steps, int kind,
                                       uint8_t *__restrict__ actions) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n)
    return;
  actions[i] = (uint8_t)heuristic_action(p, state, i, kind, steps[i]);
}

// agent::play_one_episode x episodes for every env, entirely on the device (one thread owns one
// env for the whole game; no host round trips).  Rewards are summed per block then atomically.
__global__ void heuristic_play_kernel(env_params p, int8_t *__restrict__ state,
                                      uint32_t *__restrict__ draws, uint32_t *__restrict__ steps,
                                      int kind, int episodes, unsigned long long *__restrict__ acc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long reward = 0, nsteps = 0;
  if (i < p.n) {
    const size_t S = p.stride;
    const int B = p.B;
    uint32_t k = draws[i], s = steps[i];
    int ep = 0;
    while (ep < episodes) {
      int a = heuristic_action(p, state, i, kind, s);
      int iw = state[(size_t)(2 * B) * S + i], ih = state[(size_t)(2 * B + 1) * S + i];
      int bw = state[(size_t)(2 * a) * S + i] - iw, bh = state[(size_t)(2 * a + 1) * S + i] - ih;
      bool over = bw < 0 || bh < 0;
      ++nsteps;
      ++s;
      if (over) {
        ++ep;
        for (int b = 0; b < B; ++b) {
          state[(size_t)(2 * b) * S + i] = (int8_t)p.cap_w;
          state[(size_t)(2 * b + 1) * S + i] = (int8_t)p.cap_h;
        }
      } else {
        ++reward;
        state[(size_t)(2 * a) * S + i] = (int8_t)bw;
        state[(size_t)(2 * a + 1) * S + i] = (int8_t)bh;
      }
      int s1 = draw_shape1(p, i, k++);
      state[(size_t)(2 * B) * S + i] = (int8_t)(s1 ? p.iw0 : p.iw1);
      state[(size_t)(2 * B + 1) * S + i] = (int8_t)(s1 ? p.ih0 : p.ih1);
    }
    draws[i] = k;
    steps[i] = s;
  }
  // block reduction
  for (int o = 16; o > 0; o >>= 1) {
    reward += __shfl_down_sync(0xffffffffu, reward, o);
    nsteps += __shfl_down_sync(0xffffffffu, nsteps, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&acc[0], reward);
    atomicAdd(&acc[1], nsteps);
  }
}

// Per-id forms of the reference triple environment::apply / view / reset (bin_packing.h:53-70):
// one thread, one environment. apply() leaves an overflowed bin negative and draws no item
// (bin_packing.h:59-61); reset() refills every bin and draws the next item.
__global__ void env_apply_one_kernel(env_params p, int8_t *__restrict__ state, uint32_t *__restrict__ draws,
                                     uint32_t *__restrict__ steps, int i, int a) {
  const size_t S = p.stride;
  int iw = state[(size_t)(2 * p.B) * S + i], ih = state[(size_t)(2 * p.B + 1) * S + i];
  int bw = state[(size_t)(2 * a) * S + i] - iw, bh = state[(size_t)(2 * a + 1) * S + i] - ih;
  state[(size_t)(2 * a) * S + i] = (int8_t)bw;
  state[(size_t)(2 * a + 1) * S + i] = (int8_t)bh;
  steps[i] += 1;
  if (bw < 0 || bh < 0)
    return;
  int s1 = draw_shape1(p, i, draws[i]);
  draws[i] += 1;
  state[(size_t)(2 * p.B) * S + i] = (int8_t)(s1 ? p.iw0 : p.iw1);
  state[(size_t)(2 * p.B + 1) * S + i] = (int8_t)(s1 ? p.ih0 : p.ih1);
}
__global__ void env_reset_one_kernel(env_params p, int8_t *__restrict__ state, uint32_t *__restrict__ draws, int i) {
  const size_t S = p.stride;
  for (int b = 0; b < p.B; ++b) {
    state[(size_t)(2 * b) * S + i] = (int8_t)p.cap_w;
    state[(size_t)(2 * b + 1) * S + i] = (int8_t)p.cap_h;
  }
  int s1 = draw_shape1(p, i, draws[i]);
  draws[i] += 1;
  state[(size_t)(2 * p.B) * S + i] = (int8_t)(s1 ? p.iw0 : p.iw1);
  state[(size_t)(2 * p.B + 1) * S + i] = (int8_t)(s1 ? p.ih0 : p.ih1);
}

}  // namespace

extern "C" void dfrl_env_config_default(dfrl_env_config *cfg) {
  if (!cfg)
    return;
  cfg->n_envs = 0;
  cfg->n_bins = 8;   // bin_packing.h:12
  cfg->cap_w = 8;    // bin_packing.h:19
  cfg->cap_h = 8;
  cfg->item_w[0] = 4;  // shape1, bin_packing.h:73
  cfg->item_h[0] = 2;
  cfg->item_w[1] = 1;  // shape2, bin_packing.h:74
  cfg->item_h[1] = 2;
  cfg->p_shape1 = 0.4f;  // bin_packing.h:50
  cfg->seed = 1234;
  cfg->env_offset = 0;
}

extern "C" int dfrl_env_create(dfrl_ctx *ctx, const dfrl_env_config *cfg, dfrl_env **out) {
  DFRL_CHECK(ctx && cfg && out, "null argument");
  DFRL_CHECK(cfg->n_envs > 0, "n_envs must be positive");
  DFRL_CHECK(cfg->n_bins >= 1 && cfg->n_bins <= 64, "n_bins %d out of range 1..64", cfg->n_bins);
  DFRL_CHECK(cfg->cap_w > 0 && cfg->cap_w <= 127 && cfg->cap_h > 0 && cfg->cap_h <= 127,
             "capacity must fit int8");
  for (int k = 0; k < 2; ++k)
    DFRL_CHECK(cfg->item_w[k] >= 0 && cfg->item_w[k] <= 127 && cfg->item_h[k] >= 0 &&
                   cfg->item_h[k] <= 127, "item shape must fit int8");
  dfrl_env *e = new dfrl_env();
  e->ctx = ctx;
  e->cfg = *cfg;
  e->n = cfg->n_envs;
  e->B = cfg->n_bins;
  e->P = 2 * e->B + 2;
  e->stride = (int)round_up((size_t)e->n, 16);
  e->tape = nullptr;
  e->tape_len = 0;
  DFRL_CUDA(cudaMalloc(&e->state, (size_t)e->P * e->stride));
  DFRL_CUDA(cudaMemsetAsync(e->state, 0, (size_t)e->P * e->stride, ctx->stream));
  DFRL_CUDA(cudaMalloc(&e->draws, sizeof(uint32_t) * e->stride));
  DFRL_CUDA(cudaMalloc(&e->steps, sizeof(uint32_t) * e->stride));
  *out = e;
  return dfrl_env_reset(e);
}

extern "C" int dfrl_env_destroy(dfrl_env *e) {
  if (!e)
    return DFRL_OK;
  cudaStreamSynchronize(e->ctx->stream);
  cudaFree(e->state);
  cudaFree(e->draws);
  cudaFree(e->steps);
  if (e->tape)
    cudaFree(e->tape);
  delete e;
  return DFRL_OK;
}

extern "C" int dfrl_env_reset(dfrl_env *e) {
  DFRL_CHECK(e, "null env");
  env_params p = make_params(e);
  DFRL_LAUNCH(e->ctx, env_reset_kernel, ceil_div(e->n, 256), 256, 0, p, e->state, e->draws, e->steps);
  return DFRL_OK;
}

extern "C" int dfrl_env_load_item_tape(dfrl_env *e, const uint8_t *tape_host, int len) {
  DFRL_CHECK(e, "null env");
  DFRL_CUDA(cudaStreamSynchronize(e->ctx->stream));
  if (e->tape) {
    DFRL_CUDA(cudaFree(e->tape));
    e->tape = nullptr;
    e->tape_len = 0;
  }
  if (tape_host && len > 0) {
    DFRL_CUDA(cudaMalloc(&e->tape, (size_t)e->n * len));
    DFRL_CUDA(cudaMemcpy(e->tape, tape_host, (size_t)e->n * len, cudaMemcpyHostToDevice));
    e->tape_len = len;
  }
  return DFRL_OK;
}

int dfrl_env_step_internal(dfrl_env *e, const uint8_t *actions_dev, uint8_t *done_dev,
                           int8_t *terminal_dev, bool actions_padded) {
  env_params p = make_params(e);
  bool vec = actions_padded && (((uintptr_t)actions_dev) % 4 == 0) &&
             (!done_dev || ((uintptr_t)done_dev) % 4 == 0);
  int groups = ceil_div(e->n, 4);
#define DFRL_ENV_VEC4(BINS)                                                                                     \
  do {                                                                                                          \
    if (terminal_dev)                                                                                           \
      DFRL_LAUNCH(e->ctx, (env_step_vec4_kernel<BINS, true>), ceil_div(groups, 256), 256, 0, p, e->state, e->draws, \
                  e->steps, actions_dev, done_dev, terminal_dev);                                               \
    else                                                                                                        \
      DFRL_LAUNCH(e->ctx, (env_step_vec4_kernel<BINS, false>), ceil_div(groups, 256), 256, 0, p, e->state, e->draws, \
                  e->steps, actions_dev, done_dev, terminal_dev);                                               \
  } while (0)
  if (vec && e->B == 8) {
    DFRL_ENV_VEC4(8);
  } else if (vec && e->B == 16) {
    DFRL_ENV_VEC4(16);
  } else if (vec && e->B == 32) {
    DFRL_ENV_VEC4(32);
#undef DFRL_ENV_VEC4
  } else {
    DFRL_LAUNCH(e->ctx, env_step_scalar_kernel, ceil_div(e->n, 256), 256, 0, p, e->state, e->draws,
                e->steps, actions_dev, done_dev, terminal_dev);
  }
  return DFRL_OK;
}

extern "C" int dfrl_env_step(dfrl_env *e, const uint8_t *actions_dev, uint8_t *done_dev,
                             int8_t *terminal_state_dev) {
  DFRL_CHECK(e && actions_dev, "null argument");
  // Public buffers are [N] (done) / [2B+2][stride] (terminal). The vector kernel reads and
  // writes whole 4-env words, which is safe when N is a multiple of 4.
  return dfrl_env_step_internal(e, actions_dev, done_dev, terminal_state_dev, e->n % 4 == 0);
}

extern "C" int dfrl_env_apply_one(dfrl_env *e, int id, int action) {
  DFRL_CHECK(e, "null argument");
  DFRL_CHECK(id >= 0 && id < e->n, "environment id %d out of range", id);
  DFRL_CHECK(action >= 0 && action < e->B, "action %d out of range", action);
  env_params p = make_params(e);
  DFRL_LAUNCH(e->ctx, env_apply_one_kernel, 1, 1, 0, p, e->state, e->draws, e->steps, id, action);
  return DFRL_OK;
}

extern "C" int dfrl_env_reset_one(dfrl_env *e, int id) {
  DFRL_CHECK(e, "null argument");
  DFRL_CHECK(id >= 0 && id < e->n, "environment id %d out of range", id);
  env_params p = make_params(e);
  DFRL_LAUNCH(e->ctx, env_reset_one_kernel, 1, 1, 0, p, e->state, e->draws, id);
  return DFRL_OK;
}

extern "C" int dfrl_env_view_one(dfrl_env *e, int id, int8_t *state_host) {
  DFRL_CHECK(e && state_host, "null argument");
  DFRL_CHECK(id >= 0 && id < e->n, "environment id %d out of range", id);
  DFRL_CUDA(cudaMemcpy2DAsync(state_host, 1, e->state + id, e->stride, 1, e->P, cudaMemcpyDeviceToHost,
                              e->ctx->stream));
  DFRL_CUDA(cudaStreamSynchronize(e->ctx->stream));
  return DFRL_OK;
}

extern "C" int8_t *dfrl_env_state_dev(dfrl_env *e) { return e ? e->state : nullptr; }
extern "C" int dfrl_env_state_stride(dfrl_env *e) { return e ? e->stride : 0; }

extern "C" int dfrl_env_get_state(dfrl_env *e, int8_t *state_host) {
  DFRL_CHECK(e && state_host, "null argument");
  DFRL_CUDA(cudaMemcpy2DAsync(state_host, e->n, e->state, e->stride, e->n, e->P,
                              cudaMemcpyDeviceToHost, e->ctx->stream));
  DFRL_CUDA(cudaStreamSynchronize(e->ctx->stream));
  return DFRL_OK;
}
extern "C" int dfrl_env_set_state(dfrl_env *e, const int8_t *state_host) {
  DFRL_CHECK(e && state_host, "null argument");
  DFRL_CUDA(cudaMemcpy2DAsync(e->state, e->stride, state_host, e->n, e->n, e->P,
                              cudaMemcpyHostToDevice, e->ctx->stream));
  DFRL_CUDA(cudaStreamSynchronize(e->ctx->stream));
  return DFRL_OK;
}

extern "C" int dfrl_obs_encode(dfrl_ctx *ctx, const int8_t *state_dev, int rows, int stride,
                               int n_bins, int cap_w, int cap_h, float *obs_dev) {
  DFRL_CHECK(ctx && state_dev && obs_dev, "null argument");
  DFRL_CHECK(rows >= 0 && stride >= rows && n_bins >= 1 && cap_w > 0 && cap_h > 0, "bad shape");
  if (rows == 0)
    return DFRL_OK;
  long long total = (long long)rows * n_bins;
  DFRL_LAUNCH(ctx, obs_encode_kernel, ceil_div(total, 256), 256, 0, state_dev, rows, stride, n_bins,
              (float)cap_w, (float)cap_h, reinterpret_cast<float4 *>(obs_dev));
  return DFRL_OK;
}

extern "C" int dfrl_heuristic_react(dfrl_env *e, int kind, uint8_t *actions_dev) {
  DFRL_CHECK(e && actions_dev, "null argument");
  DFRL_CHECK(kind >= 0 && kind <= 3, "unknown heuristic %d", kind);
  env_params p = make_params(e);
  DFRL_LAUNCH(e->ctx, heuristic_react_kernel, ceil_div(e->n, 256), 256, 0, p, e->state, e->steps,
              kind, actions_dev);
  return DFRL_OK;
}

extern "C" int dfrl_heuristic_play(dfrl_env *e, int kind, int episodes, double *total_reward,
                                   long long *env_steps) {
  DFRL_CHECK(e, "null env");
  DFRL_CHECK(kind >= 0 && kind <= 3, "unknown heuristic %d", kind);
  DFRL_CHECK(episodes > 0, "episodes must be positive");
  env_params p = make_params(e);
  void *acc;
  DFRL_TRY(dfrl_scratch(e->ctx, 16, &acc));
  DFRL_CUDA(cudaMemsetAsync(acc, 0, 16, e->ctx->stream));
  DFRL_LAUNCH(e->ctx, heuristic_play_kernel, ceil_div(e->n, 128), 128, 0, p, e->state, e->draws,
              e->steps, kind, episodes, (unsigned long long *)acc);
  unsigned long long h[2];
  DFRL_CUDA(cudaMemcpyAsync(h, acc, 16, cudaMemcpyDeviceToHost, e->ctx->stream));
  DFRL_CUDA(cudaStreamSynchronize(e->ctx->stream));
  if (total_reward)
    *total_reward = (double)h[0];
  if (env_steps)
    *env_steps = (long long)h[1];
  return DFRL_OK;
}
