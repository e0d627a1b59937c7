// env_dev.cuh -- device-side pieces of the batched bin-packing environment shared by env.cu,
// trainer.cu and fused.cu (reference apps/bin_packing/bin_packing.h:46-107).
#pragma once

#include "common.cuh"

struct env_params {
  int n, B, stride;
  int cap_w, cap_h;
  int iw0, ih0, iw1, ih1;  // shape1 (toss true), shape2
  uint32_t thr;            // shape1 iff philox word < thr
  uint64_t seed;
  int64_t env_offset;
  const uint8_t *tape;
  int tape_len;
};

// Item stream (bin_packing.h:73-81: bernoulli(0.4) ? shape1 : shape2): one Philox4x32-10 block per
// (group of 4 consecutive GLOBAL env ids, draw index), env g takes word g % 4 of the block of group
// g / 4 -- still a pure function of (seed, global env id, draw index), so results do not depend on how
// the environments are sharded over ranks, and a thread that steps 4 consecutive environments needs
// one block instead of four.
__device__ __forceinline__ philox4 item_block(const env_params &p, uint64_t global_group, uint32_t k) {
  return philox4x32_10(p.seed, global_group, k, DFRL_STREAM_ITEM);
}
__device__ __forceinline__ uint32_t block_word(const philox4 &r, int lane) {
  return lane == 0 ? r.x : lane == 1 ? r.y : lane == 2 ? r.z : r.w;
}
__device__ __forceinline__ int draw_shape1(const env_params &p, int i, uint32_t k) {
  if (p.tape)
    return p.tape[(size_t)i * p.tape_len + (k < (uint32_t)p.tape_len ? k : p.tape_len - 1)] != 0;
  const uint64_t g = (uint64_t)(p.env_offset + i);
  return block_word(item_block(p, g >> 2, k), (int)(g & 3)) < p.thr;
}


// environment::apply + agent::game_over + reset-on-done for env i on the global SoA planes
// (bin_packing.h:53-70, 94-106; rl.h:341-346). Returns done. Draws the next item.
__device__ __forceinline__ bool env_apply_global(const env_params &p, int8_t *__restrict__ state,
                                                 int i, int a, uint32_t draw_index) {
  const size_t S = p.stride;
  int iw = state[(size_t)(2 * p.B) * S + i], ih = state[(size_t)(2 * p.B + 1) * S + i];
  int bw = state[(size_t)(2 * a) * S + i] - iw;
  int bh = state[(size_t)(2 * a + 1) * S + i] - ih;
  bool over = bw < 0 || bh < 0;
  int s1 = draw_shape1(p, i, draw_index);
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
  return over;
}

static inline env_params make_params(const dfrl_env *e) {
  env_params p;
  p.n = e->n;
  p.B = e->B;
  p.stride = e->stride;
  p.cap_w = e->cfg.cap_w;
  p.cap_h = e->cfg.cap_h;
  p.iw0 = e->cfg.item_w[0];
  p.ih0 = e->cfg.item_h[0];
  p.iw1 = e->cfg.item_w[1];
  p.ih1 = e->cfg.item_h[1];
  double t = (double)e->cfg.p_shape1 * 4294967296.0;
  p.thr = t >= 4294967295.0 ? 0xffffffffu : (uint32_t)t;
  p.seed = e->cfg.seed;
  p.env_offset = e->cfg.env_offset;
  p.tape = e->tape;
  p.tape_len = e->tape_len;
  return p;
}


// Longest possible episode: every bin filled with the smallest item, plus the overflowing step.
static inline int env_max_episode_len(const dfrl_env_config &c) {
  long long per_bin = 1 << 20;
  int mw = c.item_w[0] < c.item_w[1] ? c.item_w[0] : c.item_w[1];
  int mh = c.item_h[0] < c.item_h[1] ? c.item_h[0] : c.item_h[1];
  if (mw > 0 && c.cap_w / mw < per_bin) per_bin = c.cap_w / mw;
  if (mh > 0 && c.cap_h / mh < per_bin) per_bin = c.cap_h / mh;
  if (per_bin >= (1 << 20)) return -1;  // zero-sized items never overflow
  return (int)(per_bin * c.n_bins + 1);
}
