// fused.cu -- fused small-MLP kernels on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// For 3-dense-layer nets  D0 -> D1 -relu-> D2 -relu-> D3  (C2/C4: 32-64-64-{8,1}) one CTA owns a
// tile of 128 learner rows and runs the WHOLE optimizer::step body for them without touching HBM
// for activations:
//   forward  H1 = relu(X0 W1^T + b1), H2 = relu(H1 W2^T + b2), out = H2 W3^T + b3     (3 UMMA GEMMs)
//   loss gradient at the output (PPO clipped surrogate / policy_loss / square loss), softmax
//   Jacobian                                                                         (registers)
//   input gradients dH2 = (dY W3) . relu', dH1 = (dH2 W2) . relu'                     (2 UMMA GEMMs)
//   weight gradients dW_l += dY_l^T X_l, accumulated over all tiles of the CTA in TMEM (3 UMMA GEMMs)
// Activations live in shared memory as SWIZZLE_128B bf16 panels (umma.cuh): the same panel is
// the K-major A operand of the forward GEMM and the MN-major operand of the weight-gradient GEMM.
// FP32-level accuracy: bf16 hi/lo split operands, three products per GEMM, FP32 accumulation.
//
// Row tiles are env-blocked: a tile holds E = 128 / T environments x all T steps (row = t * E + e)
// so that the critic target r + gamma V(s_{t+1}) and GAE only need values of the same tile.
//
// Reference semantics reproduced: rl.h:54-74 (clipped_gradient), rl.h:45-52, nn.h:393-417 (softmax
// Jacobian), nn.h:85-100 (dW = SUM over rows), policy_gradient.h:196-281 (targets, GAE).
#include <math.h>
#include <string.h>

#include "device_fns.cuh"
#include "env_dev.cuh"
#include "trainer.h"
#include "umma.cuh"

namespace {

constexpr int TILE = 128;
constexpr uint32_t PANEL = 128 * 128;  // bytes of one 128-row panel

struct net3 {
  int d0, d1, d2, d3;
  int o_w1, o_b1, o_w2, o_b2, o_w3, o_b3;  // offsets into the flat parameter vector
  int n_params;
};

struct tid_t {
  int wg, w, lane, row;
  uint32_t lane_base;  // TMEM lane field of this thread's warp
};
__device__ __forceinline__ tid_t thread_id() {
  tid_t t;
  t.wg = threadIdx.x >> 7;
  t.w = (threadIdx.x >> 5) & 3;
  t.lane = threadIdx.x & 31;
  t.row = t.w * 32 + t.lane;
  t.lane_base = (uint32_t)(t.w * 32) << 16;
  return t;
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float *v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i)
    v[i] = __uint_as_float(r[i]);
}
template <int DC>
__device__ __forceinline__ void tmem_load(uint32_t taddr, float (&v)[DC]) {
  static_assert(DC % 8 == 0, "column count per thread must be a multiple of 8");
#pragma unroll
  for (int j = 0; j < DC; j += 8)
    tmem_ld8(taddr + j, &v[j]);
  umma::tmem_ld_wait();
}

// fp32 [N][K] row-major (global) -> hi / lo panels [rows_alloc][64], zero padded.
__device__ void stage_weight(const float *__restrict__ W, int N, int K, int rows_alloc, uint8_t *hi,
                             uint8_t *lo) {
  for (int c = threadIdx.x; c < rows_alloc * 8; c += blockDim.x) {
    int row = c >> 3, chunk = c & 7;
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int col = chunk * 8 + j;
      x[j] = (row < N && col < K) ? W[(size_t)row * K + col] : 0.f;
    }
    uint4 h, l;
    umma::split8(x, h, l);
    uint32_t off = umma::panel_chunk_off(row, chunk);
    *reinterpret_cast<uint4 *>(hi + off) = h;
    *reinterpret_cast<uint4 *>(lo + off) = l;
  }
}

__device__ void zero_bytes(uint8_t *p, uint32_t bytes) {
  for (uint32_t o = threadIdx.x * 16; o < bytes; o += blockDim.x * 16)
    *reinterpret_cast<uint4 *>(p + o) = make_uint4(0, 0, 0, 0);
}

// One GEMM = ksteps x {hi.hi, hi.lo, lo.hi} tcgen05.mma instructions, issued by one thread.
// a_lo / b_lo == 0: that operand is exact in bf16 (no lo pass).
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi,
                                           uint32_t b_lo, int ksteps, bool a_mn, bool b_mn,
                                           uint32_t idesc, bool accumulate) {
  uint32_t acc = accumulate ? 1u : 0u;
  const uint32_t a_step = a_mn ? umma::KSTEP_BYTES_MNMAJOR : umma::KSTEP_BYTES_KMAJOR;
  const uint32_t b_step = b_mn ? umma::KSTEP_BYTES_MNMAJOR : umma::KSTEP_BYTES_KMAJOR;
  const uint32_t a_lbo = a_mn ? PANEL : 16, b_lbo = b_mn ? PANEL : 16;
  for (int k = 0; k < ksteps; ++k) {
    uint64_t ah = umma::make_desc_sw128(a_hi + k * a_step, a_lbo, 1024);
    uint64_t bh = umma::make_desc_sw128(b_hi + k * b_step, b_lbo, 1024);
    umma::mma_bf16(tmem_d, ah, bh, idesc, acc);
    acc = 1;
    if (b_lo) {
      uint64_t bl = umma::make_desc_sw128(b_lo + k * b_step, b_lbo, 1024);
      umma::mma_bf16(tmem_d, ah, bl, idesc, 1);
    }
    if (a_lo) {
      uint64_t al = umma::make_desc_sw128(a_lo + k * a_step, a_lbo, 1024);
      umma::mma_bf16(tmem_d, al, bh, idesc, 1);
    }
  }
}

// Shared-memory map of the learner kernels (all panel buffers 1024-byte aligned).
template <int D1, int D2>
struct smem_map {
  static constexpr uint32_t W1_HI = 0, W1_LO = W1_HI + D1 * 128;
  static constexpr uint32_t W2_HI = W1_LO + D1 * 128, W2_LO = W2_HI + D2 * 128;
  static constexpr uint32_t W3_HI = W2_LO + D2 * 128, W3_LO = W3_HI + 16 * 128;
  static constexpr uint32_t X0 = W3_LO + 16 * 128;          // hi only (k/8 is exact in bf16)
  static constexpr uint32_t H_HI = X0 + PANEL;              // [H1 | H2]
  static constexpr uint32_t H_LO = H_HI + 2 * PANEL;
  static constexpr uint32_t DH_HI = H_LO + 2 * PANEL;       // [dH1 | dH2]
  static constexpr uint32_t DH_LO = DH_HI + 2 * PANEL;
  static constexpr uint32_t DY_HI = DH_LO + 2 * PANEL;
  static constexpr uint32_t DY_LO = DY_HI + PANEL;
  static constexpr uint32_t FLOATS = DY_LO + PANEL;         // biases, w3 (fp32), scratch
  static constexpr uint32_t N_FLOATS = D1 + D2 + 16 + 64 + 128 * 10;
  static constexpr uint32_t BARS = FLOATS + N_FLOATS * 4;
  static constexpr uint32_t TOTAL = BARS + 64;
  static_assert((D1 * 128) % 1024 == 0 && (D2 * 128) % 1024 == 0, "panel alignment");
};

// TMEM column map
constexpr uint32_t TC_L1 = 0, TC_L2 = 64, TC_L3 = 128, TC_DH2 = 160, TC_DH1 = 224, TC_DA = 288,
                   TC_DB = 352, TC_DC = 416, TC_COLS = 512;

// observation::to_vector of learner row r of a tile into the X0 panel (bin_packing.h:31-40).
// One task = one 16-byte chunk = the 8 floats of two bins.
__device__ __forceinline__ uint4 obs_chunk(const int8_t *__restrict__ st, int stride, int i, int B,
                                           int chunk, float inv_w, float inv_h, int over_bin, bool valid) {
  if (!valid)
    return make_uint4(0, 0, 0, 0);
  int iw = st[(size_t)(2 * B) * stride + i], ih = st[(size_t)(2 * B + 1) * stride + i];
  uint32_t out[4];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    int b = 2 * chunk + q;
    int bw = st[(size_t)(2 * b) * stride + i], bh = st[(size_t)(2 * b + 1) * stride + i];
    if (b == over_bin) {  // terminal state of a done step: bin[a] -= item (bin_packing.h:54-61)
      bw -= iw;
      bh -= ih;
    }
    __nv_bfloat16 x0 = __float2bfloat16_rn((float)bw * inv_w), x1 = __float2bfloat16_rn((float)bh * inv_h);
    __nv_bfloat16 x2 = __float2bfloat16_rn((float)iw * inv_w), x3 = __float2bfloat16_rn((float)ih * inv_h);
    out[2 * q] = (uint32_t)__bfloat16_as_ushort(x0) | ((uint32_t)__bfloat16_as_ushort(x1) << 16);
    out[2 * q + 1] = (uint32_t)__bfloat16_as_ushort(x2) | ((uint32_t)__bfloat16_as_ushort(x3) << 16);
  }
  return make_uint4(out[0], out[1], out[2], out[3]);
}

// TMEM accumulator -> (+bias, relu) -> bf16 hi/lo panel. Returns the relu mask of this thread's
// columns (bit j = column col0 + j was > 0).
template <int D>
__device__ __forceinline__ uint32_t epi_hidden_fwd(uint32_t tmem_acc, const tid_t &t,
                                                   const float *__restrict__ bias, uint8_t *hi,
                                                   uint8_t *lo) {
  constexpr int DC = D / 2;
  const int col0 = t.wg * DC;
  float v[DC];
  tmem_load<DC>(tmem_acc + t.lane_base + col0, v);
  uint32_t mask = 0;
#pragma unroll
  for (int j = 0; j < DC; ++j) {
    float x = v[j] + bias[col0 + j];
    if (x > 0.f)
      mask |= 1u << j;
    else
      x = 0.f;
    v[j] = x;
  }
#pragma unroll
  for (int c = 0; c < DC / 8; ++c) {
    uint4 h, l;
    umma::split8(&v[8 * c], h, l);
    uint32_t off = umma::panel_chunk_off(t.row, col0 / 8 + c);
    *reinterpret_cast<uint4 *>(hi + off) = h;
    *reinterpret_cast<uint4 *>(lo + off) = l;
  }
  return mask;
}
// TMEM accumulator -> (. relu mask) -> bf16 hi/lo panel (input-gradient epilogue).
template <int D>
__device__ __forceinline__ void epi_hidden_bwd(uint32_t tmem_acc, const tid_t &t, uint32_t mask,
                                               uint8_t *hi, uint8_t *lo) {
  constexpr int DC = D / 2;
  const int col0 = t.wg * DC;
  float v[DC];
  tmem_load<DC>(tmem_acc + t.lane_base + col0, v);
#pragma unroll
  for (int j = 0; j < DC; ++j)
    v[j] = (mask >> j) & 1u ? v[j] : 0.f;
#pragma unroll
  for (int c = 0; c < DC / 8; ++c) {
    uint4 h, l;
    umma::split8(&v[8 * c], h, l);
    uint32_t off = umma::panel_chunk_off(t.row, col0 / 8 + c);
    *reinterpret_cast<uint4 *>(hi + off) = h;
    *reinterpret_cast<uint4 *>(lo + off) = l;
  }
}

__device__ __forceinline__ void sync_after_smem_writes() {
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
}


// Per-CTA context of the fused kernels.
struct tile_ctx {
  uint8_t *smem;
  uint32_t sbase, tmem;
  uint64_t *bar;
  uint32_t phase;
  tid_t t;
  __device__ __forceinline__ void wait() {
    umma::mbar_wait(bar, phase);
    phase ^= 1;
    umma::fence_after_sync();
  }
};

// As epi_hidden_fwd, but also hands the fp32 activations of this thread's columns back.
template <int D, bool WRITE_PANEL>
__device__ __forceinline__ uint32_t epi_hidden_fwd_keep(uint32_t tmem_acc, const tid_t &t,
                                                        const float *__restrict__ bias, uint8_t *hi,
                                                        uint8_t *lo, float (&v)[D / 2]) {
  constexpr int DC = D / 2;
  const int col0 = t.wg * DC;
  tmem_load<DC>(tmem_acc + t.lane_base + col0, v);
  uint32_t mask = 0;
#pragma unroll
  for (int j = 0; j < DC; ++j) {
    float x = v[j] + bias[col0 + j];
    if (x > 0.f)
      mask |= 1u << j;
    else
      x = 0.f;
    v[j] = x;
  }
  if (WRITE_PANEL) {
#pragma unroll
    for (int c = 0; c < DC / 8; ++c) {
      uint4 h, l;
      umma::split8(&v[8 * c], h, l);
      uint32_t off = umma::panel_chunk_off(t.row, col0 / 8 + c);
      *reinterpret_cast<uint4 *>(hi + off) = h;
      *reinterpret_cast<uint4 *>(lo + off) = l;
    }
  }
  return mask;
}

// Layers 1 and 2 of a tile whose X0 panel is staged and synchronised. Leaves H1 (and, if
// WRITE_H2, H2) panels written but NOT yet synchronised.
template <int D0, int D1, int D2, bool WRITE_H2>
__device__ __forceinline__ void fwd_hidden(tile_ctx &c, const float *b1, const float *b2,
                                           uint32_t &mask1, uint32_t &mask2, float (&h2)[D2 / 2]) {
  using SM = smem_map<D1, D2>;
  constexpr uint32_t ID_L1 = umma::make_idesc_bf16(128, D1, 0, 0);
  constexpr uint32_t ID_L2 = umma::make_idesc_bf16(128, D2, 0, 0);
  if (threadIdx.x == 0) {
    issue_gemm(c.tmem + TC_L1, c.sbase + SM::X0, 0, c.sbase + SM::W1_HI, c.sbase + SM::W1_LO, D0 / 16,
               false, false, ID_L1, false);
    umma::commit(c.bar);
  }
  c.wait();
  mask1 = epi_hidden_fwd<D1>(c.tmem + TC_L1, c.t, b1, c.smem + SM::H_HI, c.smem + SM::H_LO);
  sync_after_smem_writes();
  if (threadIdx.x == 0) {
    issue_gemm(c.tmem + TC_L2, c.sbase + SM::H_HI, c.sbase + SM::H_LO, c.sbase + SM::W2_HI,
               c.sbase + SM::W2_LO, D1 / 16, false, false, ID_L2, false);
    umma::commit(c.bar);
  }
  c.wait();
  mask2 = epi_hidden_fwd_keep<D2, WRITE_H2>(c.tmem + TC_L2, c.t, b2, c.smem + SM::H_HI + PANEL,
                                            c.smem + SM::H_LO + PANEL, h2);
}

// Value head (D2 -> 1) in registers: each row is held by two threads (one per warpgroup, D2/2
// columns each); partial dot products meet in shared memory. Contains one __syncthreads.
template <int D2>
__device__ __forceinline__ float value_head(const tid_t &t, const float (&h2)[D2 / 2],
                                            const float *__restrict__ w3, float b3, float *vpart) {
  constexpr int DC = D2 / 2;
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < DC; ++j)
    s = fmaf(h2[j], w3[t.wg * DC + j], s);
  vpart[t.wg * TILE + t.row] = s;
  __syncthreads();
  return vpart[t.row] + vpart[TILE + t.row] + b3;
}

// Common one-time setup of the learner / rollout kernels.
template <int D0, int D1, int D2>
__device__ __forceinline__ void setup_common(tile_ctx &c, uint8_t *smem, const float *params, const net3 &net,
                                             float *b1, float *b2, uint32_t *tmem_slot, uint64_t *bar) {
  using SM = smem_map<D1, D2>;
  c.smem = smem;
  c.sbase = umma::smem_u32(smem);
  c.bar = bar;
  c.phase = 0;
  c.t = thread_id();
  if (threadIdx.x < 32)
    umma::tmem_alloc(tmem_slot, TC_COLS);
  if (threadIdx.x == 0) {
    umma::mbar_init(bar, 1);
    umma::fence_mbar_init();
  }
  stage_weight(params + net.o_w1, D1, D0, D1, smem + SM::W1_HI, smem + SM::W1_LO);
  stage_weight(params + net.o_w2, D2, D1, D2, smem + SM::W2_HI, smem + SM::W2_LO);
  for (int i = threadIdx.x; i < D1; i += blockDim.x) b1[i] = params[net.o_b1 + i];
  for (int i = threadIdx.x; i < D2; i += blockDim.x) b2[i] = params[net.o_b2 + i];
  zero_bytes(smem + SM::X0, PANEL);
  zero_bytes(smem + SM::H_HI, 4 * PANEL);
  zero_bytes(smem + SM::DH_HI, 4 * PANEL);
  zero_bytes(smem + SM::DY_HI, 2 * PANEL);
  __syncthreads();
  if (threadIdx.x < TILE)
    *reinterpret_cast<uint16_t *>(smem + SM::X0 + umma::panel_off(threadIdx.x, D0)) = 0x3F80;  // bf16 1.0
}

// Stage the X0 panel of a learner tile (rows = t * E + e). end_rows: observation of the END state
// of step (t, e): overflowed terminal state when done, live env state at the rollout's last step,
// zeros (unused) otherwise. Global loads happen before `pre_store` (a wait) runs.
struct learner_rows {
  const int8_t *rec_state, *live_state;
  const uint8_t *rec_action, *rec_done;
  int n, stride, T, E, B;
  float inv_w, inv_h;
};
template <typename F>
__device__ __forceinline__ void stage_x0(uint8_t *x0, const learner_rows &L, int tile, bool end_rows,
                                         F pre_store) {
  const int P = 2 * L.B + 2, cpr = L.B / 2;
  uint4 xc[2];
  int xrow[2], xchunk[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    int task = threadIdx.x + q * 256;
    int r = task / cpr, c = task % cpr;
    bool in = task < TILE * cpr;
    int tt = r / L.E, e = r % L.E;
    int i = tile * L.E + e;
    bool valid = in && tt < L.T && i < L.n;
    const int8_t *src = L.rec_state + (size_t)tt * P * L.stride;
    int over_bin = -1;
    if (end_rows && valid) {
      size_t k = (size_t)tt * L.n + i;
      if (L.rec_done[k])
        over_bin = L.rec_action[k];
      else if (tt == L.T - 1)
        src = L.live_state;
      else
        valid = false;
    }
    xc[q] = obs_chunk(src, L.stride, i, L.B, c, L.inv_w, L.inv_h, over_bin, valid);
    xrow[q] = in ? r : -1;
    xchunk[q] = c;
  }
  pre_store();
#pragma unroll
  for (int q = 0; q < 2; ++q)
    if (xrow[q] >= 0)
      *reinterpret_cast<uint4 *>(x0 + umma::panel_chunk_off(xrow[q], xchunk[q])) = xc[q];
}

struct critic_args {
  const float *params;
  net3 net;
  learner_rows rows;
  int n_tiles;
  float gamma, lambda;
  float *targets_out;  // [T][n] (critic step) -- introspection + parity
  float *adv_out;      // [T][n] (GAE kernel)
  float *partials;
};

// ---------------------------------------------------------------------------------------------
// update_value_model (policy_gradient.h:196-218) minus the optimizer update: V on start and end
// rows with the current critic, targets r + gamma V_next (unmasked), dY = V - target, backward,
// dW partials.
template <int D0, int D1, int D2>
__global__ void __launch_bounds__(256, 1) fused_critic_step_kernel(critic_args a) {
  using SM = smem_map<D1, D2>;
  constexpr int DC2 = D2 / 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float *fl = reinterpret_cast<float *>(smem + SM::FLOATS);
  float *b1 = fl, *b2 = fl + D1, *w3 = fl + D1 + D2 + 16, *vpart = w3 + 64, *ve = vpart + 2 * TILE,
        *vs = ve + TILE, *dys = vs + TILE;
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + SM::BARS);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + SM::BARS + 16);
  const net3 net = a.net;
  tile_ctx c;
  setup_common<D0, D1, D2>(c, smem, a.params, net, b1, b2, tmem_slot, bar);
  for (int i = threadIdx.x; i < D2; i += blockDim.x) w3[i] = a.params[net.o_w3 + i];
  const float b3 = a.params[net.o_b3];
  sync_after_smem_writes();
  c.tmem = *tmem_slot;
  const tid_t t = c.t;
  const learner_rows &L = a.rows;

  constexpr uint32_t ID_DH1 = umma::make_idesc_bf16(128, D1, 0, 1);
  constexpr uint32_t ID_DA = umma::make_idesc_bf16(128, 64, 1, 1);
  constexpr uint32_t ID_DB = umma::make_idesc_bf16(128, D0 + 16, 1, 1);

  bool dw_pending = false, first_tile = true;
  float dw3[DC2];
#pragma unroll
  for (int j = 0; j < DC2; ++j)
    dw3[j] = 0.f;
  float db3 = 0.f;

  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    uint32_t m1, m2;
    float h2[DC2];
    // ---- pass 1: V of the end rows
    stage_x0(smem + SM::X0, L, tile, true, [&]() {
      if (dw_pending) {
        c.wait();
        dw_pending = false;
      }
    });
    sync_after_smem_writes();
    fwd_hidden<D0, D1, D2, false>(c, b1, b2, m1, m2, h2);
    float v_end = value_head<D2>(t, h2, w3, b3, vpart);
    if (t.wg == 0)
      ve[t.row] = v_end;
    __syncthreads();
    // ---- pass 2: start rows, activations kept
    stage_x0(smem + SM::X0, L, tile, false, []() {});
    sync_after_smem_writes();
    fwd_hidden<D0, D1, D2, true>(c, b1, b2, m1, m2, h2);
    float v = value_head<D2>(t, h2, w3, b3, vpart);
    if (t.wg == 0)
      vs[t.row] = v;
    __syncthreads();
    // ---- targets and dY = V - target (square_loss_grad, nn.h:548-550)
    {
      int tt = t.row / L.E, e = t.row % L.E;
      int i = tile * L.E + e;
      bool valid = tt < L.T && i < L.n;
      float dy = 0.f;
      if (valid) {
        size_t k = (size_t)tt * L.n + i;
        int d = L.rec_done[k];
        bool ends = d || tt == L.T - 1;
        float vn = ends ? ve[t.row] : vs[t.row + L.E];
        float tgt = (d ? 0.f : 1.f) + a.gamma * vn;  // not masked at terminals (quirk 6)
        dy = v - tgt;
        if (t.wg == 0 && a.targets_out)
          a.targets_out[k] = tgt;
      }
      // dH2 = dY w3 . relu'(H2) (rank-1: no GEMM); dW3 += dY H2; db3 += dY
      float g[DC2];
#pragma unroll
      for (int j = 0; j < DC2; ++j) {
        g[j] = (m2 >> j) & 1u ? dy * w3[t.wg * DC2 + j] : 0.f;
        dw3[j] = fmaf(dy, h2[j], dw3[j]);
      }
      if (t.wg == 0)
        db3 += dy;
#pragma unroll
      for (int cc = 0; cc < DC2 / 8; ++cc) {
        uint4 h, l;
        umma::split8(&g[8 * cc], h, l);
        uint32_t off = umma::panel_chunk_off(t.row, (t.wg * DC2) / 8 + cc);
        *reinterpret_cast<uint4 *>(smem + SM::DH_HI + PANEL + off) = h;
        *reinterpret_cast<uint4 *>(smem + SM::DH_LO + PANEL + off) = l;
      }
    }
    sync_after_smem_writes();
    // ---- dH1 = dH2 . W2, relu mask
    if (threadIdx.x == 0) {
      issue_gemm(c.tmem + TC_DH1, c.sbase + SM::DH_HI + PANEL, c.sbase + SM::DH_LO + PANEL,
                 c.sbase + SM::W2_HI, c.sbase + SM::W2_LO, D2 / 16, false, true, ID_DH1, false);
      umma::commit(c.bar);
    }
    c.wait();
    epi_hidden_bwd<D1>(c.tmem + TC_DH1, t, m1, smem + SM::DH_HI, smem + SM::DH_LO);
    sync_after_smem_writes();
    if (threadIdx.x == 0) {
      issue_gemm(c.tmem + TC_DA, c.sbase + SM::DH_HI, c.sbase + SM::DH_LO, c.sbase + SM::H_HI,
                 c.sbase + SM::H_LO, 8, true, true, ID_DA, !first_tile);
      issue_gemm(c.tmem + TC_DB, c.sbase + SM::DH_HI, c.sbase + SM::DH_LO, c.sbase + SM::X0, 0, 8, true,
                 true, ID_DB, !first_tile);
      umma::commit(c.bar);
    }
    dw_pending = true;
    first_tile = false;
  }

  float *part = a.partials + (size_t)blockIdx.x * net.n_params;
  if (dw_pending)
    c.wait();
  {
    constexpr int DC = D1 / 2;
    float v[DC];
    tmem_load<DC>(c.tmem + TC_DA + t.lane_base + t.wg * DC, v);
    int nrow = t.row - 64;
    if (nrow >= 0 && nrow < D2)
#pragma unroll
      for (int j = 0; j < DC; ++j)
        part[net.o_w2 + nrow * D1 + t.wg * DC + j] = v[j];
  }
  {
    constexpr int DC = (D0 + 16) / 2;
    float v[DC];
    tmem_load<DC>(c.tmem + TC_DB + t.lane_base + t.wg * DC, v);
#pragma unroll
    for (int j = 0; j < DC; ++j) {
      int col = t.wg * DC + j;
      if (t.row < D1) {
        if (col < D0)
          part[net.o_w1 + t.row * D0 + col] = v[j];
        else if (col == D0)
          part[net.o_b1 + t.row] = v[j];
      } else if (t.row >= 64 && t.row - 64 < D2 && col == D0) {
        part[net.o_b2 + t.row - 64] = v[j];
      }
    }
  }
  // dW3 / db3: thread-local sums -> fixed-order column sums through shared memory (reuse the H
  // panels as fp32 scratch: [128 rows][D2 + 1])
  __syncthreads();
  float *scr = reinterpret_cast<float *>(smem + SM::H_HI);
#pragma unroll
  for (int j = 0; j < DC2; ++j)
    scr[t.row * (D2 + 1) + t.wg * DC2 + j] = dw3[j];
  if (t.wg == 0)
    scr[t.row * (D2 + 1) + D2] = db3;
  __syncthreads();
  if (threadIdx.x <= D2) {
    float s = 0.f;
    for (int r = 0; r < TILE; ++r)
      s += scr[r * (D2 + 1) + threadIdx.x];
    if (threadIdx.x < D2)
      part[net.o_w3 + threadIdx.x] = s;
    else
      part[net.o_b3] = s;
  }
  umma::fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32)
    umma::tmem_dealloc(c.tmem, TC_COLS);
}

// calculate_advantage (policy_gradient.h:220-281) with the updated critic: V of start / end
// rows, then GAE per environment (all T steps of an env live in the tile).
template <int D0, int D1, int D2>
__global__ void __launch_bounds__(256, 1) fused_gae_kernel(critic_args a) {
  using SM = smem_map<D1, D2>;
  constexpr int DC2 = D2 / 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float *fl = reinterpret_cast<float *>(smem + SM::FLOATS);
  float *b1 = fl, *b2 = fl + D1, *w3 = fl + D1 + D2 + 16, *vpart = w3 + 64, *ve = vpart + 2 * TILE,
        *vs = ve + TILE;
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + SM::BARS);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + SM::BARS + 16);
  const net3 net = a.net;
  tile_ctx c;
  setup_common<D0, D1, D2>(c, smem, a.params, net, b1, b2, tmem_slot, bar);
  for (int i = threadIdx.x; i < D2; i += blockDim.x) w3[i] = a.params[net.o_w3 + i];
  const float b3 = a.params[net.o_b3];
  sync_after_smem_writes();
  c.tmem = *tmem_slot;
  const tid_t t = c.t;
  const learner_rows &L = a.rows;
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    uint32_t m1, m2;
    float h2[DC2];
    stage_x0(smem + SM::X0, L, tile, true, []() {});
    sync_after_smem_writes();
    fwd_hidden<D0, D1, D2, false>(c, b1, b2, m1, m2, h2);
    float v_end = value_head<D2>(t, h2, w3, b3, vpart);
    if (t.wg == 0)
      ve[t.row] = v_end;
    __syncthreads();
    stage_x0(smem + SM::X0, L, tile, false, []() {});
    sync_after_smem_writes();
    fwd_hidden<D0, D1, D2, false>(c, b1, b2, m1, m2, h2);
    float v = value_head<D2>(t, h2, w3, b3, vpart);
    if (t.wg == 0)
      vs[t.row] = v;
    __syncthreads();
    // GAE: thread e < E walks its env backwards (device_fns.cuh gae_env on shared-memory values)
    if (threadIdx.x < L.E) {
      int e = threadIdx.x, i = tile * L.E + e;
      if (i < L.n) {
        float a_next = 0.f;
        for (int tt = L.T - 1; tt >= 0; --tt) {
          size_t k = (size_t)tt * L.n + i;
          int r = tt * L.E + e;
          int d = L.rec_done[k];
          bool ends = d || tt == L.T - 1;
          float vn = ends ? ve[r] : vs[r + L.E];
          float vn_adv = d ? 0.f : vn;
          float delta = (d ? 0.f : 1.f) + a.gamma * vn_adv - vs[r];
          float adv = delta + (ends ? 0.f : a.lambda * a.gamma * a_next);
          a.adv_out[k] = adv;
          a_next = adv;
        }
      }
    }
    __syncthreads();
  }
  umma::fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32)
    umma::tmem_dealloc(c.tmem, TC_COLS);
}

enum { HEAD_JACOBIAN = 0, HEAD_IDENTITY = 1 };

struct policy_step_args {
  const float *params;        // flat fp32 policy parameters
  net3 net;
  const int8_t *rec_state;    // [T][P][stride]
  const uint8_t *rec_action;  // [T][n]
  const float *adv;           // [T][n]
  const float *p_old;         // [T][n][B]
  int n, stride, T, E, n_tiles, B;
  float inv_w, inv_h;
  int loss_kind, head_bwd;
  float *partials;            // [gridDim.x][n_params]
};

// ---------------------------------------------------------------------------------------------
// One policy optimizer::step minus the update: forward + loss gradient + backward + dW partials.
template <int D0, int D1, int D2, int NOUT>
__global__ void __launch_bounds__(256, 1) fused_policy_step_kernel(policy_step_args a) {
  using SM = smem_map<D1, D2>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = umma::smem_u32(smem);
  float *fl = reinterpret_cast<float *>(smem + SM::FLOATS);
  float *b1 = fl, *b2 = fl + D1, *b3 = fl + D1 + D2, *red = fl + D1 + D2 + 16 + 64;
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + SM::BARS);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + SM::BARS + 16);
  const tid_t t = thread_id();
  const net3 net = a.net;

  // ---- one-time setup: TMEM, barrier, weights -> bf16 hi/lo panels, constant panel parts
  if (threadIdx.x < 32)
    umma::tmem_alloc(tmem_slot, TC_COLS);
  if (threadIdx.x == 0) {
    umma::mbar_init(bar, 1);
    umma::fence_mbar_init();
  }
  stage_weight(a.params + net.o_w1, D1, D0, D1, smem + SM::W1_HI, smem + SM::W1_LO);
  stage_weight(a.params + net.o_w2, D2, D1, D2, smem + SM::W2_HI, smem + SM::W2_LO);
  stage_weight(a.params + net.o_w3, NOUT, D2, 16, smem + SM::W3_HI, smem + SM::W3_LO);
  for (int i = threadIdx.x; i < D1; i += blockDim.x) b1[i] = a.params[net.o_b1 + i];
  for (int i = threadIdx.x; i < D2; i += blockDim.x) b2[i] = a.params[net.o_b2 + i];
  for (int i = threadIdx.x; i < 16; i += blockDim.x) b3[i] = i < NOUT ? a.params[net.o_b3 + i] : 0.f;
  zero_bytes(smem + SM::X0, PANEL);
  zero_bytes(smem + SM::H_HI, 4 * PANEL);
  zero_bytes(smem + SM::DH_HI, 4 * PANEL);
  zero_bytes(smem + SM::DY_HI, 2 * PANEL);
  __syncthreads();
  // ones column (col D0) of the X0 panel: [dH1|dH2]^T . 1 = bias gradients for free
  if (threadIdx.x < TILE)
    *reinterpret_cast<uint16_t *>(smem + SM::X0 + umma::panel_off(threadIdx.x, D0)) = 0x3F80;  // bf16 1.0
  sync_after_smem_writes();
  const uint32_t tmem = *tmem_slot;

  constexpr uint32_t ID_L1 = umma::make_idesc_bf16(128, D1, 0, 0);
  constexpr uint32_t ID_L2 = umma::make_idesc_bf16(128, D2, 0, 0);
  constexpr uint32_t ID_L3 = umma::make_idesc_bf16(128, 16, 0, 0);
  constexpr uint32_t ID_DH2 = umma::make_idesc_bf16(128, D2, 0, 1);
  constexpr uint32_t ID_DH1 = umma::make_idesc_bf16(128, D1, 0, 1);
  constexpr uint32_t ID_DA = umma::make_idesc_bf16(128, 64, 1, 1);
  constexpr uint32_t ID_DB = umma::make_idesc_bf16(128, D0 + 16, 1, 1);
  constexpr uint32_t ID_DC = umma::make_idesc_bf16(128, 16, 1, 1);

  uint32_t phase = 0;
  bool dw_pending = false, first_tile = true;
  float db3[NOUT];
#pragma unroll
  for (int j = 0; j < NOUT; ++j)
    db3[j] = 0.f;
  const int P = 2 * a.B + 2;
  const int chunks_per_row = a.B / 2;

  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    // ---- stage X0 (global loads first, then wait for the previous tile's dW GEMMs which still
    // read the panels, then store)
    uint4 xc[2];
    int xrow[2], xchunk[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      int task = threadIdx.x + q * 256;
      int r = task / chunks_per_row, c = task % chunks_per_row;
      xrow[q] = r;
      xchunk[q] = c;
      bool in = task < TILE * chunks_per_row;
      int tt = r / a.E, e = r % a.E;
      int i = tile * a.E + e;
      bool valid = in && tt < a.T && i < a.n;
      xc[q] = obs_chunk(a.rec_state + (size_t)tt * P * a.stride, a.stride, i, a.B, c, a.inv_w, a.inv_h,
                        -1, valid);
      if (!in)
        xrow[q] = -1;
    }
    if (dw_pending) {
      umma::mbar_wait(bar, phase);
      phase ^= 1;
      umma::fence_after_sync();
      dw_pending = false;
    }
#pragma unroll
    for (int q = 0; q < 2; ++q)
      if (xrow[q] >= 0)
        *reinterpret_cast<uint4 *>(smem + SM::X0 + umma::panel_chunk_off(xrow[q], xchunk[q])) = xc[q];
    sync_after_smem_writes();

    // ---- layer 1
    if (threadIdx.x == 0) {
      issue_gemm(tmem + TC_L1, sbase + SM::X0, 0, sbase + SM::W1_HI, sbase + SM::W1_LO, D0 / 16, false,
                 false, ID_L1, false);
      umma::commit(bar);
    }
    umma::mbar_wait(bar, phase);
    phase ^= 1;
    umma::fence_after_sync();
    uint32_t mask1 = epi_hidden_fwd<D1>(tmem + TC_L1, t, b1, smem + SM::H_HI, smem + SM::H_LO);
    sync_after_smem_writes();
    // ---- layer 2
    if (threadIdx.x == 0) {
      issue_gemm(tmem + TC_L2, sbase + SM::H_HI, sbase + SM::H_LO, sbase + SM::W2_HI, sbase + SM::W2_LO,
                 D1 / 16, false, false, ID_L2, false);
      umma::commit(bar);
    }
    umma::mbar_wait(bar, phase);
    phase ^= 1;
    umma::fence_after_sync();
    uint32_t mask2 = epi_hidden_fwd<D2>(tmem + TC_L2, t, b2, smem + SM::H_HI + PANEL, smem + SM::H_LO + PANEL);
    sync_after_smem_writes();
    // ---- layer 3 (head, N padded to 16)
    if (threadIdx.x == 0) {
      issue_gemm(tmem + TC_L3, sbase + SM::H_HI + PANEL, sbase + SM::H_LO + PANEL, sbase + SM::W3_HI,
                 sbase + SM::W3_LO, D2 / 16, false, false, ID_L3, false);
      umma::commit(bar);
    }
    umma::mbar_wait(bar, phase);
    phase ^= 1;
    umma::fence_after_sync();
    // ---- head epilogue: softmax, loss gradient, softmax backward -> dY panel (cols 0..NOUT-1)
    if (t.wg == 0) {
      float v[8];
      tmem_load<8>(tmem + TC_L3 + t.lane_base, v);
      int tt = t.row / a.E, e = t.row % a.E;
      int i = tile * a.E + e;
      bool valid = tt < a.T && i < a.n;
      float dl[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        dl[j] = 0.f;
      if (valid) {
        size_t k = (size_t)tt * a.n + i;
        float p[NOUT], s = 0.f;
#pragma unroll
        for (int j = 0; j < NOUT; ++j) {
          p[j] = expf(v[j] + b3[j]);  // no max subtraction (nn.h:382-392)
          s += p[j];
        }
#pragma unroll
        for (int j = 0; j < NOUT; ++j)
          p[j] = p[j] / s;
        int act = a.rec_action[k];
        float A = a.adv[k];
        float g[NOUT];
        if (a.loss_kind == DFRL_LOSS_CLIPPED) {
          float pa = 0.f;
#pragma unroll
          for (int j = 0; j < NOUT; ++j)
            pa = (j == act) ? p[j] : pa;
          float gc = clipped_grad(pa, a.p_old[k * a.B + act], A);
#pragma unroll
          for (int j = 0; j < NOUT; ++j)
            g[j] = (j == act) ? gc : 0.f;
        } else {
#pragma unroll
          for (int j = 0; j < NOUT; ++j)
            g[j] = p[j] * A - (j == act ? A : 0.f);
        }
        if (a.head_bwd == HEAD_JACOBIAN) {
          float dot = 0.f;
#pragma unroll
          for (int j = 0; j < NOUT; ++j)
            dot = fmaf(p[j], g[j], dot);
#pragma unroll
          for (int j = 0; j < NOUT; ++j)
            dl[j] = p[j] * (g[j] - dot);
        } else {
#pragma unroll
          for (int j = 0; j < NOUT; ++j)
            dl[j] = g[j];
        }
#pragma unroll
        for (int j = 0; j < NOUT; ++j)
          db3[j] += dl[j];
      }
      uint4 h, l;
      umma::split8(dl, h, l);
      uint32_t off = umma::panel_chunk_off(t.row, 0);
      *reinterpret_cast<uint4 *>(smem + SM::DY_HI + off) = h;
      *reinterpret_cast<uint4 *>(smem + SM::DY_LO + off) = l;
    }
    sync_after_smem_writes();
    // ---- dH2 = dY . W3 (contraction over the 16 padded outputs), then relu mask
    if (threadIdx.x == 0) {
      issue_gemm(tmem + TC_DH2, sbase + SM::DY_HI, sbase + SM::DY_LO, sbase + SM::W3_HI, sbase + SM::W3_LO,
                 1, false, true, ID_DH2, false);
      umma::commit(bar);
    }
    umma::mbar_wait(bar, phase);
    phase ^= 1;
    umma::fence_after_sync();
    epi_hidden_bwd<D2>(tmem + TC_DH2, t, mask2, smem + SM::DH_HI + PANEL, smem + SM::DH_LO + PANEL);
    sync_after_smem_writes();
    // ---- dH1 = dH2 . W2, relu mask
    if (threadIdx.x == 0) {
      issue_gemm(tmem + TC_DH1, sbase + SM::DH_HI + PANEL, sbase + SM::DH_LO + PANEL, sbase + SM::W2_HI,
                 sbase + SM::W2_LO, D2 / 16, false, true, ID_DH1, false);
      umma::commit(bar);
    }
    umma::mbar_wait(bar, phase);
    phase ^= 1;
    umma::fence_after_sync();
    epi_hidden_bwd<D1>(tmem + TC_DH1, t, mask1, smem + SM::DH_HI, smem + SM::DH_LO);
    sync_after_smem_writes();
    // ---- weight gradients, accumulated in TMEM across the CTA's tiles (contraction = 128 rows):
    //   DA[128 x 64]      += [dH1|dH2]^T . H1        rows 64.. = dW2
    //   DB[128 x D0+16]   += [dH1|dH2]^T . [X0|1]    rows 0..  = [dW1 | db1], rows 64.. col D0 = db2
    //   DC[128 x 16]      += [H1|H2]^T . dY          rows 64.. = dW3^T
    if (threadIdx.x == 0) {
      issue_gemm(tmem + TC_DA, sbase + SM::DH_HI, sbase + SM::DH_LO, sbase + SM::H_HI, sbase + SM::H_LO, 8,
                 true, true, ID_DA, !first_tile);
      issue_gemm(tmem + TC_DB, sbase + SM::DH_HI, sbase + SM::DH_LO, sbase + SM::X0, 0, 8, true, true, ID_DB,
                 !first_tile);
      issue_gemm(tmem + TC_DC, sbase + SM::H_HI, sbase + SM::H_LO, sbase + SM::DY_HI, sbase + SM::DY_LO, 8,
                 true, true, ID_DC, !first_tile);
      umma::commit(bar);
    }
    dw_pending = true;
    first_tile = false;
  }

  // ---- drain: partial gradient of this CTA -> global, in the flat parameter layout
  float *part = a.partials + (size_t)blockIdx.x * net.n_params;
  if (dw_pending) {
    umma::mbar_wait(bar, phase);
    phase ^= 1;
    umma::fence_after_sync();
  }
  if (first_tile) {  // CTA had no tile: contribute zeros
    for (int i = threadIdx.x; i < net.n_params; i += blockDim.x)
      part[i] = 0.f;
  } else {
    // dW2[n][k]: DA row 64 + n, col k
    {
      constexpr int DC = D1 / 2;
      float v[DC];
      tmem_load<DC>(tmem + TC_DA + t.lane_base + t.wg * DC, v);
      int nrow = t.row - 64;
      if (nrow >= 0 && nrow < D2)
#pragma unroll
        for (int j = 0; j < DC; ++j)
          part[net.o_w2 + nrow * D1 + t.wg * DC + j] = v[j];
    }
    // dW1[n][k] + db1[n]: DB row n, cols 0..D0-1 and D0; db2[n]: DB row 64 + n, col D0
    {
      constexpr int DC = (D0 + 16) / 2;
      float v[DC];
      tmem_load<DC>(tmem + TC_DB + t.lane_base + t.wg * DC, v);
#pragma unroll
      for (int j = 0; j < DC; ++j) {
        int col = t.wg * DC + j;
        if (t.row < D1) {
          if (col < D0)
            part[net.o_w1 + t.row * D0 + col] = v[j];
          else if (col == D0)
            part[net.o_b1 + t.row] = v[j];
        } else if (t.row >= 64 && t.row - 64 < D2 && col == D0) {
          part[net.o_b2 + t.row - 64] = v[j];
        }
      }
    }
    // dW3[n][k] = DC row 64 + k, col n
    if (t.wg == 0) {
      float v[8];
      tmem_load<8>(tmem + TC_DC + t.lane_base, v);
      int krow = t.row - 64;
      if (krow >= 0 && krow < D2)
#pragma unroll
        for (int j = 0; j < NOUT; ++j)
          part[net.o_w3 + j * D2 + krow] = v[j];
    }
    // db3: per-thread partial sums -> fixed-order block sum
    if (t.wg == 0)
#pragma unroll
      for (int j = 0; j < NOUT; ++j)
        red[t.row * 8 + j] = db3[j];
    __syncthreads();
    if (threadIdx.x < NOUT) {
      float s = 0.f;
      for (int r = 0; r < TILE; ++r)
        s += red[r * 8 + threadIdx.x];
      part[net.o_b3 + threadIdx.x] = s;
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32)
    umma::tmem_dealloc(tmem, TC_COLS);
}

// Second stage of the gradient: fixed-order sum of the per-CTA partials.
__global__ void fused_reduce_partials_kernel(const float *__restrict__ part, int ctas, int n,
                                             float *__restrict__ grad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  float s = 0.f;
  for (int c = 0; c < ctas; ++c)
    s += part[(size_t)c * n + i];
  grad[i] = s;
}

// ---------------------------------------------------------------------------------------------
// Rollout: agent::play_steps(T) (rl.h:325-360) for a tile of 128 environments per CTA iteration.
// The tile's int8 state planes stay in shared memory for all T steps; per step the CTA records the
// start state, encodes the observation panel (exact in bf16: multiples of 1/cap), runs the three
// forward GEMMs on the tensor cores, and one thread per environment does softmax -> action
// (sample / argmax / forced) -> environment::apply -> reward / done / reset / next item.
// Two CTAs per SM (84 KB shared memory, 256 TMEM columns each) so that one CTA's epilogue
// overlaps the other's MMAs.
template <int D1, int D2>
struct smem_fwd {
  static constexpr uint32_t W1_HI = 0, W1_LO = W1_HI + D1 * 128;
  static constexpr uint32_t W2_HI = W1_LO + D1 * 128, W2_LO = W2_HI + D2 * 128;
  static constexpr uint32_t W3_HI = W2_LO + D2 * 128, W3_LO = W3_HI + 16 * 128;
  static constexpr uint32_t X0 = W3_LO + 16 * 128;
  static constexpr uint32_t H_HI = X0 + PANEL, H_LO = H_HI + PANEL;  // H1, then H2 in place
  static constexpr uint32_t FLOATS = H_LO + PANEL;                   // b1, b2, b3 / w3
  static constexpr uint32_t N_FLOATS = D1 + D2 + 16 + 64 + 4 * TILE;
  static constexpr uint32_t STATE = FLOATS + N_FLOATS * 4;           // int8 [18][128]
  static constexpr uint32_t BARS = STATE + 18 * TILE;
  static constexpr uint32_t TOTAL = BARS + 64;
};
constexpr uint32_t TF_L1 = 0, TF_L2 = 64, TF_L3 = 128, TF_COLS = 256;

struct rollout_args {
  const float *params;
  net3 net;
  env_params ep;
  int8_t *state;             // live planes [P][stride]
  uint32_t *draws, *steps;
  int T, n_tiles, mode;
  const uint8_t *forced;     // [T][n] or null
  const double *u_tape;      // [T][n] or null
  const uint8_t *item_tape;  // [T][n] or null (item drawn after step t)
  int8_t *rec_state;         // [T][P][stride]
  uint8_t *rec_action, *rec_done;
  float *rec_probs;          // [T][n][B]
  unsigned long long *counters;
  float inv_w, inv_h;
};

template <int D0, int D1, int D2, int NOUT>
__global__ void __launch_bounds__(256, 2) fused_rollout_kernel(rollout_args a) {
  using SM = smem_fwd<D1, D2>;
  constexpr int B = NOUT, P = 2 * B + 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = umma::smem_u32(smem);
  float *fl = reinterpret_cast<float *>(smem + SM::FLOATS);
  float *b1 = fl, *b2 = fl + D1, *b3 = fl + D1 + D2;
  int8_t *sst = reinterpret_cast<int8_t *>(smem + SM::STATE);
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + SM::BARS);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + SM::BARS + 16);
  const tid_t t = thread_id();
  const net3 net = a.net;
  const env_params &ep = a.ep;

  if (threadIdx.x < 32)
    umma::tmem_alloc(tmem_slot, TF_COLS);
  if (threadIdx.x == 0) {
    umma::mbar_init(bar, 1);
    umma::fence_mbar_init();
  }
  stage_weight(a.params + net.o_w1, D1, D0, D1, smem + SM::W1_HI, smem + SM::W1_LO);
  stage_weight(a.params + net.o_w2, D2, D1, D2, smem + SM::W2_HI, smem + SM::W2_LO);
  stage_weight(a.params + net.o_w3, NOUT, D2, 16, smem + SM::W3_HI, smem + SM::W3_LO);
  for (int i = threadIdx.x; i < D1; i += blockDim.x) b1[i] = a.params[net.o_b1 + i];
  for (int i = threadIdx.x; i < D2; i += blockDim.x) b2[i] = a.params[net.o_b2 + i];
  for (int i = threadIdx.x; i < 16; i += blockDim.x) b3[i] = i < NOUT ? a.params[net.o_b3 + i] : 0.f;
  zero_bytes(smem + SM::X0, 3 * PANEL);
  sync_after_smem_writes();
  const uint32_t tmem = *tmem_slot;

  constexpr uint32_t ID_L1 = umma::make_idesc_bf16(128, D1, 0, 0);
  constexpr uint32_t ID_L2 = umma::make_idesc_bf16(128, D2, 0, 0);
  constexpr uint32_t ID_L3 = umma::make_idesc_bf16(128, 16, 0, 0);
  uint32_t phase = 0;
  unsigned long long c_eps = 0, c_reward = 0, c_steps = 0;
  const size_t S = ep.stride;
  // plane copies: thread (q, c) moves the 16 environments [16c, 16c+16) of plane q
  const int cq = threadIdx.x >> 3, cc = threadIdx.x & 7;

  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int i0 = tile * TILE;
    const bool copier = cq < P && i0 + 16 * cc < ep.stride;
    if (cq < P) {
      uint4 v = make_uint4(0, 0, 0, 0);
      if (copier)
        v = *reinterpret_cast<const uint4 *>(a.state + (size_t)cq * S + i0 + 16 * cc);
      *reinterpret_cast<uint4 *>(sst + cq * TILE + 16 * cc) = v;
    }
    const int r = t.row, i = i0 + r;
    const bool owner = t.wg == 0 && i < ep.n;
    uint32_t my_draws = 0, my_steps = 0;
    if (owner) {
      my_draws = a.draws[i];
      my_steps = a.steps[i];
    }
    __syncthreads();
    for (int tt = 0; tt < a.T; ++tt) {
      // ---- record the start state of step tt, encode the observation panel
      if (copier)
        *reinterpret_cast<uint4 *>(a.rec_state + ((size_t)tt * P + cq) * S + i0 + 16 * cc) =
            *reinterpret_cast<const uint4 *>(sst + cq * TILE + 16 * cc);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        int task = threadIdx.x + q * 256;
        int row = task >> 2, c = task & 3;
        float iw = (float)sst[(2 * B) * TILE + row] * a.inv_w, ih = (float)sst[(2 * B + 1) * TILE + row] * a.inv_h;
        uint32_t out[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          int b = 2 * c + h;
          float bw = (float)sst[(2 * b) * TILE + row] * a.inv_w, bh = (float)sst[(2 * b + 1) * TILE + row] * a.inv_h;
          out[2 * h] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(bw)) |
                       ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(bh)) << 16);
          out[2 * h + 1] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(iw)) |
                           ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(ih)) << 16);
        }
        *reinterpret_cast<uint4 *>(smem + SM::X0 + umma::panel_chunk_off(row, c)) =
            make_uint4(out[0], out[1], out[2], out[3]);
      }
      sync_after_smem_writes();
      // ---- forward
      if (threadIdx.x == 0) {
        issue_gemm(tmem + TF_L1, sbase + SM::X0, 0, sbase + SM::W1_HI, sbase + SM::W1_LO, D0 / 16, false,
                   false, ID_L1, false);
        umma::commit(bar);
      }
      umma::mbar_wait(bar, phase);
      phase ^= 1;
      umma::fence_after_sync();
      epi_hidden_fwd<D1>(tmem + TF_L1, t, b1, smem + SM::H_HI, smem + SM::H_LO);
      sync_after_smem_writes();
      if (threadIdx.x == 0) {
        issue_gemm(tmem + TF_L2, sbase + SM::H_HI, sbase + SM::H_LO, sbase + SM::W2_HI, sbase + SM::W2_LO,
                   D1 / 16, false, false, ID_L2, false);
        umma::commit(bar);
      }
      umma::mbar_wait(bar, phase);
      phase ^= 1;
      umma::fence_after_sync();
      epi_hidden_fwd<D2>(tmem + TF_L2, t, b2, smem + SM::H_HI, smem + SM::H_LO);  // H2 over H1
      sync_after_smem_writes();
      if (threadIdx.x == 0) {
        issue_gemm(tmem + TF_L3, sbase + SM::H_HI, sbase + SM::H_LO, sbase + SM::W3_HI, sbase + SM::W3_LO,
                   D2 / 16, false, false, ID_L3, false);
        umma::commit(bar);
      }
      umma::mbar_wait(bar, phase);
      phase ^= 1;
      umma::fence_after_sync();
      // ---- head: softmax (no max subtraction, nn.h:382-392), action, environment::apply
      if (t.wg == 0) {
        float v[8];
        tmem_load<8>(tmem + TF_L3 + t.lane_base, v);
        if (owner) {
          const size_t k = (size_t)tt * ep.n + i;
          float p[NOUT], s = 0.f;
#pragma unroll
          for (int j = 0; j < NOUT; ++j) {
            p[j] = expf(v[j] + b3[j]);
            s += p[j];
          }
#pragma unroll
          for (int j = 0; j < NOUT; ++j)
            p[j] = p[j] / s;
          float4 *pr = reinterpret_cast<float4 *>(a.rec_probs + k * B);
#pragma unroll
          for (int j = 0; j < NOUT / 4; ++j)
            pr[j] = make_float4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
          int act;
          if (a.mode == DFRL_ACT_FORCED) {
            act = a.forced[k];
          } else if (a.mode == DFRL_ACT_ARGMAX) {
            act = argmax_first(p, B);
          } else {
            double u;
            if (a.u_tape) {
              u = a.u_tape[k];
            } else {
              philox4 rr = philox4x32_10(ep.seed, (uint64_t)(ep.env_offset + i), my_steps, DFRL_STREAM_ACTION);
              u = philox_u53(rr.x, rr.y);
            }
            act = discrete_sample(p, B, u);
          }
          act = act < B ? act : B - 1;
          a.rec_action[k] = (uint8_t)act;
          // environment::apply (bin_packing.h:53-64) on this thread's column of the tile
          int iw = sst[(2 * B) * TILE + r], ih = sst[(2 * B + 1) * TILE + r];
          int bw = sst[(2 * act) * TILE + r] - iw, bh = sst[(2 * act + 1) * TILE + r] - ih;
          bool over = bw < 0 || bh < 0;
          int s1 = a.item_tape ? (a.item_tape[k] != 0) : draw_shape1(ep, i, my_draws);
          if (over) {
#pragma unroll
            for (int b = 0; b < B; ++b) {
              sst[(2 * b) * TILE + r] = (int8_t)ep.cap_w;
              sst[(2 * b + 1) * TILE + r] = (int8_t)ep.cap_h;
            }
          } else {
            sst[(2 * act) * TILE + r] = (int8_t)bw;
            sst[(2 * act + 1) * TILE + r] = (int8_t)bh;
          }
          sst[(2 * B) * TILE + r] = (int8_t)(s1 ? ep.iw0 : ep.iw1);
          sst[(2 * B + 1) * TILE + r] = (int8_t)(s1 ? ep.ih0 : ep.ih1);
          a.rec_done[k] = over;
          my_draws += 1;
          my_steps += 1;
          c_steps += 1;
          c_eps += over ? 1 : 0;
          c_reward += over ? 0 : 1;
        }
      }
      umma::fence_before_sync();
      __syncthreads();
      umma::fence_after_sync();
    }
    // ---- live state back to the environment
    if (copier)
      *reinterpret_cast<uint4 *>(a.state + (size_t)cq * S + i0 + 16 * cc) =
          *reinterpret_cast<const uint4 *>(sst + cq * TILE + 16 * cc);
    if (owner) {
      a.draws[i] = my_draws;
      a.steps[i] = my_steps;
    }
    __syncthreads();
  }
  for (int o = 16; o > 0; o >>= 1) {
    c_steps += __shfl_down_sync(0xffffffffu, c_steps, o);
    c_eps += __shfl_down_sync(0xffffffffu, c_eps, o);
    c_reward += __shfl_down_sync(0xffffffffu, c_reward, o);
  }
  if (t.lane == 0 && c_steps) {
    atomicAdd(&a.counters[0], c_steps);
    atomicAdd(&a.counters[1], c_eps);
    atomicAdd(&a.counters[2], c_reward);
  }
  umma::fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32)
    umma::tmem_dealloc(tmem, TF_COLS);
}

struct fused_state {
  net3 pnet, vnet;
  bool policy_ok, value_ok;
  int head_bwd;
  float *partials;  // [ctas][max params]
  int ctas;
  bool rollout_ok;
};

// D - R - D - R - D (- softmax / softmax_ce)
bool parse_net3(const dfrl_mlp *m, net3 *out, int *tail_kind) {
  const auto &L = m->layers;
  size_t n = L.size();
  if (n != 5 && n != 6)
    return false;
  if (L[0].kind != DFRL_LAYER_DENSE || L[1].kind != DFRL_LAYER_RELU || L[2].kind != DFRL_LAYER_DENSE ||
      L[3].kind != DFRL_LAYER_RELU || L[4].kind != DFRL_LAYER_DENSE)
    return false;
  *tail_kind = -1;
  if (n == 6) {
    if (L[5].kind != DFRL_LAYER_SOFTMAX && L[5].kind != DFRL_LAYER_SOFTMAX_CE)
      return false;
    *tail_kind = L[5].kind;
  }
  out->d0 = L[0].in;
  out->d1 = L[0].out;
  out->d2 = L[2].out;
  out->d3 = L[4].out;
  out->o_w1 = (int)L[0].param_off;
  out->o_b1 = out->o_w1 + out->d0 * out->d1;
  out->o_w2 = (int)L[2].param_off;
  out->o_b2 = out->o_w2 + out->d1 * out->d2;
  out->o_w3 = (int)L[4].param_off;
  out->o_b3 = out->o_w3 + out->d2 * out->d3;
  out->n_params = m->n_params;
  return true;
}

bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

template <typename K>
int set_smem_once(K kernel, int smem, bool *done) {
  if (!*done) {
    DFRL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    *done = true;
  }
  return DFRL_OK;
}

template <int D0, int D1, int D2, int NOUT>
int launch_policy_step(dfrl_ctx *ctx, const policy_step_args &a, int ctas) {
  constexpr int smem = smem_map<D1, D2>::TOTAL + 1024;
  static bool attr = false;
  DFRL_TRY(set_smem_once(fused_policy_step_kernel<D0, D1, D2, NOUT>, smem, &attr));
  DFRL_LAUNCH(ctx, (fused_policy_step_kernel<D0, D1, D2, NOUT>), ctas, 256, smem, a);
  return DFRL_OK;
}

template <int D0, int D1, int D2>
int launch_critic_step(dfrl_ctx *ctx, const critic_args &a, int ctas) {
  constexpr int smem = smem_map<D1, D2>::TOTAL + 1024;
  static bool attr = false;
  DFRL_TRY(set_smem_once(fused_critic_step_kernel<D0, D1, D2>, smem, &attr));
  DFRL_LAUNCH(ctx, (fused_critic_step_kernel<D0, D1, D2>), ctas, 256, smem, a);
  return DFRL_OK;
}

template <int D0, int D1, int D2>
int launch_gae(dfrl_ctx *ctx, const critic_args &a, int ctas) {
  constexpr int smem = smem_map<D1, D2>::TOTAL + 1024;
  static bool attr = false;
  DFRL_TRY(set_smem_once(fused_gae_kernel<D0, D1, D2>, smem, &attr));
  DFRL_LAUNCH(ctx, (fused_gae_kernel<D0, D1, D2>), ctas, 256, smem, a);
  return DFRL_OK;
}

template <int D0, int D1, int D2, int NOUT>
int launch_rollout(dfrl_ctx *ctx, const rollout_args &a, int ctas) {
  constexpr int smem = smem_fwd<D1, D2>::TOTAL + 1024;
  static bool attr = false;
  DFRL_TRY(set_smem_once(fused_rollout_kernel<D0, D1, D2, NOUT>, smem, &attr));
  DFRL_LAUNCH(ctx, (fused_rollout_kernel<D0, D1, D2, NOUT>), ctas, 256, smem, a);
  return DFRL_OK;
}

bool widths_ok(const net3 &n) {
  return n.d0 == 32 && ((n.d1 == 64 && n.d2 == 64) || (n.d1 == 16 && n.d2 == 16));
}

critic_args make_critic_args(dfrl_trainer *t, fused_state *f) {
  critic_args a;
  a.params = t->value->params;
  a.net = f->vnet;
  a.rows.rec_state = t->rec_state;
  a.rows.live_state = t->env->state;
  a.rows.rec_action = t->rec_action;
  a.rows.rec_done = t->rec_done;
  a.rows.n = t->n;
  a.rows.stride = t->stride;
  a.rows.T = t->L;
  a.rows.E = TILE / t->L;
  a.rows.B = t->B;
  a.rows.inv_w = 1.0f / (float)t->env->cfg.cap_w;
  a.rows.inv_h = 1.0f / (float)t->env->cfg.cap_h;
  a.n_tiles = ceil_div(t->n, a.rows.E);
  a.gamma = t->cfg.gamma;
  a.lambda = t->cfg.lambda;
  a.targets_out = t->targets;
  a.adv_out = t->adv;
  a.partials = f->partials;
  return a;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
int dfrl_fused_try_attach(dfrl_trainer *t) {
  t->fused_impl = nullptr;
  const dfrl_trainer_config &c = t->cfg;
  if (c.algo == DFRL_ALGO_REINFORCE)
    return DFRL_ERR_UNSUPPORTED;
  if (t->L > TILE || t->B != 8)
    return DFRL_ERR_UNSUPPORTED;
  if (!is_pow2(t->env->cfg.cap_w) || !is_pow2(t->env->cfg.cap_h) || t->env->cfg.cap_w > 64 ||
      t->env->cfg.cap_h > 64)
    return DFRL_ERR_UNSUPPORTED;  // observations must be exact in bf16
  fused_state *f = new fused_state();
  int ptail = -1, vtail = -1;
  f->policy_ok = parse_net3(t->policy, &f->pnet, &ptail) && ptail != -1 && f->pnet.d3 == 8 && widths_ok(f->pnet);
  f->value_ok = t->value && parse_net3(t->value, &f->vnet, &vtail) && vtail == -1 && f->vnet.d3 == 1 &&
                widths_ok(f->vnet);
  f->rollout_ok = f->policy_ok;
  f->head_bwd = ptail == DFRL_LAYER_SOFTMAX ? HEAD_JACOBIAN : HEAD_IDENTITY;
  if (!f->policy_ok && !f->value_ok) {
    delete f;
    return DFRL_ERR_UNSUPPORTED;
  }
  f->ctas = t->ctx->sm_count;
  int maxp = t->policy->n_params;
  if (t->value && t->value->n_params > maxp)
    maxp = t->value->n_params;
  if (cudaMalloc(&f->partials, sizeof(float) * (size_t)f->ctas * maxp) != cudaSuccess) {
    delete f;
    return DFRL_ERR_CUDA;
  }
  t->fused_impl = f;
  return DFRL_OK;
}

void dfrl_fused_detach(dfrl_trainer *t) {
  fused_state *f = (fused_state *)t->fused_impl;
  if (f) {
    cudaFree(f->partials);
    delete f;
  }
  t->fused_impl = nullptr;
}

// One policy gradient (forward + loss + backward over all L*n rows) into grad_dev.
int dfrl_fused_policy_gradient(dfrl_trainer *t, int loss_kind, float *grad_dev) {
  fused_state *f = (fused_state *)t->fused_impl;
  if (!f || !f->policy_ok || loss_kind == DFRL_LOSS_KL)
    return DFRL_ERR_UNSUPPORTED;
  policy_step_args a;
  a.params = t->policy->params;
  a.net = f->pnet;
  a.rec_state = t->rec_state;
  a.rec_action = t->rec_action;
  a.adv = t->adv;
  a.p_old = t->rec_probs;
  a.n = t->n;
  a.stride = t->stride;
  a.T = t->L;
  a.E = TILE / t->L;
  a.n_tiles = ceil_div(t->n, a.E);
  a.B = t->B;
  a.inv_w = 1.0f / (float)t->env->cfg.cap_w;
  a.inv_h = 1.0f / (float)t->env->cfg.cap_h;
  a.loss_kind = loss_kind;
  a.head_bwd = f->head_bwd;
  a.partials = f->partials;
  int ctas = a.n_tiles < f->ctas ? a.n_tiles : f->ctas;
  if (f->pnet.d1 == 64)
    DFRL_TRY((launch_policy_step<32, 64, 64, 8>(t->ctx, a, ctas)));
  else
    DFRL_TRY((launch_policy_step<32, 16, 16, 8>(t->ctx, a, ctas)));
  DFRL_LAUNCH(t->ctx, fused_reduce_partials_kernel, ceil_div(f->pnet.n_params, 256), 256, 0,
              (const float *)f->partials, ctas, f->pnet.n_params, grad_dev);
  return DFRL_OK;
}

// update_value_model (policy_gradient.h:196-218) up to the gradient: writes t->targets and grad_dev.
int dfrl_fused_critic_gradient(dfrl_trainer *t, float *grad_dev) {
  fused_state *f = (fused_state *)t->fused_impl;
  if (!f || !f->value_ok)
    return DFRL_ERR_UNSUPPORTED;
  critic_args a = make_critic_args(t, f);
  int ctas = a.n_tiles < f->ctas ? a.n_tiles : f->ctas;
  if (f->vnet.d1 == 64)
    DFRL_TRY((launch_critic_step<32, 64, 64>(t->ctx, a, ctas)));
  else
    DFRL_TRY((launch_critic_step<32, 16, 16>(t->ctx, a, ctas)));
  DFRL_LAUNCH(t->ctx, fused_reduce_partials_kernel, ceil_div(f->vnet.n_params, 256), 256, 0,
              (const float *)f->partials, ctas, f->vnet.n_params, grad_dev);
  return DFRL_OK;
}

// calculate_advantage (policy_gradient.h:220-281) with the current (updated) critic: writes t->adv.
int dfrl_fused_gae(dfrl_trainer *t) {
  fused_state *f = (fused_state *)t->fused_impl;
  if (!f || !f->value_ok)
    return DFRL_ERR_UNSUPPORTED;
  critic_args a = make_critic_args(t, f);
  int ctas = a.n_tiles < f->ctas ? a.n_tiles : f->ctas;
  if (f->vnet.d1 == 64)
    DFRL_TRY((launch_gae<32, 64, 64>(t->ctx, a, ctas)));
  else
    DFRL_TRY((launch_gae<32, 16, 16>(t->ctx, a, ctas)));
  return DFRL_OK;
}

// agent::play_steps(L) for every env in one launch (AC / PPO / KL-PPO rollouts).
int dfrl_fused_rollout(dfrl_trainer *t, const uint8_t *items_dev, const uint8_t *actions_dev,
                       const double *u_dev) {
  fused_state *f = (fused_state *)t->fused_impl;
  if (!f || !f->rollout_ok)
    return DFRL_ERR_UNSUPPORTED;
  dfrl_env *e = t->env;
  rollout_args a;
  a.params = t->policy->params;
  a.net = f->pnet;
  a.ep = make_params(e);
  a.state = e->state;
  a.draws = e->draws;
  a.steps = e->steps;
  a.T = t->L;
  a.n_tiles = ceil_div(t->n, TILE);
  a.mode = t->cfg.action_mode;
  a.forced = actions_dev;
  a.u_tape = u_dev;
  a.item_tape = items_dev;
  a.rec_state = t->rec_state;
  a.rec_action = t->rec_action;
  a.rec_done = t->rec_done;
  a.rec_probs = t->rec_probs;
  a.counters = t->counters;
  a.inv_w = 1.0f / (float)e->cfg.cap_w;
  a.inv_h = 1.0f / (float)e->cfg.cap_h;
  int ctas = a.n_tiles < 2 * f->ctas ? a.n_tiles : 2 * f->ctas;
  if (f->pnet.d1 == 64)
    DFRL_TRY((launch_rollout<32, 64, 64, 8>(t->ctx, a, ctas)));
  else
    DFRL_TRY((launch_rollout<32, 16, 16, 8>(t->ctx, a, ctas)));
  t->obs_valid = false;
  return DFRL_OK;
}

int dfrl_fused_eval_argmax(dfrl_ctx *, dfrl_env *, dfrl_mlp *, int, double *, long long *) {
  return DFRL_ERR_UNSUPPORTED;
}
