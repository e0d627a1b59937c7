// common.cuh -- shared definitions of libdfrl_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "../../include/dfrl.h"

// ------------------------------------------------------------------ error handling ----------
void dfrl_set_error(const char *fmt, ...);

#define DFRL_CUDA(expr)                                                                      \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      dfrl_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return DFRL_ERR_CUDA;                                                                  \
    }                                                                                        \
  } while (0)

#define DFRL_CHECK(cond, ...)                                                                \
  do {                                                                                       \
    if (!(cond)) {                                                                           \
      dfrl_set_error(__VA_ARGS__);                                                           \
      return DFRL_ERR_INVALID;                                                               \
    }                                                                                        \
  } while (0)

#define DFRL_TRY(expr)                                                                       \
  do {                                                                                       \
    int _s = (expr);                                                                         \
    if (_s != DFRL_OK)                                                                       \
      return _s;                                                                             \
  } while (0)

// Launch wrapper: counts launches (bench.py reports gpu_launches) and checks the launch.
#define DFRL_LAUNCH(ctx, kernel, grid, block, smem, ...)                                     \
  do {                                                                                       \
    if ((ctx)->profiling)                                                                    \
      dfrl_profile_mark(ctx, #kernel, 0);                                                    \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                         \
    (ctx)->launches++;                                                                       \
    if ((ctx)->profiling)                                                                    \
      dfrl_profile_mark(ctx, #kernel, 1);                                                    \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess) {                                                                 \
      dfrl_set_error("%s:%d: launch of %s failed: %s", __FILE__, __LINE__, #kernel,          \
                     cudaGetErrorString(_e));                                                \
      return DFRL_ERR_CUDA;                                                                  \
    }                                                                                        \
  } while (0)

// ------------------------------------------------------------------ context -----------------
struct dfrl_ctx;
void dfrl_profile_mark(dfrl_ctx *ctx, const char *name, int end);

// Flat-gradient exchange over NVLink peer memory (context.cu, fused.cu). Every rank owns one
// allocation [2 slots][cap floats] + flags, IPC-mapped by every other rank of the node.
constexpr int DFRL_P2P_MAX_RANKS = 8;
constexpr size_t DFRL_P2P_CAP = 1 << 18;  // floats per gradient (1 MB): >= the largest flat gradient
// Exchange buffer of a rank (PUSH protocol, flag-in-data: every rank stores {value, exchange number}
// pairs INTO its peers' buffers with single 8-byte stores; the receiver polls its local copy until
// the exchange number matches -- no fence, no separate flag, one NVLink store latency):
//   data   [2 slots][DFRL_P2P_MAX_RANKS source ranks][DFRL_P2P_CAP] x {float value, unsigned epoch}
//   words  16 unsigned; word 2 = this rank's exchange counter
struct dfrl_p2p {
  float *local = nullptr;
  float *peer[DFRL_P2P_MAX_RANKS];   // peer[r]: rank r's allocation mapped here (own rank: local)
  bool attached = false;
};
constexpr size_t DFRL_P2P_DATA_FLOATS = 2 * 2 * (size_t)DFRL_P2P_MAX_RANKS * DFRL_P2P_CAP;  // 8-byte entries
constexpr size_t DFRL_P2P_BYTES = sizeof(float) * DFRL_P2P_DATA_FLOATS + 64;
__host__ __device__ static inline unsigned long long *dfrl_p2p_data(float *base, int slot, int src) {
  return reinterpret_cast<unsigned long long *>(base) + ((size_t)slot * DFRL_P2P_MAX_RANKS + src) * DFRL_P2P_CAP;
}
__host__ __device__ static inline unsigned *dfrl_p2p_flags(float *base) { return reinterpret_cast<unsigned *>(base + DFRL_P2P_DATA_FLOATS); }

struct dfrl_ctx {
  int device = 0;
  int nranks = 1, rank = 0;
  int sm_count = 148;
  int cc_major = 0, cc_minor = 0;
  size_t hbm_bytes = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  void *nccl_comm = nullptr;  // ncclComm_t
  long long launches = 0;
  // scratch arena (grown on demand, never shrunk): transient kernel workspaces
  void *scratch = nullptr;
  size_t scratch_bytes = 0;
  // second arena for the tcgen05 GEMMs (their B image / partials must not alias a caller's scratch)
  void *umma_ws = nullptr;
  size_t umma_ws_bytes = 0;
  dfrl_p2p p2p;
  // per-kernel event timing (dfrl_profile_*)
  int profiling = 0;
  void *prof = nullptr;
};

struct dfrl_ctx;
void dfrl_profile_mark(dfrl_ctx *ctx, const char *name, int end);

int dfrl_scratch(dfrl_ctx *ctx, size_t bytes, void **out);
int dfrl_umma_workspace(dfrl_ctx *ctx, size_t bytes, void **out);

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t round_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// ------------------------------------------------------------------ Philox4x32-10 -----------
// Counter-based RNG (Salmon et al. 2011).  key = seed; counter = (global env id, draw index,
// stream).  stream 0 = item draws (bernoulli), 1 = action-sampling uniforms, 2 = init.
struct philox4 {
  uint32_t x, y, z, w;
};

__host__ __device__ static inline philox4 philox4x32_10(uint64_t seed, uint64_t c01, uint32_t c2,
                                                        uint32_t c3) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)c01, c1 = (uint32_t)(c01 >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  philox4 o = {c0, c1, c2, c3};
  return o;
}

// 53-bit uniform in [0,1) from two words (same resolution as generate_canonical<double,53>).
__host__ __device__ static inline double philox_u53(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

enum { DFRL_STREAM_ITEM = 0, DFRL_STREAM_ACTION = 1, DFRL_STREAM_INIT = 2, DFRL_STREAM_HEUR = 3 };

// ------------------------------------------------------------------ objects ------------------
struct dfrl_env {
  dfrl_ctx *ctx;
  dfrl_env_config cfg;
  int n, B, P;       // envs, bins, planes = 2B+2
  int stride;        // plane stride (n rounded up to 16: aligned vector access)
  int8_t *state;     // [P][stride]
  uint32_t *draws;   // [n] items drawn so far by env i (tape / Philox index of the NEXT draw)
  uint32_t *steps;   // [n] actions taken so far (Philox index of the next sampling uniform)
  uint8_t *tape;     // [n][tape_len] or null
  int tape_len;
};

struct mlp_layer {
  int kind, in, out;
  int in_cols, out_cols;  // full row widths (conv1d: points * channels)
  int points;             // conv1d points (1 for dense)
  size_t param_off;       // offset into flat params ([W out x in][b out])
  size_t wt_off;          // offset into transposed-weight cache ([in x out])
};

struct dfrl_mlp {
  dfrl_ctx *ctx;
  std::vector<mlp_layer> layers;
  int input_cols, output_cols;
  int n_params;
  float *params;   // flat, reference order
  float *wt;       // W^T per parametric layer, refreshed after every parameter change
  bool wt_dirty;
  uint64_t version = 0;  // bumped on every parameter change (fused.cu panel-image cache)
  // shared-trunk models (dfrl_mlp_create_shared): every model of a family addresses ONE flat
  // parameter vector of n_params floats owned by `share_owner` (null: this model owns it); the first
  // n_shared_layers layers of a sharer carry the owner's parameter offsets
  dfrl_mlp *share_owner = nullptr;
  std::vector<dfrl_mlp *> sharers;
  int n_shared_layers = 0;
  // kept activations of the last forward_keep()
  std::vector<float *> acts;  // acts[l] = output of layer l (device), acts.size() == layers.size()
  float *act_arena;
  size_t act_arena_bytes;
  int kept_rows;
  const float *kept_input;
};

// Parameters changed on the device: invalidate the transposed-weight caches of every model that
// addresses the same flat vector.
static inline void dfrl_mlp_params_changed(dfrl_mlp *m) {
  dfrl_mlp *o = m->share_owner ? m->share_owner : m;
  o->wt_dirty = true, o->version++;
  for (dfrl_mlp *s : o->sharers)
    s->wt_dirty = true, s->version++;
}

// gemm_umma.cu: tcgen05 GEMMs of the layered path; DFRL_ERR_UNSUPPORTED = use the FFMA kernels
int umma_gemm_nn(dfrl_ctx *ctx, const float *A, const float *Bm, const float *bias, const float *mask, float *C,
                 int M, int N, int K, int relu);
int umma_gemm_tn(dfrl_ctx *ctx, const float *dY, const float *X, int M, int N, int K, float *grad, int accumulate);

int dfrl_mlp_refresh_wt(dfrl_mlp *m);
int dfrl_mlp_forward_keep(dfrl_mlp *m, const float *x_dev, int rows, float **out_dev);
int dfrl_mlp_backward(dfrl_mlp *m, const float *dy_dev, float *grad_dev);
