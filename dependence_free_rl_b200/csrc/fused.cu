// fused.cu -- fused small-MLP kernels on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// For 3-dense-layer nets  D0 -> D1 -relu-> D2 -relu-> D3  (C2/C4: 32-64-64-{8,1}) a tile of 128
// learner rows runs the WHOLE optimizer::step body without touching HBM for activations:
//   forward  H1 = relu(X0 W1^T + b1), H2 = relu(H1 W2^T + b2), out = H2 W3^T + b3     (3 UMMA GEMMs)
//   loss gradient at the output (PPO clipped surrogate / policy_loss / square loss), softmax
//   Jacobian                                                                         (registers)
//   input gradients dH2 = (dY W3) . relu', dH1 = (dH2 W2) . relu'                     (2 UMMA GEMMs)
//   weight gradients dW_l += dY_l^T X_l, accumulated over all tiles of the CTA in TMEM (3 UMMA GEMMs)
// Activations live in shared memory as SWIZZLE_128B 16-bit panels (umma.cuh): the same panel is
// the K-major A operand of the forward GEMM and the MN-major operand of the weight-gradient GEMM.
//
// Every kernel here has the same structure: a CTA runs several tile PIPELINES; a pipeline is 128
// epilogue threads (one per tile row, all columns) plus one MMA-issuing warp, with its own
// activation panels, TMEM accumulators, mbarriers and named barriers. A tile's work is a strictly
// dependent chain MMA -> epilogue -> MMA ...; while one pipeline is in an epilogue the other
// pipelines' MMAs execute. tcgen05.mma issue blocks the issuing thread for about the pipe time of
// the instruction (tools/mma_microbench.py), hence the dedicated issuer warps.
//
// FP32-level accuracy on the 16-bit tensor pipe: every operand is split into bf16 hi + lo halves
// and a product is accumulated as hi*hi + hi*lo + lo*hi in FP32 (TMEM); the dropped lo*lo term is
// ~2^-17. (FP16 pairs with power-of-two pre-scaling, 22 significant bits, were tried for the
// forward operands in round 1c: tcgen05.mma kind::f16 raised an illegal-instruction fault on B200
// when the A and B formats differ, which the weight-gradient GEMMs would need.)
// Every CTA splits the fp32 weights (27 KB, L2 resident) into its own shared-memory panels at
// start-up: no separate preparation kernel sits on the optimizer-step critical path.
//
// Row tiles are env-blocked: a tile holds E = 128 / T environments x all T steps (row = t * E + e)
// so that the critic target r + gamma V(s_{t+1}) and GAE only need values of the same tile.
//
// Reference semantics reproduced: rl.h:54-74 (clipped_gradient), rl.h:45-52, nn.h:393-417 (softmax
// Jacobian), nn.h:85-100 (dW = SUM over rows), policy_gradient.h:196-281 (targets, GAE).
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
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
  int shared;  // shared-trunk family (dfrl_mlp_create_shared): the vector holds other heads' parameters too
};
// Does this net own entry q of the flat vector? (always true unless the vector is shared)
__host__ __device__ __forceinline__ bool net_owns(const net3 &n, int q) {
  return (q >= n.o_w1 && q < n.o_b1 + n.d1) || (q >= n.o_w2 && q < n.o_b2 + n.d2) || (q >= n.o_w3 && q < n.o_b3 + n.d3);
}

struct tid_t {
  int wg, w, lane, row;
  int warp;            // warp index in the CTA, provably warp-uniform (MMA issue branch)
  uint32_t lane_base;  // TMEM lane field of this thread's warp
};
__device__ __forceinline__ tid_t thread_id() {
  tid_t t;
  t.warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  t.wg = threadIdx.x >> 7;
  t.w = (threadIdx.x >> 5) & 3;
  t.lane = threadIdx.x & 31;
  t.row = t.w * 32 + t.lane;
  t.lane_base = (uint32_t)(t.w * 32) << 16;
  return t;
}

// Load-balance probe (debug): %globaltimer in ns and the SM a CTA runs on.
__device__ __forceinline__ long long global_ns() {
  long long v;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(v));
  return v;
}
__device__ __forceinline__ unsigned sm_id() {
  unsigned v;
  asm volatile("mov.u32 %0, %%smid;\n" : "=r"(v));
  return v;
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

// ---- fp32 -> bf16 hi/lo pairs ------------------------------------------------------------------
template <bool UNUSED>
__device__ __forceinline__ void split8(const float *x, uint4 &hi, uint4 &lo) {
  umma::split8(x, hi, lo);
}
// Two fp32 values that are exact in bf16 -> one packed word.
__device__ __forceinline__ uint32_t pack2_fwd(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&h);
}

__device__ void zero_bytes(uint8_t *p, uint32_t bytes) {
  for (uint32_t o = threadIdx.x * 16; o < bytes; o += blockDim.x * 16)
    *reinterpret_cast<uint4 *>(p + o) = make_uint4(0, 0, 0, 0);
}

// ---- tcgen05.mma issue ----------------------------------------------------------------------
// Instruction descriptor, kind::f16: c_format F32, a/b format 0 = F16, 1 = BF16.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_bf16, int b_bf16, int a_mn, int b_mn) {
  return (1u << 4) | ((uint32_t)a_bf16 << 7) | ((uint32_t)b_bf16 << 10) | ((uint32_t)a_mn << 15) |
         ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Descriptor high word: SBO = 1024 B (8-row swizzle atom), version 1, SWIZZLE_128B.
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t desc_lo_hi(uint32_t lo) { return ((uint64_t)DESC_HI << 32) | lo; }
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo) {
  return ((saddr >> 4) & 0x3FFFu) | ((lbo >> 4) << 16);
}
// One GEMM = KSTEPS x {hi.hi, hi.lo, lo.hi} tcgen05.mma instructions, issued by one elected lane
// of warp 0 (warp-uniform branch: descriptors stay in uniform registers).
// A_LO / B_LO false: that operand is exact in 16 bits (no lo pass).
template <int KSTEPS, bool A_MN, bool B_MN, bool A_LO, bool B_LO>
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi,
                                           uint32_t b_lo, uint32_t idesc, bool accumulate) {
  constexpr uint32_t a_step = (A_MN ? umma::KSTEP_BYTES_MNMAJOR : umma::KSTEP_BYTES_KMAJOR) >> 4;
  constexpr uint32_t b_step = (B_MN ? umma::KSTEP_BYTES_MNMAJOR : umma::KSTEP_BYTES_KMAJOR) >> 4;
  constexpr uint32_t a_lbo = A_MN ? PANEL : 16, b_lbo = B_MN ? PANEL : 16;
  const uint32_t ah0 = desc_lo(a_hi, a_lbo), al0 = desc_lo(a_lo, a_lbo);
  const uint32_t bh0 = desc_lo(b_hi, b_lbo), bl0 = desc_lo(b_lo, b_lbo);
#pragma unroll  // (rolling the 8-step weight-gradient GEMMs: measured 0.6 % slower -- the issue rate matters more than the code size)
  for (int k = 0; k < KSTEPS; ++k) {
    uint64_t ah = desc_lo_hi(ah0 + k * a_step), bh = desc_lo_hi(bh0 + k * b_step);
    umma::mma_bf16(tmem_d, ah, bh, idesc, (k > 0 || accumulate) ? 1u : 0u);
    if (B_LO)
      umma::mma_bf16(tmem_d, ah, desc_lo_hi(bl0 + k * b_step), idesc, 1);
    if (A_LO)
      umma::mma_bf16(tmem_d, desc_lo_hi(al0 + k * a_step), bh, idesc, 1);
  }
}

__device__ __forceinline__ void sync_after_smem_writes() {
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
}

// The flat fp32 parameter vector (<= 32 KB) -> shared-memory scratch with ONE bulk copy
// (cp.async.bulk, completion on an mbarrier): the panel staging below then reads shared memory
// instead of paying an L2 round trip per staging loop at every launch. Whole CTA; returns the
// scratch pointer; contains __syncthreads. It is the FIRST access to global memory of every fused
// kernel: the programmatic-dependent-launch wait sits here (what precedes it -- tensor-memory
// allocation, mbarrier set-up -- may run while the previous kernel of the stream drains).
// In two halves, so that a kernel can put work that does not need the parameters (zero fills, the first
// tile's state loads) between the issue and the wait: stage_params_issue, __syncthreads, stage_params_wait.
__device__ __forceinline__ const float *stage_params_issue(const float *__restrict__ params, int n, uint8_t *scratch,
                                                           uint64_t *pbar) {
  umma::pdl_wait();
  float *dst = reinterpret_cast<float *>(scratch);
  const bool aligned = (reinterpret_cast<uintptr_t>(params) & 15) == 0;
  const uint32_t bytes = aligned ? ((uint32_t)n * 4u) & ~15u : 0u;
  if (threadIdx.x == 0) {
    umma::mbar_init(pbar, 1);
    umma::fence_mbar_init();
    if (bytes) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(umma::smem_u32(pbar)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                       umma::smem_u32(dst)), "l"(params), "r"(bytes), "r"(umma::smem_u32(pbar)) : "memory");
    }
  }
  for (int i = (int)(bytes / 4) + threadIdx.x; i < n; i += blockDim.x)  // tail (or everything if misaligned)
    dst[i] = params[i];
  return dst;
}
// After a __syncthreads that follows stage_params_issue (the mbarrier is initialised, the tail is visible).
__device__ __forceinline__ void stage_params_wait(const float *__restrict__ params, int n, uint64_t *pbar) {
  const bool aligned = (reinterpret_cast<uintptr_t>(params) & 15) == 0;
  if (aligned && (((uint32_t)n * 4u) & ~15u))
    umma::mbar_wait(pbar, 0);
}
__device__ __forceinline__ const float *stage_params_bulk(const float *__restrict__ params, int n, uint8_t *scratch,
                                                          uint64_t *pbar) {
  const float *dst = stage_params_issue(params, n, scratch, pbar);
  __syncthreads();  // the mbarrier is initialised (and the tail is visible) for every thread
  stage_params_wait(params, n, pbar);
  return dst;
}

// fp32 [N][K] row-major -> forward-format hi / lo panels [rows_alloc][64] of (scale * W), zero padded.
__device__ void stage_weight_f16(const float *__restrict__ W, int N, int K, int rows_alloc, float scale,
                                 uint8_t *hi, uint8_t *lo) {
  for (int c = threadIdx.x; c < rows_alloc * 8; c += blockDim.x) {
    int row = c >> 3, chunk = c & 7;
    float x[8];
    const float *src = W + (size_t)row * K + chunk * 8;
    if (row < N && chunk * 8 + 8 <= K && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
      float4 a = *reinterpret_cast<const float4 *>(src), b = *(reinterpret_cast<const float4 *>(src) + 1);
      x[0] = a.x * scale, x[1] = a.y * scale, x[2] = a.z * scale, x[3] = a.w * scale;
      x[4] = b.x * scale, x[5] = b.y * scale, x[6] = b.z * scale, x[7] = b.w * scale;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int col = chunk * 8 + j;
        x[j] = (row < N && col < K) ? W[(size_t)row * K + col] * scale : 0.f;
      }
    }
    uint4 h, l;
    split8<false>(x, h, l);
    uint32_t off = umma::panel_chunk_off(row, chunk);
    *reinterpret_cast<uint4 *>(hi + off) = h;
    *reinterpret_cast<uint4 *>(lo + off) = l;
  }
}

// idesc shorthands: F = forward-format operand, B = bf16 operand; K = K-major, M = MN-major
template <int N> struct ID {
  static constexpr int FB = 1;  // format code of the forward operands (bf16)
  static constexpr uint32_t FK_FK = make_idesc(128, N, FB, FB, 0, 0);  // forward: A fwd K, B fwd K
  static constexpr uint32_t BK_FM = make_idesc(128, N, 1, FB, 0, 1);   // dX: A bf16 K-major, B fwd MN-major
  static constexpr uint32_t BM_FM = make_idesc(128, N, 1, FB, 1, 1);   // dW: A = dH^T (bf16), B = H / X0 (fwd)
  static constexpr uint32_t FM_BM = make_idesc(128, N, FB, 1, 1, 1);   // dW3^T: A = H^T (fwd), B = dY (bf16)
  // M = 64 forms of the two weight-gradient GEMMs whose upper 64 output rows would be unused: half
  // the A-operand shared-memory traffic. D row r lives in TMEM lane 32 (r / 16) + r % 16.
  static constexpr uint32_t BM_FM_64 = make_idesc(64, N, 1, FB, 1, 1);
  static constexpr uint32_t FM_BM_64 = make_idesc(64, N, FB, 1, 1, 1);
};

// Learner rows of a tile (row = t * E + e). end_rows: observation of the END state of step (t, e):
// overflowed terminal state when done, live env state at the rollout's last step, zeros (unused)
// otherwise.
struct learner_rows {
  const int8_t *rec_state, *live_state;
  const uint8_t *rec_action, *rec_done;
  int n, stride, T, E, B;
  float inv_w, inv_h;
};
// ---- gradient tail: cross-CTA reduction (+ peer exchange) + optimizer update INSIDE the producing
// kernel. The learner kernels are persistent (one CTA per SM, launched cooperatively): after a CTA
// has written its partial gradient it arrives on a grid barrier; every CTA then owns a contiguous
// slice of ceil(P / grid) parameters, sums the grid's partials of that slice in a fixed order (bitwise
// reproducible), exchanges the slice sums with the peer ranks (PUSH, flag in the data, see below) and
// applies sgd / momentum / adam (nn.h:616-698) to its slice. This replaces a separate reduction
// launch per optimizer step (5 per PPO iteration) and the drain -> launch -> reduce hand-over.
struct p2p_view {
  float *peer[DFRL_P2P_MAX_RANKS];
  int nranks, rank;
};
struct grad_tail {
  float *grad;        // [P] reduced flat gradient (all ranks' sum when the exchange runs)
  unsigned *bar;      // [0] arrivals, [1] generation of the grid barrier (zero at creation)
  dfrl_opt_spec opt;  // params == null: gradient only (the caller runs NCCL + optimizer kernel)
  p2p_view v;         // nranks > 1 (and opt given): exchange over NVLink peer memory
};

struct critic_args {
  const float *params;  // flat fp32 parameters of the net
  net3 net;
  learner_rows rows;
  int n_tiles;
  float gamma, lambda;
  float *targets_out;  // [T][n] (critic step) -- introspection + parity
  float *adv_out;      // [T][n] (GAE kernel)
  const float *v_end;  // [T][n] V(end state) where a trajectory ends, from fused_vend_kernel; null: the
                       // kernel runs its own end pass over all rows of every tile
  float *v_end_out;    // GAE kernel, non-null (== v_end): every pipeline first evaluates the end rows of ITS
                       // OWN tiles, compacted into passes of 128 rows, and writes them here -- no pre-pass launch
  unsigned *adv_maxbits;  // GAE kernel, optional: atomicMax of the bits of max |advantage| (the fixed-point scale of
                          // the table path, conv_table.cuh; zeroed by the caller)
  float *partials;
  grad_tail tail;      // critic step only
  long long *clk;      // optional: phase clocks of CTA 0, pipeline 0 (debug; critic step only)
};

enum { HEAD_JACOBIAN = 0, HEAD_IDENTITY = 1 };

struct policy_step_args {
  const float *params;  // flat fp32 parameters of the net
  net3 net;
  learner_rows rows;
  const float *adv;    // [T][n]
  const float *p_old;  // [T][n][B]
  int n_tiles;
  int loss_kind, head_bwd;
  float *partials;     // [gridDim.x][n_params]
  grad_tail tail;
  long long *clk;      // optional: phase clocks of CTA clk_cta (debug)
  int clk_cta;
};

// ---- gradient tail (see grad_tail) ------------------------------------------------------------
// What a thread needs from global memory BEFORE its CTA arrives on the grid barrier: the barrier
// generation, the exchange number of this launch and the Adam step counter -- all three are advanced
// by the CTA that releases the barrier, i.e. only after every thread of every CTA has read them
// (the reads sit in front of the __syncthreads that precedes the arrival).
struct tail_ctx {
  unsigned gen, epoch;
  float c1, c2;
};
__device__ __forceinline__ tail_ctx tail_begin(const grad_tail &tl) {
  tail_ctx c;
  c.gen = *reinterpret_cast<volatile unsigned *>(tl.bar + 1);
  c.epoch = 0;
  if (tl.v.nranks > 1)
    c.epoch = reinterpret_cast<volatile unsigned *>(dfrl_p2p_flags(tl.v.peer[tl.v.rank]))[2] + 1;
  c.c1 = tl.opt.c1, c.c2 = tl.opt.c2;
  if (tl.opt.t_dev) {  // adam bias corrections from the device-side step counter (nn.h:683-684)
    const float t = *reinterpret_cast<volatile float *>(tl.opt.t_dev);
    c.c1 = 1.f - powf(tl.opt.beta1, t);
    c.c2 = 1.f - powf(tl.opt.beta2, t);
  }
  return c;
}
// Exchange over peer memory, PUSH protocol with the flag in the data: the owner of gradient entry i
// stores {value, exchange number} -- one 8-byte store, single-copy atomic -- into every rank's
// exchange buffer over NVLink (own rank included), then polls its LOCAL copies of all source ranks
// until their exchange number matches and sums them in rank order (bit-identical on every rank). No
// fence, no separate flag, nothing on the receive side crosses NVLink. Two slots: a rank is never more
// than one exchange ahead of a peer (it needs the peer's contribution to finish one). Time-bounded
// wait (a peer that died must turn into a reported launch failure, not a hung GPU).
__device__ __forceinline__ float exchange_entry(const p2p_view &v, unsigned epoch, int i, float r) {
  const int slot = (int)(epoch & 1u);
  const unsigned long long w = ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(r);
  for (int q = 0; q < v.nranks; ++q)
    *reinterpret_cast<volatile unsigned long long *>(dfrl_p2p_data(v.peer[q], slot, v.rank) + i) = w;
  float *local = v.peer[v.rank];
  unsigned long long x[DFRL_P2P_MAX_RANKS];
  unsigned long long t0 = 0;
  unsigned spins = 0;
  bool all;
  do {  // all loads in flight; repeat until every source rank's pair carries this exchange number
    all = true;
#pragma unroll
    for (int q = 0; q < DFRL_P2P_MAX_RANKS; ++q) {
      x[q] = q < v.nranks ? *reinterpret_cast<const volatile unsigned long long *>(dfrl_p2p_data(local, slot, q) + i)
                          : ((unsigned long long)epoch << 32);
      all = all && (unsigned)(x[q] >> 32) == epoch;
    }
    if (!all && (++spins & 0xfffu) == 0) {  // ranks may skew by seconds (graph instantiation, host tapes)
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(now));
      if (t0 == 0)
        t0 = now;
      else if (now - t0 > 120ull * 1000000000ull)
        __trap();
    }
  } while (!all);
  float g = 0.f;
#pragma unroll
  for (int q = 0; q < DFRL_P2P_MAX_RANKS; ++q)  // rank order; ranks >= nranks add +0
    g += __uint_as_float((unsigned)x[q]);
  return g;
}
// Row stride (floats) of the per-CTA partial gradients: rows are 16-byte aligned for the tail's loads.
__host__ __device__ __forceinline__ int partial_stride(int n_params) { return (n_params + 3) & ~3; }

// Whole CTA, after its partial gradient has been written to partials[blockIdx.x * partial_stride(n)][0..n).
// scratch: (blockDim.x / 16) * 64 floats of shared memory. Contains __syncthreads.
//
// The tail runs once per launch, i.e. on a cold instruction cache: round 2 measured 8 800 .. 17 000 cycles
// for "30 loads per thread + sum" when that was 3 KB of unrolled straight-line code (the whole kernel is
// 105 KB of SASS). Hence few, wide instructions: the CTA's slice is a multiple of 4 parameters, a thread
// loads 16-byte words (4 parameters of one CTA's partial) for <= 10 source CTAs -- all in flight at once
// -- and the source-CTA groups are combined through shared memory in a fixed order.
__device__ __forceinline__ void gradient_tail(const float *__restrict__ partials, const net3 &net, const grad_tail &tl,
                                              float *scratch, long long *clk = nullptr) {
  const int n = net.n_params, ps = partial_stride(n);
  const tail_ctx tc = tail_begin(tl);
  const int G = (int)gridDim.x, p = (int)threadIdx.x & 63;
  const int R = (int)blockDim.x >> 4, rg = (int)threadIdx.x >> 4, q4 = (int)threadIdx.x & 15;  // source-CTA group, 16-byte column
  const int per = (((n + G - 1) / G) + 3) & ~3;
  const int lo = (int)blockIdx.x * per, hi = min(lo + per, n);
  const dfrl_opt_spec &opt = tl.opt;
  const bool exchange = opt.params && tl.v.nranks > 1;
  // The parameter / optimizer-state entries of this CTA's slice belong to this CTA alone: the first
  // 64 are loaded BEFORE the grid barrier, so that the update behind the reduction costs no further
  // L2 round trip (the tail is a chain of round trips: barrier, partials, update).
  const bool updates = opt.params != nullptr;
  const bool owner = threadIdx.x < 64;
  float p0 = 0.f, m0 = 0.f, v0 = 0.f;
  if (updates && owner && lo + p < hi) {
    p0 = __ldcg(opt.params + lo + p);
    if (opt.kind != DFRL_OPT_SGD)
      m0 = __ldcg(opt.state + lo + p);
    if (opt.kind == DFRL_OPT_ADAM)
      v0 = __ldcg(opt.state + n + lo + p);
  }
  if (clk)
    clk[0] = clock64();
  __syncthreads();  // every thread's partial-gradient stores precede thread 0's fence (cumulativity)
  if (threadIdx.x == 0) {
    __threadfence();  // ... and are visible device-wide before this CTA arrives
    if (atomicAdd(tl.bar, 1u) == gridDim.x - 1) {  // last arrival: every CTA has read gen / epoch / t
      tl.bar[0] = 0;
      if (tl.v.nranks > 1)
        reinterpret_cast<volatile unsigned *>(dfrl_p2p_flags(tl.v.peer[tl.v.rank]))[2] = tc.epoch;
      if (tl.opt.params && tl.opt.t_dev)
        *tl.opt.t_dev += 1.f;  // nn.h:686
      __threadfence();
      atomicExch(tl.bar + 1, tc.gen + 1);  // release
    } else {
      unsigned spins = 0;
      while (*reinterpret_cast<volatile unsigned *>(tl.bar + 1) == tc.gen)
        if (++spins > (1u << 28))
          __trap();  // (one CTA per SM, grid <= SM count: every CTA is resident; bounded by the kernel's duration)
    }
    __threadfence();
  }
  __syncthreads();
  if (clk)
    clk[1] = clock64();
  constexpr int MAXR = 10;  // 16-byte loads in flight per thread: one L2 round trip for <= 10 R source CTAs (R >= 18)
  for (int base = lo; base < hi; base += 64) {
    const int i = base + p;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (base + 4 * q4 < hi) {
      const float4 *src = reinterpret_cast<const float4 *>(partials + base + 4 * q4);
      for (int c0 = rg; c0 < G; c0 += MAXR * R) {  // fixed order
        float4 x[MAXR];
#pragma unroll
        for (int q = 0; q < MAXR; ++q) {
          const int c = c0 + q * R;
          x[q] = c < G ? __ldcg(src + (size_t)c * (ps >> 2)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int q = 0; q < MAXR; ++q)
          acc.x += x[q].x, acc.y += x[q].y, acc.z += x[q].z, acc.w += x[q].w;
      }
    }
    reinterpret_cast<float4 *>(scratch)[rg * 16 + q4] = acc;
    __syncthreads();
    if (clk && base == lo)
      clk[2] = clock64();
    if (owner && i < hi) {
      float g = 0.f;
      for (int q = 0; q < R; ++q)  // source-CTA groups in order
        g += scratch[q * 64 + p];
      if (exchange)
        g = exchange_entry(tl.v, tc.epoch, i, g);
      tl.grad[i] = g;
      if (updates && (!net.shared || net_owns(net, i))) {  // (another head's slots: zero gradient, left alone)
        float pv = p0, mv = m0, vv = v0;
        if (base != lo) {
          pv = opt.params[i];
          if (opt.kind != DFRL_OPT_SGD)
            mv = opt.state[i];
          if (opt.kind == DFRL_OPT_ADAM)
            vv = opt.state[n + i];
        }
        opt_update_vals(opt.kind, pv, g, mv, vv, opt.lr, opt.wd, opt.beta1, opt.beta2, tc.c1, tc.c2);
        opt.params[i] = pv;
        if (opt.kind != DFRL_OPT_SGD)
          opt.state[i] = mv;
        if (opt.kind == DFRL_OPT_ADAM)
          opt.state[n + i] = vv;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// Policy step, two tile pipelines per CTA.
//
// A tile's work is a strictly dependent chain  MMA -> epilogue -> MMA -> ...  (8 GEMMs, 5 epilogues),
// so a CTA that runs one tile at a time leaves the tensor pipe idle during the epilogues and the
// ALUs idle during the GEMMs. Here each of the two warpgroups owns a whole tile (one thread per
// row, all columns) and runs the chain on its own: while one warpgroup is in an epilogue the
// other one's MMAs execute. Each warpgroup has its own MMA-issuing lane, mbarrier, named barrier,
// 256 TMEM columns (accumulators incl. its own dW partial sums: the drain adds the two in a fixed
// order) and five activation panels; what makes two tiles fit in 227 KB:
//   * X0 | 1 | dY share ONE panel: observations in bytes 0..63 of a row, the ones column (bias
//     gradients) at column 32, dY as [hi(8) | lo(8)] in bytes 96..127 = a single K = 16 step whose
//     B operand stacks [hi(W3); hi(W3)] (and [lo(W3); 0]): 2 MMAs instead of 3; the dW3 GEMM
//     (N = 16) yields H2^T dY_hi and H2^T dY_lo in separate columns, summed in the drain;
//   * dH2 overwrites H2 (after the dW3 GEMM, which is therefore issued BEFORE the dH2 GEMM under
//     the same commit);
//   * dH1 lives in one slot shared by both warpgroups, handed over by an mbarrier that the dW1
//     GEMM's tcgen05.commit arrives on (the slot is held for ~1/6 of a tile);
//   * hi(W1) / lo(W1) share a panel (K = 32 = 64 bytes each).
template <int D1, int D2>
struct pmap {
  static constexpr uint32_t W1P = 0;  // [D1 rows]: hi in bytes 0..63, lo in bytes 64..127
  static constexpr uint32_t W2_HI = W1P + D1 * 128, W2_LO = W2_HI + D2 * 128;
  static constexpr uint32_t W3A = W2_LO + D2 * 128;  // rows 0..7 = rows 8..15 = hi(W3)
  static constexpr uint32_t W3B = W3A + 16 * 128;    // rows 0..7 = lo(W3), rows 8..15 = 0
  static constexpr uint32_t W3C = W3B + 16 * 128;    // rows 0..7 = hi(W3), rows 8..15 = lo(W3): forward head
  static constexpr uint32_t FLOATS = W3C + 16 * 128;
  static constexpr int F_B1 = 0, F_B2 = D1, F_B3 = D1 + D2, N_FLOATS = D1 + D2 + 16;
  static constexpr uint32_t DH1_HI = (FLOATS + N_FLOATS * 4 + 1023) / 1024 * 1024;  // shared slot
  static constexpr uint32_t DH1_LO = DH1_HI + PANEL;
  static constexpr uint32_t WG0 = DH1_LO + PANEL;  // per-warpgroup blocks follow
  static constexpr uint32_t XD = 0, H1_HI = PANEL, H1_LO = 2 * PANEL, H2_HI = 3 * PANEL, H2_LO = 4 * PANEL;
  static constexpr uint32_t WG_BYTES = 5 * PANEL;
  static constexpr uint32_t BARS = WG0 + 2 * WG_BYTES;
  static constexpr uint32_t TOTAL = BARS + 128;
  static constexpr uint32_t DY_OFF = 96;  // byte offset of [dY_hi | dY_lo] in a row of the XD panel
  // Epilogue threads per row. Two (one 32-column chunk each, 576 threads) were measured SLOWER here:
  // 103 us vs 95.5 us per launch -- unlike the critic step, this kernel is bound by shared-memory
  // bandwidth (MMA operand fetch + epilogue stores) and the tensor pipe, not by instruction issue.
  static constexpr int NH = 1;
  static constexpr int THREADS = 32 * (4 * NH + 1) * 2;
  static_assert((D1 * 128) % 1024 == 0 && (D2 * 128) % 1024 == 0, "panel alignment");
  static_assert(TOTAL + 1024 <= 232448, "exceeds the 227 KB shared memory of an SM");
};
// TMEM columns of one warpgroup (base = 256 * wg)
constexpr uint32_t P2_ACC0 = 0, P2_ACC1 = 64, P2_DA = 128, P2_DB = 192, P2_DC = 240;

// fp32 W1 [D1][32] -> ONE panel: hi(W1) in bytes 0..63 of a row, lo(W1) in bytes 64..127.
template <int D1>
__device__ void stage_w1_packed(const float *__restrict__ W1, uint8_t *panel) {
  constexpr int D0 = 32;
  for (int c = threadIdx.x; c < D1 * 4; c += blockDim.x) {
    int row = c >> 2, chunk = c & 3;
    float x[8];
    const float4 *src = reinterpret_cast<const float4 *>(W1 + (size_t)row * D0 + chunk * 8);
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
      float4 p = *src, q = *(src + 1);
      x[0] = p.x, x[1] = p.y, x[2] = p.z, x[3] = p.w, x[4] = q.x, x[5] = q.y, x[6] = q.z, x[7] = q.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        x[j] = W1[(size_t)row * D0 + chunk * 8 + j];
    }
    uint4 h, l;
    split8<false>(x, h, l);
    *reinterpret_cast<uint4 *>(panel + umma::panel_chunk_off(row, chunk)) = h;
    *reinterpret_cast<uint4 *>(panel + umma::panel_chunk_off(row, 4 + chunk)) = l;
  }
}

template <int D0, int D1, int D2, int NOUT>
__device__ void build_policy_image(const float *__restrict__ params, const net3 &net, uint8_t *smem) {
  using PM = pmap<D1, D2>;
  const float *W1 = params + net.o_w1, *W2 = params + net.o_w2, *W3 = params + net.o_w3;
  static_assert(D0 == 32 && NOUT == 8, "packed W1 / stacked W3 panels assume 32 inputs, 8 outputs");
  stage_w1_packed<D1>(W1, smem + PM::W1P);
  stage_weight_f16(W2, D2, D1, D2, 1.f, smem + PM::W2_HI, smem + PM::W2_LO);
  for (int c = threadIdx.x; c < 16 * 8; c += blockDim.x) {
    int row = c >> 3, chunk = c & 7;
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int col = chunk * 8 + j;
      x[j] = col < D2 ? W3[(size_t)(row & 7) * D2 + col] : 0.f;
    }
    uint4 h, l;
    split8<false>(x, h, l);
    uint32_t off = umma::panel_chunk_off(row, chunk);
    *reinterpret_cast<uint4 *>(smem + PM::W3A + off) = h;
    *reinterpret_cast<uint4 *>(smem + PM::W3B + off) = row < 8 ? l : make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4 *>(smem + PM::W3C + off) = row < 8 ? h : l;
  }
  float *fl = reinterpret_cast<float *>(smem + PM::FLOATS);
  for (int i = threadIdx.x; i < D1; i += blockDim.x) fl[PM::F_B1 + i] = params[net.o_b1 + i];
  for (int i = threadIdx.x; i < D2; i += blockDim.x) fl[PM::F_B2 + i] = params[net.o_b2 + i];
  for (int i = threadIdx.x; i < 16; i += blockDim.x) fl[PM::F_B3 + i] = i < net.d3 ? params[net.o_b3 + i] : 0.f;
}

// issue_gemm with a run-time leading byte offset of the A operand (M = 128 spanning two panels
// that are not adjacent) and no lo pass of B.
template <int KSTEPS>
__device__ __forceinline__ void issue_gemm_mn_lbo(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t a_lbo,
                                                  uint32_t b_hi, uint32_t idesc, bool accumulate) {
  constexpr uint32_t step = umma::KSTEP_BYTES_MNMAJOR >> 4;
  const uint32_t ah0 = desc_lo(a_hi, a_lbo), al0 = desc_lo(a_lo, a_lbo), bh0 = desc_lo(b_hi, PANEL);
#pragma unroll
  for (int k = 0; k < KSTEPS; ++k) {
    uint64_t bh = desc_lo_hi(bh0 + k * step);
    umma::mma_bf16(tmem_d, desc_lo_hi(ah0 + k * step), bh, idesc, (k > 0 || accumulate) ? 1u : 0u);
    umma::mma_bf16(tmem_d, desc_lo_hi(al0 + k * step), bh, idesc, 1);
  }
}

// TMEM accumulator (this thread's row, all D columns) -> + bias, relu -> hi/lo panels. No relu mask
// is kept: the backward epilogue of the same thread reads it off the hi panel (H > 0 <=> hi(H) > 0).
// TMEM copy of an activation (A operand of the next GEMM read from tensor memory instead of shared
// memory: 2 KB instead of 6 KB of shared-memory traffic per M128.N64 instruction): the packed hi / lo
// words of a 32-column chunk h overwrite the chunk's own (already read) accumulator columns, hi of
// K step kk at column h * CH + 8 kk, lo at + CH / 2.
template <int CH>
__device__ __forceinline__ void tmem_put_chunk(uint32_t taddr, int cc, const uint4 &hh, const uint4 &ll) {
  // 8 columns of the tile = 4 packed words; cc = 8-column group inside the chunk
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(taddr + 4 * cc), "r"(hh.x),
               "r"(hh.y), "r"(hh.z), "r"(hh.w) : "memory");
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(taddr + CH / 2 + 4 * cc),
               "r"(ll.x), "r"(ll.y), "r"(ll.z), "r"(ll.w) : "memory");
}
// GEMM with the A operand (hi / lo pairs, layout of tmem_put_chunk) in tensor memory.
template <int D, int KSTEPS, bool B_MN, bool B_LO = true>
__device__ __forceinline__ void issue_gemm_ta(uint32_t tmem_d, uint32_t a_tmem, uint32_t b_hi, uint32_t b_lo,
                                              uint32_t idesc) {
  constexpr int CH = D < 32 ? D : 32;
  constexpr uint32_t b_step = (B_MN ? umma::KSTEP_BYTES_MNMAJOR : umma::KSTEP_BYTES_KMAJOR) >> 4;
  constexpr uint32_t b_lbo = B_MN ? PANEL : 16;
  const uint32_t bh0 = desc_lo(b_hi, b_lbo), bl0 = desc_lo(b_lo, b_lbo);
#pragma unroll
  for (int k = 0; k < KSTEPS; ++k) {
    const uint32_t ah = a_tmem + (16 * k / CH) * CH + ((16 * k % CH) / 16) * 8, al = ah + CH / 2;
    const uint64_t bh = desc_lo_hi(bh0 + k * b_step), bl = desc_lo_hi(bl0 + k * b_step);
    umma::mma_bf16_ta(tmem_d, ah, bh, idesc, k > 0 ? 1u : 0u);
    if (B_LO)
      umma::mma_bf16_ta(tmem_d, ah, bl, idesc, 1);
    umma::mma_bf16_ta(tmem_d, al, bh, idesc, 1);
  }
}

// [h0, h1): the 32-column chunks this thread handles (all of them with one thread per row; chunk
// `half` with two threads per row).
template <int D, bool TMEM_COPY = false, bool SMEM_STORE = true>
__device__ __forceinline__ void epi2_fwd(uint32_t acc, const tid_t &t, const float *__restrict__ bias, uint8_t *hi,
                                         uint8_t *lo, int h0 = 0, int h1 = D / (D < 32 ? D : 32)) {
  constexpr int CH = D < 32 ? D : 32;
#pragma unroll  // (rolling this loop was measured slower: the chunks' TMEM loads / stores no longer overlap)
  for (int h = h0; h < h1; ++h) {
    float v[CH];
    tmem_load<CH>(acc + t.lane_base + h * CH, v);
#pragma unroll
    for (int j4 = 0; j4 < CH; j4 += 4) {
      float4 b = *reinterpret_cast<const float4 *>(bias + h * CH + j4);
      v[j4] = fmaxf(v[j4] + b.x, 0.f);
      v[j4 + 1] = fmaxf(v[j4 + 1] + b.y, 0.f);
      v[j4 + 2] = fmaxf(v[j4 + 2] + b.z, 0.f);
      v[j4 + 3] = fmaxf(v[j4 + 3] + b.w, 0.f);
    }
#pragma unroll
    for (int cc = 0; cc < CH / 8; ++cc) {
      uint4 hh, ll;
      split8<false>(&v[8 * cc], hh, ll);
      if (SMEM_STORE) {
        const int c = h * (CH / 8) + cc;  // 16-byte chunk of the row; layers wider than 64: consecutive panels
        uint32_t off = (uint32_t)(c >> 3) * PANEL + umma::panel_chunk_off(t.row, c & 7);
        *reinterpret_cast<uint4 *>(hi + off) = hh;
        *reinterpret_cast<uint4 *>(lo + off) = ll;
      }
      if (TMEM_COPY)
        tmem_put_chunk<CH>(acc + t.lane_base + h * CH, cc, hh, ll);
    }
  }
  if (TMEM_COPY)
    umma::tmem_st_wait();
}
// 0xffff in each half whose bf16 value is > 0
__device__ __forceinline__ uint32_t pos_mask2(uint32_t w) {
  uint32_t m;
  asm("set.gt.u32.bf16x2 %0, %1, %2;\n" : "=r"(m) : "r"(w), "r"(0u));
  return m;
}
// TMEM accumulator -> . relu mask (taken from the forward activation's hi panel `act_hi`) -> hi/lo
// panels (input-gradient epilogue). `act_hi` may be the destination `hi` itself: every thread reads
// a 16-byte chunk before it overwrites that same chunk.
template <int D, bool TMEM_COPY = false>
__device__ __forceinline__ void epi2_bwd(uint32_t acc, const tid_t &t, const uint8_t *act_hi, uint8_t *hi,
                                         uint8_t *lo, int h0 = 0, int h1 = D / (D < 32 ? D : 32)) {
  constexpr int CH = D < 32 ? D : 32;
#pragma unroll
  for (int h = h0; h < h1; ++h) {
    float v[CH];
    tmem_load<CH>(acc + t.lane_base + h * CH, v);
#pragma unroll
    for (int cc = 0; cc < CH / 8; ++cc) {
      const int c = h * (CH / 8) + cc;
      uint32_t off = (uint32_t)(c >> 3) * PANEL + umma::panel_chunk_off(t.row, c & 7);
      const uint4 aw = *reinterpret_cast<const uint4 *>(act_hi + off);
      uint4 hh, ll;
      split8<false>(&v[8 * cc], hh, ll);
      const uint32_t m0 = pos_mask2(aw.x), m1 = pos_mask2(aw.y), m2 = pos_mask2(aw.z), m3 = pos_mask2(aw.w);
      hh = make_uint4(hh.x & m0, hh.y & m1, hh.z & m2, hh.w & m3);
      ll = make_uint4(ll.x & m0, ll.y & m1, ll.z & m2, ll.w & m3);
      *reinterpret_cast<uint4 *>(hi + off) = hh;
      *reinterpret_cast<uint4 *>(lo + off) = ll;
      if (TMEM_COPY)
        tmem_put_chunk<CH>(acc + t.lane_base + h * CH, cc, hh, ll);
    }
  }
  if (TMEM_COPY)
    umma::tmem_st_wait();
}

// The raw start state (2B + 2 int8 planes) of one learner row, global -> registers. Nothing touches
// the loaded values before encode_row: the loads stay in flight behind the tile's math.
template <int B>
struct row_state {
  int v[2 * B + 2];
};
template <int B>
__device__ __forceinline__ void load_row_state(const learner_rows &L, int tile, int row, row_state<B> &x) {
  constexpr int P = 2 * B + 2;
  const int tt = row / L.E, e = row % L.E, i = tile * L.E + e;
  const bool ok = tt < L.T && i < L.n;
  const int8_t *src = L.rec_state + (size_t)tt * P * L.stride + i;
#pragma unroll
  for (int q = 0; q < P; ++q) {
    x.v[q] = 0;
    if (ok)
      x.v[q] = src[(size_t)q * L.stride];
  }
}
// observation::to_vector (bin_packing.h:31-40) of this row into bytes 0..63 of its XD panel row.
template <int B>
__device__ __forceinline__ void encode_row(uint8_t *xd, int row, const row_state<B> &x, float inv_w, float inv_h) {
  float v[2 * B + 2];
#pragma unroll
  for (int q = 0; q < 2 * B + 2; ++q)
    v[q] = (float)x.v[q] * ((q & 1) ? inv_h : inv_w);
  const uint32_t it = pack2_fwd(v[2 * B], v[2 * B + 1]);
#pragma unroll
  for (int ch = 0; ch < B / 2; ++ch)
    *reinterpret_cast<uint4 *>(xd + umma::panel_chunk_off(row, ch)) =
        make_uint4(pack2_fwd(v[4 * ch], v[4 * ch + 1]), it, pack2_fwd(v[4 * ch + 2], v[4 * ch + 3]), it);
}

// Operands-ready handshake between a warpgroup's 128 epilogue threads (arrive) and its MMA-issuing
// warp (sync): named barriers 1 + 2 wg + parity, 160 threads. Two alternating ids: an epilogue
// thread is never more than one hand-over ahead of the issuer.
__device__ __forceinline__ void ready_arrive(int wg, uint32_t &parity, int threads = 160) {
  umma::fence_proxy_async();
  umma::fence_before_sync();
  asm volatile("bar.arrive %0, %1;\n" ::"r"(1 + 2 * wg + (int)parity), "r"(threads) : "memory");
  parity ^= 1;
}
__device__ __forceinline__ void ready_sync(int wg, uint32_t &parity, int threads = 160) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(1 + 2 * wg + (int)parity), "r"(threads) : "memory");
  parity ^= 1;
  umma::fence_after_sync();
}

// 320 threads: warps 0..3 / 4..7 = the epilogue threads of pipeline 0 / 1 (one thread per tile row),
// warps 8 / 9 = their MMA issuers. tcgen05.mma issue blocks the issuing thread for about the pipe
// time of the instruction (measured: tools/mma_microbench.py), so a GEMM that is meant to run behind
// an epilogue must not be issued by a thread that takes part in that epilogue.
// PROBE: the phase-clock / load-balance stamps of tools/policy_phase_clocks.py. A separate instantiation:
// the stamp code costs 1.3 % of the iteration when it is merely compiled in (measured, in-run A/B).
template <int D0, int D1, int D2, int NOUT, bool PROBE = false>
__global__ void __launch_bounds__((pmap<D1, D2>::THREADS), 1) fused_policy_step_kernel(policy_step_args a) {
  using PM = pmap<D1, D2>;
  constexpr int NH = PM::NH;          // epilogue threads per row
  constexpr int RT = 32 + 128 * NH;   // threads of an operands-ready hand-over
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);  // stays a shared-space pointer
  const float *fl = reinterpret_cast<const float *>(smem + PM::FLOATS);
  // mbarriers: [wg] MMA completion on the chain, [2] dH1 slot free, [3 + wg] dW2 GEMM done (H1 free),
  // [5 + wg] dW1 GEMM done (XD free)
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + PM::BARS);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + PM::BARS + 56);
  const net3 net = a.net;
  const learner_rows &L = a.rows;
  const tid_t t = thread_id();
  const bool issuer = t.warp >= 8 * NH;                          // warp-uniform
  const int wg = issuer ? t.warp - 8 * NH : t.warp / (4 * NH);  // pipeline index
  const int half = NH == 2 ? (t.warp >> 2) & 1 : 0;             // which 32-column chunk (NH = 2)
  const uint32_t sbase = umma::smem_u32(smem);
  const long long clk_entry = clock64();

  umma::pdl_launch_dependents();
  if (t.warp == 0)
    umma::tmem_alloc(tmem_slot, 512);
  if (threadIdx.x == 0) {
    for (int q = 0; q < 7; ++q)
      umma::mbar_init(bars + q, 1);
    umma::fence_mbar_init();
  }
  // The parameter copy is issued first; what does not need it flies behind it: the first tile's state loads
  // (an exposed L2 round trip otherwise) and the zero fill.
  const float *Pm = stage_params_issue(a.params, net.n_params, smem + PM::WG0 + PM::H1_HI, bars + 8);
  // this CTA's tiles: blockIdx.x + j * gridDim.x, j < nt; pipeline wg takes j = wg, wg + 2, ...
  const int nt = (int)blockIdx.x < a.n_tiles ? (a.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  row_state<NOUT> xr;  // raw state of the pipeline's current tile (epilogue threads)
  if (!issuer && half == NH - 1 && wg < nt)
    load_row_state<NOUT>(L, blockIdx.x + wg * gridDim.x, t.row, xr);
  // 64-wide layers: every panel byte an MMA reads is written first (epilogues cover all 128 rows x
  // 64 columns), except columns 32..47 of the XD panels ([1 | 0] for the bias gradients): zeroing
  // 32 KB instead of 192 KB takes ~0.8 us off every launch (the XD panels are not part of the scratch)
  if (D1 == 64 && D2 == 64) {
    zero_bytes(smem + PM::WG0 + PM::XD, PANEL);
    zero_bytes(smem + PM::WG0 + PM::WG_BYTES + PM::XD, PANEL);
  }
  __syncthreads();  // the copy's mbarrier is initialised, the tail of the vector and the zeros are visible
  stage_params_wait(a.params, net.n_params, bars + 8);
  build_policy_image<D0, D1, D2, NOUT>(Pm, net, smem);
  __syncthreads();  // the scratch (pipeline 0's H1 panels) is free again
  if (!(D1 == 64 && D2 == 64)) {
    zero_bytes(smem + PM::DH1_HI, PM::BARS - PM::DH1_HI);
    __syncthreads();
  }
  // ones column (col D0) of both XD panels: [dH1|dH2]^T . 1 = bias gradients for free
  if (!issuer && half == 0)
    *reinterpret_cast<uint16_t *>(smem + PM::WG0 + wg * PM::WG_BYTES + PM::XD + umma::panel_off(t.row, D0)) = 0x3F80;
  sync_after_smem_writes();
  const uint32_t tmem = *tmem_slot;

  uint8_t *wsm = smem + PM::WG0 + wg * PM::WG_BYTES;
  const uint32_t wbase = sbase + PM::WG0 + wg * PM::WG_BYTES;
  const uint32_t tm = tmem + 256u * wg;
  const uint32_t dh1_lbo = PM::WG0 + wg * PM::WG_BYTES + PM::H2_HI - PM::DH1_HI;  // dH1 panel -> own dH2 panel
  uint64_t *bar = bars + wg, *bar_dw2 = bars + 3 + wg, *bar_dw1 = bars + 5 + wg;
  uint32_t rp = 0;  // parity of the operands-ready barrier

  long long *clk = (PROBE && a.clk && (int)blockIdx.x == a.clk_cta && threadIdx.x == 96) ? a.clk : nullptr;  // pipeline 0, chunk 0
  // every CTA: %globaltimer (ns) at entry / end of the tile loop / end of the kernel, and its SM (load balance)
  long long *gclk = (PROBE && a.clk && threadIdx.x == 96 && blockIdx.x < 160) ? a.clk + 112 + 4 * blockIdx.x : nullptr;
  if (gclk)
    gclk[0] = global_ns(), gclk[3] = sm_id();
  int clk_n = 0;
#define STAMP() do { if (PROBE && clk && clk_n < 104) clk[clk_n++] = clock64(); } while (0)
  if (clk)
    clk[104] = clk_entry, clk[105] = clock64();

  float db3[NOUT];
#pragma unroll
  for (int q = 0; q < NOUT; ++q)
    db3[q] = 0.f;

  if (issuer) {
    // ================= MMA issuer of pipeline wg: one GEMM (group) per operands-ready hand-over
    bool first = true;
    if (wg < nt) {
      ready_sync(wg, rp, RT);  // X0 of the first tile staged in the H1_LO panel
      if (umma::elect_one()) {
        issue_gemm<D0 / 16, false, false, false, true>(tm + P2_ACC0, wbase + PM::H1_LO, 0, sbase + PM::W1P,
                                                       sbase + PM::W1P + 64, ID<D1>::FK_FK, false);
        umma::commit(bar);
      }
      __syncwarp();
    }
    for (int j = wg; j < nt; j += 2) {
      ready_sync(wg, rp, RT);  // H1
      if (umma::elect_one()) {  // A = H1 from tensor memory (the epilogue's copy in ACC0)
        issue_gemm_ta<D1, D1 / 16, false>(tm + P2_ACC1, tm + P2_ACC0, sbase + PM::W2_HI, sbase + PM::W2_LO,
                                          ID<D2>::FK_FK);
        umma::commit(bar);
      }
      __syncwarp();
      ready_sync(wg, rp, RT);  // H2
      if (umma::elect_one()) {
        // (hi(H2) + lo(H2)) . [hi(W3); lo(W3)], A = H2 from tensor memory (ACC1): two MMAs per K step, the
        // head adds columns q (hi.hi + lo.hi) and 8 + q (hi.lo + lo.lo)
        issue_gemm_ta<D2, D2 / 16, false, false>(tm + P2_ACC0, tm + P2_ACC1, sbase + PM::W3C, 0, ID<16>::FK_FK);
        umma::commit(bar);
      }
      __syncwarp();
      ready_sync(wg, rp, RT);  // dY
      // dW3^T += H2^T . [dY_hi | dY_lo] (M = 64, N = 16) first: the dH2 epilogue overwrites H2;
      // dH2 = [dY_hi | dY_lo] . [hi(W3); hi(W3)] + [dY_hi | dY_lo] . [lo(W3); 0]
      if (umma::elect_one()) {
        issue_gemm<8, true, true, true, false>(tm + P2_DC, wbase + PM::H2_HI, wbase + PM::H2_LO,
                                               wbase + PM::XD + PM::DY_OFF, 0, ID<16>::FM_BM_64, !first);
        issue_gemm<1, false, true, false, true>(tm + P2_ACC0, wbase + PM::XD + PM::DY_OFF, 0, sbase + PM::W3A,
                                                sbase + PM::W3B, ID<D2>::BK_FM, false);
        umma::commit(bar);
      }
      __syncwarp();
      ready_sync(wg, rp, RT);  // dH2 (in the H2 slot)
      // dH1 = dH2 . W2; dW2 += dH2^T . H1 (M = 64) runs behind the dH1 epilogue
      if (umma::elect_one()) {
        // A = dH2 from tensor memory (the epilogue's copy in ACC0)
        issue_gemm_ta<D2, D2 / 16, true>(tm + P2_ACC1, tm + P2_ACC0, sbase + PM::W2_HI, sbase + PM::W2_LO, ID<D1>::BK_FM);
        umma::commit(bar);
        issue_gemm<8, true, true, true, true>(tm + P2_DA, wbase + PM::H2_HI, wbase + PM::H2_LO, wbase + PM::H1_HI,
                                              wbase + PM::H1_LO, ID<D1>::BM_FM_64, !first);
        umma::commit(bar_dw2);
      }
      __syncwarp();
      ready_sync(wg, rp, RT);  // dH1 (shared slot) and the next tile's X0 (H1_LO panel)
      // layer 1 of the NEXT tile goes first: the in-order pipe would otherwise put this tile's dW1
      // GEMM on the next tile's critical path; dW1 runs behind the next tile's first epilogue.
      //   DB[128 x D0+16] += [dH1|dH2]^T . [X0|1]   rows 0.. = [dW1 | db1], rows 64.. col D0 = db2
      if (umma::elect_one()) {
        if (j + 2 < nt) {
          issue_gemm<D0 / 16, false, false, false, true>(tm + P2_ACC0, wbase + PM::H1_LO, 0, sbase + PM::W1P,
                                                         sbase + PM::W1P + 64, ID<D1>::FK_FK, false);
          umma::commit(bar);
        }
        issue_gemm_mn_lbo<8>(tm + P2_DB, sbase + PM::DH1_HI, sbase + PM::DH1_LO, dh1_lbo, wbase + PM::XD,
                             ID<D0 + 16>::BM_FM, !first);
        umma::commit(bar_dw1);
        umma::commit(bars + 2);
      }
      __syncwarp();
      first = false;
    }
  } else {
    // ================= epilogue threads of pipeline wg: thread = one row of the tile
    uint32_t phase = 0, phase_dw2 = 0, phase_dw1 = 0;
    auto wait_mma = [&]() {
      umma::mbar_wait(bar, phase);
      phase ^= 1;
      umma::fence_after_sync();
    };
    // Two threads per row (NH = 2): each handles one 32-column chunk in the hidden-layer epilogues;
    // chunk 0's threads also run the head (softmax, loss gradient) and keep db3, chunk 1's threads
    // prefetch the next tile's state and encode the observations.
    const int h0 = NH == 2 ? half : 0, h1d1 = NH == 2 ? half + 1 : D1 / (D1 < 32 ? D1 : 32),
              h1d2 = NH == 2 ? half + 1 : D2 / (D2 < 32 ? D2 : 32);
    const bool header = half == 0, stager = half == NH - 1;
    bool first = true;
    row_state<NOUT> xn;  // raw state of the next tile (xr: this tile's, loaded behind the parameter copy)
    // The observations of a tile are encoded twice: into the (dead) H1_LO panel for the layer-1
    // GEMM, so that the tile can start while the previous tile's dW1 GEMM still reads its XD panel,
    // and, behind the layer-2 GEMM, into the XD panel for this tile's own dW1 GEMM.
    if (wg < nt) {
      if (stager)
        encode_row<NOUT>(wsm + PM::H1_LO, t.row, xr, L.inv_w, L.inv_h);
      ready_arrive(wg, rp, RT);
    }
    for (int j = wg; j < nt; j += 2) {
      const int tile = blockIdx.x + j * gridDim.x;
      STAMP();
      // global loads that are consumed later: row data for the head, the next tile's state
      const int tt = t.row / L.E, e = t.row % L.E;
      const int i = tile * L.E + e;
      const bool valid = header && tt < L.T && i < L.n;
      const size_t k = (size_t)tt * L.n + i;
      int act = 0;
      float A = 0.f;
      float4 po[NOUT / 4];
#pragma unroll
      for (int q = 0; q < NOUT / 4; ++q)
        po[q] = make_float4(1.f, 1.f, 1.f, 1.f);
      if (valid) {
        act = L.rec_action[k];
        A = a.adv[k];
        const float4 *pr = reinterpret_cast<const float4 *>(a.p_old + k * NOUT);
#pragma unroll
        for (int q = 0; q < NOUT / 4; ++q)
          po[q] = pr[q];
      }
      const bool has_next = j + 2 < nt;
      if (has_next && stager)
        load_row_state<NOUT>(L, tile + 2 * gridDim.x, t.row, xn);
      // The two forward epilogues (and, below, the two backward ones) are ONE copy of code run twice with
      // affine arguments: the epilogue threads' loop body was 36 KB of straight-line SASS, which the SM's
      // instruction cache does not hold -- epilogue phases ran 20 - 30 % longer on the SMs far from the L2
      // (the only SM-position-dependent cost of this kernel: 2 650 .. 3 000 ns per tile across the chip).
      static_assert(D1 == D2 && P2_ACC1 == P2_ACC0 + 64 && PM::H2_HI == PM::H1_HI + 2 * PANEL && PM::H2_LO == PM::H1_LO + 2 * PANEL &&
                    PM::F_B2 == PM::F_B1 + D1, "the forward / backward epilogue loops assume equal widths and affine operand slots");
#pragma unroll 1
      for (int ph = 0; ph < 2; ++ph) {
        wait_mma();  // layer 1, layer 2
        STAMP();
        epi2_fwd<D1, true>(tm + P2_ACC0 + 64 * ph, t, fl + PM::F_B1 + D1 * ph, wsm + PM::H1_HI + 2 * PANEL * ph,
                           wsm + PM::H1_LO + 2 * PANEL * ph, h0, h1d1);
        ready_arrive(wg, rp, RT);
        if (ph == 0) {
          if (!first) {  // the previous tile's dW1 GEMM (XD, dH2 in the H2 slot) ran behind this epilogue
            umma::mbar_wait(bar_dw1, phase_dw1);
            phase_dw1 ^= 1;
          }
          if (stager)
            encode_row<NOUT>(wsm + PM::XD, t.row, xr, L.inv_w, L.inv_h);
        }
        STAMP();
      }
      wait_mma();  // layer 3
      STAMP();
      // ---- head epilogue: softmax, loss gradient, softmax backward -> dY = [hi | lo] in the XD panel
      if (header) {
        float v[16];
        tmem_load<16>(tm + P2_ACC0 + t.lane_base, v);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          v[q] += v[8 + q];
        float dl[8];
#pragma unroll
        for (int q = 0; q < 8; ++q)
          dl[q] = 0.f;
        if (valid) {
          const float *b3 = fl + PM::F_B3;
          float p[NOUT], s = 0.f;
#pragma unroll
          for (int q = 0; q < NOUT; ++q) {
            p[q] = expf(v[q] + b3[q]);  // no max subtraction (nn.h:382-392)
            s += p[q];
          }
          const float inv_s = 1.f / s;
#pragma unroll
          for (int q = 0; q < NOUT; ++q)
            p[q] = p[q] * inv_s;
          float g[NOUT];
          if (a.loss_kind == DFRL_LOSS_CLIPPED) {
            float pa = 0.f, pold = 1.f;
            const float *pof = reinterpret_cast<const float *>(po);
#pragma unroll
            for (int q = 0; q < NOUT; ++q) {
              pa = (q == act) ? p[q] : pa;
              pold = (q == act) ? pof[q] : pold;
            }
            float gc = clipped_grad(pa, pold, A);
#pragma unroll
            for (int q = 0; q < NOUT; ++q)
              g[q] = (q == act) ? gc : 0.f;
          } else {
#pragma unroll
            for (int q = 0; q < NOUT; ++q)
              g[q] = p[q] * A - (q == act ? A : 0.f);
          }
          if (a.head_bwd == HEAD_JACOBIAN) {
            float dot = 0.f;
#pragma unroll
            for (int q = 0; q < NOUT; ++q)
              dot = fmaf(p[q], g[q], dot);
#pragma unroll
            for (int q = 0; q < NOUT; ++q)
              dl[q] = p[q] * (g[q] - dot);
          } else {
#pragma unroll
            for (int q = 0; q < NOUT; ++q)
              dl[q] = g[q];
          }
#pragma unroll
          for (int q = 0; q < NOUT; ++q)
            db3[q] += dl[q];
        }
        uint4 h, l;
        split8<false>(dl, h, l);
        *reinterpret_cast<uint4 *>(wsm + PM::XD + umma::panel_chunk_off(t.row, 6)) = h;
        *reinterpret_cast<uint4 *>(wsm + PM::XD + umma::panel_chunk_off(t.row, 7)) = l;
      }
      ready_arrive(wg, rp, RT);
      STAMP();
#pragma unroll 1
      for (int ph = 0; ph < 2; ++ph) {
        wait_mma();  // dW3 + dH2; dH1
        // the shared dH1 slot: free once the previous tile of this CTA (the other pipeline's) has
        // finished its dW1 GEMM (completion j - 1 of bars[2])
        if (ph == 1 && j > 0)
          umma::mbar_wait(bars + 2, (uint32_t)(j - 1) & 1u);
        STAMP();
        // dH2 (mask from H2's hi panel) overwrites H2; dH1 (mask from H1) goes to the shared slot. Both leave
        // their packed copy in tensor memory (the A operand of the dH1 GEMM; dH1's copy is never read)
        uint8_t *dst = ph ? smem + PM::DH1_HI : wsm + PM::H2_HI;
        epi2_bwd<D1, true>(tm + P2_ACC0 + 64 * ph, t, wsm + PM::H2_HI - 2 * PANEL * ph, dst, dst + PANEL, h0, h1d1);
        if (ph == 0)
          ready_arrive(wg, rp, RT);
        STAMP();
      }
      umma::mbar_wait(bar_dw2, phase_dw2);  // H1 is free (the dW2 GEMM ran behind the dH1 epilogue)
      phase_dw2 ^= 1;
      STAMP();
      if (has_next && stager) {
        encode_row<NOUT>(wsm + PM::H1_LO, t.row, xn, L.inv_w, L.inv_h);
        xr = xn;
      }
      ready_arrive(wg, rp, RT);
      STAMP();
      first = false;
    }
    if (!first) {  // the last tile's dW1 GEMM
      umma::mbar_wait(bar_dw1, phase_dw1);
      umma::fence_after_sync();
    }
  }
#undef STAMP
  if (clk)
    clk[106] = clock64();
  if (gclk)
    gclk[1] = global_ns();

  // ---- drain: partial gradient of this CTA (warpgroup 0's sums + warpgroup 1's) -> global.
  // The TMEM accumulators are read row-wise (lane = a weight row), so direct stores would scatter 4-byte
  // words over 16 .. 32 cache lines per instruction (measured: 9 300 cycles for 27 KB). The values are
  // staged in shared memory instead (pipeline 1's dead panels, logical parameter order, one pad word per
  // 32: conflict-free for row strides 32 and 64) and leave as fully coalesced stores.
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  float *part = a.partials + (size_t)blockIdx.x * partial_stride(net.n_params);
  float *sg = reinterpret_cast<float *>(smem + PM::WG0 + PM::WG_BYTES);
  static_assert(PM::WG_BYTES >= 8448 * 4, "staging area of the partial gradient");
  auto put = [&](int i, float v) { sg[i + (i >> 5)] = v; };
  const bool two = nt > 1;  // warpgroup 1 had at least one tile
  if (nt == 0) {
    for (int q = threadIdx.x; q < net.n_params; q += blockDim.x)
      part[q] = 0.f;
  } else {
    if (net.shared) {  // the other heads' slots of the flat gradient
      for (int q = threadIdx.x; q < net.n_params; q += blockDim.x)
        if (!net_owns(net, q))
          put(q, 0.f);
    }
    // dW2[n][k]: DA (M = 64) row n = TMEM lane 32 (n / 16) + n % 16, col k
    const bool drainer = threadIdx.x < 256;  // 256 threads read the TMEM accumulators (t.wg = 0, 1)
    if (drainer) {
      constexpr int DC = D1 / 2;
      float v[DC], w[DC];
      tmem_load<DC>(tmem + P2_DA + t.lane_base + t.wg * DC, v);
      if (two) {
        tmem_load<DC>(tmem + 256 + P2_DA + t.lane_base + t.wg * DC, w);
#pragma unroll
        for (int q = 0; q < DC; ++q)
          v[q] += w[q];
      }
      int nrow = t.lane < 16 ? t.w * 16 + t.lane : -1;
      if (nrow >= 0 && nrow < D2)
#pragma unroll
        for (int q = 0; q < DC; ++q)
          put(net.o_w2 + nrow * D1 + t.wg * DC + q, v[q]);
    }
    // dW1[n][k] + db1[n]: DB row n, cols 0..D0-1 and D0; db2[n]: DB row 64 + n, col D0
    if (drainer) {
      constexpr int DC = (D0 + 16) / 2;
      float v[DC], w[DC];
      tmem_load<DC>(tmem + P2_DB + t.lane_base + t.wg * DC, v);
      if (two) {
        tmem_load<DC>(tmem + 256 + P2_DB + t.lane_base + t.wg * DC, w);
#pragma unroll
        for (int q = 0; q < DC; ++q)
          v[q] += w[q];
      }
#pragma unroll
      for (int q = 0; q < DC; ++q) {
        int col = t.wg * DC + q;
        if (t.row < D1) {
          if (col < D0)
            put(net.o_w1 + t.row * D0 + col, v[q]);
          else if (col == D0)
            put(net.o_b1 + t.row, v[q]);
        } else if (t.row >= 64 && t.row - 64 < D2 && col == D0) {
          put(net.o_b2 + t.row - 64, v[q]);
        }
      }
    }
    // dW3[n][k] = DC (M = 64) row k, cols n (H2^T dY_hi) and 8 + n (H2^T dY_lo)
    if (drainer && t.wg == 0) {
      float v[16], w[16];
      tmem_load<16>(tmem + P2_DC + t.lane_base, v);
      if (two) {
        tmem_load<16>(tmem + 256 + P2_DC + t.lane_base, w);
#pragma unroll
        for (int q = 0; q < 16; ++q)
          v[q] += w[q];
      }
      int krow = t.lane < 16 ? t.w * 16 + t.lane : -1;
      if (krow >= 0 && krow < D2)
#pragma unroll
        for (int q = 0; q < NOUT; ++q)
          put(net.o_w3 + q * D2 + krow, v[q] + v[8 + q]);
    }
    // db3: fixed-order tree inside each warp, then the epilogue warps in order (scratch = pipeline 0's H1)
    float *red = reinterpret_cast<float *>(smem + PM::WG0 + PM::H1_HI);
#pragma unroll
    for (int q = 0; q < NOUT; ++q) {
      float s = db3[q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
        s += __shfl_xor_sync(0xffffffffu, s, o);
      if (t.lane == 0 && !issuer)
        red[t.warp * 8 + q] = s;  // (warps that did not run the head contribute zeros)
    }
    __syncthreads();
    if (threadIdx.x < NOUT) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8 * NH; ++w)
        s += red[w * 8 + threadIdx.x];
      put(net.o_b3 + threadIdx.x, s);
    }
    __syncthreads();
    for (int q = threadIdx.x; q < net.n_params; q += blockDim.x)
      part[q] = sg[q + (q >> 5)];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (t.warp == 0)
    umma::tmem_dealloc(tmem, 512);
  // ---- cross-CTA reduction (+ exchange) + optimizer update of this CTA's parameter slice
  gradient_tail(a.partials, net, a.tail, reinterpret_cast<float *>(smem + PM::WG0 + PM::H1_HI), clk ? clk + 108 : nullptr);
  if (clk)
    clk[107] = clock64();
  if (gclk)
    gclk[2] = global_ns();
}

// ---------------------------------------------------------------------------------------------
// Critic step (update_value_model, policy_gradient.h:196-218, minus the optimizer update) and GAE
// (calculate_advantage, 220-281) with the structure of fused_policy_step_kernel: two tile
// pipelines per CTA, one epilogue thread per row, one MMA-issuing warp per pipeline.
// Per tile: V(end rows) [layer 1, layer 2, value head in fp32 registers], V(start rows), exchange
// of the values inside the tile, then
//   CRITIC_STEP: targets r + gamma V_next (unmasked, quirk 6), dY = V - target, dH2 = dY w3 . relu'
//                (rank 1, registers), dW3 / db3 (registers), dH1 GEMM, dW2 (M = 64) and
//                [dW1|db1|db2] GEMMs accumulated in TMEM;
//   CRITIC_GAE:  thread e < E walks its environment backwards.
// Shared memory per pipeline: XS (start observations | 1), H1 hi/lo, dH2 hi/lo; the end-row
// observations are staged in the dead H1_LO panel; dH1 in one slot shared by both pipelines.
enum { CRITIC_STEP = 0, CRITIC_GAE = 1 };
template <int D1, int D2, int MODE>
struct cmap {
  // CRITIC_STEP: 2 pipelines x 5 panels + the shared dH1 slot; CRITIC_GAE (forward only, hidden
  // activations in tensor memory): 3 pipelines x one observation panel (end rows, then start rows)
  static constexpr int NP = MODE == 0 ? 2 : 4;
  // CRITIC_STEP with 64-wide layers: TWO epilogue threads per row (one 32-column chunk each): the
  // critic's epilogues are instruction-heavy and two pipelines of 4 warps leave the SM's issue slots
  // two thirds empty
  static constexpr int NH = (MODE == 0 && D1 == 64 && D2 == 64) ? 2 : 1;
  static constexpr int THREADS = 32 * (4 * NH + 1) * NP;
  static constexpr uint32_t W1P = 0;
  static constexpr uint32_t W2_HI = W1P + D1 * 128, W2_LO = W2_HI + D2 * 128;
  static constexpr uint32_t FLOATS = W2_LO + D2 * 128;  // b1[D1] b2[D2] w3[64] b3[4]
  static constexpr int F_B1 = 0, F_B2 = D1, F_W3 = D1 + D2, F_B3 = D1 + D2 + 64, N_FLOATS = D1 + D2 + 68;
  static constexpr uint32_t SCR = FLOATS + N_FLOATS * 4;  // per pipeline: ve[NH][128], vs[NH][128] (partial values)
  static constexpr uint32_t DH1_HI = (SCR + NP * 2 * NH * TILE * 4 + 1023) / 1024 * 1024;  // shared slot
  static constexpr uint32_t DH1_LO = DH1_HI + PANEL;
  static constexpr uint32_t WG0 = MODE == 0 ? DH1_LO + PANEL : DH1_HI;
  // CRITIC_STEP: XS, H1 hi/lo (lo = staging panel of the end-row observations), dH2 hi/lo
  static constexpr uint32_t XS = 0, H1_HI = PANEL, H1_LO = MODE == 0 ? 2 * PANEL : 0, G2_HI = 3 * PANEL,
                            G2_LO = 4 * PANEL;
  static constexpr uint32_t WG_BYTES = MODE == 0 ? 5 * PANEL : PANEL;
  static constexpr uint32_t BARS = WG0 + NP * WG_BYTES;
  static constexpr uint32_t TOTAL = BARS + 256;
  static constexpr uint32_t TCOLS_PER = MODE == 0 ? 256 : 128;  // TMEM columns of a pipeline
  static_assert((D1 * 128) % 1024 == 0 && (D2 * 128) % 1024 == 0, "panel alignment");
  static_assert(TOTAL + 1024 <= 232448, "exceeds the 227 KB shared memory of an SM");
};
constexpr uint32_t C2_ACC0 = 0, C2_ACC1 = 64, C2_DA = 128, C2_DB = 192;  // TMEM columns of a pipeline

// Warp-cooperative state prefetch. The 32 rows of a warp need 2B + 2 planes x 32 bytes; with
// E % 4 == 0 four consecutive rows are four consecutive bytes of a plane, so the warp's state is
// (2B + 2) x 8 aligned 32-bit words: lane l loads words l, l + 32, ... (5 loads and 5 live registers
// per thread instead of 18 byte loads / 18 registers), and a row's bytes are gathered with warp
// shuffles when they are needed (word q * 8 + lane / 4 sits in slot q / 4 of lane (q % 4) * 8 +
// lane / 4). live = false: rec_state planes of each row's own step; live = true: the environments'
// live planes, meaningful for the rows of the rollout's last step only.
template <int B>
struct warp_state {
  uint32_t w[((2 * B + 2) * 8 + 31) / 32];
};
template <int B>
__device__ __forceinline__ void load_warp_state(const int8_t *__restrict__ base, bool live, const learner_rows &L,
                                                int tile, int warp_row0, int lane, warp_state<B> &x) {
  constexpr int P = 2 * B + 2, NW = (P * 8 + 31) / 32;
#pragma unroll
  for (int j = 0; j < NW; ++j) {
    const int w = lane + 32 * j, q = w >> 3, r0 = warp_row0 + 4 * (w & 7);
    const int tt = r0 / L.E, i0 = tile * L.E + r0 % L.E;
    x.w[j] = 0;
    if (q < P && i0 < L.stride && (live ? tt == L.T - 1 : tt < L.T))
      x.w[j] = *reinterpret_cast<const uint32_t *>(base + ((size_t)(live ? 0 : tt) * P + q) * L.stride + i0);
  }
}
template <int B>
__device__ __forceinline__ void gather_row_state(const warp_state<B> &x, int lane, bool valid, row_state<B> &out) {
  constexpr int P = 2 * B + 2;
#pragma unroll
  for (int q = 0; q < P; ++q) {
    const uint32_t word = __shfl_sync(0xffffffffu, x.w[q >> 2], ((q & 3) << 3) + (lane >> 2));
    out.v[q] = valid ? (int)(int8_t)(word >> (8 * (lane & 3))) : 0;
  }
}

// Byte-wise path (first tile of a pipeline, step counts with E % 4 != 0). END state of a row:
// overflowed terminal state when done (bin[a] -= item, item kept: bin_packing.h:54-61), the live
// state at the rollout's last step, zeros (row unused) otherwise. A row needs either its start state
// (done) or the live state, never both -- `done` is known when the loads are issued.
template <int B>
__device__ __forceinline__ void load_end_source(const learner_rows &L, int tile, int row, int done, row_state<B> &x) {
  constexpr int P = 2 * B + 2;
  const int tt = row / L.E, e = row % L.E, i = tile * L.E + e;
  const bool ok = tt < L.T && i < L.n, from_start = ok && done, from_live = ok && !done && tt == L.T - 1;
  const int8_t *src = from_start ? L.rec_state + (size_t)tt * P * L.stride + i : L.live_state + i;
#pragma unroll
  for (int q = 0; q < P; ++q) {
    x.v[q] = 0;
    if (from_start || from_live)
      x.v[q] = src[(size_t)q * L.stride];
  }
}
template <int B>
__device__ __forceinline__ void fix_end_state(row_state<B> &x, int done, int act) {
  if (done) {
#pragma unroll
    for (int b = 0; b < B; ++b)
      if (b == act) {
        x.v[2 * b] -= x.v[2 * B];
        x.v[2 * b + 1] -= x.v[2 * B + 1];
      }
  }
}
// Layer-2 accumulator -> relu(acc + b2) (registers only) -> value head in fp32.
template <int D2, bool KEEP>
__device__ __forceinline__ float epi2_value(uint32_t acc, const tid_t &t, const float *__restrict__ b2,
                                            const float *__restrict__ w3, float b3, float *keep, int h0 = 0,
                                            int h1 = D2 / (D2 < 32 ? D2 : 32)) {
  constexpr int CH = D2 < 32 ? D2 : 32;
  float s = 0.f;
#pragma unroll
  for (int h = h0; h < h1; ++h) {
    float v[CH];
    tmem_load<CH>(acc + t.lane_base + h * CH, v);
#pragma unroll
    for (int j4 = 0; j4 < CH; j4 += 4) {
      const float4 b = *reinterpret_cast<const float4 *>(b2 + h * CH + j4);
      const float4 w = *reinterpret_cast<const float4 *>(w3 + h * CH + j4);
      const float y0 = fmaxf(v[j4] + b.x, 0.f), y1 = fmaxf(v[j4 + 1] + b.y, 0.f);
      const float y2 = fmaxf(v[j4 + 2] + b.z, 0.f), y3 = fmaxf(v[j4 + 3] + b.w, 0.f);
      s = fmaf(y0, w.x, s), s = fmaf(y1, w.y, s), s = fmaf(y2, w.z, s), s = fmaf(y3, w.w, s);
      if (KEEP)
        keep[h * CH + j4] = y0, keep[h * CH + j4 + 1] = y1, keep[h * CH + j4 + 2] = y2, keep[h * CH + j4 + 3] = y3;
    }
  }
  return s + b3;
}

template <int D0, int D1, int D2, int MODE, bool EG, bool PROBE = false>
__global__ void __launch_bounds__((cmap<D1, D2, MODE>::THREADS), 1) fused_critic_kernel(critic_args a) {
  using CM = cmap<D1, D2, MODE>;
  constexpr int NP = CM::NP, NH = CM::NH;
  constexpr int RT = 32 + 128 * NH;  // threads of an operands-ready hand-over
  constexpr int NB = 8;  // bins (the fused path covers the 8-bin problem)
  static_assert(D0 == 4 * NB, "observation width");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);  // stays a shared-space pointer
  float *fl = reinterpret_cast<float *>(smem + CM::FLOATS);
  // mbarriers: [wg] MMA completion on the chain, [NP + wg] dW2 GEMM done (H1 free), [2 NP + wg] dW1
  // GEMM done (XS, dH2 free), [3 NP + wg] layer 1 of the start rows (issued right behind layer 2 of
  // the end rows: a parity wait cannot tell two outstanding completions of one mbarrier apart),
  // [4 NP] dH1 slot free
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + CM::BARS);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + CM::BARS + 248);
  const net3 net = a.net;
  const learner_rows &L = a.rows;
  const tid_t t = thread_id();
  const bool issuer = t.warp >= 4 * NH * NP;                          // warp-uniform
  const int wg = issuer ? t.warp - 4 * NH * NP : t.warp / (4 * NH);  // pipeline index
  const int half = NH == 2 ? (t.warp >> 2) & 1 : 0;                   // which 32-column chunk (NH = 2)
  const uint32_t sbase = umma::smem_u32(smem);
  // every CTA: wall clock at entry / end of the tile loop / end of the kernel, and its SM (load balance)
  long long *gclk = (PROBE && MODE == CRITIC_STEP && a.clk && threadIdx.x == 96 && blockIdx.x < 160) ? a.clk + 112 + 4 * blockIdx.x : nullptr;
  if (gclk)
    gclk[0] = global_ns(), gclk[3] = sm_id();

  umma::pdl_launch_dependents();
  if (t.warp == 0)
    umma::tmem_alloc(tmem_slot, 512);
  if (threadIdx.x == 0) {
    for (int q = 0; q < 4 * NP + 1; ++q)
      umma::mbar_init(bars + q, 1);
    umma::fence_mbar_init();
  }
  constexpr bool small_zero = MODE == CRITIC_STEP && D1 == 64 && D2 == 64;  // see fused_policy_step_kernel
  {
    const float *P = stage_params_issue(a.params, net.n_params, smem + CM::WG0 + (MODE == CRITIC_STEP ? PANEL : 0), bars + 29);
    if (small_zero) {  // (the XS panels are not part of the scratch: zeroed behind the parameter copy)
      zero_bytes(smem + CM::WG0 + CM::XS, PANEL);
      zero_bytes(smem + CM::WG0 + CM::WG_BYTES + CM::XS, PANEL);
    }
    __syncthreads();
    stage_params_wait(a.params, net.n_params, bars + 29);
    stage_w1_packed<D1>(P + net.o_w1, smem + CM::W1P);
    stage_weight_f16(P + net.o_w2, D2, D1, D2, 1.f, smem + CM::W2_HI, smem + CM::W2_LO);
    for (int i = threadIdx.x; i < D1; i += blockDim.x) fl[CM::F_B1 + i] = P[net.o_b1 + i];
    for (int i = threadIdx.x; i < D2; i += blockDim.x) fl[CM::F_B2 + i] = P[net.o_b2 + i];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) fl[CM::F_W3 + i] = i < D2 ? P[net.o_w3 + i] : 0.f;
    if (threadIdx.x == 0)
      fl[CM::F_B3] = P[net.o_b3];
    __syncthreads();  // the scratch (activation panels) is free again
  }
  if (!small_zero) {
    zero_bytes(smem + CM::DH1_HI, CM::BARS - CM::DH1_HI);
    __syncthreads();
  }
  if (!issuer && MODE == CRITIC_STEP)  // ones column (col D0) of both XS panels: bias gradients for free
    *reinterpret_cast<uint16_t *>(smem + CM::WG0 + wg * CM::WG_BYTES + CM::XS + umma::panel_off(t.row, D0)) = 0x3F80;
  sync_after_smem_writes();
  const uint32_t tmem = *tmem_slot;

  const int nt = (int)blockIdx.x < a.n_tiles ? (a.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  uint8_t *wsm = smem + CM::WG0 + wg * CM::WG_BYTES;
  const uint32_t wbase = sbase + CM::WG0 + wg * CM::WG_BYTES;
  const uint32_t tm = tmem + CM::TCOLS_PER * wg;
  const uint32_t dh1_lbo = CM::WG0 + wg * CM::WG_BYTES + CM::G2_HI - CM::DH1_HI;  // dH1 panel -> own dH2 panel
  uint64_t *bar = bars + wg, *bar_dw2 = bars + NP + wg, *bar_dw1 = bars + 2 * NP + wg, *bar_l1s = bars + 3 * NP + wg,
           *bar_slot = bars + 4 * NP;
  uint32_t rp = 0;
  // eg: V(end) comes from global memory (fused_vend_kernel ran on the compacted end rows): no end pass.
  // A tile is then  layer 1 (start rows, staged in H1_LO ahead of the previous tile's dW1 GEMM, as in
  // fused_policy_step_kernel) -> layer 2 -> values -> ...
  constexpr bool eg = EG;  // (a.v_end is set)
  // dW3 partial sums. One thread per row: this thread's row-sums of all D2 columns. Two threads per
  // row: ONE register -- after every tile the 32 lanes of a warp transpose-reduce their 32 columns
  // (butterfly, fixed order), lane l keeps column 32 half + l summed over the warp's rows (32
  // accumulators per thread would not fit in the 96 registers of the 576-thread kernel, and spills
  // go to L2 here: the shared-memory carve-out leaves no L1)
  constexpr int NDW3 = (MODE == CRITIC_STEP && NH == 1) ? D2 : 1;
  float dw3[NDW3];
#pragma unroll
  for (int q = 0; q < NDW3; ++q)
    dw3[q] = 0.f;
  float db3 = 0.f;

  if (issuer) {
    // ================= MMA issuer of pipeline wg
    auto layer1 = [&](uint32_t x0, uint64_t *done_bar) {
      issue_gemm<D0 / 16, false, false, false, true>(tm + C2_ACC0, x0, 0, sbase + CM::W1P, sbase + CM::W1P + 64,
                                                     ID<D1>::FK_FK, false);
      umma::commit(done_bar);
    };
    auto layer2 = [&]() {  // A = H1 from tensor memory (the epilogue's copy in ACC0)
      issue_gemm_ta<D1, D1 / 16, false>(tm + C2_ACC1, tm + C2_ACC0, sbase + CM::W2_HI, sbase + CM::W2_LO, ID<D2>::FK_FK);
      umma::commit(bar);
    };
    bool first = true;
    if (MODE == CRITIC_GAE && eg && a.v_end_out) {  // compacted end-row passes of this pipeline's tiles (see the epilogue side)
      volatile int *npass_slot = reinterpret_cast<int *>(smem + CM::SCR) + wg * 2 * NH * TILE + 136;
      int u = 0;
      for (int j = wg; j < nt; j += NP, u ^= 1) {
        ready_sync(wg, rp, RT);  // the pass count after this tile's end rows joined the list
        const int npass = npass_slot[u];
        for (int ps = 0; ps < npass; ++ps) {
          ready_sync(wg, rp, RT);  // observations
          if (umma::elect_one())
            layer1(wbase + CM::H1_LO, bar);
          __syncwarp();
          ready_sync(wg, rp, RT);  // H1 (tensor memory)
          if (umma::elect_one())
            layer2();
          __syncwarp();
        }
      }
    }
    if (wg < nt) {
      ready_sync(wg, rp, RT);  // observations of the first tile (end rows; eg: start rows) staged in the H1_LO panel
      if (umma::elect_one())
        layer1(wbase + CM::H1_LO, bar);
      __syncwarp();
    }
    for (int j = wg; j < nt; j += NP) {
      if (!eg) {
        ready_sync(wg, rp, RT);  // H1 (end rows)
        if (umma::elect_one())
          layer2();
        __syncwarp();
        ready_sync(wg, rp, RT);  // start-row observations in the XS panel
        if (umma::elect_one())
          layer1(wbase + CM::XS, bar_l1s);
        __syncwarp();
      }
      ready_sync(wg, rp, RT);  // H1 (start rows)
      if (umma::elect_one())
        layer2();
      __syncwarp();
      if (MODE == CRITIC_STEP) {
        ready_sync(wg, rp, RT);  // dH2
        // dH1 = dH2 . W2; dW2 += dH2^T . H1 (M = 64) runs behind the dH1 epilogue
        if (umma::elect_one()) {
          // A = dH2 from tensor memory (the targets phase's copy in ACC1)
          issue_gemm_ta<D2, D2 / 16, true>(tm + C2_ACC0, tm + C2_ACC1, sbase + CM::W2_HI, sbase + CM::W2_LO, ID<D1>::BK_FM);
          umma::commit(bar);
          issue_gemm<8, true, true, true, true>(tm + C2_DA, wbase + CM::G2_HI, wbase + CM::G2_LO, wbase + CM::H1_HI,
                                                wbase + CM::H1_LO, ID<D1>::BM_FM_64, !first);
          umma::commit(bar_dw2);
        }
        __syncwarp();
      }
      ready_sync(wg, rp, RT);  // (dH1 in the shared slot and) the next tile's observations (end rows; eg: start rows)
      if (umma::elect_one()) {
        if (j + NP < nt)  // ahead of dW1: see fused_policy_step_kernel
          layer1(wbase + CM::H1_LO, bar);
        if (MODE == CRITIC_STEP) {
          //   DB[128 x D0+16] += [dH1|dH2]^T . [X0|1]   rows 0.. = [dW1 | db1], rows 64.. col D0 = db2
          issue_gemm_mn_lbo<8>(tm + C2_DB, sbase + CM::DH1_HI, sbase + CM::DH1_LO, dh1_lbo, wbase + CM::XS,
                               ID<D0 + 16>::BM_FM, !first);
          umma::commit(bar_dw1);
          umma::commit(bar_slot);
        }
      }
      __syncwarp();
      first = false;
    }
  } else {
    // ================= epilogue threads of pipeline wg: thread = one row of the tile
    const float *b1 = fl + CM::F_B1, *b2 = fl + CM::F_B2, *w3 = fl + CM::F_W3;
    const float b3 = fl[CM::F_B3];
    // partial values of the tile's rows: ve[half][row], vs[half][row]; V = (sum of the partials) + b3
    float *ve = reinterpret_cast<float *>(smem + CM::SCR) + wg * 2 * NH * TILE, *vs = ve + NH * TILE;
    auto value_of = [&](const float *vp, int r) { return NH == 2 ? (vp[r] + vp[TILE + r]) + b3 : vp[r] + b3; };
    auto end_value = [&](int r) { return eg ? ve[r] : value_of(ve, r); };
    const int h0 = NH == 2 ? half : 0, h1d1 = NH == 2 ? half + 1 : D1 / (D1 < 32 ? D1 : 32),
              h1d2 = NH == 2 ? half + 1 : D2 / (D2 < 32 ? D2 : 32);
    // two threads per row split the state work: chunk 0's thread prefetches the next START state and
    // encodes the start-row observations (and keeps the target / db3 books), chunk 1's thread
    // prefetches the next END state and stages the end-row observations (18 raw bytes in registers
    // each; with one thread per row the same thread does both)
    const bool stager = half == 0, xe_role = half == NH - 1;
    constexpr int XT = 128 * NH;    // threads of the value-exchange barrier
    uint32_t phase = 0, phase_dw2 = 0, phase_dw1 = 0, phase_l1s = 0;
    auto wait_mma = [&]() {
      umma::mbar_wait(bar, phase);
      phase ^= 1;
      umma::fence_after_sync();
    };
    const int tt = t.row / L.E, e = t.row % L.E;
    const bool last = tt == L.T - 1;
    bool first = true;
    long long *clk = (PROBE && MODE == CRITIC_STEP && a.clk && blockIdx.x == 0 && threadIdx.x == 96) ? a.clk : nullptr;
    int clk_n = 0;
#define CSTAMP() do { if (PROBE && clk && clk_n < 112) clk[clk_n++] = clock64(); } while (0)
    // prefetched state of the rows of this warp (warp_state: 5 words per thread); `fast` = the
    // word-wise path applies (E % 4 == 0), otherwise byte loads at the point of use
    const bool fast = L.E % 4 == 0;
    const int warp_row0 = t.row - t.lane;
    if (MODE == CRITIC_GAE && eg && a.v_end_out) {
      // ---- V(end state) of the rows of this pipeline's tiles that END a trajectory (done, or the rollout's last
      //      step: about a third of the rows at T = 4), compacted into passes of 128 rows: the rows of tile after
      //      tile join a list (ballot + prefix sums, deterministic order), a pass runs whenever 128 are waiting
      //      and once more after the last tile. Replaces the second fused_vend_kernel launch of an iteration:
      //      its fixed cost (set-up, launch, tail imbalance) was larger than its work.
      //      list: uint16 (done << 15 | tile ordinal << 7 | row) x 256, then 4 warp totals + 2 pass-count slots, in the
      //      pipeline's ve / vs scratch (unused until the tile loop)
      uint16_t *lst = reinterpret_cast<uint16_t *>(ve);
      int *wtot = reinterpret_cast<int *>(ve) + 128;
      volatile int *npass_slot = wtot + 8;
      int cnt = 0, u = 0;
      for (int j = wg; j < nt; j += NP, u ^= 1) {
        const int tile = blockIdx.x + j * gridDim.x, i = tile * L.E + e;
        const bool valid = tt < L.T && i < L.n;
        const int dn = valid ? (int)(L.rec_done[(size_t)tt * L.n + i] != 0) : 0;
        const bool ends = valid && (last || dn);
        const unsigned bal = __ballot_sync(0xffffffffu, ends);
        if (t.lane == 0)
          wtot[t.w] = __popc(bal);
        asm volatile("bar.sync %0, %1;\n" ::"r"(9 + wg), "r"(128) : "memory");
        int off = cnt + __popc(bal & ((1u << t.lane) - 1u)), total = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q < t.w)
            off += wtot[q];
          total += wtot[q];
        }
        if (ends)  // done flag | tile ordinal | row: the pass below needs no dependent load
          lst[off] = (uint16_t)((dn << 15) | (((j - wg) / NP) << 7) | t.row);
        cnt += total;
        asm volatile("bar.sync %0, %1;\n" ::"r"(9 + wg), "r"(128) : "memory");  // list complete; wtot may be reused
        const bool last_tile = j + NP >= nt;
        const int npass = cnt / TILE + ((last_tile && cnt % TILE) ? 1 : 0);
        if (t.row == 0)
          npass_slot[u] = npass;
        ready_sync(wg, rp, RT);  // blocking on purpose: the issuer reads the count, these threads must not run ahead of it
        for (int ps = 0; ps < npass; ++ps) {
          const int m = cnt < TILE ? cnt : TILE;
          const bool has = t.row < m;
          const int entry = has ? (int)lst[t.row] : 0;
          const int tile2 = blockIdx.x + (wg + ((entry >> 7) & 255) * NP) * gridDim.x, r2 = entry & 127;
          const int tt2 = r2 / L.E, i2 = tile2 * L.E + r2 % L.E;
          const size_t k2 = (size_t)tt2 * L.n + i2;
          const int done2 = entry >> 15;
          int act2 = 0;
          if (done2)
            act2 = L.rec_action[k2];
          row_state<NB> x;
          {
            constexpr int P = 2 * NB + 2;
            const int8_t *src = done2 ? L.rec_state + (size_t)tt2 * P * L.stride + i2 : L.live_state + i2;
#pragma unroll
            for (int q = 0; q < P; ++q) {
              x.v[q] = 0;
              if (has)
                x.v[q] = src[(size_t)q * L.stride];
            }
          }
          fix_end_state<NB>(x, done2, act2);
          // the rest of the list moves down (every thread has read its entry of this pass)
          const int rest = cnt - m;
          const int moved = t.row < rest ? (int)lst[TILE + t.row] : 0;
          asm volatile("bar.sync %0, %1;\n" ::"r"(9 + wg), "r"(128) : "memory");
          if (t.row < rest)
            lst[t.row] = (uint16_t)moved;
          cnt = rest;
          encode_row<NB>(wsm + CM::H1_LO, t.row, x, L.inv_w, L.inv_h);
          ready_arrive(wg, rp, RT);
          wait_mma();  // layer 1
          epi2_fwd<D1, true, false>(tm + C2_ACC0, t, b1, nullptr, nullptr, h0, h1d1);  // H1 only as a TMEM A operand
          ready_arrive(wg, rp, RT);
          wait_mma();  // layer 2
          const float v = epi2_value<D2, false>(tm + C2_ACC1, t, b2, w3, b3, nullptr, h0, h1d2);
          if (has)
            a.v_end_out[k2] = v;
        }
      }
      __threadfence_block();
      asm volatile("bar.sync %0, %1;\n" ::"r"(9 + wg), "r"(128) : "memory");  // the pipeline's V(end) are visible to its threads
    }
    warp_state<NB> ws;  // start state of this tile (chunk 0's threads)
    int done = 0;
    if (wg < nt) {
      const int tile0 = blockIdx.x + wg * gridDim.x, i0 = tile0 * L.E + e;
      int act = 0;
      if (tt < L.T && i0 < L.n) {
        done = L.rec_done[(size_t)tt * L.n + i0];
        act = L.rec_action[(size_t)tt * L.n + i0];
      }
      if (xe_role && !eg) {
        row_state<NB> xe;
        load_end_source<NB>(L, tile0, t.row, done, xe);
        fix_end_state<NB>(xe, done, act);
        encode_row<NB>(wsm + CM::H1_LO, t.row, xe, L.inv_w, L.inv_h);
      }
      if (stager) {
        if (fast)
          load_warp_state<NB>(L.rec_state, false, L, tile0, warp_row0, t.lane, ws);
        if (eg) {  // the first tile's start-row observations for its layer-1 GEMM
          row_state<NB> xs;
          if (fast)
            gather_row_state<NB>(ws, t.lane, tt < L.T && i0 < L.n, xs);
          else
            load_row_state<NB>(L, tile0, t.row, xs);
          encode_row<NB>(wsm + CM::H1_LO, t.row, xs, L.inv_w, L.inv_h);
        }
      }
      ready_arrive(wg, rp, RT);
    }
    for (int j = wg; j < nt; j += NP) {
      const int tile = blockIdx.x + j * gridDim.x;
      const int i = tile * L.E + e;
      const bool valid = tt < L.T && i < L.n;
      const size_t k = (size_t)tt * L.n + i;
      const bool has_next = j + NP < nt;
      // GAE: done flags of the env this thread walks (threads < E)
      uint32_t dmask = 0;
      if (MODE == CRITIC_GAE && t.row < L.E && i < L.n)
        for (int q = 0; q < L.T && q < 32; ++q)
          dmask |= (uint32_t)(L.rec_done[(size_t)q * L.n + tile * L.E + t.row] != 0) << q;
      // V(end) of this row from the compacted pre-pass (eg): needed where the trajectory ends here
      float vend_g = 0.f;
      if (eg && valid && (done || last))
        vend_g = a.v_end[k];
      if (!eg) {
        // ---- pass 1: V of the end rows
        CSTAMP();
        wait_mma();  // layer 1 (end rows): H1 only as a TMEM A operand
        CSTAMP();
        epi2_fwd<D1, true, false>(tm + C2_ACC0, t, b1, nullptr, nullptr, h0, h1d1);
        ready_arrive(wg, rp, RT);
        CSTAMP();
        // start-row observations -> XS (the previous tile's dW1 GEMM ran behind the epilogue above)
        if (MODE == CRITIC_STEP && !first) {
          umma::mbar_wait(bar_dw1, phase_dw1);
          phase_dw1 ^= 1;
        }
        if (stager) {
          row_state<NB> xs;
          if (fast)
            gather_row_state<NB>(ws, t.lane, valid, xs);
          else
            load_row_state<NB>(L, tile, t.row, xs);  // step counts with E % 4 != 0: loaded where it is used
          encode_row<NB>(wsm + CM::XS, t.row, xs, L.inv_w, L.inv_h);
        }
        ready_arrive(wg, rp, RT);
        CSTAMP();
      }
      // the next tile's state: the start-state loads fly behind the layer-2 GEMM of the end rows and
      // are packed right after it; the end-state loads (source chosen by `done`, known by then) fly
      // behind the start rows' layer 1 / 2
      warp_state<NB> wn, wes, wel;  // next tile: start words (chunk 0) / start + live words (chunk 1)
      int ndone = 0, nact = 0;
      const int ntile = tile + NP * gridDim.x;
      if (has_next) {
        const int ni = ntile * L.E + e;
        if (fast) {
          if (stager)
            load_warp_state<NB>(L.rec_state, false, L, ntile, warp_row0, t.lane, wn);
          if (xe_role && !eg) {
            if (NH == 2)  // (one thread per row: the start words are in wn already)
              load_warp_state<NB>(L.rec_state, false, L, ntile, warp_row0, t.lane, wes);
            load_warp_state<NB>(L.live_state, true, L, ntile, warp_row0, t.lane, wel);
          }
        }
        if (tt < L.T && ni < L.n) {
          ndone = L.rec_done[(size_t)tt * L.n + ni];
          nact = L.rec_action[(size_t)tt * L.n + ni];
        }
      }
      if (!eg) {
        wait_mma();  // layer 2 (end rows)
        CSTAMP();
        ve[half * TILE + t.row] = epi2_value<D2, false>(tm + C2_ACC1, t, b2, w3, 0.f, nullptr, h0, h1d2);
        // ---- pass 2: start rows, H1 kept for the dW2 GEMM
        CSTAMP();
        umma::mbar_wait(bar_l1s, phase_l1s);  // layer 1 (start rows)
        phase_l1s ^= 1;
        umma::fence_after_sync();
      } else {
        if (half == 0)
          ve[t.row] = vend_g;  // the complete value (end_value() below)
        wait_mma();  // layer 1 (start rows; staged in H1_LO, which the epilogue below overwrites)
      }
      CSTAMP();
      epi2_fwd<D1, true, MODE == CRITIC_STEP>(tm + C2_ACC0, t, b1, wsm + CM::H1_HI, wsm + CM::H1_LO, h0, h1d1);  // panels: dW2
      ready_arrive(wg, rp, RT);
      CSTAMP();
      if (eg && MODE == CRITIC_STEP) {
        // start-row observations -> XS for this tile's dW1 GEMM (the previous tile's dW1 GEMM, which
        // read XS, ran behind the epilogue above)
        if (!first) {
          umma::mbar_wait(bar_dw1, phase_dw1);
          phase_dw1 ^= 1;
        }
        if (stager) {
          row_state<NB> xs;
          if (fast)
            gather_row_state<NB>(ws, t.lane, valid, xs);
          else
            load_row_state<NB>(L, tile, t.row, xs);
          encode_row<NB>(wsm + CM::XS, t.row, xs, L.inv_w, L.inv_h);
        }
      }

      CSTAMP();
      wait_mma();  // layer 2 (start rows)
      CSTAMP();
      vs[half * TILE + t.row] = epi2_value<D2, false>(tm + C2_ACC1, t, b2, w3, 0.f, nullptr, h0, h1d2);
      asm volatile("bar.sync %0, %1;\n" ::"r"(9 + wg), "r"(XT) : "memory");  // ve / vs of the tile visible
      CSTAMP();
      if (MODE == CRITIC_GAE) {
        // thread e < E walks its env backwards (same recurrence as device_fns.cuh gae_env)
        if (t.row < L.E && i < L.n) {
          float a_next = 0.f;
          for (int q = L.T - 1; q >= 0; --q) {
            const size_t kq = (size_t)q * L.n + i;
            const int r = q * L.E + t.row;
            const int d = L.T <= 32 ? (int)((dmask >> q) & 1u) : (int)L.rec_done[kq];
            const bool ends = d || q == L.T - 1;
            const float vn = ends ? end_value(r) : value_of(vs, r + L.E);
            const float vn_adv = d ? 0.f : vn;
            const float delta = (d ? 0.f : 1.f) + a.gamma * vn_adv - value_of(vs, r);
            const float adv = delta + (ends ? 0.f : a.lambda * a.gamma * a_next);
            a.adv_out[kq] = adv;
            a_next = adv;
            const float aa = fabsf(adv);
            if (aa < 3.0e38f)           // (false for NaN too: same filter as conv_table_absmax_kernel)
              db3 = fmaxf(db3, aa);     // db3 is free in this mode: running max |A| of this thread
          }
        }
        asm volatile("bar.sync %0, %1;\n" ::"r"(9 + wg), "r"(XT) : "memory");  // ve / vs may be overwritten
      } else {
        // ---- targets and dY = V - target (square_loss_grad, nn.h:548-550)
        // (with two threads per row both compute the same dY; thread `stager` keeps the books)
        float dy = 0.f;
        if (valid) {
          const bool ends = done || last;
          const float vn = ends ? end_value(t.row) : value_of(vs, t.row + L.E);
          const float tgt = (done ? 0.f : 1.f) + a.gamma * vn;  // not masked at terminals (quirk 6)
          dy = value_of(vs, t.row) - tgt;
          if (a.targets_out && stager)
            a.targets_out[k] = tgt;
        }
        if (stager)
          db3 += dy;
        // the next tile overwrites ve / vs only after hand-overs that every thread takes part in
        // dH2 = dY w3 . relu'(H2) (rank 1: no GEMM) -> panels; dW3 += dY H2. H2 = relu(acc + b2) is
        // recomputed from the layer-2 accumulator, which stays in TMEM until the next tile's layer 2
        // (64 registers less than keeping the row across the value exchange)
        {
          constexpr int CH = D2 < 32 ? D2 : 32;
#pragma unroll
          for (int h = h0; h < h1d2; ++h) {
            float y[CH];
            float contrib[NH == 2 ? CH : 1];  // dY . H2 of this row (NH = 2: reduced over the warp below)
            tmem_load<CH>(tm + C2_ACC1 + t.lane_base + h * CH, y);
#pragma unroll
            for (int cc = 0; cc < CH / 8; ++cc) {
              float g[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const int c = h * CH + 8 * cc + q;  // column; dw3 is indexed by the thread's own columns
                const float yy = fmaxf(y[8 * cc + q] + b2[c], 0.f);
                g[q] = yy > 0.f ? dy * w3[c] : 0.f;
                if (NH == 2)
                  contrib[NH == 2 ? 8 * cc + q : 0] = dy * yy;
                else
                  dw3[NH == 2 ? 0 : c] = fmaf(dy, yy, dw3[NH == 2 ? 0 : c]);
              }
              uint4 hh, ll;
              split8<false>(g, hh, ll);
              const uint32_t off = umma::panel_chunk_off(t.row, h * (CH / 8) + cc);
              *reinterpret_cast<uint4 *>(wsm + CM::G2_HI + off) = hh;
              *reinterpret_cast<uint4 *>(wsm + CM::G2_LO + off) = ll;
              tmem_put_chunk<CH>(tm + C2_ACC1 + t.lane_base + h * CH, cc, hh, ll);  // A operand of the dH1 GEMM
            }
            if (NH == 2) {  // transpose-reduce: lane l ends with column l of the chunk, summed over 32 rows
#pragma unroll
              for (int sft = CH / 2; sft >= 1; sft >>= 1) {
                const bool up = (t.lane & sft) != 0;
#pragma unroll
                for (int q = 0; q < sft; ++q) {
                  const float give = up ? contrib[q] : contrib[q + sft];
                  const float keep = up ? contrib[q + sft] : contrib[q];
                  contrib[q] = keep + __shfl_xor_sync(0xffffffffu, give, sft);
                }
              }
              dw3[0] += contrib[0];
            }
          }
          umma::tmem_st_wait();
        }
        ready_arrive(wg, rp, RT);
        CSTAMP();
        wait_mma();  // dH1
        if (j > 0)   // the shared dH1 slot (see fused_policy_step_kernel)
          umma::mbar_wait(bar_slot, (uint32_t)(j - 1) & 1u);
        CSTAMP();
        epi2_bwd<D1>(tm + C2_ACC0, t, wsm + CM::H1_HI, smem + CM::DH1_HI, smem + CM::DH1_LO, h0, h1d1);
        CSTAMP();
        umma::mbar_wait(bar_dw2, phase_dw2);  // H1 is free (the dW2 GEMM ran behind the dH1 epilogue)
        phase_dw2 ^= 1;
      }
      if (has_next && eg) {
        if (stager) {  // the next tile's start-row observations for its layer-1 GEMM
          row_state<NB> xs;
          const int ni = ntile * L.E + e;
          if (fast)
            gather_row_state<NB>(wn, t.lane, tt < L.T && ni < L.n, xs);
          else
            load_row_state<NB>(L, ntile, t.row, xs);
          encode_row<NB>(wsm + CM::H1_LO, t.row, xs, L.inv_w, L.inv_h);
          ws = wn;
        }
        done = ndone;
      } else if (has_next) {
        if (xe_role) {
          row_state<NB> xe;
          if (fast) {  // end state of the next tile's row: its start state (done) or the live state (last step)
            const int ni = ntile * L.E + e;
            const bool nvalid = tt < L.T && ni < L.n;
            row_state<NB> xl;
            gather_row_state<NB>(NH == 2 ? wes : wn, t.lane, nvalid && ndone, xe);
            gather_row_state<NB>(wel, t.lane, nvalid && !ndone && last, xl);
#pragma unroll
            for (int q = 0; q < 2 * NB + 2; ++q)
              xe.v[q] += xl.v[q];  // at most one of the two is non-zero
            fix_end_state<NB>(xe, ndone, nact);
          } else {
            load_end_source<NB>(L, ntile, t.row, ndone, xe);
            fix_end_state<NB>(xe, ndone, nact);
          }
          encode_row<NB>(wsm + CM::H1_LO, t.row, xe, L.inv_w, L.inv_h);
        }
        if (stager)
          ws = wn;
        done = ndone;
      }
      ready_arrive(wg, rp, RT);
      CSTAMP();
      first = false;
    }
#undef CSTAMP
    if (MODE == CRITIC_STEP && !first) {  // the last tile's dW1 GEMM
      umma::mbar_wait(bar_dw1, phase_dw1);
      umma::fence_after_sync();
    }
  }

  if (gclk)
    gclk[1] = global_ns();
  if (MODE == CRITIC_GAE && a.adv_maxbits && !issuer) {  // one atomic per epilogue warp
    const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(db3));
    if ((threadIdx.x & 31) == 0 && m)
      atomicMax(a.adv_maxbits, m);
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  if (MODE == CRITIC_STEP) {
    // ---- drain: partial gradient of this CTA (pipeline 0's sums + pipeline 1's) -> global, staged in
    //      shared memory (pipeline 1's dead panels) so that it leaves as coalesced stores (see the policy step)
    float *part = a.partials + (size_t)blockIdx.x * partial_stride(net.n_params);
    float *sg = reinterpret_cast<float *>(smem + CM::WG0 + CM::WG_BYTES);
    static_assert(MODE != CRITIC_STEP || CM::WG_BYTES >= 8448 * 4, "staging area of the partial gradient");
    static_assert(MODE != CRITIC_STEP || 257 * (D2 + 1) * 4 + 4 * (D2 + 1) * 4 <= CM::WG_BYTES, "dW3 scratch vs staging area");
    auto put = [&](int i, float v) { sg[i + (i >> 5)] = v; };
    const bool two = nt > 1;
    if (nt == 0) {
      for (int q = threadIdx.x; q < net.n_params; q += blockDim.x)
        part[q] = 0.f;
    } else {
      if (net.shared)  // the other heads' slots of the flat gradient
        for (int q = threadIdx.x; q < net.n_params; q += blockDim.x)
          if (!net_owns(net, q))
            put(q, 0.f);
      const bool drainer = threadIdx.x < 256;  // 256 threads read the TMEM accumulators (t.wg = 0, 1)
      if (drainer) {  // dW2[n][k]: DA (M = 64) row n = TMEM lane 32 (n / 16) + n % 16, col k
        constexpr int DC = D1 / 2;
        float v[DC], w[DC];
        tmem_load<DC>(tmem + C2_DA + t.lane_base + t.wg * DC, v);
        if (two) {
          tmem_load<DC>(tmem + 256 + C2_DA + t.lane_base + t.wg * DC, w);
#pragma unroll
          for (int q = 0; q < DC; ++q)
            v[q] += w[q];
        }
        int nrow = t.lane < 16 ? t.w * 16 + t.lane : -1;
        if (nrow >= 0 && nrow < D2)
#pragma unroll
          for (int q = 0; q < DC; ++q)
            put(net.o_w2 + nrow * D1 + t.wg * DC + q, v[q]);
      }
      if (drainer) {  // dW1[n][k] + db1[n]: DB row n, cols 0..D0-1 and D0; db2[n]: DB row 64 + n, col D0
        constexpr int DC = (D0 + 16) / 2;
        float v[DC], w[DC];
        tmem_load<DC>(tmem + C2_DB + t.lane_base + t.wg * DC, v);
        if (two) {
          tmem_load<DC>(tmem + 256 + C2_DB + t.lane_base + t.wg * DC, w);
#pragma unroll
          for (int q = 0; q < DC; ++q)
            v[q] += w[q];
        }
#pragma unroll
        for (int q = 0; q < DC; ++q) {
          int col = t.wg * DC + q;
          if (t.row < D1) {
            if (col < D0)
              put(net.o_w1 + t.row * D0 + col, v[q]);
            else if (col == D0)
              put(net.o_b1 + t.row, v[q]);
          } else if (t.row >= 64 && t.row - 64 < D2 && col == D0) {
            put(net.o_b2 + t.row - 64, v[q]);
          }
        }
      }
      // dW3 / db3: thread-local sums -> fixed-order column sums through shared memory (the panels
      // are free): red[256 rows][D2 + 1], 4 row quarters per column, combined in order
      float *red = reinterpret_cast<float *>(smem + CM::WG0);
      constexpr int W = D2 + 1;
      for (int q = threadIdx.x; q < 256 * W; q += blockDim.x)
        red[q] = 0.f;
      __syncthreads();
      if (!issuer) {
        const int rr = wg * TILE + t.row;  // row = pipeline * 128 + tile row
        if (NH == 2) {  // lane l of a warp holds column 32 half + l, summed over the warp's rows: row slot rr - l
          red[(rr - t.lane) * W + half * 32 + t.lane] = dw3[0];
        } else {
#pragma unroll
          for (int q = 0; q < NDW3; ++q)
            red[rr * W + q] = dw3[q];
        }
        if (half == 0)
          red[rr * W + D2] = db3;
      }
      __syncthreads();
      float *quart = red + 256 * W;
      for (int u = threadIdx.x; u < 4 * W; u += blockDim.x) {
        const int c = u % W, qr = u / W;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        for (int r = 0; r < 64; r += 4) {
          s0 += red[(qr * 64 + r) * W + c];
          s1 += red[(qr * 64 + r + 1) * W + c];
          s2 += red[(qr * 64 + r + 2) * W + c];
          s3 += red[(qr * 64 + r + 3) * W + c];
        }
        quart[u] = (s0 + s1) + (s2 + s3);
      }
      __syncthreads();
      if (threadIdx.x < W) {
        const float s = (quart[threadIdx.x] + quart[W + threadIdx.x]) + (quart[2 * W + threadIdx.x] + quart[3 * W + threadIdx.x]);
        if (threadIdx.x < D2)
          put(net.o_w3 + threadIdx.x, s);
        else
          put(net.o_b3, s);
      }
      __syncthreads();
      for (int q = threadIdx.x; q < net.n_params; q += blockDim.x)
        part[q] = sg[q + (q >> 5)];
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (t.warp == 0)
    umma::tmem_dealloc(tmem, 512);
  if (MODE == CRITIC_STEP)  // cross-CTA reduction (+ exchange) + optimizer update of this CTA's parameter slice
    gradient_tail(a.partials, net, a.tail, reinterpret_cast<float *>(smem + CM::WG0));
  if (gclk)
    gclk[2] = global_ns();
}

// ---------------------------------------------------------------------------------------------
// V(end states) on COMPACTED rows. The critic target r + gamma V(s') and the advantage pass need the
// value of a transition's END state only where it is not the next recorded start state: at the
// rollout's last step (the live environment state) and where an episode ended (the overflowed
// terminal state, bin_packing.h:54-61) -- about a third of the rows at T = 4. The critic-step and
// GAE kernels used to push ALL 128 rows of every tile through an end pass; for large batches this
// forward-only kernel evaluates just the needed rows first and the learner kernels read v_end[T][n].
//   live units  u < ceil(n / 128): the 128 environments [128 u, 128 u + 128) at step T - 1
//   scan units: 1024 consecutive (t, i) entries of rec_done with t < T - 1 each; the done entries are
//               compacted (warp ballots + prefix sums, deterministic) into passes of 128 rows
// Structure as the rollout kernel: 4 pipelines of 128 epilogue threads + one MMA-issuing warp, hidden
// activations only in tensor memory.
struct vend_args {
  const float *params;
  net3 net;
  learner_rows rows;
  float *v_end;  // [T][n]
  int n_live_units, n_units;
};
template <int D1, int D2>
struct vmap {
  static constexpr int NP = 4, CH = 1024;
  static constexpr uint32_t W1P = 0;
  static constexpr uint32_t W2_HI = W1P + D1 * 128, W2_LO = W2_HI + D2 * 128;
  static constexpr uint32_t FLOATS = W2_LO + D2 * 128;  // b1[D1] b2[D2] w3[64] b3[4]
  static constexpr int F_B1 = 0, F_B2 = D1, F_W3 = D1 + D2, F_B3 = D1 + D2 + 64, N_FLOATS = D1 + D2 + 68;
  static constexpr uint32_t LISTS = FLOATS + N_FLOATS * 4;  // per pipeline: uint16 lst[CH], int wtot[4], int npass[2]
  static constexpr uint32_t LIST_BYTES = CH * 2 + 32;
  static constexpr uint32_t WG0 = (LISTS + NP * LIST_BYTES + 1023) / 1024 * 1024;
  static constexpr uint32_t BARS = WG0 + NP * PANEL;
  static constexpr uint32_t TOTAL = BARS + 64;
  static constexpr int THREADS = 160 * NP;
};

template <int D0, int D1, int D2>
__global__ void __launch_bounds__((vmap<D1, D2>::THREADS), 1) fused_vend_kernel(vend_args a) {
  using VM = vmap<D1, D2>;
  constexpr int NP = VM::NP, CH = VM::CH, NB = 8, P = 2 * NB + 2;
  static_assert(D0 == 4 * NB, "observation width");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
  float *fl = reinterpret_cast<float *>(smem + VM::FLOATS);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + VM::BARS);  // [wg]: MMA completion
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + VM::BARS + 56);
  const net3 net = a.net;
  const learner_rows &L = a.rows;
  const tid_t t = thread_id();
  const bool issuer = t.warp >= 4 * NP;
  const int wg = issuer ? t.warp - 4 * NP : t.warp >> 2;
  const uint32_t sbase = umma::smem_u32(smem);

  umma::pdl_launch_dependents();
  if (t.warp == 0)
    umma::tmem_alloc(tmem_slot, 512);
  if (threadIdx.x == 0) {
    for (int q = 0; q < NP; ++q)
      umma::mbar_init(bars + q, 1);
    umma::fence_mbar_init();
  }
  {
    const float *Pm = stage_params_bulk(a.params, net.n_params, smem + VM::WG0, bars + 5);
    stage_w1_packed<D1>(Pm + net.o_w1, smem + VM::W1P);
    stage_weight_f16(Pm + net.o_w2, D2, D1, D2, 1.f, smem + VM::W2_HI, smem + VM::W2_LO);
    for (int i = threadIdx.x; i < D1; i += blockDim.x) fl[VM::F_B1 + i] = Pm[net.o_b1 + i];
    for (int i = threadIdx.x; i < D2; i += blockDim.x) fl[VM::F_B2 + i] = Pm[net.o_b2 + i];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) fl[VM::F_W3 + i] = i < D2 ? Pm[net.o_w3 + i] : 0.f;
    if (threadIdx.x == 0)
      fl[VM::F_B3] = Pm[net.o_b3];
  }
  __syncthreads();  // the scratch (observation panels) is free again
  zero_bytes(smem + VM::WG0, VM::BARS - VM::WG0);
  sync_after_smem_writes();
  const uint32_t tmem = *tmem_slot;

  const int nt = (int)blockIdx.x < a.n_units ? (a.n_units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  uint8_t *wsm = smem + VM::WG0 + wg * PANEL;
  const uint32_t wbase = sbase + VM::WG0 + wg * PANEL;
  const uint32_t tm = tmem + 128u * wg;  // ACC0 (layer 1) at +0, ACC1 (layer 2) at +64
  uint64_t *bar = bars + wg;
  uint16_t *lst = reinterpret_cast<uint16_t *>(smem + VM::LISTS + wg * VM::LIST_BYTES);
  int *wtot = reinterpret_cast<int *>(smem + VM::LISTS + wg * VM::LIST_BYTES + CH * 2);
  volatile int *npass_slot = wtot + 4;  // [2], alternating per unit
  uint32_t rp = 0;

  if (issuer) {
    int u = 0;
    for (int j = wg; j < nt; j += NP, u ^= 1) {
      ready_sync(wg, rp);  // the unit's pass count is published (the epilogue threads block on this one too)
      const int npass = npass_slot[u];
      for (int ps = 0; ps < npass; ++ps) {
        ready_sync(wg, rp);  // observations
        if (umma::elect_one()) {
          issue_gemm<D0 / 16, false, false, false, true>(tm, wbase, 0, sbase + VM::W1P, sbase + VM::W1P + 64, ID<D1>::FK_FK, false);
          umma::commit(bar);
        }
        __syncwarp();
        ready_sync(wg, rp);  // H1 (tensor memory)
        if (umma::elect_one()) {
          issue_gemm_ta<D1, D1 / 16, false>(tm + 64, tm, sbase + VM::W2_HI, sbase + VM::W2_LO, ID<D2>::FK_FK);
          umma::commit(bar);
        }
        __syncwarp();
      }
    }
  } else {
    const float *b1 = fl + VM::F_B1, *b2 = fl + VM::F_B2, *w3 = fl + VM::F_W3;
    const float b3 = fl[VM::F_B3];
    uint32_t phase = 0;
    auto wait_mma = [&]() {
      umma::mbar_wait(bar, phase);
      phase ^= 1;
      umma::fence_after_sync();
    };
    const long long scan_rows = (long long)(L.T - 1) * L.n;
    int u = 0;
    for (int j = wg; j < nt; j += NP, u ^= 1) {
      const int unit = blockIdx.x + j * gridDim.x;
      const bool live = unit < a.n_live_units;
      long long base = 0;
      int cnt;
      if (live) {
        cnt = min(TILE, L.n - TILE * unit);
      } else {
        // ---- compaction of the done entries of this chunk: thread r scans entries EPT r .. EPT r + EPT - 1
        constexpr int EPT = CH / TILE;
        static_assert(EPT == 8 || EPT == 16, "one 8- or 16-byte load of done flags per thread");
        base = (long long)(unit - a.n_live_units) * CH;
        const long long k0 = base + EPT * t.row;
        uint32_t bits = 0;
        if (k0 + EPT <= scan_rows && ((reinterpret_cast<uintptr_t>(L.rec_done) + k0) & (EPT - 1)) == 0) {
          uint32_t wv[4] = {0, 0, 0, 0};
          if (EPT == 16) {
            const uint4 f = *reinterpret_cast<const uint4 *>(L.rec_done + k0);
            wv[0] = f.x, wv[1] = f.y, wv[2] = f.z, wv[3] = f.w;
          } else {
            const uint2 f = *reinterpret_cast<const uint2 *>(L.rec_done + k0);
            wv[0] = f.x, wv[1] = f.y;
          }
#pragma unroll
          for (int q = 0; q < EPT / 4; ++q)
#pragma unroll
            for (int e = 0; e < 4; ++e)
              bits |= (((wv[q] >> (8 * e)) & 0xffu) ? 1u : 0u) << (4 * q + e);
        } else {
          for (int e = 0; e < EPT; ++e)
            if (k0 + e < scan_rows && L.rec_done[k0 + e])
              bits |= 1u << e;
        }
        const int mine = __popc(bits);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, o);
          if (t.lane >= o)
            incl += v;
        }
        if (t.lane == 31)
          wtot[t.w] = incl;
        asm volatile("bar.sync %0, %1;\n" ::"r"(9 + wg), "r"(128) : "memory");
        int off = incl - mine;
        cnt = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q < t.w)
            off += wtot[q];
          cnt += wtot[q];
        }
        for (uint32_t b = bits; b; b &= b - 1)
          lst[off++] = (uint16_t)(EPT * t.row + (__ffs(b) - 1));
        asm volatile("bar.sync %0, %1;\n" ::"r"(9 + wg), "r"(128) : "memory");  // list complete; wtot may be reused
      }
      const int npass = (cnt + TILE - 1) / TILE;
      if (t.row == 0)
        npass_slot[u] = npass;
      ready_sync(wg, rp);  // blocking on purpose: an empty unit must not let these threads run ahead of the issuer
      for (int ps = 0; ps < npass; ++ps) {
        const int slot = TILE * ps + t.row;
        const bool has = slot < cnt;
        long long k = 0;
        int tt = L.T - 1, i = TILE * unit + t.row;
        if (has && !live) {
          k = base + lst[slot];
          tt = (int)(k / L.n);
          i = (int)(k - (long long)tt * L.n);
        } else if (has) {
          k = (long long)tt * L.n + i;
        }
        int done = 0, act = 0;
        if (has) {
          done = live ? (int)L.rec_done[k] : 1;
          act = L.rec_action[k];
        }
        // end state: the overflowed terminal state (done) or the live state (last step)
        row_state<NB> x;
        const int8_t *src = done ? L.rec_state + (size_t)tt * P * L.stride + i : L.live_state + i;
#pragma unroll
        for (int q = 0; q < P; ++q) {
          x.v[q] = 0;
          if (has)
            x.v[q] = src[(size_t)q * L.stride];
        }
        fix_end_state<NB>(x, done, act);
        encode_row<NB>(wsm, t.row, x, L.inv_w, L.inv_h);
        ready_arrive(wg, rp);
        wait_mma();  // layer 1
        epi2_fwd<D1, true, false>(tm, t, b1, nullptr, nullptr);  // H1 only as a TMEM A operand
        ready_arrive(wg, rp);
        wait_mma();  // layer 2
        const float v = epi2_value<D2, false>(tm + 64, t, b2, w3, b3, nullptr);
        if (has)
          a.v_end[k] = v;
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (t.warp == 0)
    umma::tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// Rollout: agent::play_steps(T) (rl.h:325-360) for a tile of 128 environments per CTA iteration.
// The tile's int8 state planes stay in shared memory for all T steps; per step the CTA records the
// start state, encodes the observation panel (exact in 16 bits: multiples of 1/cap), runs the three
// forward GEMMs on the tensor cores, and one thread per environment does softmax -> action
// (sample / argmax / forced) -> environment::apply -> reward / done / reset / next item.
struct rollout_args {
  const float *params;  // flat fp32 parameters of the net
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

// ---------------------------------------------------------------------------------------------
// Rollout with NP tile pipelines per CTA (same structure as fused_policy_step_kernel): a pipeline
// = 128 environments, one epilogue thread per environment that keeps the env state in REGISTERS for
// all T steps, plus one MMA-issuing warp. The hidden activations exist ONLY in tensor memory: the
// epilogues write their packed bf16 hi / lo words over the accumulator columns they just read and
// the next GEMM takes its A operand from there; shared memory per pipeline = the observation panel.
// The layer-3 B operand stacks [hi(W3); lo(W3)] (N = 16: two MMAs per K step, the head adds columns
// j and 8 + j).
template <int D1, int D2, int NP>
struct rmap {
  static constexpr uint32_t W1P = 0;
  static constexpr uint32_t W2_HI = W1P + D1 * 128, W2_LO = W2_HI + D2 * 128;
  static constexpr uint32_t W3C = W2_LO + D2 * 128;  // rows 0..7 = hi(W3), rows 8..15 = lo(W3)
  static constexpr uint32_t FLOATS = W3C + 16 * 128;  // b1[D1] b2[D2] b3[16]
  static constexpr int F_B1 = 0, F_B2 = D1, F_B3 = D1 + D2, N_FLOATS = D1 + D2 + 16;
  static constexpr uint32_t WG0 = (FLOATS + N_FLOATS * 4 + 1023) / 1024 * 1024;
  static constexpr uint32_t X0 = 0, WG_BYTES = PANEL;  // per pipeline: the observation panel only
  static constexpr uint32_t BARS = WG0 + NP * WG_BYTES;
  static constexpr uint32_t TOTAL = BARS + 64;
  static_assert(TOTAL + 1024 <= 232448, "exceeds the 227 KB shared memory of an SM");
};

template <int D0, int D1, int D2, int NOUT, int NP>
__global__ void __launch_bounds__(160 * NP, 1) fused_rollout_kernel(rollout_args a) {
  using RM = rmap<D1, D2, NP>;
  constexpr int B = NOUT, P = 2 * B + 2;
  static_assert(D0 == 4 * B && NOUT == 8, "observation width");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);  // stays a shared-space pointer
  float *fl = reinterpret_cast<float *>(smem + RM::FLOATS);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + RM::BARS);  // [wg]: MMA completion
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + RM::BARS + 56);
  const net3 net = a.net;
  const env_params &ep = a.ep;
  const tid_t t = thread_id();
  const bool issuer = t.warp >= 4 * NP;                    // warp-uniform
  const int wg = issuer ? t.warp - 4 * NP : t.warp >> 2;   // pipeline index
  const uint32_t sbase = umma::smem_u32(smem);
  constexpr uint32_t TCOLS = NP <= 2 ? 256 : 512;

  umma::pdl_launch_dependents();
  if (t.warp == 0)
    umma::tmem_alloc(tmem_slot, TCOLS);
  if (threadIdx.x == 0) {
    for (int q = 0; q < NP; ++q)
      umma::mbar_init(bars + q, 1);
    umma::fence_mbar_init();
  }
  {
    const float *Pm = stage_params_bulk(a.params, net.n_params, smem + RM::WG0, bars + 5);
    const float *W3 = Pm + net.o_w3;
    stage_w1_packed<D1>(Pm + net.o_w1, smem + RM::W1P);
    stage_weight_f16(Pm + net.o_w2, D2, D1, D2, 1.f, smem + RM::W2_HI, smem + RM::W2_LO);
    for (int c = threadIdx.x; c < 16 * 8; c += blockDim.x) {
      int row = c >> 3, chunk = c & 7;
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int col = chunk * 8 + j;
        x[j] = col < D2 ? W3[(size_t)(row & 7) * D2 + col] : 0.f;
      }
      uint4 h, l;
      split8<false>(x, h, l);
      uint32_t off = umma::panel_chunk_off(row, chunk);
      *reinterpret_cast<uint4 *>(smem + RM::W3C + off) = row < 8 ? h : l;
    }
    for (int i = threadIdx.x; i < D1; i += blockDim.x) fl[RM::F_B1 + i] = Pm[net.o_b1 + i];
    for (int i = threadIdx.x; i < D2; i += blockDim.x) fl[RM::F_B2 + i] = Pm[net.o_b2 + i];
    for (int i = threadIdx.x; i < 16; i += blockDim.x) fl[RM::F_B3 + i] = i < net.d3 ? Pm[net.o_b3 + i] : 0.f;
  }
  __syncthreads();  // the scratch (observation panels) is free again
  zero_bytes(smem + RM::WG0, RM::BARS - RM::WG0);
  sync_after_smem_writes();
  const uint32_t tmem = *tmem_slot;

  // this CTA's tiles: blockIdx.x + j * gridDim.x, j < nt; pipeline wg takes j = wg, wg + NP, ...
  const int nt = (int)blockIdx.x < a.n_tiles ? (a.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  uint8_t *wsm = smem + RM::WG0 + wg * RM::WG_BYTES;
  const uint32_t wbase = sbase + RM::WG0 + wg * RM::WG_BYTES;
  const uint32_t tm = tmem + 128u * wg;  // ACC0 (layers 1, 3) at +0, ACC1 (layer 2) at +64
  uint64_t *bar = bars + wg;
  uint32_t rp = 0;

  if (issuer) {
    for (int j = wg; j < nt; j += NP)
      for (int tt = 0; tt < a.T; ++tt) {
        ready_sync(wg, rp);  // observations staged in the lo panel
        if (umma::elect_one()) {
          issue_gemm<D0 / 16, false, false, false, true>(tm, wbase + RM::X0, 0, sbase + RM::W1P,
                                                         sbase + RM::W1P + 64, ID<D1>::FK_FK, false);
          umma::commit(bar);
        }
        __syncwarp();
        ready_sync(wg, rp);  // H1
        if (umma::elect_one()) {
          // A = H1 from tensor memory (the epilogue's copy in ACC0)
          issue_gemm_ta<D1, D1 / 16, false>(tm + 64, tm, sbase + RM::W2_HI, sbase + RM::W2_LO, ID<D2>::FK_FK);
          umma::commit(bar);
        }
        __syncwarp();
        ready_sync(wg, rp);  // H2 (in place)
        if (umma::elect_one()) {
          // (hi(H2) + lo(H2)) . [hi(W3); lo(W3)], A = H2 from tensor memory (ACC1)
          issue_gemm_ta<D2, D2 / 16, false, false>(tm, tm + 64, sbase + RM::W3C, 0, ID<16>::FK_FK);
          umma::commit(bar);
        }
        __syncwarp();
      }
  } else {
    const float *b1 = fl + RM::F_B1, *b2 = fl + RM::F_B2, *b3 = fl + RM::F_B3;
    uint32_t phase = 0;
    auto wait_mma = [&]() {
      umma::mbar_wait(bar, phase);
      phase ^= 1;
      umma::fence_after_sync();
    };
    unsigned long long c_eps = 0, c_reward = 0, c_steps = 0;
    const size_t S = ep.stride;
    for (int j = wg; j < nt; j += NP) {
      const int tile = blockIdx.x + j * gridDim.x;
      const int i = tile * TILE + t.row;
      const bool owner = i < ep.n;
      row_state<B> st;  // live state of this thread's environment
      uint32_t my_draws = 0, my_steps = 0;
#pragma unroll
      for (int q = 0; q < P; ++q)
        st.v[q] = 0;
      if (owner) {
#pragma unroll
        for (int q = 0; q < P; ++q)
          st.v[q] = a.state[(size_t)q * S + i];
        my_draws = a.draws[i];
        my_steps = a.steps[i];
      }
      for (int tt = 0; tt < a.T; ++tt) {
        // ---- record the start state of step tt, stage the observations
        if (owner) {
#pragma unroll
          for (int q = 0; q < P; ++q)
            a.rec_state[((size_t)tt * P + q) * S + i] = (int8_t)st.v[q];
        }
        encode_row<B>(wsm + RM::X0, t.row, st, a.inv_w, a.inv_h);
        ready_arrive(wg, rp);
        // tape entries of this step (loads issued before the GEMMs, used in the head)
        const size_t k = (size_t)tt * ep.n + i;
        int forced_a = 0, tape_item = 0;
        double tape_u = 0.0;
        if (owner) {
          if (a.mode == DFRL_ACT_FORCED)
            forced_a = a.forced[k];
          else if (a.mode == DFRL_ACT_SAMPLE && a.u_tape)
            tape_u = a.u_tape[k];
          if (a.item_tape)
            tape_item = a.item_tape[k];
        }
        // (running these two epilogues as one copy of code in a 2-trip loop, as the policy step does, was measured
        //  0.3 % slower at 131 072 envs: this kernel's loop is 25 KB either way)
        wait_mma();  // layer 1
        epi2_fwd<D1, true, false>(tm, t, b1, nullptr, nullptr);  // H1 only as a TMEM A operand
        ready_arrive(wg, rp);
        wait_mma();  // layer 2
        epi2_fwd<D2, true, false>(tm + 64, t, b2, nullptr, nullptr);
        ready_arrive(wg, rp);
        wait_mma();  // layer 3
        // ---- head: softmax (no max subtraction, nn.h:382-392), action, environment::apply
        float v[16];
        tmem_load<16>(tm + t.lane_base, v);
        if (owner) {
          float p[NOUT], s = 0.f;
#pragma unroll
          for (int q = 0; q < NOUT; ++q) {
            p[q] = expf((v[q] + v[8 + q]) + b3[q]);
            s += p[q];
          }
#pragma unroll
          for (int q = 0; q < NOUT; ++q)
            p[q] = p[q] / s;
          float4 *pr = reinterpret_cast<float4 *>(a.rec_probs + k * B);
#pragma unroll
          for (int q = 0; q < NOUT / 4; ++q)
            pr[q] = make_float4(p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]);
          int act;
          if (a.mode == DFRL_ACT_FORCED) {
            act = forced_a;
          } else if (a.mode == DFRL_ACT_ARGMAX) {
            act = argmax_first(p, B);
          } else {
            double u = tape_u;
            if (!a.u_tape) {
              philox4 rr = philox4x32_10(ep.seed, (uint64_t)(ep.env_offset + i), my_steps, DFRL_STREAM_ACTION);
              u = philox_u53(rr.x, rr.y);
            }
            act = discrete_sample(p, B, u);
          }
          act = act < B ? act : B - 1;
          a.rec_action[k] = (uint8_t)act;
          // environment::apply (bin_packing.h:53-64) on the registers of this environment
          const int iw = st.v[2 * B], ih = st.v[2 * B + 1];
          int bw = 0, bh = 0;
#pragma unroll
          for (int b = 0; b < B; ++b)
            if (b == act) {
              bw = st.v[2 * b] - iw;
              bh = st.v[2 * b + 1] - ih;
            }
          const bool over = bw < 0 || bh < 0;
          const int s1 = a.item_tape ? (tape_item != 0) : draw_shape1(ep, i, my_draws);
#pragma unroll
          for (int b = 0; b < B; ++b) {
            if (over) {
              st.v[2 * b] = ep.cap_w;
              st.v[2 * b + 1] = ep.cap_h;
            } else if (b == act) {
              st.v[2 * b] = bw;
              st.v[2 * b + 1] = bh;
            }
          }
          st.v[2 * B] = s1 ? ep.iw0 : ep.iw1;
          st.v[2 * B + 1] = s1 ? ep.ih0 : ep.ih1;
          a.rec_done[k] = over;
          my_draws += 1;
          my_steps += 1;
          c_steps += 1;
          c_eps += over ? 1 : 0;
          c_reward += over ? 0 : 1;
        }
      }
      // ---- live state back to the environment
      if (owner) {
#pragma unroll
        for (int q = 0; q < P; ++q)
          a.state[(size_t)q * S + i] = (int8_t)st.v[q];
        a.draws[i] = my_draws;
        a.steps[i] = my_steps;
      }
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
  }
  umma::fence_before_sync();
  __syncthreads();
  if (t.warp == 0)
    umma::tmem_dealloc(tmem, TCOLS);
}

#include "fused_conv.cuh"
#include "conv_table.cuh"

// ---------------------------------------------------------------------------------------------
struct fused_state {
  net3 pnet, vnet;
  bool policy_ok, value_ok, rollout_ok;
  bool policy_conv;  // the policy is a conv1d_1 net 4-D1-D2-1 over the 8 bins (fused_conv.cuh)
  // conv1d_1 policy on its finite input domain (conv_table.cuh): logit table, fixed-point histogram of dY, max |A|
  bool policy_table;   // the optimizer steps run on the table (large batches)
  bool rollout_table;  // the rollout runs on the table (same range)
  int tbl_Dw, tbl_Dh;
  float *tbl_logits;
  unsigned long long *tbl_hist;  // [D] + one word: the bit pattern of max |A| in its low half
  bool tbl_reuse;                // a policy step may use a table that an earlier launch computed from the same parameters
  bool tbl_have;                 // tbl_logits holds the table of the policy parameters at version tbl_version
  uint64_t tbl_version;
  bool tbl_scale_valid;          // max |A| of the current advantages is known (reset by every rollout / GAE pass)
  int head_bwd;
  float *partials;  // [ctas][max params]
  unsigned *gridbar;  // [2] grid barrier of the gradient tail (arrivals, generation)
  int vend_mode;      // -1: by batch size, 0 / 1: never / always use the compacted V(end) pre-pass
  int ctas;
  long long *clk;  // [112] phase clocks of the last policy step (allocated on first request)
  long long *clk_critic;  // [112] the same for the critic step
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
  out->shared = (m->share_owner || !m->sharers.empty()) ? 1 : 0;
  return true;
}

// C - R - C - R - C - softmax / softmax_ce with 4 input channels, one output channel, 8 points
bool parse_conv3(const dfrl_mlp *m, net3 *out, int *tail_kind) {
  const auto &L = m->layers;
  if (L.size() != 6)
    return false;
  if (L[0].kind != DFRL_LAYER_CONV1D_1 || L[1].kind != DFRL_LAYER_RELU || L[2].kind != DFRL_LAYER_CONV1D_1 ||
      L[3].kind != DFRL_LAYER_RELU || L[4].kind != DFRL_LAYER_CONV1D_1)
    return false;
  if (L[5].kind != DFRL_LAYER_SOFTMAX && L[5].kind != DFRL_LAYER_SOFTMAX_CE)
    return false;
  if (m->share_owner || !m->sharers.empty())
    return false;
  if (L[0].in != 4 || L[0].points != 8 || L[2].in != L[0].out || L[4].in != L[2].out || L[4].out != 1)
    return false;
  *tail_kind = L[5].kind;
  out->d0 = 4;
  out->d1 = L[0].out;
  out->d2 = L[2].out;
  out->d3 = 1;
  out->o_w1 = (int)L[0].param_off;
  out->o_b1 = out->o_w1 + out->d0 * out->d1;
  out->o_w2 = (int)L[2].param_off;
  out->o_b2 = out->o_w2 + out->d1 * out->d2;
  out->o_w3 = (int)L[4].param_off;
  out->o_b3 = out->o_w3 + out->d2 * out->d3;
  out->n_params = m->n_params;
  out->shared = 0;
  return true;
}
bool conv_widths_ok(const net3 &n) { return (n.d1 == 128 && n.d2 == 64) || (n.d1 == 64 && n.d2 == 32); }

bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

// cudaFuncSetAttribute is a per-device setting: one bit per device ordinal (a process may own contexts
// on several GPUs).
template <typename K>
int set_smem_once(const dfrl_ctx *ctx, K kernel, int smem, unsigned long long *done) {
  const unsigned long long bit = 1ull << (ctx->device & 63);
  if (!(*done & bit)) {
    DFRL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    *done |= bit;
  }
  return DFRL_OK;
}

// Launch of a fused kernel with programmatic stream serialization (see umma::pdl_*), counted /
// profiled like DFRL_LAUNCH. The learner kernels' grid barrier needs every CTA resident at once:
// grid <= SM count at one CTA per SM, and a dependent grid is scheduled only after EVERY CTA of its
// predecessor has started (each issues launch_dependents first), so it can never take an SM from one.
// Measured (B200, 131 072 envs, CUDA-graphed learn phase): 0.5821 ms / iteration with the attribute,
// 0.5837 without -- inside noise: the gaps between the graph's kernel nodes are already short, and a
// dependent CTA cannot become resident before its predecessor releases the SM's shared memory. OFF by
// default; DFRL_PDL=1 enables it.
template <typename K, typename A>
int launch_fused(dfrl_ctx *ctx, K kernel, const char *name, int grid, int block, int smem, const A &args) {
  if (ctx->profiling)
    dfrl_profile_mark(ctx, name, 0);
  static const bool pdl = getenv("DFRL_PDL") && atoi(getenv("DFRL_PDL")) != 0;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args);
  ctx->launches++;
  if (ctx->profiling)
    dfrl_profile_mark(ctx, name, 1);
  if (e != cudaSuccess) {
    dfrl_set_error("launch of %s failed: %s", name, cudaGetErrorString(e));
    return DFRL_ERR_CUDA;
  }
  return DFRL_OK;
}

template <int D0, int D1, int D2, int NOUT, bool PROBE>
int launch_policy_step_p(dfrl_ctx *ctx, const policy_step_args &a, int ctas) {
  constexpr int smem = pmap<D1, D2>::TOTAL + 1024;
  static unsigned long long attr = 0;
  DFRL_TRY(set_smem_once(ctx, fused_policy_step_kernel<D0, D1, D2, NOUT, PROBE>, smem, &attr));
  return launch_fused(ctx, fused_policy_step_kernel<D0, D1, D2, NOUT, PROBE>, "(fused_policy_step_kernel<D0, D1, D2, NOUT>)",
                            ctas, pmap<D1, D2>::THREADS, smem, a);
}
template <int D0, int D1, int D2, int NOUT>
int launch_policy_step(dfrl_ctx *ctx, const policy_step_args &a, int ctas) {
  if (a.clk && D1 == 64)  // the phase-clock tools asked for stamps (64-wide nets only)
    return launch_policy_step_p<D0, D1, D2, NOUT, D1 == 64>(ctx, a, ctas);
  return launch_policy_step_p<D0, D1, D2, NOUT, false>(ctx, a, ctas);
}

template <int D0, int D1, int D2, bool EG, bool PROBE>
int launch_critic_step_eg(dfrl_ctx *ctx, const critic_args &a, int ctas) {
  constexpr int smem = cmap<D1, D2, CRITIC_STEP>::TOTAL + 1024;
  static unsigned long long attr = 0;
  DFRL_TRY(set_smem_once(ctx, fused_critic_kernel<D0, D1, D2, CRITIC_STEP, EG, PROBE>, smem, &attr));
  return launch_fused(ctx, fused_critic_kernel<D0, D1, D2, CRITIC_STEP, EG, PROBE>,
                            "(fused_critic_kernel<D0, D1, D2, CRITIC_STEP, EG>)",
                            ctas, cmap<D1, D2, CRITIC_STEP>::THREADS, smem, a);
}
template <int D0, int D1, int D2>
int launch_critic_step(dfrl_ctx *ctx, const critic_args &a, int ctas) {
  if (a.clk && a.v_end && D1 == 64 && D2 == 64)  // tools/critic_phase_clocks.py (64-wide nets, compacted end rows)
    return launch_critic_step_eg<D0, D1, D2, true, D1 == 64 && D2 == 64>(ctx, a, ctas);
  return a.v_end ? launch_critic_step_eg<D0, D1, D2, true, false>(ctx, a, ctas) : launch_critic_step_eg<D0, D1, D2, false, false>(ctx, a, ctas);
}

template <int D0, int D1, int D2, bool EG>
int launch_gae_eg(dfrl_ctx *ctx, const critic_args &a, int ctas) {
  constexpr int smem = cmap<D1, D2, CRITIC_GAE>::TOTAL + 1024;
  static unsigned long long attr = 0;
  DFRL_TRY(set_smem_once(ctx, fused_critic_kernel<D0, D1, D2, CRITIC_GAE, EG>, smem, &attr));
  return launch_fused(ctx, fused_critic_kernel<D0, D1, D2, CRITIC_GAE, EG>, "(fused_critic_kernel<D0, D1, D2, CRITIC_GAE, EG>)", ctas,
                      cmap<D1, D2, CRITIC_GAE>::THREADS, smem, a);
}
template <int D0, int D1, int D2>
int launch_gae(dfrl_ctx *ctx, const critic_args &a, int ctas) {
  return a.v_end ? launch_gae_eg<D0, D1, D2, true>(ctx, a, ctas) : launch_gae_eg<D0, D1, D2, false>(ctx, a, ctas);
}

template <int D0, int D1, int D2>
int launch_vend(dfrl_ctx *ctx, const vend_args &a, int ctas) {
  constexpr int smem = vmap<D1, D2>::TOTAL + 1024;
  static unsigned long long attr = 0;
  DFRL_TRY(set_smem_once(ctx, fused_vend_kernel<D0, D1, D2>, smem, &attr));
  return launch_fused(ctx, fused_vend_kernel<D0, D1, D2>, "(fused_vend_kernel<D0, D1, D2>)", ctas, vmap<D1, D2>::THREADS, smem, a);
}

// Four rollout pipelines: 131 072 envs = 1024 tiles = 6.9 per SM, i.e. two tile rounds (three
// pipelines need three rounds: measured 74 us vs 64 us).
template <int D0, int D1, int D2, int NOUT>
int launch_rollout(dfrl_ctx *ctx, const rollout_args &a, int ctas) {
  constexpr int NP = 4;
  constexpr int smem = rmap<D1, D2, NP>::TOTAL + 1024;
  static unsigned long long attr = 0;
  DFRL_TRY(set_smem_once(ctx, fused_rollout_kernel<D0, D1, D2, NOUT, NP>, smem, &attr));
  return launch_fused(ctx, fused_rollout_kernel<D0, D1, D2, NOUT, NP>, "(fused_rollout_kernel<D0, D1, D2, NOUT, NP>)", ctas,
                      160 * NP, smem, a);
}

template <int D1, int D2, bool PROBE>
int launch_conv_policy_step_p(dfrl_ctx *ctx, const conv_step_args &a, int ctas) {
  constexpr int smem = cvmap<D1, D2>::TOTAL + 1024;
  static unsigned long long attr = 0;
  DFRL_TRY(set_smem_once(ctx, fused_conv_policy_step_kernel<D1, D2, PROBE>, smem, &attr));
  return launch_fused(ctx, fused_conv_policy_step_kernel<D1, D2, PROBE>, "(fused_conv_policy_step_kernel<D1, D2>)", ctas,
                      cvmap<D1, D2>::THREADS, smem, a);
}
template <int D1, int D2>
int launch_conv_policy_step(dfrl_ctx *ctx, const conv_step_args &a, int ctas) {
  if (a.clk && D1 == 128)  // tools/conv_phase_clocks.py asked for stamps (reference PPO net only)
    return launch_conv_policy_step_p<D1, D2, D1 == 128>(ctx, a, ctas);
  return launch_conv_policy_step_p<D1, D2, false>(ctx, a, ctas);
}
template <int D1, int D2>
int launch_conv_rollout(dfrl_ctx *ctx, const rollout_args &a, int ctas) {
  constexpr int NP = 2;
  constexpr int smem = cvrmap<D1, D2, NP>::TOTAL + 1024;
  static unsigned long long attr = 0;
  DFRL_TRY(set_smem_once(ctx, fused_conv_rollout_kernel<D1, D2, NP>, smem, &attr));
  return launch_fused(ctx, fused_conv_rollout_kernel<D1, D2, NP>, "(fused_conv_rollout_kernel<D1, D2, NP>)", ctas, 160 * NP,
                      smem, a);
}

// conv1d_1 policy on its input domain (conv_table.cuh)
template <int D1, int D2>
int launch_conv_table_forward(dfrl_ctx *ctx, const float *params, const net3 &net, float inv_w, float inv_h, int Dw, int Dh,
                              float *logits, const uint8_t *present) {
  const int D = Dw * Dh * Dw * Dh, grid = ceil_div(D, 16) < 4 * ctx->sm_count ? (int)ceil_div(D, 16) : 4 * ctx->sm_count;  // 8 warps x 2 entries
  DFRL_LAUNCH(ctx, (conv_table_forward_kernel<D1, D2>), grid, 256, 0, params, net, inv_w, inv_h, Dw, Dh, logits, present);
  return DFRL_OK;
}
template <int D1, int D2>
int launch_conv_table_step(dfrl_ctx *ctx, const conv_table_args &ta, int ctas, bool scan_adv, bool forward) {
  const conv_step_args &a = ta.s;
  const int D = ta.Dw * ta.Dh * ta.Dw * ta.Dh;
  const long long rows = (long long)a.T * a.n;
  // the histogram is cleared for every step; max |A| (the fixed-point scale) only when the advantages are new: the
  // k steps of a PPO iteration share them
  DFRL_CUDA(cudaMemsetAsync(ta.hist, 0, sizeof(unsigned long long) * (D + (scan_adv ? 1 : 0)), ctx->stream));
  if (scan_adv)
    DFRL_LAUNCH(ctx, conv_table_absmax_kernel, ctx->sm_count, 256, 0, a.adv, rows, const_cast<unsigned *>(ta.maxbits));
  if (forward)  // (not after a rollout on the same parameters: PPO's first step sees the rollout's table)
    DFRL_TRY((launch_conv_table_forward<D1, D2>(ctx, a.params, a.net, a.inv_w, a.inv_h, ta.Dw, ta.Dh, const_cast<float *>(ta.logits),
                                                nullptr)));
  const int smem = 12 * D;
  static unsigned long long attr = 0;
  DFRL_TRY(set_smem_once(ctx, conv_table_head_kernel, smem, &attr));
  DFRL_LAUNCH(ctx, conv_table_head_kernel, 2 * ctx->sm_count, 256, smem, ta);
  DFRL_LAUNCH(ctx, (conv_table_backward_kernel<D1, D2>), ctas, 256, 0, ta);
  return DFRL_OK;
}
int launch_conv_table_rollout(dfrl_ctx *ctx, const rollout_args &a, const float *logits, int Dw, int Dh) {
  const int smem = 4 * Dw * Dh * Dw * Dh;
  static unsigned long long attr = 0;
  DFRL_TRY(set_smem_once(ctx, conv_table_rollout_kernel, smem, &attr));
  DFRL_LAUNCH(ctx, conv_table_rollout_kernel, (int)ceil_div(a.ep.n, 128), 128, smem, a, logits, Dw, Dh);
  return DFRL_OK;
}

bool widths_ok(const net3 &n) {
  return n.d0 == 32 && ((n.d1 == 64 && n.d2 == 64) || (n.d1 == 16 && n.d2 == 16));
}
// Value nets: also 32-64-32-1, the reference's own critic (ppo_training.cc:19-26, ac_training.cc:18-25).
bool value_widths_ok(const net3 &n) { return widths_ok(n) || (n.d0 == 32 && n.d1 == 64 && n.d2 == 32); }
// CALL(D0, D1, D2) for the value net's widths
#define DFRL_VNET_DISPATCH(net, CALL)      \
  do {                                     \
    if ((net).d1 == 64 && (net).d2 == 64)  \
      CALL(32, 64, 64);                    \
    else if ((net).d1 == 64)               \
      CALL(32, 64, 32);                    \
    else                                   \
      CALL(32, 16, 16);                    \
  } while (0)

learner_rows make_rows(dfrl_trainer *t) {
  learner_rows r;
  r.rec_state = t->rec_state;
  r.live_state = t->env->state;
  r.rec_action = t->rec_action;
  r.rec_done = t->rec_done;
  r.n = t->n;
  r.stride = t->stride;
  r.T = t->L;
  r.E = TILE / t->L;
  r.B = t->B;
  r.inv_w = 1.0f / (float)t->env->cfg.cap_w;
  r.inv_h = 1.0f / (float)t->env->cfg.cap_h;
  return r;
}

critic_args make_critic_args(dfrl_trainer *t, fused_state *f) {
  critic_args a;
  a.params = t->value->params;
  a.net = f->vnet;
  a.rows = make_rows(t);
  a.n_tiles = ceil_div(t->n, a.rows.E);
  a.gamma = t->cfg.gamma;
  a.lambda = t->cfg.lambda;
  a.targets_out = t->targets;
  a.adv_out = t->adv;
  a.partials = f->partials;
  memset(&a.tail, 0, sizeof(a.tail));
  a.v_end = nullptr;
  a.v_end_out = nullptr;
  a.adv_maxbits = nullptr;
  a.clk = f->clk_critic;
  return a;
}

// The gradient tail of a learner kernel: reduction into grad_dev, and -- when `opt` is given --
// the optimizer update (single rank) or exchange over NVLink peer memory + update (several ranks:
// opt given means the peers are attached). opt == null: the caller exchanges / updates itself.
int make_tail(dfrl_trainer *t, fused_state *f, const net3 &net, int ctas, float *grad_dev, const dfrl_opt_spec *opt,
              grad_tail *tail) {
  dfrl_ctx *ctx = t->ctx;
  // a CTA's slice of ceil(P / ctas) parameters is walked 64 entries at a time: any ctas works
  memset(tail, 0, sizeof(*tail));
  tail->grad = grad_dev;
  tail->bar = f->gridbar;
  if (opt)
    tail->opt = *opt;
  if (opt && ctx->nranks > 1) {
    DFRL_CHECK((size_t)net.n_params <= DFRL_P2P_CAP, "flat gradient exceeds the exchange slot");
    DFRL_CHECK(ctx->p2p.attached, "gradient exchange without attached peers");
    for (int r = 0; r < ctx->nranks; ++r)
      tail->v.peer[r] = ctx->p2p.peer[r];
    tail->v.nranks = ctx->nranks;
    tail->v.rank = ctx->rank;
  }
  (void)ctas;
  return DFRL_OK;
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
  const dfrl_env_config &ec = t->env->cfg;
  // observations must be exact in 16 bits and bounded by 1 in magnitude (prep kernel's bounds)
  if (!is_pow2(ec.cap_w) || !is_pow2(ec.cap_h) || ec.cap_w > 64 || ec.cap_h > 64)
    return DFRL_ERR_UNSUPPORTED;
  for (int s = 0; s < 2; ++s)
    if (ec.item_w[s] < 0 || ec.item_h[s] < 0 || ec.item_w[s] > ec.cap_w || ec.item_h[s] > ec.cap_h)
      return DFRL_ERR_UNSUPPORTED;
  fused_state *f = new fused_state();
  memset(f, 0, sizeof(*f));
  int ptail = -1, vtail = -1;
  f->policy_ok = parse_net3(t->policy, &f->pnet, &ptail) && ptail != -1 && f->pnet.d3 == 8 && widths_ok(f->pnet);
  if (!f->policy_ok && parse_conv3(t->policy, &f->pnet, &ptail) && conv_widths_ok(f->pnet))
    f->policy_ok = f->policy_conv = true;
  f->value_ok = t->value && parse_net3(t->value, &f->vnet, &vtail) && vtail == -1 && f->vnet.d3 == 1 &&
                value_widths_ok(f->vnet);
  f->rollout_ok = f->policy_ok;
  f->head_bwd = ptail == DFRL_LAYER_SOFTMAX ? HEAD_JACOBIAN : HEAD_IDENTITY;
  if (!f->policy_ok && !f->value_ok) {
    delete f;
    return DFRL_ERR_UNSUPPORTED;
  }
  // (the learner kernels stage a CTA's partial gradient in 33 KB of shared memory)
  if ((f->policy_ok && !f->policy_conv && t->policy->n_params > 8192) || (f->value_ok && t->value->n_params > 8192)) {
    delete f;
    return DFRL_ERR_UNSUPPORTED;
  }
  f->ctas = t->ctx->sm_count;
  f->vend_mode = -1;
  int maxp = t->policy->n_params;
  if (t->value && t->value->n_params > maxp)
    maxp = t->value->n_params;
  bool ok = cudaMalloc(&f->partials, sizeof(float) * (size_t)f->ctas * partial_stride(maxp)) == cudaSuccess;
  if (ok)
    ok = cudaMalloc(&f->gridbar, 2 * sizeof(unsigned)) == cudaSuccess &&
         cudaMemsetAsync(f->gridbar, 0, 2 * sizeof(unsigned), t->ctx->stream) == cudaSuccess;
  if (!ok) {
    cudaFree(f->gridbar);
    cudaFree(f->partials);
    delete f;
    return DFRL_ERR_CUDA;
  }
  if (f->policy_conv) {
    const char *e = getenv("DFRL_CONV_TABLE");
    const long long Dw = t->env->cfg.cap_w + 1, Dh = t->env->cfg.cap_h + 1, D = Dw * Dh * Dw * Dh;
    // (from 8 192 environments x 4 steps on: below that the table's fixed cost per optimizer step -- 6 561 entries
    //  forward, histogram flush, a backward pass of few CTAs -- exceeds the tensor-core kernels' whole step;
    //  DFRL_CONV_TABLE=1 / 0 forces / forbids it)
    const bool big = (long long)t->n * t->L >= 32768;
    if ((!e || atoi(e) != 0) && D <= 16384) {
      f->tbl_Dw = (int)Dw, f->tbl_Dh = (int)Dh;
      if (cudaMalloc(&f->tbl_logits, sizeof(float) * D) == cudaSuccess &&
          cudaMalloc(&f->tbl_hist, sizeof(unsigned long long) * (D + 1) + (size_t)D) == cudaSuccess) {
        // (the table rollout alone was measured slower than the tensor-core rollout at 1 024 / 4 096 envs:
        //  two launches and 6 561 table entries against one launch)
        f->policy_table = f->rollout_table = e ? true : big;
        const char *r = getenv("DFRL_TABLE_REUSE");  // 0: every policy step computes its own table (tests)
        f->tbl_reuse = !r || atoi(r) != 0;
      }
    }
  }
  if (f->policy_conv)
    if (const char *e = getenv("DFRL_CONV_LOLO")) {
      const int v = atoi(e) != 0;
      cudaMemcpyToSymbolAsync(c_conv_lolo, &v, sizeof(int), 0, cudaMemcpyHostToDevice, t->ctx->stream);
      cudaStreamSynchronize(t->ctx->stream);
    }
  t->fused_impl = f;
  return DFRL_OK;
}

bool dfrl_fused_covers_iteration(const dfrl_trainer *t) {
  const fused_state *f = (const fused_state *)t->fused_impl;
  return f && f->policy_ok && f->value_ok && f->rollout_ok && t->cfg.algo != DFRL_ALGO_KL_PPO;
}

bool dfrl_fused_covers_critic(const dfrl_trainer *t) {
  const fused_state *f = (const fused_state *)t->fused_impl;
  return f && f->value_ok && f->rollout_ok;
}

void dfrl_fused_detach(dfrl_trainer *t) {
  fused_state *f = (fused_state *)t->fused_impl;
  if (f) {
    cudaFree(f->partials);
    cudaFree(f->tbl_logits);
    cudaFree(f->tbl_hist);
    cudaFree(f->clk);
    cudaFree(f->clk_critic);
    cudaFree(f->gridbar);
    delete f;
  }
  t->fused_impl = nullptr;
}

// Test hook: caps the grid of the fused learner kernels (persistent CTAs) so that small problems
// exercise the multi-tile steady state of both tile pipelines. ctas <= 0 restores one CTA per SM.
extern "C" int dfrl_debug_set_fused_ctas(dfrl_trainer *t, int ctas) {
  DFRL_CHECK(t, "null trainer");
  fused_state *f = (fused_state *)t->fused_impl;
  DFRL_CHECK(f, "fused path not attached");
  f->ctas = (ctas > 0 && ctas < t->ctx->sm_count) ? ctas : t->ctx->sm_count;
  return DFRL_OK;
}

// Test hook: forces (1, 2) / forbids (0) the compacted V(end) evaluation of the critic-step and GAE kernels,
// -1 restores the choice by batch size. 1: the GAE kernel's pipelines evaluate the end rows of their own
// tiles (the default when compaction is on), 2: fused_vend_kernel runs before the GAE kernel too.
extern "C" int dfrl_debug_set_vend(dfrl_trainer *t, int mode) {
  DFRL_CHECK(t, "null trainer");
  fused_state *f = (fused_state *)t->fused_impl;
  DFRL_CHECK(f, "fused path not attached");
  f->vend_mode = mode < 0 ? -1 : (mode > 2 ? 1 : mode);
  return DFRL_OK;
}

extern "C" int dfrl_trainer_fused_coverage(dfrl_trainer *t, int *mask) {
  DFRL_CHECK(t && mask, "null argument");
  const fused_state *f = (const fused_state *)t->fused_impl;
  *mask = 0;
  if (f) {
    const bool kl = t->cfg.algo == DFRL_ALGO_KL_PPO;
    *mask = (f->rollout_ok ? DFRL_FUSED_ROLLOUT : 0) | (f->value_ok ? DFRL_FUSED_CRITIC : 0) |
            (f->policy_ok && !kl ? DFRL_FUSED_POLICY : 0);
    if (dfrl_fused_covers_iteration(t) && (t->ctx->nranks == 1 || t->ctx->p2p.attached))
      *mask |= DFRL_FUSED_GRAPH;
  }
  return DFRL_OK;
}

// Phase clocks (SM cycles) of CTA 0 of the fused policy step: 12 stamps per tile, first 8 tiles.
// The first call arms the instrumentation (returns zeros); later calls return the last launch.
extern "C" int dfrl_debug_policy_clocks(dfrl_trainer *t, long long *out_host, int n) {
  DFRL_CHECK(t && out_host && n > 0 && n <= 752, "bad argument");  // 112 phase stamps of CTA 0 + 4 per CTA (160 CTAs)
  fused_state *f = (fused_state *)t->fused_impl;
  DFRL_CHECK(f, "fused path not attached");
  if (!f->clk) {
    DFRL_CUDA(cudaMalloc(&f->clk, sizeof(long long) * 752));
    DFRL_CUDA(cudaMemsetAsync(f->clk, 0, sizeof(long long) * 752, t->ctx->stream));
  }
  DFRL_CUDA(cudaMemcpyAsync(out_host, f->clk, sizeof(long long) * n, cudaMemcpyDeviceToHost, t->ctx->stream));
  DFRL_CUDA(cudaStreamSynchronize(t->ctx->stream));
  return DFRL_OK;
}

// The same for the critic step (pipeline 0 of CTA 0): 15 stamps per tile, first 7 tiles.
extern "C" int dfrl_debug_critic_clocks(dfrl_trainer *t, long long *out_host, int n) {
  DFRL_CHECK(t && out_host && n > 0 && n <= 752, "bad argument");
  fused_state *f = (fused_state *)t->fused_impl;
  DFRL_CHECK(f, "fused path not attached");
  if (!f->clk_critic) {
    DFRL_CUDA(cudaMalloc(&f->clk_critic, sizeof(long long) * 752));
    DFRL_CUDA(cudaMemsetAsync(f->clk_critic, 0, sizeof(long long) * 752, t->ctx->stream));
  }
  DFRL_CUDA(cudaMemcpyAsync(out_host, f->clk_critic, sizeof(long long) * n, cudaMemcpyDeviceToHost, t->ctx->stream));
  DFRL_CUDA(cudaStreamSynchronize(t->ctx->stream));
  return DFRL_OK;
}

// One policy gradient (forward + loss + backward over all L*n rows) into grad_dev.
int dfrl_fused_policy_gradient(dfrl_trainer *t, int loss_kind, float *grad_dev, const dfrl_opt_spec *opt) {
  fused_state *f = (fused_state *)t->fused_impl;
  if (!f || !f->policy_ok || loss_kind == DFRL_LOSS_KL)
    return DFRL_ERR_UNSUPPORTED;
  if (f->policy_conv) {
    conv_step_args c;
    c.params = t->policy->params;
    c.net = f->pnet;
    c.rec_state = t->rec_state;
    c.rec_action = t->rec_action;
    c.adv = t->adv;
    c.p_old = t->rec_probs;
    c.n = t->n;
    c.stride = t->stride;
    c.T = t->L;
    c.inv_w = 1.0f / (float)t->env->cfg.cap_w;
    c.inv_h = 1.0f / (float)t->env->cfg.cap_h;
    c.n_tiles = (int)ceil_div((long long)t->n * t->L, TILE / 8);
    c.loss_kind = loss_kind;
    c.head_bwd = f->head_bwd;
    c.partials = f->partials;
    c.clk = f->clk;
    int ctas = c.n_tiles < f->ctas ? c.n_tiles : f->ctas;
    DFRL_TRY(make_tail(t, f, f->pnet, ctas, grad_dev, opt, &c.tail));
    if (f->policy_table) {  // the net on its finite input domain: table, histogram of dY, one backward pass
      conv_table_args ta;
      ta.s = c;
      ta.logits = f->tbl_logits;
      ta.hist = f->tbl_hist;
      ta.Dw = f->tbl_Dw, ta.Dh = f->tbl_Dh;
      ta.maxbits = reinterpret_cast<const unsigned *>(f->tbl_hist + (size_t)ta.Dw * ta.Dh * ta.Dw * ta.Dh);
      const bool scan = !f->tbl_scale_valid;
      const bool fwd = !(f->tbl_reuse && f->tbl_have && f->tbl_version == t->policy->version);
      if (f->pnet.d1 == 128)
        DFRL_TRY((launch_conv_table_step<128, 64>(t->ctx, ta, ctas, scan, fwd)));
      else
        DFRL_TRY((launch_conv_table_step<64, 32>(t->ctx, ta, ctas, scan, fwd)));
      f->tbl_scale_valid = true;
      f->tbl_have = true, f->tbl_version = t->policy->version;
    } else if (f->pnet.d1 == 128)
      DFRL_TRY((launch_conv_policy_step<128, 64>(t->ctx, c, ctas)));
    else
      DFRL_TRY((launch_conv_policy_step<64, 32>(t->ctx, c, ctas)));
    if (opt)
      dfrl_mlp_params_changed(t->policy);
    return DFRL_OK;
  }
  policy_step_args a;
  a.params = t->policy->params;
  a.net = f->pnet;
  a.rows = make_rows(t);
  a.adv = t->adv;
  a.p_old = t->rec_probs;
  a.n_tiles = ceil_div(t->n, a.rows.E);
  a.loss_kind = loss_kind;
  a.head_bwd = f->head_bwd;
  a.partials = f->partials;
  a.clk = f->clk;
  a.clk_cta = getenv("DFRL_CLK_CTA") ? atoi(getenv("DFRL_CLK_CTA")) : 0;  // (debug: tools/policy_phase_clocks.py)
  int ctas = a.n_tiles < f->ctas ? a.n_tiles : f->ctas;
  DFRL_TRY(make_tail(t, f, f->pnet, ctas, grad_dev, opt, &a.tail));
  if (f->pnet.d1 == 64)
    DFRL_TRY((launch_policy_step<32, 64, 64, 8>(t->ctx, a, ctas)));
  else
    DFRL_TRY((launch_policy_step<32, 16, 16, 8>(t->ctx, a, ctas)));
  if (opt)
    dfrl_mlp_params_changed(t->policy);
  return DFRL_OK;
}

// V(end) of the rows that end a trajectory (last step, finished episodes) with the CURRENT critic, on
// compacted rows, into t->v_end. Used by the critic-step and GAE kernels of large batches instead of
// their own end pass over every row (DFRL_VEND=0 / 1 forces the choice).
static bool use_vend(const dfrl_trainer *t, const fused_state *f) {
  static const char *env = getenv("DFRL_VEND");
  if (f->vend_mode >= 0)
    return f->vend_mode != 0;
  if (env)
    return atoi(env) != 0;
  return (long long)t->n * t->L >= 65536;
}
static int fused_vend(dfrl_trainer *t, fused_state *f) {
  vend_args a;
  a.params = t->value->params;
  a.net = f->vnet;
  a.rows = make_rows(t);
  a.v_end = t->v_end;
  a.n_live_units = ceil_div(t->n, TILE);
  const long long scan_rows = (long long)(t->L - 1) * t->n;
  a.n_units = a.n_live_units + ceil_div(scan_rows, vmap<64, 64>::CH);
  int ctas = a.n_units < f->ctas ? a.n_units : f->ctas;
#define CALL(A, B, C) return launch_vend<A, B, C>(t->ctx, a, ctas)
  DFRL_VNET_DISPATCH(f->vnet, CALL);
#undef CALL
}

// update_value_model (policy_gradient.h:196-218) up to the gradient: writes t->targets and grad_dev.
int dfrl_fused_critic_gradient(dfrl_trainer *t, float *grad_dev, const dfrl_opt_spec *opt) {
  fused_state *f = (fused_state *)t->fused_impl;
  if (f)
    f->tbl_scale_valid = false;  // every VALUE phase ends in new advantages (also on the layered critic path)
  if (!f || !f->value_ok)
    return DFRL_ERR_UNSUPPORTED;
  critic_args a = make_critic_args(t, f);
  if (use_vend(t, f)) {
    DFRL_TRY(fused_vend(t, f));
    a.v_end = t->v_end;
  }
  int ctas = a.n_tiles < f->ctas ? a.n_tiles : f->ctas;
  DFRL_TRY(make_tail(t, f, f->vnet, ctas, grad_dev, opt, &a.tail));
#define CALL(A, B, C) DFRL_TRY((launch_critic_step<A, B, C>(t->ctx, a, ctas)))
  DFRL_VNET_DISPATCH(f->vnet, CALL);
#undef CALL
  if (opt)
    dfrl_mlp_params_changed(t->value);
  return DFRL_OK;
}

int dfrl_fused_learn_key(const dfrl_trainer *t) {
  const fused_state *f = (const fused_state *)t->fused_impl;
  return f && f->policy_table && f->tbl_reuse && f->tbl_have && f->tbl_version == t->policy->version ? 1 : 0;
}

// calculate_advantage (policy_gradient.h:220-281) with the current (updated) critic: writes t->adv.
int dfrl_fused_gae(dfrl_trainer *t) {
  fused_state *f = (fused_state *)t->fused_impl;
  if (f)
    f->tbl_scale_valid = false;  // new advantages
  if (!f || !f->value_ok)
    return DFRL_ERR_UNSUPPORTED;
  critic_args a = make_critic_args(t, f);
  if (use_vend(t, f)) {
    static const bool local = !getenv("DFRL_GAE_LOCAL_END") || atoi(getenv("DFRL_GAE_LOCAL_END")) != 0;
    a.v_end = t->v_end;
    const int ctas0 = a.n_tiles < f->ctas ? a.n_tiles : f->ctas;
    if (local && f->vend_mode != 2 && ceil_div(a.n_tiles, ctas0) <= 1000)  // (16-bit list entries: < 256 tiles per pipeline)
      a.v_end_out = t->v_end;  // the GAE kernel's pipelines evaluate the end rows of their own tiles first
    else
      DFRL_TRY(fused_vend(t, f));
  }
  if (f->policy_table) {  // the table path's fixed-point scale (max |A|) comes out of this kernel: no scan pass
    unsigned long long *mb = f->tbl_hist + (size_t)f->tbl_Dw * f->tbl_Dh * f->tbl_Dw * f->tbl_Dh;
    DFRL_CUDA(cudaMemsetAsync(mb, 0, sizeof(*mb), t->ctx->stream));
    a.adv_maxbits = reinterpret_cast<unsigned *>(mb);
  }
  int ctas = a.n_tiles < f->ctas ? a.n_tiles : f->ctas;
#define CALL(A, B, C) DFRL_TRY((launch_gae<A, B, C>(t->ctx, a, ctas)))
  DFRL_VNET_DISPATCH(f->vnet, CALL);
#undef CALL
  if (a.adv_maxbits)
    f->tbl_scale_valid = true;
  return DFRL_OK;
}

// agent::play_steps(L) for every env in one launch (AC / PPO / KL-PPO rollouts).
int dfrl_fused_rollout(dfrl_trainer *t, const uint8_t *items_dev, const uint8_t *actions_dev,
                       const double *u_dev) {
  fused_state *f = (fused_state *)t->fused_impl;
  if (f)
    f->tbl_scale_valid = false;  // (the layered GAE path does not pass through dfrl_fused_gae)
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
  if (f->policy_conv)
    a.n_tiles = ceil_div(t->n, TILE / 8);
  int ctas = a.n_tiles < f->ctas ? a.n_tiles : f->ctas;
  if (f->policy_conv && f->rollout_table) {
    if (f->pnet.d1 == 128)
      DFRL_TRY((launch_conv_table_forward<128, 64>(t->ctx, a.params, a.net, a.inv_w, a.inv_h, f->tbl_Dw, f->tbl_Dh, f->tbl_logits, nullptr)));
    else
      DFRL_TRY((launch_conv_table_forward<64, 32>(t->ctx, a.params, a.net, a.inv_w, a.inv_h, f->tbl_Dw, f->tbl_Dh, f->tbl_logits, nullptr)));
    f->tbl_have = true, f->tbl_version = t->policy->version;
    DFRL_TRY(launch_conv_table_rollout(t->ctx, a, f->tbl_logits, f->tbl_Dw, f->tbl_Dh));
  } else if (f->policy_conv && f->pnet.d1 == 128)
    DFRL_TRY((launch_conv_rollout<128, 64>(t->ctx, a, ctas)));
  else if (f->policy_conv)
    DFRL_TRY((launch_conv_rollout<64, 32>(t->ctx, a, ctas)));
  else if (f->pnet.d1 == 64)
    DFRL_TRY((launch_rollout<32, 64, 64, 8>(t->ctx, a, ctas)));
  else
    DFRL_TRY((launch_rollout<32, 16, 16, 8>(t->ctx, a, ctas)));
  t->obs_valid = false;
  return DFRL_OK;
}
