// layers.cu -- K2/K6 as per-layer kernels: the xylo::layer interface (reference nn.h:20-33) on
// the device, and xylo::model (nn.h:467-542) composed from them ("layered path").
//
// These kernels serve (a) the layer-level C ABI the C++ mirror of xylo/nn.h binds to, (b) nets
// too wide for the fused small-MLP kernels of fused.cu (C5: 128->256->256->256->32), and (c) the
// GPU-side cross-check of the fused kernels.  Plain FP32 FFMA (1e-4 parity needs more than
// TF32); shared-memory tiled, register micro-tiles, float4 traffic.
//
//   gemm_nn : C[M x N] = A[M x K] . B[K x N] (+ bias) (relu) (mask)   forward uses B = W^T
//                                                                      (cached transposed copy),
//                                                                      backward uses B = W
//   gemm_tn : dW[N x K] = dY^T . X, db = sum dY, SUM over rows (nn.h:94-98), deterministic
//             two-stage reduction (per-CTA partials in fixed order, then a fixed-order sum)
#include <math.h>

#include "common.cuh"

namespace {

// ------------------------------------------------------------------------- gemm_nn ----------
// 256 threads. Thread (tx, ty): tx = tid % (BN/4) owns 4 consecutive columns, ty = tid / (BN/4)
// owns TM rows {ty + RG * i}, RG = 256 / (BN/4) row groups (interleaved rows: the two row
// groups of a warp read adjacent shared-memory rows -> conflict-free float4 broadcasts).
template <int BM, int BN, int BK>
struct gemm_cfg {
  static constexpr int CG = BN / 4;        // column groups
  static constexpr int RG = 256 / CG;      // row groups
  static constexpr int TM = BM / RG;       // rows per thread
  static constexpr int LDA = BK + 4;       // padded A tile row
  static constexpr int SMEM = (BM * LDA + BK * BN) * 4;
};

template <int BM, int BN, int BK>
__global__ void __launch_bounds__(256)
gemm_nn_kernel(const float *__restrict__ A, const float *__restrict__ Bm,
               const float *__restrict__ bias, const float *__restrict__ mask,
               float *__restrict__ C, int M, int N, int K, int relu, int vecA, int vecB) {
  using cfg = gemm_cfg<BM, BN, BK>;
  constexpr int CG = cfg::CG, RG = cfg::RG, TM = cfg::TM, LDA = cfg::LDA;
  extern __shared__ __align__(16) float smem[];
  float *As = smem;             // [BM][LDA]
  float *Bs = smem + BM * LDA;  // [BK][BN]
  const int tid = threadIdx.x;
  const int tx = tid % CG, ty = tid / CG;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  float acc[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    // ---- stage A tile [BM][BK]
    constexpr int A_F4 = BM * BK / 4;
#pragma unroll
    for (int f = tid; f < A_F4; f += 256) {
      int r = f / (BK / 4), q = f % (BK / 4);
      int gr = m0 + r, gk = k0 + q * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < M) {
        const float *src = A + (size_t)gr * K + gk;
        if (vecA && gk + 3 < K) {
          v = *reinterpret_cast<const float4 *>(src);
        } else {
          if (gk < K) v.x = src[0];
          if (gk + 1 < K) v.y = src[1];
          if (gk + 2 < K) v.z = src[2];
          if (gk + 3 < K) v.w = src[3];
        }
      }
      *reinterpret_cast<float4 *>(As + r * LDA + q * 4) = v;
    }
    // ---- stage B tile [BK][BN]
    constexpr int B_F4 = BK * BN / 4;
#pragma unroll
    for (int f = tid; f < B_F4; f += 256) {
      int r = f / (BN / 4), q = f % (BN / 4);
      int gk = k0 + r, gn = n0 + q * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gk < K) {
        const float *src = Bm + (size_t)gk * N + gn;
        if (vecB && gn + 3 < N) {
          v = *reinterpret_cast<const float4 *>(src);
        } else {
          if (gn < N) v.x = src[0];
          if (gn + 1 < N) v.y = src[1];
          if (gn + 2 < N) v.z = src[2];
          if (gn + 3 < N) v.w = src[3];
        }
      }
      *reinterpret_cast<float4 *>(Bs + r * BN + q * 4) = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      float4 a[TM];
#pragma unroll
      for (int i = 0; i < TM; ++i)
        a[i] = *reinterpret_cast<const float4 *>(As + (ty + RG * i) * LDA + kk);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4 b = *reinterpret_cast<const float4 *>(Bs + (kk + j) * BN + tx * 4);
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          float av = j == 0 ? a[i].x : j == 1 ? a[i].y : j == 2 ? a[i].z : a[i].w;
          acc[i][0] = fmaf(av, b.x, acc[i][0]);
          acc[i][1] = fmaf(av, b.y, acc[i][1]);
          acc[i][2] = fmaf(av, b.z, acc[i][2]);
          acc[i][3] = fmaf(av, b.w, acc[i][3]);
        }
      }
    }
    __syncthreads();
  }
  // ---- epilogue
  const int gn = n0 + tx * 4;
  float bv[4] = {0.f, 0.f, 0.f, 0.f};
  if (bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (gn + j < N)
        bv[j] = bias[gn + j];
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int gr = m0 + ty + RG * i;
    if (gr >= M)
      continue;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = acc[i][j] + bv[j];
      if (relu)
        v = v > 0.f ? v : 0.f;
      o[j] = v;
    }
    float *dst = C + (size_t)gr * N + gn;
    if (vecB && gn + 3 < N) {
      if (mask) {
        float4 mk = *reinterpret_cast<const float4 *>(mask + (size_t)gr * N + gn);
        o[0] = mk.x > 0.f ? o[0] : 0.f;
        o[1] = mk.y > 0.f ? o[1] : 0.f;
        o[2] = mk.z > 0.f ? o[2] : 0.f;
        o[3] = mk.w > 0.f ? o[3] : 0.f;
      }
      *reinterpret_cast<float4 *>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (gn + j < N) {
          float v = o[j];
          if (mask && !(mask[(size_t)gr * N + gn + j] > 0.f))
            v = 0.f;
          dst[j] = v;
        }
    }
  }
}

template <int BM, int BN, int BK>
int launch_gemm_nn(dfrl_ctx *ctx, const float *A, const float *Bm, const float *bias,
                   const float *mask, float *C, int M, int N, int K, int relu) {
  using cfg = gemm_cfg<BM, BN, BK>;
  int vecA = (K % 4 == 0) && ((uintptr_t)A % 16 == 0);
  int vecB = (N % 4 == 0) && ((uintptr_t)Bm % 16 == 0) && ((uintptr_t)C % 16 == 0) &&
             (!mask || (uintptr_t)mask % 16 == 0);
  static bool attr_set = false;
  if (!attr_set) {
    DFRL_CUDA(cudaFuncSetAttribute(gemm_nn_kernel<BM, BN, BK>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, cfg::SMEM));
    attr_set = true;
  }
  dim3 grid(ceil_div(M, BM), ceil_div(N, BN));
  DFRL_LAUNCH(ctx, (gemm_nn_kernel<BM, BN, BK>), grid, 256, cfg::SMEM, A, Bm, bias, mask, C, M, N, K,
              relu, vecA, vecB);
  return DFRL_OK;
}

int gemm_nn(dfrl_ctx *ctx, const float *A, const float *Bm, const float *bias, const float *mask,
            float *C, int M, int N, int K, int relu) {
  if (M <= 0 || N <= 0)
    return DFRL_OK;
  {  // large, GEMM-shaped problems go to the tensor cores (gemm_umma.cu)
    int rc = umma_gemm_nn(ctx, A, Bm, bias, mask, C, M, N, K, relu);
    if (rc != DFRL_ERR_UNSUPPORTED)
      return rc;
  }
  if (K <= 8) {
    if (N > 32) return launch_gemm_nn<128, 64, 8>(ctx, A, Bm, bias, mask, C, M, N, K, relu);
    if (N > 8) return launch_gemm_nn<256, 32, 8>(ctx, A, Bm, bias, mask, C, M, N, K, relu);
    return launch_gemm_nn<256, 8, 8>(ctx, A, Bm, bias, mask, C, M, N, K, relu);
  }
  if (N > 32) return launch_gemm_nn<128, 64, 32>(ctx, A, Bm, bias, mask, C, M, N, K, relu);
  if (N > 16) return launch_gemm_nn<256, 32, 32>(ctx, A, Bm, bias, mask, C, M, N, K, relu);
  if (N > 8) return launch_gemm_nn<256, 16, 32>(ctx, A, Bm, bias, mask, C, M, N, K, relu);
  return launch_gemm_nn<256, 8, 32>(ctx, A, Bm, bias, mask, C, M, N, K, relu);
}

// ------------------------------------------------------------------------- gemm_tn ----------
// Partial dW over a chunk of rows. CTA output tile TN x TK (n x k), 8x8 register micro-tile,
// TPG = (TN/8)(TK/8) threads per group, G = 256 / TPG groups striding the rows; groups are
// summed in fixed order through shared memory, the CTA writes part[chunk][tile].
template <int TN, int TK>
__global__ void __launch_bounds__(256)
gemm_tn_kernel(const float *__restrict__ dY, const float *__restrict__ X,
               float *__restrict__ part, int M, int N, int K, int rows_per_cta, int vecY, int vecX) {
  constexpr int TPG = (TN / 8) * (TK / 8);
  constexpr int G = 256 / TPG;
  constexpr int RS = 32;                     // rows staged per step
  constexpr int LDY = TN + 4, LDX = TK + 4;
  extern __shared__ __align__(16) float smem[];
  float *Ys = smem;              // [RS][LDY]
  float *Xs = smem + RS * LDY;   // [RS][LDX]
  const int tid = threadIdx.x;
  const int g = tid / TPG, t = tid % TPG;
  const int tn = t % (TN / 8), tk = t / (TN / 8);
  const int n0 = blockIdx.y * TN, k0 = blockIdx.z * TK;
  const int r_begin = blockIdx.x * rows_per_cta;
  const int r_end = min(M, r_begin + rows_per_cta);

  float acc[8][8];
  float bsum[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    bsum[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      acc[i][j] = 0.f;
  }

  for (int r0 = r_begin; r0 < r_end; r0 += RS) {
    for (int f = tid; f < RS * TN / 4; f += 256) {
      int r = f / (TN / 4), q = f % (TN / 4);
      int gr = r0 + r, gn = n0 + q * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < r_end) {
        const float *src = dY + (size_t)gr * N + gn;
        if (vecY && gn + 3 < N)
          v = *reinterpret_cast<const float4 *>(src);
        else {
          if (gn < N) v.x = src[0];
          if (gn + 1 < N) v.y = src[1];
          if (gn + 2 < N) v.z = src[2];
          if (gn + 3 < N) v.w = src[3];
        }
      }
      *reinterpret_cast<float4 *>(Ys + r * LDY + q * 4) = v;
    }
    for (int f = tid; f < RS * TK / 4; f += 256) {
      int r = f / (TK / 4), q = f % (TK / 4);
      int gr = r0 + r, gk = k0 + q * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < r_end) {
        const float *src = X + (size_t)gr * K + gk;
        if (vecX && gk + 3 < K)
          v = *reinterpret_cast<const float4 *>(src);
        else {
          if (gk < K) v.x = src[0];
          if (gk + 1 < K) v.y = src[1];
          if (gk + 2 < K) v.z = src[2];
          if (gk + 3 < K) v.w = src[3];
        }
      }
      *reinterpret_cast<float4 *>(Xs + r * LDX + q * 4) = v;
    }
    __syncthreads();
#pragma unroll 2
    for (int r = g; r < RS; r += G) {
      float4 y0 = *reinterpret_cast<const float4 *>(Ys + r * LDY + tn * 8);
      float4 y1 = *reinterpret_cast<const float4 *>(Ys + r * LDY + tn * 8 + 4);
      float4 x0 = *reinterpret_cast<const float4 *>(Xs + r * LDX + tk * 8);
      float4 x1 = *reinterpret_cast<const float4 *>(Xs + r * LDX + tk * 8 + 4);
      float yv[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
      float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        bsum[i] += yv[i];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          acc[i][j] = fmaf(yv[i], xv[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
  // ---- fixed-order reduction over the G groups through shared memory (64 KB: [256][64])
  float *red = smem;  // reuse: needs 256 * 64 floats (+ 256 * 8 for bias)
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      red[(size_t)(i * 8 + j) * 256 + tid] = acc[i][j];
  float *redb = red + 64 * 256;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    redb[i * 256 + tid] = bsum[i];
  __syncthreads();
  // tile element (n, k) lives in thread-slot t = (k/8)*(TN/8) + n/8, register (n%8, k%8)
  const size_t tile_elems = (size_t)TN * TK + TN;
  float *dst = part + ((size_t)(blockIdx.x * gridDim.y + blockIdx.y) * gridDim.z + blockIdx.z) * tile_elems;
  for (int e = tid; e < TN * TK; e += 256) {
    int n = e / TK, k = e % TK;
    int slot = (k / 8) * (TN / 8) + n / 8, reg = (n % 8) * 8 + (k % 8);
    float s = 0.f;
    for (int gg = 0; gg < G; ++gg)
      s += red[(size_t)reg * 256 + gg * TPG + slot];
    dst[e] = s;
  }
  for (int n = tid; n < TN; n += 256) {
    // bias partial: threads with tk == 0 hold the row sums of their 8 columns
    int slot = n / 8, reg = n % 8;
    float s = 0.f;
    for (int gg = 0; gg < G; ++gg)
      s += redb[reg * 256 + gg * TPG + slot];
    dst[TN * TK + n] = s;
  }
}

// Second stage: grad[W n x k][b n] (+)= sum over chunks, fixed order.
__global__ void reduce_partials_kernel(const float *__restrict__ part, int chunks, int tiles_n,
                                       int tiles_k, int TN, int TK, int N, int K,
                                       float *__restrict__ grad, int accumulate) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  int total = N * K + N;
  if (e >= total)
    return;
  const size_t tile_elems = (size_t)TN * TK + TN;
  float s = 0.f;
  if (e < N * K) {
    int n = e / K, k = e % K;
    int tn = n / TN, tk = k / TK;
    size_t off = (size_t)(n % TN) * TK + (k % TK);
    for (int c = 0; c < chunks; ++c)
      s += part[((size_t)(c * tiles_n + tn) * tiles_k + tk) * tile_elems + off];
  } else {
    int n = e - N * K;
    int tn = n / TN;
    size_t off = (size_t)TN * TK + (n % TN);
    for (int c = 0; c < chunks; ++c)
      s += part[((size_t)(c * tiles_n + tn) * tiles_k + 0) * tile_elems + off];
  }
  grad[e] = accumulate ? grad[e] + s : s;
}

template <int TN, int TK>
int launch_gemm_tn(dfrl_ctx *ctx, const float *dY, const float *X, int M, int N, int K, float *grad,
                   int accumulate) {
  constexpr int SMEM = (64 * 256 + 8 * 256) * 4;  // reduction buffer dominates staging
  static bool attr_set = false;
  if (!attr_set) {
    DFRL_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<TN, TK>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  int tiles_n = ceil_div(N, TN), tiles_k = ceil_div(K, TK);
  int want_ctas = 2 * ctx->sm_count;
  int chunks = want_ctas / (tiles_n * tiles_k);
  if (chunks < 1) chunks = 1;
  int rows_per_cta = (int)round_up((size_t)ceil_div(M, chunks), 32);
  if (rows_per_cta < 256) rows_per_cta = 256;
  chunks = ceil_div(M, rows_per_cta);
  size_t tile_elems = (size_t)TN * TK + TN;
  void *part;
  DFRL_TRY(dfrl_scratch(ctx, sizeof(float) * tile_elems * chunks * tiles_n * tiles_k, &part));
  int vecY = (N % 4 == 0) && ((uintptr_t)dY % 16 == 0);
  int vecX = (K % 4 == 0) && ((uintptr_t)X % 16 == 0);
  dim3 grid(chunks, tiles_n, tiles_k);
  DFRL_LAUNCH(ctx, (gemm_tn_kernel<TN, TK>), grid, 256, SMEM, dY, X, (float *)part, M, N, K,
              rows_per_cta, vecY, vecX);
  int total = N * K + N;
  DFRL_LAUNCH(ctx, reduce_partials_kernel, ceil_div(total, 256), 256, 0, (const float *)part, chunks,
              tiles_n, tiles_k, TN, TK, N, K, grad, accumulate);
  return DFRL_OK;
}

int gemm_tn(dfrl_ctx *ctx, const float *dY, const float *X, int M, int N, int K, float *grad,
            int accumulate) {
  {
    int rc = umma_gemm_tn(ctx, dY, X, M, N, K, grad, accumulate);
    if (rc != DFRL_ERR_UNSUPPORTED)
      return rc;
  }
  // pick the smallest tile covering N x K (capped at 64 x 64)
  int tn = N > 32 ? 64 : N > 16 ? 32 : N > 8 ? 16 : 8;
  int tk = K > 32 ? 64 : K > 16 ? 32 : K > 8 ? 16 : 8;
#define TN_CASE(a, b) if (tn == a && tk == b) return launch_gemm_tn<a, b>(ctx, dY, X, M, N, K, grad, accumulate);
  TN_CASE(64, 64) TN_CASE(64, 32) TN_CASE(64, 16) TN_CASE(64, 8)
  TN_CASE(32, 64) TN_CASE(32, 32) TN_CASE(32, 16) TN_CASE(32, 8)
  TN_CASE(16, 64) TN_CASE(16, 32) TN_CASE(16, 16) TN_CASE(16, 8)
  TN_CASE(8, 64) TN_CASE(8, 32) TN_CASE(8, 16) TN_CASE(8, 8)
#undef TN_CASE
  dfrl_set_error("gemm_tn: no tile for %d x %d", N, K);
  return DFRL_ERR_INVALID;
}

// ------------------------------------------------------------------------- elementwise ------
__global__ void transpose_kernel(const float *__restrict__ W, int rows, int cols,
                                 float *__restrict__ Wt) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * cols)
    return;
  int r = e / cols, c = e % cols;
  Wt[(size_t)c * rows + r] = W[e];
}

__global__ void relu_fwd_kernel(const float *__restrict__ x, size_t n, float *__restrict__ y) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float v = x[i];
    y[i] = v > 0.f ? v : 0.f;
  }
}
__global__ void relu_bwd_kernel(const float *__restrict__ x, const float *__restrict__ dy, size_t n,
                                float *__restrict__ dx) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    dx[i] = x[i] > 0.f ? dy[i] : 0.f;
}
// softmax_layer::forward (nn.h:382-392): expf / sum, no max subtraction.
__global__ void softmax_fwd_kernel(const float *__restrict__ x, int rows, int cols,
                                   float *__restrict__ y) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows)
    return;
  const float *xi = x + (size_t)r * cols;
  float *yi = y + (size_t)r * cols;
  float s = 0.f;
  for (int c = 0; c < cols; ++c) {
    float e = expf(xi[c]);
    yi[c] = e;
    s += e;
  }
  for (int c = 0; c < cols; ++c)
    yi[c] = yi[c] / s;
}
// softmax_layer::backward (nn.h:393-417) from the probabilities s: dx_j = s_j (g_j - sum s_k g_k).
// Wide rows (cols > 8, e.g. 32 bins): 8 lanes per row, lane j owns columns j, j + 8, ... so that a
// warp reads 4 rows as contiguous 32-byte pieces; fixed-order xor-shuffle sums (deterministic).
__global__ void softmax_fwd_wide_kernel(const float *__restrict__ x, int rows, int cols,
                                        float *__restrict__ y) {
  long long g = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int j = threadIdx.x & 7;
  const bool ok = g < rows;
  const float *xi = x + (size_t)(ok ? g : 0) * cols;
  float s = 0.f;
  for (int c = j; c < cols && ok; c += 8)
    s += expf(xi[c]);
  for (int o = 4; o > 0; o >>= 1)
    s += __shfl_xor_sync(0xffffffffu, s, o);
  if (ok)
    for (int c = j; c < cols; c += 8)
      y[(size_t)g * cols + c] = expf(xi[c]) / s;
}
__global__ void softmax_bwd_probs_wide_kernel(const float *__restrict__ s, const float *__restrict__ dy,
                                              int rows, int cols, float *__restrict__ dx) {
  long long g = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int j = threadIdx.x & 7;
  const bool ok = g < rows;
  const float *si = s + (size_t)(ok ? g : 0) * cols, *gi = dy + (size_t)(ok ? g : 0) * cols;
  float dot = 0.f;
  for (int c = j; c < cols && ok; c += 8)
    dot = fmaf(si[c], gi[c], dot);
  for (int o = 4; o > 0; o >>= 1)
    dot += __shfl_xor_sync(0xffffffffu, dot, o);
  if (ok)
    for (int c = j; c < cols; c += 8)
      dx[(size_t)g * cols + c] = si[c] * (gi[c] - dot);
}
__global__ void softmax_bwd_probs_kernel(const float *__restrict__ s, const float *__restrict__ dy,
                                         int rows, int cols, float *__restrict__ dx) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows)
    return;
  const float *si = s + (size_t)r * cols;
  const float *gi = dy + (size_t)r * cols;
  float dot = 0.f;
  for (int c = 0; c < cols; ++c)
    dot = fmaf(si[c], gi[c], dot);
  for (int c = 0; c < cols; ++c)
    dx[(size_t)r * cols + c] = si[c] * (gi[c] - dot);
}
__global__ void softmax_bwd_kernel(const float *__restrict__ x, const float *__restrict__ dy,
                                   int rows, int cols, float *__restrict__ dx) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows)
    return;
  const float *xi = x + (size_t)r * cols;
  const float *gi = dy + (size_t)r * cols;
  float sum = 0.f;
  for (int c = 0; c < cols; ++c)
    sum += expf(xi[c]);
  float dot = 0.f;
  for (int c = 0; c < cols; ++c)
    dot = fmaf(expf(xi[c]) / sum, gi[c], dot);
  for (int c = 0; c < cols; ++c)
    dx[(size_t)r * cols + c] = expf(xi[c]) / sum * (gi[c] - dot);
}

// nn.h:12-18: N(0, stddev) weights via Box-Muller on Philox; biases zero.
__global__ void init_normal_kernel(float *__restrict__ p, int n, float stddev, uint64_t seed,
                                   uint32_t layer) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  philox4 r = philox4x32_10(seed, (uint64_t)i, layer, DFRL_STREAM_INIT);
  float u1 = ((float)(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
  float u2 = ((float)(r.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
  p[i] = stddev * sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
}

}  // namespace

// ------------------------------------------------------------------ layer-level C ABI -------
extern "C" int dfrl_dense_forward(dfrl_ctx *ctx, const float *params_dev, int in, int out,
                                  const float *x_dev, int rows, float *y_dev, int fuse_relu) {
  DFRL_CHECK(ctx && params_dev && x_dev && y_dev, "null argument");
  DFRL_CHECK(in > 0 && out > 0 && rows >= 0, "bad shape");
  // transposed copy of W in scratch (one-off; the model object caches it instead)
  void *wt;
  DFRL_TRY(dfrl_scratch(ctx, sizeof(float) * (size_t)in * out, &wt));
  DFRL_LAUNCH(ctx, transpose_kernel, ceil_div(in * out, 256), 256, 0, params_dev, out, in, (float *)wt);
  return gemm_nn(ctx, x_dev, (const float *)wt, params_dev + (size_t)in * out, nullptr, y_dev, rows,
                 out, in, fuse_relu);
}

extern "C" int dfrl_dense_backward(dfrl_ctx *ctx, const float *params_dev, int in, int out,
                                   const float *dy_dev, int rows, const float *relu_mask_dev,
                                   float *dx_dev) {
  DFRL_CHECK(ctx && params_dev && dy_dev && dx_dev, "null argument");
  DFRL_CHECK(in > 0 && out > 0 && rows >= 0, "bad shape");
  return gemm_nn(ctx, dy_dev, params_dev, nullptr, relu_mask_dev, dx_dev, rows, in, out, 0);
}

extern "C" int dfrl_dense_gradient(dfrl_ctx *ctx, int in, int out, const float *x_dev,
                                   const float *dy_dev, int rows, float *grad_dev, int accumulate) {
  DFRL_CHECK(ctx && x_dev && dy_dev && grad_dev, "null argument");
  DFRL_CHECK(in > 0 && out > 0 && rows > 0, "bad shape");
  return gemm_tn(ctx, dy_dev, x_dev, rows, out, in, grad_dev, accumulate);
}

extern "C" int dfrl_relu_forward(dfrl_ctx *ctx, const float *x_dev, size_t n, float *y_dev) {
  DFRL_CHECK(ctx && x_dev && y_dev, "null argument");
  if (n)
    DFRL_LAUNCH(ctx, relu_fwd_kernel, ceil_div((long long)n, 256), 256, 0, x_dev, n, y_dev);
  return DFRL_OK;
}
extern "C" int dfrl_relu_backward(dfrl_ctx *ctx, const float *x_dev, const float *dy_dev, size_t n,
                                  float *dx_dev) {
  DFRL_CHECK(ctx && x_dev && dy_dev && dx_dev, "null argument");
  if (n)
    DFRL_LAUNCH(ctx, relu_bwd_kernel, ceil_div((long long)n, 256), 256, 0, x_dev, dy_dev, n, dx_dev);
  return DFRL_OK;
}
extern "C" int dfrl_softmax_forward(dfrl_ctx *ctx, const float *x_dev, int rows, int cols,
                                    float *y_dev) {
  DFRL_CHECK(ctx && x_dev && y_dev, "null argument");
  DFRL_CHECK(rows >= 0 && cols > 0, "bad shape");
  if (rows)
    if (cols > 8)
      DFRL_LAUNCH(ctx, softmax_fwd_wide_kernel, ceil_div((long long)rows * 8, 256), 256, 0, x_dev, rows, cols, y_dev);
    else
      DFRL_LAUNCH(ctx, softmax_fwd_kernel, ceil_div(rows, 128), 128, 0, x_dev, rows, cols, y_dev);
  return DFRL_OK;
}
extern "C" int dfrl_softmax_backward(dfrl_ctx *ctx, const float *x_dev, const float *dy_dev,
                                     int rows, int cols, float *dx_dev) {
  DFRL_CHECK(ctx && x_dev && dy_dev && dx_dev, "null argument");
  DFRL_CHECK(rows >= 0 && cols > 0, "bad shape");
  if (rows)
    DFRL_LAUNCH(ctx, softmax_bwd_kernel, ceil_div(rows, 128), 128, 0, x_dev, dy_dev, rows, cols, dx_dev);
  return DFRL_OK;
}

// ------------------------------------------------------------------ xylo::model -------------
static bool is_param(int kind) { return kind == DFRL_LAYER_DENSE || kind == DFRL_LAYER_CONV1D_1; }

// Appends `n_layers` layers to m (incoming width `cols`), parameters laid out from *poff / *woff on.
static int append_layers(dfrl_mlp *m, int n_layers, const int *kinds, const int *ins, const int *outs, int cols,
                         size_t *poff, size_t *woff) {
  for (int l = 0; l < n_layers; ++l) {
    mlp_layer L;
    L.kind = kinds[l];
    L.in = ins[l];
    L.out = outs[l];
    L.in_cols = cols;
    L.points = 1;
    L.param_off = *poff;
    L.wt_off = *woff;
    if (L.kind == DFRL_LAYER_DENSE) {
      DFRL_CHECK(L.in == cols, "layer %d: dense input %d != incoming width %d", l, L.in, cols);
      L.out_cols = L.out;
    } else if (L.kind == DFRL_LAYER_CONV1D_1) {
      DFRL_CHECK(L.in > 0 && cols % L.in == 0, "layer %d: conv1d channels %d do not divide width %d", l, L.in, cols);
      L.points = cols / L.in;
      L.out_cols = L.points * L.out;
    } else if (L.kind == DFRL_LAYER_RELU || L.kind == DFRL_LAYER_SOFTMAX || L.kind == DFRL_LAYER_SOFTMAX_CE) {
      L.out_cols = cols;
    } else {
      dfrl_set_error("layer %d: unknown kind %d", l, L.kind);
      return DFRL_ERR_INVALID;
    }
    if (is_param(L.kind)) {
      DFRL_CHECK(L.out > 0, "layer %d: bad output size", l);
      *poff += (size_t)(L.in + 1) * L.out;
      *woff += (size_t)L.in * L.out;
    }
    cols = L.out_cols;
    m->layers.push_back(L);
  }
  m->output_cols = cols;
  return DFRL_OK;
}

static void init_mlp_fields(dfrl_mlp *m, dfrl_ctx *ctx, int input_cols) {
  m->ctx = ctx;
  m->input_cols = input_cols;
  m->params = nullptr;
  m->wt = nullptr;
  m->wt_dirty = true, m->version++;
  m->act_arena = nullptr;
  m->act_arena_bytes = 0;
  m->kept_rows = 0;
  m->kept_input = nullptr;
}

extern "C" int dfrl_mlp_create(dfrl_ctx *ctx, int n_layers, const int *kinds, const int *ins,
                               const int *outs, int input_cols, dfrl_mlp **out) {
  DFRL_CHECK(ctx && kinds && ins && outs && out, "null argument");
  DFRL_CHECK(n_layers > 0 && n_layers <= 64 && input_cols > 0, "bad layer count / input width");
  dfrl_mlp *m = new dfrl_mlp();
  init_mlp_fields(m, ctx, input_cols);
  size_t poff = 0, woff = 0;
  int rc = append_layers(m, n_layers, kinds, ins, outs, input_cols, &poff, &woff);
  if (rc != DFRL_OK) {
    delete m;
    return rc;
  }
  m->n_params = (int)poff;
  m->acts.assign(n_layers, nullptr);
  if (cudaMalloc(&m->params, sizeof(float) * (poff ? poff : 1)) != cudaSuccess ||
      cudaMemsetAsync(m->params, 0, sizeof(float) * (poff ? poff : 1), ctx->stream) != cudaSuccess ||
      cudaMalloc(&m->wt, sizeof(float) * (woff ? woff : 1)) != cudaSuccess) {
    dfrl_set_error("dfrl_mlp_create: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(m->params);
    cudaFree(m->wt);
    delete m;
    return DFRL_ERR_CUDA;
  }
  *out = m;
  return DFRL_OK;
}

// Shared-trunk model (BASELINE configs[2]: "shared-trunk policy/value MLP"). The reference's `model`
// is strictly sequential (nn.h:467-542) and its actor-critic mains build two separate nets
// (ac_training.cc:9-25); this is the extension SURVEY section 8d asks for: a second model whose first
// `n_shared` layers ARE the first layers of `trunk` -- same parameters, same memory -- followed by
// its own head layers. Both models address one flat parameter vector
//   [ trunk model's layers in the reference order | head layers of the sharer | ... ]
// (dfrl_mlp_param_count / get / set_params of either model see the whole vector; a model's flat
// gradient has zeros in the other heads' slots), so the learners' sequence "critic step, advantages
// with the UPDATED critic, actor step" (policy_gradient.h:159-185) moves the trunk twice per
// iteration, each time through the head whose loss is being optimised.
extern "C" int dfrl_mlp_create_shared(dfrl_mlp *trunk, int n_shared, int n_layers, const int *kinds, const int *ins,
                                      const int *outs, dfrl_mlp **out) {
  DFRL_CHECK(trunk && kinds && ins && outs && out, "null argument");
  dfrl_mlp *owner = trunk->share_owner ? trunk->share_owner : trunk;
  DFRL_CHECK(n_shared > 0 && n_shared <= (int)trunk->layers.size(), "n_shared must be in 1..%d", (int)trunk->layers.size());
  DFRL_CHECK(n_layers > 0 && n_shared + n_layers <= 64, "bad head layer count");
  dfrl_ctx *ctx = owner->ctx;
  dfrl_mlp *m = new dfrl_mlp();
  init_mlp_fields(m, ctx, trunk->input_cols);
  size_t woff = 0;
  for (int l = 0; l < n_shared; ++l) {  // the trunk's layers: same parameter offsets, own transposed-weight cache
    mlp_layer L = trunk->layers[l];
    L.wt_off = woff;
    if (is_param(L.kind))
      woff += (size_t)L.in * L.out;
    m->layers.push_back(L);
  }
  size_t poff = (size_t)owner->n_params;
  int rc = append_layers(m, n_layers, kinds, ins, outs, trunk->layers[n_shared - 1].out_cols, &poff, &woff);
  if (rc != DFRL_OK) {
    delete m;
    return rc;
  }
  // grow the family's flat vector by the head's parameters
  float *grown = nullptr;
  cudaStreamSynchronize(ctx->stream);
  if (cudaMalloc(&grown, sizeof(float) * poff) != cudaSuccess || cudaMalloc(&m->wt, sizeof(float) * (woff ? woff : 1)) != cudaSuccess ||
      cudaMemsetAsync(grown, 0, sizeof(float) * poff, ctx->stream) != cudaSuccess ||
      cudaMemcpyAsync(grown, owner->params, sizeof(float) * (size_t)owner->n_params, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess ||
      cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    dfrl_set_error("dfrl_mlp_create_shared: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(grown);
    cudaFree(m->wt);
    delete m;
    return DFRL_ERR_CUDA;
  }
  cudaFree(owner->params);
  m->share_owner = owner;
  m->n_shared_layers = n_shared;
  owner->sharers.push_back(m);
  owner->params = grown;
  owner->n_params = (int)poff;
  for (dfrl_mlp *s : owner->sharers) {
    s->params = grown;
    s->n_params = (int)poff;
  }
  m->acts.assign(m->layers.size(), nullptr);
  dfrl_mlp_params_changed(owner);
  *out = m;
  return DFRL_OK;
}

extern "C" int dfrl_mlp_destroy(dfrl_mlp *m) {
  if (!m)
    return DFRL_OK;
  DFRL_CHECK(m->sharers.empty(), "destroy the %d model(s) that share this model's trunk first", (int)m->sharers.size());
  cudaStreamSynchronize(m->ctx->stream);
  if (m->share_owner) {  // the family's parameter vector stays with its owner
    std::vector<dfrl_mlp *> &v = m->share_owner->sharers;
    for (size_t i = 0; i < v.size(); ++i)
      if (v[i] == m) {
        v.erase(v.begin() + i);
        break;
      }
  } else {
    cudaFree(m->params);
  }
  cudaFree(m->wt);
  if (m->act_arena)
    cudaFree(m->act_arena);
  delete m;
  return DFRL_OK;
}

extern "C" int dfrl_mlp_param_count(dfrl_mlp *m) { return m ? m->n_params : 0; }
extern "C" int dfrl_mlp_output_cols(dfrl_mlp *m) { return m ? m->output_cols : 0; }
extern "C" float *dfrl_mlp_params_dev(dfrl_mlp *m) {
  if (m)
    dfrl_mlp_params_changed(m);  // the caller may write through the pointer
  return m ? m->params : nullptr;
}

extern "C" int dfrl_mlp_set_params(dfrl_mlp *m, const float *params_host, int n) {
  DFRL_CHECK(m && params_host, "null argument");
  DFRL_CHECK(n == m->n_params, "parameter count %d != model's %d", n, m->n_params);
  DFRL_CUDA(cudaMemcpyAsync(m->params, params_host, sizeof(float) * n, cudaMemcpyHostToDevice,
                            m->ctx->stream));
  DFRL_CUDA(cudaStreamSynchronize(m->ctx->stream));
  dfrl_mlp_params_changed(m);
  return DFRL_OK;
}
extern "C" int dfrl_mlp_get_params(dfrl_mlp *m, float *params_host, int n) {
  DFRL_CHECK(m && params_host, "null argument");
  DFRL_CHECK(n == m->n_params, "parameter count %d != model's %d", n, m->n_params);
  DFRL_CUDA(cudaMemcpyAsync(params_host, m->params, sizeof(float) * n, cudaMemcpyDeviceToHost,
                            m->ctx->stream));
  DFRL_CUDA(cudaStreamSynchronize(m->ctx->stream));
  return DFRL_OK;
}

extern "C" int dfrl_mlp_init_params(dfrl_mlp *m, uint64_t seed) {
  DFRL_CHECK(m, "null model");
  // (a sharer initialises its own head layers only; the trunk belongs to the model it was built on)
  for (size_t l = (size_t)m->n_shared_layers; l < m->layers.size(); ++l) {
    const mlp_layer &L = m->layers[l];
    if (!is_param(L.kind))
      continue;
    DFRL_CUDA(cudaMemsetAsync(m->params + L.param_off, 0, sizeof(float) * (size_t)(L.in + 1) * L.out, m->ctx->stream));
    // normal_initialize: N(0, 0.01) regardless of fan-in (nn.h:12-14);
    // he_initialize: N(0, sqrt(2 / in_channels)) (nn.h:16-18)
    float sd = L.kind == DFRL_LAYER_DENSE ? 0.01f : sqrtf(2.0f / (float)L.in);
    int n = L.in * L.out;
    DFRL_LAUNCH(m->ctx, init_normal_kernel, ceil_div(n, 256), 256, 0, m->params + L.param_off, n, sd,
                seed, (uint32_t)l);
  }
  dfrl_mlp_params_changed(m);
  return DFRL_OK;
}

int dfrl_mlp_refresh_wt(dfrl_mlp *m) {
  if (!m->wt_dirty)
    return DFRL_OK;
  for (const mlp_layer &L : m->layers) {
    if (!is_param(L.kind))
      continue;
    DFRL_LAUNCH(m->ctx, transpose_kernel, ceil_div(L.in * L.out, 256), 256, 0,
                m->params + L.param_off, L.out, L.in, m->wt + L.wt_off);
  }
  m->wt_dirty = false;
  return DFRL_OK;
}

// model::forward (nn.h:481-488), keeping every activation for the reverse sweep. A relu that
// directly follows a parametric layer is fused into that layer's epilogue (the pre-activation is
// not materialised: relu's backward mask x > 0 equals relu(x) > 0).
int dfrl_mlp_forward_keep(dfrl_mlp *m, const float *x_dev, int rows, float **out_dev) {
  DFRL_CHECK(m && x_dev, "null argument");
  DFRL_CHECK(rows > 0, "rows must be positive");
  dfrl_ctx *ctx = m->ctx;
  DFRL_TRY(dfrl_mlp_refresh_wt(m));
  const int n = (int)m->layers.size();
  // arena: one buffer per materialised activation + two backprop ping-pong buffers
  size_t need = 0;
  int max_cols = m->input_cols;
  for (int l = 0; l < n; ++l) {
    bool fused = is_param(m->layers[l].kind) && l + 1 < n && m->layers[l + 1].kind == DFRL_LAYER_RELU;
    if (!fused)
      need += round_up((size_t)rows * m->layers[l].out_cols * sizeof(float), 256);
    if (m->layers[l].out_cols > max_cols)
      max_cols = m->layers[l].out_cols;
  }
  size_t pp = round_up((size_t)rows * max_cols * sizeof(float), 256);
  need += 2 * pp;
  if (need > m->act_arena_bytes) {
    DFRL_CUDA(cudaStreamSynchronize(ctx->stream));
    if (m->act_arena)
      DFRL_CUDA(cudaFree(m->act_arena));
    m->act_arena = nullptr;
    DFRL_CUDA(cudaMalloc(&m->act_arena, need));
    m->act_arena_bytes = need;
  }
  char *cur = (char *)m->act_arena + 2 * pp;
  const float *in = x_dev;
  bool prev_fused = false;
  for (int l = 0; l < n; ++l) {
    const mlp_layer &L = m->layers[l];
    if (prev_fused) {
      // this relu was applied in the epilogue of layer l-1; its output is already in acts[l]
      prev_fused = false;
      in = m->acts[l];
      continue;
    }
    bool fused = is_param(L.kind) && l + 1 < n && m->layers[l + 1].kind == DFRL_LAYER_RELU;
    float *dst = (float *)cur;
    cur += round_up((size_t)rows * L.out_cols * sizeof(float), 256);
    if (fused) {
      m->acts[l] = nullptr;  // pre-activation not materialised
      m->acts[l + 1] = dst;
      prev_fused = true;
    } else {
      m->acts[l] = dst;
    }
    if (is_param(L.kind)) {
      DFRL_TRY(gemm_nn(ctx, in, m->wt + L.wt_off, m->params + L.param_off + (size_t)L.in * L.out,
                       nullptr, dst, rows * L.points, L.out, L.in, fused ? 1 : 0));
    } else if (L.kind == DFRL_LAYER_RELU) {
      size_t cnt = (size_t)rows * L.out_cols;
      DFRL_LAUNCH(ctx, relu_fwd_kernel, ceil_div((long long)cnt, 256), 256, 0, in, cnt, dst);
    } else {
      if (L.out_cols > 8)
        DFRL_LAUNCH(ctx, softmax_fwd_wide_kernel, ceil_div((long long)rows * 8, 256), 256, 0, in, rows, L.out_cols, dst);
      else
        DFRL_LAUNCH(ctx, softmax_fwd_kernel, ceil_div(rows, 128), 128, 0, in, rows, L.out_cols, dst);
    }
    in = dst;
  }
  m->kept_rows = rows;
  m->kept_input = x_dev;
  if (out_dev)
    *out_dev = const_cast<float *>(in);
  return DFRL_OK;
}

// model::gradient (nn.h:510-528): reverse sweep; each layer's gradient lands in its slice of the
// flat vector; the first layer gets no dX.
int dfrl_mlp_backward(dfrl_mlp *m, const float *dy_dev, float *grad_dev) {
  DFRL_CHECK(m && dy_dev && grad_dev, "null argument");
  DFRL_CHECK(m->kept_rows > 0, "forward_keep must run first");
  dfrl_ctx *ctx = m->ctx;
  const int n = (int)m->layers.size();
  const int rows = m->kept_rows;
  if (m->share_owner || !m->sharers.empty())  // the other heads' slots of the flat gradient are zero
    DFRL_CUDA(cudaMemsetAsync(grad_dev, 0, sizeof(float) * (size_t)m->n_params, ctx->stream));
  int max_cols = m->input_cols;
  for (int l = 0; l < n; ++l)
    if (m->layers[l].out_cols > max_cols)
      max_cols = m->layers[l].out_cols;
  size_t pp = round_up((size_t)rows * max_cols * sizeof(float), 256);
  float *buf[2] = {(float *)m->act_arena, (float *)((char *)m->act_arena + pp)};
  int which = 0;
  const float *back = dy_dev;
  auto layer_input = [&](int l) -> const float * {
    for (int j = l - 1; j >= 0; --j)
      if (m->acts[j])
        return m->acts[j];
    return m->kept_input;
  };
  for (int l = n - 1; l >= 0; --l) {
    const mlp_layer &L = m->layers[l];
    if (is_param(L.kind)) {
      const float *in = layer_input(l);
      DFRL_TRY(gemm_tn(ctx, back, in, rows * L.points, L.out, L.in, grad_dev + L.param_off, 0));
      if (l == 0)
        break;
      // fuse the preceding relu's backward as an output mask
      const float *mask = nullptr;
      if (m->layers[l - 1].kind == DFRL_LAYER_RELU)
        mask = m->acts[l - 1];
      float *dst = buf[which];
      which ^= 1;
      DFRL_TRY(gemm_nn(ctx, back, m->params + L.param_off, nullptr, mask, dst, rows * L.points, L.in,
                       L.out, 0));
      back = dst;
      if (mask)
        --l;  // relu handled
    } else if (L.kind == DFRL_LAYER_RELU) {
      if (l == 0)
        break;
      float *dst = buf[which];
      which ^= 1;
      size_t cnt = (size_t)rows * L.out_cols;
      // mask by the relu OUTPUT (> 0 iff the input was > 0)
      DFRL_LAUNCH(ctx, relu_bwd_kernel, ceil_div((long long)cnt, 256), 256, 0, m->acts[l], back, cnt, dst);
      back = dst;
    } else if (L.kind == DFRL_LAYER_SOFTMAX) {
      if (l == 0)
        break;
      float *dst = buf[which];
      which ^= 1;
      if (L.out_cols > 8)
        DFRL_LAUNCH(ctx, softmax_bwd_probs_wide_kernel, ceil_div((long long)rows * 8, 256), 256, 0, m->acts[l], back,
                    rows, L.out_cols, dst);
      else
        DFRL_LAUNCH(ctx, softmax_bwd_probs_kernel, ceil_div(rows, 128), 128, 0, m->acts[l], back, rows,
                    L.out_cols, dst);
      back = dst;
    } else {
      // softmax_cross_entropy_layer::backward is the identity (nn.h:428-430)
    }
  }
  return DFRL_OK;
}

extern "C" int dfrl_mlp_eval(dfrl_mlp *m, const float *x_dev, int rows, float *y_dev) {
  DFRL_CHECK(m && x_dev && y_dev, "null argument");
  if (rows <= 0)
    return DFRL_OK;
  float *out = nullptr;
  DFRL_TRY(dfrl_mlp_forward_keep(m, x_dev, rows, &out));
  DFRL_CUDA(cudaMemcpyAsync(y_dev, out, sizeof(float) * (size_t)rows * m->output_cols,
                            cudaMemcpyDeviceToDevice, m->ctx->stream));
  return DFRL_OK;
}

extern "C" int dfrl_mlp_forward_gradient(dfrl_mlp *m, const float *x_dev, int rows,
                                         const float *dy_dev, float *grad_dev, float *out_dev) {
  DFRL_CHECK(m && x_dev && dy_dev && grad_dev, "null argument");
  DFRL_CHECK(rows > 0, "rows must be positive");
  float *out = nullptr;
  DFRL_TRY(dfrl_mlp_forward_keep(m, x_dev, rows, &out));
  if (out_dev)
    DFRL_CUDA(cudaMemcpyAsync(out_dev, out, sizeof(float) * (size_t)rows * m->output_cols,
                              cudaMemcpyDeviceToDevice, m->ctx->stream));
  return dfrl_mlp_backward(m, dy_dev, grad_dev);
}
