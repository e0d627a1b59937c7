// umma_selftest.cu -- validates the tcgen05 building blocks of umma.cuh against a host fp64
// reference: one 128-row tile, bf16 hi/lo split operands in SWIZZLE_128B panels, three operand
// major-ness combinations (the three GEMMs of an MLP layer):
//   variant 0  D[128 x N]  = X[128 x K] . W[N x K]^T      A K-major,  B K-major   (forward)
//   variant 1  D[128 x Ko] = dY[128 x N] . W[N x Ko]      A K-major,  B MN-major  (input gradient)
//   variant 2  D[128 x N2] = G[128 x 128]^T . H[128 x N2] A MN-major, B MN-major  (weight gradient)
// and the "packed" operand used for the 16-column loss-gradient panel: hi in columns 0..15 and lo
// in columns 16..31 of ONE panel, addressed by descriptor start offsets of 0 and 32 bytes:
//   variant 3  D[128 x 16] = G[128 x 128]^T . Y[128 x 16]  B = packed Y, MN-major
//   variant 4  D[128 x N]  = Y[128 x 16] . W[16 x N]       A = packed Y, K-major
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "umma.cuh"

namespace {

constexpr int ROWS = 128;

// fp32 row-major [rows][cols] (global) -> hi/lo panels (64 columns each) in shared memory
__device__ void stage_panels(const float *__restrict__ src, int rows, int cols, uint8_t *hi, uint8_t *lo,
                             int panel_rows) {
  int npanels = (cols + 63) / 64;
  int chunks = rows * npanels * 8;
  for (int c = threadIdx.x; c < chunks; c += blockDim.x) {
    int row = c / (npanels * 8), rem = c % (npanels * 8);
    int panel = rem / 8, chunk = rem % 8;
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int col = panel * 64 + chunk * 8 + j;
      x[j] = col < cols ? src[(size_t)row * cols + col] : 0.f;
    }
    uint4 h, l;
    umma::split8(x, h, l);
    uint32_t off = panel * panel_rows * 128 + umma::panel_chunk_off(row, chunk);
    *reinterpret_cast<uint4 *>(hi + off) = h;
    *reinterpret_cast<uint4 *>(lo + off) = l;
  }
}

// fp32 [rows][16] -> ONE panel: hi in columns 0..15, lo in columns 16..31
__device__ void stage_packed16(const float *__restrict__ src, int rows, uint8_t *panel) {
  for (int c = threadIdx.x; c < rows * 2; c += blockDim.x) {
    int row = c >> 1, half = c & 1;
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      x[j] = src[(size_t)row * 16 + half * 8 + j];
    uint4 h, l;
    umma::split8(x, h, l);
    *reinterpret_cast<uint4 *>(panel + umma::panel_chunk_off(row, half)) = h;
    *reinterpret_cast<uint4 *>(panel + umma::panel_chunk_off(row, 2 + half)) = l;
  }
}

__global__ void __launch_bounds__(128)
umma_selftest_kernel(int variant, const float *__restrict__ A, int a_rows, int a_cols,
                     const float *__restrict__ Bm, int b_rows, int b_cols, float *__restrict__ D,
                     int n_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  // layout: A_hi [2 panels x 128 rows], A_lo, B_hi [2 panels x 128 rows], B_lo
  const uint32_t PANEL = 128 * 128;  // bytes of a 128-row panel
  uint8_t *a_hi = smem, *a_lo = smem + 2 * PANEL, *b_hi = smem + 4 * PANEL, *b_lo = smem + 6 * PANEL;
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x / 32;
  if (warp == 0)
    umma::tmem_alloc(&tmem_base_slot, 128);
  if (threadIdx.x == 0) {
    umma::mbar_init(&mbar, 1);
    umma::fence_mbar_init();
  }
  if (variant == 4) {
    stage_packed16(A, a_rows, a_hi);
    a_lo = a_hi + 32;
  } else {
    stage_panels(A, a_rows, a_cols, a_hi, a_lo, 128);
  }
  if (variant == 3) {
    stage_packed16(Bm, b_rows, b_hi);
    b_lo = b_hi + 32;
  } else {
    stage_panels(Bm, b_rows, b_cols, b_hi, b_lo, 128);
  }
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_slot;

  if (threadIdx.x == 0) {
    uint32_t idesc;
    int ksteps;
    if (variant == 0) {
      idesc = umma::make_idesc_bf16(128, n_out, 0, 0);
      ksteps = a_cols / 16;
    } else if (variant == 1 || variant == 4) {
      idesc = umma::make_idesc_bf16(128, n_out, 0, 1);
      ksteps = a_cols / 16;
    } else {
      idesc = umma::make_idesc_bf16(128, n_out, 1, 1);
      ksteps = a_rows / 16;
    }
    uint32_t acc = 0;
    for (int k = 0; k < ksteps; ++k) {
      uint32_t a_off, b_off;
      if (variant == 0) {  // both K-major: K runs along the panel row
        a_off = (k / 4) * PANEL + (k % 4) * umma::KSTEP_BYTES_KMAJOR;
        b_off = (k / 4) * PANEL + (k % 4) * umma::KSTEP_BYTES_KMAJOR;
      } else if (variant == 1 || variant == 4) {  // A K-major, B MN-major: B's K runs along panel rows
        a_off = (k / 4) * PANEL + (k % 4) * umma::KSTEP_BYTES_KMAJOR;
        b_off = k * umma::KSTEP_BYTES_MNMAJOR;
      } else {  // both MN-major
        a_off = k * umma::KSTEP_BYTES_MNMAJOR;
        b_off = k * umma::KSTEP_BYTES_MNMAJOR;
      }
      // LBO: K-major swizzled -> 16 B (unused); MN-major -> stride between 64-element MN blocks
      uint32_t a_lbo = (variant == 2 || variant == 3) ? PANEL : 16, b_lbo = variant == 0 ? 16 : PANEL;
      uint64_t ah = umma::make_desc_sw128(umma::smem_u32(a_hi) + a_off, a_lbo, 1024);
      uint64_t al = umma::make_desc_sw128(umma::smem_u32(a_lo) + a_off, a_lbo, 1024);
      uint64_t bh = umma::make_desc_sw128(umma::smem_u32(b_hi) + b_off, b_lbo, 1024);
      uint64_t bl = umma::make_desc_sw128(umma::smem_u32(b_lo) + b_off, b_lbo, 1024);
      umma::mma_bf16(tmem, ah, bh, idesc, acc);
      acc = 1;
      umma::mma_bf16(tmem, ah, bl, idesc, 1);
      umma::mma_bf16(tmem, al, bh, idesc, 1);
    }
    umma::commit(&mbar);
  }
  umma::mbar_wait(&mbar, 0);
  umma::fence_after_sync();
  const int row = warp * 32 + (threadIdx.x & 31);
  for (int c0 = 0; c0 < n_out; c0 += 16) {
    float v[16];
    umma::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    umma::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j)
      D[(size_t)row * n_out + c0 + j] = v[j];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0)
    umma::tmem_dealloc(tmem, 128);
}

// A operand from TENSOR MEMORY: D[128 x N] = X[128 x K] . W[N x K]^T, X = hi + lo packed bf16 pairs
// written with tcgen05.st (hi in columns 128.., lo in columns 160..), W in shared-memory panels.
// b_mn = 1: D[128 x N] = X[128 x K] . W[K x N] with W as an MN-major operand (input gradient form).
__global__ void __launch_bounds__(128)
umma_tmem_a_selftest_kernel(const float *__restrict__ A, int k, const float *__restrict__ Bm, int n, int b_mn,
                            float *__restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t PANEL = 128 * 128;
  uint8_t *b_hi = smem, *b_lo = smem + PANEL;
  __shared__ uint64_t mbar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x / 32, lane = threadIdx.x & 31, row = warp * 32 + lane;
  if (warp == 0)
    umma::tmem_alloc(&slot, 256);
  if (threadIdx.x == 0) {
    umma::mbar_init(&mbar, 1);
    umma::fence_mbar_init();
  }
  for (uint32_t o = threadIdx.x * 16; o < 2 * PANEL; o += blockDim.x * 16)
    *reinterpret_cast<uint4 *>(smem + o) = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if (b_mn)
    stage_panels(Bm, k, n, b_hi, b_lo, 128);  // rows = K, columns = N
  else
    stage_panels(Bm, n, k, b_hi, b_lo, 128);  // rows = N, columns = K
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = slot, lane_base = (uint32_t)(warp * 32) << 16;
  // this thread's row of X -> packed hi / lo words -> TMEM
  for (int c = 0; c < k; c += 16) {
    float x[16];
#pragma unroll
    for (int j = 0; j < 16; ++j)
      x[j] = A[(size_t)row * k + c + j];
    uint4 h0, l0, h1, l1;
    umma::split8(x, h0, l0);
    umma::split8(x + 8, h1, l1);
    uint32_t hw[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
    uint32_t lw[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
    umma::tmem_st8(tmem + lane_base + 128 + c / 2, hw);
    umma::tmem_st8(tmem + lane_base + 160 + c / 2, lw);
  }
  umma::tmem_st_wait();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma::make_idesc_bf16(128, n, 0, b_mn);
    for (int q = 0; q < k / 16; ++q) {
      const uint32_t b_off = b_mn ? q * umma::KSTEP_BYTES_MNMAJOR : q * umma::KSTEP_BYTES_KMAJOR;
      const uint32_t b_lbo = b_mn ? PANEL : 16;
      uint64_t bh = umma::make_desc_sw128(umma::smem_u32(b_hi) + b_off, b_lbo, 1024);
      uint64_t bl = umma::make_desc_sw128(umma::smem_u32(b_lo) + b_off, b_lbo, 1024);
      umma::mma_bf16_ta(tmem, tmem + 128 + 8 * q, bh, idesc, q > 0);
      umma::mma_bf16_ta(tmem, tmem + 128 + 8 * q, bl, idesc, 1);
      umma::mma_bf16_ta(tmem, tmem + 160 + 8 * q, bh, idesc, 1);
    }
    umma::commit(&mbar);
  }
  umma::mbar_wait(&mbar, 0);
  umma::fence_after_sync();
  for (int c0 = 0; c0 < n; c0 += 16) {
    float v[16];
    umma::tmem_ld16(tmem + lane_base + c0, v);
    umma::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j)
      D[(size_t)row * n + c0 + j] = v[j];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0)
    umma::tmem_dealloc(tmem, 256);
}

// M = 64 accumulator + tcgen05.ld.16x256b: D[64 x N] = X[64 x K] . W[N x K]^T (both K-major).
// Assumed layouts (what this test pins): D row r lives in TMEM lane 32 (r / 16) + r % 16; a
// 16x256b.x1 load by warp w returns to thread t the 2 x 2 block rows {t / 4, t / 4 + 8} x columns
// {2 (t % 4), 2 (t % 4) + 1} of the 16 lanes x 8 columns it addresses.
__global__ void __launch_bounds__(128)
umma_m64_selftest_kernel(const float *__restrict__ A, int k, const float *__restrict__ Bm, int n, float *__restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t PANEL = 128 * 128;
  uint8_t *a_hi = smem, *a_lo = smem + 2 * PANEL, *b_hi = smem + 4 * PANEL, *b_lo = smem + 6 * PANEL;
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_slot;
  const int warp = threadIdx.x / 32, t = threadIdx.x & 31;
  if (warp == 0)
    umma::tmem_alloc(&tmem_base_slot, 128);
  if (threadIdx.x == 0) {
    umma::mbar_init(&mbar, 1);
    umma::fence_mbar_init();
  }
  stage_panels(A, 64, k, a_hi, a_lo, 128);
  stage_panels(Bm, n, k, b_hi, b_lo, 128);
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma::make_idesc_bf16(64, n, 0, 0);
    uint32_t acc = 0;
    for (int ks = 0; ks < k / 16; ++ks) {
      uint32_t off = (ks / 4) * PANEL + (ks % 4) * umma::KSTEP_BYTES_KMAJOR;
      uint64_t ah = umma::make_desc_sw128(umma::smem_u32(a_hi) + off, 16, 1024);
      uint64_t al = umma::make_desc_sw128(umma::smem_u32(a_lo) + off, 16, 1024);
      uint64_t bh = umma::make_desc_sw128(umma::smem_u32(b_hi) + off, 16, 1024);
      uint64_t bl = umma::make_desc_sw128(umma::smem_u32(b_lo) + off, 16, 1024);
      umma::mma_bf16(tmem, ah, bh, idesc, acc);
      acc = 1;
      umma::mma_bf16(tmem, ah, bl, idesc, 1);
      umma::mma_bf16(tmem, al, bh, idesc, 1);
    }
    umma::commit(&mbar);
  }
  umma::mbar_wait(&mbar, 0);
  umma::fence_after_sync();
  for (int c0 = 0; c0 < n; c0 += 8) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0) : "memory");
    umma::tmem_ld_wait();
    const int row = warp * 16 + t / 4, col = c0 + 2 * (t % 4);
    D[(size_t)row * n + col] = __uint_as_float(r[0]);
    D[(size_t)row * n + col + 1] = __uint_as_float(r[1]);
    D[(size_t)(row + 8) * n + col] = __uint_as_float(r[2]);
    D[(size_t)(row + 8) * n + col + 1] = __uint_as_float(r[3]);
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0)
    umma::tmem_dealloc(tmem, 128);
}

float lcg(uint32_t &s) {
  s = s * 1664525u + 1013904223u;
  return ((s >> 8) & 0xffff) / 32768.0f - 1.0f;
}

}  // namespace

// variant 0..2, see file header. k = contraction length (multiple of 16, <= 128), n = output
// columns (multiple of 16, <= 64). Returns max |D - ref| / max |ref| through *rel_err.
extern "C" int dfrl_umma_selftest(dfrl_ctx *ctx, int variant, int k, int n, float *rel_err) {
  DFRL_CHECK(ctx && rel_err, "null argument");
  if (variant == 6 || variant == 7) {  // A operand from tensor memory (6: B K-major, 7: B MN-major)
    DFRL_CHECK(k % 16 == 0 && k >= 16 && k <= 64 && n % 16 == 0 && n >= 16 && n <= 64, "bad k / n");
    const int b_mn = variant == 7;
    std::vector<float> A((size_t)128 * k), B((size_t)n * k), D((size_t)128 * n);
    uint32_t s = 4242u + k * 3 + n + variant;
    for (float &x : A) x = lcg(s) * 1.7f;
    for (float &x : B) x = lcg(s) * 0.9f;
    float *dA, *dB, *dD;
    DFRL_CUDA(cudaMalloc(&dA, A.size() * 4));
    DFRL_CUDA(cudaMalloc(&dB, B.size() * 4));
    DFRL_CUDA(cudaMalloc(&dD, D.size() * 4));
    DFRL_CUDA(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    DFRL_CUDA(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    DFRL_CUDA(cudaMemset(dD, 0, D.size() * 4));
    const int smem = 2 * 128 * 128 + 1024;
    DFRL_LAUNCH(ctx, umma_tmem_a_selftest_kernel, 1, 128, smem, dA, k, dB, n, b_mn, dD);
    DFRL_CUDA(cudaStreamSynchronize(ctx->stream));
    DFRL_CUDA(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    double max_ref = 0, max_err = 0;
    for (int i = 0; i < 128; ++i)
      for (int j = 0; j < n; ++j) {
        double ref = 0;
        for (int q = 0; q < k; ++q)  // B: [n][k] (K-major) or [k][n] (MN-major)
          ref += (double)A[(size_t)i * k + q] * (b_mn ? B[(size_t)q * n + j] : B[(size_t)j * k + q]);
        double err = fabs(ref - (double)D[(size_t)i * n + j]);
        if (fabs(ref) > max_ref) max_ref = fabs(ref);
        if (err > max_err) max_err = err;
      }
    *rel_err = (float)(max_err / (max_ref + 1e-30));
    return DFRL_OK;
  }
  if (variant == 5) {  // M = 64 accumulator read with 16x256b loads
    DFRL_CHECK(k % 16 == 0 && k >= 16 && k <= 128 && n % 16 == 0 && n >= 16 && n <= 64, "bad k / n");
    std::vector<float> A((size_t)64 * k), B((size_t)n * k), D((size_t)64 * n);
    uint32_t s = 999u + k * 3 + n;
    for (float &x : A) x = lcg(s) * 1.7f;
    for (float &x : B) x = lcg(s) * 0.9f;
    float *dA, *dB, *dD;
    DFRL_CUDA(cudaMalloc(&dA, A.size() * 4));
    DFRL_CUDA(cudaMalloc(&dB, B.size() * 4));
    DFRL_CUDA(cudaMalloc(&dD, D.size() * 4));
    DFRL_CUDA(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    DFRL_CUDA(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    DFRL_CUDA(cudaMemset(dD, 0, D.size() * 4));
    const int smem = 8 * 128 * 128 + 1024;
    DFRL_CUDA(cudaFuncSetAttribute(umma_m64_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    DFRL_LAUNCH(ctx, umma_m64_selftest_kernel, 1, 128, smem, dA, k, dB, n, dD);
    DFRL_CUDA(cudaStreamSynchronize(ctx->stream));
    DFRL_CUDA(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    double max_ref = 0, max_err = 0;
    for (int i = 0; i < 64; ++i)
      for (int j = 0; j < n; ++j) {
        double ref = 0;
        for (int q = 0; q < k; ++q) ref += (double)A[(size_t)i * k + q] * B[(size_t)j * k + q];
        double err = fabs(ref - (double)D[(size_t)i * n + j]);
        if (fabs(ref) > max_ref) max_ref = fabs(ref);
        if (err > max_err) max_err = err;
      }
    *rel_err = (float)(max_err / (max_ref + 1e-30));
    return DFRL_OK;
  }
  DFRL_CHECK(variant >= 0 && variant <= 4, "variant 0..4");
  DFRL_CHECK(k % 16 == 0 && k >= 16 && k <= 128 && n % 16 == 0 && n >= 16 && n <= 64, "bad k / n");
  DFRL_CHECK((variant != 2 && variant != 3) || k == 128, "variants 2, 3 contract over the 128 tile rows");
  DFRL_CHECK(variant != 3 || n == 16, "variant 3 has 16 output columns");
  DFRL_CHECK(variant != 4 || k == 16, "variant 4 contracts over 16 columns");
  int a_rows, a_cols, b_rows, b_cols;
  if (variant == 0) { a_rows = ROWS; a_cols = k; b_rows = n; b_cols = k; }
  else if (variant == 1 || variant == 4) { a_rows = ROWS; a_cols = k; b_rows = k; b_cols = n; }
  else { a_rows = ROWS; a_cols = 128; b_rows = ROWS; b_cols = n; }
  std::vector<float> A((size_t)a_rows * a_cols), B((size_t)b_rows * b_cols), D((size_t)ROWS * n);
  uint32_t s = 12345u + variant * 77 + k * 3 + n;
  for (float &x : A) x = lcg(s) * 1.7f;
  for (float &x : B) x = lcg(s) * 0.9f;
  float *dA, *dB, *dD;
  DFRL_CUDA(cudaMalloc(&dA, A.size() * 4));
  DFRL_CUDA(cudaMalloc(&dB, B.size() * 4));
  DFRL_CUDA(cudaMalloc(&dD, D.size() * 4));
  DFRL_CUDA(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  DFRL_CUDA(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  DFRL_CUDA(cudaMemset(dD, 0, D.size() * 4));
  const int smem = 8 * 128 * 128 + 1024;
  DFRL_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  DFRL_LAUNCH(ctx, umma_selftest_kernel, 1, 128, smem, variant, dA, a_rows, a_cols, dB, b_rows, b_cols, dD, n);
  DFRL_CUDA(cudaStreamSynchronize(ctx->stream));
  DFRL_CUDA(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  double max_ref = 0, max_err = 0;
  for (int i = 0; i < ROWS; ++i)
    for (int j = 0; j < n; ++j) {
      double ref = 0;
      if (variant == 0)
        for (int q = 0; q < k; ++q) ref += (double)A[(size_t)i * k + q] * B[(size_t)j * k + q];
      else if (variant == 1 || variant == 4)
        for (int q = 0; q < k; ++q) ref += (double)A[(size_t)i * k + q] * B[(size_t)q * n + j];
      else
        for (int r = 0; r < ROWS; ++r) ref += (double)A[(size_t)r * 128 + i] * B[(size_t)r * n + j];
      double err = fabs(ref - (double)D[(size_t)i * n + j]);
      if (fabs(ref) > max_ref) max_ref = fabs(ref);
      if (err > max_err) max_err = err;
    }
  *rel_err = (float)(max_err / (max_ref + 1e-30));
  return DFRL_OK;
}

// ---------------------------------------------------------------------------------------------
// Pipe cost of one tcgen05.mma (kind::f16, bf16, K = 16) per operand form: `count` instructions
// issued back to back by one lane over zeroed panels, one commit, SM cycles from the first issue
// to the mbarrier flip. Two counts give the per-instruction slope.
namespace {
__global__ void __launch_bounds__(128) umma_microbench_kernel(int M, int N, int a_mn, int b_mn, int count,
                                                              long long *out) {
  extern __shared__ __align__(1024) uint8_t mb_raw[];
  uint8_t *smem = mb_raw + ((1024u - (umma::smem_u32(mb_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (uint32_t o = threadIdx.x * 16; o < 8 * 16384; o += blockDim.x * 16)
    *reinterpret_cast<uint4 *>(smem + o) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x < 32)
    umma::tmem_alloc(&slot, 256);
  if (threadIdx.x == 0) {
    umma::mbar_init(&bar, 1);
    umma::fence_mbar_init();
  }
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = slot, sb = umma::smem_u32(smem);
  const uint32_t idesc = umma::make_idesc_bf16(M, N, a_mn == 1, b_mn);
  long long t0 = 0;
  if (threadIdx.x < 32 && umma::elect_one()) {
    // A: panels 0..3 (MN-major M = 128 spans two panels, LBO = 16 KB), B: panels 4..7
    const uint32_t a_lbo = a_mn ? 16384u : 16u, b_lbo = b_mn ? 16384u : 16u;
    const uint32_t a_step = a_mn ? 2048u : 32u, b_step = b_mn ? 2048u : 32u;
    t0 = clock64();
    for (int i = 0; i < count; ++i) {
      const uint32_t k = (uint32_t)(i & 3);
      uint64_t ad = umma::make_desc_sw128(sb + k * a_step, a_lbo, 1024);
      uint64_t bd = umma::make_desc_sw128(sb + 4 * 16384 + k * b_step, b_lbo, 1024);
      if (a_mn == 2)  // A operand from tensor memory (columns 192.., whatever they hold)
        umma::mma_bf16_ta(tmem, tmem + 192 + 8 * k, bd, idesc, i > 0);
      else
        umma::mma_bf16(tmem, ad, bd, idesc, i > 0);
    }
    umma::commit(&bar);
    out[1] = clock64() - t0;  // issue time
  }
  umma::mbar_wait(&bar, 0);
  if (threadIdx.x < 32 && t0 != 0)
    out[0] = clock64() - t0;
  umma::fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32)
    umma::tmem_dealloc(tmem, 256);
}
}  // namespace

extern "C" int dfrl_umma_microbench(dfrl_ctx *ctx, int M, int N, int a_mn, int b_mn, int count, long long *cycles2) {
  DFRL_CHECK(ctx && cycles2 && (M == 64 || M == 128) && N >= 8 && N <= 256 && N % 8 == 0 && count > 0, "bad argument");
  long long *d;
  DFRL_CUDA(cudaMalloc(&d, 16));
  DFRL_CUDA(cudaMemset(d, 0, 16));
  const int smem = 8 * 16384 + 1024;
  DFRL_CUDA(cudaFuncSetAttribute(umma_microbench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  DFRL_LAUNCH(ctx, umma_microbench_kernel, 1, 128, smem, M, N, a_mn, b_mn, count, d);
  DFRL_CUDA(cudaStreamSynchronize(ctx->stream));
  DFRL_CUDA(cudaMemcpy(cycles2, d, 16, cudaMemcpyDeviceToHost));
  cudaFree(d);
  return DFRL_OK;
}
