// gemm_umma.cu -- tcgen05 GEMMs of the layered path (any net the fused kernels do not cover, in
// particular the 256-wide C5 nets): the same three products per layer as layers.cu
//   nn : C[M x N] = A[M x K] . B[K x N] (+ bias) (relu) (mask)   matmul_layer::forward / backward
//                                                                (nn.h:72-83), B = W^T or W
//   tn : dW[N x K] = dY[M x N]^T . X[M x K], db[N] = sum dY      matmul_layer::gradient (nn.h:85-100)
// with FP32-grade results on the bf16 tensor pipe (hi/lo split operands, hi.hi + hi.lo + lo.hi,
// FP32 accumulation in TMEM; see umma.cuh).
//
// Both kernels are persistent and warp specialised:
//   nn : warps 0-7 epilogue (TMEM -> bias / relu / mask -> global), warps 8-15 convert the fp32 A
//        tile to bf16 hi/lo SWIZZLE_128B panels, warp 16 issues cp.async.bulk copies of the
//        pre-split B chunks and the tcgen05.mma instructions. Two shared-memory stages and two
//        TMEM accumulators: the epilogue of tile i runs under the main loop of tile i+1.
//   tn : warps 0-15 convert 32-row chunks of dY and X (both MN-major operands: the contraction
//        runs over rows) and keep the column sums of dY for db, warp 16 issues the MMAs; the
//        [N x K] accumulator stays in TMEM for the CTA's whole row range (up to 2 x 256 columns),
//        per-CTA partials are summed in a fixed order by a second kernel (deterministic).
#include <string.h>

#include "common.cuh"
#include "umma.cuh"

namespace {

constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO 1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint64_t mk_desc(uint32_t saddr, uint32_t lbo) {
  return ((uint64_t)DESC_HI << 32) | ((saddr >> 4) & 0x3FFFu) | ((lbo >> 4) << 16);
}
__host__ __device__ constexpr uint32_t mk_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(umma::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(umma::smem_u32(bar)), "r"(bytes)
               : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on the mbarrier (UBLKCP).
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   umma::smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(umma::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16v(uint32_t taddr, float *v) {
  float t[16];
  umma::tmem_ld16(taddr, reinterpret_cast<float(&)[16]>(t));
#pragma unroll
  for (int i = 0; i < 16; ++i)
    v[i] = t[i];
}
// 8 consecutive fp32 (two 16-byte global words, or zeros) -> hi / lo chunk
__device__ __forceinline__ void load8(const float *p, bool ok, float *x) {
  if (ok) {
    float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
    x[0] = a.x, x[1] = a.y, x[2] = a.z, x[3] = a.w, x[4] = b.x, x[5] = b.y, x[6] = b.z, x[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      x[j] = 0.f;
  }
}

// 17 warps: with one warp per scheduler every instruction latency is exposed; 8 epilogue + 8 loader
// (nn) or 16 loader (tn) warps + the MMA warp keep 4 warps per scheduler in flight.
constexpr int nn_threads = 544, tn_threads = 544;

// ------------------------------------------------------------------ B image (nn) -------------
// fp32 B[K][N] -> per 64-row K chunk: [hi panels][lo panels], panel = 64 N-columns x 64 k-rows
// (MN-major operand: k along the panel rows), zero padded. One 16-byte chunk per thread.
__global__ void umma_prep_b_kernel(const float *__restrict__ B, int K, int N, int n_chunks, int NP,
                                   uint8_t *__restrict__ image) {
  const int per_chunk = 64 * NP * 8;  // 16-byte chunks of one hi (or lo) half
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_chunks * per_chunk)
    return;
  int c = idx / per_chunk, rem = idx % per_chunk;
  int p = rem / (64 * 8), kr = (rem / 8) % 64, ch = rem % 8;
  int k = c * 64 + kr, n0 = p * 64 + ch * 8;
  float x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
    x[j] = (k < K && n0 + j < N) ? B[(size_t)k * N + n0 + j] : 0.f;
  uint4 h, l;
  umma::split8(x, h, l);
  const size_t half = (size_t)NP * 8192;
  uint8_t *base = image + (size_t)c * 2 * half + (size_t)p * 8192 + umma::panel_chunk_off(kr, ch);
  *reinterpret_cast<uint4 *>(base) = h;
  *reinterpret_cast<uint4 *>(base + half) = l;
}

// ------------------------------------------------------------------ nn -----------------------
struct nn_args {
  const float *A, *bias, *mask;
  const uint8_t *image;  // pre-split B
  float *C;
  int M, N, K, relu, n_tiles, n_chunks, NP, acc_cols;
};

__global__ void __launch_bounds__(nn_threads, 1) umma_gemm_nn_kernel(nn_args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t b_half = (uint32_t)a.NP * 8192, stage_bytes = 2 * 16384 + 2 * b_half;
  // stage s: A_hi, A_lo (128 rows x 128 B each), B_hi, B_lo (NP panels of 64 rows each)
  uint8_t *tail = smem + 2 * stage_bytes;
  float *bias_s = reinterpret_cast<float *>(tail);
  uint64_t *bars = reinterpret_cast<uint64_t *>(tail + 1024);
  float *tpose = reinterpret_cast<float *>(tail + 1280);  // 8 warps x [32][20] floats = 20 KB
  uint64_t *a_full = bars, *b_full = bars + 2, *st_free = bars + 4, *acc_full = bars + 6, *acc_free = bars + 8;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 10);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  if (warp == 16)
    umma::tmem_alloc(tmem_slot, 2 * a.acc_cols);
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      umma::mbar_init(&a_full[s], 256);
      umma::mbar_init(&b_full[s], 1);
      umma::mbar_init(&st_free[s], 1);
      umma::mbar_init(&acc_full[s], 1);
      umma::mbar_init(&acc_free[s], 256);
    }
    umma::fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    bias_s[i] = (a.bias && i < a.N) ? a.bias[i] : 0.f;
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot, sbase = umma::smem_u32(smem);
  const uint32_t idesc = mk_idesc(128, a.N, 0, 1);
  const int my_tiles = ((int)blockIdx.x < a.n_tiles) ? (a.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 16) {
    // ---------------- B copies + MMA issue: one elected lane
    if (umma::elect_one()) {
      const uint32_t chunk_bytes = 2 * b_half;
      auto copy_b = [&](int g) {  // g = CTA-local chunk counter
        int s = g & 1, c = g % a.n_chunks;
        mbar_expect_tx(&b_full[s], chunk_bytes);
        bulk_g2s(smem + s * stage_bytes + 32768, a.image + (size_t)c * chunk_bytes, chunk_bytes, &b_full[s]);
      };
      const int total = my_tiles * a.n_chunks;
      if (total > 0)
        copy_b(0);
      int g = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int acc = i & 1;
        umma::mbar_wait(&acc_free[acc], ((i >> 1) & 1) ^ 1);
        umma::fence_after_sync();
        const uint32_t d = tmem + acc * a.acc_cols;
        for (int c = 0; c < a.n_chunks; ++c, ++g) {
          const int s = g & 1;
          const uint32_t par = (g >> 1) & 1;
          umma::mbar_wait(&a_full[s], par);
          umma::mbar_wait(&b_full[s], par);
          umma::fence_after_sync();
          const uint32_t st = sbase + s * stage_bytes;
          const int ks_n = min(4, (a.K - c * 64 + 15) / 16);
          for (int ks = 0; ks < ks_n; ++ks) {
            uint64_t ah = mk_desc(st + ks * 32, 16), al = mk_desc(st + 16384 + ks * 32, 16);
            uint64_t bh = mk_desc(st + 32768 + ks * 2048, 8192), bl = mk_desc(st + 32768 + b_half + ks * 2048, 8192);
            umma::mma_bf16(d, ah, bh, idesc, (c > 0 || ks > 0) ? 1u : 0u);
            umma::mma_bf16(d, ah, bl, idesc, 1);
            umma::mma_bf16(d, al, bh, idesc, 1);
          }
          umma::commit(&st_free[s]);
          if (g + 1 < total) {  // prefetch the next B chunk into the other stage once its MMAs are done
            umma::mbar_wait(&st_free[s ^ 1], (((g + 1) >> 1) & 1) ^ 1);
            copy_b(g + 1);
          }
        }
        umma::commit(&acc_full[acc]);
      }
    }
    __syncwarp();
  } else if (warp >= 8) {
    // ---------------- A loaders: fp32 [128 rows][64 k] -> hi / lo panels, 4 tasks of 8 floats per
    // thread and chunk; the loads of chunk g + 1 are in flight while chunk g is split and stored
    const int t = threadIdx.x - 256;  // 0..255
    const bool vec = (a.K % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.A) & 15) == 0);
    const int total = my_tiles * a.n_chunks;
    auto load_chunk = [&](int g, float (&x)[4][8]) {
      int i = g / a.n_chunks, c = g % a.n_chunks;
      const int row0 = ((int)blockIdx.x + i * (int)gridDim.x) * 128;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int task = t + 256 * q, r = task >> 3, ch = task & 7;
        int gr = row0 + r, k0 = c * 64 + ch * 8;
        const float *src = a.A + (size_t)gr * a.K + k0;
        if (gr < a.M && k0 + 8 <= a.K && vec) {
          load8(src, true, x[q]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            x[q][j] = (gr < a.M && k0 + j < a.K) ? src[j] : 0.f;
        }
      }
    };
    auto store_chunk = [&](int g, float (&x)[4][8]) {
      const int s = g & 1;
      umma::mbar_wait(&st_free[s], ((g >> 1) & 1) ^ 1);
      uint8_t *st = smem + s * stage_bytes;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int task = t + 256 * q, r = task >> 3, ch = task & 7;
        uint4 h, l;
        umma::split8(x[q], h, l);
        uint32_t off = umma::panel_chunk_off(r, ch);
        *reinterpret_cast<uint4 *>(st + off) = h;
        *reinterpret_cast<uint4 *>(st + 16384 + off) = l;
      }
      umma::fence_proxy_async();
      mbar_arrive(&a_full[s]);
    };
    float xa[4][8], xb[4][8];
    if (total > 0)
      load_chunk(0, xa);
    for (int g = 0; g < total; g += 2) {
      if (g + 1 < total)
        load_chunk(g + 1, xb);
      store_chunk(g, xa);
      if (g + 2 < total)
        load_chunk(g + 2, xa);
      if (g + 1 < total)
        store_chunk(g + 1, xb);
    }
  } else {
    // ---------------- epilogue: TMEM -> (+bias, relu, mask) -> C
    // 8 warps: warp e reads TMEM lanes 32 (e % 4).. and one half of the 16-column groups. A thread
    // owns one row of the accumulator; global traffic goes through a per-warp [32 rows][16 cols]
    // shared-memory block so that 4 lanes cover 64 contiguous bytes of a row.
    const int q4 = warp & 3, part = warp >> 2;
    const int groups = a.N / 16, gsplit = (groups + 1) / 2;
    const int g_begin = part == 0 ? 0 : gsplit, g_end = part == 0 ? gsplit : groups;
    float *tb = tpose + warp * (32 * 20);
    const int lr = lane >> 2, lc = (lane & 3) * 4;  // coalesced view: rows lr + 8 i, columns lc..lc+3
    for (int i = 0; i < my_tiles; ++i) {
      const int acc = i & 1;
      const int row0 = ((int)blockIdx.x + i * (int)gridDim.x) * 128 + q4 * 32;
      umma::mbar_wait(&acc_full[acc], (i >> 1) & 1);
      umma::fence_after_sync();
      const uint32_t d = tmem + acc * a.acc_cols + ((uint32_t)(q4 * 32) << 16);
      for (int gi = g_begin; gi < g_end; ++gi) {
        const int n0 = gi * 16;
        float v[16], mk[16];
        tmem_ld16v(d + n0, v);
        if (a.mask) {  // coalesced read of the [32 x 16] mask block, then every thread takes its row
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            int gr = row0 + lr + 8 * q;
            float4 m = gr < a.M ? __ldg(reinterpret_cast<const float4 *>(a.mask + (size_t)gr * a.N + n0 + lc))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4 *>(tb + (lr + 8 * q) * 20 + lc) = m;
          }
          __syncwarp();
#pragma unroll
          for (int j4 = 0; j4 < 16; j4 += 4) {
            float4 m = *reinterpret_cast<const float4 *>(tb + lane * 20 + j4);
            mk[j4] = m.x, mk[j4 + 1] = m.y, mk[j4 + 2] = m.z, mk[j4 + 3] = m.w;
          }
          __syncwarp();
        }
        umma::tmem_ld_wait();
#pragma unroll
        for (int j4 = 0; j4 < 16; j4 += 4) {
          float o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float y = v[j4 + j] + bias_s[n0 + j4 + j];
            if (a.relu)
              y = y > 0.f ? y : 0.f;
            if (a.mask)
              y = mk[j4 + j] > 0.f ? y : 0.f;
            o[j] = y;
          }
          *reinterpret_cast<float4 *>(tb + lane * 20 + j4) = make_float4(o[0], o[1], o[2], o[3]);
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          int gr = row0 + lr + 8 * q;
          if (gr < a.M)
            *reinterpret_cast<float4 *>(a.C + (size_t)gr * a.N + n0 + lc) =
                *reinterpret_cast<const float4 *>(tb + (lr + 8 * q) * 20 + lc);
        }
        __syncwarp();
      }
      umma::fence_before_sync();
      mbar_arrive(&acc_free[acc]);
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 16)
    umma::tmem_dealloc(tmem, 2 * a.acc_cols);
}

// ------------------------------------------------------------------ tn -----------------------
struct tn_args {
  const float *dY, *X;
  float *part;  // [gridDim.x][N * K + N]
  int M, N, K, rows_per_cta, NPA, NPB, n_stages, kcols, tmem_cols, m_mma, sh_n, sh_k;
};

__global__ void __launch_bounds__(tn_threads, 1) umma_gemm_tn_kernel(tn_args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
  // stage: dY hi panels (NPA x 4 KB), dY lo, X hi panels (NPB x 4 KB), X lo; a panel = 32 rows x 128 B
  const uint32_t a_half = (uint32_t)a.NPA * 4096, b_half = (uint32_t)a.NPB * 4096, stage_bytes = 2 * (a_half + b_half);
  uint8_t *tail = smem + (size_t)a.n_stages * stage_bytes;
  float *dbred = reinterpret_cast<float *>(tail);  // [512 threads][8]
  uint64_t *bars = reinterpret_cast<uint64_t *>(tail + 16384);
  uint64_t *full = bars, *st_free = bars + 4, *done = bars + 8;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 9);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (warp == 16)
    umma::tmem_alloc(tmem_slot, a.tmem_cols);
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.n_stages; ++s) {
      umma::mbar_init(&full[s], 512);
      umma::mbar_init(&st_free[s], 1);
    }
    umma::mbar_init(done, 1);
    umma::fence_mbar_init();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot, sbase = umma::smem_u32(smem);
  const int r0 = (int)blockIdx.x * a.rows_per_cta, r1 = min(a.M, r0 + a.rows_per_cta);
  const int n_chunks = r1 > r0 ? (r1 - r0 + 31) / 32 : 0;
  const int NB = (a.N + 127) / 128;  // 128-row blocks of the [N x K] output

  if (warp == 16) {
    if (umma::elect_one()) {
      const uint32_t idesc = mk_idesc(a.m_mma, a.kcols, 1, 1);
      for (int g = 0; g < n_chunks; ++g) {
        const int s = g % a.n_stages;
        umma::mbar_wait(&full[s], (g / a.n_stages) & 1);
        umma::fence_after_sync();
        const uint32_t st = sbase + s * stage_bytes;
        for (int mb = 0; mb < NB; ++mb) {
          const uint32_t d = tmem + mb * a.kcols;
          for (int ks = 0; ks < 2; ++ks) {
            uint64_t ah = mk_desc(st + mb * 8192 + ks * 2048, 4096), al = mk_desc(st + a_half + mb * 8192 + ks * 2048, 4096);
            uint64_t bh = mk_desc(st + 2 * a_half + ks * 2048, 4096), bl = mk_desc(st + 2 * a_half + b_half + ks * 2048, 4096);
            umma::mma_bf16(d, ah, bh, idesc, (g > 0 || ks > 0) ? 1u : 0u);
            umma::mma_bf16(d, ah, bl, idesc, 1);
            umma::mma_bf16(d, al, bh, idesc, 1);
          }
        }
        umma::commit(&st_free[s]);
      }
      umma::commit(done);
    }
    __syncwarp();
  } else {
    // ---------------- loaders: 32-row chunks of dY [32 x N] and X [32 x K] -> MN-major panels.
    // <= 4 tasks of 8 floats per thread and chunk (the first ta tasks are dY, the rest X); the
    // loads of chunk g + 1 are in flight while chunk g is split and stored.
    const int t = threadIdx.x;  // 0..511
    const int sa = a.sh_n, sb = a.sh_k;       // log2 of the 8-column groups per row (N / 8, K / 8)
    const int ta = 32 << sa, tb = 32 << sb;   // tasks per chunk
    const bool veca = (reinterpret_cast<uintptr_t>(a.dY) & 15) == 0, vecb = (reinterpret_cast<uintptr_t>(a.X) & 15) == 0;
    float colsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // of dY, column group t % (N / 8)
    auto load_chunk = [&](int g, float (&x)[4][8]) {
      const int row0 = r0 + g * 32;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int task = t + 512 * q;
        if (task < ta) {
          int r = task >> sa, cg = task & ((1 << sa) - 1), gr = row0 + r;
          load8(a.dY + (size_t)gr * a.N + cg * 8, gr < r1 && veca, x[q]);
          if (gr < r1 && !veca)
            for (int j = 0; j < 8; ++j)
              x[q][j] = a.dY[(size_t)gr * a.N + cg * 8 + j];
        } else if (task - ta < tb) {
          int tk = task - ta, r = tk >> sb, cg = tk & ((1 << sb) - 1), gr = row0 + r;
          load8(a.X + (size_t)gr * a.K + cg * 8, gr < r1 && vecb, x[q]);
          if (gr < r1 && !vecb)
            for (int j = 0; j < 8; ++j)
              x[q][j] = a.X[(size_t)gr * a.K + cg * 8 + j];
        }
      }
    };
    auto store_chunk = [&](int g, float (&x)[4][8]) {
      const int s = g % a.n_stages;
      umma::mbar_wait(&st_free[s], ((g / a.n_stages) & 1) ^ 1);
      uint8_t *st = smem + s * stage_bytes;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int task = t + 512 * q;
        uint4 h, l;
        if (task < ta) {
          int r = task >> sa, cg = task & ((1 << sa) - 1);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            colsum[j] += x[q][j];
          umma::split8(x[q], h, l);
          uint32_t off = (cg >> 3) * 4096 + umma::panel_chunk_off(r, cg & 7);
          *reinterpret_cast<uint4 *>(st + off) = h;
          *reinterpret_cast<uint4 *>(st + a_half + off) = l;
        } else if (task - ta < tb) {
          int tk = task - ta, r = tk >> sb, cg = tk & ((1 << sb) - 1);
          umma::split8(x[q], h, l);
          uint32_t off = (cg >> 3) * 4096 + umma::panel_chunk_off(r, cg & 7);
          *reinterpret_cast<uint4 *>(st + 2 * a_half + off) = h;
          *reinterpret_cast<uint4 *>(st + 2 * a_half + b_half + off) = l;
        }
      }
      umma::fence_proxy_async();
      mbar_arrive(&full[s]);
    };
    float xa[4][8], xb[4][8];
    if (n_chunks > 0)
      load_chunk(0, xa);
    for (int g = 0; g < n_chunks; g += 2) {
      if (g + 1 < n_chunks)
        load_chunk(g + 1, xb);
      store_chunk(g, xa);
      if (g + 2 < n_chunks)
        load_chunk(g + 2, xa);
      if (g + 1 < n_chunks)
        store_chunk(g + 1, xb);
    }
    // ---------------- drain: db (fixed-order sum over the threads sharing a column group), dW
    float *part = a.part + (size_t)blockIdx.x * ((size_t)a.N * a.K + a.N);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      dbred[t * 8 + j] = colsum[j];
    asm volatile("bar.sync 1, 512;\n" ::: "memory");
    if (t < a.N) {
      int cg = t >> 3, j = t & 7;
      float sum = 0.f;
      for (int u = cg; u < 512; u += 1 << sa)  // the threads whose dY tasks all belong to column group cg
        sum += dbred[u * 8 + j];
      part[(size_t)a.N * a.K + t] = sum;
    }
    if (n_chunks > 0) {
      umma::mbar_wait(done, 0);
      umma::fence_after_sync();
    }
    // 4 warps per TMEM lane quarter, each takes every fourth 16-column group
    const int q = warp & 3, cpart = warp >> 2;
    for (int mb = 0; mb < NB; ++mb) {
      int nrow;  // output row (dY column) held by this thread's TMEM lane
      if (a.m_mma == 128)
        nrow = mb * 128 + q * 32 + lane;
      else
        nrow = lane < 16 ? q * 16 + lane : -1;  // M = 64: row r in lane 32 (r / 16) + r % 16
      for (int c0 = cpart * 16; c0 < a.kcols; c0 += 64) {
        float v[16];
        if (n_chunks > 0) {
          tmem_ld16v(tmem + ((uint32_t)(q * 32) << 16) + mb * a.kcols + c0, v);
          umma::tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            v[j] = 0.f;
        }
        if (nrow >= 0 && nrow < a.N)
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c0 + j < a.K)
              part[(size_t)nrow * a.K + c0 + j] = v[j];
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 16)
    umma::tmem_dealloc(tmem, a.tmem_cols);
}

__global__ void umma_reduce_kernel(const float *__restrict__ part, int ctas, int n, float *__restrict__ grad,
                                   int accumulate) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  float s = 0.f;
  for (int c = 0; c < ctas; ++c)
    s += part[(size_t)c * n + i];
  grad[i] = accumulate ? grad[i] + s : s;
}

int pow2_cols(int c) {
  int p = 32;
  while (p < c)
    p <<= 1;
  return p;
}

}  // namespace

// C = A . B (+bias)(relu)(mask). DFRL_ERR_UNSUPPORTED when the shape is better served by the FFMA kernels.
int umma_gemm_nn(dfrl_ctx *ctx, const float *A, const float *Bm, const float *bias, const float *mask, float *C,
                 int M, int N, int K, int relu) {
  if (M < 4096 || N % 16 != 0 || N < 16 || N > 256 || K % 16 != 0 || K < 32 || K > 4096)
    return DFRL_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(C) & 15) != 0 || (mask && (reinterpret_cast<uintptr_t>(mask) & 15) != 0))
    return DFRL_ERR_UNSUPPORTED;
  nn_args a;
  a.A = A;
  a.bias = bias;
  a.mask = mask;
  a.C = C;
  a.M = M;
  a.N = N;
  a.K = K;
  a.relu = relu;
  a.n_tiles = ceil_div(M, 128);
  a.n_chunks = ceil_div(K, 64);
  a.NP = ceil_div(N, 64);
  a.acc_cols = pow2_cols(N);
  const size_t image_bytes = (size_t)a.n_chunks * 2 * a.NP * 8192;
  void *ws = nullptr;
  DFRL_TRY(dfrl_umma_workspace(ctx, image_bytes, &ws));
  a.image = (const uint8_t *)ws;
  const int prep_chunks = a.n_chunks * 64 * a.NP * 8;
  DFRL_LAUNCH(ctx, umma_prep_b_kernel, ceil_div(prep_chunks, 256), 256, 0, Bm, K, N, a.n_chunks, a.NP, (uint8_t *)ws);
  const int smem = 2 * (32768 + 2 * a.NP * 8192) + 1024 + 256 + 8 * 32 * 20 * 4 + 1024;
  static int smem_set = 0;
  if (smem > smem_set) {
    DFRL_CUDA(cudaFuncSetAttribute(umma_gemm_nn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    smem_set = smem;
  }
  const int grid = a.n_tiles < ctx->sm_count ? a.n_tiles : ctx->sm_count;
  DFRL_LAUNCH(ctx, umma_gemm_nn_kernel, grid, nn_threads, smem, a);
  return DFRL_OK;
}

// grad = [dW = dY^T . X (N x K)][db = sum dY (N)], SUM over rows (nn.h:94-98).
int umma_gemm_tn(dfrl_ctx *ctx, const float *dY, const float *X, int M, int N, int K, float *grad, int accumulate) {
  const bool n_ok = N == 16 || N == 32 || N == 64 || N == 128 || N == 256;
  const bool k_ok = K == 16 || K == 32 || K == 64 || K == 128 || K == 256;
  if (M < 4096 || !n_ok || !k_ok)
    return DFRL_ERR_UNSUPPORTED;
  tn_args a;
  a.dY = dY;
  a.X = X;
  a.M = M;
  a.N = N;
  a.K = K;
  a.NPA = ceil_div(N, 64);
  a.NPB = ceil_div(K, 64);
  if (N > 64 && a.NPA % 2)  // a 128-row output block spans two 64-column panels
    a.NPA += 1;
  a.kcols = K;  // MMA N dimension (multiple of 16)
  a.sh_n = a.sh_k = 0;
  while ((8 << a.sh_n) < N) ++a.sh_n;
  while ((8 << a.sh_k) < K) ++a.sh_k;
  a.m_mma = N <= 64 ? 64 : 128;
  const int NB = ceil_div(N, 128);
  a.tmem_cols = pow2_cols(NB * a.kcols);
  const int stage_bytes = 2 * (a.NPA + a.NPB) * 4096;
  a.n_stages = (192 * 1024) / stage_bytes;
  if (a.n_stages > 4)
    a.n_stages = 4;
  const int grid = ctx->sm_count;
  int rows = ceil_div(M, grid);
  a.rows_per_cta = (rows + 31) / 32 * 32;
  const size_t per = (size_t)N * K + N;
  void *ws = nullptr;
  DFRL_TRY(dfrl_umma_workspace(ctx, sizeof(float) * per * grid, &ws));
  a.part = (float *)ws;
  const int smem = a.n_stages * stage_bytes + 16384 + 256 + 1024;
  static int smem_set = 0;
  if (smem > smem_set) {
    DFRL_CUDA(cudaFuncSetAttribute(umma_gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    smem_set = smem;
  }
  DFRL_LAUNCH(ctx, umma_gemm_tn_kernel, grid, tn_threads, smem, a);
  DFRL_LAUNCH(ctx, umma_reduce_kernel, ceil_div((int)per, 256), 256, 0, (const float *)ws, grid, (int)per, grad, accumulate);
  return DFRL_OK;
}
