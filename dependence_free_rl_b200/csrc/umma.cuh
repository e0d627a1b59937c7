// umma.cuh -- hand-written tcgen05 / TMEM / mbarrier helpers for sm_100a (inline PTX).
//
// Operand convention used by every kernel here: a "panel" is a shared-memory tile of
// [rows][64 bf16] with a 128-byte row pitch, 1024-byte aligned, in the canonical
// SWIZZLE_128B layout (the 16-byte chunk index of an element is XORed with row % 8).
// The same physical panel can be fed to tcgen05.mma
//   * as a K-major operand:  rows = M/N index, the 64 columns = K (4 K-steps of 16), or
//   * as an MN-major operand: rows = K (contraction) index, the 64 columns = M/N index,
// which is what lets one copy of an activation tile serve the forward GEMM (X . W^T), the
// input-gradient GEMM (dY . W) and the weight-gradient GEMM (dY^T . X).
//
// FP32-level accuracy on the BF16 tensor pipe: every fp32 value x is split into
// hi = bf16(x), lo = bf16(x - hi) and a product is accumulated as hi*hi + hi*lo + lo*hi
// (fp32 accumulate in TMEM); the dropped lo*lo term is ~2^-18 relative.
#pragma once

#include <cuda_bf16.h>
#include <stdint.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- shared-memory matrix descriptor (SWIZZLE_128B, sm_100 version bits) ----
// bits [0,14) start address >> 4, [16,30) leading byte offset >> 4, [32,46) stride byte offset >> 4,
// [46,48) version = 1, [61,64) layout type = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t lbo_bytes,
                                                    uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// ---- instruction descriptor, kind::f16 with BF16 inputs and FP32 accumulation ----
// [4,6) c_format = 1 (F32), [7,10) a_format = 1 (BF16), [10,13) b_format = 1 (BF16),
// [15] a_major, [16] b_major (0 = K-major, 1 = MN-major), [17,23) N >> 3, [24,29) M >> 4.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Byte offset of bf16 element (row, col < 64) inside a SWIZZLE_128B panel.
__device__ __forceinline__ uint32_t panel_off(int row, int col) {
  return (uint32_t)(row * 128 + ((((col >> 3) ^ (row & 7)) & 7) << 4) + ((col & 7) << 1));
}
// Byte offset of the 16-byte chunk holding columns [8c, 8c+8) of `row`.
__device__ __forceinline__ uint32_t panel_chunk_off(int row, int chunk) {
  return (uint32_t)(row * 128 + (((chunk ^ (row & 7)) & 7) << 4));
}

constexpr uint32_t PANEL_ROW_BYTES = 128;
constexpr uint32_t PANEL_ATOM_BYTES = 1024;  // 8 rows
constexpr uint32_t KSTEP_BYTES_KMAJOR = 32;  // 16 bf16 along the row
constexpr uint32_t KSTEP_BYTES_MNMAJOR = 16 * 128;  // 16 rows

// ---- TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                   smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
}
__device__ __forceinline__ void fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}

// D[tmem] (+)= A[smem] . B[smem]; issued by ONE thread.
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]: A = [128 rows = lanes][K] packed 2 x bf16 per 32-bit column
// (low half = even k), a K = 16 step = 8 columns starting at a_taddr.
__device__ __forceinline__ void mma_bf16_ta(uint32_t tmem_d, uint32_t a_taddr, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(a_taddr), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// registers -> TMEM: thread (warp w, lane l) writes TMEM lane 32 (w % 4) + l, 8 consecutive columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}
// All previously issued MMAs of this thread arrive on the mbarrier when they complete.
__device__ __forceinline__ void commit(uint64_t *mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
                   smem_u32(mbar)) : "memory");
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t *mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *mbar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n" : "=r"(ok) : "r"(smem_u32(mbar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded spin: a tensor-core operation that never completes (malformed descriptor) must turn
// into a reported launch failure, never into a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *mbar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(mbar, parity)) {
    if (++spins > (1u << 26))
      __trap();
  }
}

// ---- TMEM -> registers: thread (warp w, lane l) reads TMEM lane 32 (w % 4) + l,
// 16 / 8 consecutive 32-bit columns starting at `taddr`'s column. ----
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i)
    v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// ---- fp32 -> (hi, lo) bf16 split ----
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16 &hi, __nv_bfloat16 &lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
// Two fp32 -> packed bf16x2 hi and lo words (low half = a): one cvt.rn.bf16x2.f32 per word.
__device__ __forceinline__ void split2(float a, float b, uint32_t &hi, uint32_t &lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  hi = *reinterpret_cast<uint32_t *>(&h);
  float ah = __uint_as_float(hi << 16), bh = __uint_as_float(hi & 0xffff0000u);
  __nv_bfloat162 l = __floats2bfloat162_rn(a - ah, b - bh);
  lo = *reinterpret_cast<uint32_t *>(&l);
}
// 8 consecutive fp32 -> one 16-byte chunk of hi and one of lo
__device__ __forceinline__ void split8(const float *x, uint4 &hi, uint4 &lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    split2(x[2 * i], x[2 * i + 1], h[i], l[i]);
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// One lane of a fully converged warp.
// Programmatic dependent launch (the fused kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization): launch_dependents lets the NEXT kernel of the
// stream be scheduled as soon as every CTA of this grid has issued it (its CTAs still need this grid's
// shared memory / tensor memory to be released, but the launch latency is off the critical path);
// wait blocks until the PREVIOUS grid has completed and its memory is visible. No-ops without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n" : "=r"(pred));
  return pred != 0;
}

}  // namespace umma
