// fused_conv.cuh -- fused tcgen05 kernels for the reference's conv1d policy nets; included by fused.cu
// inside its anonymous namespace (shares its panel / TMEM / gradient-tail helpers).
//
// Net: convolution1d_1_layer(4 -> D1) relu convolution1d_1_layer(D1 -> D2) relu
//      convolution1d_1_layer(D2 -> 1) softmax over the B = 8 points
// (ppo_training.cc:10-16: 4-128-64-1; ac_training.cc:9-16: 4-64-32-1; nn.h:113-194). A conv1d_1 layer
// is the dense layer applied to every (sample, bin) row [bin.w/8, bin.h/8, item.w/8, item.h/8]
// (bin_packing.h:31-40), so the net is a 4-D1-D2-1 MLP over 8 x as many rows whose 8 scalar outputs per
// sample are the logits. A tile = 128 rows = 16 samples x 8 bins, row = 8 * sample + bin: the 8 rows
// of a sample are 8 adjacent lanes of one warp (softmax / sampling through warp shuffles).
//
//   forward   H1 = relu(X W1^T + b1)        K = 16 (4 inputs | 1 | zeros), N = D1     2 MMAs (X exact in bf16)
//             H2 = relu(H1 W2^T + b2)       A = H1 from tensor memory, K = D1         3 MMAs per K step
//             logit = H2 . w3 + b3          fp32 registers
//   head      softmax over the sample's 8 lanes, loss gradient, softmax Jacobian -> dY (a scalar per row)
//   backward  dH2 = dY w3 . relu'           rank 1, registers; dW3 / db3 accumulate in registers
//             db2 += [hi(dH2); lo(dH2)]^T . [X | 1]   M = 128 (stacked), N = 16: column 4 (the ones column)
//             dH1 = (dH2 W2) . relu'        A = dH2 from tensor memory, B = W2 as MN-major operand, N = D1
//             dW2 += [hi(dH2); lo(dH2)]^T . [hi(H1) + lo(H1)]   M = 128 (stacked), N = D1: 2 MMAs per K step
//                                           carry all four hi / lo products
//             [dW1 | db1] += dH1^T . [X | 1]   M = D1, N = 16
// Weight-gradient sums stay in tensor memory for all tiles of the CTA (fixed order: bitwise
// reproducible), then gradient_tail (grid barrier, slice reduction, exchange, optimizer).

// Operand precision: FP16 pairs. The 8 logit gradients of a sample sum to zero, so every weight
// gradient of a conv1d policy depends only on how the activations DIFFER between a sample's bins
// (often by 10 % or less of their size: most bins of a young episode are identical): the absolute
// rounding error of an activation is amplified by that ratio, and the bf16 hi/lo pairs of the dense
// kernels (2^-17 relative) left a 16-sample reference trace 1.1e-4 off (golden `ppo_refnet_conv`).
// Here every MMA operand is split x = hi + lo with hi, lo in FP16 (11 + 11 significant bits, 2^-22
// relative) and a product is hi.hi + hi.lo + lo.hi (+ lo.lo where it is free: the stacked weight-gradient
// GEMMs; see c_conv_lolo). FP16's narrow exponent range is handled by
// power-of-two scales (exact): weights are staged as W * S_W with max|W * S_W| in [2^12, 2^13) and the
// epilogues multiply the accumulators by 1 / S_W; the gradient chain runs on dY * S_g with max|A| * S_g
// in [8, 16) over the CTA's own rows (the CTA's partial gradient is multiplied by 1 / S_g in the
// drain). Activations are used unscaled (|h| < 65504 assumed; an overflow would surface as inf / NaN
// gradients, not silently).
// Whether the A-from-TMEM GEMMs (layer 2, dH1) also issue the lo.lo product (2^-22 of a term; the stacked
// weight-gradient GEMMs carry it for free). OFF: measured on B200 (profiles/r02c_conv_lolo_ab.log) the
// fourth product costs 4.8 % of the policy step (982 -> 935 us at 131 072 envs) and 4 % of the rollout and
// changes no parity figure (conv policy gradients 0.5e-6 .. 1.4e-5 from the fp64 yardstick of the tests either way: the
// error floor is the fp32 accumulation order, not the operand split). DFRL_CONV_LOLO=1 turns it on.
__constant__ int c_conv_lolo = 0;

__device__ __forceinline__ void split2_h(float a, float b, uint32_t &hi, uint32_t &lo) {
  __half2 h = __floats2half2_rn(a, b);
  hi = *reinterpret_cast<uint32_t *>(&h);
  const float2 hf = __half22float2(h);
  __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  lo = *reinterpret_cast<uint32_t *>(&l);
}
__device__ __forceinline__ void split8_h(const float *x, uint4 &hi, uint4 &lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    split2_h(x[2 * i], x[2 * i + 1], h[i], l[i]);
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ uint32_t pack2_h(float a, float b) {  // two values that are exact in fp16
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&h);
}
// 0xffff in each half whose fp16 value is > 0
__device__ __forceinline__ uint32_t pos_mask2_h(uint32_t w) {
  uint32_t m;
  asm("set.gt.u32.f16x2 %0, %1, %2;\n" : "=r"(m) : "r"(w), "r"(0u));
  return m;
}
// Power-of-two scale S with max * S in [target / 2, target) (1 when max == 0 or not finite).
__device__ __forceinline__ float pow2_scale(float mx, int target_log2) {
  if (!(mx > 0.f) || !(mx < 3.0e38f))
    return 1.f;
  int e;
  frexpf(mx, &e);  // mx = m 2^e, m in [0.5, 1)
  return ldexpf(1.f, target_log2 - e);
}
// idesc with FP16 operands (a_format = b_format = 0), FP32 accumulation
template <int N> struct IDH {
  static constexpr uint32_t FK_FK = make_idesc(128, N, 0, 0, 0, 0);  // forward: A K-major, B K-major
  static constexpr uint32_t BK_FM = make_idesc(128, N, 0, 0, 0, 1);  // dX: A K-major (tensor memory), B MN-major
  static constexpr uint32_t BM_FM = make_idesc(128, N, 0, 0, 1, 1);  // dW: both MN-major
};
// Block-wide maximum of non-negative values (whole CTA; `red` = 32 floats of shared memory).
__device__ __forceinline__ float block_max(float v, float *red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0)
    red[threadIdx.x >> 5] = v;
  __syncthreads();
  float m = 0.f;
  for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w)
    m = fmaxf(m, red[w]);
  __syncthreads();
  return m;
}
// Hidden-layer epilogues for FP16-pair operands: as epi2_fwd / epi2_bwd with the weight scale undone.
template <int D, bool SMEM_STORE>
__device__ __forceinline__ void conv_epi_fwd(uint32_t acc, const tid_t &t, const float *__restrict__ bias, float inv_s, uint8_t *hi,
                                             uint8_t *lo, int h0, int h1) {
#pragma unroll
  for (int h = h0; h < h1; ++h) {
    float v[32];
    tmem_load<32>(acc + t.lane_base + h * 32, v);
#pragma unroll
    for (int j = 0; j < 32; ++j)
      v[j] = fmaxf(fmaf(v[j], inv_s, bias[h * 32 + j]), 0.f);
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      uint4 hh, ll;
      split8_h(&v[8 * cc], hh, ll);
      if (SMEM_STORE) {
        const int c = h * 4 + cc;
        const uint32_t off = (uint32_t)(c >> 3) * PANEL + umma::panel_chunk_off(t.row, c & 7);
        *reinterpret_cast<uint4 *>(hi + off) = hh;
        *reinterpret_cast<uint4 *>(lo + off) = ll;
      }
      tmem_put_chunk<32>(acc + t.lane_base + h * 32, cc, hh, ll);
    }
  }
  umma::tmem_st_wait();
}
template <int D>
__device__ __forceinline__ void conv_epi_bwd(uint32_t acc, const tid_t &t, float inv_s, const uint8_t *act_hi, uint8_t *hi, uint8_t *lo,
                                             int h0, int h1) {
#pragma unroll
  for (int h = h0; h < h1; ++h) {
    float v[32];
    tmem_load<32>(acc + t.lane_base + h * 32, v);
#pragma unroll
    for (int j = 0; j < 32; ++j)
      v[j] *= inv_s;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int c = h * 4 + cc;
      const uint32_t off = (uint32_t)(c >> 3) * PANEL + umma::panel_chunk_off(t.row, c & 7);
      const uint4 aw = *reinterpret_cast<const uint4 *>(act_hi + off);
      uint4 hh, ll;
      split8_h(&v[8 * cc], hh, ll);
      const uint32_t m0 = pos_mask2_h(aw.x), m1 = pos_mask2_h(aw.y), m2 = pos_mask2_h(aw.z), m3 = pos_mask2_h(aw.w);
      *reinterpret_cast<uint4 *>(hi + off) = make_uint4(hh.x & m0, hh.y & m1, hh.z & m2, hh.w & m3);
      *reinterpret_cast<uint4 *>(lo + off) = make_uint4(ll.x & m0, ll.y & m1, ll.z & m2, ll.w & m3);
    }
  }
}

template <int D1, int D2>
struct cvmap {
  static_assert((D1 == 128 && D2 == 64) || (D1 == 64 && D2 == 32), "conv policy widths: 4-128-64-1 or 4-64-32-1");
  static constexpr int NP1 = D1 / 64;  // 64-column panels of a D1-wide activation
  // [128 rows]: hi(W1) in bytes 0..31 of a row (K = 16: 4 inputs, zeros), lo(W1) in 32..63,
  // observation slot 0 in 64..95, slot 1 in 96..127
  static constexpr uint32_t WX = 0;
  static constexpr uint32_t W2SUB = D2 * 128;  // one 64-column sub-panel of W2: [D2 rows][64 columns]
  static constexpr uint32_t W2_HI = WX + PANEL, W2_LO = W2_HI + NP1 * W2SUB;
  static constexpr uint32_t FLOATS = W2_LO + NP1 * W2SUB;  // b1[D1] b2[D2] w3[D2] b3
  // ... b3, 1 / S_W1, 1 / S_W2, S_g, 1 / S_g; 32 floats of reduction scratch
  static constexpr int F_B1 = 0, F_B2 = D1, F_W3 = D1 + D2, F_B3 = D1 + 2 * D2, F_IS1 = F_B3 + 1, F_IS2 = F_B3 + 2, F_SG = F_B3 + 3,
                       F_ISG = F_B3 + 4, F_RED = F_B3 + 8, N_FLOATS = F_RED + 32;
  static constexpr uint32_t SCR = FLOATS + N_FLOATS * 4;  // partial logits [2][128]
  static constexpr uint32_t H1_HI = (SCR + 2 * TILE * 4 + 1023) / 1024 * 1024;
  static constexpr uint32_t H1_LO = H1_HI + NP1 * PANEL;
  static constexpr uint32_t G2_HI = H1_LO + NP1 * PANEL, G2_LO = G2_HI + PANEL;  // dH2 (adjacent: stacked M = 128)
  static constexpr uint32_t DH1_HI = G2_LO + PANEL, DH1_LO = DH1_HI + NP1 * PANEL;
  static constexpr uint32_t BARS = DH1_LO + NP1 * PANEL;
  static constexpr uint32_t TOTAL = BARS + 128;
  static constexpr int THREADS = 288;  // 2 epilogue threads per row (a 32-column chunk group each) + the MMA issuer warp
  // tensor memory columns
  static constexpr uint32_t ACC0 = 0, ACC1 = 128, DA = 192, DB = 320, DB2 = 336;
  static_assert(TOTAL + 1024 <= 232448, "exceeds the 227 KB shared memory of an SM");
};

struct conv_step_args {
  const float *params;  // flat fp32 parameters of the net
  net3 net;             // d0 = 4, d3 = 1
  const int8_t *rec_state;  // [T][2B + 2][stride]
  const uint8_t *rec_action;
  const float *adv;    // [T][n]
  const float *p_old;  // [T][n][8]
  int n, stride, T;
  float inv_w, inv_h;
  int n_tiles;  // ceil(T n / 16)
  int loss_kind, head_bwd;
  float *partials;
  grad_tail tail;
  long long *clk;  // optional: phase clocks of CTA 0 (debug): 7 stamps per tile, first 14 tiles
};

// W1 [D1][4] and W2 [D2][D1] (fp32, staged copy `P`) -> FP16-pair operand panels of W * S_W (S_W a power
// of two from max|W|, identical in every CTA); biases / head weights / 1 / S_W -> floats. Whole CTA.
template <int D1, int D2, typename CM>
__device__ void stage_conv_weights(const float *__restrict__ P, const net3 &net, uint8_t *smem, float *fl) {
  const float *W1 = P + net.o_w1, *W2 = P + net.o_w2;
  float m1 = 0.f, m2 = 0.f;
  for (int i = threadIdx.x; i < D1 * 4; i += blockDim.x)
    m1 = fmaxf(m1, fabsf(W1[i]));
  for (int i = threadIdx.x; i < D2 * D1; i += blockDim.x)
    m2 = fmaxf(m2, fabsf(W2[i]));
  const float s1 = pow2_scale(block_max(m1, fl + CM::F_RED), 13), s2 = pow2_scale(block_max(m2, fl + CM::F_RED), 13);
  for (int n = threadIdx.x; n < D1; n += blockDim.x) {
    float x[8] = {W1[n * 4] * s1, W1[n * 4 + 1] * s1, W1[n * 4 + 2] * s1, W1[n * 4 + 3] * s1, 0.f, 0.f, 0.f, 0.f};
    uint4 h, l;
    split8_h(x, h, l);
    *reinterpret_cast<uint4 *>(smem + CM::WX + umma::panel_chunk_off(n, 0)) = h;
    *reinterpret_cast<uint4 *>(smem + CM::WX + umma::panel_chunk_off(n, 2)) = l;
  }
  for (int c = threadIdx.x; c < D2 * (D1 / 8); c += blockDim.x) {
    const int n = c / (D1 / 8), ch = c % (D1 / 8);
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      x[j] = W2[(size_t)n * D1 + ch * 8 + j] * s2;
    uint4 h, l;
    split8_h(x, h, l);
    const uint32_t off = (uint32_t)(ch >> 3) * CM::W2SUB + umma::panel_chunk_off(n, ch & 7);
    *reinterpret_cast<uint4 *>(smem + CM::W2_HI + off) = h;
    *reinterpret_cast<uint4 *>(smem + CM::W2_LO + off) = l;
  }
  for (int i = threadIdx.x; i < D1; i += blockDim.x) fl[CM::F_B1 + i] = P[net.o_b1 + i];
  for (int i = threadIdx.x; i < D2; i += blockDim.x) fl[CM::F_B2 + i] = P[net.o_b2 + i];
  for (int i = threadIdx.x; i < D2; i += blockDim.x) fl[CM::F_W3 + i] = P[net.o_w3 + i];
  if (threadIdx.x == 0) {
    fl[CM::F_B3] = P[net.o_b3];
    fl[CM::F_IS1] = 1.f / s1;
    fl[CM::F_IS2] = 1.f / s2;
  }
}

// Layer 2: ACC1 = H1 (tensor memory, tmem_put_chunk layout in ACC0) . W2^T, B = W2 sub-panels K-major.
template <int D1, int D2, typename CM>
__device__ __forceinline__ void conv_issue_layer2(uint32_t tm, uint32_t sbase) {
#pragma unroll
  for (int k = 0; k < D1 / 16; ++k) {
    const uint32_t ah = tm + CM::ACC0 + (k / 2) * 32 + (k % 2) * 8, al = ah + 16;
    const uint32_t boff = (uint32_t)(k / 4) * CM::W2SUB + (uint32_t)(k % 4) * umma::KSTEP_BYTES_KMAJOR;
    const uint64_t bh = desc_lo_hi(desc_lo(sbase + CM::W2_HI + boff, 16)), bl = desc_lo_hi(desc_lo(sbase + CM::W2_LO + boff, 16));
    umma::mma_bf16_ta(tm + CM::ACC1, ah, bh, IDH<D2>::FK_FK, k > 0 ? 1u : 0u);
    umma::mma_bf16_ta(tm + CM::ACC1, ah, bl, IDH<D2>::FK_FK, 1);
    umma::mma_bf16_ta(tm + CM::ACC1, al, bh, IDH<D2>::FK_FK, 1);
    if (c_conv_lolo)
      umma::mma_bf16_ta(tm + CM::ACC1, al, bl, IDH<D2>::FK_FK, 1);  // lo.lo too: see the precision note in the header
  }
}

// Observation row [bin.w / cap_w, bin.h / cap_h, item.w / cap_w, item.h / cap_h, 1, 0, 0, 0] of one
// (sample, bin) row into observation slot `slot` of the WX panel (exact in fp16).
template <typename CM>
__device__ __forceinline__ void conv_encode_row(uint8_t *smem, int row, int slot, int bw, int bh, int iw, int ih, float inv_w,
                                                float inv_h) {
  *reinterpret_cast<uint4 *>(smem + CM::WX + umma::panel_chunk_off(row, 4 + 2 * slot)) =
      make_uint4(pack2_h((float)bw * inv_w, (float)bh * inv_h), pack2_h((float)iw * inv_w, (float)ih * inv_h), 0x00003C00u, 0u);
}

template <int D1, int D2, bool PROBE = false>  // PROBE: phase stamps of tools/conv_phase_clocks.py (a separate instantiation)
__global__ void __launch_bounds__((cvmap<D1, D2>::THREADS), 1) fused_conv_policy_step_kernel(conv_step_args a) {
  using CM = cvmap<D1, D2>;
  constexpr int NB = 8, P = 2 * NB + 2, SPT = TILE / NB;  // bins, state planes, samples per tile
  constexpr int RT = CM::THREADS;                           // threads of an operands-ready hand-over
  constexpr int NC1 = D1 / 32, NC2 = D2 / 32;               // 32-column chunks of the hidden layers
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
  float *fl = reinterpret_cast<float *>(smem + CM::FLOATS);
  float *scr = reinterpret_cast<float *>(smem + CM::SCR);
  // mbarriers: [0] MMA completion on the chain, [1] dW2 GEMM done (H1, dH2 panels free), [2] dW1 GEMM
  // done (dH1 panels, observation slot free); [8] parameter staging
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + CM::BARS);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + CM::BARS + 96);
  const net3 net = a.net;
  const tid_t t = thread_id();
  const bool issuer = t.warp >= 8;  // warp-uniform
  const int half = issuer ? 0 : (t.warp >> 2) & 1;
  const uint32_t sbase = umma::smem_u32(smem);

  umma::pdl_launch_dependents();
  if (t.warp == 0)
    umma::tmem_alloc(tmem_slot, 512);
  if (threadIdx.x == 0) {
    for (int q = 0; q < 3; ++q)
      umma::mbar_init(bars + q, 1);
    umma::fence_mbar_init();
  }
  // zeros that no epilogue writes: the K padding of W1 / X in the WX panel, the columns >= D2 of dH2
  zero_bytes(smem + CM::WX, PANEL);
  zero_bytes(smem + CM::G2_HI, 2 * PANEL);
  __syncthreads();
  stage_conv_weights<D1, D2, CM>(stage_params_bulk(a.params, net.n_params, smem + CM::H1_HI, bars + 8), net, smem, fl);
  sync_after_smem_writes();  // (the scratch -- the H1 panels -- is free again)
  const uint32_t tm = *tmem_slot;

  const int nt = (int)blockIdx.x < a.n_tiles ? (a.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  uint64_t *bar = bars, *bar_dw2 = bars + 1, *bar_dw1 = bars + 2;
  uint32_t rp = 0;
  const long long rows_total = (long long)a.T * a.n;
  {  // S_g: the gradient chain of this CTA runs on dY * S_g, max|A| * S_g in [8, 16) over the CTA's own samples
    float mx = 0.f;
    for (int q = threadIdx.x; q < nt * SPT; q += blockDim.x) {
      const long long k = (long long)(blockIdx.x + (q / SPT) * gridDim.x) * SPT + q % SPT;
      if (k < rows_total)
        mx = fmaxf(mx, fabsf(a.adv[k]));
    }
    const float sg = pow2_scale(block_max(mx, fl + CM::F_RED), 4);
    if (threadIdx.x == 0) {
      fl[CM::F_SG] = sg;
      fl[CM::F_ISG] = 1.f / sg;
    }
    __syncthreads();
  }
  const float inv_s1 = fl[CM::F_IS1], inv_s2 = fl[CM::F_IS2], s_g = fl[CM::F_SG], inv_sg = fl[CM::F_ISG];

  // per-thread partial sums (drained after the last tile)
  float dw3[32], db3 = 0.f;
#pragma unroll
  for (int q = 0; q < 32; ++q)
    dw3[q] = 0.f;

  if (issuer) {
    // ================= MMA issuer
    auto layer1 = [&](int slot) {
      issue_gemm<1, false, false, false, true>(tm + CM::ACC0, sbase + CM::WX + 64 + 32 * slot, 0, sbase + CM::WX,
                                               sbase + CM::WX + 32, IDH<D1>::FK_FK, false);
      umma::commit(bar);
    };
    bool first = true;
    if (nt > 0) {
      ready_sync(0, rp, RT);  // observations of the first tile
      if (umma::elect_one())
        layer1(0);
      __syncwarp();
    }
    for (int j = 0; j < nt; ++j) {
      ready_sync(0, rp, RT);  // H1
      if (umma::elect_one()) {
        conv_issue_layer2<D1, D2, CM>(tm, sbase);
        umma::commit(bar);
      }
      __syncwarp();
      ready_sync(0, rp, RT);  // dH2
      if (umma::elect_one()) {
        // dH1 = dH2 . W2: A = dH2 from tensor memory (the head's copy in ACC1), B = W2 as MN-major operand
        // (contraction over W2's rows, N = D1 spans its 64-column sub-panels: LBO = W2SUB)
        constexpr int CH2 = 32;
#pragma unroll
        for (int k = 0; k < D2 / 16; ++k) {
          const uint32_t ah = tm + CM::ACC1 + (16 * k / CH2) * CH2 + ((16 * k % CH2) / 16) * 8, al = ah + CH2 / 2;
          const uint64_t bh = desc_lo_hi(desc_lo(sbase + CM::W2_HI + k * umma::KSTEP_BYTES_MNMAJOR, CM::W2SUB));
          const uint64_t bl = desc_lo_hi(desc_lo(sbase + CM::W2_LO + k * umma::KSTEP_BYTES_MNMAJOR, CM::W2SUB));
          umma::mma_bf16_ta(tm + CM::ACC0, ah, bh, IDH<D1>::BK_FM, k > 0 ? 1u : 0u);
          umma::mma_bf16_ta(tm + CM::ACC0, ah, bl, IDH<D1>::BK_FM, 1);
          umma::mma_bf16_ta(tm + CM::ACC0, al, bh, IDH<D1>::BK_FM, 1);
          if (c_conv_lolo)
            umma::mma_bf16_ta(tm + CM::ACC0, al, bl, IDH<D1>::BK_FM, 1);
        }
        umma::commit(bar);
        // dW2 (+)= [hi(dH2); lo(dH2)]^T . hi(H1) + [hi(dH2); lo(dH2)]^T . lo(H1): runs behind the dH1 epilogue
        issue_gemm<8, true, true, false, true>(tm + CM::DA, sbase + CM::G2_HI, 0, sbase + CM::H1_HI, sbase + CM::H1_LO,
                                               IDH<D1>::BM_FM, !first);
        // db2 (+)= [hi(dH2); lo(dH2)]^T . [X | 1]: column 4 = the sums over rows (the other columns are not used)
        issue_gemm<8, true, true, false, false>(tm + CM::DB2, sbase + CM::G2_HI, 0, sbase + CM::WX + 64 + 32 * (j & 1), 0,
                                                IDH<16>::BM_FM, !first);
        umma::commit(bar_dw2);
      }
      __syncwarp();
      ready_sync(0, rp, RT);  // dH1, and the next tile's observations in the other slot
      if (umma::elect_one()) {
        if (j + 1 < nt)  // ahead of dW1: the pipe is in order, dW1 runs behind the next tile's first epilogue
          layer1((j + 1) & 1);
        const uint32_t xb = sbase + CM::WX + 64 + 32 * (j & 1);
        if (D1 == 128)  // M = D1 = 128 over the two column panels of dH1, hi and lo passes
          issue_gemm_mn_lbo<8>(tm + CM::DB, sbase + CM::DH1_HI, sbase + CM::DH1_LO, PANEL, xb, IDH<16>::BM_FM, !first);
        else  // D1 = 64: [hi(dH1); lo(dH1)] stacked (the lo panel follows the hi panel), one MMA per K step
          issue_gemm<8, true, true, false, false>(tm + CM::DB, sbase + CM::DH1_HI, 0, xb, 0, IDH<16>::BM_FM, !first);
        umma::commit(bar_dw1);
      }
      __syncwarp();
      first = false;
    }
  } else {
    // ================= epilogue threads: (row, half)
    uint32_t phase = 0, phase_dw2 = 0, phase_dw1 = 0;
    auto wait_mma = [&]() {
      umma::mbar_wait(bar, phase);
      phase ^= 1;
      umma::fence_after_sync();
    };
    const float *b1 = fl + CM::F_B1, *b2 = fl + CM::F_B2, *w3 = fl + CM::F_W3;
    const int bin = t.row & 7, g0 = t.lane & ~7;
    const int c1_0 = half * (NC1 / 2), c1_1 = c1_0 + NC1 / 2;  // this thread's chunks of a D1-wide layer
    const bool has2 = half < NC2;                               // ... and its chunk (index `half`) of the D2-wide layer
    // raw state of this row's (sample, bin) of a tile: bin w / h, item w / h
    auto load_state = [&](int tile, int &bw, int &bh, int &iw, int &ih) {
      const long long k = (long long)tile * SPT + (t.row >> 3);
      bw = bh = iw = ih = 0;
      if (k < rows_total) {
        const int tt = (int)(k / a.n), i = (int)(k % a.n);
        const int8_t *src = a.rec_state + (size_t)tt * P * a.stride + i;
        bw = src[(size_t)(2 * bin) * a.stride];
        bh = src[(size_t)(2 * bin + 1) * a.stride];
        iw = src[(size_t)(2 * NB) * a.stride];
        ih = src[(size_t)(2 * NB + 1) * a.stride];
      }
    };
    long long *clk = (PROBE && a.clk && blockIdx.x == 0 && threadIdx.x == 96) ? a.clk : nullptr;
    int clk_n = 0;
#define STAMP() do { if (PROBE && clk && clk_n < 98) clk[clk_n++] = clock64(); } while (0)
    int nbw = 0, nbh = 0, niw = 0, nih = 0;
    if (nt > 0) {
      if (half == 0) {
        load_state(blockIdx.x, nbw, nbh, niw, nih);
        conv_encode_row<CM>(smem, t.row, 0, nbw, nbh, niw, nih, a.inv_w, a.inv_h);
      }
      ready_arrive(0, rp, RT);
    }
    for (int j = 0; j < nt; ++j) {
      const int tile = blockIdx.x + j * gridDim.x;
      // global loads consumed in the head / at the end of the tile
      const long long k = (long long)tile * SPT + (t.row >> 3);
      const bool valid = k < rows_total;
      int act = 0;
      float A = 0.f, pold_b = 1.f;
      if (valid) {
        act = a.rec_action[k];
        A = a.adv[k];
        pold_b = a.p_old[k * NB + bin];
      }
      act = act < NB ? act : NB - 1;
      const bool has_next = j + 1 < nt;
      if (has_next && half == 0)
        load_state(tile + gridDim.x, nbw, nbh, niw, nih);
      STAMP();
      wait_mma();  // layer 1
      if (j > 0) {  // the previous tile's dW2 GEMM (reads the H1 and dH2 panels) ran behind its dH1 epilogue
        umma::mbar_wait(bar_dw2, phase_dw2);
        phase_dw2 ^= 1;
      }
      STAMP();
      conv_epi_fwd<D1, true>(tm + CM::ACC0, t, b1, inv_s1, smem + CM::H1_HI, smem + CM::H1_LO, c1_0, c1_1);
      ready_arrive(0, rp, RT);
      STAMP();
      wait_mma();  // layer 2
      STAMP();
      // ---- head: H2 = relu(acc + b2) (registers), logit = H2 . w3 + b3 (the row's two threads exchange
      //      their partial sums), softmax over the sample's 8 rows, loss gradient, Jacobian -> dY
      float y[32], s = 0.f;
      if (has2) {
        float v[32];
        tmem_load<32>(tm + CM::ACC1 + t.lane_base + half * 32, v);
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          y[q] = fmaxf(fmaf(v[q], inv_s2, b2[half * 32 + q]), 0.f);
          s = fmaf(y[q], w3[half * 32 + q], s);
        }
      } else {
#pragma unroll
        for (int q = 0; q < 32; ++q)
          y[q] = 0.f;
      }
      scr[half * TILE + t.row] = s;
      asm volatile("bar.sync 5, 256;\n" ::: "memory");
      const float logit = (scr[t.row] + scr[TILE + t.row]) + fl[CM::F_B3];
      const float e = expf(logit);  // no max subtraction (nn.h:382-392)
      float p[NB], ssum = 0.f;
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        p[q] = __shfl_sync(0xffffffffu, e, g0 + q);
        ssum += p[q];
      }
      const float inv_s = 1.f / ssum;
#pragma unroll
      for (int q = 0; q < NB; ++q)
        p[q] = p[q] * inv_s;
      const float pold = __shfl_sync(0xffffffffu, pold_b, g0 + act);
      float g[NB];
      if (a.loss_kind == DFRL_LOSS_CLIPPED) {
        float pa = 0.f;
#pragma unroll
        for (int q = 0; q < NB; ++q)
          pa = (q == act) ? p[q] : pa;
        const float gc = clipped_grad(pa, pold, A);
#pragma unroll
        for (int q = 0; q < NB; ++q)
          g[q] = (q == act) ? gc : 0.f;
      } else {
#pragma unroll
        for (int q = 0; q < NB; ++q)
          g[q] = p[q] * A - (q == act ? A : 0.f);
      }
      float dY = 0.f;
      if (a.head_bwd == HEAD_JACOBIAN) {
        float dot = 0.f;
#pragma unroll
        for (int q = 0; q < NB; ++q)
          dot = fmaf(p[q], g[q], dot);
#pragma unroll
        for (int q = 0; q < NB; ++q)
          dY = (q == bin) ? p[q] * (g[q] - dot) : dY;
      } else {
#pragma unroll
        for (int q = 0; q < NB; ++q)
          dY = (q == bin) ? g[q] : dY;
      }
      dY = valid ? dY : 0.f;
      if (half == 0)
        db3 += dY;
      if (has2) {
        // dH2 = dY w3 . relu' (rank 1), dW3 += dY H2; dH2 -> its panels (A of the dW2 / db2 GEMMs) and, in
        // place over the accumulator columns just read, tensor memory (A of the dH1 GEMM)
        float v[32];
        const float dYs = dY * s_g;
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          v[q] = y[q] > 0.f ? dYs * w3[half * 32 + q] : 0.f;
          dw3[q] = fmaf(dY, y[q], dw3[q]);
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          uint4 hh, ll;
          split8_h(&v[8 * cc], hh, ll);
          const uint32_t off = umma::panel_chunk_off(t.row, half * 4 + cc);
          *reinterpret_cast<uint4 *>(smem + CM::G2_HI + off) = hh;
          *reinterpret_cast<uint4 *>(smem + CM::G2_LO + off) = ll;
          tmem_put_chunk<32>(tm + CM::ACC1 + t.lane_base + half * 32, cc, hh, ll);
        }
        umma::tmem_st_wait();
      }
      ready_arrive(0, rp, RT);
      STAMP();
      wait_mma();  // dH1
      if (j > 0) {  // the previous tile's dW1 GEMM (reads the dH1 panels and its observation slot)
        umma::mbar_wait(bar_dw1, phase_dw1);
        phase_dw1 ^= 1;
      }
      STAMP();
      conv_epi_bwd<D1>(tm + CM::ACC0, t, inv_s2, smem + CM::H1_HI, smem + CM::DH1_HI, smem + CM::DH1_LO, c1_0, c1_1);
      if (has_next && half == 0)
        conv_encode_row<CM>(smem, t.row, (j + 1) & 1, nbw, nbh, niw, nih, a.inv_w, a.inv_h);
      ready_arrive(0, rp, RT);
      STAMP();
    }
#undef STAMP
    if (nt > 0) {  // the last tile's weight-gradient GEMMs
      umma::mbar_wait(bar_dw2, phase_dw2);
      umma::mbar_wait(bar_dw1, phase_dw1);
      umma::fence_after_sync();
    }
  }

  // ---- drain: this CTA's partial gradient -> global (scratch = the dead activation panels)
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  float *part = a.partials + (size_t)blockIdx.x * partial_stride(net.n_params);
  float *sc = reinterpret_cast<float *>(smem + CM::H1_HI);
  // the partial gradient is staged in shared memory (the dead dH1 panels behind the tail's 8 KB of scratch; one pad
  // word per 32) and leaves as coalesced stores: the accumulators are read row-wise, direct stores scatter
  float *sg = reinterpret_cast<float *>(smem + CM::DH1_HI + 8192);
  static_assert(2 * CM::NP1 * PANEL - 8192 >= (D1 == 128 ? 9280u : 2560u) * 4, "staging area of the partial gradient");
  auto put = [&](int i, float v) { sg[i + (i >> 5)] = v; };
  if (nt == 0) {
    for (int q = threadIdx.x; q < net.n_params; q += blockDim.x)
      part[q] = 0.f;
  } else {
    constexpr int HC = D1 / 2;  // columns of DA per (row, half) thread
    // dW2[n][k] = DA[lane n][k] + DA[lane 64 + n][k]: the upper lanes go through shared memory
    if (!issuer && t.row >= 64) {
#pragma unroll
      for (int c0 = 0; c0 < HC; c0 += 32) {
        float v[32];
        tmem_load<32>(tm + CM::DA + t.lane_base + half * HC + c0, v);
#pragma unroll
        for (int q = 0; q < 32; ++q)
          sc[(t.row - 64) * (D1 + 1) + half * HC + c0 + q] = v[q];
      }
    }
    __syncthreads();
    if (!issuer && t.row < D2) {
#pragma unroll
      for (int c0 = 0; c0 < HC; c0 += 32) {
        float v[32];
        tmem_load<32>(tm + CM::DA + t.lane_base + half * HC + c0, v);
#pragma unroll
        for (int q = 0; q < 32; ++q)
          put(net.o_w2 + t.row * D1 + half * HC + c0 + q, (v[q] + sc[t.row * (D1 + 1) + half * HC + c0 + q]) * inv_sg);
      }
    }
    // [dW1 | db1]: DB lane n, columns 0..3 and 4 (D1 = 64: the lo part in lanes 64..127)
    if (!issuer && half == 0) {
      float w[8];
      tmem_load<8>(tm + CM::DB + t.lane_base, w);
      if (D1 == 128) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          put(net.o_w1 + t.row * 4 + q, w[q] * inv_sg);
        put(net.o_b1 + t.row, w[4] * inv_sg);
      } else {
        float *s2 = sc + 64 * (D1 + 1);
        if (t.row >= 64)
#pragma unroll
          for (int q = 0; q < 5; ++q)
            s2[(t.row - 64) * 8 + q] = w[q];
        asm volatile("bar.sync 6, 128;\n" ::: "memory");
        if (t.row < 64) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            put(net.o_w1 + t.row * 4 + q, (w[q] + s2[t.row * 8 + q]) * inv_sg);
          put(net.o_b1 + t.row, (w[4] + s2[t.row * 8 + 4]) * inv_sg);
        }
      }
    }
    __syncthreads();
    // db2[n] = DB2[lane n][4] + DB2[lane 64 + n][4]; dW3 / db3: fixed-order sums over the 128 rows
    float *r3 = sc, *r2 = sc + TILE * (D2 + 1), *r1 = r2 + TILE;
    if (!issuer) {
      if (half < NC2) {
#pragma unroll
        for (int q = 0; q < 32; ++q)
          r3[t.row * (D2 + 1) + half * 32 + q] = dw3[q];
      }
      if (half == 0) {
        float w[8];
        tmem_load<8>(tm + CM::DB2 + t.lane_base, w);
        r2[t.row] = w[4];
        r1[t.row] = db3;
      }
    }
    __syncthreads();
    if (threadIdx.x < D2) {
      float s3 = 0.f;
      for (int r = 0; r < TILE; ++r)
        s3 += r3[r * (D2 + 1) + threadIdx.x];
      put(net.o_w3 + threadIdx.x, s3);
      put(net.o_b2 + threadIdx.x, (r2[threadIdx.x] + r2[64 + threadIdx.x]) * inv_sg);
    } else if (threadIdx.x == D2) {
      float s1 = 0.f;
      for (int r = 0; r < TILE; ++r)
        s1 += r1[r];
      put(net.o_b3, s1);
    }
    __syncthreads();
    for (int q = threadIdx.x; q < net.n_params; q += blockDim.x)
      part[q] = sg[q + (q >> 5)];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (t.warp == 0)
    umma::tmem_dealloc(tm, 512);
  gradient_tail(a.partials, net, a.tail, reinterpret_cast<float *>(smem + CM::DH1_HI));
}

// ---------------------------------------------------------------------------------------------
// Rollout (agent::play_steps(T), rl.h:325-360) for the conv1d policy: NP tile pipelines of 16
// environments x 8 bins; thread (env, bin) keeps ITS bin and the env's item in registers for all T
// steps; softmax, sampling (libstdc++ discrete_distribution semantics, device_fns.cuh) and
// environment::apply / reset / next item are evaluated redundantly by the 8 lanes of an environment
// (identical inputs, identical results), lane `bin` records its own planes. Hidden activations only in
// tensor memory.
template <int D1, int D2, int NP>
struct cvrmap {
  static constexpr int NP1 = D1 / 64;
  static constexpr uint32_t WX = 0;  // as cvmap::WX, observation slot = pipeline index
  static constexpr uint32_t W2SUB = D2 * 128;
  static constexpr uint32_t W2_HI = WX + PANEL, W2_LO = W2_HI + NP1 * W2SUB;
  static constexpr uint32_t FLOATS = W2_LO + NP1 * W2SUB;
  static constexpr int F_B1 = 0, F_B2 = D1, F_W3 = D1 + D2, F_B3 = D1 + 2 * D2, F_IS1 = F_B3 + 1, F_IS2 = F_B3 + 2, F_RED = F_B3 + 8,
                       N_FLOATS = F_RED + 32;
  static constexpr uint32_t SCRATCH = (FLOATS + N_FLOATS * 4 + 1023) / 1024 * 1024;  // parameter staging (<= 36 KB)
  static constexpr uint32_t BARS = SCRATCH + 36 * 1024;
  static constexpr uint32_t TOTAL = BARS + 128;
  static constexpr uint32_t ACC0 = 0, ACC1 = 128;  // per pipeline (256 columns each)
  static_assert(NP == 2, "two observation slots in the WX panel, 2 x 256 tensor-memory columns");
};

template <int D1, int D2, int NP>
__global__ void __launch_bounds__(160 * NP, 1) fused_conv_rollout_kernel(rollout_args a) {
  using RM = cvrmap<D1, D2, NP>;
  constexpr int B = 8, P = 2 * B + 2, EPT = TILE / B;  // bins, planes, environments per tile
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
  float *fl = reinterpret_cast<float *>(smem + RM::FLOATS);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + RM::BARS);  // [wg]: MMA completion
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + RM::BARS + 96);
  const net3 net = a.net;
  const env_params &ep = a.ep;
  const tid_t t = thread_id();
  const bool issuer = t.warp >= 4 * NP;                   // warp-uniform
  const int wg = issuer ? t.warp - 4 * NP : t.warp >> 2;  // pipeline index
  const uint32_t sbase = umma::smem_u32(smem);

  umma::pdl_launch_dependents();
  if (t.warp == 0)
    umma::tmem_alloc(tmem_slot, 512);
  if (threadIdx.x == 0) {
    for (int q = 0; q < NP; ++q)
      umma::mbar_init(bars + q, 1);
    umma::fence_mbar_init();
  }
  zero_bytes(smem + RM::WX, PANEL);
  __syncthreads();
  stage_conv_weights<D1, D2, RM>(stage_params_bulk(a.params, net.n_params, smem + RM::SCRATCH, bars + 8), net, smem, fl);
  sync_after_smem_writes();
  const uint32_t tmem = *tmem_slot;

  const int nt = (int)blockIdx.x < a.n_tiles ? (a.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const uint32_t tm = tmem + 256u * wg;
  uint64_t *bar = bars + wg;
  uint32_t rp = 0;

  if (issuer) {
    for (int j = wg; j < nt; j += NP)
      for (int tt = 0; tt < a.T; ++tt) {
        ready_sync(wg, rp);  // observations staged in this pipeline's slot
        if (umma::elect_one()) {
          issue_gemm<1, false, false, false, true>(tm + RM::ACC0, sbase + RM::WX + 64 + 32 * wg, 0, sbase + RM::WX,
                                                   sbase + RM::WX + 32, IDH<D1>::FK_FK, false);
          umma::commit(bar);
        }
        __syncwarp();
        ready_sync(wg, rp);  // H1 (tensor memory)
        if (umma::elect_one()) {
          conv_issue_layer2<D1, D2, RM>(tm, sbase);
          umma::commit(bar);
        }
        __syncwarp();
      }
  } else {
    const float *b1 = fl + RM::F_B1, *b2 = fl + RM::F_B2, *w3 = fl + RM::F_W3;
    const float b3 = fl[RM::F_B3], inv_s1 = fl[RM::F_IS1], inv_s2 = fl[RM::F_IS2];
    uint32_t phase = 0;
    auto wait_mma = [&]() {
      umma::mbar_wait(bar, phase);
      phase ^= 1;
      umma::fence_after_sync();
    };
    unsigned long long c_eps = 0, c_reward = 0, c_steps = 0;
    const size_t S = ep.stride;
    const int bin = t.row & 7, g0 = t.lane & ~7;
    const uint32_t gmask = 0xffu << g0;
    for (int j = wg; j < nt; j += NP) {
      const int tile = blockIdx.x + j * gridDim.x;
      const int i = tile * EPT + (t.row >> 3);
      const bool owner = i < ep.n;
      int bw = 0, bh = 0, iw = 0, ih = 0;  // this thread's bin, the environment's item
      uint32_t my_draws = 0, my_steps = 0;
      if (owner) {
        bw = a.state[(size_t)(2 * bin) * S + i];
        bh = a.state[(size_t)(2 * bin + 1) * S + i];
        iw = a.state[(size_t)(2 * B) * S + i];
        ih = a.state[(size_t)(2 * B + 1) * S + i];
        my_draws = a.draws[i];
        my_steps = a.steps[i];
      }
      for (int tt = 0; tt < a.T; ++tt) {
        // ---- record the start state of step tt, stage the observation rows
        if (owner) {
          a.rec_state[((size_t)tt * P + 2 * bin) * S + i] = (int8_t)bw;
          a.rec_state[((size_t)tt * P + 2 * bin + 1) * S + i] = (int8_t)bh;
          if (bin == 0) {
            a.rec_state[((size_t)tt * P + 2 * B) * S + i] = (int8_t)iw;
            a.rec_state[((size_t)tt * P + 2 * B + 1) * S + i] = (int8_t)ih;
          }
        }
        conv_encode_row<RM>(smem, t.row, wg, bw, bh, iw, ih, a.inv_w, a.inv_h);
        ready_arrive(wg, rp);
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
        wait_mma();  // layer 1
        conv_epi_fwd<D1, false>(tm + RM::ACC0, t, b1, inv_s1, nullptr, nullptr, 0, D1 / 32);  // H1 only as a TMEM A operand
        ready_arrive(wg, rp);
        wait_mma();  // layer 2
        float logit = 0.f;
#pragma unroll
        for (int h = 0; h < D2 / 32; ++h) {
          float v[32];
          tmem_load<32>(tm + RM::ACC1 + t.lane_base + h * 32, v);
#pragma unroll
          for (int q = 0; q < 32; ++q)
            logit = fmaf(fmaxf(fmaf(v[q], inv_s2, b2[h * 32 + q]), 0.f), w3[h * 32 + q], logit);
        }
        logit += b3;
        // ---- head: softmax over the environment's 8 rows (no max subtraction, nn.h:382-392), action, apply
        const float e = expf(logit);
        float p[B], s = 0.f;
#pragma unroll
        for (int q = 0; q < B; ++q) {
          p[q] = __shfl_sync(0xffffffffu, e, g0 + q);
          s += p[q];
        }
#pragma unroll
        for (int q = 0; q < B; ++q)
          p[q] = p[q] / s;
        int act = 0;
        if (owner) {
          float mine = 0.f;
#pragma unroll
          for (int q = 0; q < B; ++q)
            mine = (q == bin) ? p[q] : mine;
          a.rec_probs[k * B + bin] = mine;
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
        }
        // environment::apply (bin_packing.h:53-64): the lane of bin `act` places the item
        const bool mine_hit = owner && bin == act;
        const int nw = bw - iw, nh = bh - ih;
        const bool over = (__ballot_sync(0xffffffffu, mine_hit && (nw < 0 || nh < 0)) & gmask) != 0;
        if (owner) {
          const int s1 = a.item_tape ? (tape_item != 0) : draw_shape1(ep, i, my_draws);
          if (over) {
            bw = ep.cap_w;
            bh = ep.cap_h;
          } else if (mine_hit) {
            bw = nw;
            bh = nh;
          }
          iw = s1 ? ep.iw0 : ep.iw1;
          ih = s1 ? ep.ih0 : ep.ih1;
          my_draws += 1;
          my_steps += 1;
          if (bin == 0) {
            a.rec_action[k] = (uint8_t)act;
            a.rec_done[k] = over;
            c_steps += 1;
            c_eps += over ? 1 : 0;
            c_reward += over ? 0 : 1;
          }
        }
      }
      // ---- live state back to the environment
      if (owner) {
        a.state[(size_t)(2 * bin) * S + i] = (int8_t)bw;
        a.state[(size_t)(2 * bin + 1) * S + i] = (int8_t)bh;
        if (bin == 0) {
          a.state[(size_t)(2 * B) * S + i] = (int8_t)iw;
          a.state[(size_t)(2 * B + 1) * S + i] = (int8_t)ih;
          a.draws[i] = my_draws;
          a.steps[i] = my_steps;
        }
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
    umma::tmem_dealloc(tmem, 512);
}
