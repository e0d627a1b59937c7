// conv_table.cuh -- the reference's conv1d_1 policies evaluated on their FINITE input domain; included by
// fused.cu inside its anonymous namespace (shares net3, rollout_args, conv_step_args, the gradient tail).
//
// A convolution1d_1_layer net over the bins (ppo_training.cc:10-16, nn.h:113-194) maps every (sample, bin) row
//     x = [bin.w / cap_w, bin.h / cap_h, item.w / cap_w, item.h / cap_h]          (bin_packing.h:31-40)
// through the SAME 4-D1-D2-1 MLP, and x takes at most (cap_w + 1)(cap_h + 1) x (cap_w + 1)(cap_h + 1) distinct
// values (6 561 for the reference's 8 x 8 bins; 90 of them are reachable with its two item shapes). So
//   forward:   logit(row) = L[d(row)] with a table L of the net's output on every domain entry d, recomputed
//              whenever the parameters change (6 561 rows instead of 8 x 524 288 per pass);
//   backward:  dTheta = sum_rows dY_row dlogit(x_row)/dTheta = sum_d G[d] dlogit(x_d)/dTheta with
//              G[d] = sum of dY over the rows whose input is d -- a histogram -- then ONE backward pass over the
//              entries with G[d] != 0.
// Same function, same gradient (the sums are regrouped, nothing is approximated); everything in fp32 FMA
// arithmetic, closer to the reference's own than the tensor-core kernels' 16-bit pairs. What is left per row is
// the head (softmax over the sample's 8 looked-up logits, loss gradient, Jacobian) and the record traffic: the
// kernels are HBM / L2-bound byte work instead of GEMMs (983 us -> about 30 us per policy step at 131 072 envs).
//
// Determinism: G accumulates in 64-bit FIXED POINT (dY * S rounded to an integer, S = a power of two with
// max|A| * S in [2^23, 2^24)): integer atomics commute, so the histogram -- and with it the gradient -- is
// bitwise reproducible whatever the order of the atomic adds.

__host__ __device__ __forceinline__ int tbl_index(int bw, int bh, int iw, int ih, int Dw, int Dh) {
  return ((iw * Dh + ih) * Dw + bw) * Dh + bh;
}
__device__ __forceinline__ int tbl_clamp(int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); }

// ---- L[d] = exp(logit) for every domain entry: one warp per pair of entries. W2 sits TRANSPOSED in shared memory
//      (row i = the weights of hidden-1 unit i, padded: conflict-free both ways), a lane owns D2 / 32 hidden-2 units of
//      both entries and walks the D1 inputs (H1 of the pair broadcast from shared memory): 8 instructions per input
//      and ONE warp reduction per pair. (A lane per input with a warp reduction per hidden-2 unit took 40
//      instructions per unit: 24 us per table instead of 7.)
template <int D1, int D2>
__global__ void __launch_bounds__(256) conv_table_forward_kernel(const float *__restrict__ params, net3 net, float inv_w,
                                                                 float inv_h, int Dw, int Dh, float *__restrict__ logits,
                                                                 const uint8_t *__restrict__ present) {
  constexpr int C = D1 / 32, C2 = D2 / 32, E = 2, LD = D2 + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = Dw * Dh * Dw * Dh;
  __shared__ float W2T[D1 * LD];
  __shared__ float H1s[8][E][D1];
  __shared__ float b2[D2], w3[D2];
  {
    const float *W2g = params + net.o_w2;  // [D2][D1]: read along i (coalesced), stored transposed
    for (int q = threadIdx.x; q < D2 * D1; q += blockDim.x)
      W2T[(q % D1) * LD + q / D1] = W2g[q];
    for (int q = threadIdx.x; q < D2; q += blockDim.x)
      b2[q] = params[net.o_b2 + q], w3[q] = params[net.o_w3 + q];
  }
  __syncthreads();
  const float *W1 = params + net.o_w1, *b1 = params + net.o_b1;
  const float b3 = params[net.o_b3];
  for (int d0 = E * (blockIdx.x * 8 + warp); d0 < D; d0 += E * gridDim.x * 8) {
    if (present) {  // only the entries that occur in a batch (unused: marking them costs more than the whole table)
      bool any = false;
      for (int e = 0; e < E; ++e)
        any = any || (d0 + e < D && present[d0 + e]);
      if (!any)
        continue;
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int d = min(d0 + e, D - 1);
      const int bh = d % Dh, bw = (d / Dh) % Dw, ih = (d / (Dh * Dw)) % Dh, iw = d / (Dh * Dw * Dh);
      const float x0 = (float)bw * inv_w, x1 = (float)bh * inv_h, x2 = (float)iw * inv_w, x3 = (float)ih * inv_h;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int j = lane + 32 * c;
        const float4 w = *reinterpret_cast<const float4 *>(W1 + 4 * j);
        H1s[warp][e][j] = fmaxf(fmaf(x3, w.w, fmaf(x2, w.z, fmaf(x1, w.y, fmaf(x0, w.x, b1[j])))), 0.f);
      }
    }
    __syncwarp();
    float acc[E][C2];
#pragma unroll
    for (int e = 0; e < E; ++e)
#pragma unroll
      for (int c = 0; c < C2; ++c)
        acc[e][c] = b2[lane + 32 * c];
#pragma unroll 8
    for (int i = 0; i < D1; ++i) {
      float w[C2];
#pragma unroll
      for (int c = 0; c < C2; ++c)
        w[c] = W2T[i * LD + lane + 32 * c];
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const float h = H1s[warp][e][i];
#pragma unroll
        for (int c = 0; c < C2; ++c)
          acc[e][c] = fmaf(h, w[c], acc[e][c]);
      }
    }
    float l[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      l[e] = 0.f;
#pragma unroll
      for (int c = 0; c < C2; ++c)
        l[e] = fmaf(fmaxf(acc[e][c], 0.f), w3[lane + 32 * c], l[e]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
        l[e] += __shfl_xor_sync(0xffffffffu, l[e], o);
    }
    if (lane == 0) {  // the table holds exp(logit): what the softmax of every row needs (no max subtraction, nn.h:382-392)
#pragma unroll
      for (int e = 0; e < E; ++e)
        if (d0 + e < D && (!present || present[d0 + e]))
          logits[d0 + e] = expf(l[e] + b3);
    }
    __syncwarp();
  }
}

// ---- max |A| over the rows (order-independent: unsigned max on the bit patterns of non-negative floats).
//      (Marking the domain entries that occur in the batch here, to restrict the table to them, was measured
//      slower than computing the whole table: 58 us of contended byte flags against 6 us.)
__global__ void conv_table_absmax_kernel(const float *__restrict__ adv, long long rows, unsigned *__restrict__ maxbits) {
  unsigned m = 0;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < rows; k += (long long)gridDim.x * blockDim.x) {
    const float a = fabsf(adv[k]);
    if (a == a && a < 3.0e38f)
      m = max(m, __float_as_uint(a));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m)
    atomicMax(maxbits, m);
}
__device__ __forceinline__ float tbl_scale(unsigned maxbits) {  // S = 2^(24 - e), max|A| = m 2^e with m in [0.5, 1)
  const float mx = __uint_as_float(maxbits);
  if (!(mx > 0.f))
    return 1.f;
  int e;
  frexpf(mx, &e);
  return ldexpf(1.f, 24 - e);
}

struct conv_table_args {
  conv_step_args s;            // the policy step's own arguments (records, loss kind, partials, tail)
  const float *logits;         // [D]
  unsigned long long *hist;    // [D] fixed-point G
  const unsigned *maxbits;     // max |A| (bit pattern)
  int Dw, Dh;
};

// Adds the lanes' (entry, fixed-point value) pairs of one slot to the shared histogram. Most rows of a batch hit
// the same few entries (the untouched bins of young episodes), and 32 atomics of a warp on one address serialise:
// the entry of the first active lane is summed over the lanes that share it (butterfly) and added once, twice
// over (more passes were measured slower); what is left goes one atomic per lane. Whole warp; d < 0: nothing to add.
// (a 64-bit shared-memory atomic add is a compare-and-swap loop on sm_100: the 64-bit sum is kept as two 32-bit
//  words, the carry of the low word's native atomic add -- known from the value it returns -- goes into the high one)
__device__ __forceinline__ void tbl_atomic_add64(unsigned *lo, unsigned *hi, int d, long long v) {
  const unsigned vl = (unsigned)(unsigned long long)v, vh = (unsigned)((unsigned long long)v >> 32);
  const unsigned old = atomicAdd(lo + d, vl);
  const unsigned carry = (old + vl) < old ? 1u : 0u;
  if (vh + carry)
    atomicAdd(hi + d, vh + carry);
}
__device__ __forceinline__ void tbl_add(unsigned *lo, unsigned *hi, int d, long long v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const unsigned act = __ballot_sync(0xffffffffu, d >= 0);
    if (act == 0)
      return;
    const int leader = __ffs(act) - 1;
    const int dl = __shfl_sync(0xffffffffu, d, leader);
    const bool mine = d == dl;
    const unsigned grp = __ballot_sync(0xffffffffu, mine);
    if (__popc(grp) < 4)
      break;
    long long c = mine ? v : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == leader && c != 0)
      tbl_atomic_add64(lo, hi, dl, c);
    if (mine)
      d = -1;
  }
  if (d >= 0 && v != 0)
    tbl_atomic_add64(lo, hi, d, v);
}

// ---- per row: head on the looked-up logits -> dY -> histogram. One thread per sample.
__global__ void __launch_bounds__(256) conv_table_head_kernel(conv_table_args ta) {
  extern __shared__ __align__(16) uint8_t tsm[];
  const conv_step_args &a = ta.s;
  constexpr int NB = 8, P = 2 * NB + 2;
  const int Dw = ta.Dw, Dh = ta.Dh, D = Dw * Dh * Dw * Dh;
  unsigned *hlo = reinterpret_cast<unsigned *>(tsm), *hhi = hlo + D;  // low / high words of the fixed-point sums
  float *ls = reinterpret_cast<float *>(tsm + (size_t)D * 8);
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    hlo[d] = 0u, hhi[d] = 0u;
    ls[d] = ta.logits[d];
  }
  __syncthreads();
  const float S = tbl_scale(*ta.maxbits);
  const long long rows = (long long)a.T * a.n;
  for (long long base = (long long)blockIdx.x * blockDim.x; base < rows; base += (long long)gridDim.x * blockDim.x) {
    const long long k = base + threadIdx.x;
    int d[NB];
    long long fx[NB];
#pragma unroll
    for (int q = 0; q < NB; ++q)
      d[q] = -1, fx[q] = 0;
    if (k < rows) {
      const int tt = (int)(k / a.n), i = (int)(k % a.n);
      const int8_t *src = a.rec_state + (size_t)tt * P * a.stride + i;
      int v[P];
#pragma unroll
      for (int q = 0; q < P; ++q)
        v[q] = src[(size_t)q * a.stride];
      int act = a.rec_action[k];
      act = act < NB ? act : NB - 1;
      const float A = a.adv[k];
      const float4 po0 = *reinterpret_cast<const float4 *>(a.p_old + k * NB), po1 = *reinterpret_cast<const float4 *>(a.p_old + k * NB + 4);
      const float pold[NB] = {po0.x, po0.y, po0.z, po0.w, po1.x, po1.y, po1.z, po1.w};
      const int iw = tbl_clamp(v[2 * NB], Dw - 1), ih = tbl_clamp(v[2 * NB + 1], Dh - 1);
      float p[NB], ssum = 0.f;
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        d[q] = tbl_index(tbl_clamp(v[2 * q], Dw - 1), tbl_clamp(v[2 * q + 1], Dh - 1), iw, ih, Dw, Dh);
        p[q] = ls[d[q]];  // exp(logit) from the table
        ssum += p[q];
      }
      const float inv_s = 1.f / ssum;
#pragma unroll
      for (int q = 0; q < NB; ++q)
        p[q] = p[q] * inv_s;
      float g[NB];
      if (a.loss_kind == DFRL_LOSS_CLIPPED) {
        float pa = 0.f, po = 1.f;
#pragma unroll
        for (int q = 0; q < NB; ++q) {
          pa = (q == act) ? p[q] : pa;
          po = (q == act) ? pold[q] : po;
        }
        const float gc = clipped_grad(pa, po, A);
#pragma unroll
        for (int q = 0; q < NB; ++q)
          g[q] = (q == act) ? gc : 0.f;
      } else {
#pragma unroll
        for (int q = 0; q < NB; ++q)
          g[q] = p[q] * A - (q == act ? A : 0.f);
      }
      float dot = 0.f;
      if (a.head_bwd == HEAD_JACOBIAN) {
#pragma unroll
        for (int q = 0; q < NB; ++q)
          dot = fmaf(p[q], g[q], dot);
      }
      // bins of the sample that share an entry (all the untouched bins of a young episode) are merged first, in
      // fp32 and in bin order (a fixed order inside the thread: still reproducible), then converted to fixed point
      float dY[NB];
#pragma unroll
      for (int q = 0; q < NB; ++q)
        dY[q] = a.head_bwd == HEAD_JACOBIAN ? p[q] * (g[q] - dot) : g[q];
#pragma unroll
      for (int q = 1; q < NB; ++q)
#pragma unroll
        for (int r = 0; r < q; ++r)
          if (d[q] >= 0 && d[q] == d[r]) {
            dY[r] += dY[q];
            d[q] = -1;
          }
#pragma unroll
      for (int q = 0; q < NB; ++q)
        fx[q] = d[q] >= 0 ? __float2ll_rn(dY[q] * S) : 0;
    }
#pragma unroll
    for (int q = 0; q < NB; ++q)
      tbl_add(hlo, hhi, d[q], fx[q]);
  }
  __syncthreads();
  for (int q = threadIdx.x; q < D; q += blockDim.x) {
    const unsigned long long h = ((unsigned long long)hhi[q] << 32) | hlo[q];
    if (h)
      atomicAdd(&ta.hist[q], h);
  }
}

// ---- backward over the entries with G[d] != 0, partial gradient per CTA, gradient tail (grid barrier, slice
//      reduction, exchange, optimizer) as in the other learner kernels. 256 threads, grid <= SM count.
template <int D1, int D2>
__global__ void __launch_bounds__(256, 1) conv_table_backward_kernel(conv_table_args ta) {
  constexpr int JG = 256 / D1, JPT = D2 / JG;  // thread (i = column of W2, jg): rows jg * JPT .. of dW2
  static_assert(256 % D1 == 0 && D2 % JG == 0, "thread layout");
  const conv_step_args &a = ta.s;
  const net3 net = a.net;
  __shared__ float H1[D1], H2[D2], G2[D2], G1[D1], X[4];
  __shared__ __align__(16) float scratch[(256 / 16) * 64];
  const int Dw = ta.Dw, Dh = ta.Dh, D = Dw * Dh * Dw * Dh;
  const float *P = a.params;
  const float *W1 = P + net.o_w1, *b1 = P + net.o_b1, *W2 = P + net.o_w2, *b2 = P + net.o_b2, *w3 = P + net.o_w3;
  const float inv_S = 1.f / tbl_scale(*ta.maxbits);
  const int tid = threadIdx.x, ci = tid % D1, jg = tid / D1;
  float dw2[JPT], dw1[4] = {0.f, 0.f, 0.f, 0.f}, db1 = 0.f, db2 = 0.f, dw3 = 0.f, db3 = 0.f;
#pragma unroll
  for (int q = 0; q < JPT; ++q)
    dw2[q] = 0.f;
  // this CTA's histogram entries, all loads in flight (read one by one in the loop below they were a chain of
  // ~ 44 L2 round trips: half of the kernel's 33 us)
  __shared__ long long myh[256];
  for (int d0 = blockIdx.x; d0 < D; d0 += 256 * gridDim.x) {
  {
    const int dd = d0 + tid * (int)gridDim.x;
    __syncthreads();
    myh[tid] = dd < D ? (long long)ta.hist[dd] : 0;
    __syncthreads();
  }
  for (int slot = 0; slot < 256; ++slot) {
    const int d = d0 + slot * (int)gridDim.x;
    if (d >= D)
      break;
    const long long fx = myh[slot];
    if (fx == 0)
      continue;  // (uniform: every thread reads the same entry)
    const float g = (float)((double)fx * (double)inv_S);
    const int bh = d % Dh, bw = (d / Dh) % Dw, ih = (d / (Dh * Dw)) % Dh, iw = d / (Dh * Dw * Dh);
    if (tid == 0) {
      X[0] = (float)bw * a.inv_w, X[1] = (float)bh * a.inv_h, X[2] = (float)iw * a.inv_w, X[3] = (float)ih * a.inv_h;
    }
    __syncthreads();
    if (tid < D1)
      H1[tid] = fmaxf(fmaf(X[3], W1[4 * tid + 3], fmaf(X[2], W1[4 * tid + 2], fmaf(X[1], W1[4 * tid + 1], fmaf(X[0], W1[4 * tid], b1[tid])))), 0.f);
    __syncthreads();
    {  // hidden 2: the dot products in 256 / D2 parts per unit (a thread per unit walked a row of W2 alone:
       // D1 dependent, uncoalesced loads on the kernel's critical path)
      constexpr int PARTS = 256 / D2, LEN = D1 / PARTS;
      const int j = tid % D2, part = tid / D2;
      const float *wr = W2 + (size_t)j * D1 + part * LEN;
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < LEN; ++i)
        acc = fmaf(H1[part * LEN + i], wr[i], acc);
      scratch[part * D2 + j] = acc;
    }
    __syncthreads();
    if (tid < D2) {
      constexpr int PARTS = 256 / D2;
      float acc = b2[tid];
#pragma unroll
      for (int q = 0; q < PARTS; ++q)
        acc += scratch[q * D2 + tid];
      const float h2 = fmaxf(acc, 0.f);
      H2[tid] = h2;
      const float g2 = h2 > 0.f ? g * w3[tid] : 0.f;  // dH2 = dY w3 . relu'
      G2[tid] = g2;
      dw3 = fmaf(g, h2, dw3);
      db2 += g2;
    }
    if (tid == 0)
      db3 += g;
    __syncthreads();
    if (tid < D1) {
      float acc = 0.f;
#pragma unroll 8
      for (int j = 0; j < D2; ++j)
        acc = fmaf(G2[j], W2[(size_t)j * D1 + tid], acc);
      const float g1 = H1[tid] > 0.f ? acc : 0.f;
      G1[tid] = g1;
      db1 += g1;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        dw1[c] = fmaf(g1, X[c], dw1[c]);
    }
    {
      const float h1 = H1[ci];
#pragma unroll
      for (int q = 0; q < JPT; ++q)
        dw2[q] = fmaf(G2[jg * JPT + q], h1, dw2[q]);
    }
    __syncthreads();
  }
  }
  // ---- this CTA's partial gradient (flat parameter order), then the tail
  float *part = a.partials + (size_t)blockIdx.x * partial_stride(net.n_params);
#pragma unroll
  for (int q = 0; q < JPT; ++q)
    part[net.o_w2 + (jg * JPT + q) * D1 + ci] = dw2[q];
  if (tid < D1) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
      part[net.o_w1 + 4 * tid + c] = dw1[c];
    part[net.o_b1 + tid] = db1;
  }
  if (tid < D2) {
    part[net.o_b2 + tid] = db2;
    part[net.o_w3 + tid] = dw3;
  }
  if (tid == 0)
    part[net.o_b3] = db3;
  __syncthreads();
  gradient_tail(a.partials, net, a.tail, scratch);
}

// ---- rollout (agent::play_steps(T), rl.h:325-360) on the table: thread = environment, state in registers
__global__ void __launch_bounds__(128) conv_table_rollout_kernel(rollout_args a, const float *__restrict__ logits, int Dw, int Dh) {
  extern __shared__ __align__(16) uint8_t tsm[];
  constexpr int B = 8, P = 2 * B + 2;
  float *ls = reinterpret_cast<float *>(tsm);
  const int D = Dw * Dh * Dw * Dh;
  for (int d = threadIdx.x; d < D; d += blockDim.x)
    ls[d] = logits[d];
  __syncthreads();
  const env_params &ep = a.ep;
  const size_t S = ep.stride;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool owner = i < ep.n;
  unsigned long long c_eps = 0, c_reward = 0, c_steps = 0;
  if (owner) {
    int st[P];
#pragma unroll
    for (int q = 0; q < P; ++q)
      st[q] = a.state[(size_t)q * S + i];
    uint32_t my_draws = a.draws[i], my_steps = a.steps[i];
    for (int tt = 0; tt < a.T; ++tt) {
#pragma unroll
      for (int q = 0; q < P; ++q)
        a.rec_state[((size_t)tt * P + q) * S + i] = (int8_t)st[q];
      const size_t k = (size_t)tt * ep.n + i;
      const int iw = st[2 * B], ih = st[2 * B + 1];
      const int ciw = tbl_clamp(iw, Dw - 1), cih = tbl_clamp(ih, Dh - 1);
      float p[B], s = 0.f;
#pragma unroll
      for (int q = 0; q < B; ++q) {
        p[q] = ls[tbl_index(tbl_clamp(st[2 * q], Dw - 1), tbl_clamp(st[2 * q + 1], Dh - 1), ciw, cih, Dw, Dh)];  // exp(logit)
        s += p[q];  // sequential sum in bin order, no max subtraction (nn.h:382-392)
      }
#pragma unroll
      for (int q = 0; q < B; ++q)
        p[q] = p[q] / s;
      float4 *pr = reinterpret_cast<float4 *>(a.rec_probs + k * B);
      pr[0] = make_float4(p[0], p[1], p[2], p[3]);
      pr[1] = make_float4(p[4], p[5], p[6], p[7]);
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
      // environment::apply (bin_packing.h:53-64), game over, reset, next item
      int bw = 0, bh = 0;
#pragma unroll
      for (int b = 0; b < B; ++b)
        if (b == act) {
          bw = st[2 * b] - iw;
          bh = st[2 * b + 1] - ih;
        }
      const bool over = bw < 0 || bh < 0;
      const int s1 = a.item_tape ? (a.item_tape[k] != 0) : draw_shape1(ep, i, my_draws);
#pragma unroll
      for (int b = 0; b < B; ++b) {
        if (over) {
          st[2 * b] = ep.cap_w;
          st[2 * b + 1] = ep.cap_h;
        } else if (b == act) {
          st[2 * b] = bw;
          st[2 * b + 1] = bh;
        }
      }
      st[2 * B] = s1 ? ep.iw0 : ep.iw1;
      st[2 * B + 1] = s1 ? ep.ih0 : ep.ih1;
      a.rec_done[k] = over;
      my_draws += 1;
      my_steps += 1;
      c_steps += 1;
      c_eps += over ? 1 : 0;
      c_reward += over ? 0 : 1;
    }
#pragma unroll
    for (int q = 0; q < P; ++q)
      a.state[(size_t)q * S + i] = (int8_t)st[q];
    a.draws[i] = my_draws;
    a.steps[i] = my_steps;
  }
  for (int o = 16; o > 0; o >>= 1) {
    c_steps += __shfl_down_sync(0xffffffffu, c_steps, o);
    c_eps += __shfl_down_sync(0xffffffffu, c_eps, o);
    c_reward += __shfl_down_sync(0xffffffffu, c_reward, o);
  }
  if ((threadIdx.x & 31) == 0 && c_steps) {
    atomicAdd(&a.counters[0], c_steps);
    atomicAdd(&a.counters[1], c_eps);
    atomicAdd(&a.counters[2], c_reward);
  }
}
