// Discriminator blocks on the pixel-major / tensor-core path (code/dmcnet_GAN/model.py:254-438).
//
// Every convolution of Discriminator / Discriminator2 / 3 / 5 runs as a tap GEMM of gemm_tc.cu on
// bf16 hi/lo pixel-major operands with 64 or 128 channels:
//   * 16-channel maps at 112x112 are held in SPACE-TO-DEPTH form (56x56 grid, channel = phase*16 + c),
//     the 2-channel 224x224 input in 4x4 space-to-depth form (56x56 grid, channel = (a4*4+b4)*2 + c),
//     32-channel maps are zero-padded to 64 channels;
//   * a stride-1 conv between two space-to-depth maps is a 9-tap GEMM whose 64x64 weight slices are
//     block-sparse, a stride-2 conv out of a space-to-depth map is a 4-tap (2x2) GEMM -- which
//     weight goes where is a host-built index table (dmcnet_b200/disc_plan.py), so ONE gather kernel
//     prepares all operand layouts and ONE gather kernel maps GEMM-space gradients back to OIHW;
//   * Conv bias + LeakyReLU(0.2) + Dropout2d live in the GEMM epilogue (ActFuse), BatchNorm(eps 0.8)
//     statistics in the same epilogue, the BatchNorm-backward reductions in the epilogue of the
//     data-gradient GEMM of the layer above (BwFuse).
// The kernels here are the glue: table-driven weight gather / gradient scatter, BatchNorm finalize
// with channel folding (a true channel owns up to four GEMM columns), the fused
// BatchNorm-backward + Dropout2d + LeakyReLU' pass producing the hi/lo gradient operand, the layout
// converters at the two ends, and the flatten + Linear(.., 2) head.
#include "common.cuh"

namespace dmc {

// ------------------------------------------------------------------ weights
// Wg[t][n][k] = map[..] >= 0 ? w[map[..]] : 0 as bf16 hi/lo, plus the transpose [t][k][n] (dgrad
// operand) and the bias expanded to GEMM columns.
__global__ void __launch_bounds__(256)
weight_gather_prep_kernel(const float* __restrict__ w, const int* __restrict__ map, int T, int N, int K,
                          bf16* __restrict__ W_hi, bf16* __restrict__ W_lo, bf16* __restrict__ Wt_hi,
                          bf16* __restrict__ Wt_lo, const float* __restrict__ bias,
                          const int* __restrict__ bmap, float* __restrict__ bias_exp) {
  const int total = T * N * K;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int m = map[i];
    const float v = m >= 0 ? w[m] : 0.f;
    bf16 h, l;
    split_bf16(v, h, l);
    W_hi[i] = h;
    W_lo[i] = l;
    if (Wt_hi) {
      const int k = i % K, n = (i / K) % N, t = i / (K * N);
      const int o = (t * K + k) * N + n;
      Wt_hi[o] = h;
      Wt_lo[o] = l;
    }
  }
  if (bias_exp && blockIdx.x == 0)
    for (int j = threadIdx.x; j < N; j += blockDim.x) bias_exp[j] = bmap[j] >= 0 ? bias[bmap[j]] : 0.f;
}

// Inference: BatchNorm folded into the conv operand.  W[t][co][ci] = w_oihw[co][ci][t] * scale[co] as bf16
// hi/lo, zero-padded to [T][Np][Kp]; bias_out[co] = shift[co] (0 on padding columns).
__global__ void __launch_bounds__(256)
weight_fold_prep_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                        const float* __restrict__ shift, int Cout, int Cin, int T, int Np, int Kp,
                        bf16* __restrict__ W_hi, bf16* __restrict__ W_lo, float* __restrict__ bias_out) {
  const int total = T * Np * Kp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i % Kp, n = (i / Kp) % Np, t = i / (Kp * Np);
    float v = 0.f;
    if (n < Cout && k < Cin) v = w[((long)n * Cin + k) * T + t] * scale[n];
    bf16 h, l;
    split_bf16(v, h, l);
    W_hi[i] = h;
    W_lo[i] = l;
  }
  if (blockIdx.x == 0)
    for (int j = threadIdx.x; j < Np; j += blockDim.x) bias_out[j] = j < Cout ? shift[j] : 0.f;
}

// dW[e] += sum_r dWg[inv[e][r]] (inv < 0: no entry), dbias[c] += sum_r dbias_exp[binv[c][r]].
// Fixed summation order: deterministic.
__global__ void __launch_bounds__(256)
weight_grad_gather_kernel(const float* __restrict__ dWg, const int* __restrict__ inv, int n_w, int R,
                          float* __restrict__ dW, const double* __restrict__ dbias_exp,
                          const int* __restrict__ binv, int C, float* __restrict__ dbias) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_w; e += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < R; ++r) {
      const int g = inv[e * R + r];
      if (g >= 0) s += dWg[g];
    }
    dW[e] += s;
  }
  if (dbias && blockIdx.x == 0)
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      double s = 0.0;
      for (int r = 0; r < R; ++r) {
        const int j = binv[c * R + r];
        if (j >= 0) s += dbias_exp[j];
      }
      dbias[c] += (float)s;
    }
}

// ------------------------------------------------------------------ BatchNorm with channel folding
// cmap[j] = true channel of GEMM column j (or -1 for a padding column).  Train mode (sums != null):
// torch BatchNorm2d semantics over count = frames * H * W samples per TRUE channel (biased variance
// to normalise, unbiased into running_var, momentum update, num_batches_tracked); eval mode
// (sums == null): running statistics.  Outputs are per GEMM column; padding columns get zeros.
__global__ void pm_bn_finalize_kernel(const double* __restrict__ sums, const int* __restrict__ cmap, int Cp,
                                      int C, double count, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, float* __restrict__ running_mean,
                                      float* __restrict__ running_var, long long* __restrict__ nbt,
                                      float momentum, float eps, float* __restrict__ scale,
                                      float* __restrict__ shift, float* __restrict__ mean_out,
                                      float* __restrict__ invstd_out) {
  __shared__ float s_mean[128], s_invstd[128];
  const int c = threadIdx.x;
  if (c < C) {
    if (sums) {
      double s1 = 0.0, s2 = 0.0;
      for (int j = 0; j < Cp; ++j)
        if (cmap[j] == c) { s1 += sums[j]; s2 += sums[Cp + j]; }
      const double mean = s1 / count;
      double var = s2 / count - mean * mean;
      if (var < 0) var = 0;
      s_mean[c] = (float)mean;
      s_invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
      if (running_mean) {
        const double unbiased = count > 1 ? var * count / (count - 1) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
      }
    } else {
      s_mean[c] = running_mean[c];
      s_invstd[c] = 1.0f / sqrtf(running_var[c] + eps);
    }
  }
  if (c == 0 && nbt && sums) *nbt += 1;
  __syncthreads();
  for (int j = threadIdx.x; j < Cp; j += blockDim.x) {
    const int t = cmap[j];
    float sc = 0.f, sh = 0.f, mu = 0.f, is = 0.f;
    if (t >= 0) {
      mu = s_mean[t];
      is = s_invstd[t];
      sc = gamma[t] * is;
      sh = beta[t] - mu * sc;
    }
    scale[j] = sc;
    shift[j] = sh;
    if (mean_out) { mean_out[j] = mu; invstd_out[j] = is; }
  }
}

// Backward coefficients per GEMM column from the two reductions S1 = sum dZ, S2 = sum dZ * xhat
// (accumulated per column, folded here per true channel):
//   dA = k1 * (dZ - k2 - xhat * k3),  k1 = gamma * invstd, k2 = S1 / count, k3 = S2 / count
// and the BatchNorm parameter gradients dgamma = S2, dbeta = S1 (written when non-null).
__global__ void pm_bn_bwd_fold_kernel(const double* __restrict__ sums2, const int* __restrict__ cmap, int Cp,
                                      int C, double count, const float* __restrict__ gamma,
                                      const float* __restrict__ invstd, float* __restrict__ coef,
                                      float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ double s1[128], s2[128];
  const int c = threadIdx.x;
  if (c < C) {
    double a = 0.0, b = 0.0;
    for (int j = 0; j < Cp; ++j)
      if (cmap[j] == c) { a += sums2[j]; b += sums2[Cp + j]; }
    s1[c] = a;
    s2[c] = b;
    if (dgamma) { dgamma[c] = (float)b; dbeta[c] = (float)a; }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < Cp; j += blockDim.x) {
    const int t = cmap[j];
    float k1 = 0.f, k2 = 0.f, k3 = 0.f;
    if (t >= 0) {
      k1 = gamma[t] * invstd[j];
      k2 = (float)(s1[t] / count);
      k3 = (float)(s2[t] / count);
    }
    coef[j] = k1;
    coef[Cp + j] = k2;
    coef[2 * Cp + j] = k3;
  }
}

// dPre = dA * mask[frame][j] * (A > 0 ? 1 : slope) with dA as above (coef == null: dA = dZ, the
// block without BatchNorm), written as bf16 hi/lo planes with a zero ring; dbias_exp[j] += sum dPre.
// Thread = 4 consecutive columns, striding over rows (coalesced 16-byte accesses).
__global__ void __launch_bounds__(256)
pm_act_bwd_kernel(const float* __restrict__ dZ, const float* __restrict__ A, const float* __restrict__ mean,
                  const float* __restrict__ invstd, const float* __restrict__ coef,
                  const float* __restrict__ mask, float slope, long P, int C, int Hp, int Wp,
                  bf16* __restrict__ G_hi, bf16* __restrict__ G_lo, double* __restrict__ dbias, int ring) {
  const int c4n = C / 4;
  const int rows_per_iter = blockDim.x / c4n;
  const int c = (threadIdx.x % c4n) * 4;
  const int r = threadIdx.x / c4n;
  float4 k1 = make_float4(1.f, 1.f, 1.f, 1.f), k2 = make_float4(0.f, 0.f, 0.f, 0.f), k3 = k2, mu = k2, is = k2;
  if (coef) {
    k1 = *reinterpret_cast<const float4*>(coef + c);
    k2 = *reinterpret_cast<const float4*>(coef + C + c);
    k3 = *reinterpret_cast<const float4*>(coef + 2 * C + c);
    mu = *reinterpret_cast<const float4*>(mean + c);
    is = *reinterpret_cast<const float4*>(invstd + c);
  }
  const unsigned frame_rows = (unsigned)(Hp * Wp);
  double acc[4] = {0, 0, 0, 0};
  float part[4] = {0, 0, 0, 0};
  int n = 0;
  const unsigned step = gridDim.x * rows_per_iter;
  for (unsigned q = blockIdx.x * rows_per_iter + r; q < (unsigned)P; q += step) {
    const long off = (long)q * C + c;
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    if (interior_r(q, Hp, Wp, ring)) {
      const float4 dz = *reinterpret_cast<const float4*>(dZ + off);
      const float4 a = *reinterpret_cast<const float4*>(A + off);
      float4 m = make_float4(1.f, 1.f, 1.f, 1.f);
      if (mask) m = *reinterpret_cast<const float4*>(mask + (long)(q / frame_rows) * C + c);
      o[0] = k1.x * (dz.x - k2.x - (a.x - mu.x) * is.x * k3.x) * m.x * (a.x > 0.f ? 1.f : slope);
      o[1] = k1.y * (dz.y - k2.y - (a.y - mu.y) * is.y * k3.y) * m.y * (a.y > 0.f ? 1.f : slope);
      o[2] = k1.z * (dz.z - k2.z - (a.z - mu.z) * is.z * k3.z) * m.z * (a.z > 0.f ? 1.f : slope);
      o[3] = k1.w * (dz.w - k2.w - (a.w - mu.w) * is.w * k3.w) * m.w * (a.w > 0.f ? 1.f : slope);
#pragma unroll
      for (int k = 0; k < 4; ++k) part[k] += o[k];
    }
    bf16 h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) split_bf16(o[k], h[k], l[k]);
    *reinterpret_cast<uint2*>(G_hi + off) = *reinterpret_cast<uint2*>(h);
    *reinterpret_cast<uint2*>(G_lo + off) = *reinterpret_cast<uint2*>(l);
    if (++n == 32) {
#pragma unroll
      for (int k = 0; k < 4; ++k) { acc[k] += part[k]; part[k] = 0.f; }
      n = 0;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) acc[k] += part[k];
  extern __shared__ double red[];                 // [blockDim.x][4]
#pragma unroll
  for (int k = 0; k < 4; ++k) red[threadIdx.x * 4 + k] = acc[k];
  __syncthreads();
  if (threadIdx.x < c4n && dbias) {
    for (int rr = 1; rr < rows_per_iter; ++rr)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] += red[(rr * c4n + threadIdx.x) * 4 + k];
#pragma unroll
    for (int k = 0; k < 4; ++k) atomicAdd(dbias + c + k, acc[k]);
  }
}

// ------------------------------------------------------------------ layout converters at the two ends
// planar x [M][2][H][W] fp32 -> 4x4 space-to-depth pixel-major hi/lo [M][H/4+1][W/4+1][64]:
// channel (a*4+b)*2 + c holds x[c][4i+a][4j+b]; channels 32..63 and the ring are zero.
// Thread = one 8-channel group of one grid pixel: group g < 4 is input row 4i+g, four columns, both
// channels (two float4 loads, one 16-byte store per plane).
__global__ void __launch_bounds__(256)
planar_to_s2d4_kernel(const float* __restrict__ x, long x_ns, int H, int W, int M, bf16* __restrict__ out_hi,
                      bf16* __restrict__ out_lo) {
  const int Hg = H / 4, Wg = W / 4, Hp = dmc_padded(Hg), Wp = dmc_padded(Wg);
  const long total = (long)M * Hp * Wp * 8;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int g = (int)(i & 7);
    long q = i >> 3;
    const int wp = (int)(q % Wp);
    const int hp = (int)((q / Wp) % Hp);
    const int m = (int)(q / ((long)Wp * Hp));
    bf16 h[8], l[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { h[k] = __float2bfloat16_rn(0.f); l[k] = h[k]; }
    if (g < 4 && hp >= 1 && wp >= 1) {
      const float* p = x + (long)m * x_ns + (long)(4 * (hp - 1) + g) * W + 4 * (wp - 1);
      const float4 c0 = *reinterpret_cast<const float4*>(p);
      const float4 c1 = *reinterpret_cast<const float4*>(p + (long)H * W);
      const float v[8] = {c0.x, c1.x, c0.y, c1.y, c0.z, c1.z, c0.w, c1.w};     // (b, c) pairs
#pragma unroll
      for (int k = 0; k < 8; ++k) split_bf16(v[k], h[k], l[k]);
    }
    *reinterpret_cast<uint4*>(out_hi + q * 64 + g * 8) = *reinterpret_cast<uint4*>(h);
    *reinterpret_cast<uint4*>(out_lo + q * 64 + g * 8) = *reinterpret_cast<uint4*>(l);
  }
}

// dX[m][c][4i+a][4j+b] (+)= dS[m][i+1][j+1][(a*4+b)*2 + c]  (dS fp32, 64 columns per pixel).
__global__ void __launch_bounds__(256)
s2d4_to_planar_kernel(const float* __restrict__ dS, int H, int W, int M, float* __restrict__ dX, long dx_ns,
                      int accumulate) {
  const int Hg = H / 4, Wg = W / 4, Hp = dmc_padded(Hg), Wp = dmc_padded(Wg);
  const long total = (long)M * 2 * H * Wg;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int j = (int)(i % Wg);
    long r = i / Wg;
    const int y = (int)(r % H); r /= H;
    const int c = (int)(r & 1);
    const int m = (int)(r >> 1);
    const float* s = dS + ((((long)m * Hp + (y >> 2) + 1) * Wp + j + 1) * 64) + (y & 3) * 8 + c;
    float4 v = make_float4(s[0], s[2], s[4], s[6]);
    float* d = dX + (long)m * dx_ns + ((long)c * H + y) * W + 4 * j;
    if (accumulate) {
      const float4 o = *reinterpret_cast<const float4*>(d);
      v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
    }
    *reinterpret_cast<float4*>(d) = v;
  }
}

// ------------------------------------------------------------------ flatten + Linear(C*H*W, 2)
// The reference flattens NCHW: feature index = c*H*W + h*W + w (code/dmcnet_GAN/model.py:296-299).
// out[m][o] = b[o] + sum Z[m][h+1][w+1][c] * Wl[o][c*HW + h*W + w],  Z = hi + lo.  Block = frame.
__global__ void __launch_bounds__(256)
pm_linear_fwd_kernel(const bf16* __restrict__ Z_hi, const bf16* __restrict__ Z_lo, const float* __restrict__ Wl,
                     const float* __restrict__ b, int C, int H, int W, float* __restrict__ out) {
  const int m = blockIdx.x, HW = H * W, K = C * HW;
  const int Hp = dmc_padded(H), Wp = dmc_padded(W);
  float s0 = 0.f, s1 = 0.f;
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    const int c = i % C, hw = i / C;
    const int h = hw / W, w = hw - h * W;
    const long off = (((long)m * Hp + h + 1) * Wp + w + 1) * C + c;
    const float z = join_bf16(Z_hi[off], Z_lo[off]);
    s0 = fmaf(z, Wl[c * HW + hw], s0);
    s1 = fmaf(z, Wl[K + c * HW + hw], s1);
  }
  __shared__ float r0[8], r1[8];
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  if ((threadIdx.x & 31) == 0) { r0[threadIdx.x >> 5] = s0; r1[threadIdx.x >> 5] = s1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c2 = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { a += r0[k]; c2 += r1[k]; }
    out[m * 2 + 0] = a + b[0];
    out[m * 2 + 1] = c2 + b[1];
  }
}

// dZ[m][h+1][w+1][c] = sum_o dv[m][o] * Wl[o][c*HW + hw]   (fp32, ring = 0)
__global__ void __launch_bounds__(256)
pm_linear_dx_kernel(const float* __restrict__ dv, const float* __restrict__ Wl, int M, int C, int H, int W,
                    float* __restrict__ dZ) {
  const int HW = H * W, K = C * HW, Hp = dmc_padded(H), Wp = dmc_padded(W);
  const long total = (long)M * Hp * Wp * C;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long q = i / C;
    const int wp = (int)(q % Wp);
    const int hp = (int)((q / Wp) % Hp);
    const int m = (int)(q / ((long)Wp * Hp));
    float v = 0.f;
    if (hp >= 1 && wp >= 1) {
      const int f = c * HW + (hp - 1) * W + (wp - 1);
      v = dv[m * 2] * Wl[f] + dv[m * 2 + 1] * Wl[K + f];
    }
    dZ[i] = v;
  }
}

// dWl[o][c*HW + hw] += sum_m dv[m][o] * Z[m][hw][c];  db[o] += sum_m dv[m][o]
__global__ void __launch_bounds__(256)
pm_linear_dw_kernel(const float* __restrict__ dv, const bf16* __restrict__ Z_hi, const bf16* __restrict__ Z_lo,
                    int M, int C, int H, int W, float* __restrict__ dWl, float* __restrict__ db) {
  const int HW = H * W, K = C * HW, Hp = dmc_padded(H), Wp = dmc_padded(W);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K) {
    const int c = i % C, hw = i / C;
    const int h = hw / W, w = hw - h * W;
    float s0 = 0.f, s1 = 0.f;
    for (int m = 0; m < M; ++m) {
      const long off = (((long)m * Hp + h + 1) * Wp + w + 1) * C + c;
      const float z = join_bf16(Z_hi[off], Z_lo[off]);
      s0 = fmaf(dv[m * 2], z, s0);
      s1 = fmaf(dv[m * 2 + 1], z, s1);
    }
    dWl[c * HW + hw] += s0;
    dWl[K + c * HW + hw] += s1;
  }
  if (blockIdx.x == 0 && threadIdx.x < 2 && db) {
    float s = 0.f;
    for (int m = 0; m < M; ++m) s += dv[m * 2 + threadIdx.x];
    db[threadIdx.x] += s;
  }
}

static inline unsigned grid_for(long n, int block, long cap = 148L * 16) {
  long g = cdiv(n, block);
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

}  // namespace dmc

using namespace dmc;

// GEMM operands of one discriminator conv from its fp32 OIHW weight through an index table:
// W_hi/lo [T][N][K] (fprop B operand), Wt_hi/lo [T][K][N] (dgrad B operand; may be NULL),
// bias_exp[j] = bias[bmap[j]] (0 where bmap < 0).  map / bmap: int32 device tables.
extern "C" int dmc_weight_gather_prep(const float* w, const int* map, int T, int N, int K, void* W_hi,
                                      void* W_lo, void* Wt_hi, void* Wt_lo, const float* bias,
                                      const int* bmap, float* bias_exp, void* stream) {
  DMC_REQUIRE(w && map && W_hi && W_lo && T > 0 && N > 0 && K > 0, "weight_gather_prep: bad arguments");
  DMC_REQUIRE((Wt_hi == nullptr) == (Wt_lo == nullptr), "weight_gather_prep: Wt_hi / Wt_lo");
  DMC_REQUIRE(bias_exp == nullptr || (bias && bmap), "weight_gather_prep: bias / bmap");
  weight_gather_prep_kernel<<<grid_for((long)T * N * K, 256), 256, 0, (cudaStream_t)stream>>>(
      w, map, T, N, K, (bf16*)W_hi, (bf16*)W_lo, (bf16*)Wt_hi, (bf16*)Wt_lo, bias, bmap, bias_exp);
  return dmc_check_launch("weight_gather_prep_kernel");
}

// Inference operands with eval-mode BatchNorm folded in: W_hi/lo [T][Np][Kp] = bf16 split of
// w_oihw[co][ci][t] * scale[co] (zero padding beyond Cout / Cin), bias_out [Np] = shift (0 on padding).
// scale / shift are the eval coefficients (dmc_bn_eval_coeffs / dmc_pm_bn_finalize in eval mode).
extern "C" int dmc_weight_fold_prep(const float* w_oihw, const float* scale, const float* shift, int Cout,
                                    int Cin, int T, int Np, int Kp, void* W_hi, void* W_lo, float* bias_out,
                                    void* stream) {
  DMC_REQUIRE(w_oihw && scale && shift && W_hi && W_lo && bias_out, "weight_fold_prep: null argument");
  DMC_REQUIRE(Cout >= 1 && Cin >= 1 && T >= 1 && Np >= Cout && Kp >= Cin, "weight_fold_prep: bad shape");
  weight_fold_prep_kernel<<<grid_for((long)T * Np * Kp, 256), 256, 0, (cudaStream_t)stream>>>(
      w_oihw, scale, shift, Cout, Cin, T, Np, Kp, (bf16*)W_hi, (bf16*)W_lo, bias_out);
  return dmc_check_launch("weight_fold_prep_kernel");
}

// OIHW gradient from the GEMM-space gradient: dW[e] += sum_{r<R} dWg[inv[e*R + r]] (entries < 0
// skipped); dbias[c] += sum_r dbias_exp[binv[c*R + r]] (dbias may be NULL).  Deterministic.
extern "C" int dmc_weight_grad_gather(const float* dWg, const int* inv, int n_w, int R, float* dW,
                                      const double* dbias_exp, const int* binv, int C, float* dbias,
                                      void* stream) {
  DMC_REQUIRE(dWg && inv && dW && n_w > 0 && R > 0, "weight_grad_gather: bad arguments");
  DMC_REQUIRE(dbias == nullptr || (dbias_exp && binv && C > 0), "weight_grad_gather: bias arguments");
  weight_grad_gather_kernel<<<grid_for(n_w, 256), 256, 0, (cudaStream_t)stream>>>(dWg, inv, n_w, R, dW,
                                                                                  dbias_exp, binv, C, dbias);
  return dmc_check_launch("weight_grad_gather_kernel");
}

// BatchNorm2d over GEMM columns with channel folding (cmap[j] = true channel of column j, -1 = padding).
// sums != NULL: train mode from the per-column sums / sums of squares (double [2][Cp]) over `count`
// samples per true channel, running statistics updated (torch semantics); sums == NULL: eval mode.
// Outputs per column: scale, shift (and mean, invstd when non-NULL); zeros on padding columns.
extern "C" int dmc_pm_bn_finalize(const double* sums, const int* cmap, int Cp, int C, double count,
                                  const float* gamma, const float* beta, float* running_mean,
                                  float* running_var, long long* num_batches_tracked, float momentum,
                                  float eps, float* scale, float* shift, float* mean, float* invstd,
                                  void* stream) {
  DMC_REQUIRE(cmap && gamma && beta && scale && shift, "pm_bn_finalize: null argument");
  DMC_REQUIRE(C >= 1 && C <= 128 && Cp >= C && Cp <= 512, "pm_bn_finalize: C=%d Cp=%d", C, Cp);
  DMC_REQUIRE(sums != nullptr || (running_mean && running_var), "pm_bn_finalize: eval mode needs running stats");
  DMC_REQUIRE((mean == nullptr) == (invstd == nullptr), "pm_bn_finalize: mean / invstd");
  pm_bn_finalize_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(sums, cmap, Cp, C, count, gamma, beta, running_mean,
                                                            running_var, num_batches_tracked, momentum, eps,
                                                            scale, shift, mean, invstd);
  return dmc_check_launch("pm_bn_finalize_kernel");
}

// coef [3][Cp] = (gamma*invstd, S1/count, S2/count) per column from sums2 = (sum dZ, sum dZ*xhat) per
// column, folded per true channel; dgamma[c] = S2, dbeta[c] = S1 when non-NULL.
extern "C" int dmc_pm_bn_bwd_fold(const double* sums2, const int* cmap, int Cp, int C, double count,
                                  const float* gamma, const float* invstd, float* coef, float* dgamma,
                                  float* dbeta, void* stream) {
  DMC_REQUIRE(sums2 && cmap && gamma && invstd && coef, "pm_bn_bwd_fold: null argument");
  DMC_REQUIRE(C >= 1 && C <= 128 && Cp >= C && Cp <= 512, "pm_bn_bwd_fold: C=%d Cp=%d", C, Cp);
  DMC_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "pm_bn_bwd_fold: dgamma / dbeta");
  pm_bn_bwd_fold_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(sums2, cmap, Cp, C, count, gamma, invstd, coef,
                                                            dgamma, dbeta);
  return dmc_check_launch("pm_bn_bwd_fold_kernel");
}

// Gradient of a discriminator block's pre-activation, as the hi/lo operand of its dgrad / wgrad GEMMs:
// dPre = dA * mask[frame][j] * (A > 0 ? 1 : slope),  dA = coef ? k1*(dZ - k2 - (A-mean)*invstd*k3) : dZ;
// zero on the ring; dbias_exp[j] += sum dPre (double, may be NULL).  A = the block's saved
// post-dropout activation [P][C]; mask [frames][C] or NULL.
extern "C" int dmc_pm_act_bwd(const float* dZ, const float* A, const float* mean, const float* invstd,
                              const float* coef, const float* mask, float slope, long P, int C, int Hp,
                              int Wp, void* G_hi, void* G_lo, double* dbias_exp, void* stream) {
  DMC_REQUIRE(dZ && A && G_hi && G_lo, "pm_act_bwd: null argument");
  DMC_REQUIRE(coef == nullptr || (mean && invstd), "pm_act_bwd: coef needs mean and invstd");
  DMC_REQUIRE(C % 4 == 0 && C >= 4 && C <= 1024 && 256 % (C / 4) == 0, "pm_act_bwd: C=%d", C);
  DMC_REQUIRE(P > 0 && P < (1L << 31) && Hp > 0 && Wp > 0, "pm_act_bwd: bad geometry");
  const int rows_per_iter = 256 / (C / 4);
  pm_act_bwd_kernel<<<grid_for(cdiv(P, rows_per_iter), 1, 148L * 8), 256, 256 * 4 * sizeof(double),
                      (cudaStream_t)stream>>>(dZ, A, mean, invstd, coef, mask, slope, P, C, Hp, Wp,
                                              (bf16*)G_hi, (bf16*)G_lo, dbias_exp, 1);
  return dmc_check_launch("pm_act_bwd_kernel");
}

// BatchNorm backward alone, from the folded coefficients, on a layout with a `ring`-wide zero ring:
// G = k1 * (dz - k2 - (Y - mean) * invstd * k3) as bf16 hi/lo (ContextNetwork blocks, where the
// activation gradient was already applied to dz by the data-gradient GEMM's epilogue).
extern "C" int dmc_pm_bn_bwd_apply(const float* dz, const float* Y, const float* mean, const float* invstd,
                                   const float* coef, long P, int C, int Hp, int Wp, int ring, void* G_hi,
                                   void* G_lo, void* stream) {
  DMC_REQUIRE(dz && Y && mean && invstd && coef && G_hi && G_lo, "pm_bn_bwd_apply: null argument");
  DMC_REQUIRE(C % 4 == 0 && C >= 4 && C <= 1024 && 256 % (C / 4) == 0, "pm_bn_bwd_apply: C=%d", C);
  DMC_REQUIRE(P > 0 && P < (1L << 31) && ring >= 1 && ring < Hp && ring < Wp, "pm_bn_bwd_apply: bad geometry");
  const int rows_per_iter = 256 / (C / 4);
  pm_act_bwd_kernel<<<grid_for(cdiv(P, rows_per_iter), 1, 148L * 8), 256, 256 * 4 * sizeof(double),
                      (cudaStream_t)stream>>>(dz, Y, mean, invstd, coef, nullptr, 1.f, P, C, Hp, Wp,
                                              (bf16*)G_hi, (bf16*)G_lo, nullptr, ring);
  return dmc_check_launch("pm_act_bwd_kernel");
}

// Discriminator input: planar x [M][2][H][W] (frame stride x_ns) -> 4x4 space-to-depth pixel-major
// hi/lo [M][H/4+1][W/4+1][64] (channel (a*4+b)*2+c = x[c][4i+a][4j+b]; 32..63 and the ring zero).
extern "C" int dmc_planar_to_s2d4(const float* x, long x_ns, int H, int W, int M, void* out_hi, void* out_lo,
                                  void* stream) {
  DMC_REQUIRE(x && out_hi && out_lo && M > 0, "planar_to_s2d4: null argument");
  DMC_REQUIRE(H % 4 == 0 && W % 4 == 0 && x_ns % 4 == 0, "planar_to_s2d4: H=%d W=%d must be multiples of 4", H, W);
  const long total = (long)M * dmc_padded(H / 4) * dmc_padded(W / 4) * 8;
  planar_to_s2d4_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, x_ns, H, W, M, (bf16*)out_hi,
                                                                               (bf16*)out_lo);
  return dmc_check_launch("planar_to_s2d4_kernel");
}

// Inverse for gradients: dX[m][c][4i+a][4j+b] (+)= dS[m][i+1][j+1][(a*4+b)*2+c], dS fp32 [..][64].
extern "C" int dmc_s2d4_to_planar(const float* dS, int H, int W, int M, float* dX, long dx_ns, int accumulate,
                                  void* stream) {
  DMC_REQUIRE(dS && dX && M > 0, "s2d4_to_planar: null argument");
  DMC_REQUIRE(H % 4 == 0 && W % 4 == 0 && dx_ns % 4 == 0, "s2d4_to_planar: H=%d W=%d", H, W);
  s2d4_to_planar_kernel<<<grid_for((long)M * 2 * H * (W / 4), 256), 256, 0, (cudaStream_t)stream>>>(
      dS, H, W, M, dX, dx_ns, accumulate);
  return dmc_check_launch("s2d4_to_planar_kernel");
}

// adv_layer of the discriminators on a pixel-major map: out[m][o] = b[o] + <flatten_NCHW(Z[m]), Wl[o]>,
// Z = hi + lo [M][H+1][W+1][C], Wl [2][C*H*W] (code/dmcnet_GAN/model.py:296-299).
extern "C" int dmc_pm_linear_fwd(const void* Z_hi, const void* Z_lo, const float* Wl, const float* b, int M,
                                 int C, int H, int W, float* out, void* stream) {
  DMC_REQUIRE(Z_hi && Z_lo && Wl && b && out && M > 0 && C > 0 && H > 0 && W > 0, "pm_linear_fwd: bad arguments");
  pm_linear_fwd_kernel<<<M, 256, 0, (cudaStream_t)stream>>>((const bf16*)Z_hi, (const bf16*)Z_lo, Wl, b, C, H, W,
                                                            out);
  return dmc_check_launch("pm_linear_fwd_kernel");
}

// Backward of dmc_pm_linear_fwd from dv [M][2]: dZ fp32 [M][H+1][W+1][C] (ring zero; may be NULL),
// dWl += , db += (both may be NULL together).
extern "C" int dmc_pm_linear_bwd(const float* dv, const void* Z_hi, const void* Z_lo, const float* Wl, int M,
                                 int C, int H, int W, float* dZ, float* dWl, float* db, void* stream) {
  DMC_REQUIRE(dv && Wl && M > 0 && C > 0 && H > 0 && W > 0, "pm_linear_bwd: bad arguments");
  DMC_REQUIRE(dWl == nullptr || (Z_hi && Z_lo), "pm_linear_bwd: dWl needs Z");
  cudaStream_t st = (cudaStream_t)stream;
  if (dZ) {
    const long total = (long)M * dmc_padded(H) * dmc_padded(W) * C;
    pm_linear_dx_kernel<<<grid_for(total, 256), 256, 0, st>>>(dv, Wl, M, C, H, W, dZ);
    int rc = dmc_check_launch("pm_linear_dx_kernel");
    if (rc) return rc;
  }
  if (dWl) {
    pm_linear_dw_kernel<<<(unsigned)cdiv((long)C * H * W, 256), 256, 0, st>>>(dv, (const bf16*)Z_hi,
                                                                              (const bf16*)Z_lo, M, C, H, W, dWl, db);
    return dmc_check_launch("pm_linear_dw_kernel");
  }
  return DMC_OK;
}

// ------------------------------------------------------------------ stem data gradient on the tensor cores
// d(loss)/d(gen_flow) through the classifier's 7x7/2 stem conv (GAN G-step; code/dmcnet_GAN/model.py:560:
// the classifier input is NOT detached).  With Y = 2u + a, X = 2v + b the gradient of input pixel
// (Y, X) reads dZ at (u + di, v + dj), di, dj in {-1, 0, 1, 2}, through kernel row r = a + 3 - 2 di:
// a 16-tap GEMM over the 112x112 grid with K = 64 output channels and N = (a, b, c) = 8 columns
// (padded to 32), operands from dmc_weight_gather_prep.  The two kernels below move the data:
// planar dZ -> pixel-major hi/lo with a ring of TWO zero rows / columns (offset +2 must not reach
// the next frame), and the GEMM result -> planar input gradient.
namespace dmc {

// in  [N][C][H][W] fp32 (frame stride in_ns)  ->  out [N][H+R][W+R][Cp], pixel (h, w) at (h+R, w+R),
// rows / columns < R and channels >= C zero.  Output either bf16 hi/lo planes (out_f32 == null) or one
// fp32 plane; act_hi != null multiplies by the LeakyReLU derivative (act > 0 ? 1 : slope) of a saved
// activation in the OUTPUT layout (the gradient entering the last ContextNetwork block).
// Block = one output row of one frame; 32-pixel x C tiles through shared memory.
__global__ void __launch_bounds__(256)
planar_to_pm_ring_kernel(const float* __restrict__ in, long in_ns, int C, int Cp, int H, int W, int R,
                         bf16* __restrict__ out_hi, bf16* __restrict__ out_lo, float* __restrict__ out_f32,
                         const bf16* __restrict__ act_hi, float slope) {
  extern __shared__ float tile[];                       // [C][33]
  const int hp = blockIdx.x, n = blockIdx.y;
  const int Wp = W + R;
  const long row_base = (((long)n * (H + R) + hp) * Wp) * Cp;
  const int zero_px = hp < R ? Wp : R;                  // whole ring row, or the ring columns of a data row
  if (out_f32) {
    for (long i = threadIdx.x; i < (long)zero_px * Cp / 4; i += blockDim.x)
      reinterpret_cast<float4*>(out_f32 + row_base)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    for (long i = threadIdx.x; i < (long)zero_px * Cp / 8; i += blockDim.x) {
      reinterpret_cast<uint4*>(out_hi + row_base)[i] = make_uint4(0, 0, 0, 0);
      reinterpret_cast<uint4*>(out_lo + row_base)[i] = make_uint4(0, 0, 0, 0);
    }
  }
  if (hp < R) return;
  const int h = hp - R;
  for (int w0 = 0; w0 < W; w0 += 32) {
    for (int i = threadIdx.x; i < C * 32; i += blockDim.x) {
      const int c = i >> 5, dw = i & 31;
      tile[c * 33 + dw] = (w0 + dw < W) ? in[(long)n * in_ns + ((long)c * H + h) * W + w0 + dw] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * Cp / 4; i += blockDim.x) {
      const int c = (i % (Cp / 4)) * 4, dw = i / (Cp / 4);
      if (w0 + dw < W) {
        const long o = row_base + (long)(w0 + dw + R) * Cp + c;
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (c + k < C) ? tile[(c + k) * 33 + dw] : 0.f;
        if (act_hi) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (!(__bfloat162float(act_hi[o + k]) > 0.f)) v[k] *= slope;
        }
        if (out_f32) {
          *reinterpret_cast<float4*>(out_f32 + o) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
          bf16 hh[4], ll[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) split_bf16(v[k], hh[k], ll[k]);
          *reinterpret_cast<uint2*>(out_hi + o) = *reinterpret_cast<uint2*>(hh);
          *reinterpret_cast<uint2*>(out_lo + o) = *reinterpret_cast<uint2*>(ll);
        }
      }
    }
    __syncthreads();
  }
}

// out[n][c][h][w] = Z[n][h+R][w+R][c] (+ add[n][c][h][w]),  Z = hi + lo with Cp columns, c < C.
// Block = one row of one frame.
__global__ void __launch_bounds__(256)
pm_ring_to_planar_kernel(const bf16* __restrict__ Z_hi, const bf16* __restrict__ Z_lo, int Cp, int C, int H,
                         int W, int R, const float* __restrict__ add, long add_ns, float* __restrict__ out,
                         long out_ns) {
  const int h = blockIdx.x, n = blockIdx.y;
  const long row = (((long)n * (H + R) + h + R) * (W + R) + R) * Cp;
  for (int i = threadIdx.x; i < C * W; i += blockDim.x) {
    const int c = i / W, w = i - c * W;
    float v = join_bf16(Z_hi[row + (long)w * Cp + c], Z_lo[row + (long)w * Cp + c]);
    if (add) v += add[(long)n * add_ns + ((long)c * H + h) * W + w];
    out[(long)n * out_ns + ((long)c * H + h) * W + w] = v;
  }
}

// dX[n][c][2u+a][2v+b] (+)= D[n][u+2][v+2][(a*2+b)*2 + c]   (D fp32 with ldD columns, ring 2)
__global__ void __launch_bounds__(256)
s2d2_ring2_to_planar_kernel(const float* __restrict__ D, int ldD, int H, int W, int N, float* __restrict__ dX,
                            long dx_ns, int accumulate) {
  const int Hg = H / 2, Wg = W / 2, Wp = Wg + 2;
  const long total = (long)N * 2 * H * Wg;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int v = (int)(i % Wg);
    long r = i / Wg;
    const int y = (int)(r % H); r /= H;
    const int c = (int)(r & 1);
    const int n = (int)(r >> 1);
    const float* s = D + (((long)n * (Hg + 2) + (y >> 1) + 2) * Wp + v + 2) * ldD + (y & 1) * 4 + c;
    float2 o = make_float2(s[0], s[2]);                 // b = 0, 1
    float* d = dX + (long)n * dx_ns + ((long)c * H + y) * W + 2 * v;
    if (accumulate) {
      const float2 p = *reinterpret_cast<const float2*>(d);
      o.x += p.x; o.y += p.y;
    }
    *reinterpret_cast<float2*>(d) = o;
  }
}

}  // namespace dmc

// planar fp32 [N][C][H][W] (frame stride in_ns) -> pixel-major bf16 hi/lo [N][H+2][W+2][C] with a
// two-pixel zero ring on the low sides (the A operand of the stem data-gradient GEMM).
extern "C" int dmc_planar_to_pm_ring2(const float* in, long in_ns, int C, int H, int W, int N, void* out_hi,
                                      void* out_lo, void* stream) {
  DMC_REQUIRE(in && out_hi && out_lo && N > 0, "planar_to_pm_ring2: null argument");
  DMC_REQUIRE(C % 8 == 0 && C <= 256 && H > 0 && W > 0, "planar_to_pm_ring2: C=%d", C);
  dim3 grid((unsigned)(H + 2), (unsigned)N);
  planar_to_pm_ring_kernel<<<grid, 256, C * 33 * sizeof(float), (cudaStream_t)stream>>>(
      in, in_ns, C, C, H, W, 2, (bf16*)out_hi, (bf16*)out_lo, nullptr, nullptr, 1.f);
  return dmc_check_launch("planar_to_pm_ring_kernel");
}

// planar fp32 [N][C][H][W] (frame stride in_ns) -> pixel-major [N][H+R][W+R][Cp] with an R-wide zero
// ring and channels C..Cp-1 zero: bf16 hi/lo planes, or (out_f32 != NULL) one fp32 plane.
// act_hi != NULL (same layout as the output): values are multiplied by (act > 0 ? 1 : slope), the
// LeakyReLU derivative -- the loss gradient entering ContextNetwork's last block.
extern "C" int dmc_planar_to_pm_ring(const float* in, long in_ns, int C, int Cp, int H, int W, int R, int N,
                                     void* out_hi, void* out_lo, float* out_f32, const void* act_hi,
                                     float slope, void* stream) {
  DMC_REQUIRE(in && N > 0 && ((out_hi && out_lo) || out_f32), "planar_to_pm_ring: null argument");
  DMC_REQUIRE(C >= 1 && C <= Cp && Cp % 8 == 0 && C <= 256 && R >= 1 && H > 0 && W > 0,
              "planar_to_pm_ring: C=%d Cp=%d R=%d", C, Cp, R);
  dim3 grid((unsigned)(H + R), (unsigned)N);
  planar_to_pm_ring_kernel<<<grid, 256, C * 33 * sizeof(float), (cudaStream_t)stream>>>(
      in, in_ns, C, Cp, H, W, R, (bf16*)out_hi, (bf16*)out_lo, out_f32, (const bf16*)act_hi, slope);
  return dmc_check_launch("planar_to_pm_ring_kernel");
}

// Inverse for activations: out[n][c][h][w] = (hi + lo)[n][h+R][w+R][c] (+ add[n][c][h][w]) for c < C
// (ContextNetwork output -> planar gen_flow, "+ input_mv" of code/dmcnet/model.py:346 fused).
extern "C" int dmc_pm_ring_to_planar(const void* Z_hi, const void* Z_lo, int Cp, int C, int H, int W, int R,
                                     int N, const float* add, long add_ns, float* out, long out_ns,
                                     void* stream) {
  DMC_REQUIRE(Z_hi && Z_lo && out && N > 0 && C >= 1 && C <= Cp && R >= 1, "pm_ring_to_planar: bad arguments");
  dim3 grid((unsigned)H, (unsigned)N);
  pm_ring_to_planar_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)Z_hi, (const bf16*)Z_lo, Cp, C,
                                                                   H, W, R, add, add_ns, out, out_ns);
  return dmc_check_launch("pm_ring_to_planar_kernel");
}

// Result of the stem data-gradient GEMM (2x2 space-to-depth columns (a*2+b)*2+c on the H/2 x W/2 grid,
// ring 2, ldD columns per pixel) -> planar dX [N][2][H][W] (frame stride dx_ns), optionally accumulating.
extern "C" int dmc_s2d2_ring2_to_planar(const float* D, int ldD, int H, int W, int N, float* dX, long dx_ns,
                                        int accumulate, void* stream) {
  DMC_REQUIRE(D && dX && N > 0 && ldD >= 8, "s2d2_ring2_to_planar: bad arguments");
  DMC_REQUIRE(H % 2 == 0 && W % 2 == 0 && dx_ns % 2 == 0, "s2d2_ring2_to_planar: H=%d W=%d", H, W);
  s2d2_ring2_to_planar_kernel<<<grid_for((long)N * 2 * H * (W / 2), 256), 256, 0, (cudaStream_t)stream>>>(
      D, ldD, H, W, N, dX, dx_ns, accumulate);
  return dmc_check_launch("s2d2_ring2_to_planar_kernel");
}
