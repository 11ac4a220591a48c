// Small-channel convolutions on planar NCHW fp32 tensors (CUDA cores).
//
// These serve the layers whose channel counts are far below a tensor-core tile:
// the DMC generator (EstimatorDenseNetTiny, code/dmcnet/model.py:172-194: Cin
// 5..33 -> Cout 8,8,6,4,2,2), the discriminator blocks (code/dmcnet_GAN/model.py:
// 254-279) and the classifier's 7x7/2 stem conv on the 2-channel DMC map
// (code/dmcnet/model.py:289-294).  Every tensor argument is (pointer, per-image
// element stride), so a layer can read / write a channel sub-range of the dense
// concat buffer without any torch.cat copy.
//
//   conv_fwd   : out = [mask *] lrelu_slope(bias + conv(in, w)) [+ add] [+ out]
//   conv_dgrad : dX (+)= conv_transpose(dY, w)           (stride 1 or 2)
//   conv_wgrad : dW += sum_n,pixels dY * in ; dbias += sum dY   (register-resident
//                accumulators over many tiles, one atomic flush per CTA)
// plus the planar BatchNorm / activation helpers the discriminator needs.
#include "common.cuh"

namespace dmc {

// ------------------------------------------------------------------ forward
template <int KS, int ST, int COG, int CK>
__global__ void __launch_bounds__(256)
conv_fwd_kernel(const float* __restrict__ in, long in_ns, int Cin, int H, int W,
                const float* __restrict__ w, const float* __restrict__ bias, int Cout,
                float* __restrict__ out, long out_ns, int Ho, int Wo, float slope,
                const float* __restrict__ mask, const float* __restrict__ add, long add_ns,
                int accumulate, int tiles_x) {
  constexpr int PAD = KS / 2;
  constexpr int IT = 15 * ST + KS;
  constexpr int ITP = IT + 1;
  __shared__ float in_s[CK][IT][ITP];
  __shared__ __align__(16) float w_s[CK][KS * KS][COG];
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  const int th0 = (blockIdx.x / tiles_x) * 16, tw0 = (blockIdx.x % tiles_x) * 16;
  const int co0 = blockIdx.y * COG;
  const int n = blockIdx.z;
  const int ih0 = th0 * ST - PAD, iw0 = tw0 * ST - PAD;
  const float* inp = in + (long)n * in_ns;
  float acc[COG];
#pragma unroll
  for (int g = 0; g < COG; ++g) acc[g] = 0.f;

  for (int c0 = 0; c0 < Cin; c0 += CK) {
    for (int i = tid; i < CK * IT * IT; i += 256) {
      const int c = i / (IT * IT), r = (i / IT) % IT, s = i % IT;
      const int ih = ih0 + r, iw = iw0 + s;
      float v = 0.f;
      if (c0 + c < Cin && ih >= 0 && ih < H && iw >= 0 && iw < W)
        v = inp[((long)(c0 + c) * H + ih) * W + iw];
      in_s[c][r][s] = v;
    }
    for (int i = tid; i < CK * KS * KS * COG; i += 256) {
      const int g = i % COG, t = (i / COG) % (KS * KS), c = i / (COG * KS * KS);
      float v = 0.f;
      if (c0 + c < Cin && co0 + g < Cout) v = w[((long)(co0 + g) * Cin + c0 + c) * (KS * KS) + t];
      w_s[c][t][g] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int c = 0; c < CK; ++c) {
#pragma unroll
      for (int r = 0; r < KS; ++r)
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          const float v = in_s[c][ty * ST + r][tx * ST + s];
#pragma unroll
          for (int g = 0; g < COG; ++g) acc[g] = fmaf(v, w_s[c][r * KS + s][g], acc[g]);
        }
    }
    __syncthreads();
  }
  const int oh = th0 + ty, ow = tw0 + tx;
  if (oh >= Ho || ow >= Wo) return;
#pragma unroll
  for (int g = 0; g < COG; ++g) {
    const int co = co0 + g;
    if (co >= Cout) break;
    float v = acc[g] + (bias ? bias[co] : 0.f);
    v = v < 0.f ? v * slope : v;
    if (mask) v *= mask[(long)n * Cout + co];
    const long o = (long)n * out_ns + ((long)co * Ho + oh) * Wo + ow;
    if (add) v += add[(long)n * add_ns + ((long)co * Ho + oh) * Wo + ow];
    if (accumulate) v += out[o];
    out[o] = v;
  }
}

// ------------------------------------------------------------------ data gradient
template <int KS, int ST, int CIG, int CK>
__global__ void __launch_bounds__(256)
conv_dgrad_kernel(const float* __restrict__ dY, long dy_ns, int Cout, int Ho, int Wo,
                  const float* __restrict__ w, int Cin, int ci_count, float* __restrict__ dX,
                  long dx_ns, int H, int W, int accumulate, int tiles_x) {
  constexpr int PAD = KS / 2;
  constexpr int DT = (15 + KS - 1) / ST + 2;
  constexpr int DTP = DT + 1;
  __shared__ float dy_s[CK][DT][DTP];
  __shared__ __align__(16) float w_s[CK][KS * KS][CIG];
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  const int h0 = (blockIdx.x / tiles_x) * 16, w0 = (blockIdx.x % tiles_x) * 16;
  const int ci0 = blockIdx.y * CIG;
  const int n = blockIdx.z;
  // first output row/col any pixel of this tile can touch (floor division)
  const int a0 = h0 + PAD - (KS - 1), b0 = w0 + PAD - (KS - 1);
  const int oh_lo = a0 >= 0 ? a0 / ST : -((-a0 + ST - 1) / ST);
  const int ow_lo = b0 >= 0 ? b0 / ST : -((-b0 + ST - 1) / ST);
  const float* dyp = dY + (long)n * dy_ns;
  float acc[CIG];
#pragma unroll
  for (int g = 0; g < CIG; ++g) acc[g] = 0.f;
  const int h = h0 + ty, x = w0 + tx;

  for (int c0 = 0; c0 < Cout; c0 += CK) {
    for (int i = tid; i < CK * DT * DT; i += 256) {
      const int c = i / (DT * DT), r = (i / DT) % DT, s = i % DT;
      const int oh = oh_lo + r, ow = ow_lo + s;
      float v = 0.f;
      if (c0 + c < Cout && oh >= 0 && oh < Ho && ow >= 0 && ow < Wo)
        v = dyp[((long)(c0 + c) * Ho + oh) * Wo + ow];
      dy_s[c][r][s] = v;
    }
    for (int i = tid; i < CK * KS * KS * CIG; i += 256) {
      const int g = i % CIG, t = (i / CIG) % (KS * KS), c = i / (CIG * KS * KS);
      float v = 0.f;
      if (c0 + c < Cout && ci0 + g < ci_count)
        v = w[((long)(c0 + c) * Cin + ci0 + g) * (KS * KS) + t];
      w_s[c][t][g] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int c = 0; c < CK; ++c) {
#pragma unroll
      for (int r = 0; r < KS; ++r) {
        const int hh = h + PAD - r;
        if (ST > 1 && (hh & (ST - 1))) continue;
        const int oh = (ST > 1 ? hh / ST : hh) - oh_lo;
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          const int ww = x + PAD - s;
          if (ST > 1 && (ww & (ST - 1))) continue;
          const int ow = (ST > 1 ? ww / ST : ww) - ow_lo;
          const float v = dy_s[c][oh][ow];
#pragma unroll
          for (int g = 0; g < CIG; ++g) acc[g] = fmaf(v, w_s[c][r * KS + s][g], acc[g]);
        }
      }
    }
    __syncthreads();
  }
  if (h >= H || x >= W) return;
#pragma unroll
  for (int g = 0; g < CIG; ++g) {
    const int ci = ci0 + g;
    if (ci >= ci_count) break;
    const long o = (long)n * dx_ns + ((long)ci * H + h) * W + x;
    dX[o] = accumulate ? dX[o] + acc[g] : acc[g];
  }
}

// ------------------------------------------------------------------ weight gradient
// Thread = one (co, ci) pair with KS*KS register accumulators; the CTA walks its
// share of (image, tile) work items with dY / input tiles staged in shared
// memory and a KS x KS sliding register window along each output row.
template <int KS, int ST, int COC, int CIC>
__global__ void __launch_bounds__(COC * CIC)
conv_wgrad_kernel(const float* __restrict__ in, long in_ns, int Cin, int H, int W,
                  const float* __restrict__ dY, long dy_ns, int Cout, int Ho, int Wo,
                  float* __restrict__ dW, float* __restrict__ dbias, int N, int ci_groups) {
  constexpr int PAD = KS / 2;
  constexpr int TO = 16 / ST;                    // output tile side
  constexpr int IT = (TO - 1) * ST + KS;         // input tile side
  constexpr int ICH = (IT * IT) | 1;             // odd channel stride -> conflict-free
  constexpr int DCH = (TO * TO) | 1;
  constexpr int NT = COC * CIC;
  extern __shared__ float smem[];
  float* in_s = smem;                            // [CIC][ICH]
  float* dy_s = smem + CIC * ICH;                // [COC][DCH]
  const int tid = threadIdx.x;
  const int lco = tid / CIC, lci = tid % CIC;
  const int co0 = (blockIdx.y / ci_groups) * COC, ci0 = (blockIdx.y % ci_groups) * CIC;
  const int co = co0 + lco, ci = ci0 + lci;
  const int tiles_x = (Wo + TO - 1) / TO, tiles_y = (Ho + TO - 1) / TO;
  const long items = (long)N * tiles_x * tiles_y;
  float acc[KS][KS];
#pragma unroll
  for (int r = 0; r < KS; ++r)
#pragma unroll
    for (int s = 0; s < KS; ++s) acc[r][s] = 0.f;
  float bsum = 0.f;

  for (long item = blockIdx.x; item < items; item += gridDim.x) {
    const int n = (int)(item / (tiles_x * tiles_y));
    const int t = (int)(item % (tiles_x * tiles_y));
    const int oh0 = (t / tiles_x) * TO, ow0 = (t % tiles_x) * TO;
    const int ih0 = oh0 * ST - PAD, iw0 = ow0 * ST - PAD;
    const float* inp = in + (long)n * in_ns;
    const float* dyp = dY + (long)n * dy_ns;
    for (int i = tid; i < CIC * IT * IT; i += NT) {
      const int c = i / (IT * IT), r = (i / IT) % IT, s = i % IT;
      const int ih = ih0 + r, iw = iw0 + s;
      float v = 0.f;
      if (ci0 + c < Cin && ih >= 0 && ih < H && iw >= 0 && iw < W)
        v = inp[((long)(ci0 + c) * H + ih) * W + iw];
      in_s[c * ICH + r * IT + s] = v;
    }
    for (int i = tid; i < COC * TO * TO; i += NT) {
      const int c = i / (TO * TO), r = (i / TO) % TO, s = i % TO;
      const int oh = oh0 + r, ow = ow0 + s;
      float v = 0.f;
      if (co0 + c < Cout && oh < Ho && ow < Wo) v = dyp[((long)(co0 + c) * Ho + oh) * Wo + ow];
      dy_s[c * DCH + r * TO + s] = v;
    }
    __syncthreads();
    const float* ip = in_s + lci * ICH;
    const float* dp = dy_s + lco * DCH;
#pragma unroll 1
    for (int y = 0; y < TO; ++y) {
      float win[KS][KS];
#pragma unroll
      for (int x = 0; x < TO; ++x) {
        if (x == 0) {
#pragma unroll
          for (int r = 0; r < KS; ++r)
#pragma unroll
            for (int s = 0; s < KS; ++s) win[r][s] = ip[(y * ST + r) * IT + s];
        } else {
#pragma unroll
          for (int r = 0; r < KS; ++r) {
#pragma unroll
            for (int s = 0; s + ST < KS; ++s) win[r][s] = win[r][s + ST];
#pragma unroll
            for (int s = (KS - ST > 0 ? KS - ST : 0); s < KS; ++s)
              win[r][s] = ip[(y * ST + r) * IT + x * ST + s];
          }
        }
        const float d = dp[y * TO + x];
        bsum += d;
#pragma unroll
        for (int r = 0; r < KS; ++r)
#pragma unroll
          for (int s = 0; s < KS; ++s) acc[r][s] = fmaf(d, win[r][s], acc[r][s]);
      }
    }
    __syncthreads();
  }
  if (co < Cout && ci < Cin) {
    float* wp = dW + ((long)co * Cin + ci) * (KS * KS);
#pragma unroll
    for (int r = 0; r < KS; ++r)
#pragma unroll
      for (int s = 0; s < KS; ++s) atomicAdd(wp + r * KS + s, acc[r][s]);
    if (dbias && ci == 0) atomicAdd(dbias + co, bsum);
  }
}

// ------------------------------------------------------------------ planar elementwise / BatchNorm
// dPre = dA * mask[n][c] * (A > 0 ? 1 : slope)     (LeakyReLU + Dropout2d backward)
__global__ void act_bwd_planar_kernel(const float* __restrict__ dA, long da_ns,
                                      const float* __restrict__ A, long a_ns,
                                      const float* __restrict__ mask, float slope, int C, long HW,
                                      long total, float* __restrict__ dPre, long dp_ns) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long chw = (long)C * HW;
    const long n = i / chw, r = i % chw;
    const int c = (int)(r / HW);
    float g = dA[n * da_ns + r];
    const float a = A[n * a_ns + r];
    g *= (a > 0.f) ? 1.f : slope;
    if (mask) g *= mask[n * C + c];
    dPre[n * dp_ns + r] = g;
  }
}

// per-channel sum / sum of squares over N x HW of a planar tensor
__global__ void __launch_bounds__(256)
bn_stats_planar_kernel(const float* __restrict__ X, long x_ns, int C, long HW, int N,
                       double* __restrict__ sums) {
  const int c = blockIdx.x;
  double s0 = 0, s1 = 0;
  for (int n = blockIdx.y; n < N; n += gridDim.y) {
    const float* p = X + (long)n * x_ns + (long)c * HW;
    float a = 0.f, b = 0.f;
    for (long i = threadIdx.x; i < HW; i += blockDim.x) {
      const float v = p[i];
      a += v;
      b += v * v;
    }
    s0 += a;
    s1 += b;
  }
  __shared__ double r0[8], r1[8];
  s0 = warp_sum_d(s0);
  s1 = warp_sum_d(s1);
  if ((threadIdx.x & 31) == 0) { r0[threadIdx.x >> 5] = s0; r1[threadIdx.x >> 5] = s1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) { s0 += r0[i]; s1 += r1[i]; }
    atomicAdd(sums + c, s0);
    atomicAdd(sums + C + c, s1);
  }
}

// out = [relu](X*scale + shift), planar
__global__ void bn_apply_planar_kernel(const float* __restrict__ X, long x_ns,
                                       const float* __restrict__ scale,
                                       const float* __restrict__ shift, int C, long HW, long total,
                                       int relu, float* __restrict__ out, long o_ns) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long chw = (long)C * HW;
    const long n = i / chw, r = i % chw;
    const int c = (int)(r / HW);
    float v = fmaf(X[n * x_ns + r], scale[c], shift[c]);
    if (relu) v = fmaxf(v, 0.f);
    out[n * o_ns + r] = v;
  }
}

// sums2[0][c] = sum dZ, sums2[1][c] = sum dZ * xhat
__global__ void __launch_bounds__(256)
bn_bwd_reduce_planar_kernel(const float* __restrict__ dZ, long dz_ns, const float* __restrict__ X,
                            long x_ns, const float* __restrict__ mean,
                            const float* __restrict__ invstd, int C, long HW, int N,
                            double* __restrict__ sums2) {
  const int c = blockIdx.x;
  const float m = mean[c], is = invstd[c];
  double s0 = 0, s1 = 0;
  for (int n = blockIdx.y; n < N; n += gridDim.y) {
    const float* g = dZ + (long)n * dz_ns + (long)c * HW;
    const float* p = X + (long)n * x_ns + (long)c * HW;
    float a = 0.f, b = 0.f;
    for (long i = threadIdx.x; i < HW; i += blockDim.x) {
      const float d = g[i];
      a += d;
      b += d * (p[i] - m) * is;
    }
    s0 += a;
    s1 += b;
  }
  __shared__ double r0[8], r1[8];
  s0 = warp_sum_d(s0);
  s1 = warp_sum_d(s1);
  if ((threadIdx.x & 31) == 0) { r0[threadIdx.x >> 5] = s0; r1[threadIdx.x >> 5] = s1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) { s0 += r0[i]; s1 += r1[i]; }
    atomicAdd(sums2 + c, s0);
    atomicAdd(sums2 + C + c, s1);
  }
}

// dX = gamma*invstd*(dZ - mean(dZ) - xhat*mean(dZ*xhat)); block 0 writes dgamma/dbeta
__global__ void bn_bwd_apply_planar_kernel(const float* __restrict__ dZ, long dz_ns,
                                           const float* __restrict__ X, long x_ns,
                                           const float* __restrict__ mean,
                                           const float* __restrict__ invstd,
                                           const float* __restrict__ gamma,
                                           const double* __restrict__ sums2, double count, int C,
                                           long HW, long total, float* __restrict__ dX, long dx_ns,
                                           float* __restrict__ dgamma, float* __restrict__ dbeta) {
  if (blockIdx.x == 0 && dgamma) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      dbeta[c] = (float)sums2[c];
      dgamma[c] = (float)sums2[C + c];
    }
  }
  const float inv_count = (float)(1.0 / count);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long chw = (long)C * HW;
    const long n = i / chw, r = i % chw;
    const int c = (int)(r / HW);
    const float is = invstd[c];
    const float xhat = (X[n * x_ns + r] - mean[c]) * is;
    const float m1 = (float)sums2[c] * inv_count, m2 = (float)sums2[C + c] * inv_count;
    dX[n * dx_ns + r] = gamma[c] * is * (dZ[n * dz_ns + r] - m1 - xhat * m2);
  }
}

// dst[n][0:count] = src[n][0:count] with independent per-image strides
__global__ void copy_planar_kernel(const float* __restrict__ src, long s_ns, float* __restrict__ dst,
                                   long d_ns, long count, long total) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long n = i / count, r = i % count;
    dst[n * d_ns + r] = src[n * s_ns + r];
  }
}

static int ew_grid(long total) {
  long b = cdiv(total, 256);
  if (b > 148L * 16) b = 148L * 16;
  if (b < 1) b = 1;
  return (int)b;
}

template <int KS, int ST, int COC, int CIC>
static int launch_wgrad(const float* in, long in_ns, int Cin, int H, int W, const float* dY,
                        long dy_ns, int Cout, int Ho, int Wo, float* dW, float* dbias, int N,
                        cudaStream_t st) {
  constexpr int TO = 16 / ST, IT = (TO - 1) * ST + KS;
  constexpr int ICH = (IT * IT) | 1, DCH = (TO * TO) | 1;
  const int smem = (CIC * ICH + COC * DCH) * (int)sizeof(float);
  auto kern = conv_wgrad_kernel<KS, ST, COC, CIC>;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return dmc_check_launch("conv_wgrad smem attribute");
    attr = true;
  }
  const int co_groups = (int)cdiv(Cout, COC), ci_groups = (int)cdiv(Cin, CIC);
  const long items = (long)N * cdiv(Wo, TO) * cdiv(Ho, TO);
  long gx = cdiv(148L * 4, (long)co_groups * ci_groups);
  if (gx < 1) gx = 1;
  if (gx > items) gx = items;
  dim3 grid((unsigned)gx, (unsigned)(co_groups * ci_groups));
  kern<<<grid, COC * CIC, smem, st>>>(in, in_ns, Cin, H, W, dY, dy_ns, Cout, Ho, Wo, dW, dbias, N,
                                      ci_groups);
  return dmc_check_launch("conv_wgrad_kernel");
}

}  // namespace dmc

using namespace dmc;
#define ST_(s) reinterpret_cast<cudaStream_t>(s)

// out[n][co][Ho][Wo] = mask * lrelu(bias + conv_{ks x ks, stride, pad ks/2}(in, w)) + add (+ out)
extern "C" int dmc_conv_fwd(const float* in, long in_ns, int Cin, int H, int W, const float* w,
                            const float* bias, int Cout, int ks, int stride, float* out,
                            long out_ns, float slope, const float* mask, const float* add,
                            long add_ns, int accumulate, int N, void* stream) {
  DMC_REQUIRE((ks == 3 && (stride == 1 || stride == 2)) || (ks == 7 && stride == 2),
              "conv_fwd: unsupported ks=%d stride=%d", ks, stride);
  const int pad = ks / 2;
  const int Ho = (H + 2 * pad - ks) / stride + 1, Wo = (W + 2 * pad - ks) / stride + 1;
  const int tiles_x = (int)cdiv(Wo, 16), tiles_y = (int)cdiv(Ho, 16);
  cudaStream_t st = ST_(stream);
  if (ks == 3 && stride == 1) {
    dim3 grid(tiles_x * tiles_y, (unsigned)cdiv(Cout, 8), N);
    conv_fwd_kernel<3, 1, 8, 8><<<grid, 256, 0, st>>>(in, in_ns, Cin, H, W, w, bias, Cout, out,
                                                      out_ns, Ho, Wo, slope, mask, add, add_ns,
                                                      accumulate, tiles_x);
  } else if (ks == 3) {
    dim3 grid(tiles_x * tiles_y, (unsigned)cdiv(Cout, 8), N);
    conv_fwd_kernel<3, 2, 8, 8><<<grid, 256, 0, st>>>(in, in_ns, Cin, H, W, w, bias, Cout, out,
                                                      out_ns, Ho, Wo, slope, mask, add, add_ns,
                                                      accumulate, tiles_x);
  } else {
    dim3 grid(tiles_x * tiles_y, (unsigned)cdiv(Cout, 16), N);
    conv_fwd_kernel<7, 2, 16, 2><<<grid, 256, 0, st>>>(in, in_ns, Cin, H, W, w, bias, Cout, out,
                                                       out_ns, Ho, Wo, slope, mask, add, add_ns,
                                                       accumulate, tiles_x);
  }
  return dmc_check_launch("conv_fwd_kernel");
}

// dX[n][ci][H][W] (+)= conv_transpose(dY[n][co][Ho][Wo], w[co][ci][ks][ks])
// Only input channels [0, ci_count) are produced (the weight keeps its full Cin stride).
extern "C" int dmc_conv_dgrad(const float* dY, long dy_ns, int Cout, const float* w, int Cin,
                              int ci_count, int ks, int stride, float* dX, long dx_ns, int H, int W,
                              int accumulate, int N, void* stream) {
  DMC_REQUIRE(ci_count >= 1 && ci_count <= Cin, "conv_dgrad: ci_count=%d Cin=%d", ci_count, Cin);
  DMC_REQUIRE((ks == 3 && (stride == 1 || stride == 2)) || (ks == 7 && stride == 2),
              "conv_dgrad: unsupported ks=%d stride=%d", ks, stride);
  const int pad = ks / 2;
  const int Ho = (H + 2 * pad - ks) / stride + 1, Wo = (W + 2 * pad - ks) / stride + 1;
  const int tiles_x = (int)cdiv(W, 16), tiles_y = (int)cdiv(H, 16);
  cudaStream_t st = ST_(stream);
  if (ks == 3 && stride == 1) {
    dim3 grid(tiles_x * tiles_y, (unsigned)cdiv(ci_count, 8), N);
    conv_dgrad_kernel<3, 1, 8, 8><<<grid, 256, 0, st>>>(dY, dy_ns, Cout, Ho, Wo, w, Cin, ci_count, dX,
                                                        dx_ns, H, W, accumulate, tiles_x);
  } else if (ks == 3) {
    dim3 grid(tiles_x * tiles_y, (unsigned)cdiv(ci_count, 8), N);
    conv_dgrad_kernel<3, 2, 8, 8><<<grid, 256, 0, st>>>(dY, dy_ns, Cout, Ho, Wo, w, Cin, ci_count, dX,
                                                        dx_ns, H, W, accumulate, tiles_x);
  } else {
    dim3 grid(tiles_x * tiles_y, (unsigned)cdiv(ci_count, 2), N);
    conv_dgrad_kernel<7, 2, 2, 8><<<grid, 256, 0, st>>>(dY, dy_ns, Cout, Ho, Wo, w, Cin, ci_count, dX,
                                                        dx_ns, H, W, accumulate, tiles_x);
  }
  return dmc_check_launch("conv_dgrad_kernel");
}

// dW[co][ci][ks][ks] += ..., dbias[co] += ...   (caller zeroes both)
extern "C" int dmc_conv_wgrad(const float* in, long in_ns, int Cin, int H, int W, const float* dY,
                              long dy_ns, int Cout, int ks, int stride, float* dW, float* dbias,
                              int N, void* stream) {
  DMC_REQUIRE((ks == 3 && (stride == 1 || stride == 2)) || (ks == 7 && stride == 2),
              "conv_wgrad: unsupported ks=%d stride=%d", ks, stride);
  const int pad = ks / 2;
  const int Ho = (H + 2 * pad - ks) / stride + 1, Wo = (W + 2 * pad - ks) / stride + 1;
  cudaStream_t st = ST_(stream);
  if (ks == 3 && stride == 1)
    return launch_wgrad<3, 1, 8, 32>(in, in_ns, Cin, H, W, dY, dy_ns, Cout, Ho, Wo, dW, dbias, N, st);
  if (ks == 3)
    return launch_wgrad<3, 2, 8, 32>(in, in_ns, Cin, H, W, dY, dy_ns, Cout, Ho, Wo, dW, dbias, N, st);
  return launch_wgrad<7, 2, 64, 2>(in, in_ns, Cin, H, W, dY, dy_ns, Cout, Ho, Wo, dW, dbias, N, st);
}

extern "C" int dmc_act_bwd_planar(const float* dA, long da_ns, const float* A, long a_ns,
                                  const float* mask, float slope, int C, long HW, int N, float* dPre,
                                  long dp_ns, void* stream) {
  const long total = (long)N * C * HW;
  act_bwd_planar_kernel<<<ew_grid(total), 256, 0, ST_(stream)>>>(dA, da_ns, A, a_ns, mask, slope, C,
                                                                 HW, total, dPre, dp_ns);
  return dmc_check_launch("act_bwd_planar_kernel");
}

extern "C" int dmc_bn_stats_planar(const float* X, long x_ns, int C, long HW, int N, double* sums,
                                   void* stream) {
  if (cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, ST_(stream)) != cudaSuccess)
    return dmc_check_launch("bn_stats_planar memset");
  int gy = (int)cdiv(148L * 8, C);
  if (gy > N) gy = N;
  if (gy < 1) gy = 1;
  bn_stats_planar_kernel<<<dim3(C, gy), 256, 0, ST_(stream)>>>(X, x_ns, C, HW, N, sums);
  return dmc_check_launch("bn_stats_planar_kernel");
}

extern "C" int dmc_bn_apply_planar(const float* X, long x_ns, const float* scale, const float* shift,
                                   int C, long HW, int N, int relu, float* out, long o_ns,
                                   void* stream) {
  const long total = (long)N * C * HW;
  bn_apply_planar_kernel<<<ew_grid(total), 256, 0, ST_(stream)>>>(X, x_ns, scale, shift, C, HW,
                                                                  total, relu, out, o_ns);
  return dmc_check_launch("bn_apply_planar_kernel");
}

extern "C" int dmc_bn_bwd_reduce_planar(const float* dZ, long dz_ns, const float* X, long x_ns,
                                        const float* mean, const float* invstd, int C, long HW,
                                        int N, double* sums2, void* stream) {
  if (cudaMemsetAsync(sums2, 0, sizeof(double) * 2 * C, ST_(stream)) != cudaSuccess)
    return dmc_check_launch("bn_bwd_reduce_planar memset");
  int gy = (int)cdiv(148L * 8, C);
  if (gy > N) gy = N;
  if (gy < 1) gy = 1;
  bn_bwd_reduce_planar_kernel<<<dim3(C, gy), 256, 0, ST_(stream)>>>(dZ, dz_ns, X, x_ns, mean, invstd,
                                                                    C, HW, N, sums2);
  return dmc_check_launch("bn_bwd_reduce_planar_kernel");
}

extern "C" int dmc_bn_bwd_apply_planar(const float* dZ, long dz_ns, const float* X, long x_ns,
                                       const float* mean, const float* invstd, const float* gamma,
                                       const double* sums2, double count, int C, long HW, int N,
                                       float* dX, long dx_ns, float* dgamma, float* dbeta,
                                       void* stream) {
  const long total = (long)N * C * HW;
  bn_bwd_apply_planar_kernel<<<ew_grid(total), 256, 0, ST_(stream)>>>(
      dZ, dz_ns, X, x_ns, mean, invstd, gamma, sums2, count, C, HW, total, dX, dx_ns, dgamma, dbeta);
  return dmc_check_launch("bn_bwd_apply_planar_kernel");
}

extern "C" int dmc_copy_planar(const float* src, long s_ns, float* dst, long d_ns, long count, int N,
                               void* stream) {
  const long total = count * N;
  copy_planar_kernel<<<ew_grid(total), 256, 0, ST_(stream)>>>(src, s_ns, dst, d_ns, count, total);
  return dmc_check_launch("copy_planar_kernel");
}
