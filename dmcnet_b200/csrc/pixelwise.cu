// Per-pixel / per-channel kernels of the classifier over the padded pixel-major
// layout [frames][Hp][Wp][C] (C a multiple of 4, channel fastest):
// weight repacking for the tap GEMMs, train-mode BatchNorm forward (batch
// statistics, running-stat update, apply + residual + ReLU -> bf16 hi/lo
// planes), BatchNorm backward (two reductions + apply), the space-to-depth
// "phase" layouts used by stride-2 convolutions, and global average pooling.
// All of them are HBM-bound streaming kernels: 16-byte accesses, channel index
// on the fastest thread dimension, grid sized in multiples of the SM count.
#include "common.cuh"

namespace dmc {

static int grid_for(long work_items, int per_block) {
  long b = cdiv(work_items, per_block);
  const long cap = 148L * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

struct __align__(8) bf16x4 {
  bf16 v[4];
};

__device__ __forceinline__ void store_split4(bf16* hi, bf16* lo, long off, const float (&x)[4]) {
  bf16x4 h, l;
#pragma unroll
  for (int i = 0; i < 4; ++i) split_bf16(x[i], h.v[i], l.v[i]);
  *reinterpret_cast<bf16x4*>(hi + off) = h;
  if (lo) *reinterpret_cast<bf16x4*>(lo + off) = l;    // lo == NULL: bf16 (round-to-nearest) only
}
__device__ __forceinline__ void load_join4(const bf16* hi, const bf16* lo, long off, float (&x)[4]) {
  const bf16x4 h = *reinterpret_cast<const bf16x4*>(hi + off);
  const bf16x4 l = *reinterpret_cast<const bf16x4*>(lo + off);
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = join_bf16(h.v[i], l.v[i]);
}

// ------------------------------------------------------------------ weights
// OIHW fp32 -> [tap][Cout][Cin] and [tap][Cin][Cout] bf16 hi/lo (GEMM B operands
// of fprop and dgrad).  code/dmcnet/model.py:305 (torchvision conv weights).
__global__ void weight_prep_kernel(const float* __restrict__ w, int Cout, int Cin, int taps,
                                   bf16* __restrict__ Wh, bf16* __restrict__ Wl,
                                   bf16* __restrict__ Th, bf16* __restrict__ Tl) {
  const long n = (long)Cout * Cin * taps;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    // i enumerates the OUTPUT [tap][co][ci] so the bf16 writes coalesce
    const int ci = (int)(i % Cin);
    const int co = (int)((i / Cin) % Cout);
    const int t = (int)(i / ((long)Cin * Cout));
    const float v = w[((long)co * Cin + ci) * taps + t];
    bf16 h, l;
    split_bf16(v, h, l);
    Wh[i] = h;
    Wl[i] = l;
    if (Th) {
      const long j = ((long)t * Cin + ci) * Cout + co;
      Th[j] = h;
      Tl[j] = l;
    }
  }
}

// All conv weights of the classifier in ONE launch: a chunk table maps 32 x 32 (co, ci) tiles
// to (tensor, destination planes).  A CTA reads its tile's taps as 32 contiguous runs of the
// OIHW tensor into shared memory and writes both operand layouts with the fastest index on
// the lanes ([tap][co][ci]: ci, [tap][ci][co]: co), so every global access is a full run.
struct PrepChunk {
  long long w_off;             // element offset of the OIHW tensor inside the parameter bucket
  long long Wh, Wl, Th, Tl;    // destination addresses (bf16), Th/Tl may be 0
  int Cout, Cin, taps, start;  // start = tile index: (co tile) * (Cin / 32) + (ci tile)
};

__global__ void __launch_bounds__(256)
weight_prep_multi_kernel(const float* __restrict__ params, const PrepChunk* __restrict__ chunks) {
  __shared__ float tile[32][32 * 9 + 1];
  const PrepChunk ch = chunks[blockIdx.x];
  const int taps = ch.taps, Cin = ch.Cin, Cout = ch.Cout;
  const int tiles_ci = Cin / 32;
  const int co0 = (ch.start / tiles_ci) * 32, ci0 = (ch.start % tiles_ci) * 32;
  const float* w = params + ch.w_off;
  bf16* Wh = reinterpret_cast<bf16*>(ch.Wh);
  bf16* Wl = reinterpret_cast<bf16*>(ch.Wl);
  bf16* Th = reinterpret_cast<bf16*>(ch.Th);
  bf16* Tl = reinterpret_cast<bf16*>(ch.Tl);
  const int run = 32 * taps;                       // contiguous floats per output channel
  for (int i = threadIdx.x; i < 32 * run; i += 256) {
    const int co = i / run, r = i - co * run;
    tile[co][r] = w[((long)(co0 + co) * Cin + ci0) * taps + r];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < taps * 1024; i += 256) {
    const int x = i & 31, y = (i >> 5) & 31, t = i >> 10;
    {   // [tap][co][ci]: x = ci
      bf16 h, l;
      split_bf16(tile[y][x * taps + t], h, l);
      const long o = ((long)t * Cout + co0 + y) * Cin + ci0 + x;
      Wh[o] = h;
      Wl[o] = l;
    }
    if (Th) {   // [tap][ci][co]: x = co
      bf16 h, l;
      split_bf16(tile[x][y * taps + t], h, l);
      const long o = ((long)t * Cin + ci0 + y) * Cout + co0 + x;
      Th[o] = h;
      Tl[o] = l;
    }
  }
}

// [tap][Cout][Cin] fp32 (wgrad accumulator) -> OIHW gradient.
__global__ void wgrad_unpack_kernel(const float* __restrict__ dWs, float* __restrict__ g, int Cout,
                                    int Cin, int taps) {
  const long n = (long)Cout * Cin * taps;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int t = (int)(i % taps);
    const int ci = (int)((i / taps) % Cin);
    const int co = (int)(i / ((long)taps * Cin));
    g[i] = dWs[((long)t * Cout + co) * Cin + ci];
  }
}

// ------------------------------------------------------------------ channel reductions
// Each thread owns 4 consecutive channels (c4) and strides over rows; partials
// are combined in shared memory in double and added to global double sums.
template <class F>
__device__ __forceinline__ void channel_reduce2(long P, int C, double* __restrict__ out0,
                                                double* __restrict__ out1, F&& f) {
  const int c4n = C / 4;                       // threads per row
  const int rows_per_iter = blockDim.x / c4n;  // blockDim.x is a multiple of c4n
  const int c4 = threadIdx.x % c4n;
  const int r = threadIdx.x / c4n;
  float s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
  double d0[4] = {0, 0, 0, 0}, d1[4] = {0, 0, 0, 0};
  const unsigned step = gridDim.x * rows_per_iter;
  const unsigned uP = (unsigned)P;
  unsigned q = blockIdx.x * rows_per_iter + r;
  int n = 0;
  // four independent rows per trip keep enough loads in flight to stream HBM
  for (; q + 3u * step < uP; q += 4u * step) {
    f(q, c4 * 4, s0, s1);
    f(q + step, c4 * 4, s0, s1);
    f(q + 2u * step, c4 * 4, s0, s1);
    f(q + 3u * step, c4 * 4, s0, s1);
    if (++n == 16) {                           // flush fp32 partials to double regularly
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        d0[i] += s0[i]; d1[i] += s1[i]; s0[i] = 0.f; s1[i] = 0.f;
      }
      n = 0;
    }
  }
  for (; q < uP; q += step) f(q, c4 * 4, s0, s1);
#pragma unroll
  for (int i = 0; i < 4; ++i) { d0[i] += s0[i]; d1[i] += s1[i]; }
  extern __shared__ double red[];              // [2][blockDim.x][4]
  double* r0 = red;
  double* r1 = red + blockDim.x * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    r0[threadIdx.x * 4 + i] = d0[i];
    r1[threadIdx.x * 4 + i] = d1[i];
  }
  __syncthreads();
  if (threadIdx.x < c4n) {
    for (int rr = 1; rr < rows_per_iter; ++rr) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        d0[i] += r0[(rr * c4n + threadIdx.x) * 4 + i];
        d1[i] += r1[(rr * c4n + threadIdx.x) * 4 + i];
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      atomicAdd(out0 + threadIdx.x * 4 + i, d0[i]);
      atomicAdd(out1 + threadIdx.x * 4 + i, d1[i]);
    }
  }
}

// sums[0][c] = sum_q Y[q][c], sums[1][c] = sum_q Y[q][c]^2  (border rows are zero).
__global__ void __launch_bounds__(256, 4)
bn_stats_kernel(const float* __restrict__ Y, long P, int C, double* __restrict__ sums) {
  channel_reduce2(P, C, sums, sums + C, [&](unsigned q, int c, float (&s0)[4], float (&s1)[4]) {
    const float4 v = *reinterpret_cast<const float4*>(Y + (long)q * C + c);
    s0[0] += v.x; s0[1] += v.y; s0[2] += v.z; s0[3] += v.w;
    s1[0] += v.x * v.x; s1[1] += v.y * v.y; s1[2] += v.z * v.z; s1[3] += v.w * v.w;
  });
}

// nn.BatchNorm2d training-mode statistics (torch semantics: biased variance for
// normalisation, unbiased for running_var, momentum update, num_batches_tracked).
__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   long long* __restrict__ nbt, float momentum, float eps, int C,
                                   float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mean_out, float* __restrict__ invstd_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && nbt) *nbt += 1;
  if (c >= C) return;
  const double mean = sums[c] / count;
  double var = sums[C + c] / count - mean * mean;
  if (var < 0) var = 0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = gamma[c] * invstd;
  scale[c] = sc;
  shift[c] = beta[c] - (float)mean * sc;
  mean_out[c] = (float)mean;
  invstd_out[c] = invstd;
  if (running_mean) {
    const double unbiased = count > 1 ? var * count / (count - 1) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// Eval-mode BatchNorm: scale/shift from the running statistics.
__global__ void bn_eval_coeffs_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ running_mean,
                                      const float* __restrict__ running_var, float eps, int C,
                                      float* __restrict__ scale, float* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invstd = 1.0f / sqrtf(running_var[c] + eps);
  const float sc = gamma[c] * invstd;
  scale[c] = sc;
  shift[c] = beta[c] - running_mean[c] * sc;
}

// out = [relu]( Y*scale + shift  [+ residual] ), zero on the padding ring,
// written as bf16 hi/lo planes.  residual is either a hi/lo activation or a
// second raw conv output with its own BN coefficients (downsample branch).
__global__ void __launch_bounds__(256)
bn_apply_kernel(const float* __restrict__ Y, const float* __restrict__ scale,
                const float* __restrict__ shift, long P, int C, int Hp, int Wp, int relu,
                const bf16* __restrict__ res_hi, const bf16* __restrict__ res_lo,
                const float* __restrict__ resY, const float* __restrict__ res_scale,
                const float* __restrict__ res_shift, bf16* __restrict__ out_hi,
                bf16* __restrict__ out_lo, float slope, int ring) {
  const FastDiv fd((unsigned)(C / 4));
  const unsigned n = (unsigned)(P * (C / 4));
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned q = fd.div(i);
    const int c = (int)fd.mod(i) * 4;
    const long off = (long)q * C + c;
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    if (interior_r(q, Hp, Wp, ring)) {
      const float4 y = *reinterpret_cast<const float4*>(Y + off);
      const float4 sc = *reinterpret_cast<const float4*>(scale + c);
      const float4 sh = *reinterpret_cast<const float4*>(shift + c);
      o[0] = fmaf(y.x, sc.x, sh.x); o[1] = fmaf(y.y, sc.y, sh.y);
      o[2] = fmaf(y.z, sc.z, sh.z); o[3] = fmaf(y.w, sc.w, sh.w);
      if (res_hi) {
        float r[4];
        load_join4(res_hi, res_lo, off, r);
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] += r[k];
      } else if (resY) {
        const float4 y2 = *reinterpret_cast<const float4*>(resY + off);
        const float4 s2 = *reinterpret_cast<const float4*>(res_scale + c);
        const float4 h2 = *reinterpret_cast<const float4*>(res_shift + c);
        o[0] += fmaf(y2.x, s2.x, h2.x); o[1] += fmaf(y2.y, s2.y, h2.y);
        o[2] += fmaf(y2.z, s2.z, h2.z); o[3] += fmaf(y2.w, s2.w, h2.w);
      }
      if (relu) {                       // slope 0: ReLU; otherwise LeakyReLU(slope)
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = o[k] > 0.f ? o[k] : o[k] * slope;
      }
    }
    store_split4(out_hi, out_lo, off, o);
  }
}

// ------------------------------------------------------------------ BatchNorm backward
// dz = (g_a [+ g_b]) * [act > 0];  sums2[0][c] = sum dz, sums2[1][c] = sum dz * xhat.
__device__ __forceinline__ void load_dz(const float* g_a, const float* g_b, const bf16* act_hi,
                                        long off, float (&dz)[4]) {
  const float4 a = *reinterpret_cast<const float4*>(g_a + off);
  dz[0] = a.x; dz[1] = a.y; dz[2] = a.z; dz[3] = a.w;
  if (g_b) {
    const float4 b = *reinterpret_cast<const float4*>(g_b + off);
    dz[0] += b.x; dz[1] += b.y; dz[2] += b.z; dz[3] += b.w;
  }
  if (act_hi) {
    const bf16x4 h = *reinterpret_cast<const bf16x4*>(act_hi + off);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (!(__bfloat162float(h.v[k]) > 0.f)) dz[k] = 0.f;
  }
}

__global__ void __launch_bounds__(256, 4)
bn_bwd_reduce_kernel(const float* __restrict__ g_a, const float* __restrict__ g_b,
                     const bf16* __restrict__ act_hi, const float* __restrict__ Y,
                     const float* __restrict__ mean, const float* __restrict__ invstd, long P, int C,
                     int Hp, int Wp, double* __restrict__ sums2) {
  channel_reduce2(P, C, sums2, sums2 + C, [&](unsigned q, int c, float (&s0)[4], float (&s1)[4]) {
    if (!interior(q, Hp, Wp)) return;
    const long off = (long)q * C + c;
    float dz[4];
    load_dz(g_a, g_b, act_hi, off, dz);
    const float4 y = *reinterpret_cast<const float4*>(Y + off);
    const float4 m = *reinterpret_cast<const float4*>(mean + c);
    const float4 is = *reinterpret_cast<const float4*>(invstd + c);
    s0[0] += dz[0]; s0[1] += dz[1]; s0[2] += dz[2]; s0[3] += dz[3];
    s1[0] += dz[0] * (y.x - m.x) * is.x; s1[1] += dz[1] * (y.y - m.y) * is.y;
    s1[2] += dz[2] * (y.z - m.z) * is.z; s1[3] += dz[3] * (y.w - m.w) * is.w;
  });
}

// dY = gamma*invstd * (dz - mean(dz) - xhat*mean(dz*xhat)) as bf16 hi/lo planes
// (zero ring); optionally also dz itself (fp32) for the residual branch; block 0
// writes dgamma / dbeta.
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ g_a, const float* __restrict__ g_b,
                    const bf16* __restrict__ act_hi, const float* __restrict__ Y,
                    const float* __restrict__ mean, const float* __restrict__ invstd,
                    const float* __restrict__ gamma, const double* __restrict__ sums2, double count,
                    long P, int C, int Hp, int Wp, bf16* __restrict__ G_hi, bf16* __restrict__ G_lo,
                    float* __restrict__ dz_out, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  // dY = A*dz + B*y + K per channel:  A = gamma*invstd, B = -A*invstd*mean(dz*xhat),
  // K = -A*mean(dz) - B*mean
  extern __shared__ __align__(16) float coef[];            // [3][C]
  float* cA = coef;
  float* cB = coef + C;
  float* cK = coef + 2 * C;
  const double inv_count = 1.0 / count;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float a = gamma[c] * invstd[c];
    const float m1 = (float)(sums2[c] * inv_count), m2 = (float)(sums2[C + c] * inv_count);
    const float b = -a * invstd[c] * m2;
    cA[c] = a;
    cB[c] = b;
    cK[c] = -a * m1 - b * mean[c];
    if (blockIdx.x == 0 && dgamma) {
      dbeta[c] = (float)sums2[c];
      dgamma[c] = (float)sums2[C + c];
    }
  }
  __syncthreads();
  const FastDiv fd((unsigned)(C / 4));
  const unsigned n = (unsigned)(P * (C / 4));
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned q = fd.div(i);
    const int c = (int)fd.mod(i) * 4;
    const long off = (long)q * C + c;
    float dy[4] = {0.f, 0.f, 0.f, 0.f};
    float dz[4] = {0.f, 0.f, 0.f, 0.f};
    if (interior(q, Hp, Wp)) {
      load_dz(g_a, g_b, act_hi, off, dz);
      const float4 y = *reinterpret_cast<const float4*>(Y + off);
      const float4 a = *reinterpret_cast<const float4*>(cA + c);
      const float4 b = *reinterpret_cast<const float4*>(cB + c);
      const float4 k = *reinterpret_cast<const float4*>(cK + c);
      dy[0] = fmaf(a.x, dz[0], fmaf(b.x, y.x, k.x));
      dy[1] = fmaf(a.y, dz[1], fmaf(b.y, y.y, k.y));
      dy[2] = fmaf(a.z, dz[2], fmaf(b.z, y.z, k.z));
      dy[3] = fmaf(a.w, dz[3], fmaf(b.w, y.w, k.w));
    }
    store_split4(G_hi, G_lo, off, dy);
    if (dz_out) *reinterpret_cast<float4*>(dz_out + off) = make_float4(dz[0], dz[1], dz[2], dz[3]);
  }
}

// ------------------------------------------------------------------ stride-2 phase layouts
// in  [frames][H+2][W+2][C] hi/lo  ->  out [4][frames][H/2+2][W/2+2][C] hi/lo with
// out[ph*2+pw][n][i+1][j+1] = in[n][2i+ph+1][2j+pw+1]   (unpadded in(2i+ph, 2j+pw)).
__global__ void __launch_bounds__(256)
phase_split_kernel(const bf16* __restrict__ in_hi, const bf16* __restrict__ in_lo, int frames, int H,
                   int W, int C, bf16* __restrict__ out_hi, bf16* __restrict__ out_lo) {
  const int Hq = dmc_padded(H / 2), Wq = dmc_padded(W / 2), c8n = C / 8;
  const long Pq = (long)frames * Hq * Wq;
  const long n = 4 * Pq * c8n;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c8n) * 8;
    long r = i / c8n;
    const int wq = (int)(r % Wq); r /= Wq;
    const int hq = (int)(r % Hq); r /= Hq;
    const int f = (int)(r % frames);
    const int ph = (int)(r / frames);
    uint4 vh = make_uint4(0, 0, 0, 0), vl = make_uint4(0, 0, 0, 0);
    if (hq >= 1 && hq <= H / 2 && wq >= 1 && wq <= W / 2) {
      const int h = 2 * (hq - 1) + (ph >> 1) + 1, w = 2 * (wq - 1) + (ph & 1) + 1;
      const long src = (((long)f * dmc_padded(H) + h) * dmc_padded(W) + w) * C + c;
      vh = *reinterpret_cast<const uint4*>(in_hi + src);
      vl = *reinterpret_cast<const uint4*>(in_lo + src);
    }
    const long dst = ((((long)ph * frames + f) * Hq + hq) * Wq + wq) * C + c;
    *reinterpret_cast<uint4*>(out_hi + dst) = vh;
    *reinterpret_cast<uint4*>(out_lo + dst) = vl;
  }
}

// fp32 gradient in phase layout -> padded full-resolution layout (ring = 0).
__global__ void __launch_bounds__(256)
phase_unsplit_kernel(const float* __restrict__ in, int frames, int H, int W, int C,
                     float* __restrict__ out) {
  const int Hq = dmc_padded(H / 2), Wq = dmc_padded(W / 2), Hp = dmc_padded(H), Wp = dmc_padded(W), c4n = C / 4;
  const long n = (long)frames * Hp * Wp * c4n;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    long r = i / c4n;
    const int wp = (int)(r % Wp); r /= Wp;
    const int hp = (int)(r % Hp);
    const int f = (int)(r / Hp);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (hp >= 1 && hp <= H && wp >= 1 && wp <= W) {
      const int h = hp - 1, w = wp - 1;
      const int ph = (h & 1) * 2 + (w & 1);
      const long src = ((((long)ph * frames + f) * Hq + (h >> 1) + 1) * Wq + (w >> 1) + 1) * C + c;
      v = *reinterpret_cast<const float4*>(in + src);
    }
    *reinterpret_cast<float4*>(out + (((long)f * Hp + hp) * Wp + wp) * C + c) = v;
  }
}

// phase_unsplit fused with the ReLU mask of the activation the gradient belongs to and with the two
// BatchNorm-backward reductions of the conv + BN that produced it (the block in front of a stride-2
// block): out = dz = unsplit(in) * [act > 0], sums2[0][c] += sum dz, sums2[1][c] += sum dz * xhat.
// One pass instead of unsplit + a separate reduction pass over the same tensor.
__global__ void __launch_bounds__(256, 4)
phase_unsplit_reduce_kernel(const float* __restrict__ in, int frames, int H, int W, int C,
                            const bf16* __restrict__ act_hi, const float* __restrict__ Y,
                            const float* __restrict__ mean, const float* __restrict__ invstd,
                            float* __restrict__ out, double* __restrict__ sums2) {
  const int Hq = dmc_padded(H / 2), Wq = dmc_padded(W / 2), Hp = dmc_padded(H), Wp = dmc_padded(W);
  const long P = (long)frames * Hp * Wp;
  channel_reduce2(P, C, sums2, sums2 + C, [&](unsigned q, int c, float (&s0)[4], float (&s1)[4]) {
    const int wp = (int)(q % (unsigned)Wp);
    const unsigned r = q / (unsigned)Wp;
    const int hp = (int)(r % (unsigned)Hp);
    const int f = (int)(r / (unsigned)Hp);
    const long off = (long)q * C + c;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (hp >= 1 && wp >= 1) {
      const int h = hp - 1, w = wp - 1;
      const int ph = (h & 1) * 2 + (w & 1);
      const long src = ((((long)ph * frames + f) * Hq + (h >> 1) + 1) * Wq + (w >> 1) + 1) * C + c;
      v = *reinterpret_cast<const float4*>(in + src);
      const bf16x4 a = *reinterpret_cast<const bf16x4*>(act_hi + off);
      if (!(__bfloat162float(a.v[0]) > 0.f)) v.x = 0.f;
      if (!(__bfloat162float(a.v[1]) > 0.f)) v.y = 0.f;
      if (!(__bfloat162float(a.v[2]) > 0.f)) v.z = 0.f;
      if (!(__bfloat162float(a.v[3]) > 0.f)) v.w = 0.f;
      const float4 y = *reinterpret_cast<const float4*>(Y + off);
      const float4 m = *reinterpret_cast<const float4*>(mean + c);
      const float4 is = *reinterpret_cast<const float4*>(invstd + c);
      s0[0] += v.x; s0[1] += v.y; s0[2] += v.z; s0[3] += v.w;
      s1[0] += v.x * (y.x - m.x) * is.x; s1[1] += v.y * (y.y - m.y) * is.y;
      s1[2] += v.z * (y.z - m.z) * is.z; s1[3] += v.w * (y.w - m.w) * is.w;
    }
    *reinterpret_cast<float4*>(out + off) = v;
  });
}

// ------------------------------------------------------------------ global average pool
__global__ void avgpool_kernel(const bf16* __restrict__ hi, const bf16* __restrict__ lo, int frames,
                               int Hp, int Wp, int C, float* __restrict__ pooled) {
  const int f = blockIdx.x;
  const int H = dmc_unpadded(Hp), W = dmc_unpadded(Wp);
  const float inv = 1.f / (float)(H * W);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int h = 1; h <= H; ++h)
      for (int w = 1; w <= W; ++w) {
        const long off = (((long)f * Hp + h) * Wp + w) * C + c;
        s += join_bf16(hi[off], lo[off]);
      }
    pooled[(long)f * C + c] = s * inv;
  }
}

__global__ void avgpool_bwd_kernel(const float* __restrict__ dpooled, int frames, int Hp, int Wp,
                                   int C, float* __restrict__ dX) {
  const long n = (long)frames * Hp * Wp * C;
  const float inv = 1.f / (float)(dmc_unpadded(Hp) * dmc_unpadded(Wp));
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long q = i / C;
    const int f = (int)(q / ((long)Hp * Wp));
    dX[i] = interior(q, Hp, Wp) ? dpooled[(long)f * C + c] * inv : 0.f;
  }
}

// fp32 padded pixel-major tensor -> bf16 hi/lo planes (ring forced to zero).
__global__ void __launch_bounds__(256)
split_planes_kernel(const float* __restrict__ X, long P, int C, int Hp, int Wp,
                    bf16* __restrict__ hi, bf16* __restrict__ lo) {
  const FastDiv fd((unsigned)(C / 4));
  const unsigned n = (unsigned)(P * (C / 4));
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned q = fd.div(i);
    const long off = (long)q * C + fd.mod(i) * 4;
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    if (Hp == 0 || interior(q, Hp, Wp)) {
      const float4 v = *reinterpret_cast<const float4*>(X + off);
      o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    }
    store_split4(hi, lo, off, o);
  }
}

static int reduce_block(int C) {
  // threads per block: a multiple of C/4 close to 256
  const int c4n = C / 4;
  int rows = 256 / c4n;
  if (rows < 1) rows = 1;
  return rows * c4n;
}

}  // namespace dmc

using namespace dmc;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int dmc_memset_zero(void* p, long bytes, void* stream) {
  if (bytes <= 0) return DMC_OK;
  if (cudaMemsetAsync(p, 0, (size_t)bytes, ST(stream)) != cudaSuccess)
    return dmc_check_launch("memset");
  return DMC_OK;
}

extern "C" int dmc_weight_prep(const float* w_oihw, int Cout, int Cin, int taps, void* W_hi,
                               void* W_lo, void* Wt_hi, void* Wt_lo, void* stream) {
  DMC_REQUIRE(Cout > 0 && Cin > 0 && taps > 0, "weight_prep: bad shape");
  const long n = (long)Cout * Cin * taps;
  weight_prep_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(
      w_oihw, Cout, Cin, taps, (bf16*)W_hi, (bf16*)W_lo, (bf16*)Wt_hi, (bf16*)Wt_lo);
  return dmc_check_launch("weight_prep_kernel");
}

// chunks: device array of PrepChunk {int64 w_off, Wh, Wl, Th, Tl; int32 Cout, Cin, taps, tile}, one per
// 32 x 32 (co, ci) tile: tile = (co / 32) * (Cin / 32) + ci / 32; Cout and Cin multiples of 32, taps <= 9
extern "C" int dmc_weight_prep_multi(const float* params, const void* chunks, int nchunks,
                                     void* stream) {
  if (nchunks <= 0) return DMC_OK;
  weight_prep_multi_kernel<<<nchunks, 256, 0, ST(stream)>>>(params,
                                                            reinterpret_cast<const PrepChunk*>(chunks));
  return dmc_check_launch("weight_prep_multi_kernel");
}

extern "C" int dmc_wgrad_unpack(const float* dWs, float* grad_oihw, int Cout, int Cin, int taps,
                                void* stream) {
  const long n = (long)Cout * Cin * taps;
  wgrad_unpack_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(dWs, grad_oihw, Cout, Cin, taps);
  return dmc_check_launch("wgrad_unpack_kernel");
}

extern "C" int dmc_bn_stats(const float* Y, long P, int C, double* sums, void* stream) {
  DMC_REQUIRE(C % 4 == 0 && C >= 4 && C <= 1024, "bn_stats: C=%d", C);
  if (cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, ST(stream)) != cudaSuccess)
    return dmc_check_launch("bn_stats memset");
  const int threads = reduce_block(C);
  const int rows = threads / (C / 4);
  int blocks = (int)cdiv(P, (long)rows * 16);
  if (blocks > 148 * 4) blocks = 148 * 4;      // 4 resident blocks per SM: the kernel is latency-bound
  if (blocks < 1) blocks = 1;
  bn_stats_kernel<<<blocks, threads, sizeof(double) * 8 * threads, ST(stream)>>>(Y, P, C, sums);
  return dmc_check_launch("bn_stats_kernel");
}

extern "C" int dmc_bn_finalize(const double* sums, double count, const float* gamma,
                               const float* beta, float* running_mean, float* running_var,
                               long long* num_batches_tracked, float momentum, float eps, int C,
                               float* scale, float* shift, float* mean, float* invstd,
                               void* stream) {
  bn_finalize_kernel<<<(int)cdiv(C, 128), 128, 0, ST(stream)>>>(
      sums, count, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps, C,
      scale, shift, mean, invstd);
  return dmc_check_launch("bn_finalize_kernel");
}

extern "C" int dmc_bn_eval_coeffs(const float* gamma, const float* beta, const float* running_mean,
                                  const float* running_var, float eps, int C, float* scale,
                                  float* shift, void* stream) {
  bn_eval_coeffs_kernel<<<(int)cdiv(C, 128), 128, 0, ST(stream)>>>(gamma, beta, running_mean,
                                                                   running_var, eps, C, scale, shift);
  return dmc_check_launch("bn_eval_coeffs_kernel");
}

extern "C" int dmc_bn_apply(const float* Y, const float* scale, const float* shift, long P, int C,
                            int Hp, int Wp, int relu, const void* res_hi, const void* res_lo,
                            const float* resY, const float* res_scale, const float* res_shift,
                            void* out_hi, void* out_lo, void* stream) {
  DMC_REQUIRE(C % 4 == 0, "bn_apply: C=%d", C);
  bn_apply_kernel<<<grid_for(P * (C / 4), 256), 256, 0, ST(stream)>>>(
      Y, scale, shift, P, C, Hp, Wp, relu, (const bf16*)res_hi, (const bf16*)res_lo, resY, res_scale,
      res_shift, (bf16*)out_hi, (bf16*)out_lo, 0.f, 1);
  return dmc_check_launch("bn_apply_kernel");
}

// out = LeakyReLU_slope(Y*scale + shift) as bf16 hi/lo planes on a layout with a zero ring of `ring`
// rows / columns (ContextNetwork blocks: dilated conv -> BatchNorm -> LeakyReLU(0.1),
// code/dmcnet/model.py:31-42).  slope = 1 applies no activation.
extern "C" int dmc_bn_apply_lrelu(const float* Y, const float* scale, const float* shift, long P, int C,
                                  int Hp, int Wp, int ring, float slope, void* out_hi, void* out_lo,
                                  void* stream) {
  DMC_REQUIRE(C % 4 == 0 && ring >= 1 && ring < Hp && ring < Wp, "bn_apply_lrelu: C=%d ring=%d", C, ring);
  bn_apply_kernel<<<grid_for(P * (C / 4), 256), 256, 0, ST(stream)>>>(
      Y, scale, shift, P, C, Hp, Wp, 1, nullptr, nullptr, nullptr, nullptr, nullptr, (bf16*)out_hi,
      (bf16*)out_lo, slope, ring);
  return dmc_check_launch("bn_apply_kernel");
}

extern "C" int dmc_bn_bwd_reduce(const float* g_a, const float* g_b, const void* act_hi,
                                 const float* Y, const float* mean, const float* invstd, long P,
                                 int C, int Hp, int Wp, double* sums2, void* stream) {
  DMC_REQUIRE(C % 4 == 0 && C >= 4 && C <= 1024, "bn_bwd_reduce: C=%d", C);
  if (cudaMemsetAsync(sums2, 0, sizeof(double) * 2 * C, ST(stream)) != cudaSuccess)
    return dmc_check_launch("bn_bwd_reduce memset");
  const int threads = reduce_block(C);
  const int rows = threads / (C / 4);
  int blocks = (int)cdiv(P, (long)rows * 16);
  if (blocks > 148 * 4) blocks = 148 * 4;      // 4 resident blocks per SM: the kernel is latency-bound
  if (blocks < 1) blocks = 1;
  bn_bwd_reduce_kernel<<<blocks, threads, sizeof(double) * 8 * threads, ST(stream)>>>(
      g_a, g_b, (const bf16*)act_hi, Y, mean, invstd, P, C, Hp, Wp, sums2);
  return dmc_check_launch("bn_bwd_reduce_kernel");
}

extern "C" int dmc_bn_bwd_apply(const float* g_a, const float* g_b, const void* act_hi,
                                const float* Y, const float* mean, const float* invstd,
                                const float* gamma, const double* sums2, double count, long P, int C,
                                int Hp, int Wp, void* G_hi, void* G_lo, float* dz_out, float* dgamma,
                                float* dbeta, void* stream) {
  DMC_REQUIRE(C % 4 == 0, "bn_bwd_apply: C=%d", C);
  bn_bwd_apply_kernel<<<grid_for(P * (C / 4), 256), 256, 3 * C * sizeof(float), ST(stream)>>>(
      g_a, g_b, (const bf16*)act_hi, Y, mean, invstd, gamma, sums2, count, P, C, Hp, Wp, (bf16*)G_hi,
      (bf16*)G_lo, dz_out, dgamma, dbeta);
  return dmc_check_launch("bn_bwd_apply_kernel");
}

extern "C" int dmc_phase_split(const void* in_hi, const void* in_lo, int frames, int H, int W, int C,
                               void* out_hi, void* out_lo, void* stream) {
  DMC_REQUIRE(H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "phase_split: H=%d W=%d C=%d", H, W, C);
  const long n = 4L * frames * dmc_padded(H / 2) * dmc_padded(W / 2) * (C / 8);
  phase_split_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(
      (const bf16*)in_hi, (const bf16*)in_lo, frames, H, W, C, (bf16*)out_hi, (bf16*)out_lo);
  return dmc_check_launch("phase_split_kernel");
}

extern "C" int dmc_phase_unsplit(const float* in, int frames, int H, int W, int C, float* out,
                                 void* stream) {
  DMC_REQUIRE(H % 2 == 0 && W % 2 == 0 && C % 4 == 0, "phase_unsplit: H=%d W=%d C=%d", H, W, C);
  const long n = (long)frames * dmc_padded(H) * dmc_padded(W) * (C / 4);
  phase_unsplit_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(in, frames, H, W, C, out);
  return dmc_check_launch("phase_unsplit_kernel");
}

// dmc_phase_unsplit fused with the ReLU mask and the two BatchNorm-backward reductions of the unit the
// gradient belongs to: out [frames][H+1][W+1][C] = dz = unsplit(in) * [act_hi > 0] (ring zero),
// sums2 (double [2][C], caller-zeroed) += (sum dz, sum dz * (Y - mean) * invstd).
extern "C" int dmc_phase_unsplit_reduce(const float* in, int frames, int H, int W, int C, const void* act_hi,
                                        const float* Y, const float* mean, const float* invstd, float* out,
                                        double* sums2, void* stream) {
  DMC_REQUIRE(H % 2 == 0 && W % 2 == 0 && C % 4 == 0, "phase_unsplit_reduce: H=%d W=%d C=%d", H, W, C);
  DMC_REQUIRE(act_hi && Y && mean && invstd && sums2, "phase_unsplit_reduce: null argument");
  const long P = (long)frames * dmc_padded(H) * dmc_padded(W);
  DMC_REQUIRE(P < (1L << 31), "phase_unsplit_reduce: too many rows");
  const int threads = reduce_block(C);
  const int rows_per_iter = threads / (C / 4);
  long blocks = cdiv(P, (long)rows_per_iter * 8);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  phase_unsplit_reduce_kernel<<<(unsigned)blocks, threads, 2 * threads * 4 * sizeof(double), ST(stream)>>>(
      in, frames, H, W, C, (const bf16*)act_hi, Y, mean, invstd, out, sums2);
  return dmc_check_launch("phase_unsplit_reduce_kernel");
}

extern "C" int dmc_avgpool(const void* hi, const void* lo, int frames, int Hp, int Wp, int C,
                           float* pooled, void* stream) {
  avgpool_kernel<<<frames, 256, 0, ST(stream)>>>((const bf16*)hi, (const bf16*)lo, frames, Hp, Wp, C,
                                                 pooled);
  return dmc_check_launch("avgpool_kernel");
}

extern "C" int dmc_avgpool_bwd(const float* dpooled, int frames, int Hp, int Wp, int C, float* dX,
                               void* stream) {
  const long n = (long)frames * Hp * Wp * C;
  avgpool_bwd_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(dpooled, frames, Hp, Wp, C, dX);
  return dmc_check_launch("avgpool_bwd_kernel");
}

extern "C" int dmc_split_planes(const float* X, long P, int C, int Hp, int Wp, void* hi, void* lo,
                                void* stream) {
  DMC_REQUIRE(C % 4 == 0, "split_planes: C=%d", C);
  split_planes_kernel<<<grid_for(P * (C / 4), 256), 256, 0, ST(stream)>>>(X, P, C, Hp, Wp, (bf16*)hi,
                                                                          (bf16*)lo);
  return dmc_check_launch("split_planes_kernel");
}
