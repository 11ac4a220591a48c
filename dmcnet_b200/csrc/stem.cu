// Classifier stem glue: BatchNorm + ReLU + MaxPool(3x3, stride 2, pad 1) fused,
// converting the planar NCHW stem-conv output into the padded pixel-major bf16
// hi/lo layout the tensor-core layers consume (torchvision resnet18 stem:
// conv1 -> bn1 -> relu -> maxpool, used by code/dmcnet/model.py:305).
// The pooling argmax (first maximum in row-major window order, as ATen) is
// kept as one byte per output so the backward pass is a pure gather.
#include "common.cuh"

namespace dmc {

// grid (Hq, N); Y [N][C][H][W] -> hi/lo [N][Hq+2][Wq+2][C] interior, idx [N][Hq][Wq][C].
// One warp streams one (channel, input row) at a time with 128-bit loads (no per-element
// index arithmetic); the pooled row is then produced channel-fastest so the pixel-major
// stores coalesce.
__global__ void __launch_bounds__(256)
stem_pool_fwd_kernel(const float* __restrict__ Y, const float* __restrict__ scale,
                     const float* __restrict__ shift, int C, int H, int W, bf16* __restrict__ out_hi,
                     bf16* __restrict__ out_lo, unsigned char* __restrict__ idx) {
  constexpr int CG = 32;                       // channels per pass
  extern __shared__ float rows[];              // [CG][3][W+1]
  const int Hq = H / 2, Wq = W / 2;
  const int ph = blockIdx.x, n = blockIdx.y;
  const int WP = W + 1, W4 = W / 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int cg = 0; cg < C; cg += CG) {
    for (int cr = warp; cr < CG * 3; cr += nwarps) {
      const int c = cr / 3, r = cr - 3 * c;
      const int h = 2 * ph - 1 + r;
      float* dst = rows + cr * WP;
      if (h >= 0 && h < H) {
        const float sc = scale[cg + c], sh = shift[cg + c];
        const float4* src = reinterpret_cast<const float4*>(Y + (((long)n * C + cg + c) * H + h) * W);
        for (int j = lane; j < W4; j += 32) {
          const float4 v = src[j];
          dst[4 * j + 0] = fmaxf(fmaf(v.x, sc, sh), 0.f);
          dst[4 * j + 1] = fmaxf(fmaf(v.y, sc, sh), 0.f);
          dst[4 * j + 2] = fmaxf(fmaf(v.z, sc, sh), 0.f);
          dst[4 * j + 3] = fmaxf(fmaf(v.w, sc, sh), 0.f);
        }
      } else {
        for (int j = lane; j < W; j += 32) dst[j] = -1.f;     // marks "outside the image"
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Wq * CG; i += blockDim.x) {
      const int c = i % CG, pw = i / CG;
      float best = -1.f;
      int bi = 0;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s3 = 0; s3 < 3; ++s3) {
          const int w = 2 * pw - 1 + s3;
          if (w < 0 || w >= W) continue;
          const float v = rows[(c * 3 + r) * WP + w];
          if (v > best) { best = v; bi = r * 3 + s3; }
        }
      const long o = (((long)n * (Hq + 2) + ph + 1) * (Wq + 2) + pw + 1) * C + cg + c;
      bf16 h, l;
      split_bf16(best, h, l);
      out_hi[o] = h;
      out_lo[o] = l;
      idx[(((long)n * Hq + ph) * Wq + pw) * C + cg + c] = (unsigned char)bi;
    }
    __syncthreads();
  }
}

// grid (H, N); dZ[n][c][h][w] = relu'(bn(Y)) * sum_{windows whose argmax is (h,w)} dA[...].
// The pooled gradient rows are staged channel-transposed once; every warp then streams one
// channel row of Y / dZ with 128-bit accesses.
__global__ void __launch_bounds__(256)
stem_pool_bwd_kernel(const float* __restrict__ g_a, const float* __restrict__ g_b,
                     const unsigned char* __restrict__ idx, const float* __restrict__ Y,
                     const float* __restrict__ scale, const float* __restrict__ shift, int C, int H,
                     int W, float* __restrict__ dZ) {
  extern __shared__ float sm[];                // g [2][C][Wq+1] floats, then idx bytes [2][C][Wq+1]
  const int Hq = H / 2, Wq = W / 2, WQP = Wq + 1, W4 = W / 4;
  const int h = blockIdx.x, n = blockIdx.y;
  float* g_s = sm;
  unsigned char* i_s = reinterpret_cast<unsigned char*>(sm + 2 * C * WQP);
  const int ph_lo = h / 2;                      // h even -> {h/2}; h odd -> {(h-1)/2, (h+1)/2}
  const int nph = (h & 1) ? 2 : 1;
  for (int i = threadIdx.x; i < nph * Wq * C; i += blockDim.x) {
    const int c = i % C, pw = (i / C) % Wq, k = i / (C * Wq);
    const int ph = ph_lo + k;
    float g = 0.f;
    unsigned char id = 255;
    if (ph < Hq) {
      const long o = (((long)n * (Hq + 2) + ph + 1) * (Wq + 2) + pw + 1) * C + c;
      g = g_a[o];
      if (g_b) g += g_b[o];
      id = idx[(((long)n * Hq + ph) * Wq + pw) * C + c];
    }
    g_s[(k * C + c) * WQP + pw] = g;
    i_s[(k * C + c) * WQP + pw] = id;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int c = warp; c < C; c += nwarps) {
    const float sc = scale[c], sh = shift[c];
    const long base = (((long)n * C + c) * H + h) * W;
    const float4* ysrc = reinterpret_cast<const float4*>(Y + base);
    float4* zdst = reinterpret_cast<float4*>(dZ + base);
    for (int j = lane; j < W4; j += 32) {
      const float4 y = ysrc[j];
      const float yv[4] = {y.x, y.y, y.z, y.w};
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int w = 4 * j + e;
        float acc = 0.f;
        if (fmaf(yv[e], sc, sh) > 0.f) {
          const int pw_lo = w / 2, npw = (w & 1) ? 2 : 1;
          for (int k = 0; k < nph; ++k) {
            const int ph = ph_lo + k;
            if (ph >= Hq) continue;
            const int r = h - (2 * ph - 1);
            for (int q = 0; q < npw; ++q) {
              const int pw = pw_lo + q;
              if (pw >= Wq) continue;
              const int s3 = w - (2 * pw - 1);
              if (i_s[(k * C + c) * WQP + pw] == r * 3 + s3) acc += g_s[(k * C + c) * WQP + pw];
            }
          }
        }
        o[e] = acc;
      }
      zdst[j] = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}

}  // namespace dmc

using namespace dmc;

extern "C" int dmc_stem_pool_fwd(const float* Y, const float* scale, const float* shift, int N, int C,
                                 int H, int W, void* out_hi, void* out_lo, unsigned char* idx,
                                 void* stream) {
  DMC_REQUIRE(C % 32 == 0 && H % 2 == 0 && W % 4 == 0, "stem_pool_fwd: C=%d H=%d W=%d", C, H, W);
  const int smem = 32 * 3 * (W + 1) * (int)sizeof(float);
  DMC_REQUIRE(smem <= 48 * 1024, "stem_pool_fwd: W=%d too wide", W);
  stem_pool_fwd_kernel<<<dim3(H / 2, N), 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      Y, scale, shift, C, H, W, (bf16*)out_hi, (bf16*)out_lo, idx);
  return dmc_check_launch("stem_pool_fwd_kernel");
}

extern "C" int dmc_stem_pool_bwd(const float* g_a, const float* g_b, const unsigned char* idx,
                                 const float* Y, const float* scale, const float* shift, int N, int C,
                                 int H, int W, float* dZ, void* stream) {
  DMC_REQUIRE(W % 4 == 0, "stem_pool_bwd: W=%d must be a multiple of 4", W);
  const int WQP = W / 2 + 1;
  const int smem = 2 * C * WQP * (int)sizeof(float) + 2 * C * WQP;
  DMC_REQUIRE(smem <= 48 * 1024, "stem_pool_bwd: smem %d", smem);
  stem_pool_bwd_kernel<<<dim3(H, N), 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      g_a, g_b, idx, Y, scale, shift, C, H, W, dZ);
  return dmc_check_launch("stem_pool_bwd_kernel");
}
