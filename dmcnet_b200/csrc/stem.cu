// Classifier stem glue: BatchNorm + ReLU + MaxPool(3x3, stride 2, pad 1) fused,
// converting the planar NCHW stem-conv output into the padded pixel-major bf16
// hi/lo layout the tensor-core layers consume (torchvision resnet18 stem:
// conv1 -> bn1 -> relu -> maxpool, used by code/dmcnet/model.py:305).
// The pooling argmax (first maximum in row-major window order, as ATen) is
// kept as one byte per output so the backward pass is a pure gather.
#include "common.cuh"

namespace dmc {

// grid (Hq, N); Y [N][C][H][W] -> hi/lo [N][Hq+2][Wq+2][C] interior, idx [N][Hq][Wq][C]
__global__ void __launch_bounds__(256)
stem_pool_fwd_kernel(const float* __restrict__ Y, const float* __restrict__ scale,
                     const float* __restrict__ shift, int C, int H, int W, bf16* __restrict__ out_hi,
                     bf16* __restrict__ out_lo, unsigned char* __restrict__ idx) {
  constexpr int CG = 32;                       // channels per pass
  extern __shared__ float rows[];              // [CG][3][W+1]
  const int Hq = H / 2, Wq = W / 2;
  const int ph = blockIdx.x, n = blockIdx.y;
  const int WP = W + 1;
  for (int cg = 0; cg < C; cg += CG) {
    for (int i = threadIdx.x; i < CG * 3 * W; i += blockDim.x) {
      const int w = i % W, r = (i / W) % 3, c = i / (3 * W);
      const int h = 2 * ph - 1 + r;
      float v = -1.f;                          // marks "outside the image"
      if (h >= 0 && h < H) {
        const float y = Y[(((long)n * C + cg + c) * H + h) * W + w];
        v = fmaxf(fmaf(y, scale[cg + c], shift[cg + c]), 0.f);
      }
      rows[(c * 3 + r) * WP + w] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Wq * CG; i += blockDim.x) {
      const int c = i % CG, pw = i / CG;
      float best = -1.f;
      int bi = 0;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int w = 2 * pw - 1 + s;
          if (w < 0 || w >= W) continue;
          const float v = rows[(c * 3 + r) * WP + w];
          if (v > best) { best = v; bi = r * 3 + s; }
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

// grid (H, N); dZ[n][c][h][w] = relu'(bn(Y)) * sum_{windows whose argmax is (h,w)} dA[...]
__global__ void __launch_bounds__(256)
stem_pool_bwd_kernel(const float* __restrict__ g_a, const float* __restrict__ g_b,
                     const unsigned char* __restrict__ idx, const float* __restrict__ Y,
                     const float* __restrict__ scale, const float* __restrict__ shift, int C, int H,
                     int W, float* __restrict__ dZ) {
  extern __shared__ float sm[];                // g [2][C][Wq+1] floats, then idx bytes [2][C][Wq+1]
  const int Hq = H / 2, Wq = W / 2, WQP = Wq + 1;
  const int h = blockIdx.x, n = blockIdx.y;
  float* g_s = sm;
  unsigned char* i_s = reinterpret_cast<unsigned char*>(sm + 2 * C * WQP);
  // candidate pooled rows: 2*ph-1 <= h <= 2*ph+1
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
  for (int i = threadIdx.x; i < C * W; i += blockDim.x) {
    const int w = i % W, c = i / W;
    const long o = (((long)n * C + c) * H + h) * W + w;
    const float a = fmaf(Y[o], scale[c], shift[c]);
    float acc = 0.f;
    if (a > 0.f) {
      const int pw_lo = w / 2, npw = (w & 1) ? 2 : 1;
      for (int k = 0; k < nph; ++k) {
        const int ph = ph_lo + k;
        if (ph >= Hq) continue;
        const int r = h - (2 * ph - 1);
        for (int j = 0; j < npw; ++j) {
          const int pw = pw_lo + j;
          if (pw >= Wq) continue;
          const int s = w - (2 * pw - 1);
          if (i_s[(k * C + c) * WQP + pw] == r * 3 + s) acc += g_s[(k * C + c) * WQP + pw];
        }
      }
    }
    dZ[o] = acc;
  }
}

}  // namespace dmc

using namespace dmc;

extern "C" int dmc_stem_pool_fwd(const float* Y, const float* scale, const float* shift, int N, int C,
                                 int H, int W, void* out_hi, void* out_lo, unsigned char* idx,
                                 void* stream) {
  DMC_REQUIRE(C % 32 == 0 && H % 2 == 0 && W % 2 == 0, "stem_pool_fwd: C=%d H=%d W=%d", C, H, W);
  const int smem = 32 * 3 * (W + 1) * (int)sizeof(float);
  DMC_REQUIRE(smem <= 48 * 1024, "stem_pool_fwd: W=%d too wide", W);
  stem_pool_fwd_kernel<<<dim3(H / 2, N), 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      Y, scale, shift, C, H, W, (bf16*)out_hi, (bf16*)out_lo, idx);
  return dmc_check_launch("stem_pool_fwd_kernel");
}

extern "C" int dmc_stem_pool_bwd(const float* g_a, const float* g_b, const unsigned char* idx,
                                 const float* Y, const float* scale, const float* shift, int N, int C,
                                 int H, int W, float* dZ, void* stream) {
  const int WQP = W / 2 + 1;
  const int smem = 2 * C * WQP * (int)sizeof(float) + 2 * C * WQP;
  DMC_REQUIRE(smem <= 48 * 1024, "stem_pool_bwd: smem %d", smem);
  stem_pool_bwd_kernel<<<dim3(H, N), 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      g_a, g_b, idx, Y, scale, shift, C, H, W, dZ);
  return dmc_check_launch("stem_pool_bwd_kernel");
}
