// Classifier stem glue: BatchNorm + ReLU + MaxPool(3x3, stride 2, pad 1) fused,
// converting the planar NCHW stem-conv output into the padded pixel-major bf16
// hi/lo layout the tensor-core layers consume (torchvision resnet18 stem:
// conv1 -> bn1 -> relu -> maxpool, used by code/dmcnet/model.py:305).
// The pooling argmax (first maximum in row-major window order, as ATen) is
// kept as one byte per output so the backward pass is a pure gather.
#include "common.cuh"

namespace dmc {

// grid (Hq, N); Y [N][C][H][W] -> hi/lo padded pixel-major [N][Hq+1][Wq+1][C] interior (common.cuh),
// idx [N][Hq][Wq][C].
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
    // four (channel, row) strips per warp in flight: the loads are issued before the first use
    for (int cr0 = warp; cr0 < CG * 3; cr0 += 4 * nwarps) {
      float4 v[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int cr = cr0 + u * nwarps;
        const int c = cr / 3, r = cr - 3 * c;
        const int h = 2 * ph - 1 + r;
        ok[u] = cr < CG * 3 && h >= 0 && h < H && lane < W4;
        if (ok[u])
          v[u] = reinterpret_cast<const float4*>(Y + (((long)n * C + cg + c) * H + h) * W)[lane];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int cr = cr0 + u * nwarps;
        if (cr >= CG * 3) break;
        const int c = cr / 3, r = cr - 3 * c;
        const int h = 2 * ph - 1 + r;
        float* dst = rows + cr * WP;
        if (h >= 0 && h < H) {
          if (ok[u]) {
            const float sc = scale[cg + c], sh = shift[cg + c];
            dst[4 * lane + 0] = fmaxf(fmaf(v[u].x, sc, sh), 0.f);
            dst[4 * lane + 1] = fmaxf(fmaf(v[u].y, sc, sh), 0.f);
            dst[4 * lane + 2] = fmaxf(fmaf(v[u].z, sc, sh), 0.f);
            dst[4 * lane + 3] = fmaxf(fmaf(v[u].w, sc, sh), 0.f);
          }
          for (int j = lane + 32; j < W4; j += 32) {      // rows wider than 128 pixels
            const float sc = scale[cg + c], sh = shift[cg + c];
            const float4 t = reinterpret_cast<const float4*>(Y + (((long)n * C + cg + c) * H + h) * W)[j];
            dst[4 * j + 0] = fmaxf(fmaf(t.x, sc, sh), 0.f);
            dst[4 * j + 1] = fmaxf(fmaf(t.y, sc, sh), 0.f);
            dst[4 * j + 2] = fmaxf(fmaf(t.z, sc, sh), 0.f);
            dst[4 * j + 3] = fmaxf(fmaf(t.w, sc, sh), 0.f);
          }
        } else {
          for (int j = lane; j < W; j += 32) dst[j] = -1.f;     // marks "outside the image"
        }
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
      const long o = (((long)n * dmc_padded(Hq) + ph + 1) * dmc_padded(Wq) + pw + 1) * C + cg + c;
      bf16 h, l;
      split_bf16(best, h, l);
      out_hi[o] = h;
      out_lo[o] = l;
      idx[(((long)n * Hq + ph) * Wq + pw) * C + cg + c] = (unsigned char)bi;
    }
    __syncthreads();
  }
}

// dZ[n][c][h][w] = relu'(bn(Y)) * sum_{windows whose argmax is (h,w)} dA[...].
// Scatter form: a CTA owns 2*PB input rows of one frame; for CG channels at a time it zeroes a
// shared-memory tile, adds every pooled gradient of rows ph0 .. ph0+PB into the tile cell its
// argmax byte names (cells outside the owned rows belong to the neighbour band), then
// streams the tile out through the ReLU mask with 128-bit accesses.
// ~1/4 of the instructions of the gather form, and every global access is a full sector.
constexpr int SPB_PB = 2;      // pooled rows per band
constexpr int SPB_CG = 32;     // channels per pass
__global__ void __launch_bounds__(256)
stem_pool_bwd_kernel(const float* __restrict__ g_a, const float* __restrict__ g_b,
                     const unsigned char* __restrict__ idx, const float* __restrict__ Y,
                     const float* __restrict__ scale, const float* __restrict__ shift, int C, int H,
                     int W, float* __restrict__ dZ) {
  constexpr int PB = SPB_PB, CG = SPB_CG, RO = 2 * PB;
  extern __shared__ float tile[];                // [CG][RO][WP] (+4 floats between planes)
  const int Hq = H / 2, Wq = W / 2;
  const int WP = W + 4, PLANE = RO * WP + 4;
  const int ph0 = blockIdx.x * PB, n = blockIdx.y;
  const int h0 = 2 * ph0;
  const int nph = (ph0 + PB < Hq) ? PB + 1 : Hq - ph0;        // pooled rows that reach owned rows
  const int W4 = W / 4;
  for (int cg = 0; cg < C; cg += CG) {
    for (int i = threadIdx.x; i < CG * PLANE; i += blockDim.x) tile[i] = 0.f;
    __syncthreads();
    // windows of equal (row, column) parity are disjoint (3x3, stride 2): four race-free passes
    // of plain read-modify-write in a fixed order -> deterministic, no atomics
    for (int pass = 0; pass < 4; ++pass) {
      const int kp = pass >> 1, wp = pass & 1;
      const int nk = (nph - kp + 1) / 2, nw = (Wq - wp + 1) / 2;
      const int cnt = nk * nw * CG;
      for (int i0 = threadIdx.x; i0 < cnt; i0 += 4 * blockDim.x) {
        float g[4];
        int id[4], kk[4], pww[4], cc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * blockDim.x;
          const int ii = i < cnt ? i : i0;
          cc[u] = ii % CG;
          pww[u] = 2 * ((ii / CG) % nw) + wp;
          kk[u] = i < cnt ? 2 * (ii / (CG * nw)) + kp : -1;
          const int ph = ph0 + (kk[u] < 0 ? kp : kk[u]);
          const long o = (((long)n * dmc_padded(Hq) + ph + 1) * dmc_padded(Wq) + pww[u] + 1) * C + cg + cc[u];
          g[u] = g_a[o];
          if (g_b) g[u] += g_b[o];
          id[u] = idx[(((long)n * Hq + ph) * Wq + pww[u]) * C + cg + cc[u]];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (kk[u] < 0) continue;
          const int r = (id[u] * 11) >> 5, s3 = id[u] - 3 * r;     // id / 3, id % 3 for id in 0..8
          const int hl = 2 * kk[u] - 1 + r, w = 2 * pww[u] - 1 + s3;   // row relative to h0
          if (hl >= 0 && hl < RO) tile[cc[u] * PLANE + hl * WP + w] += g[u];
        }
      }
      __syncthreads();
    }
    // stream the tile out; four independent 128-bit loads in flight per thread
    const int items = CG * RO * W4;
    for (int i0 = threadIdx.x; i0 < items; i0 += 4 * blockDim.x) {
      float4 y[4];
      long base[4];
      int toff[4];
      float sc[4], sh[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * blockDim.x;
        const int ii = i < items ? i : i0;
        const int j = ii % W4, hl = (ii / W4) % RO, c = ii / (W4 * RO);
        const int h = h0 + hl;
        base[u] = (i < items && h < H) ? (((long)n * C + cg + c) * H + h) * W + 4 * j : -1;
        toff[u] = c * PLANE + hl * WP + 4 * j;
        sc[u] = scale[cg + c];
        sh[u] = shift[cg + c];
        if (base[u] >= 0) y[u] = *reinterpret_cast<const float4*>(Y + base[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (base[u] < 0) continue;
        const float4 t = *reinterpret_cast<const float4*>(&tile[toff[u]]);
        float4 o;
        o.x = fmaf(y[u].x, sc[u], sh[u]) > 0.f ? t.x : 0.f;
        o.y = fmaf(y[u].y, sc[u], sh[u]) > 0.f ? t.y : 0.f;
        o.z = fmaf(y[u].z, sc[u], sh[u]) > 0.f ? t.z : 0.f;
        o.w = fmaf(y[u].w, sc[u], sh[u]) > 0.f ? t.w : 0.f;
        *reinterpret_cast<float4*>(dZ + base[u]) = o;
      }
    }
    __syncthreads();
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
  DMC_REQUIRE(W % 4 == 0 && H % 2 == 0 && C % SPB_CG == 0, "stem_pool_bwd: C=%d H=%d W=%d", C, H, W);
  const int smem = SPB_CG * (2 * SPB_PB * (W + 4) + 4) * (int)sizeof(float);
  static int attr_bytes = 0;
  if (smem > attr_bytes) {
    if (cudaFuncSetAttribute(stem_pool_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
        cudaSuccess)
      return dmc_check_launch("stem_pool_bwd smem attribute");
    attr_bytes = smem;
  }
  const int bands = (H / 2 + SPB_PB - 1) / SPB_PB;
  stem_pool_bwd_kernel<<<dim3(bands, N), 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      g_a, g_b, idx, Y, scale, shift, C, H, W, dZ);
  return dmc_check_launch("stem_pool_bwd_kernel");
}
