// Micro-benchmark: FP32 FMA issue rates on sm_100a (scalar FFMA vs packed FFMA2, register vs
// shared-memory operands).  Prints FMA/clk/SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 fma.cu -o fma
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITER = 4096;
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, const float* in, long long* clk) {
  __shared__ __align__(16) float ws[64];
  if (threadIdx.x < 64) ws[threadIdx.x] = in[threadIdx.x];
  __syncthreads();
  float2 a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = make_float2(in[i] + threadIdx.x, in[i + 8]);
  float2 b = make_float2(in[16], in[17]), c = make_float2(in[18], in[19]);
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
    if (MODE == 0) {          // 16 scalar FFMA, 3 register operands
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }
    } else if (MODE == 1) {   // 8 FFMA2 (16 FMA)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = __ffma2_rn(a[i], b, c);
    } else if (MODE == 2) {   // conv-like: acc += v * w, w from smem (LDS.64 broadcast), 8 FFMA2 per 8 LDS.64
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 w = *reinterpret_cast<const float2*>(&ws[(2 * i + it) & 62]);
        a[i] = __ffma2_rn(b, w, a[i]);
      }
    } else if (MODE == 3) {   // conv-like: 4 pixels x 4 channel pairs, 2 LDS.128 weights -> 16 FFMA2
      const float4 w0 = *reinterpret_cast<const float4*>(&ws[(it * 8) & 56]);
      const float4 w1 = *reinterpret_cast<const float4*>(&ws[(it * 8 + 4) & 60]);
      const float2 wa = make_float2(w0.x, w0.y), wb = make_float2(w0.z, w0.w);
      const float2 wc = make_float2(w1.x, w1.y), wd = make_float2(w1.z, w1.w);
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const float2 v = p ? c : b;
        a[p * 4 + 0] = __ffma2_rn(v, wa, a[p * 4 + 0]);
        a[p * 4 + 1] = __ffma2_rn(v, wb, a[p * 4 + 1]);
        a[p * 4 + 2] = __ffma2_rn(v, wc, a[p * 4 + 2]);
        a[p * 4 + 3] = __ffma2_rn(v, wd, a[p * 4 + 3]);
      }
    } else if (MODE == 4) {   // scalar: acc += v * w with w from smem, 16 FFMA per 4 LDS.128
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 w = *reinterpret_cast<const float4*>(&ws[(it * 4 + i * 4) & 60]);
        a[2 * i].x = fmaf(b.x, w.x, a[2 * i].x);
        a[2 * i].y = fmaf(b.x, w.y, a[2 * i].y);
        a[2 * i + 1].x = fmaf(b.x, w.z, a[2 * i + 1].x);
        a[2 * i + 1].y = fmaf(b.x, w.w, a[2 * i + 1].y);
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, int fma_per_iter, int ctas_per_sm) {
  float *out, *in; long long* clk;
  int nb = 148 * ctas_per_sm;
  cudaMalloc(&out, nb * 256 * 4); cudaMalloc(&in, 256); cudaMalloc(&clk, nb * 8);
  cudaMemset(in, 0, 256);
  k<MODE><<<nb, 256>>>(out, in, clk);
  k<MODE><<<nb, 256>>>(out, in, clk);
  cudaDeviceSynchronize();
  long long h[148 * 8];
  cudaMemcpy(h, clk, nb * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < nb; ++i) avg += h[i]; avg /= nb;
  double fma = (double)ITER * fma_per_iter * 256 * ctas_per_sm;
  printf("%-44s ctas/SM=%d  %.1f FMA/clk/SM\n", name, ctas_per_sm, fma / avg);
  cudaFree(out); cudaFree(in); cudaFree(clk);
}
int main() {
  for (int c : {2, 4}) {
    run<0>("FFMA reg,reg,reg (16 chains)", 16, c);
    run<1>("FFMA2 reg (8 chains)", 16, c);
    run<2>("FFMA2 + LDS.64 weight per FFMA2", 16, c);
    run<3>("FFMA2 x8 per 2 LDS.128", 16, c);
    run<4>("FFMA x16 per 4 LDS.128", 16, c);
  }
  return 0;
}
