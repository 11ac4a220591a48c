// Micro-benchmark of the sm_100a issue/data pipes the small-channel conv kernels lean on:
// FP32 FMA (scalar / packed), shared-memory loads under different lane patterns, shuffles.
// Reports warp-instructions per clock per SM (clock64 ticks calibrated against globaltimer).
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITER = 2048;
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
}
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, const float* in, long long* clk, int stride) {
  __shared__ __align__(16) float ws[4096];
  for (int i = threadIdx.x; i < 4096; i += 256) ws[i] = in[i & 63];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = in[i] + threadIdx.x;
  const float b = in[16], c = in[17];
  float2 a2[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a2[i] = make_float2(a[i], a[i] + 1.f);
  const float2 b2 = make_float2(b, c), c2 = make_float2(c, b);
  int off = (lane * stride) & 2047;      // float index
  long long t0 = clock64();
  unsigned long long g0 = gtimer();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a2[i] = __ffma2_rn(a2[i], b2, c2);
    } else if (MODE == 2) {          // 16 LDS.32
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] += ws[off + i * 64 + (it & 1)];
    } else if (MODE == 3) {          // 16 LDS.128
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 f = *reinterpret_cast<const float4*>(&ws[(off + i * 128 + (it & 1) * 4) & 4092]);
        a[i] += (f.x + f.y) + (f.z + f.w);
      }
    } else if (MODE == 4) {          // 16 SHFL
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = __shfl_down_sync(0xffffffffu, a[i], 1) + b;
    } else if (MODE == 5) {          // 16 LDS.64
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float2 f = *reinterpret_cast<const float2*>(&ws[(off + i * 128 + (it & 1) * 2) & 4094]);
        a[i] += f.x + f.y;
      }
    } else if (MODE == 6) {          // 4 LDS.32 + 32 FFMA (8 FMA per loaded float)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float v = ws[off + i * 64 + (it & 1)];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[(i * 8 + j) & 15] = fmaf(v, b, a[(i * 8 + j) & 15]);
      }
    } else if (MODE == 7) {          // 8 LDS.32 + 32 FFMA (4 FMA per loaded float)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float v = ws[off + i * 64 + (it & 1)];
#pragma unroll
        for (int j = 0; j < 4; ++j) a[(i * 4 + j) & 15] = fmaf(v, b, a[(i * 4 + j) & 15]);
      }
    }
  }
  long long t1 = clock64();
  unsigned long long g1 = gtimer();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i] + a2[i].x + a2[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) { clk[2 * blockIdx.x] = t1 - t0; clk[2 * blockIdx.x + 1] = (long long)(g1 - g0); }
}
template <int MODE>
void run(const char* name, int inst_per_iter, int stride) {
  const int ctas_per_sm = 4, nb = 148 * ctas_per_sm;
  float *out, *in; long long* clk;
  cudaMalloc(&out, nb * 256 * 4); cudaMalloc(&in, 1024); cudaMalloc(&clk, nb * 16);
  cudaMemset(in, 0, 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<nb, 256>>>(out, in, clk, stride);
  cudaEventRecord(e0);
  k<MODE><<<nb, 256>>>(out, in, clk, stride);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  static long long h[148 * 8];
  cudaMemcpy(h, clk, nb * 16, cudaMemcpyDeviceToHost);
  double ck = 0, ns = 0; for (int i = 0; i < nb; ++i) { ck += h[2 * i]; ns += h[2 * i + 1]; } ck /= nb; ns /= nb;
  const double ghz = ck / ns;
  const double winst = (double)ITER * inst_per_iter * 8 * ctas_per_sm;     // warp-instructions per SM
  printf("%-46s stride=%2d  clk/ns=%.3f  kernel %.3f ms  cta %.0f clk | %.2f warp-inst/clk/SM (cta clock), %.2f (event time @ clk/ns)\n",
         name, stride, ghz, ms, ck, winst / ck, winst / (ms * 1e6 * ghz));
  cudaFree(out); cudaFree(in); cudaFree(clk);
}
int main() {
  run<0>("FFMA x16", 16, 0);
  run<1>("FFMA2 x16", 16, 0);
  run<2>("LDS.32 lanes consecutive", 16, 1);
  run<2>("LDS.32 lanes stride 8 floats (4 banks groups)", 16, 8);
  run<2>("LDS.32 broadcast", 16, 0);
  run<3>("LDS.128 lanes consecutive 16B", 16, 4);
  run<3>("LDS.128 lanes stride 32B", 16, 8);
  run<3>("LDS.128 broadcast", 16, 0);
  run<5>("LDS.64 lanes consecutive 8B", 16, 2);
  run<5>("LDS.64 broadcast", 16, 0);
  run<4>("SHFL.DOWN", 16, 0);
  run<6>("4 LDS.32 + 32 FFMA", 36, 1);
  run<7>("8 LDS.32 + 32 FFMA", 40, 1);
  return 0;
}
