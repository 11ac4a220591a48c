// CUDA-core versions of the two tap GEMMs in gemm_tc.cu, same operands and
// semantics (bf16 hi/lo pairs joined to fp32, fp32 FMA).  Used for layer shapes
// the tensor-core tiles do not cover (channel counts that are not multiples of
// 64) and as the on-device cross-check of the tcgen05 kernels in the GPU tests.
#include "common.cuh"

namespace dmc {

struct SimtTaps {
  int ntaps;
  int shift[16];
  int phase[16];
  int bsel[16];
};

// 64 x 64 output tile, 256 threads, 4 x 4 per thread, K step 16.
__device__ __forceinline__ void tile_fma(const float (*As)[68], const float (*Bs)[68], float (&acc)[4][4],
                                         int ty, int tx) {
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    float a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

__global__ void __launch_bounds__(256)
simt_tap_gemm_kernel(const bf16* __restrict__ Ah, const bf16* __restrict__ Al, long a_rows, int K,
                     const bf16* __restrict__ Bh, const bf16* __restrict__ Bl, int N,
                     float* __restrict__ D, long M, int ldD, int Hp, int Wp, SimtTaps taps,
                     double* __restrict__ stats) {
  __shared__ float As[16][68];
  __shared__ float Bs[16][68];
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  const long m0 = (long)blockIdx.x * 64;
  const int n0 = blockIdx.y * 64;
  float acc[4][4] = {};
  const int lr = tid / 4;          // 0..63 : row within tile
  const int lk = (tid % 4) * 4;    // 0,4,8,12 : k offset
  for (int t = 0; t < taps.ntaps; ++t) {
    const long arow = m0 + lr + taps.shift[t];
    const bool a_ok = arow >= 0 && arow < a_rows;
    const bf16* ah = Ah + ((long)taps.phase[t] * a_rows + (a_ok ? arow : 0)) * K;
    const bf16* al = Al ? Al + ((long)taps.phase[t] * a_rows + (a_ok ? arow : 0)) * K : nullptr;
    const int brow = n0 + lr;
    const bool b_ok = brow < N;
    const bf16* bh = Bh + ((long)taps.bsel[t] * N + (b_ok ? brow : 0)) * K;
    const bf16* bl = Bl + ((long)taps.bsel[t] * N + (b_ok ? brow : 0)) * K;
    for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + lk + j;
        As[lk + j][lr] = (a_ok && k < K) ? (al ? join_bf16(ah[k], al[k]) : __bfloat162float(ah[k])) : 0.f;
        Bs[lk + j][lr] = (b_ok && k < K) ? join_bf16(bh[k], bl[k]) : 0.f;
      }
      __syncthreads();
      tile_fma(As, Bs, acc, ty, tx);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long q = m0 + ty * 4 + i;
    if (q >= M) continue;
    const bool keep = Hp == 0 || interior(q, Hp, Wp);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) {
        const float v = keep ? acc[i][j] : 0.f;
        D[q * ldD + n] = v;
        if (stats && keep) {
          atomicAdd(stats + n, (double)v);
          atomicAdd(stats + N + n, (double)v * v);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256)
simt_wgrad_kernel(const bf16* __restrict__ Gh, const bf16* __restrict__ Gl, long P, int Cout,
                  const bf16* __restrict__ Xh, const bf16* __restrict__ Xl, int Cin,
                  float* __restrict__ dW, SimtTaps taps, int n_tiles, long rows_per_split,
                  int oihw_taps) {
  __shared__ float As[16][68];
  __shared__ float Bs[16][68];
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  const int t = blockIdx.z;
  const int m0 = (blockIdx.y / n_tiles) * 64, n0 = (blockIdx.y % n_tiles) * 64;
  const long q0 = (long)blockIdx.x * rows_per_split;
  long q1 = q0 + rows_per_split;
  if (q1 > P) q1 = P;
  const int shift = taps.shift[t];
  const bf16* xh = Xh + (long)taps.phase[t] * P * Cin;
  const bf16* xl = Xl + (long)taps.phase[t] * P * Cin;
  float acc[4][4] = {};
  const int lk = tid / 16;          // 0..15 : pixel within chunk
  const int lc = (tid % 16) * 4;    // 0..60 : channel offset
  for (long q = q0; q < q1; q += 16) {
    const long gq = q + lk;
    const long xq = gq + shift;
    const bool g_ok = gq < q1;
    const bool x_ok = g_ok && xq >= 0 && xq < P;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = m0 + lc + j, ci = n0 + lc + j;
      As[lk][lc + j] = (g_ok && co < Cout)
                           ? (Gl ? join_bf16(Gh[gq * Cout + co], Gl[gq * Cout + co]) : __bfloat162float(Gh[gq * Cout + co]))
                           : 0.f;
      Bs[lk][lc + j] = (x_ok && ci < Cin) ? join_bf16(xh[xq * Cin + ci], xl[xq * Cin + ci]) : 0.f;
    }
    __syncthreads();
    tile_fma(As, Bs, acc, ty, tx);
    __syncthreads();
  }
  float* w = dW + (oihw_taps ? (long)taps.bsel[t] : (long)taps.bsel[t] * Cout * Cin);
  const int wstep = oihw_taps ? oihw_taps : 1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = m0 + ty * 4 + i;
    if (co >= Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = n0 + tx * 4 + j;
      if (ci < Cin) atomicAdd(w + ((long)co * Cin + ci) * wstep, acc[i][j]);
    }
  }
}

static void fill(SimtTaps& tt, int ntaps, const int* shift, const int* phase, const int* bsel) {
  tt.ntaps = ntaps;
  for (int i = 0; i < 16; ++i) {
    tt.shift[i] = i < ntaps ? shift[i] : 0;
    tt.phase[i] = i < ntaps ? phase[i] : 0;
    tt.bsel[i] = i < ntaps ? bsel[i] : 0;
  }
}

}  // namespace dmc

using namespace dmc;

extern "C" int dmc_simt_tap_gemm(const void* A_hi, const void* A_lo, int a_phases, long a_rows, int K,
                                 const void* B_hi, const void* B_lo, int b_slices, int N, float* D,
                                 long M, int ldD, int Hp, int Wp, int ntaps, const int* shift,
                                 const int* phase, const int* bsel, double* stats, const float* bw_Y,
                                 const void* bw_act_hi, const float* bw_gb, const float* bw_mean,
                                 const float* bw_invstd, void* stream) {
  DMC_REQUIRE(ntaps >= 1 && ntaps <= 16, "simt_tap_gemm: ntaps=%d", ntaps);
  DMC_REQUIRE(bw_Y == nullptr, "simt_tap_gemm: the fused BN-backward epilogue exists only in the tensor-core kernel");
  (void)bw_act_hi; (void)bw_gb; (void)bw_mean; (void)bw_invstd;
  DMC_REQUIRE(K > 0 && N > 0 && M > 0 && ldD >= N, "simt_tap_gemm: bad shape");
  for (int i = 0; i < ntaps; ++i)
    DMC_REQUIRE(phase[i] >= 0 && phase[i] < a_phases && bsel[i] >= 0 && bsel[i] < b_slices,
                "simt_tap_gemm: tap %d out of range", i);
  SimtTaps tt;
  fill(tt, ntaps, shift, phase, bsel);
  dim3 grid((unsigned)cdiv(M, 64), (unsigned)cdiv(N, 64));
  simt_tap_gemm_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const bf16*)A_hi, (const bf16*)A_lo, a_rows, K, (const bf16*)B_hi, (const bf16*)B_lo, N, D, M,
      ldD, Hp, Wp, tt, stats);
  return dmc_check_launch("simt_tap_gemm_kernel");
}

extern "C" int dmc_simt_wgrad(const void* G_hi, const void* G_lo, long P, int Cout, const void* X_hi,
                              const void* X_lo, int x_phases, int Cin, float* dW, int ntaps,
                              const int* shift, const int* phase, const int* bsel, int oihw_taps,
                              float* workspace, long workspace_floats, void* stream) {
  (void)workspace; (void)workspace_floats;      // the CUDA-core twin always accumulates atomically
  DMC_REQUIRE(ntaps >= 1 && ntaps <= 16, "simt_wgrad: ntaps=%d", ntaps);
  for (int i = 0; i < ntaps; ++i)
    DMC_REQUIRE(phase[i] >= 0 && phase[i] < x_phases, "simt_wgrad: tap %d phase out of range", i);
  SimtTaps tt;
  fill(tt, ntaps, shift, phase, bsel);
  const int m_tiles = (int)cdiv(Cout, 64), n_tiles = (int)cdiv(Cin, 64);
  long splits = cdiv(148L * 8, (long)m_tiles * n_tiles * ntaps);
  if (splits < 1) splits = 1;
  long rows = cdiv(cdiv(P, splits), 16) * 16;
  if (rows < 256) rows = 256;
  splits = cdiv(P, rows);
  dim3 grid((unsigned)splits, (unsigned)(m_tiles * n_tiles), (unsigned)ntaps);
  simt_wgrad_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const bf16*)G_hi, (const bf16*)G_lo, P, Cout, (const bf16*)X_hi, (const bf16*)X_lo, Cin, dW, tt,
      n_tiles, rows, oihw_taps);
  return dmc_check_launch("simt_wgrad_kernel");
}
