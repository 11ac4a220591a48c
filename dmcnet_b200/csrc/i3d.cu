// Kernels either side of the tap GEMMs of the I3D classifier (code/dmcnet_I3D/network/i3d.py:299-601).
//
// 3-D maps are pixel-major [clips][T+1][H+1][W+1][C] with the shared zero ring of common.cuh on the low
// side of every dimension (Hp arguments of the pixelwise / GEMM kernels carry the temporal extent, see
// dmc_pack_hp); activations are bf16 hi/lo plane pairs, raw conv outputs and gradients fp32.  A 3x3x3
// "SAME" convolution is then a 27-tap GEMM with flat row shifts dt*Hp*Wp + dh*Wp + dw, a 1x1x1
// convolution a one-tap GEMM (csrc/gemm_tc.cu).  What is left for this file:
//   * the sample layout [B][7][T][H][W] -> per-frame planar mv / residual / flow,
//   * the 7x7x7 stride-2 stem: spatial patches of every frame (K = 98 -> 128) for a 7-tap temporal GEMM, and
//     the transpose of that gather,
//   * MaxPool3dTFPadding (TF "SAME": zeros appended on the HIGH side, ceil_mode) forward / backward,
//   * AvgPool3d((2,7,7)) + the mean over the remaining temporal positions, folded into one weighted mean,
//   * dropout mask multiply, SGD with Nesterov momentum over the flat bucket (train_model.py:129-142).
#include "common.cuh"

namespace dmc {

static unsigned grid_1d(long n, int block, long cap = 148L * 16) {
  long g = cdiv(n, block);
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

// ------------------------------------------------------------------ sample layout
// data [B][Cd][T][HW] (Cd = 5 or 7: mv 2 | residual 3 | flow 2, code/dmcnet_I3D/train/model.py:139-158)
// -> mv [B*T][2][HW], res [B*T][3][HW], flow [B*T][2][HW] (flow may be NULL).
__global__ void __launch_bounds__(256)
i3d_unpack_kernel(const float* __restrict__ data, int B, int Cd, int T, int HW4, float* __restrict__ mv,
                  float* __restrict__ res, float* __restrict__ flow) {
  const long total = (long)B * Cd * T * HW4;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW4);
    long r = i / HW4;
    const int t = (int)(r % T);
    r /= T;
    const int c = (int)(r % Cd);
    const int b = (int)(r / Cd);
    const float4 v = reinterpret_cast<const float4*>(data)[i];
    const long f = (long)b * T + t;
    float* dst;
    if (c < 2) dst = mv + (f * 2 + c) * (long)HW4 * 4;
    else if (c < 5) dst = res + (f * 3 + (c - 2)) * (long)HW4 * 4;
    else if (flow) dst = flow + (f * 2 + (c - 5)) * (long)HW4 * 4;
    else continue;
    reinterpret_cast<float4*>(dst)[p] = v;
  }
}

// ------------------------------------------------------------------ stem operand
// Conv3d(2, 64, 7, stride 2) with TF "SAME" padding (2 zeros in front, 3 behind, i3d.py:299-315) as a
// 7-tap GEMM over the temporal kernel index: the spatial 7x7x2 patch of every INPUT frame at the output's
// spatial positions is the K dimension (98 -> 128 columns), and output frame `to` reads input frames
// 2*to + kt - 2.  Frames are split by parity into two phases so that the stride-2 walk becomes a unit
// frame shift:  A2[phase][clip][tp][ho+1][wo+1][(kh*7+kw)*2+ci] = x[ci][t = 2*(tp-1)+phase][2ho+kh-2][2wo+kw-2],
// tp in 1..T/2; frames tp = 0 and tp = T/2+1 stay zero (the taps reach one frame back and two ahead).
// 3x less operand traffic than a full 686-column im2col, and the data gradient comes back in the same form.
__global__ void __launch_bounds__(256)
i3d_stem_patches_kernel(const float* __restrict__ x, long x_ns, int clips, int T, int H, int W,
                        bf16* __restrict__ A_hi, bf16* __restrict__ A_lo) {
  const int Tq = T / 2, Ho = H / 2, Wo = W / 2;
  const int Tp = Tq + 2, Hp = Ho + 1, Wp = Wo + 1;
  const unsigned rows = (unsigned)(clips * T * Ho * Wo);        // (input frame, ho, wo)
  const long phase_rows = (long)clips * Tp * Hp * Wp;
  const int kk = threadIdx.x & 63, sub = threadIdx.x >> 6;      // 49 of 64 lanes: (kh, kw); 4 rows per block
  if (kk >= 49) return;
  const int kh = kk / 7, kw = kk % 7;
  for (unsigned r = blockIdx.x * 4 + sub; r < rows; r += gridDim.x * 4) {
    const unsigned wo = r % (unsigned)Wo, r1 = r / (unsigned)Wo;
    const unsigned ho = r1 % (unsigned)Ho, f = r1 / (unsigned)Ho;      // f = clip * T + t
    const unsigned t = f % (unsigned)T, n = f / (unsigned)T;
    const int h = 2 * (int)ho + kh - 2, w = 2 * (int)wo + kw - 2;
    float v0 = 0.f, v1 = 0.f;
    if (h >= 0 && h < H && w >= 0 && w < W) {
      const float* f0 = x + (long)f * x_ns + (long)h * W + w;
      v0 = __ldg(f0);
      v1 = __ldg(f0 + (long)H * W);
    }
    bf16 h0, l0, h1, l1;
    split_bf16(v0, h0, l0);
    split_bf16(v1, h1, l1);
    __nv_bfloat162 hh, ll;
    hh.x = h0; hh.y = h1; ll.x = l0; ll.y = l1;
    const long row = (t & 1u) * phase_rows + (((long)n * Tp + (t >> 1) + 1) * Hp + ho + 1) * Wp + wo + 1;
    *reinterpret_cast<__nv_bfloat162*>(A_hi + row * 128 + kk * 2) = hh;
    *reinterpret_cast<__nv_bfloat162*>(A_lo + row * 128 + kk * 2) = ll;
  }
}

// Transpose for the data gradient: dX[clip*T + t][ci][h][w] (+)= sum over the (kh, kw) whose output position
// ((h+2-kh)/2, (w+2-kw)/2) is integral and inside the map of dA2[t & 1][row(clip, t/2+1, ho+1, wo+1)][(kh*7+kw)*2+ci].
__global__ void __launch_bounds__(256)
i3d_stem_patches_bwd_kernel(const float* __restrict__ dA, int clips, int T, int H, int W, float* __restrict__ dX,
                            long dx_ns, int accumulate) {
  const int Tq = T / 2, Ho = H / 2, Wo = W / 2;
  const int Tp = Tq + 2, Hp = Ho + 1, Wp = Wo + 1;
  const long phase_rows = (long)clips * Tp * Hp * Wp;
  const unsigned total = (unsigned)(clips * T * H * W);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned w = i % (unsigned)W, r1 = i / (unsigned)W;
    const unsigned h = r1 % (unsigned)H, f = r1 / (unsigned)H;
    const unsigned t = f % (unsigned)T, n = f / (unsigned)T;
    const long base = (t & 1u) * phase_rows + (((long)n * Tp + (t >> 1) + 1) * Hp) * Wp;
    float a0 = 0.f, a1 = 0.f;
    for (int kh = (int)(h & 1u); kh < 7; kh += 2) {
      const int ho = ((int)h + 2 - kh) / 2;
      if ((int)h + 2 - kh < 0 || ho >= Ho) continue;
      for (int kw = (int)(w & 1u); kw < 7; kw += 2) {
        const int wo = ((int)w + 2 - kw) / 2;
        if ((int)w + 2 - kw < 0 || wo >= Wo) continue;
        const long row = base + (long)(ho + 1) * Wp + wo + 1;
        const float2 v = __ldg(reinterpret_cast<const float2*>(dA + row * 128 + (kh * 7 + kw) * 2));
        a0 += v.x;
        a1 += v.y;
      }
    }
    float* o0 = dX + (long)f * dx_ns + (long)h * W + w;
    float* o1 = o0 + (long)H * W;
    if (accumulate) { *o0 += a0; *o1 += a1; } else { *o0 = a0; *o1 = a1; }
  }
}

// ------------------------------------------------------------------ MaxPool3dTFPadding
struct Pool3 {
  int Ti, Hi, Wi, To, Ho, Wo;      // interior extents of the input / output maps
  int kt, kh, kw, st, sh, sw;      // window, stride
  int pt, ph, pw;                  // zeros in FRONT of each dimension (TF "SAME": pad_along / 2)
  int t_hi;                        // extra zero frames behind each clip of the INPUT map (the stem map has one)
};

// Work decomposition of the STRIDED forward pools (COMPACT in the kernel; the stride-1 pools and every backward
// keep the linear order, channel quads fastest).  Thread index -> (channel quad, column block, row, frame, clip):
// the low 8 bits (one CTA of 256 threads) are 16 channel quads x 2 column blocks x POOL_HT rows x POOL_TT
// frames (every channel count of I3D is a multiple of 64, so no lane idles; a half-warp still moves whole
// 128-byte lines).
constexpr int POOL_HT = 4, POOL_TT = 2;
struct PoolBlocks { int cchunks, wpairs, wblocks, hblocks, tblocks; long units; };
struct PoolUnit { int c4, wblk, h, t, n; };
__host__ __device__ inline PoolBlocks pool_blocks(int clips, int C4, int T, int H, int wblocks) {
  PoolBlocks b;
  b.cchunks = (C4 + 15) / 16; b.wblocks = wblocks; b.wpairs = (wblocks + 1) / 2;
  b.hblocks = (H + POOL_HT - 1) / POOL_HT; b.tblocks = (T + POOL_TT - 1) / POOL_TT;
  b.units = (long)clips * b.tblocks * b.hblocks * b.wpairs * b.cchunks * 256;
  return b;
}
__device__ __forceinline__ bool pool_unit(const PoolBlocks& b, unsigned i, int C4, int T, int H, PoolUnit& u) {
  static_assert(16 * 2 * POOL_HT * POOL_TT == 256, "one CTA = 256 threads");
  unsigned r = i >> 8;
  u.c4 = (int)(r % (unsigned)b.cchunks) * 16 + (int)(i & 15u); r /= (unsigned)b.cchunks;
  u.wblk = (int)(r % (unsigned)b.wpairs) * 2 + (int)((i >> 4) & 1u); r /= (unsigned)b.wpairs;
  u.h = (int)(r % (unsigned)b.hblocks) * POOL_HT + (int)((i >> 5) & (POOL_HT - 1)); r /= (unsigned)b.hblocks;
  u.t = (int)(r % (unsigned)b.tblocks) * POOL_TT + (int)((i >> 7) & (POOL_TT - 1));
  u.n = (int)(r / (unsigned)b.tblocks);
  return u.c4 < C4 && u.wblk < b.wblocks && u.h < H && u.t < T;
}

// 8-byte read-only load under a predicate (zeros otherwise) without a branch around it
__device__ __forceinline__ uint2 ldg_u2_if(const void* p, bool on) {
  uint2 v;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\tmov.u32 %0, 0;\n\tmov.u32 %1, 0;\n\t"
               "@p ld.global.nc.v2.u32 {%0, %1}, [%2];\n\t}"
               : "=r"(v.x), "=r"(v.y) : "l"(p), "r"((int)on));
  return v;
}

// Pool geometry: compile-time for the four pools I3D uses (i3d.py:452-476 and the inception branch), so the
// window loops unroll and every window code is an immediate; the Pool3 fields otherwise (KT = 0).
template <int KT, int KH, int KW, int ST, int SH, int SW>
struct PoolGeo {
  static constexpr bool FIXED = KT > 0;
  const Pool3& g;
  __device__ __forceinline__ int kt() const { return FIXED ? KT : g.kt; }
  __device__ __forceinline__ int kh() const { return FIXED ? KH : g.kh; }
  __device__ __forceinline__ int kw() const { return FIXED ? KW : g.kw; }
  __device__ __forceinline__ int st() const { return FIXED ? ST : g.st; }
  __device__ __forceinline__ int sh() const { return FIXED ? SH : g.sh; }
  __device__ __forceinline__ int sw() const { return FIXED ? SW : g.sw; }
  __device__ __forceinline__ int pt() const { return FIXED ? (KT > ST ? (KT - ST) / 2 : 0) : g.pt; }
  __device__ __forceinline__ int ph() const { return FIXED ? (KH > SH ? (KH - SH) / 2 : 0) : g.ph; }
  __device__ __forceinline__ int pw() const { return FIXED ? (KW > SW ? (KW - SW) / 2 : 0) : g.pw; }
};

// in hi/lo [clips][Ti+1][Hi+1][Wi+1][C] -> out hi/lo [clips][To+1][Ho+1][Wo+1][C] (interior rows only; the
// output buffers are allocated zeroed) and idx [rows_out][C] = window position (dt*kh + dh)*kw + dw of the
// maximum, 255 when a padding zero wins.  Scan order and strict '>' as ATen's max_pool3d: the first
// maximum wins.  Positions outside the map are the ConstantPad3d zeros of i3d.py:380-388.
// One thread = 4 channels x WB consecutive outputs of a row: every input of the union of their windows is
// loaded once (WB + 2 instead of 3 * WB positions per (dt, dh) at stride 1).  The kernel is bound by
// instruction issue (ncu: sm 70 %, DRAM 3 %), so the running maximum of a (output, channel) is three
// registers -- value, the hi/lo bf16 pair packed in one word, window code -- and an update is one compare
// and three selects.
template <int WB, int KT, int KH, int KW, int ST, int SH, int SW>
__global__ void __launch_bounds__(256)
maxpool3d_fwd_kernel(const bf16* __restrict__ in_hi, const bf16* __restrict__ in_lo, int clips, int C,
                     const Pool3 g, bf16* __restrict__ out_hi, bf16* __restrict__ out_lo,
                     uint8_t* __restrict__ idx) {
  const PoolGeo<KT, KH, KW, ST, SH, SW> q{g};
  const int kt = q.kt(), kh = q.kh(), kw = q.kw(), st = q.st(), sh = q.sh(), sw = q.sw();
  const int pt = q.pt(), ph = q.ph(), pw = q.pw();
  const int C4 = C / 4;
  const int wblocks = (g.Wo + WB - 1) / WB;
  const int Hpi = g.Hi + 1, Wpi = g.Wi + 1, Tpi = g.Ti + 1 + g.t_hi, Hpo = g.Ho + 1, Wpo = g.Wo + 1, Tpo = g.To + 1;
  const int nw = (WB - 1) * sw + kw;          // input columns under the WB windows
  // Strided pools: a CTA (256 threads) = 16 channel quads x 2 column blocks x POOL_HT rows x POOL_TT frames,
  // so the overlapping windows of h / t neighbours are re-read from L1 (measured 1.4-1.6x faster than channels
  // running over the whole CTA, where a block is a sliver of ONE row).  Stride-1 pools measured 2x SLOWER
  // that way and keep the linear order (channel quads fastest).
  constexpr bool COMPACT = ST > 1 || SH > 1;
  const PoolBlocks pb = pool_blocks(clips, C4, g.To, g.Ho, wblocks);
  const long units = COMPACT ? pb.units : (long)clips * g.To * g.Ho * wblocks * C4;
  // 32-bit index arithmetic (the host checks units < 2^32): 64-bit divisions cost more than the loads
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (unsigned)units; i += gridDim.x * blockDim.x) {
    int c, wo0, ho, to, n;
    if constexpr (COMPACT) {
      PoolUnit u;
      if (!pool_unit(pb, i, C4, g.To, g.Ho, u)) continue;
      c = u.c4 * 4; wo0 = u.wblk * WB; ho = u.h; to = u.t; n = u.n;
    } else {
      c = (int)(i % (unsigned)C4) * 4;
      unsigned r = i / (unsigned)C4;
      wo0 = (int)(r % (unsigned)wblocks) * WB; r /= (unsigned)wblocks;
      ho = (int)(r % (unsigned)g.Ho); r /= (unsigned)g.Ho;
      to = (int)(r % (unsigned)g.To);
      n = (int)(r / (unsigned)g.To);
    }
    float best[WB][4];
    uint32_t bhl[WB][4];                     // (hi bits << 16) | lo bits of the running maximum
    uint32_t bcode[WB][4];
#pragma unroll
    for (int j = 0; j < WB; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) { best[j][k] = -INFINITY; bhl[j][k] = 0u; bcode[j][k] = 255u; }
    const int w_lo = wo0 * sw - pw;
#pragma unroll
    for (int dt = 0; dt < kt; ++dt) {
      const int t = to * st + dt - pt;
#pragma unroll
      for (int dh = 0; dh < kh; ++dh) {
        const int h = ho * sh + dh - ph;
        const bool in_th = t >= 0 && t < g.Ti && h >= 0 && h < g.Hi;
        // element offset of column w_lo of the row; 32-bit (the host checks rows * C < 2^31 - 2^24) so that an
        // address is one multiply-add, not a rematerialised 64-bit product per load
        const int e0 = (((n * Tpi + t + 1) * Hpi + h + 1) * Wpi + 1 + w_lo) * C + c;
        const uint32_t cbase = (uint32_t)((dt * kh + dh) * kw);
#pragma unroll
        for (int iw = 0; iw < nw; ++iw) {
          const int w = w_lo + iw;
          const bool in = in_th && w >= 0 && w < g.Wi;
          const uint2 vh = ldg_u2_if(in_hi + (e0 + iw * C), in);
          const uint2 vl = ldg_u2_if(in_lo + (e0 + iw * C), in);
          const uint32_t p[4] = {__byte_perm(vl.x, vh.x, 0x5410), __byte_perm(vl.x, vh.x, 0x7632),
                                 __byte_perm(vl.y, vh.y, 0x5410), __byte_perm(vl.y, vh.y, 0x7632)};
          float v[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) v[k] = __uint_as_float(p[k] & 0xffff0000u) + __uint_as_float(p[k] << 16);
#pragma unroll
          for (int j = 0; j < WB; ++j) {
            const int dw = iw - j * sw;
            if (dw < 0 || dw >= kw) continue;
            const uint32_t code = in ? cbase + (uint32_t)dw : 255u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const bool up = v[k] > best[j][k];
              best[j][k] = up ? v[k] : best[j][k];
              bhl[j][k] = up ? p[k] : bhl[j][k];
              bcode[j][k] = up ? code : bcode[j][k];
            }
          }
        }
      }
    }
    const long qo = (((long)n * Tpo + to + 1) * Hpo + ho + 1) * Wpo + wo0 + 1;
#pragma unroll
    for (int j = 0; j < WB; ++j) {
      if (wo0 + j >= g.Wo) break;
      *reinterpret_cast<uint2*>(out_hi + (qo + j) * C + c) =
          make_uint2(__byte_perm(bhl[j][0], bhl[j][1], 0x7632), __byte_perm(bhl[j][2], bhl[j][3], 0x7632));
      *reinterpret_cast<uint2*>(out_lo + (qo + j) * C + c) =
          make_uint2(__byte_perm(bhl[j][0], bhl[j][1], 0x5410), __byte_perm(bhl[j][2], bhl[j][3], 0x5410));
      *reinterpret_cast<uint32_t*>(idx + (qo + j) * C + c) =
          bcode[j][0] | (bcode[j][1] << 8) | (bcode[j][2] << 16) | (bcode[j][3] << 24);
    }
  }
}

// dX[row_in][c] = add[row_in][c] + sum over the windows containing the position whose idx points at it of
// g[row_out][c]; every row of dX is written (ring rows: 0).  One thread = 4 channels x WB consecutive input
// columns: the idx words of the union of their windows are loaded once.
template <int WB, int KT, int KH, int KW, int ST, int SH, int SW>
__global__ void __launch_bounds__(256)
maxpool3d_bwd_kernel(const float* __restrict__ gout, const uint8_t* __restrict__ idx, int clips, int C,
                     const Pool3 g, const float* __restrict__ add, float* __restrict__ dX) {
  const PoolGeo<KT, KH, KW, ST, SH, SW> q{g};
  const int kt = q.kt(), kh = q.kh(), kw = q.kw(), st = q.st(), sh = q.sh(), sw = q.sw();
  const int pt = q.pt(), ph = q.ph(), pw = q.pw();
  const int C4 = C / 4;
  const int Hpi = g.Hi + 1, Wpi = g.Wi + 1, Tpi = g.Ti + 1 + g.t_hi, Hpo = g.Ho + 1, Wpo = g.Wo + 1, Tpo = g.To + 1;
  const int wblocks = (Wpi + WB - 1) / WB;
  const long units = (long)clips * Tpi * Hpi * wblocks * C4;
  // at most this many windows contain a position / overlap the WB columns, per dimension
  const int nto = (kt + st - 1) / st, nho = (kh + sh - 1) / sh, nwo = (WB + kw - 2) / sw + 1;
  const uint32_t kwu = (uint32_t)kw;
  // linear order (channel quads fastest): the compact CTA shape of the strided forward pools measured
  // 5-15 % slower here
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (unsigned)units; i += gridDim.x * blockDim.x) {
    const int c = (int)(i % (unsigned)C4) * 4;
    unsigned r = i / (unsigned)C4;
    const int wp0 = (int)(r % (unsigned)wblocks) * WB; r /= (unsigned)wblocks;
    const int hp = (int)(r % (unsigned)Hpi); r /= (unsigned)Hpi;
    const int tp = (int)(r % (unsigned)Tpi);
    const int n = (int)(r / (unsigned)Tpi);
    const long q0 = (((long)n * Tpi + tp) * Hpi + hp) * Wpi + wp0;
    float4 acc[WB];
#pragma unroll
    for (int j = 0; j < WB; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tp >= 1 && tp <= g.Ti && hp >= 1) {
      const int t = tp - 1, h = hp - 1;
      if (add) {
#pragma unroll
        for (int j = 0; j < WB; ++j)
          if (wp0 + j >= 1 && wp0 + j < Wpi) acc[j] = __ldg(reinterpret_cast<const float4*>(add + (q0 + j) * C + c));
      }
      // windows along w of the WB columns w0 .. w0 + WB - 1 (w = wp - 1): wo * sw - pw <= w <= wo * sw - pw + kw - 1
      const int w0 = wp0 - 1;
      int wo_lo = (w0 + pw - kw + 1 + sw - 1);
      wo_lo = wo_lo <= 0 ? 0 : wo_lo / sw;
      int wo_hi = (w0 + WB - 1 + pw) / sw;
      if (wo_hi >= g.Wo) wo_hi = g.Wo - 1;
      const int to_hi = (t + pt) / st, ho_hi = (h + ph) / sh;
#pragma unroll
      for (int a = 0; a < nto; ++a) {
        const int to = to_hi - a;
        const int dt = t + pt - to * st;                    // grows with a: past kt - 1 no window reaches t
        if (to < 0 || dt >= kt || to >= g.To) continue;
#pragma unroll
        for (int b = 0; b < nho; ++b) {
          const int ho = ho_hi - b;
          const int dh = h + ph - ho * sh;
          if (ho < 0 || dh >= kh || ho >= g.Ho) continue;
          const uint32_t cbase = (uint32_t)((dt * kh + dh) * kw);
          // 32-bit element offsets (the host checks rows * C < 2^31 - 2^24)
          const int eo = (((n * Tpo + to + 1) * Hpo + ho + 1) * Wpo + 1 + wo_lo) * C + c;
#pragma unroll
          for (int e = 0; e < nwo; ++e) {
            const int wo = wo_lo + e;
            if (wo > wo_hi) continue;
            const uint32_t id = __ldg(reinterpret_cast<const uint32_t*>(idx + (eo + e * C)));
            // the codes this output can hold for our columns: cbase + dw, dw = w - (wo * sw - pw) in [0, kw)
            const uint32_t d0 = (id & 0xffu) - cbase, d1 = ((id >> 8) & 0xffu) - cbase,
                           d2 = ((id >> 16) & 0xffu) - cbase, d3 = (id >> 24) - cbase;
            if (d0 >= kwu && d1 >= kwu && d2 >= kwu && d3 >= kwu) continue;
            const float4 gv = __ldg(reinterpret_cast<const float4*>(gout + (eo + e * C)));
            const int wbase = wo * sw - pw - w0;            // column index j of dw = 0
#pragma unroll
            for (int j = 0; j < WB; ++j) {
              const uint32_t dw = (uint32_t)(j - wbase);
              if (dw >= kwu || wp0 + j < 1 || wp0 + j >= Wpi) continue;
              if (d0 == dw) acc[j].x += gv.x;
              if (d1 == dw) acc[j].y += gv.y;
              if (d2 == dw) acc[j].z += gv.z;
              if (d3 == dw) acc[j].w += gv.w;
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < WB; ++j) {
      if (wp0 + j >= Wpi) break;
      *reinterpret_cast<float4*>(dX + (q0 + j) * C + c) = (wp0 + j >= 1) ? acc[j] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// ------------------------------------------------------------------ head
// AvgPool3d((2,7,7), stride 1) over a [T5][7][7] map leaves T5-1 temporal positions; the 1x1x1 logits
// convolution and the mean over them are linear, so mean_t'(conv(avg_t')) = conv(weighted mean):
// pooled[n][c] = sum_{t,h,w} wt(t) * x / ((T5-1) * 2 * H * W), wt(t) = number of windows containing t
// (i3d.py:484,521-525 with Unit3Dpy's squeeze / mean).
__device__ __forceinline__ float head_wt(int t, int T5) {
  const int lo = t - 1 > 0 ? t - 1 : 0, hi = t < T5 - 2 ? t : T5 - 2;
  return (float)(hi - lo + 1);
}

__global__ void __launch_bounds__(256)
i3d_head_pool_fwd_kernel(const bf16* __restrict__ hi, const bf16* __restrict__ lo, int T5, int H, int W,
                         int C, float* __restrict__ pooled) {
  const int n = blockIdx.x;
  const int Hp = H + 1, Wp = W + 1, Tp = T5 + 1;
  const float inv = 1.f / (float)((T5 - 1) * 2 * H * W);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int t = 0; t < T5; ++t) {
      float s = 0.f;
      for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w) {
          const long q = (((long)n * Tp + t + 1) * Hp + h + 1) * Wp + w + 1;
          s += join_bf16(hi[q * C + c], lo[q * C + c]);
        }
      acc += head_wt(t, T5) * s;
    }
    pooled[(long)n * C + c] = acc * inv;
  }
}

__global__ void __launch_bounds__(256)
i3d_head_pool_bwd_kernel(const float* __restrict__ dpooled, int clips, int T5, int H, int W, int C,
                         float* __restrict__ dX) {
  const int C4 = C / 4;
  const int Hp = H + 1, Wp = W + 1, Tp = T5 + 1;
  const long units = (long)clips * Tp * Hp * Wp * C4;
  const float inv = 1.f / (float)((T5 - 1) * 2 * H * W);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < units; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const long q = i / C4;
    long r = q;
    const int wp = (int)(r % Wp); r /= Wp;
    const int hp = (int)(r % Hp); r /= Hp;
    const int tp = (int)(r % Tp);
    const int n = (int)(r / Tp);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tp >= 1 && hp >= 1 && wp >= 1) {
      const float s = head_wt(tp - 1, T5) * inv;
      const float4 d = __ldg(reinterpret_cast<const float4*>(dpooled + (long)n * C + c));
      v = make_float4(d.x * s, d.y * s, d.z * s, d.w * s);
    }
    *reinterpret_cast<float4*>(dX + q * C + c) = v;
  }
}

__global__ void mul_kernel(const float* __restrict__ a, const float* __restrict__ m, long n, float* __restrict__ out) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    out[i] = a[i] * m[i];
}

__global__ void axpy_kernel(float* __restrict__ y, const float* __restrict__ x, float a, long n4) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 yv = reinterpret_cast<float4*>(y)[i];
    const float4 xv = reinterpret_cast<const float4*>(x)[i];
    yv.x = fmaf(a, xv.x, yv.x); yv.y = fmaf(a, xv.y, yv.y); yv.z = fmaf(a, xv.z, yv.z); yv.w = fmaf(a, xv.w, yv.w);
    reinterpret_cast<float4*>(y)[i] = yv;
  }
}

__global__ void add_i64_kernel(long long* __restrict__ p, int n, long long v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] += v;
}

// ------------------------------------------------------------------ SGD, Nesterov momentum
struct SgdChunk { int offset, count, tensor, pad; };     // same table as dmc_adam_step

// torch.optim.SGD(momentum, nesterov=True, dampening=0), `_single_tensor_sgd`:
//   g += wd * p;  buf = momentum * buf + g  (first step: buf = g, identical with buf = 0);
//   p -= lr * (g + momentum * buf)
__global__ void __launch_bounds__(256)
sgd_nesterov_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf,
                    const SgdChunk* __restrict__ chunks, const float* __restrict__ hyper, float momentum,
                    float grad_scale) {
  const SgdChunk ch = chunks[blockIdx.x];
  const float lr = hyper[2 * ch.tensor], wd = hyper[2 * ch.tensor + 1];
  for (int i = threadIdx.x; i < ch.count; i += blockDim.x) {
    const long o = (long)ch.offset + i;
    float gk = fmaf(wd, p[o], g[o] * grad_scale);
    const float b = momentum * buf[o] + gk;
    buf[o] = b;
    p[o] = p[o] - lr * (gk + momentum * b);
  }
}

}  // namespace dmc

using namespace dmc;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

// data [B][Cd][T][H*W] fp32 (Cd = 5 or 7) -> per-frame planar mv [B*T][2][HW], res [B*T][3][HW] and, with
// Cd = 7, flow [B*T][2][HW] (code/dmcnet_I3D/train/model.py:139-158: input[:, :5] feeds the generator,
// input[:, 5:7] is the flow target; i3d.py:504-506 folds T into the batch).
extern "C" int dmc_i3d_unpack(const float* data, int B, int Cd, int T, long HW, float* mv, float* res,
                              float* flow, void* stream) {
  DMC_REQUIRE(data && mv && res && B > 0 && T > 0 && (Cd == 5 || Cd == 7) && HW % 4 == 0, "i3d_unpack: bad arguments");
  const long total = (long)B * Cd * T * (HW / 4);
  i3d_unpack_kernel<<<grid_1d(total, 256), 256, 0, ST(stream)>>>(data, B, Cd, T, (int)(HW / 4), mv, res,
                                                                Cd == 7 ? flow : nullptr);
  return dmc_check_launch("i3d_unpack_kernel");
}

// A_hi / A_lo [2][clips][T/2+2][H/2+1][W/2+1][128] (zero-initialised by the caller once; ring rows, the two
// zero frames of every clip and columns >= 98 are never written): the spatial 7x7x2 patches of every frame of
// x [clips*T][2][H][W] at the stride-2 output positions, frames split by parity (see the kernel).
extern "C" int dmc_i3d_stem_patches(const float* x, long x_ns, int clips, int T, int H, int W, void* A_hi,
                                    void* A_lo, void* stream) {
  DMC_REQUIRE(x && A_hi && A_lo && clips > 0, "i3d_stem_patches: null argument");
  DMC_REQUIRE(T % 2 == 0 && H % 2 == 0 && W % 2 == 0, "i3d_stem_patches: T=%d H=%d W=%d", T, H, W);
  const long rows = (long)clips * T * (H / 2) * (W / 2);
  DMC_REQUIRE(rows < (1L << 31), "i3d_stem_patches: too many rows");
  i3d_stem_patches_kernel<<<grid_1d(rows, 4, 148L * 16), 256, 0, ST(stream)>>>(x, x_ns, clips, T, H, W, (bf16*)A_hi,
                                                                              (bf16*)A_lo);
  return dmc_check_launch("i3d_stem_patches_kernel");
}

// Transpose of dmc_i3d_stem_patches for the data gradient: dX [clips*T][2][H][W] (frame stride dx_ns)
// (+)= gather(dA [2][rows][128] fp32).
extern "C" int dmc_i3d_stem_patches_bwd(const float* dA, int clips, int T, int H, int W, float* dX, long dx_ns,
                                        int accumulate, void* stream) {
  DMC_REQUIRE(dA && dX && clips > 0 && T % 2 == 0 && H % 2 == 0 && W % 2 == 0, "i3d_stem_patches_bwd: bad arguments");
  const long total = (long)clips * T * H * W;
  DMC_REQUIRE(total < (1L << 32), "i3d_stem_patches_bwd: too many elements");
  i3d_stem_patches_bwd_kernel<<<grid_1d(total, 256, 148L * 32), 256, 0, ST(stream)>>>(dA, clips, T, H, W, dX, dx_ns,
                                                                                     accumulate);
  return dmc_check_launch("i3d_stem_patches_bwd_kernel");
}

static int fill_pool(Pool3& g, const int* in_thw, const int* kernel, const int* stride, int t_hi = 0) {
  g.t_hi = t_hi;
  g.Ti = in_thw[0]; g.Hi = in_thw[1]; g.Wi = in_thw[2];
  g.kt = kernel[0]; g.kh = kernel[1]; g.kw = kernel[2];
  g.st = stride[0]; g.sh = stride[1]; g.sw = stride[2];
  const int k[3] = {g.kt, g.kh, g.kw}, s[3] = {g.st, g.sh, g.sw}, in[3] = {g.Ti, g.Hi, g.Wi};
  int out[3], pf[3];
  for (int d = 0; d < 3; ++d) {
    if (k[d] < 1 || s[d] < 1 || in[d] < 1 || k[d] > 3) return -1;
    const int pad = k[d] - s[d] > 0 ? k[d] - s[d] : 0;            // get_padding_shape, i3d.py:299-315
    pf[d] = pad / 2;
    // MaxPool3d(kernel, stride, ceil_mode=True) on the padded extent
    const int ext = in[d] + pad;
    int o = (ext - k[d] + s[d] - 1) / s[d] + 1;
    if ((o - 1) * s[d] >= ext) --o;                               // last window must start inside the input
    out[d] = o;
  }
  g.pt = pf[0]; g.ph = pf[1]; g.pw = pf[2];
  g.To = out[0]; g.Ho = out[1]; g.Wo = out[2];
  return 0;
}

// The four pools of I3D (i3d.py:452-476, Mixed branch 3) have kernels with the geometry compiled in;
// 0 = the generic kernel.
static int pool_variant(const Pool3& g) {
  const int k = g.kt * 100 + g.kh * 10 + g.kw, s = g.st * 100 + g.sh * 10 + g.sw;
  if (k == 133 && s == 122) return 1;
  if (k == 333 && s == 222) return 2;
  if (k == 222 && s == 222) return 3;
  if (k == 333 && s == 111) return 4;
  return 0;
}

// Output extents of MaxPool3dTFPadding(kernel, stride) on a [T][H][W] map (i3d.py:375-388).
extern "C" int dmc_maxpool3d_out_shape(const int* in_thw, const int* kernel, const int* stride, int* out_thw) {
  Pool3 g;
  DMC_REQUIRE(fill_pool(g, in_thw, kernel, stride) == 0, "maxpool3d: bad geometry");
  out_thw[0] = g.To; out_thw[1] = g.Ho; out_thw[2] = g.Wo;
  return DMC_OK;
}

// MaxPool3dTFPadding forward on pixel-major hi/lo maps (see the kernel); idx is uint8 [rows_out][C].
// in_t_hi: zero frames BEHIND each clip of the input map (0, or 1 for the stem map).
extern "C" int dmc_maxpool3d_fwd(const void* in_hi, const void* in_lo, int clips, int C, const int* in_thw,
                                 int in_t_hi, const int* kernel, const int* stride, void* out_hi, void* out_lo,
                                 void* idx, void* stream) {
  Pool3 g;
  DMC_REQUIRE(in_hi && in_lo && out_hi && out_lo && idx && clips > 0 && C % 4 == 0, "maxpool3d_fwd: bad arguments");
  DMC_REQUIRE(fill_pool(g, in_thw, kernel, stride, in_t_hi) == 0 && in_t_hi >= 0, "maxpool3d_fwd: bad geometry");
  const bool compact = g.st > 1 || g.sh > 1;        // must match COMPACT of the kernel variant (0 = generic: linear)
  const int variant = pool_variant(g);
  const long units = (variant != 0 && compact) ? pool_blocks(clips, C / 4, g.To, g.Ho, (int)cdiv(g.Wo, 4)).units
                                               : (long)clips * g.To * g.Ho * cdiv(g.Wo, 4) * (C / 4);
  DMC_REQUIRE(units < (1L << 32), "maxpool3d_fwd: map too large");
  DMC_REQUIRE((long)clips * (g.Ti + 2 + g.t_hi) * (g.Hi + 1) * (g.Wi + 1) * C < (1L << 31) - (1L << 24),
              "maxpool3d_fwd: input map too large for 32-bit element offsets");
  const unsigned grid = grid_1d(units, 256, 148L * 32);
#define DMC_POOL_FWD(...)                                                                                   \
  maxpool3d_fwd_kernel<4, __VA_ARGS__><<<grid, 256, 0, ST(stream)>>>(                                       \
      (const bf16*)in_hi, (const bf16*)in_lo, clips, C, g, (bf16*)out_hi, (bf16*)out_lo, (uint8_t*)idx)
  switch (variant) {
    case 1: DMC_POOL_FWD(1, 3, 3, 1, 2, 2); break;
    case 2: DMC_POOL_FWD(3, 3, 3, 2, 2, 2); break;
    case 3: DMC_POOL_FWD(2, 2, 2, 2, 2, 2); break;
    case 4: DMC_POOL_FWD(3, 3, 3, 1, 1, 1); break;
    default: DMC_POOL_FWD(0, 0, 0, 0, 0, 0);
  }
#undef DMC_POOL_FWD
  return dmc_check_launch("maxpool3d_fwd_kernel");
}

// Backward: dX [rows_in][C] = add (may be NULL) + routed gradient; every row written (ring rows zero).
extern "C" int dmc_maxpool3d_bwd(const float* gout, const void* idx, int clips, int C, const int* in_thw,
                                 int in_t_hi, const int* kernel, const int* stride, const float* add, float* dX,
                                 void* stream) {
  Pool3 g;
  DMC_REQUIRE(gout && idx && dX && clips > 0 && C % 4 == 0, "maxpool3d_bwd: bad arguments");
  DMC_REQUIRE(fill_pool(g, in_thw, kernel, stride, in_t_hi) == 0 && in_t_hi >= 0, "maxpool3d_bwd: bad geometry");
  const long units = (long)clips * (g.Ti + 1 + g.t_hi) * (g.Hi + 1) * cdiv(g.Wi + 1, 4) * (C / 4);
  DMC_REQUIRE(units < (1L << 32), "maxpool3d_bwd: map too large");
  DMC_REQUIRE((long)clips * (g.To + 1) * (g.Ho + 1) * (g.Wo + 1) * C < (1L << 31) - (1L << 24),
              "maxpool3d_bwd: output map too large for 32-bit element offsets");
  const unsigned grid = grid_1d(units, 256, 148L * 32);
#define DMC_POOL_BWD(...)                                                                                   \
  maxpool3d_bwd_kernel<4, __VA_ARGS__><<<grid, 256, 0, ST(stream)>>>(gout, (const uint8_t*)idx, clips, C, g, \
                                                                     add, dX)
  switch (pool_variant(g)) {
    case 1: DMC_POOL_BWD(1, 3, 3, 1, 2, 2); break;
    case 2: DMC_POOL_BWD(3, 3, 3, 2, 2, 2); break;
    case 3: DMC_POOL_BWD(2, 2, 2, 2, 2, 2); break;
    case 4: DMC_POOL_BWD(3, 3, 3, 1, 1, 1); break;
    default: DMC_POOL_BWD(0, 0, 0, 0, 0, 0);
  }
#undef DMC_POOL_BWD
  return dmc_check_launch("maxpool3d_bwd_kernel");
}

// pooled [clips][C] = AvgPool3d((2,7,7)) followed by the mean over the remaining temporal positions of
// the hi/lo map [clips][T5+1][H+1][W+1][C] (see head_wt).
extern "C" int dmc_i3d_head_pool_fwd(const void* hi, const void* lo, int clips, int T5, int H, int W, int C,
                                     float* pooled, void* stream) {
  DMC_REQUIRE(hi && lo && pooled && clips > 0 && T5 >= 2 && H >= 1 && W >= 1, "i3d_head_pool_fwd: bad arguments");
  i3d_head_pool_fwd_kernel<<<clips, 256, 0, ST(stream)>>>((const bf16*)hi, (const bf16*)lo, T5, H, W, C, pooled);
  return dmc_check_launch("i3d_head_pool_fwd_kernel");
}

extern "C" int dmc_i3d_head_pool_bwd(const float* dpooled, int clips, int T5, int H, int W, int C, float* dX,
                                     void* stream) {
  DMC_REQUIRE(dpooled && dX && clips > 0 && T5 >= 2 && C % 4 == 0, "i3d_head_pool_bwd: bad arguments");
  const long units = (long)clips * (T5 + 1) * (H + 1) * (W + 1) * (C / 4);
  i3d_head_pool_bwd_kernel<<<grid_1d(units, 256), 256, 0, ST(stream)>>>(dpooled, clips, T5, H, W, C, dX);
  return dmc_check_launch("i3d_head_pool_bwd_kernel");
}

// out = a * m elementwise (dropout mask of i3d.py:526, forward and backward).
extern "C" int dmc_mul(const float* a, const float* m, long n, float* out, void* stream) {
  DMC_REQUIRE(a && m && out && n > 0, "mul: bad arguments");
  mul_kernel<<<grid_1d(n, 256), 256, 0, ST(stream)>>>(a, m, n, out);
  return dmc_check_launch("mul_kernel");
}

// y += a * x over n floats (n a multiple of 4): gradient accumulation over the iter_size micro-batches of
// code/dmcnet_I3D/train/model.py:377-396 (loss.backward() adds into .grad).
extern "C" int dmc_axpy(float* y, const float* x, float a, long n, void* stream) {
  DMC_REQUIRE(y && x && n > 0 && n % 4 == 0, "axpy: bad arguments");
  axpy_kernel<<<grid_1d(n / 4, 256), 256, 0, ST(stream)>>>(y, x, a, n / 4);
  return dmc_check_launch("axpy_kernel");
}

// p[i] += v for an int64 array (the num_batches_tracked counters of every BatchNorm3d, kept contiguous).
extern "C" int dmc_add_i64(void* p, int n, long long v, void* stream) {
  DMC_REQUIRE(p && n > 0, "add_i64: bad arguments");
  add_i64_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ST(stream)>>>((long long*)p, n, v);
  return dmc_check_launch("add_i64_kernel");
}

// One torch.optim.SGD(momentum, nesterov=True) step over `nchunks` chunks of the flat bucket (the chunk /
// hyper tables of dmc_adam_step); grad_scale multiplies the gradient first (1 / iter_size,
// code/dmcnet_I3D/train/model.py:387-396).
extern "C" int dmc_sgd_nesterov_step(float* p, const float* g, float* buf, const void* chunks, int nchunks,
                                     const float* hyper, float momentum, float grad_scale, void* stream) {
  if (nchunks <= 0) return DMC_OK;
  DMC_REQUIRE(p && g && buf && chunks && hyper, "sgd_nesterov_step: null argument");
  sgd_nesterov_kernel<<<nchunks, 256, 0, ST(stream)>>>(p, g, buf, reinterpret_cast<const SgdChunk*>(chunks), hyper,
                                                       momentum, grad_scale);
  return dmc_check_launch("sgd_nesterov_kernel");
}
