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
  static_assert(COG % 2 == 0, "COG must be even");
  float2 acc[COG / 2];               // output-channel pairs -> packed FFMA2
#pragma unroll
  for (int g = 0; g < COG / 2; ++g) acc[g] = make_float2(0.f, 0.f);

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
          const float2 vv = make_float2(v, v);
          const float2* wp = reinterpret_cast<const float2*>(&w_s[c][r * KS + s][0]);
#pragma unroll
          for (int g = 0; g < COG / 2; ++g) acc[g] = __ffma2_rn(vv, wp[g], acc[g]);
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
    float v = ((g & 1) ? acc[g / 2].y : acc[g / 2].x) + (bias ? bias[co] : 0.f);
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

// ------------------------------------------------------------------ 3x3 stride-1, register-tiled
// Measured on B200 (tools/ubench/pipes.cu): the SM retires 1 shared-memory wavefront and
// 128 fp32 FMAs per clock; FFMA2 has the same FMA rate as FFMA but needs half the issue
// slots; a shuffle costs one wavefront; a 128-bit load whose lanes are 32 bytes apart is
// 2-way bank conflicted, halo loads 4 floats apart 4- to 8-way.  So the kernel is built
// to spend < 1 wavefront per 4 FFMA2:
//   * thread = 4 consecutive pixels x R rows x COG output channels; the 8 lanes of a
//     quarter-warp read one tile row as consecutive 128-bit words (conflict free),
//   * the left/right halo pixel comes from the neighbour lane by shuffle; only the two
//     edge lanes of a row touch shared memory for it,
//   * each input row is loaded once and feeds the (up to) three output rows it overlaps,
//   * the 9*COG weights of the current input channel live in registers.
// Tile = 32 x 32 output pixels; block = 8/R warps; input rows arrive by TMA (4-D tiled
// map, zero fill = conv padding) into a two-stage ring of CK channels.
// T0 / TN restrict the kernel to the taps [T0, T0 + TN) of every kernel row and column: the
// space-to-depth form of a stride-2 conv only has weights at taps {0,1} (forward) or {1,2}
// (its flipped data gradient), so 4 instead of 9 taps are computed (dense_conv.cu, space-to-depth).
template <int COG, int R, int CK, int NS, int T0 = 0, int TN = 3>
struct FwdV3Cfg {
  static constexpr int TH = 32, PITCH = 40, ROWS = TH + 2;
  static constexpr int NT = 32 * (8 / R);
  static constexpr int WPC = (TN * TN * COG + 3) / 4 * 4;      // weight floats per input channel
  static constexpr int BUF = (CK * ROWS * PITCH + 31) / 32 * 32;   // floats per stage (128-byte multiple)
  static int smem_bytes(int Cin) { return 128 + NS * BUF * 4 + Cin * WPC * 4 + 8 * NS + 64; }
};

// grid.x CTAs walk the (frame, tile) list with a fixed stride (one tile each when grid.x ==
// number of tiles) while the NS-stage TMA ring keeps running across tile boundaries; the
// weights of this CTA's output-channel group (grid.y) are staged once.
template <int COG, int R, int CK, int NS, int T0, int TN>
__global__ void __launch_bounds__(32 * (8 / R))
conv3x3_fwd_v3_kernel(const __grid_constant__ CUtensorMap in_map, int Cin, int ckb, int H, int W,
                      const float* __restrict__ w, const float* __restrict__ bias, int Cout,
                      float* __restrict__ out, long out_ns, float slope,
                      const float* __restrict__ mask, const float* __restrict__ add, long add_ns,
                      int accumulate, int tiles_x, int tiles_img, int total_tiles,
                      const float* __restrict__ act_src, long act_ns, int act_c1, float act_slope) {
  using C = FwdV3Cfg<COG, R, CK, NS, T0, TN>;
  constexpr int PITCH = C::PITCH, ROWS = C::ROWS, WPC = C::WPC, NT = C::NT;
  static_assert(T0 >= 0 && TN >= 1 && T0 + TN <= 3, "tap range");
  static_assert(COG % 2 == 0 && (R == 2 || R == 4), "COG even, R in {2,4}");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base_u32 = (smem_u32(smem_raw) + 127u) & ~127u;
  float* buf = reinterpret_cast<float*>(smem_raw + (base_u32 - smem_u32(smem_raw)));
  float* w_s = buf + NS * C::BUF;                      // [Cin][WPC]: tap-major, channel pairs
  const uint32_t bar0 = base_u32 + (NS * C::BUF + Cin * WPC) * 4;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sx = lane & 7;                             // 4-pixel strip within the 32-wide tile row
  const int y0 = (warp * 4 + (lane >> 3)) * R;         // first output row of this thread in the tile
  const int co0 = blockIdx.y * COG;
  const int nch = (Cin + ckb - 1) / ckb;
  const uint32_t stage_bytes = (uint32_t)(ckb * ROWS * PITCH * 4);
  const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total_steps = my_tiles * nch;

  // producer state (thread 0): next (tile, channel block) to request and its ring slot
  int p_left = total_steps, p_k = 0, p_sl = 0, p_n = 0, p_th0 = 0, p_tw0 = 0, p_tile = blockIdx.x;
  auto locate = [&]() {
    p_n = p_tile / tiles_img;
    const int t = p_tile - p_n * tiles_img;
    p_th0 = (t / tiles_x) * C::TH;
    p_tw0 = (t % tiles_x) * 32;
  };
  auto issue = [&]() {                                 // thread 0 only
    mbar_expect_tx(bar0 + 8 * p_sl, stage_bytes);
    tma_load_4d(base_u32 + p_sl * C::BUF * 4, &in_map, bar0 + 8 * p_sl, p_tw0 - 4, p_th0 - 1, p_k * ckb, p_n);
    --p_left;
    if (++p_sl == NS) p_sl = 0;
    if (++p_k == nch) {
      p_k = 0;
      p_tile += gridDim.x;
      if (p_left > 0) locate();
    }
  };
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < NS; ++i) mbar_init(bar0 + 8 * i, 1);
    fence_barrier_init();
    tma_prefetch_desc(&in_map);
    locate();
    for (int i = 0; i < NS - 1 && p_left > 0; ++i) issue();
  }
  // weights of this output group: thread = one (channel, output) filter; taps [T0, T0+TN)^2
  for (int i = tid; i < Cin * COG; i += NT) {
    const int c = i / COG, g = i - c * COG;
    float v[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) v[t] = 0.f;
    if (co0 + g < Cout) {
      const float* wp = w + ((long)(co0 + g) * Cin + c) * 9;
#pragma unroll
      for (int t = 0; t < 9; ++t) v[t] = wp[t];
    }
#pragma unroll
    for (int tr = 0; tr < TN; ++tr)
#pragma unroll
      for (int ts = 0; ts < TN; ++ts)
        w_s[c * WPC + (tr * TN + ts) * COG + g] = v[(tr + T0) * 3 + ts + T0];
  }
  if (WPC > TN * TN * COG)
    for (int c = tid; c < Cin; c += NT)
      for (int j = TN * TN * COG; j < WPC; ++j) w_s[c * WPC + j] = 0.f;
  float2 acc[R][4][COG / 2];
#pragma unroll
  for (int y = 0; y < R; ++y)
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int g = 0; g < COG / 2; ++g) acc[y][p][g] = make_float2(0.f, 0.f);
  __syncthreads();

  // edge lanes fetch the one halo pixel the shuffle cannot provide
  const bool edge = sx == 0 || sx == 7;
  const int edge_off = sx == 0 ? -1 : 4;
  const int thr_off = y0 * PITCH + 4 + 4 * sx;

  int k = 0, ti = 0, sl = 0;
  uint32_t ph = 0;
#pragma unroll 1
  for (int step = 0; step < total_steps; ++step) {
    if (tid == 0 && p_left > 0) issue();
    mbar_wait(bar0 + 8 * sl, ph);
    const float* in_s = buf + sl * C::BUF + thr_off;
    const int c0 = k * ckb;
    const int cmax = (Cin - c0) < ckb ? (Cin - c0) : ckb;
#pragma unroll 1
    for (int c = 0; c < cmax; ++c) {
      float wr[WPC];
      {
        const float4* wp = reinterpret_cast<const float4*>(w_s + (c0 + c) * WPC);
#pragma unroll
        for (int j = 0; j < WPC / 4; ++j) {
          const float4 f = wp[j];
          wr[4 * j] = f.x; wr[4 * j + 1] = f.y; wr[4 * j + 2] = f.z; wr[4 * j + 3] = f.w;
        }
      }
      const float* cp = in_s + c * (ROWS * PITCH);
#pragma unroll
      for (int rr = T0; rr < T0 + R + TN - 1; ++rr) {  // input row y0 - 1 + rr of the image tile
        const float* row = cp + rr * PITCH;
        const float4 f = *reinterpret_cast<const float4*>(row);
        float e = 0.f;
        if (edge) e = row[edge_off];
        float left = __shfl_up_sync(0xffffffffu, f.w, 1);
        float right = __shfl_down_sync(0xffffffffu, f.x, 1);
        if (sx == 0) left = e;
        if (sx == 7) right = e;
        const float v[6] = {left, f.x, f.y, f.z, f.w, right};
#pragma unroll
        for (int r = T0; r < T0 + TN; ++r) {
          const int y = rr - r;                        // output row fed through kernel row r
          if (y < 0 || y >= R) continue;
#pragma unroll
          for (int s3 = T0; s3 < T0 + TN; ++s3)
#pragma unroll
            for (int g = 0; g < COG / 2; ++g) {
              const int wi = ((r - T0) * TN + (s3 - T0)) * COG + 2 * g;
              const float2 wv = make_float2(wr[wi], wr[wi + 1]);
#pragma unroll
              for (int p = 0; p < 4; ++p)              // FFMA2 with a scalar-broadcast operand
                acc[y][p][g] = __ffma2_rn(make_float2(v[p + s3], v[p + s3]), wv, acc[y][p][g]);
            }
        }
      }
    }
    __syncthreads();          // this stage may be refilled at the next iteration
    if (++sl == NS) { sl = 0; ph ^= 1u; }
    if (++k < nch) continue;
    k = 0;
    // ---- tile finished: epilogue, then clear the accumulators
    const int tile = blockIdx.x + ti * gridDim.x;
    ++ti;
    const int n = tile / tiles_img, t = tile - n * tiles_img;
    const int th0 = (t / tiles_x) * C::TH, tw0 = (t % tiles_x) * 32;
    const int ox = tw0 + 4 * sx;
    const long hw = (long)H * W;
    const long pix = (long)(th0 + y0) * W + ox;
    float bs[COG], mk[COG];
#pragma unroll
    for (int g = 0; g < COG; ++g) {
      const bool ok = co0 + g < Cout;
      bs[g] = (bias && ok) ? bias[co0 + g] : 0.f;
      mk[g] = (mask && ok) ? mask[(long)n * Cout + co0 + g] : 1.f;
    }
#pragma unroll
    for (int y = 0; y < R; ++y) {
      const bool row_ok = th0 + y0 + y < H && ox < W;
#pragma unroll
      for (int g = 0; g < COG; ++g) {
        const int co = co0 + g;
        if (co < Cout && row_ok) {
          const long cpix = co * hw + pix + (long)y * W;
          float rv[4];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            float v = ((g & 1) ? acc[y][p][g / 2].y : acc[y][p][g / 2].x) + bs[g];
            v = v < 0.f ? v * slope : v;
            rv[p] = v * mk[g];
          }
          float4 f = make_float4(rv[0], rv[1], rv[2], rv[3]);   // W % 4 == 0 is required by the launcher
          if (add) {
            const float4 a = *reinterpret_cast<const float4*>(add + (long)n * add_ns + cpix);
            f.x += a.x; f.y += a.y; f.z += a.z; f.w += a.w;
          }
          float* op = out + (long)n * out_ns + cpix;
          if (accumulate) {
            const float4 a = *reinterpret_cast<const float4*>(op);
            f.x += a.x; f.y += a.y; f.z += a.z; f.w += a.w;
          }
          if (act_src && co < act_c1) {            // fused LeakyReLU backward on the finished slice
            const float4 a = *reinterpret_cast<const float4*>(act_src + (long)n * act_ns + cpix);
            f.x *= a.x > 0.f ? 1.f : act_slope; f.y *= a.y > 0.f ? 1.f : act_slope;
            f.z *= a.z > 0.f ? 1.f : act_slope; f.w *= a.w > 0.f ? 1.f : act_slope;
          }
          *reinterpret_cast<float4*>(op) = f;
        }
      }
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int g = 0; g < COG / 2; ++g) acc[y][p][g] = make_float2(0.f, 0.f);
    }
  }
}

template <int COG, int R, int CK, int NS, int T0 = 0, int TN = 3>
static int launch_fwd_v3(const float* in, long in_ns, int Cin, int H, int W, const float* w,
                         const float* bias, int Cout, float* out, long out_ns, float slope,
                         const float* mask, const float* add, long add_ns, int accumulate, int N,
                         cudaStream_t st, const float* act_src = nullptr, long act_ns = 0,
                         int act_c1 = 0, float act_slope = 1.f) {
  using C = FwdV3Cfg<COG, R, CK, NS, T0, TN>;
  const int ckb = Cin < CK ? Cin : CK;
  CUtensorMap map;
  const unsigned long long dims[4] = {(unsigned long long)W, (unsigned long long)H,
                                      (unsigned long long)Cin, (unsigned long long)N};
  const unsigned long long strides[3] = {(unsigned long long)W, (unsigned long long)H * W,
                                         (unsigned long long)in_ns};
  const unsigned int box[4] = {(unsigned)C::PITCH, (unsigned)C::ROWS, (unsigned)ckb, 1u};
  int rc = dmc_make_f32_map(&map, in, 4, dims, strides, box);
  if (rc) return rc;
  const int smem = C::smem_bytes(Cin);
  auto kern = conv3x3_fwd_v3_kernel<COG, R, CK, NS, T0, TN>;
  static int attr_bytes = 0;
  if (smem > attr_bytes) {
    const int want = smem > 64 * 1024 ? smem : 64 * 1024;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, want) != cudaSuccess)
      return dmc_check_launch("conv3x3_fwd_v3 smem attribute");
    attr_bytes = want;
  }
  const int tx32 = (int)cdiv(W, 32), tiles_img = tx32 * (int)cdiv(H, C::TH);
  const long total = (long)tiles_img * N;
  const int groups = (int)cdiv(Cout, COG);
  // One CTA per tile measured ~10% faster than a persistent grid of 148*occ CTAs here (the
  // hardware scheduler staggers the CTAs' prologues/epilogues); the kernel handles both.
  const long gx = total;
  dim3 grid((unsigned)gx, (unsigned)groups);
  kern<<<grid, C::NT, smem, st>>>(map, Cin, ckb, H, W, w, bias, Cout, out, out_ns, slope, mask, add,
                                  add_ns, accumulate, tx32, tiles_img, (int)total, act_src, act_ns,
                                  act_c1, act_slope);
  return dmc_check_launch("conv3x3_fwd_v3_kernel");
}

// wT[ci][co][2-r][2-s] = w[co][ci][r][s] for ci < ci_count: turns the stride-1 data
// gradient into a forward convolution of dY with wT.
__global__ void weight_flip_kernel(const float* __restrict__ w, int Cout, int Cin, int ci_count,
                                   float* __restrict__ wT) {
  const int n = ci_count * Cout * 9;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int t = i % 9, co = (i / 9) % Cout, ci = i / (9 * Cout);
    wT[i] = w[((long)co * Cin + ci) * 9 + (8 - t)];
  }
}

// 3x3 stride-1 weight gradient.  Warp = one input channel x COUT output channels
// (9*COUT register accumulators kept across ALL tiles the CTA visits), lane = one
// column of a 32 x 16 output tile, with a 3 x 3 register window sliding down the
// rows: 3 + COUT shared loads per 9*COUT FMAs.  Tiles (input with halo, dY) arrive
// by TMA into a two-stage ring.  One shuffle reduction + atomic flush per CTA.
template <int COUT>
__global__ void __launch_bounds__(256)
conv3x3_wgrad_v3_kernel(const __grid_constant__ CUtensorMap x_map, const __grid_constant__ CUtensorMap dy_map,
                        int Cin, int Cout, int H, int W, float* __restrict__ dW,
                        float* __restrict__ dbias, int N, int ci_groups, int xbox_c, int cpg) {
  constexpr int XP = 40, XR = 18, TR = 16;   // 4 + 32 + 4 columns: 16-byte aligned box start
  constexpr int DBUF = COUT * TR * 32;
  const int XBUF = (xbox_c * XR * XP + 31) & ~31, STAGE = XBUF + DBUF;   // floats; the CTA has xbox_c warps
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base_u32 = (smem_u32(smem_raw) + 127u) & ~127u;
  const float* buf = reinterpret_cast<const float*>(smem_raw + (base_u32 - smem_u32(smem_raw)));
  const uint32_t bar0 = base_u32 + 2 * STAGE * 4;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cig = blockIdx.y % ci_groups, cog = blockIdx.y / ci_groups;
  // input channels are dealt evenly to the groups (cpg <= 8 per CTA, one per warp): 33 channels
  // run as 7+7+7+6+6 instead of 8+8+8+8+1
  const int ci = warp < cpg ? cig * cpg + warp : Cin, co0 = cog * COUT;
  const int tiles_x = (W + 31) / 32, tiles_y = (H + TR - 1) / TR;
  const long items = (long)N * tiles_x * tiles_y;
  const uint32_t stage_bytes = (uint32_t)(xbox_c * XR * XP + DBUF) * 4;
  static_assert(COUT % 2 == 0, "COUT must be even");
  float2 acc[COUT / 2][3][3];       // output-channel pairs -> packed FFMA2
  float2 bs[COUT / 2];
#pragma unroll
  for (int c = 0; c < COUT / 2; ++c) {
    bs[c] = make_float2(0.f, 0.f);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) acc[c][r][s] = make_float2(0.f, 0.f);
  }
#define DMC_WG_ISSUE(item_, stage_)                                                   \
  {                                                                                   \
    const long it_ = (item_);                                                         \
    const int n_ = (int)(it_ / (tiles_x * tiles_y));                                  \
    const int t_ = (int)(it_ % (tiles_x * tiles_y));                                  \
    const int oh0_ = (t_ / tiles_x) * TR, ow0_ = (t_ % tiles_x) * 32;                 \
    const uint32_t b_ = bar0 + 8 * (stage_);                                          \
    const uint32_t dst_ = base_u32 + (stage_) * STAGE * 4;                            \
    mbar_expect_tx(b_, stage_bytes);                                                  \
    tma_load_4d(dst_, &x_map, b_, ow0_ - 4, oh0_ - 1, cig * cpg, n_);                 \
    tma_load_4d(dst_ + XBUF * 4, &dy_map, b_, ow0_, oh0_, co0, n_);                   \
  }
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    fence_barrier_init();
    tma_prefetch_desc(&x_map);
    tma_prefetch_desc(&dy_map);
    if ((long)blockIdx.x < items) DMC_WG_ISSUE(blockIdx.x, 0)
  }
  __syncthreads();
  int k = 0;
  for (long item = blockIdx.x; item < items; item += gridDim.x, ++k) {
    const long next = item + gridDim.x;
    if (tid == 0 && next < items) DMC_WG_ISSUE(next, (k + 1) & 1)
    mbar_wait(bar0 + 8 * (k & 1), (k >> 1) & 1);
    if (ci < Cin) {
      const float* xs = buf + (k & 1) * STAGE + warp * XR * XP + lane + 3;   // tile col 0 = x0 - 4
      const float* ds = buf + (k & 1) * STAGE + XBUF + lane;
      // 3 x 3 scalar window, rows rotate through three register slots (rows fully unrolled, so
      // the slot index is static and no register is ever moved); FFMA2 broadcasts the scalar
      float xw[3][3];
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        xw[0][s] = xs[s];
        xw[1][s] = xs[XP + s];
      }
#pragma unroll
      for (int y = 0; y < TR; ++y) {
#pragma unroll
        for (int s = 0; s < 3; ++s) xw[(y + 2) % 3][s] = xs[(y + 2) * XP + s];
#pragma unroll
        for (int c = 0; c < COUT / 2; ++c) {
          const float2 d = make_float2(ds[((2 * c) * TR + y) * 32], ds[((2 * c + 1) * TR + y) * 32]);
          bs[c] = __fadd2_rn(bs[c], d);
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < 3; ++s) {
              const float v = xw[(y + r) % 3][s];
              acc[c][r][s] = __ffma2_rn(make_float2(v, v), d, acc[c][r][s]);
            }
        }
      }
    }
    __syncthreads();
  }
  if (ci < Cin) {
#pragma unroll
    for (int c = 0; c < COUT; ++c) {
      const int co = co0 + c;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const float v = warp_sum((c & 1) ? acc[c / 2][r][s].y : acc[c / 2][r][s].x);
          if (lane == 0 && co < Cout) atomicAdd(dW + ((long)co * Cin + ci) * 9 + r * 3 + s, v);
        }
      if (dbias && ci == 0) {
        const float v = warp_sum((c & 1) ? bs[c / 2].y : bs[c / 2].x);
        if (lane == 0 && co < Cout) atomicAdd(dbias + co, v);
      }
    }
  }
}

#undef DMC_WG_ISSUE

template <int COUT>
static int launch_wgrad_v3(const float* in, long in_ns, int Cin, int H, int W, const float* dY,
                           long dy_ns, int Cout, float* dW, float* dbias, int N, cudaStream_t st) {
  // input channels are dealt evenly to ceil(Cin/8) groups; a CTA has one warp per channel of its
  // group (5..8 warps) and stages exactly those channels
  const int ci_groups = (int)cdiv(Cin, 8), co_groups = (int)cdiv(Cout, COUT);
  const int cpg = (int)cdiv(Cin, ci_groups);
  CUtensorMap xm, dm;
  {
    const unsigned long long dims[4] = {(unsigned long long)W, (unsigned long long)H,
                                        (unsigned long long)Cin, (unsigned long long)N};
    const unsigned long long str[3] = {(unsigned long long)W, (unsigned long long)H * W,
                                       (unsigned long long)in_ns};
    const unsigned int box[4] = {40u, 18u, (unsigned)cpg, 1u};
    int rc = dmc_make_f32_map(&xm, in, 4, dims, str, box);
    if (rc) return rc;
  }
  {
    const unsigned long long dims[4] = {(unsigned long long)W, (unsigned long long)H,
                                        (unsigned long long)Cout, (unsigned long long)N};
    const unsigned long long str[3] = {(unsigned long long)W, (unsigned long long)H * W,
                                       (unsigned long long)dy_ns};
    const unsigned int box[4] = {32u, 16u, (unsigned)COUT, 1u};
    int rc = dmc_make_f32_map(&dm, dY, 4, dims, str, box);
    if (rc) return rc;
  }
  const int STAGE = ((cpg * 18 * 40 + 31) & ~31) + COUT * 16 * 32;
  const int smem = 128 + 2 * STAGE * 4 + 64;
  auto kern = conv3x3_wgrad_v3_kernel<COUT>;
  static int attr_bytes = 0;
  if (smem > attr_bytes) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return dmc_check_launch("conv3x3_wgrad_v3 smem attribute");
    attr_bytes = smem;
  }
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * cpg, smem) != cudaSuccess || occ < 1)
    return dmc_check_launch("conv3x3_wgrad_v3 occupancy");
  const long items = (long)N * cdiv(W, 32) * cdiv(H, 16);
  // a few more CTAs than are resident: measured faster than an exact fit (shorter item lists
  // even out the tail)
  long gx = cdiv(148L * (occ > 3 ? occ : 3), (long)ci_groups * co_groups);
  if (gx < 1) gx = 1;
  if (gx > items) gx = items;
  dim3 grid((unsigned)gx, (unsigned)(ci_groups * co_groups));
  kern<<<grid, 32 * cpg, smem, st>>>(xm, dm, Cin, Cout, H, W, dW, dbias, N, ci_groups, cpg, cpg);
  return dmc_check_launch("conv3x3_wgrad_v3_kernel");
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
    if ((HW & 3) == 0) {
      for (long i = threadIdx.x * 4; i < HW; i += blockDim.x * 4) {
        const float4 v = *reinterpret_cast<const float4*>(p + i);
        a += (v.x + v.y) + (v.z + v.w);
        b += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
      }
    } else {
      for (long i = threadIdx.x; i < HW; i += blockDim.x) {
        const float v = p[i];
        a += v;
        b += v * v;
      }
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

// out = [relu](X*scale + shift), planar.  grid (chunks, C, N): one (image, channel) plane per
// block row, 128-bit accesses, no per-element index arithmetic.
__global__ void __launch_bounds__(256)
bn_apply_planar_kernel(const float* __restrict__ X, long x_ns, const float* __restrict__ scale,
                       const float* __restrict__ shift, long HW, int relu, float* __restrict__ out,
                       long o_ns) {
  const int c = blockIdx.y, n = blockIdx.z;
  const float sc = scale[c], sh = shift[c];
  const float* xp = X + (long)n * x_ns + (long)c * HW;
  float* op = out + (long)n * o_ns + (long)c * HW;
  const long i0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4, step = (long)gridDim.x * blockDim.x * 4;
  if ((HW & 3) == 0) {
    for (long i = i0; i < HW; i += step) {
      float4 v = *reinterpret_cast<const float4*>(xp + i);
      v.x = fmaf(v.x, sc, sh); v.y = fmaf(v.y, sc, sh); v.z = fmaf(v.z, sc, sh); v.w = fmaf(v.w, sc, sh);
      if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      *reinterpret_cast<float4*>(op + i) = v;
    }
  } else {
    for (long i = i0 / 4; i < HW; i += step / 4) {
      float v = fmaf(xp[i], sc, sh);
      op[i] = relu ? fmaxf(v, 0.f) : v;
    }
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
    if ((HW & 3) == 0) {
      for (long i = threadIdx.x * 4; i < HW; i += blockDim.x * 4) {
        const float4 d = *reinterpret_cast<const float4*>(g + i);
        const float4 x = *reinterpret_cast<const float4*>(p + i);
        a += (d.x + d.y) + (d.z + d.w);
        b += (d.x * (x.x - m) + d.y * (x.y - m) + d.z * (x.z - m) + d.w * (x.w - m)) * is;
      }
    } else {
      for (long i = threadIdx.x; i < HW; i += blockDim.x) {
        const float d = g[i];
        a += d;
        b += d * (p[i] - m) * is;
      }
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

// dX = gamma*invstd*(dZ - mean(dZ) - xhat*mean(dZ*xhat)) = A*dZ + B*X + K per channel;
// grid (chunks, C, N); block (0, c, 0) writes dgamma[c] / dbeta[c].
__global__ void __launch_bounds__(256)
bn_bwd_apply_planar_kernel(const float* __restrict__ dZ, long dz_ns, const float* __restrict__ X,
                           long x_ns, const float* __restrict__ mean, const float* __restrict__ invstd,
                           const float* __restrict__ gamma, const double* __restrict__ sums2,
                           double count, int C, long HW, float* __restrict__ dX, long dx_ns,
                           float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.y, n = blockIdx.z;
  const double inv_count = 1.0 / count;
  const float a = gamma[c] * invstd[c];
  const float m1 = (float)(sums2[c] * inv_count), m2 = (float)(sums2[C + c] * inv_count);
  const float b = -a * invstd[c] * m2;
  const float k = -a * m1 - b * mean[c];
  if (blockIdx.x == 0 && n == 0 && threadIdx.x == 0 && dgamma) {
    dbeta[c] = (float)sums2[c];
    dgamma[c] = (float)sums2[C + c];
  }
  const float* zp = dZ + (long)n * dz_ns + (long)c * HW;
  const float* xp = X + (long)n * x_ns + (long)c * HW;
  float* op = dX + (long)n * dx_ns + (long)c * HW;
  const long i0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4, step = (long)gridDim.x * blockDim.x * 4;
  if ((HW & 3) == 0) {
    for (long i = i0; i < HW; i += step) {
      const float4 z = *reinterpret_cast<const float4*>(zp + i);
      const float4 x = *reinterpret_cast<const float4*>(xp + i);
      float4 o;
      o.x = fmaf(a, z.x, fmaf(b, x.x, k)); o.y = fmaf(a, z.y, fmaf(b, x.y, k));
      o.z = fmaf(a, z.z, fmaf(b, x.z, k)); o.w = fmaf(a, z.w, fmaf(b, x.w, k));
      *reinterpret_cast<float4*>(op + i) = o;
    }
  } else {
    for (long i = i0 / 4; i < HW; i += step / 4) op[i] = fmaf(a, zp[i], fmaf(b, xp[i], k));
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

// ------------------------------------------------------------------ 3x3 stride-2 via space-to-depth
// A 3x3 stride-2 pad-1 convolution equals a 3x3 stride-1 pad-1 convolution of the space-to-depth
// input S[(pr,pc,ci)][i][j] = in[ci][2i+pr][2j+pc] with the weight
//   W3[co][(pr,pc,ci)][a][b] = w[co][ci][r(pr,a)][s(pc,b)],   r(1,0)=0, r(0,1)=1, r(1,1)=2, else 0
// (kernel row r reads input row 2y+r-1 = phase pr at i = y + a - 1).  The backward pass of the
// stride-2 discriminator layers runs on S through the fast stride-1 kernels: the weight
// gradient of W3 is gathered back into OIHW, the data gradient of S is re-interleaved.
__device__ __forceinline__ int s2_tap(int p, int a) {      // kernel index for (phase, window position)
  return p == 0 ? (a == 1 ? 1 : -1) : (a == 0 ? 0 : (a == 1 ? 2 : -1));
}

// S[n][p*C + c][H/2][W/2] <- in[n][c][H][W];  W % 8 == 0
__global__ void s2d_planar_kernel(const float* __restrict__ in, long in_ns, int C, int H, int W,
                                  float* __restrict__ out, long out_ns, long total8) {
  const int W8 = W / 8, Ho = H / 2, Wo = W / 2;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total8; i += (long)gridDim.x * blockDim.x) {
    const int j = (int)(i % W8);
    long r = i / W8;
    const int h = (int)(r % H); r /= H;
    const int c = (int)(r % C);
    const long n = r / C;
    const float4* src = reinterpret_cast<const float4*>(in + n * in_ns + ((long)c * H + h) * W + 8 * j);
    const float4 a = src[0], b = src[1];
    const int pr = h & 1;
    float* o0 = out + n * out_ns + (((long)(pr * 2 + 0) * C + c) * Ho + (h >> 1)) * Wo + 4 * j;
    float* o1 = out + n * out_ns + (((long)(pr * 2 + 1) * C + c) * Ho + (h >> 1)) * Wo + 4 * j;
    *reinterpret_cast<float4*>(o0) = make_float4(a.x, a.z, b.x, b.z);
    *reinterpret_cast<float4*>(o1) = make_float4(a.y, a.w, b.y, b.w);
  }
}

// out[n][c][H][W] (+)= S[n][p*C + c][H/2][W/2]
__global__ void d2s_planar_kernel(const float* __restrict__ S, long s_ns, int C, int H, int W,
                                  float* __restrict__ out, long out_ns, long total8, int accumulate) {
  const int W8 = W / 8, Ho = H / 2, Wo = W / 2;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total8; i += (long)gridDim.x * blockDim.x) {
    const int j = (int)(i % W8);
    long r = i / W8;
    const int h = (int)(r % H); r /= H;
    const int c = (int)(r % C);
    const long n = r / C;
    const int pr = h & 1;
    const float4 e = *reinterpret_cast<const float4*>(
        S + n * s_ns + (((long)(pr * 2 + 0) * C + c) * Ho + (h >> 1)) * Wo + 4 * j);
    const float4 o = *reinterpret_cast<const float4*>(
        S + n * s_ns + (((long)(pr * 2 + 1) * C + c) * Ho + (h >> 1)) * Wo + 4 * j);
    float4* dst = reinterpret_cast<float4*>(out + n * out_ns + ((long)c * H + h) * W + 8 * j);
    float4 a = make_float4(e.x, o.x, e.y, o.y), b = make_float4(e.z, o.z, e.w, o.w);
    if (accumulate) {
      const float4 pa = dst[0], pb = dst[1];
      a.x += pa.x; a.y += pa.y; a.z += pa.z; a.w += pa.w;
      b.x += pb.x; b.y += pb.y; b.z += pb.z; b.w += pb.w;
    }
    dst[0] = a;
    dst[1] = b;
  }
}

// mode 0: W3[co][p*Cin + ci][a][b] = w[co][ci][r][s] (zero where no tap);  mode 1: dW[co][ci][r][s] += dW3[...]
__global__ void s2_weight_map_kernel(const float* __restrict__ src, float* __restrict__ dst, int Cout,
                                     int Cin, int mode) {
  const int n = Cout * 4 * Cin * 9;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int t = i % 9, pc_ci = (i / 9) % (4 * Cin), co = i / (9 * 4 * Cin);
    const int p = pc_ci / Cin, ci = pc_ci % Cin;
    const int r = s2_tap(p >> 1, t / 3), s3 = s2_tap(p & 1, t % 3);
    const bool ok = r >= 0 && s3 >= 0;
    const long o = ((long)co * Cin + ci) * 9 + r * 3 + s3;
    if (mode == 0) dst[i] = ok ? src[o] : 0.f;
    else if (ok) dst[o] += src[i];                // every OIHW element has exactly one source
  }
}

// Stride-2 variant: moving one output pixel to the right shifts the input window by exactly
// one float2, so the window and the accumulators are kept as column PAIRS and every update
// is a packed FFMA2 (KS padded to the next even width; the pad column is discarded).
template <int KS, int COC, int CIC>
__global__ void __launch_bounds__(COC * CIC)
conv_wgrad_s2_kernel(const float* __restrict__ in, long in_ns, int Cin, int H, int W,
                     const float* __restrict__ dY, long dy_ns, int Cout, int Ho, int Wo,
                     float* __restrict__ dW, float* __restrict__ dbias, int N, int ci_groups) {
  constexpr int PAD = KS / 2;
  constexpr int TO = 8;                          // output tile side
  constexpr int IT = (TO - 1) * 2 + KS;          // input tile side
  constexpr int ITP = IT + (IT & 1);             // even row pitch -> 8-byte aligned pairs
  constexpr int ICH = ITP * IT;                  // = 2 * odd for KS = 3, 7 -> conflict-free LDS.64
  constexpr int DCH = (TO * TO) | 1;
  constexpr int NP = (KS + 1) / 2;               // column pairs per window row
  constexpr int NT = COC * CIC;
  extern __shared__ __align__(16) float smem[];
  float* in_s = smem;                            // [CIC][ICH]
  float* dy_s = smem + CIC * ICH;                // [COC][DCH]
  const int tid = threadIdx.x;
  const int lco = tid / CIC, lci = tid % CIC;
  const int co0 = (blockIdx.y / ci_groups) * COC, ci0 = (blockIdx.y % ci_groups) * CIC;
  const int co = co0 + lco, ci = ci0 + lci;
  const int tiles_x = (Wo + TO - 1) / TO, tiles_y = (Ho + TO - 1) / TO;
  const long items = (long)N * tiles_x * tiles_y;
  float2 acc[KS][NP];
#pragma unroll
  for (int r = 0; r < KS; ++r)
#pragma unroll
    for (int j = 0; j < NP; ++j) acc[r][j] = make_float2(0.f, 0.f);
  float bsum = 0.f;

  for (long item = blockIdx.x; item < items; item += gridDim.x) {
    const int n = (int)(item / (tiles_x * tiles_y));
    const int t = (int)(item % (tiles_x * tiles_y));
    const int oh0 = (t / tiles_x) * TO, ow0 = (t % tiles_x) * TO;
    const int ih0 = oh0 * 2 - PAD, iw0 = ow0 * 2 - PAD;
    const float* inp = in + (long)n * in_ns;
    const float* dyp = dY + (long)n * dy_ns;
    for (int i = tid; i < CIC * IT * ITP; i += NT) {
      const int c = i / (IT * ITP), r = (i / ITP) % IT, s3 = i % ITP;
      const int ih = ih0 + r, iw = iw0 + s3;
      float v = 0.f;
      if (s3 < IT && ci0 + c < Cin && ih >= 0 && ih < H && iw >= 0 && iw < W)
        v = inp[((long)(ci0 + c) * H + ih) * W + iw];
      in_s[c * ICH + r * ITP + s3] = v;
    }
    for (int i = tid; i < COC * TO * TO; i += NT) {
      const int c = i / (TO * TO), r = (i / TO) % TO, s3 = i % TO;
      const int oh = oh0 + r, ow = ow0 + s3;
      float v = 0.f;
      if (co0 + c < Cout && oh < Ho && ow < Wo) v = dyp[((long)(co0 + c) * Ho + oh) * Wo + ow];
      dy_s[c * DCH + r * TO + s3] = v;
    }
    __syncthreads();
    const float* ip = in_s + lci * ICH;
    const float* dp = dy_s + lco * DCH;
#pragma unroll 1
    for (int y = 0; y < TO; ++y) {
      float2 win[KS][NP];
#pragma unroll
      for (int r = 0; r < KS; ++r)
#pragma unroll
        for (int j = 0; j + 1 < NP; ++j)
          win[r][j + 1] = *reinterpret_cast<const float2*>(ip + (y * 2 + r) * ITP + 2 * j);
#pragma unroll
      for (int x = 0; x < TO; ++x) {
#pragma unroll
        for (int r = 0; r < KS; ++r) {
#pragma unroll
          for (int j = 0; j + 1 < NP; ++j) win[r][j] = win[r][j + 1];
          win[r][NP - 1] = *reinterpret_cast<const float2*>(ip + (y * 2 + r) * ITP + 2 * x + 2 * (NP - 1));
        }
        const float d = dp[y * TO + x];
        bsum += d;
        const float2 dd = make_float2(d, d);
#pragma unroll
        for (int r = 0; r < KS; ++r)
#pragma unroll
          for (int j = 0; j < NP; ++j) acc[r][j] = __ffma2_rn(dd, win[r][j], acc[r][j]);
      }
    }
    __syncthreads();
  }
  if (co < Cout && ci < Cin) {
    float* wp = dW + ((long)co * Cin + ci) * (KS * KS);
#pragma unroll
    for (int r = 0; r < KS; ++r)
#pragma unroll
      for (int s3 = 0; s3 < KS; ++s3)
        atomicAdd(wp + r * KS + s3, (s3 & 1) ? acc[r][s3 / 2].y : acc[r][s3 / 2].x);
    if (dbias && ci == 0) atomicAdd(dbias + co, bsum);
  }
}

template <int KS, int COC, int CIC>
static int launch_wgrad_s2(const float* in, long in_ns, int Cin, int H, int W, const float* dY,
                           long dy_ns, int Cout, int Ho, int Wo, float* dW, float* dbias, int N,
                           cudaStream_t st) {
  constexpr int TO = 8, IT = (TO - 1) * 2 + KS, ITP = IT + (IT & 1);
  constexpr int ICH = ITP * IT, DCH = (TO * TO) | 1;
  const int smem = (CIC * ICH + COC * DCH) * (int)sizeof(float);
  auto kern = conv_wgrad_s2_kernel<KS, COC, CIC>;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return dmc_check_launch("conv_wgrad_s2 smem attribute");
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
  return dmc_check_launch("conv_wgrad_s2_kernel");
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
  if (ks == 3 && stride == 1 && W % 4 == 0 && in_ns % 4 == 0 && out_ns % 4 == 0 && add_ns % 4 == 0 &&
      (H * W) % 4 == 0) {
#define DMC_FW(COG, R) \
  return launch_fwd_v3<COG, R, 2, 2>(in, in_ns, Cin, H, W, w, bias, Cout, out, out_ns, slope, mask, add, add_ns, accumulate, N, st)
    if (Cout == 2) DMC_FW(2, 4);
    if (Cout == 6) DMC_FW(6, 2);
    DMC_FW(4, 4);
#undef DMC_FW
  } else if (ks == 3 && stride == 1) {
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
  if (ks == 3 && stride == 1 && W % 4 == 0 && in_ns % 4 == 0 && dy_ns % 4 == 0 && (H * W) % 4 == 0) {
    if (Cout == 2) return launch_wgrad_v3<2>(in, in_ns, Cin, H, W, dY, dy_ns, Cout, dW, dbias, N, st);
    if (Cout <= 4) return launch_wgrad_v3<4>(in, in_ns, Cin, H, W, dY, dy_ns, Cout, dW, dbias, N, st);
    if (Cout == 6) return launch_wgrad_v3<6>(in, in_ns, Cin, H, W, dY, dy_ns, Cout, dW, dbias, N, st);
    return launch_wgrad_v3<8>(in, in_ns, Cin, H, W, dY, dy_ns, Cout, dW, dbias, N, st);
  }
  if (ks == 3 && stride == 1)
    return launch_wgrad<3, 1, 8, 32>(in, in_ns, Cin, H, W, dY, dy_ns, Cout, Ho, Wo, dW, dbias, N, st);
  if (ks == 3)
    return launch_wgrad_s2<3, 8, 32>(in, in_ns, Cin, H, W, dY, dy_ns, Cout, Ho, Wo, dW, dbias, N, st);
  return launch_wgrad_s2<7, 64, 2>(in, in_ns, Cin, H, W, dY, dy_ns, Cout, Ho, Wo, dW, dbias, N, st);
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
  const int chunks = (int)cdiv(HW, 256 * 4 * 4) < 1 ? 1 : (int)cdiv(HW, 256 * 4 * 4);
  bn_apply_planar_kernel<<<dim3(chunks, C, N), 256, 0, ST_(stream)>>>(X, x_ns, scale, shift, HW, relu,
                                                                      out, o_ns);
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
  const int chunks = (int)cdiv(HW, 256 * 4 * 4) < 1 ? 1 : (int)cdiv(HW, 256 * 4 * 4);
  bn_bwd_apply_planar_kernel<<<dim3(chunks, C, N), 256, 0, ST_(stream)>>>(
      dZ, dz_ns, X, x_ns, mean, invstd, gamma, sums2, count, C, HW, dX, dx_ns, dgamma, dbeta);
  return dmc_check_launch("bn_bwd_apply_planar_kernel");
}

extern "C" int dmc_copy_planar(const float* src, long s_ns, float* dst, long d_ns, long count, int N,
                               void* stream) {
  const long total = count * N;
  copy_planar_kernel<<<ew_grid(total), 256, 0, ST_(stream)>>>(src, s_ns, dst, d_ns, count, total);
  return dmc_check_launch("copy_planar_kernel");
}

// wT[ci][co][3][3] (ci < ci_count) = spatially flipped, transposed w[co][ci][3][3]
extern "C" int dmc_weight_flip(const float* w, int Cout, int Cin, int ci_count, float* wT,
                               void* stream) {
  DMC_REQUIRE(ci_count >= 1 && ci_count <= Cin, "weight_flip: ci_count=%d Cin=%d", ci_count, Cin);
  const int n = ci_count * Cout * 9;
  weight_flip_kernel<<<(int)cdiv(n, 256), 256, 0, ST_(stream)>>>(w, Cout, Cin, ci_count, wT);
  return dmc_check_launch("weight_flip_kernel");
}

// 3x3 stride-1 data gradient expressed as a forward convolution of dY with the flipped,
// transposed weight wT (dmc_weight_flip): dX[n][0:Cx] (+)= conv(dY, wT); output channels
// [0, act_c1) are then multiplied by LeakyReLU'(act_src) -- the slice of the dense-concat
// gradient that receives its last contribution here becomes the pre-activation gradient.
extern "C" int dmc_conv3x3_dgrad_fused(const float* dY, long dy_ns, int Cy, int H, int W,
                                       const float* wT, int Cx, float* dX, long dx_ns,
                                       int accumulate, const float* act_src, long act_ns, int act_c1,
                                       float act_slope, int N, void* stream) {
  DMC_REQUIRE(W % 4 == 0 && dy_ns % 4 == 0 && dx_ns % 4 == 0 && act_ns % 4 == 0 && (H * W) % 4 == 0,
              "conv3x3_dgrad_fused: W=%d and strides must be multiples of 4", W);
  cudaStream_t st = ST_(stream);
#define DMC_DG(COG, R)                                                                              \
  return launch_fwd_v3<COG, R, 2, 2>(dY, dy_ns, Cy, H, W, wT, nullptr, Cx, dX, dx_ns, 1.f, nullptr, nullptr, \
                                     0, accumulate, N, st, act_src, act_ns, act_c1, act_slope)
  if (Cx == 2) DMC_DG(2, 4);
  if (Cx == 6) DMC_DG(6, 2);
  DMC_DG(4, 4);
#undef DMC_DG
}

// ------------------------------------------------------------------ dense-block data gradient
// Gradient buffer layout [dOut(c_out) | new_{L-1} | ... | new_0]: the gradient of slice k is ONE
// forward convolution over all channels in front of it (every later layer's pre-activation
// gradient), with a weight tensor assembled from the flipped weights of those layers:
//   Wc_k[ci'][c'][t] = w_j[co_j][(x_k + ci') - in_off_j][8 - t],  c' <-> (layer j, output co_j).
struct DenseSeg {
  int c0, cnt;        // channels [c0, c0+cnt) of the gradient buffer belong to layer j
  int w_off;          // offset of w_j (OIHW) in the parameter bucket
  int cin_j;          // input channels of layer j
  int ci_off;         // x_k - in_off_j: where slice k starts inside layer j's input
};
struct DenseSlice {
  int out_off;        // offset of Wc_k in the output buffer
  int gk, cin_s, nseg;
  DenseSeg seg[6];
};
struct DenseTable {
  int nslices;
  DenseSlice sl[6];
};

__global__ void dense_dgrad_weights_kernel(const float* __restrict__ params, DenseTable tab,
                                           float* __restrict__ out) {
  const DenseSlice& sl = tab.sl[blockIdx.x];
  const int n = sl.gk * sl.cin_s * 9;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int t = i % 9, c = (i / 9) % sl.cin_s, ci = i / (9 * sl.cin_s);
    float v = 0.f;
    for (int g = 0; g < sl.nseg; ++g) {
      const DenseSeg& sg = sl.seg[g];
      if (c >= sg.c0 && c < sg.c0 + sg.cnt)
        v = params[sg.w_off + ((long)(c - sg.c0) * sg.cin_j + sg.ci_off + ci) * 9 + (8 - t)];
    }
    out[sl.out_off + i] = v;
  }
}

// table: int32 array [1 + nslices * (4 + 6*5)] = {nslices, per slice: out_off, gk, cin_s, nseg,
// 6 x (c0, cnt, w_off, cin_j, ci_off)} (host memory).
extern "C" int dmc_dense_dgrad_weights(const float* params, const int* table, float* out,
                                       void* stream) {
  DenseTable tab;
  tab.nslices = table[0];
  DMC_REQUIRE(tab.nslices >= 1 && tab.nslices <= 6, "dense_dgrad_weights: nslices=%d", tab.nslices);
  const int* p = table + 1;
  for (int k = 0; k < tab.nslices; ++k) {
    DenseSlice& sl = tab.sl[k];
    sl.out_off = p[0]; sl.gk = p[1]; sl.cin_s = p[2]; sl.nseg = p[3];
    DMC_REQUIRE(sl.nseg >= 1 && sl.nseg <= 6, "dense_dgrad_weights: nseg=%d", sl.nseg);
    for (int g = 0; g < 6; ++g) {
      const int* q = p + 4 + 5 * g;
      sl.seg[g].c0 = q[0]; sl.seg[g].cnt = q[1]; sl.seg[g].w_off = q[2];
      sl.seg[g].cin_j = q[3]; sl.seg[g].ci_off = q[4];
    }
    p += 4 + 30;
  }
  dense_dgrad_weights_kernel<<<tab.nslices, 256, 0, ST_(stream)>>>(params, tab, out);
  return dmc_check_launch("dense_dgrad_weights_kernel");
}

// Space-to-depth of a planar tensor: S[n][(pr*2+pc)*C + c][H/2][W/2] = in[n][c][2i+pr][2j+pc]; W % 8 == 0.
extern "C" int dmc_s2d_planar(const float* in, long in_ns, int C, int H, int W, float* S, long s_ns,
                              int N, void* stream) {
  DMC_REQUIRE(W % 8 == 0 && H % 2 == 0 && in_ns % 4 == 0 && s_ns % 4 == 0, "s2d_planar: H=%d W=%d", H, W);
  const long total8 = (long)N * C * H * (W / 8);
  s2d_planar_kernel<<<ew_grid(total8), 256, 0, ST_(stream)>>>(in, in_ns, C, H, W, S, s_ns, total8);
  return dmc_check_launch("s2d_planar_kernel");
}

// Inverse of dmc_s2d_planar: out[n][c][H][W] (+)= S[n][(pr*2+pc)*C + c][H/2][W/2].
extern "C" int dmc_d2s_planar(const float* S, long s_ns, int C, int H, int W, float* out, long out_ns,
                              int accumulate, int N, void* stream) {
  DMC_REQUIRE(W % 8 == 0 && H % 2 == 0 && out_ns % 4 == 0 && s_ns % 4 == 0, "d2s_planar: H=%d W=%d", H, W);
  const long total8 = (long)N * C * H * (W / 8);
  d2s_planar_kernel<<<ew_grid(total8), 256, 0, ST_(stream)>>>(S, s_ns, C, H, W, out, out_ns, total8,
                                                               accumulate);
  return dmc_check_launch("d2s_planar_kernel");
}

// Weight of the stride-1 convolution on the space-to-depth input that equals a 3x3 stride-2 conv:
// W3[Cout][4*Cin][3][3] from w[Cout][Cin][3][3] (to_s2d = 1), or the gather of its gradient back:
// w_or_dw[Cout][Cin][3][3] += W3 (to_s2d = 0).
extern "C" int dmc_s2_weight_map(const float* src, float* dst, int Cout, int Cin, int to_s2d,
                                 void* stream) {
  const int n = Cout * 4 * Cin * 9;
  s2_weight_map_kernel<<<(int)cdiv(n, 256), 256, 0, ST_(stream)>>>(src, dst, Cout, Cin, to_s2d ? 0 : 1);
  return dmc_check_launch("s2_weight_map_kernel");
}

// 3x3 stride-1 pad-1 convolution restricted to the taps [t0, t0 + 2) of every kernel row and
// column of w (OIHW 3x3; the other taps are taken as zero): t0 = 0 is the forward and t0 = 1 the
// flipped data gradient of a stride-2 conv in space-to-depth form.  Same epilogue as dmc_conv_fwd.
extern "C" int dmc_conv3x3_taps2(const float* in, long in_ns, int Cin, int H, int W, const float* w,
                                 const float* bias, int Cout, int t0, float* out, long out_ns,
                                 float slope, const float* mask, int N, void* stream) {
  DMC_REQUIRE(t0 == 0 || t0 == 1, "conv3x3_taps2: t0=%d", t0);
  DMC_REQUIRE(W % 4 == 0 && in_ns % 4 == 0 && out_ns % 4 == 0 && (H * W) % 4 == 0,
              "conv3x3_taps2: W=%d and strides must be multiples of 4", W);
  cudaStream_t st = ST_(stream);
#define DMC_T2(COG, R, T0)                                                                          \
  return launch_fwd_v3<COG, R, 2, 2, T0, 2>(in, in_ns, Cin, H, W, w, bias, Cout, out, out_ns, slope, \
                                            mask, nullptr, 0, 0, N, st)
  if (t0 == 0) {
    if (Cout == 2) DMC_T2(2, 4, 0);
    DMC_T2(4, 4, 0);
  }
  if (Cout == 2) DMC_T2(2, 4, 1);
  DMC_T2(4, 4, 1);
#undef DMC_T2
}

