// Classifier stem convolution (torchvision resnet18 conv1 with the 2-channel input of
// code/dmcnet/model.py:305-312: 7x7, stride 2, pad 3, 2 -> 64, no bias) on the tensor cores.
//
// A 2-channel input cannot feed TMA-tiled implicit GEMM (4-byte pixels), so the CTA builds
// the im2col operand itself: K = 128 = 8 kernel rows (7 real) x 2 channels x 8 columns
// (7 real), which makes every 16-byte chunk of an operand row simply 8 consecutive input
// pixels of one (kernel row, channel).  Input rows are staged once as bf16 hi/lo planes in a
// ring (loads for the row after next are in flight in registers); 128 builder threads (one
// output pixel each) copy 14 chunks per plane into the SWIZZLE_128B K-major layout tcgen05
// reads while the previous row's MMAs run and the other 128 threads drain its accumulator.  One tile = one output row (<= 128 pixels) x 64 channels.
//
// Precision: both operands are bf16 hi/lo pairs and all four cross products are accumulated
// (fp32 in TMEM): the weight operand is the stack [W_hi ; W_lo] (N = 128), so
//   D[:, 0:64] += A_p . W_hi,  D[:, 64:128] += A_p . W_lo     for p in {hi, lo}
// takes two N=128 MMAs per k-step and the epilogue adds the column halves.
#include "common.cuh"

namespace dmc {

constexpr int ST_RING = 16;                 // staged input rows per (plane, channel)
constexpr int ST_KB = 16 * 1024;            // one 64-wide k-block of a 128-row operand tile
constexpr int ST_PLANE = 2 * ST_KB;         // K = 128
constexpr int ST_A_BYTES = 2 * ST_PLANE;    // hi + lo
constexpr int ST_B_BYTES = ST_PLANE;        // [W_hi ; W_lo] x K

__device__ __forceinline__ uint32_t sw128_off(uint32_t row, uint32_t chunk) {
  return row * 128u + ((chunk ^ (row & 7u)) << 4);
}

// Wb[row][k] bf16: rows 0..63 = hi(w[row]), 64..127 = lo(w[row - 64]); k = r*16 + ci*8 + s.
__global__ void stem_weight_prep_kernel(const float* __restrict__ w, bf16* __restrict__ Wb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 128) return;
  const int co = i / 128, k = i % 128;
  const int r = k / 16, ci = (k / 8) & 1, s = k & 7;
  float v = 0.f;
  if (r < 7 && s < 7) v = w[((co * 2 + ci) * 7 + r) * 7 + s];
  bf16 h, l;
  split_bf16(v, h, l);
  Wb[co * 128 + k] = h;
  Wb[(64 + co) * 128 + k] = l;
}

struct StemSmem {
  uint32_t a[2], b, stg, bar[2], tmem_slot;
};

// Staging of input rows, split in two halves so the global-load latency spans a whole
// pipeline step: stem_rows_load issues the 128-bit loads of `cnt` (<= 7) consecutive image rows
// iy0.. of frame n into registers (zeros outside the image); stem_rows_store splits them to bf16
// hi/lo and writes ring slots slot0..  stg[plane][ci][slot][SW], column index = image column + 3.
// Thread t < 2 * W / 4 owns one 128-bit column group of one channel.
__device__ __forceinline__ void stem_rows_load(const float* __restrict__ in, long in_ns, int H, int W,
                                               int n, int iy0, int cnt, int t, float4 (&v)[7]) {
  const int W4 = W / 4;
  if (t >= 2 * W4) return;
  const int ci = t / W4, j = t - ci * W4;
  const float* src = in + (long)n * in_ns + (long)ci * H * W + 4 * j;
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const int iy = iy0 + r;
    v[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < cnt && iy >= 0 && iy < H) v[r] = *reinterpret_cast<const float4*>(src + (long)iy * W);
  }
}
__device__ __forceinline__ void stem_rows_store(const float4 (&v)[7], int W, int cnt, int slot0,
                                                uint16_t* stg, int SW, int t) {
  const int W4 = W / 4;
  if (t >= 2 * W4) return;
  const int ci = t / W4, j = t - ci * W4;
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    if (r >= cnt) break;
    const int slot = (slot0 + r) & (ST_RING - 1);
    const float f[4] = {v[r].x, v[r].y, v[r].z, v[r].w};
    uint16_t* hi = stg + ((0 * 2 + ci) * ST_RING + slot) * SW + 4 * j + 3;
    uint16_t* lo = stg + ((1 * 2 + ci) * ST_RING + slot) * SW + 4 * j + 3;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      bf16 h, l;
      split_bf16(f[e], h, l);
      hi[e] = __bfloat16_as_ushort(h);
      lo[e] = __bfloat16_as_ushort(l);
    }
  }
}

// Builder thread m copies the 14 real chunks per plane of output pixel x (ring rows sb..sb+6)
// into row m of the im2col tile.
__device__ __forceinline__ void stem_build_row(uint8_t* A, const uint16_t* stg, int SW, int sb, int m,
                                               int x) {
#pragma unroll
  for (int plane = 0; plane < 2; ++plane) {
    uint4 v[14];                      // all loads of a plane first: one shared-memory latency, not 14
#pragma unroll
    for (int c = 0; c < 14; ++c) {
      const int r = c >> 1, ci = c & 1;
      const int slot = (sb + r) & (ST_RING - 1);
      const uint32_t* src =
          reinterpret_cast<const uint32_t*>(stg + ((plane * 2 + ci) * ST_RING + slot) * SW) + x;
      v[c].x = src[0]; v[c].y = src[1]; v[c].z = src[2]; v[c].w = src[3];
    }
#pragma unroll
    for (int c = 0; c < 14; ++c) {    // chunk c = (kernel row c/2, channel c%2) -> k-block c/8, chunk c%8
      *reinterpret_cast<uint4*>(A + plane * ST_PLANE + (c >> 3) * ST_KB + sw128_off(m, c & 7)) = v[c];
    }
  }
}

__global__ void __launch_bounds__(256, 1)
stem_conv_tc_fwd_kernel(const float* __restrict__ in, long in_ns, int H, int W,
                        const bf16* __restrict__ Wb, float* __restrict__ Y, long y_ns, int N) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const int Ho = H / 2, Wo = W / 2, SW = W + 8;
  uint8_t* A0 = gen;                                   // two im2col buffers
  uint8_t* Bs = gen + 2 * ST_A_BYTES;
  uint16_t* stg = reinterpret_cast<uint16_t*>(gen + 2 * ST_A_BYTES + ST_B_BYTES);
  const int stg_bytes = 2 * 2 * ST_RING * SW * 2;
  const uint32_t bar0 = base + 2 * ST_A_BYTES + ST_B_BYTES + ((stg_bytes + 15) & ~15);
  const uint32_t tmem_slot = bar0 + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(gen + (tmem_slot - base));
  const int tid = threadIdx.x, warp = tid >> 5;

  const long total = (long)N * Ho;
  const long rpc = cdiv(total, gridDim.x);
  const long q0 = (long)blockIdx.x * rpc;
  const long q1 = (q0 + rpc < total) ? q0 + rpc : total;
  if (q0 >= q1) return;                                // uniform per CTA

  // ---- one-time setup: zero staging and A (pad chunks stay zero), weights, barriers, TMEM
  for (int i = tid; i < stg_bytes / 4; i += 256) reinterpret_cast<uint32_t*>(stg)[i] = 0u;
  for (int i = tid; i < 2 * ST_A_BYTES / 16; i += 256)
    reinterpret_cast<uint4*>(A0)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < 128 * 16; i += 256) {          // 16-byte chunks of Wb[128][128]
    const int row = i >> 4, c = i & 15;
    const uint4 v = *reinterpret_cast<const uint4*>(Wb + row * 128 + c * 8);
    *reinterpret_cast<uint4*>(Bs + (c >> 3) * ST_KB + sw128_off(row, c & 7)) = v;
  }
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 256);
  __syncthreads();
  // prime: the seven input rows of the first output row
  int sb = 0;                                           // ring slot of kernel row 0 of the current row
  {
    const int n = (int)(q0 / Ho), y = (int)(q0 % Ho);
    float4 v[7];
    stem_rows_load(in, in_ns, H, W, n, 2 * y - 3, 7, tid, v);
    stem_rows_store(v, W, 7, 0, stg, SW, tid);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot_ptr;
  const uint32_t idesc = umma_idesc_bf16(128, 0, 0);

  // Roles: warps 0-3 run the epilogue (thread = output pixel = TMEM lane); warps 4-7 build the
  // im2col tile, stage input rows and issue the MMAs.  The builders never store to global
  // memory, so the generic->async proxy fence after a build does not wait on output stores.
  const int m = tid & 127;
  auto epilogue = [&](int n, int y, uint32_t acc) {
    const long cstride = (long)Ho * Wo;
    float* d = Y + (long)n * y_ns + (long)y * Wo + m;
#pragma unroll 1
    for (int hf = 0; hf < 2; ++hf) {
      uint32_t a[32], b[32];
      tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + acc + hf * 32, a);
      tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + acc + 64 + hf * 32, b);
      tmem_ld_wait();
      if (m < Wo) {
#pragma unroll
        for (int c = 0; c < 32; ++c, d += cstride) *d = __uint_as_float(a[c]) + __uint_as_float(b[c]);
      }
    }
  };

  const int rows = (int)(q1 - q0);
  // loader pipeline: rows of output row it+1 are loaded (registers) during step it-1 and stored
  // to the ring during step it
  float4 pv[7];
  int p_cnt = 0, p_slot = 0;
  // (frame, row) of output rows q-1 (epilogue), q, q+1, q+2, advanced without divisions
  auto advance = [&](int& n, int& y) { if (++y == Ho) { y = 0; ++n; } };
  auto next_rows = [&](int yn, int sbn, int& iy0, int& cnt, int& slot) {   // rows an output row adds
    if (yn == 0) { iy0 = -3; cnt = 7; slot = sbn; }
    else { iy0 = 2 * yn + 2; cnt = 2; slot = sbn + 5; }
  };
  int n0 = (int)(q0 / Ho), y0 = (int)(q0 % Ho);       // current row
  int np = n0, yp = y0;                               // previous row (epilogue)
  int n1 = n0, y1 = y0; advance(n1, y1);
  int n2 = n1, y2 = y1; advance(n2, y2);
  if (warp >= 4 && rows > 1) {
    int iy0;
    next_rows(y1, y1 == 0 ? sb + 7 : sb + 2, iy0, p_cnt, p_slot);
    stem_rows_load(in, in_ns, H, W, n1, iy0, p_cnt, tid - 128, pv);
  }
  for (int it = 0; it < rows; ++it) {
    const int buf = it & 1;
    uint8_t* A = A0 + buf * ST_A_BYTES;
    // ring base of the next row: +2 rows inside a frame, a fresh 7-row window at a frame start
    const int sb_next = y1 == 0 ? sb + 7 : sb + 2;
    if (warp >= 4) {
      if (it >= 2) mbar_wait(bar0 + 8 * buf, ((it - 2) >> 1) & 1);   // MMAs of row it-2 have read A[buf]
      stem_build_row(A, stg, SW, sb, m, m < Wo ? m : Wo - 1);
      fence_proxy_async();
      // loader duty after the fence (it would otherwise wait for the loads issued here)
      const int t = tid - 128;
      if (it + 1 < rows) stem_rows_store(pv, W, p_cnt, p_slot, stg, SW, t);
      if (it + 2 < rows) {
        int iy0;
        next_rows(y2, y2 == 0 ? sb_next + 7 : sb_next + 2, iy0, p_cnt, p_slot);
        stem_rows_load(in, in_ns, H, W, n2, iy0, p_cnt, t, pv);
      }
    }
    tc_fence_before();
    __syncthreads();
    // issued from a builder warp: tcgen05 operations of one warp execute in order, so an
    // epilogue warp that issued the MMAs would have its TMEM loads queue behind them
    if (warp == 4 && elect_one_sync()) {
      tc_fence_after();
      const uint32_t acc = tmem_d + buf * 128;
      const uint32_t a_u32 = base + buf * ST_A_BYTES, b_u32 = base + 2 * ST_A_BYTES;
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t ah = umma_desc_sw128(a_u32 + kb * ST_KB + ks * 32, 16, 1024);
          const uint64_t al = umma_desc_sw128(a_u32 + ST_PLANE + kb * ST_KB + ks * 32, 16, 1024);
          const uint64_t bd = umma_desc_sw128(b_u32 + kb * ST_KB + ks * 32, 16, 1024);
          umma_bf16(acc, ah, bd, idesc, (kb | ks) != 0);
          umma_bf16(acc, al, bd, idesc, 1);
        }
      umma_commit(bar0 + 8 * buf);
    }
    if (warp < 4 && it > 0) {
      mbar_wait(bar0 + 8 * (buf ^ 1), ((it - 1) >> 1) & 1);
      tc_fence_after();
      epilogue(np, yp, (buf ^ 1) * 128);
    }
    sb = sb_next;
    np = n0; yp = y0;
    n0 = n1; y0 = y1;
    n1 = n2; y1 = y2;
    advance(n2, y2);
  }
  if (warp < 4) {
    const int last = rows - 1;
    mbar_wait(bar0 + 8 * (last & 1), (last >> 1) & 1);
    tc_fence_after();
    epilogue(np, yp, (last & 1) * 128);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_d, 256);
}

// ------------------------------------------------------------------ weight gradient
// dW[co][k] = sum over pixels dP[co][px] * im2col[px][k].  Per output row: the im2col tile is
// built exactly as in the forward kernel (here it is the MN-major B operand, N = k), the
// gradient row dP[n][0:64][y][0:Wo] becomes the K-major A operand with the hi and lo planes
// STACKED along M (rows 0..63 = hi, 64..127 = lo, pixels >= Wo zero), and
//   D[128][128] += [dP_hi ; dP_lo] . im2col_hi + [dP_hi ; dP_lo] . im2col_lo
// accumulates in ONE TMEM tile over every row the CTA visits.  The epilogue stores the tile to
// the workspace; stem_wgrad_reduce_kernel adds rows co and 64+co over all CTAs in a fixed order.
constexpr int ST_G_BYTES = 2 * ST_KB;       // dP operand: 128 rows x 128 pixels bf16

__global__ void __launch_bounds__(256, 1)
stem_conv_tc_wgrad_kernel(const float* __restrict__ in, long in_ns, int H, int W,
                          const float* __restrict__ dP, long dp_ns, float* __restrict__ ws, int N) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const int Ho = H / 2, Wo = W / 2, SW = W + 8;
  uint8_t* A0 = gen;                                   // two im2col buffers
  uint8_t* G0 = gen + 2 * ST_A_BYTES;                  // two dP buffers
  uint16_t* stg = reinterpret_cast<uint16_t*>(gen + 2 * ST_A_BYTES + 2 * ST_G_BYTES);
  const int stg_bytes = 2 * 2 * ST_RING * SW * 2;
  const uint32_t bar0 = base + 2 * ST_A_BYTES + 2 * ST_G_BYTES + ((stg_bytes + 15) & ~15);
  const uint32_t tmem_slot = bar0 + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(gen + (tmem_slot - base));
  const int tid = threadIdx.x, warp = tid >> 5;

  const long total = (long)N * Ho;
  const long rpc = cdiv(total, gridDim.x);
  const long q0 = (long)blockIdx.x * rpc;
  const long q1 = (q0 + rpc < total) ? q0 + rpc : total;
  float* wsc = ws + (long)blockIdx.x * 128 * 128;
  if (q0 >= q1) {                                      // uniform per CTA: contributes zeros
    for (int i = tid; i < 128 * 128 / 4; i += 256)
      reinterpret_cast<float4*>(wsc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  for (int i = tid; i < stg_bytes / 4; i += 256) reinterpret_cast<uint32_t*>(stg)[i] = 0u;
  for (int i = tid; i < (2 * ST_A_BYTES + 2 * ST_G_BYTES) / 16; i += 256)
    reinterpret_cast<uint4*>(A0)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 128);
  __syncthreads();
  int sb = 0;
  int n0 = (int)(q0 / Ho), y0 = (int)(q0 % Ho);
  if (warp >= 4) {
    float4 v[7];
    stem_rows_load(in, in_ns, H, W, n0, 2 * y0 - 3, 7, tid - 128, v);
    stem_rows_store(v, W, 7, 0, stg, SW, tid - 128);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot_ptr;
  const uint32_t idesc = umma_idesc_bf16(128, 0, 1);   // A K-major, B MN-major

  const int rows = (int)(q1 - q0);
  const int m = tid & 127;
  auto advance = [&](int& n, int& y) { if (++y == Ho) { y = 0; ++n; } };
  auto next_rows = [&](int yn, int sbn, int& iy0, int& cnt, int& slot) {
    if (yn == 0) { iy0 = -3; cnt = 7; slot = sbn; }
    else { iy0 = 2 * yn + 2; cnt = 2; slot = sbn + 5; }
  };
  int n1 = n0, y1 = y0; advance(n1, y1);
  int n2 = n1, y2 = y1; advance(n2, y2);
  // builders (warps 4-7): input-row pipeline; gradient warps (0-3): thread = (co, 64-pixel half),
  // 16 x 128-bit loads of the NEXT row in flight in registers
  float4 pv[7];
  int p_cnt = 0, p_slot = 0;
  const int gco = m >> 1, ghalf = m & 1;
  float4 gv[16];
  auto grad_load = [&](int n, int y) {
    const float* src = dP + (long)n * dp_ns + ((long)gco * Ho + y) * Wo + ghalf * 64;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const int px = ghalf * 64 + 4 * c;
      gv[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (px < Wo) gv[c] = *reinterpret_cast<const float4*>(src + 4 * c);   // Wo % 4 == 0
    }
  };
  // L2 prefetch of a later gradient row: the register prefetch above keeps only one row
  // (32 KB per SM) in flight, which at DRAM latency caps the kernel near 1.5 TB/s
  auto grad_prefetch = [&](int n, int y) {
    const float* src = dP + (long)n * dp_ns + ((long)gco * Ho + y) * Wo + ghalf * 64;
    if (ghalf * 64 < Wo) asm volatile("prefetch.global.L2 [%0];" ::"l"(src));
    if (ghalf * 64 + 32 < Wo) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + 32));
  };
  auto grad_store = [&](uint8_t* G) {                  // 8 chunks (8 pixels each) of row gco, hi and lo
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
      const float f[8] = {gv[2 * c8].x, gv[2 * c8].y, gv[2 * c8].z, gv[2 * c8].w,
                          gv[2 * c8 + 1].x, gv[2 * c8 + 1].y, gv[2 * c8 + 1].z, gv[2 * c8 + 1].w};
      uint32_t hw[4], lw[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        bf16 h0, l0, h1, l1;
        split_bf16(f[2 * e], h0, l0);
        split_bf16(f[2 * e + 1], h1, l1);
        hw[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        lw[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
      }
      uint8_t* kbp = G + ghalf * ST_KB;                // pixels 0..63 -> k-block 0, 64..127 -> k-block 1
      *reinterpret_cast<uint4*>(kbp + sw128_off(gco, c8)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
      *reinterpret_cast<uint4*>(kbp + sw128_off(64 + gco, c8)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    }
  };
  if (warp >= 4) {
    if (rows > 1) {
      int iy0;
      next_rows(y1, y1 == 0 ? sb + 7 : sb + 2, iy0, p_cnt, p_slot);
      stem_rows_load(in, in_ns, H, W, n1, iy0, p_cnt, tid - 128, pv);
    }
  } else {
    grad_load(n0, y0);
  }
  for (int it = 0; it < rows; ++it) {
    const int buf = it & 1;
    uint8_t* A = A0 + buf * ST_A_BYTES;
    uint8_t* G = G0 + buf * ST_G_BYTES;
    const int sb_next = y1 == 0 ? sb + 7 : sb + 2;
    if (it >= 2) mbar_wait(bar0 + 8 * buf, ((it - 2) >> 1) & 1);   // MMAs of row it-2 are done with [buf]
    if (warp >= 4) {
      stem_build_row(A, stg, SW, sb, m, m < Wo ? m : Wo - 1);
      fence_proxy_async();
      const int t = tid - 128;
      if (it + 1 < rows) stem_rows_store(pv, W, p_cnt, p_slot, stg, SW, t);
      if (it + 2 < rows) {
        int iy0;
        next_rows(y2, y2 == 0 ? sb_next + 7 : sb_next + 2, iy0, p_cnt, p_slot);
        stem_rows_load(in, in_ns, H, W, n2, iy0, p_cnt, t, pv);
      }
    } else {
      grad_store(G);
      fence_proxy_async();
      if (it + 1 < rows) grad_load(n1, y1);
      if (it + 4 < rows) {                             // row it+4: three steps before its register load
        int n4 = n2, y4 = y2;
        advance(n4, y4);
        advance(n4, y4);
        grad_prefetch(n4, y4);
      }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4 && elect_one_sync()) {
      tc_fence_after();
      const uint32_t a_u32 = base + buf * ST_A_BYTES;
      const uint32_t g_u32 = base + 2 * ST_A_BYTES + buf * ST_G_BYTES;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {                 // 16 pixels per MMA
        const uint64_t gd = umma_desc_sw128(g_u32 + (ks >> 2) * ST_KB + (ks & 3) * 32, 16, 1024);
        const uint64_t bh = umma_desc_sw128(a_u32 + ks * 2048, ST_KB, 1024);
        const uint64_t bl = umma_desc_sw128(a_u32 + ST_PLANE + ks * 2048, ST_KB, 1024);
        umma_bf16(tmem_d, gd, bh, idesc, (it | ks) != 0);
        umma_bf16(tmem_d, gd, bl, idesc, 1);
      }
      umma_commit(bar0 + 8 * buf);
    }
    sb = sb_next;
    n0 = n1; y0 = y1;
    n1 = n2; y1 = y2;
    advance(n2, y2);
  }
  // all MMAs retire in order: the last commit covers the whole accumulation
  {
    const int last = rows - 1;
    mbar_wait(bar0 + 8 * (last & 1), (last >> 1) & 1);
    tc_fence_after();
  }
  if (warp < 4) {
    float* dst = wsc + (long)m * 128;
#pragma unroll 1
    for (int c = 0; c < 128; c += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + c, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(dst + c + j) =
            make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                        __uint_as_float(r[j + 3]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_d, 128);
}

// dW[co][ci][r][s] += sum_cta ws[cta][co][k] + ws[cta][64 + co][k],  k = r*16 + ci*8 + s
__global__ void __launch_bounds__(128)
stem_wgrad_reduce_kernel(const float* __restrict__ ws, int ctas, float* __restrict__ dW) {
  const int co = blockIdx.x, k = threadIdx.x;
  float s = 0.f;
  for (int c = 0; c < ctas; ++c) {
    const float* p = ws + ((long)c * 128 + co) * 128 + k;
    s += p[0] + p[64 * 128];
  }
  const int r = k / 16, ci = (k / 8) & 1, s3 = k & 7;
  if (r < 7 && s3 < 7) dW[((co * 2 + ci) * 7 + r) * 7 + s3] += s;
}

static int stem_wgrad_smem_bytes(int W) {
  return 1024 + 2 * ST_A_BYTES + 2 * ST_G_BYTES + ((2 * 2 * ST_RING * (W + 8) * 2 + 15) & ~15) + 64;
}

static int stem_smem_bytes(int W) {
  return 1024 + 2 * ST_A_BYTES + ST_B_BYTES + ((2 * 2 * ST_RING * (W + 8) * 2 + 15) & ~15) + 64;
}

}  // namespace dmc

using namespace dmc;

// Y[n][64][H/2][W/2] = conv7x7/2 pad 3 (in[n][2][H][W], w[64][2][7][7]); wb = 128*128 bf16 scratch
// (the split, re-ordered weight operand, rebuilt on every call).
extern "C" int dmc_stem_conv_tc_fwd(const float* in, long in_ns, int H, int W, const float* w, void* wb,
                                    float* Y, long y_ns, int N, void* stream) {
  DMC_REQUIRE(H % 2 == 0 && W % 8 == 0 && W / 2 <= 128 && W / 2 >= 4 && in_ns % 4 == 0 && W <= 256,
              "stem_conv_tc_fwd: H=%d W=%d (needs W/2 <= 128, W %% 8 == 0)", H, W);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  stem_weight_prep_kernel<<<(64 * 128 + 255) / 256, 256, 0, st>>>(w, (bf16*)wb);
  int rc = dmc_check_launch("stem_weight_prep_kernel");
  if (rc) return rc;
  const int smem = stem_smem_bytes(W);
  static int attr_bytes = 0;
  if (smem > attr_bytes) {
    if (cudaFuncSetAttribute(stem_conv_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
        cudaSuccess)
      return dmc_check_launch("stem_conv_tc_fwd smem attribute");
    attr_bytes = smem;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long grid = (long)N * (H / 2);
  if (grid > sms) grid = sms;
  stem_conv_tc_fwd_kernel<<<(unsigned)grid, 256, smem, st>>>(in, in_ns, H, W, (const bf16*)wb, Y, y_ns, N);
  return dmc_check_launch("stem_conv_tc_fwd_kernel");
}

// dW[64][2][7][7] += weight gradient of the stem conv; dP = [n][64][H/2][W/2] fp32 (dp_ns per frame);
// ws = workspace of dmc_stem_conv_tc_wgrad_workspace() floats.
extern "C" long dmc_stem_conv_tc_wgrad_workspace(void) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (long)sms * 128 * 128;
}

extern "C" int dmc_stem_conv_tc_wgrad(const float* in, long in_ns, int H, int W, const float* dP,
                                      long dp_ns, float* dW, float* ws, int N, void* stream) {
  DMC_REQUIRE(H % 2 == 0 && W % 8 == 0 && W / 2 <= 128 && W / 2 >= 4 && in_ns % 4 == 0 && W <= 256 &&
                  dp_ns % 4 == 0,
              "stem_conv_tc_wgrad: H=%d W=%d (needs W/2 <= 128, W %% 8 == 0)", H, W);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int smem = stem_wgrad_smem_bytes(W);
  static int attr_bytes = 0;
  if (smem > attr_bytes) {
    if (cudaFuncSetAttribute(stem_conv_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
        cudaSuccess)
      return dmc_check_launch("stem_conv_tc_wgrad smem attribute");
    attr_bytes = smem;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long grid = (long)N * (H / 2);
  if (grid > sms) grid = sms;
  stem_conv_tc_wgrad_kernel<<<(unsigned)grid, 256, smem, st>>>(in, in_ns, H, W, dP, dp_ns, ws, N);
  int rc = dmc_check_launch("stem_conv_tc_wgrad_kernel");
  if (rc) return rc;
  stem_wgrad_reduce_kernel<<<64, 128, 0, st>>>(ws, (int)grid, dW);
  return dmc_check_launch("stem_wgrad_reduce_kernel");
}
