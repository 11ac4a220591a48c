// tcgen05 tensor-core kernels for the classifier / discriminator convolutions.
//
// A 3x3 convolution over the padded pixel-major layout [frames][Hp][Wp][C] is a
// sum of "row-shifted" GEMMs: out[q][co] = sum_tap sum_ci act[q + shift(tap)][ci] *
// w[tap][co][ci], where shift(tap) = dr*Wp + ds is a flat offset.  Border rows
// produce garbage and are masked to zero in the epilogue, which keeps the zero
// ring intact for the next layer.  Stride-2 convolutions read four parity
// "phases" of the input, each stored in the OUTPUT geometry, so they use the
// same kernel with a per-tap phase index (see phase_split in pixelwise.cu).
//
// Precision: operands are bf16 hi/lo pairs; each k-step issues hi*hi, lo*hi and
// hi*lo MMAs into one fp32 TMEM accumulator (~2^-16 relative error).
//
//   tap_gemm_ws_kernel: D[P][N] = sum_t A_ph(t)[q+s_t][K] * B_t[N][K]^T    (fprop, dgrad)
//                      K-major operands, TMA 3-D tiled loads, persistent warp-specialised
//                      CTAs (128 x BN tiles, two TMEM accumulators).
//   wgrad_gemm_kernel: dW_t[M][N] += sum_q G[q][M] * A_ph(t)[q+s_t][N]     (split-K, atomics)
//                      MN-major operands (pixel index is K).
#include "common.cuh"
#include <cudaTypedefs.h>
#include <mutex>
#include <cstdlib>

namespace dmc {

static constexpr int MAX_TAPS = 27;      // 3 x 3 x 3 (I3D)
struct TapTable {
  int ntaps;
  int shift[MAX_TAPS];   // row shift applied to the A operand
  int phase[MAX_TAPS];   // A phase (outer TMA coordinate)
  int bsel[MAX_TAPS];    // which weight slice
};

// "Row window" schedule of a K = 64 tap GEMM: taps whose row shifts differ by a few rows (the three
// columns of one kernel row) read almost the same 128 activation rows, so the producer loads ONE
// window of 136 rows per group and the MMAs of each tap start their A descriptor `rowoff` rows into
// it (a SWIZZLE_128B descriptor may start at any 128-byte row of a 1024-byte aligned tile):
// the L2 -> shared traffic of the A operand, which bounds 64-channel layers, drops ~3x.
static constexpr int RW_GROUPS = 8, RW_GT = 3, RW_ROWS = 136;
static constexpr int EPILOGUE_WARPS_DEFAULT = 8;
struct RwTable {
  int ngroups;
  int gshift[RW_GROUPS];          // window start = m0 + gshift
  int gcount[RW_GROUPS];          // taps in the group (<= RW_GT)
  int rowoff[RW_GROUPS][RW_GT];   // tap's first row inside the window (0 .. RW_ROWS - 128)
  int bsel[RW_GROUPS][RW_GT];
};

// ------------------------------------------------------------------ tensor maps
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static std::once_flag g_encode_once;

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  });
  return g_encode;
}

// bf16 tensor [d2][d1][d0] (d0 contiguous), box [1][box1][64], SWIZZLE_128B.
// pitch (elements, 0 = d0): row pitch of the tensor, so that a column sub-range [c0, c0 + d0) of a wider
// [d1][pitch] matrix is an operand (base = its first element; the inception branches of I3D read and write
// slices of the concatenated maps).
static int make_map_3d(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                       uint32_t box0, uint32_t box1, uint64_t pitch = 0) {
  if (pitch == 0) pitch = d0;
  auto enc = get_encode();
  if (!enc) {
    dmc_set_error("cuTensorMapEncodeTiled entry point not available");
    return DMC_ERR_CUDA;
  }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {pitch * sizeof(bf16), pitch * d1 * sizeof(bf16)};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    dmc_set_error("cuTensorMapEncodeTiled failed (%d) dims=%llu,%llu,%llu box=%u,%u", (int)r,
                  (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, box0,
                  box1);
    return DMC_ERR_CUDA;
  }
  return DMC_OK;
}

}  // namespace dmc

int dmc_make_f32_map(CUtensorMap* m, const void* base, int rank, const unsigned long long* dims,
                     const unsigned long long* strides_elems, const unsigned int* box) {
  auto enc = dmc::get_encode();
  if (!enc) {
    dmc_set_error("cuTensorMapEncodeTiled entry point not available");
    return DMC_ERR_CUDA;
  }
  cuuint64_t d[5], strides[4];
  cuuint32_t b[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    estr[i] = 1;
    if (i < rank - 1) strides[i] = strides_elems[i] * sizeof(float);
  }
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<void*>(base), d, strides, b,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    dmc_set_error("cuTensorMapEncodeTiled(f32, rank %d) failed (%d)", rank, (int)r);
    return DMC_ERR_CUDA;
  }
  return DMC_OK;
}

namespace dmc {

static int sm_count();

// ------------------------------------------------------------------ persistent warp-specialised fprop / dgrad
// One CTA per SM walks the tile list (m fastest, so concurrent CTAs share the
// weight tile through L2).  Warp 0 = TMA producer, warp 1 = MMA issuer (one
// thread each), warps 2..5 = epilogue.  Two fp32 accumulators live in TMEM, so
// the epilogue of tile j (tcgen05.ld -> ring mask -> shared-memory transpose ->
// 128-byte coalesced global stores) overlaps the MMAs of tile j+1.
// Optional fused BatchNorm-backward reduction in the epilogue of a data-gradient GEMM: the tile
// D = dX (gradient w.r.t. the activation a = relu(bn(Y) [+ residual])) is turned into
//   dz = (D + gb) * [act > 0]            (written instead of D)
// and sums[0][c] += sum dz, sums[1][c] += sum dz * (Y - mean) * invstd are accumulated, which is
// everything the BN backward needs besides one more elementwise pass.
struct BwFuse {
  const float* Y;        // raw conv output of the BN being differentiated, [M][N]; null = off
  const bf16* act_hi;    // hi plane of the activation (ReLU mask)
  const float* gb;       // second gradient source (residual branch) or null
  const float* mean;
  const float* invstd;
  float slope;           // LeakyReLU slope of the activation (0 = ReLU): dz = D * (act > 0 ? 1 : slope)
};

// Optional fused activation epilogue of a forward GEMM (discriminator blocks, code/dmcnet_GAN/
// model.py:254-279: Conv(bias) -> LeakyReLU -> Dropout2d): D = mask[frame][n] * lrelu(D + bias[n]),
// and the fused column statistics are then those of the activated, masked value (the BatchNorm that
// follows the dropout).  frame = q / frame_rows.  mask may be null (eval mode / no dropout).
struct ActFuse {
  const float* bias;     // [N]; null = off
  const float* mask;     // [frames][N] or null
  float slope;
  unsigned frame_rows;   // Hp * Wp
  // inference with BatchNorm folded into the weights (bias = folded shift): optional residual added
  // BEFORE the activation (fp32 [M][N], or a bf16 hi/lo activation), and the result written as bf16
  // hi/lo planes instead of fp32 D -- the operand of the next GEMM, with no elementwise pass in between
  const float* res_f32;
  const bf16* res_hi;
  const bf16* res_lo;
  bf16* out_hi;
  bf16* out_lo;
};

template <int BN, int STAGES, bool RW = false, int EPW = 4>
struct TapGemmWsSmem {
  static constexpr int A_BYTES = (RW ? RW_ROWS : 128) * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + (RW ? RW_GT : 1) * 2 * B_BYTES;
  static constexpr int EPI_PITCH = 36;                         // floats; STS.128 conflict-free
  static constexpr int EPI_BYTES = EPW * 32 * EPI_PITCH * 4;   // one staging tile per epilogue warp
  static constexpr int RED_BYTES = 4 * 2 * BN * 4;             // per-warp column sums (BN statistics)
  static constexpr int TOTAL = STAGES * STAGE_BYTES + EPI_BYTES + RED_BYTES + 1024 + 256;
};

// EPW = 4 or 8 epilogue warps.  With 8, two warps share a TMEM lane quarter (warp % 4) and split the 32-column
// chunks of a tile between them: the fused BN-backward epilogue of a 64-wide launch (three global loads per
// element behind a 96-cycle k-step) is otherwise the critical path (ncu: 212 vs 144 us).
template <int BN, int STAGES, bool RW = false, int EPW = 4>
__global__ void __launch_bounds__(64 + 32 * EPW, 1)
tap_gemm_ws_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                   const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl,
                   const __grid_constant__ TapTable taps, float* __restrict__ D, long M, int N, int ldD,
                   int K, int Hp, int Wp, int tiles_m, int tiles_n, double* __restrict__ stats,
                   const BwFuse bw, int a_lo_on, const ActFuse act, int ring,
                   const __grid_constant__ RwTable rw, int stats_ld) {
  using S = TapGemmWsSmem<BN, STAGES, RW, EPW>;
  static_assert(EPW == 4 || EPW == 8, "epilogue warps");
  static_assert((BN / 32) % (EPW / 4) == 0, "chunks must divide between the warps of a lane quarter");
  // Two accumulators of 2*BN columns each: columns [0, BN) collect hi*hi + lo*hi, columns [BN, 2BN) the
  // hi*lo term, because A_hi is multiplied with the STACKED operand [B_hi ; B_lo] (adjacent in the stage) in
  // ONE N = 2*BN MMA: two MMAs per k-step instead of three, and 20 KB instead of 24 KB of shared-memory
  // operand reads per k-step of a 128-wide tile (the MMA unit's limiter here).  The epilogue adds the halves.
  constexpr uint32_t ACC_COLS = 2 * BN;
  constexpr uint32_t TMEM_COLS = (2 * ACC_COLS) < 32 ? 32 : 2 * ACC_COLS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  float* epi = reinterpret_cast<float*>(smem_gen + STAGES * S::STAGE_BYTES);
  float* red = reinterpret_cast<float*>(smem_gen + STAGES * S::STAGE_BYTES + S::EPI_BYTES);
  const uint32_t bar_base = smem_base + STAGES * S::STAGE_BYTES + S::EPI_BYTES + S::RED_BYTES;
  const uint32_t bar_full = bar_base, bar_empty = bar_base + 8 * STAGES;
  const uint32_t bar_tfull = bar_base + 16 * STAGES, bar_tempty = bar_tfull + 16;
  const uint32_t tmem_slot = bar_tempty + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = K / 64;
  const int iters = RW ? rw.ngroups : taps.ntaps * kblocks;     // pipeline stages per tile
  const int total_tiles = tiles_m * tiles_n;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, EPW);
    }
    fence_barrier_init();
    tma_prefetch_desc(&mapAh);
    tma_prefetch_desc(&mapAl);
    tma_prefetch_desc(&mapBh);
    tma_prefetch_desc(&mapBl);
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot_ptr;

  if (warp == 0) {
    if (elect_one_sync()) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const long m0 = (long)(tile % tiles_m) * 128;
        const int n0 = (tile / tiles_m) * BN;
        for (int i = 0; i < iters; ++i, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          const uint32_t full = bar_full + 8 * s;
          const uint32_t st = smem_base + s * S::STAGE_BYTES;
          if constexpr (RW) {
            // one 136-row window of A (hi [+ lo]) and the weight tiles of the group's taps
            const int cnt = rw.gcount[i];
            mbar_expect_tx(full, (uint32_t)((a_lo_on ? 2 : 1) * S::A_BYTES + cnt * 2 * S::B_BYTES));
            const int row = (int)(m0 + rw.gshift[i]);
            tma_load_3d(st, &mapAh, full, 0, row, 0);
            if (a_lo_on) tma_load_3d(st + S::A_BYTES, &mapAl, full, 0, row, 0);
            for (int j = 0; j < cnt; ++j) {
              const uint32_t bt = st + 2 * S::A_BYTES + j * 2 * S::B_BYTES;
              tma_load_3d(bt, &mapBh, full, 0, n0, rw.bsel[i][j]);
              tma_load_3d(bt + S::B_BYTES, &mapBl, full, 0, n0, rw.bsel[i][j]);
            }
            continue;
          }
          const int t = i / kblocks, kb = i - t * kblocks;
          mbar_expect_tx(full, a_lo_on ? S::STAGE_BYTES : S::STAGE_BYTES - S::A_BYTES);
          const int row = (int)(m0 + taps.shift[t]);
          tma_load_3d(st, &mapAh, full, kb * 64, row, taps.phase[t]);
          if (a_lo_on) tma_load_3d(st + S::A_BYTES, &mapAl, full, kb * 64, row, taps.phase[t]);
          tma_load_3d(st + 2 * S::A_BYTES, &mapBh, full, kb * 64, n0, taps.bsel[t]);
          tma_load_3d(st + 2 * S::A_BYTES + S::B_BYTES, &mapBl, full, kb * 64, n0, taps.bsel[t]);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      const uint32_t idesc = umma_idesc_bf16(BN, 0, 0);            // A_lo * B_hi            -> columns [0, BN)
      const uint32_t idesc2 = umma_idesc_bf16(2 * BN, 0, 0);       // A_hi * [B_hi ; B_lo]   -> columns [0, 2BN)
      uint32_t it = 0, j = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++j) {
        const uint32_t a = j & 1, aph = (j >> 1) & 1;
        mbar_wait(bar_tempty + 8 * a, aph ^ 1);          // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem_d + a * ACC_COLS;
        for (int i = 0; i < iters; ++i, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(bar_full + 8 * s, ph);
          tc_fence_after();
          const uint32_t st = smem_base + s * S::STAGE_BYTES;
          if constexpr (RW) {
            const int cnt = rw.gcount[i];
            for (int j = 0; j < cnt; ++j) {
              const uint32_t a0 = st + (uint32_t)rw.rowoff[i][j] * 128u;
              const uint32_t bt = st + 2 * S::A_BYTES + j * 2 * S::B_BYTES;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint64_t ah = umma_desc_sw128(a0 + ks * 32, 16, 1024);
                const uint64_t al = umma_desc_sw128(a0 + S::A_BYTES + ks * 32, 16, 1024);
                const uint64_t bh = umma_desc_sw128(bt + ks * 32, 16, 1024);     // B_lo follows B_hi in the stage
                umma_bf16(acc, ah, bh, idesc2, (uint32_t)((i | j | ks) != 0));
                if (a_lo_on) umma_bf16(acc, al, bh, idesc, 1);
              }
            }
            umma_commit(bar_empty + 8 * s);
            continue;
          }
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ah = umma_desc_sw128(st + ks * 32, 16, 1024);
            const uint64_t al = umma_desc_sw128(st + S::A_BYTES + ks * 32, 16, 1024);
            const uint64_t bh = umma_desc_sw128(st + 2 * S::A_BYTES + ks * 32, 16, 1024);   // [B_hi ; B_lo]
            umma_bf16(acc, ah, bh, idesc2, (uint32_t)((i | ks) != 0));
            if (a_lo_on) umma_bf16(acc, al, bh, idesc, 1);
          }
          umma_commit(bar_empty + 8 * s);
        }
        umma_commit(bar_tfull + 8 * a);
      }
    }
  } else {
    // ---- epilogue warps 2..(2 + EPW - 1): TMEM lane quarter = warp % 4; with EPW = 8 the warps ew and ew + 4
    // share a quarter and take the chunks ci = cj * NH + half
    const int wq = warp & 3;
    const int ew = warp - 2;
    constexpr int NH = EPW / 4;
    const int half = ew >> 2;
    float* stage = epi + ew * 32 * S::EPI_PITCH;
    // fused BatchNorm statistics: lane = one column of each 32-column chunk; partial sums
    // stay in registers while the CTA walks tiles of the same output-channel block
    constexpr int NCH = (BN / 32) / NH;               // chunks of THIS warp
    float cs[NCH], css[NCH];          // forward statistics: lane = column of the chunk
    float bs1[NCH][4], bs2[NCH][4];   // fused BN backward: lane owns 4 columns ((lane & 7) * 4 ..)
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      cs[i] = 0.f; css[i] = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) { bs1[i][k] = 0.f; bs2[i][k] = 0.f; }
    }
    const bool bwd = bw.Y != nullptr;
    int stat_n0 = -1;
    auto flush_stats = [&](int n0f) {
      // combine the four epilogue warps in shared memory, then one double atomic per column
      float* mine = red + wq * 2 * BN;
      if (bwd) {
#pragma unroll
        for (int i = 0; i < NCH; ++i)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float a = bs1[i][k], b = bs2[i][k];
            a += __shfl_xor_sync(0xffffffffu, a, 8);  b += __shfl_xor_sync(0xffffffffu, b, 8);
            a += __shfl_xor_sync(0xffffffffu, a, 16); b += __shfl_xor_sync(0xffffffffu, b, 16);
            if (lane < 8) {
              mine[(i * NH + half) * 32 + lane * 4 + k] = a;
              mine[BN + (i * NH + half) * 32 + lane * 4 + k] = b;
            }
            bs1[i][k] = 0.f; bs2[i][k] = 0.f;
          }
      } else {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          mine[(i * NH + half) * 32 + lane] = cs[i];
          mine[BN + (i * NH + half) * 32 + lane] = css[i];
          cs[i] = 0.f;
          css[i] = 0.f;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EPW) : "memory");
      for (int c = ew * 32 + lane; c < 2 * BN; c += 32 * EPW) {
        const float v = red[c] + red[2 * BN + c] + red[4 * BN + c] + red[6 * BN + c];
        const int col = c < BN ? c : c - BN;
        if (n0f + col < N) atomicAdd(stats + (c < BN ? 0 : stats_ld) + n0f + col, (double)v);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EPW) : "memory");
    };
    uint32_t j = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++j) {
      const long m0 = (long)(tile % tiles_m) * 128;
      const int n0 = (tile / tiles_m) * BN;
      if (stats && stat_n0 != n0) {
        if (stat_n0 >= 0) flush_stats(stat_n0);
        stat_n0 = n0;
      }
      const uint32_t a = j & 1, aph = (j >> 1) & 1;
      mbar_wait(bar_tfull + 8 * a, aph);
      tc_fence_after();
      const long q = m0 + wq * 32 + lane;
      const bool keep = q < M && (Hp == 0 || interior_r(q, Hp, Wp, ring));
#pragma unroll
      for (int ci = 0; ci < NCH; ++ci) {
        const int c = (ci * NH + half) * 32;
        uint32_t r[32], r2[32];
        tmem_ld32(tmem_d + a * ACC_COLS + ((uint32_t)(wq * 32) << 16) + c, r);
        tmem_ld32(tmem_d + a * ACC_COLS + ((uint32_t)(wq * 32) << 16) + BN + c, r2);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(__uint_as_float(r[k]) + __uint_as_float(r2[k]));
        if (ci == NCH - 1) {                      // accumulator fully read: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty + 8 * a);
        }
        if (act.bias != nullptr) {
          // bias + LeakyReLU + per-(frame, channel) dropout scale on this thread's row
          const float* mrow = (act.mask && keep)
                                  ? act.mask + (long)((unsigned)q / act.frame_rows) * N + n0 + c : nullptr;
          const long roff = q * (long)ldD + n0 + c;        // this thread's row in a [M][ldD] residual
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(act.bias + n0 + c + 4 * g));
            float4 m4 = make_float4(1.f, 1.f, 1.f, 1.f);
            if (mrow) m4 = __ldg(reinterpret_cast<const float4*>(mrow + 4 * g));
            float4 v;
            v.x = __uint_as_float(r[4 * g + 0]) + b4.x;
            v.y = __uint_as_float(r[4 * g + 1]) + b4.y;
            v.z = __uint_as_float(r[4 * g + 2]) + b4.z;
            v.w = __uint_as_float(r[4 * g + 3]) + b4.w;
            if (keep && act.res_f32) {
              const float4 rr = __ldg(reinterpret_cast<const float4*>(act.res_f32 + roff + 4 * g));
              v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
            } else if (keep && act.res_hi) {
              const uint2 h = __ldg(reinterpret_cast<const uint2*>(act.res_hi + roff + 4 * g));
              const uint2 l = __ldg(reinterpret_cast<const uint2*>(act.res_lo + roff + 4 * g));
              v.x += __uint_as_float(h.x << 16) + __uint_as_float(l.x << 16);
              v.y += __uint_as_float(h.x & 0xffff0000u) + __uint_as_float(l.x & 0xffff0000u);
              v.z += __uint_as_float(h.y << 16) + __uint_as_float(l.y << 16);
              v.w += __uint_as_float(h.y & 0xffff0000u) + __uint_as_float(l.y & 0xffff0000u);
            }
            v.x = (v.x > 0.f ? v.x : v.x * act.slope) * m4.x;
            v.y = (v.y > 0.f ? v.y : v.y * act.slope) * m4.y;
            v.z = (v.z > 0.f ? v.z : v.z * act.slope) * m4.z;
            v.w = (v.w > 0.f ? v.w : v.w * act.slope) * m4.w;
            if (!keep) v = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(stage + lane * S::EPI_PITCH + 4 * g) = v;
          }
        } else {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float4 v;
            v.x = keep ? __uint_as_float(r[4 * g + 0]) : 0.f;
            v.y = keep ? __uint_as_float(r[4 * g + 1]) : 0.f;
            v.z = keep ? __uint_as_float(r[4 * g + 2]) : 0.f;
            v.w = keep ? __uint_as_float(r[4 * g + 3]) : 0.f;
            *reinterpret_cast<float4*>(stage + lane * S::EPI_PITCH + 4 * g) = v;
          }
        }
        __syncwarp();
        if (stats && !bwd) {
          float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
          for (int row = 0; row < 32; ++row) {
            const float v = stage[row * S::EPI_PITCH + lane];
            s1 += v;
            s2 = fmaf(v, v, s2);
          }
          cs[ci] += s1;
          css[ci] += s2;
        }
        if (n0 + c < N) {
          const int c4 = (lane & 7) * 4;
          float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), is = mu;
          if (bwd) {
            mu = *reinterpret_cast<const float4*>(bw.mean + n0 + c + c4);
            is = *reinterpret_cast<const float4*>(bw.invstd + n0 + c + c4);
          }
          if (!bwd) {
            // two batches of four rows; the loads of a second gradient source (bw.gb) are all issued before
            // the first use -- one dependent load per row made this path latency-bound (20 us per tile)
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {
              float4 g4[4];
#pragma unroll
              for (int ii = 0; ii < 4; ++ii) {
                const int row = (hb * 4 + ii) * 4 + (lane >> 3);
                const long gq = m0 + wq * 32 + row;
                g4[ii] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (bw.gb && gq < M) g4[ii] = __ldg(reinterpret_cast<const float4*>(bw.gb + gq * (long)ldD + n0 + c + c4));
              }
#pragma unroll
              for (int ii = 0; ii < 4; ++ii) {
                const int row = (hb * 4 + ii) * 4 + (lane >> 3);
                const long gq = m0 + wq * 32 + row;
                if (gq < M) {
                  float4 v = *reinterpret_cast<const float4*>(stage + row * S::EPI_PITCH + c4);
                  const long o = gq * (long)ldD + n0 + c + c4;
                  v.x += g4[ii].x; v.y += g4[ii].y; v.z += g4[ii].z; v.w += g4[ii].w;
                  if (act.out_hi) {
                    bf16 hh[4], ll[4];
                    split_bf16(v.x, hh[0], ll[0]); split_bf16(v.y, hh[1], ll[1]);
                    split_bf16(v.z, hh[2], ll[2]); split_bf16(v.w, hh[3], ll[3]);
                    *reinterpret_cast<uint2*>(act.out_hi + o) = *reinterpret_cast<uint2*>(hh);
                    *reinterpret_cast<uint2*>(act.out_lo + o) = *reinterpret_cast<uint2*>(ll);
                  } else {
                    *reinterpret_cast<float4*>(D + o) = v;
                  }
                }
              }
            }
          } else {
            // two batches of four rows: all global loads of a batch are issued before any use,
            // so the epilogue warp keeps 12 independent requests in flight
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {
              float4 yv[4], gv[4];
              uint2 hv[4];
#pragma unroll
              for (int ii = 0; ii < 4; ++ii) {
                const int row = (hb * 4 + ii) * 4 + (lane >> 3);
                const long gq = m0 + wq * 32 + row;
                const long off = gq * (long)ldD + n0 + c + c4;
                yv[ii] = make_float4(0.f, 0.f, 0.f, 0.f);
                gv[ii] = yv[ii];
                hv[ii] = make_uint2(0x3f803f80u, 0x3f803f80u);      // act_hi == null: no ReLU mask (all "> 0")
                if (gq < M) {
                  yv[ii] = __ldg(reinterpret_cast<const float4*>(bw.Y + off));
                  if (bw.act_hi) hv[ii] = __ldg(reinterpret_cast<const uint2*>(bw.act_hi + off));
                  if (bw.gb) gv[ii] = __ldg(reinterpret_cast<const float4*>(bw.gb + off));
                }
              }
#pragma unroll
              for (int ii = 0; ii < 4; ++ii) {
                const int row = (hb * 4 + ii) * 4 + (lane >> 3);
                const long gq = m0 + wq * 32 + row;
                if (gq < M) {
                  float4 v = *reinterpret_cast<const float4*>(stage + row * S::EPI_PITCH + c4);
                  const long off = gq * (long)ldD + n0 + c + c4;
                  v.x += gv[ii].x; v.y += gv[ii].y; v.z += gv[ii].z; v.w += gv[ii].w;
                  const uint2 hraw = hv[ii];
                  // bf16 > 0  <=>  sign bit clear and magnitude non-zero
                  v.x = ((hraw.x & 0x8000u) == 0 && (hraw.x & 0x7fffu) != 0) ? v.x : v.x * bw.slope;
                  v.y = ((hraw.x & 0x80000000u) == 0 && (hraw.x & 0x7fff0000u) != 0) ? v.y : v.y * bw.slope;
                  v.z = ((hraw.y & 0x8000u) == 0 && (hraw.y & 0x7fffu) != 0) ? v.z : v.z * bw.slope;
                  v.w = ((hraw.y & 0x80000000u) == 0 && (hraw.y & 0x7fff0000u) != 0) ? v.w : v.w * bw.slope;
                  const float4 y = yv[ii];
                  bs1[ci][0] += v.x; bs1[ci][1] += v.y; bs1[ci][2] += v.z; bs1[ci][3] += v.w;
                  bs2[ci][0] = fmaf(v.x, (y.x - mu.x) * is.x, bs2[ci][0]);
                  bs2[ci][1] = fmaf(v.y, (y.y - mu.y) * is.y, bs2[ci][1]);
                  bs2[ci][2] = fmaf(v.z, (y.z - mu.z) * is.z, bs2[ci][2]);
                  bs2[ci][3] = fmaf(v.w, (y.w - mu.w) * is.w, bs2[ci][3]);
                  *reinterpret_cast<float4*>(D + off) = v;
                }
              }
            }
          }
        }
        __syncwarp();
      }
    }
    if (stats && stat_n0 >= 0) flush_stats(stat_n0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_d, TMEM_COLS);
}

template <int BN, int STAGES, bool RW = false, int EPW = 4>
static int launch_tap_gemm_ws(const CUtensorMap& mAh, const CUtensorMap& mAl, const CUtensorMap& mBh,
                              const CUtensorMap& mBl, const TapTable& taps, float* D, long M, int N,
                              int ldD, int K, int Hp, int Wp, int sms, double* stats,
                              const BwFuse& bw, int a_lo_on, const ActFuse& act, int ring,
                              cudaStream_t stream, const RwTable* rw = nullptr, int stats_ld = 0) {
  using S = TapGemmWsSmem<BN, STAGES, RW, EPW>;
  auto kern = tap_gemm_ws_kernel<BN, STAGES, RW, EPW>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) !=
        cudaSuccess)
      return dmc_check_launch("tap_gemm_ws smem attribute");
    attr_set = true;
  }
  const int tiles_m = (int)cdiv(M, 128), tiles_n = (int)cdiv(N, BN);
  long grid = (long)tiles_m * tiles_n;
  if (grid > sms) grid = sms;
  RwTable none;
  none.ngroups = 0;
  kern<<<(unsigned)grid, 64 + 32 * EPW, S::TOTAL, stream>>>(mAh, mAl, mBh, mBl, taps, D, M, N, ldD, K, Hp, Wp,
                                                  tiles_m, tiles_n, stats, bw, a_lo_on, act, ring,
                                                  rw ? *rw : none, stats_ld > 0 ? stats_ld : N);
  return dmc_check_launch("tap_gemm_ws_kernel");
}

// Groups the taps of a K = 64, single-phase tap GEMM into row windows (see RwTable); false when the tap
// set does not fit (more than RW_GROUPS groups).
static bool build_row_windows(const TapTable& tt, RwTable& rw) {
  int order[MAX_TAPS];
  for (int i = 0; i < tt.ntaps; ++i) order[i] = i;
  for (int i = 1; i < tt.ntaps; ++i)                      // insertion sort by shift
    for (int j = i; j > 0 && tt.shift[order[j]] < tt.shift[order[j - 1]]; --j) {
      const int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t;
    }
  rw.ngroups = 0;
  for (int g = 0; g < RW_GROUPS; ++g) {
    rw.gshift[g] = 0; rw.gcount[g] = 0;
    for (int j = 0; j < RW_GT; ++j) { rw.rowoff[g][j] = 0; rw.bsel[g][j] = 0; }
  }
  int i = 0;
  while (i < tt.ntaps) {
    if (rw.ngroups == RW_GROUPS) return false;
    const int g = rw.ngroups++;
    const int base = tt.shift[order[i]];
    rw.gshift[g] = base;
    while (i < tt.ntaps && rw.gcount[g] < RW_GT && tt.shift[order[i]] - base <= RW_ROWS - 128) {
      const int j = rw.gcount[g]++;
      rw.rowoff[g][j] = tt.shift[order[i]] - base;
      rw.bsel[g][j] = tt.bsel[order[i]];
      ++i;
    }
  }
  return true;
}

// ------------------------------------------------------------------ wgrad (split-K, tap groups)
// One CTA accumulates TG taps of one (128 x BN) weight tile over its pixel range: the dY
// operand (M side) is loaded ONCE per k-block and multiplied with TG row-shifted activation
// tiles into TG TMEM accumulators, instead of being re-read by nine single-tap CTAs.
template <int BN, int TG, int BKP, int STAGES>
struct WgradSmem {
  static constexpr int BOX_BYTES = BKP * 128;      // BKP pixels x 64 channels bf16
  static constexpr int G_BYTES = 2 * BOX_BYTES;    // M = 128 channels of dY (one plane)
  static constexpr int X_BYTES = (BN / 64) * BOX_BYTES;   // one tap, one plane
  static constexpr int STAGE_BYTES = 2 * G_BYTES + TG * 2 * X_BYTES;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 + 256;
  // STACK: G_hi is multiplied with the stacked operand [X_hi | X_lo] (adjacent boxes of the stage) in one
  // N = 2*BN MMA, as in tap_gemm_ws_kernel; needs 2*BN accumulator columns per tap
  static constexpr bool STACK = TG * 2 * BN <= 512;
  static constexpr int ACC = STACK ? 2 * BN : BN;
  static constexpr uint32_t TMEM_COLS = (TG * ACC) <= 32 ? 32 : (TG * ACC) <= 64 ? 64 : (TG * ACC) <= 128 ? 128
                                        : (TG * ACC) <= 256 ? 256 : 512;
};

template <int BN, int TG, int BKP, int STAGES>
__global__ void __launch_bounds__(128)
wgrad_gemm_kernel(const __grid_constant__ CUtensorMap mapGh, const __grid_constant__ CUtensorMap mapGl,
                  const __grid_constant__ CUtensorMap mapXh, const __grid_constant__ CUtensorMap mapXl,
                  const __grid_constant__ TapTable taps, float* __restrict__ dW, int Cout, int Cin,
                  long P, int kb_per_split, int n_tiles, int oihw_taps, int g_lo_on,
                  float* __restrict__ ws) {
  using S = WgradSmem<BN, TG, BKP, STAGES>;
  static_assert(TG * S::ACC <= 512, "tap group does not fit TMEM");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * S::STAGE_BYTES;
  const uint32_t tmem_full = bar_base + 16 * STAGES;
  const uint32_t tmem_slot = tmem_full + 8;
  uint32_t* tmem_slot_ptr =
      reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = blockIdx.z * TG;
  const int nt_g = (taps.ntaps - t0) < TG ? (taps.ntaps - t0) : TG;     // taps in this group
  const int mt = blockIdx.y / n_tiles, nt = blockIdx.y % n_tiles;
  const int m0 = mt * 128, n0 = nt * BN;
  const long kb_total = cdiv(P, BKP);
  const long kb0 = (long)blockIdx.x * kb_per_split;
  long kb1 = kb0 + kb_per_split;
  if (kb1 > kb_total) kb1 = kb_total;
  const int iters = (int)(kb1 - kb0);
  const int m_boxes = (Cout - m0) >= 128 ? 2 : 1;   // second 64-channel group may not exist
  const uint32_t stage_tx = (uint32_t)((g_lo_on ? 2 : 1) * m_boxes * S::BOX_BYTES + nt_g * 2 * S::X_BYTES);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_base + 8 * s, 1);
      mbar_init(bar_base + 8 * (STAGES + s), 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&mapGh);
    tma_prefetch_desc(&mapGl);
    tma_prefetch_desc(&mapXh);
    tma_prefetch_desc(&mapXl);
  }
  if (warp == 2) tmem_alloc(tmem_slot, S::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot_ptr;
  if (iters <= 0) {   // uniform per CTA
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_d, S::TMEM_COLS);
    return;
  }

  if (warp == 0 && elect_one_sync()) {
    for (int i = 0; i < iters; ++i) {
      const int s = i % STAGES;
      if (i >= STAGES) mbar_wait(bar_base + 8 * (STAGES + s), ((i / STAGES) - 1) & 1);
      const uint32_t full = bar_base + 8 * s;
      const uint32_t st = smem_base + s * S::STAGE_BYTES;
      mbar_expect_tx(full, stage_tx);
      const int row = (int)((kb0 + i) * BKP);
      for (int b = 0; b < m_boxes; ++b) {
        tma_load_3d(st + b * S::BOX_BYTES, &mapGh, full, m0 + 64 * b, row, 0);
        if (g_lo_on) tma_load_3d(st + S::G_BYTES + b * S::BOX_BYTES, &mapGl, full, m0 + 64 * b, row, 0);
      }
      for (int g = 0; g < nt_g; ++g) {
        const uint32_t xs = st + 2 * S::G_BYTES + g * 2 * S::X_BYTES;
        const int xrow = row + taps.shift[t0 + g], ph = taps.phase[t0 + g];
#pragma unroll
        for (int b = 0; b < BN / 64; ++b) {
          tma_load_3d(xs + b * S::BOX_BYTES, &mapXh, full, n0 + 64 * b, xrow, ph);
          tma_load_3d(xs + S::X_BYTES + b * S::BOX_BYTES, &mapXl, full, n0 + 64 * b, xrow, ph);
        }
      }
    }
  } else if (warp == 1 && elect_one_sync()) {
    const uint32_t idesc = umma_idesc_bf16(BN, 1, 1);
    const uint32_t idesc2 = umma_idesc_bf16(2 * BN, 1, 1);
    for (int i = 0; i < iters; ++i) {
      const int s = i % STAGES;
      mbar_wait(bar_base + 8 * s, (i / STAGES) & 1);
      tc_fence_after();
      const uint32_t st = smem_base + s * S::STAGE_BYTES;
      for (int g = 0; g < nt_g; ++g) {
        const uint32_t xs = st + 2 * S::G_BYTES + g * 2 * S::X_BYTES;
        const uint32_t acc = tmem_d + g * S::ACC;
#pragma unroll
        for (int ks = 0; ks < BKP / 16; ++ks) {   // 16 pixels per MMA = two 8-row swizzle atoms
          const uint32_t koff = ks * 2048;
          const uint64_t gh = umma_desc_sw128(st + koff, S::BOX_BYTES, 1024);
          const uint64_t gl = umma_desc_sw128(st + S::G_BYTES + koff, S::BOX_BYTES, 1024);
          const uint64_t xh = umma_desc_sw128(xs + koff, S::BOX_BYTES, 1024);
          if constexpr (S::STACK) {
            // G_hi * [X_hi | X_lo] (the lo boxes follow the hi boxes at the same box pitch), then G_lo * X_hi
            umma_bf16(acc, gh, xh, idesc2, (uint32_t)((i | ks) != 0));
            if (g_lo_on) umma_bf16(acc, gl, xh, idesc, 1);
          } else {
            const uint64_t xl = umma_desc_sw128(xs + S::X_BYTES + koff, S::BOX_BYTES, 1024);
            if (g_lo_on) umma_bf16(acc, gl, xh, idesc, (i | ks) != 0);
            umma_bf16(acc, gh, xl, idesc, g_lo_on ? 1u : (uint32_t)((i | ks) != 0));
            umma_bf16(acc, gh, xh, idesc, 1);
          }
        }
      }
      umma_commit(bar_base + 8 * (STAGES + s));
    }
    umma_commit(tmem_full);
  }
  __syncwarp();

  mbar_wait(tmem_full, 0);
  tc_fence_after();
  const int co = m0 + warp * 32 + lane;
  const int wstep = oihw_taps ? oihw_taps : 1;
  for (int g = 0; g < nt_g; ++g) {
    if (ws) {
      // split-K partial tile -> workspace [split][tap][Cout][Cin] with plain 128-bit stores (each
      // lane owns 128 contiguous bytes of its row); wgrad_reduce_kernel sums the splits in a
      // fixed order.  No atomics: deterministic, and ~4x cheaper than scattered 4-byte REDs.
      float* wrow = ws + (((long)blockIdx.x * taps.ntaps + (t0 + g)) * Cout + co) * Cin + n0;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + g * S::ACC + c, r);
        if constexpr (S::STACK) {
          uint32_t r2[32];
          tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + g * S::ACC + BN + c, r2);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(__uint_as_float(r[k]) + __uint_as_float(r2[k]));
        }
        tmem_ld_wait();
        if (co < Cout) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (n0 + c + j < Cin)
              *reinterpret_cast<float4*>(wrow + c + j) =
                  make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                              __uint_as_float(r[j + 3]));
        }
      }
      continue;
    }
    // output layout: [slice][Cout][Cin] (oihw_taps == 0) or the OIHW gradient itself,
    // dW[co][ci][tap] with oihw_taps taps per filter (no repacking pass afterwards)
    const int bsel = taps.bsel[t0 + g];
    const long wbase = oihw_taps ? ((long)co * Cin + n0) * oihw_taps + bsel
                                 : ((long)bsel * Cout + co) * Cin + n0;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + g * S::ACC + c, r);
      if constexpr (S::STACK) {
        uint32_t r2[32];
        tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + g * S::ACC + BN + c, r2);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(__uint_as_float(r[k]) + __uint_as_float(r2[k]));
      }
      tmem_ld_wait();
      if (co < Cout) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + c + j < Cin) atomicAdd(dW + wbase + (long)(c + j) * wstep, __uint_as_float(r[j]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_d, S::TMEM_COLS);
}

// ------------------------------------------------------------------ wgrad, 64 x 64 layers
// With Cout = 64 the generic kernel fills only half of the M = 128 MMA rows.  Here the roles
// are swapped and two taps share one MMA: A (M side) = [X(tap 2p) ; X(tap 2p+1)] (2 x 64 input
// channels, each box loaded at its own row shift), B (N side) = dY (64 output channels,
// loaded once per k-block), so D_p[128][64] holds dW^T of both taps.  One CTA accumulates all
// taps (ceil(ntaps/2) TMEM accumulators of 64 columns) over its pixel range.
template <int BKP, int STAGES>
struct Wgrad64Smem {
  static constexpr int BOX_BYTES = BKP * 128;                  // BKP pixels x 64 channels bf16
  static constexpr int G_BYTES = BOX_BYTES;                    // one plane of dY
  static constexpr int XP_BYTES = 2 * BOX_BYTES;               // one plane of a tap pair
  static constexpr int MAX_PAIRS = 5;
  static constexpr int STAGE_BYTES = 2 * G_BYTES + MAX_PAIRS * 2 * XP_BYTES;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int BKP, int STAGES>
__global__ void __launch_bounds__(128)
wgrad64_kernel(const __grid_constant__ CUtensorMap mapGh, const __grid_constant__ CUtensorMap mapGl,
               const __grid_constant__ CUtensorMap mapXh, const __grid_constant__ CUtensorMap mapXl,
               const __grid_constant__ TapTable taps, float* __restrict__ dW, long P, int kb_per_split,
               int oihw_taps, int g_lo_on, float* __restrict__ ws) {
  using S = Wgrad64Smem<BKP, STAGES>;
  constexpr int Cout = 64, Cin = 64;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * S::STAGE_BYTES;
  const uint32_t tmem_full = bar_base + 16 * STAGES;
  const uint32_t tmem_slot = tmem_full + 8;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntaps = taps.ntaps, npairs = (ntaps + 1) / 2;
  const long kb_total = cdiv(P, BKP);
  const long kb0 = (long)blockIdx.x * kb_per_split;
  long kb1 = kb0 + kb_per_split;
  if (kb1 > kb_total) kb1 = kb_total;
  const int iters = (int)(kb1 - kb0);
  const uint32_t stage_tx = (uint32_t)((g_lo_on ? 2 : 1) * S::G_BYTES + ntaps * 2 * S::BOX_BYTES);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_base + 8 * s, 1);
      mbar_init(bar_base + 8 * (STAGES + s), 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&mapGh);
    tma_prefetch_desc(&mapGl);
    tma_prefetch_desc(&mapXh);
    tma_prefetch_desc(&mapXl);
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot_ptr;
  if (iters <= 0) {   // uniform per CTA (never happens: every split owns >= 1 k-block)
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_d, 512);
    return;
  }

  if (warp == 0 && elect_one_sync()) {
    for (int i = 0; i < iters; ++i) {
      const int s = i % STAGES;
      if (i >= STAGES) mbar_wait(bar_base + 8 * (STAGES + s), ((i / STAGES) - 1) & 1);
      const uint32_t full = bar_base + 8 * s;
      const uint32_t st = smem_base + s * S::STAGE_BYTES;
      mbar_expect_tx(full, stage_tx);
      const int row = (int)((kb0 + i) * BKP);
      tma_load_3d(st, &mapGh, full, 0, row, 0);
      if (g_lo_on) tma_load_3d(st + S::G_BYTES, &mapGl, full, 0, row, 0);
      for (int t = 0; t < ntaps; ++t) {
        // pair p = t/2, box b = t%2 inside the pair's M = 128 tile; planes hi | lo per pair
        const uint32_t xs = st + 2 * S::G_BYTES + (t >> 1) * 2 * S::XP_BYTES + (t & 1) * S::BOX_BYTES;
        const int xrow = row + taps.shift[t], ph = taps.phase[t];
        tma_load_3d(xs, &mapXh, full, 0, xrow, ph);
        tma_load_3d(xs + S::XP_BYTES, &mapXl, full, 0, xrow, ph);
      }
    }
  } else if (warp == 1 && elect_one_sync()) {
    const uint32_t idesc = umma_idesc_bf16(64, 1, 1);
    for (int i = 0; i < iters; ++i) {
      const int s = i % STAGES;
      mbar_wait(bar_base + 8 * s, (i / STAGES) & 1);
      tc_fence_after();
      const uint32_t st = smem_base + s * S::STAGE_BYTES;
      for (int p = 0; p < npairs; ++p) {
        const uint32_t xs = st + 2 * S::G_BYTES + p * 2 * S::XP_BYTES;
        const uint32_t acc = tmem_d + p * 64;
#pragma unroll
        for (int ks = 0; ks < BKP / 16; ++ks) {   // 16 pixels per MMA = two 8-row swizzle atoms
          const uint32_t koff = ks * 2048;
          const uint64_t xh = umma_desc_sw128(xs + koff, S::BOX_BYTES, 1024);
          const uint64_t xl = umma_desc_sw128(xs + S::XP_BYTES + koff, S::BOX_BYTES, 1024);
          const uint64_t gh = umma_desc_sw128(st + koff, S::BOX_BYTES, 1024);
          const uint64_t gl = umma_desc_sw128(st + S::G_BYTES + koff, S::BOX_BYTES, 1024);
          umma_bf16(acc, xl, gh, idesc, (i | ks) != 0);
          if (g_lo_on) umma_bf16(acc, xh, gl, idesc, 1);
          umma_bf16(acc, xh, gh, idesc, 1);
        }
      }
      umma_commit(bar_base + 8 * (STAGES + s));
    }
    umma_commit(tmem_full);
  }
  __syncwarp();

  mbar_wait(tmem_full, 0);
  tc_fence_after();
  // thread = row m of every pair tile: tap 2p + (m >= 64), input channel m & 63; columns = co
  const int m = warp * 32 + lane;
  const int ci = m & 63;
  for (int p = 0; p < npairs; ++p) {
    const int t = 2 * p + (m >> 6);
    const bool ok = t < ntaps;
    const int bsel = ok ? taps.bsel[t] : 0;
#pragma unroll 1
    for (int c = 0; c < 64; c += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + p * 64 + c, r);
      tmem_ld_wait();
      if (!ok) continue;
      if (ws) {       // [split][tap][Cout][Cin]: lanes = consecutive ci -> coalesced
        float* o = ws + (((long)blockIdx.x * ntaps + t) * Cout + c) * Cin + ci;
#pragma unroll
        for (int j = 0; j < 32; ++j) o[(long)j * Cin] = __uint_as_float(r[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int co = c + j;
          float* o = oihw_taps ? dW + ((long)co * Cin + ci) * oihw_taps + bsel
                               : dW + ((long)bsel * Cout + co) * Cin + ci;
          atomicAdd(o, __uint_as_float(r[j]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_d, 512);
}

// dW (+)= sum over splits of the workspace partials.  Block = 32 consecutive (co, ci) filters of ONE
// tap x SL split lanes (one warp each); a thread sums its splits, the lanes are combined through
// shared memory in a fixed order, and lane-warp 0 adds the tap into dW.
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ ws, int splits, const __grid_constant__ TapTable taps,
                    int Cout, int Cin, int oihw_taps, float* __restrict__ dW) {
  __shared__ float part[8][32];                   // [SL][lane]
  const int lane = threadIdx.x & 31, sl = threadIdx.x >> 5, SL = blockDim.x >> 5;
  const long filters = (long)Cout * Cin;
  const long f = (long)blockIdx.x * 32 + lane;
  const int ntaps = taps.ntaps, t = blockIdx.y;   // one tap per block row: 9 (27) x more CTAs in flight
  float acc = 0.f;
  if (f < filters) {
    const float* p = ws + (long)t * filters + f;
    const long sp_stride = (long)ntaps * filters;
#pragma unroll 4
    for (int sp = sl; sp < splits; sp += SL) acc += p[sp * sp_stride];
  }
  part[sl][lane] = acc;
  __syncthreads();
  if (sl == 0 && f < filters) {
    for (int k = 1; k < SL; ++k) acc += part[k][lane];
    const int co = (int)(f / Cin), ci = (int)(f % Cin);
    const int bsel = taps.bsel[t];
    float* o = oihw_taps ? dW + f * oihw_taps + bsel : dW + ((long)bsel * Cout + co) * Cin + ci;
    *o += acc;
  }
}

template <int TG>
static void wgrad_plan(int BN, int BKP, int Cout, int Cin, int ntaps, long P, int sm_count,
                       long* kb_per_split, long* splits) {
  const int m_tiles = (int)cdiv(Cout, 128), n_tiles = (int)cdiv(Cin, BN);
  const int groups = (int)cdiv(ntaps, TG);
  const long kb_total = cdiv(P, BKP);
  const long tiles = (long)m_tiles * n_tiles * groups;
  // whole waves: these CTAs are resident one or two per SM, so the CTA count must not spill
  // a few CTAs into an extra round
  long want_splits = ((long)sm_count * (TG > 1 ? 2 : 4)) / tiles;
  if (want_splits < 1) want_splits = 1;
  long kbs = cdiv(kb_total, want_splits);
  if (kbs < 16) kbs = 16;     // shorter pixel ranges are all prologue + epilogue (1x1 downsample layers)
  *kb_per_split = kbs;
  *splits = cdiv(kb_total, kbs);
}

template <int BN, int TG, int BKP, int STAGES>
static int launch_wgrad(const void* G_hi, const void* G_lo, const void* X_hi, const void* X_lo,
                        int x_phases, const TapTable& taps, float* dW, int Cout, int Cin, long P,
                        int sm_count, int oihw_taps, float* ws, long ws_floats, cudaStream_t stream,
                        int ldg = 0, int ldx = 0) {
  using S = WgradSmem<BN, TG, BKP, STAGES>;
  CUtensorMap mGh, mGl, mXh, mXl;
  int rc;
  if ((rc = make_map_3d(&mGh, G_hi, Cout, P, 1, 64, BKP, ldg))) return rc;
  const int g_lo_on = G_lo != nullptr;     // G_lo == NULL: dY is used at bf16 precision
  if ((rc = make_map_3d(&mGl, g_lo_on ? G_lo : G_hi, Cout, P, 1, 64, BKP, ldg))) return rc;
  if ((rc = make_map_3d(&mXh, X_hi, Cin, P, x_phases, 64, BKP, ldx))) return rc;
  if ((rc = make_map_3d(&mXl, X_lo, Cin, P, x_phases, 64, BKP, ldx))) return rc;
  auto kern = wgrad_gemm_kernel<BN, TG, BKP, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) !=
        cudaSuccess)
      return dmc_check_launch("wgrad smem attribute");
    attr_set = true;
  }
  const int m_tiles = (int)cdiv(Cout, 128), n_tiles = (int)cdiv(Cin, BN);
  const int groups = (int)cdiv(taps.ntaps, TG);
  long kb_per_split, splits;
  wgrad_plan<TG>(BN, BKP, Cout, Cin, taps.ntaps, P, sm_count, &kb_per_split, &splits);
  const long per_split = (long)taps.ntaps * Cout * Cin;
  if (ws) DMC_REQUIRE(ws_floats >= splits * per_split, "wgrad: workspace %ld floats < %ld needed", ws_floats,
                      splits * per_split);
  dim3 grid((unsigned)splits, (unsigned)(m_tiles * n_tiles), (unsigned)groups);
  kern<<<grid, 128, S::TOTAL, stream>>>(mGh, mGl, mXh, mXl, taps, dW, Cout, Cin, P,
                                        (int)kb_per_split, n_tiles, oihw_taps, g_lo_on, ws);
  if ((rc = dmc_check_launch("wgrad_gemm_kernel"))) return rc;
  if (ws) {
    const int SL = splits >= 8 ? 8 : (splits >= 4 ? 4 : (splits >= 2 ? 2 : 1));
    wgrad_reduce_kernel<<<dim3((unsigned)cdiv((long)Cout * Cin, 32), (unsigned)taps.ntaps), 32 * SL, 0, stream>>>(
        ws, (int)splits, taps, Cout, Cin, oihw_taps, dW);
    return dmc_check_launch("wgrad_reduce_kernel");
  }
  return DMC_OK;
}

static void wgrad64_plan(int BKP, long P, int sm_count, long* kb_per_split, long* splits) {
  const long kb_total = cdiv(P, BKP);
  long kbs = cdiv(kb_total, (long)sm_count * 2);      // one CTA per SM resident, two rounds
  if (kbs < 16) kbs = 16;     // shorter pixel ranges are all prologue + epilogue (1x1 downsample layers)
  *kb_per_split = kbs;
  *splits = cdiv(kb_total, kbs);
}

template <int BKP, int STAGES>
static int launch_wgrad64(const void* G_hi, const void* G_lo, const void* X_hi, const void* X_lo,
                          int x_phases, const TapTable& taps, float* dW, long P, int sm_count,
                          int oihw_taps, float* ws, long ws_floats, cudaStream_t stream, int ldg = 0,
                          int ldx = 0) {
  using S = Wgrad64Smem<BKP, STAGES>;
  CUtensorMap mGh, mGl, mXh, mXl;
  int rc;
  const int g_lo_on = G_lo != nullptr;
  if ((rc = make_map_3d(&mGh, G_hi, 64, P, 1, 64, BKP, ldg))) return rc;
  if ((rc = make_map_3d(&mGl, g_lo_on ? G_lo : G_hi, 64, P, 1, 64, BKP, ldg))) return rc;
  if ((rc = make_map_3d(&mXh, X_hi, 64, P, x_phases, 64, BKP, ldx))) return rc;
  if ((rc = make_map_3d(&mXl, X_lo, 64, P, x_phases, 64, BKP, ldx))) return rc;
  auto kern = wgrad64_kernel<BKP, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) != cudaSuccess)
      return dmc_check_launch("wgrad64 smem attribute");
    attr_set = true;
  }
  long kb_per_split, splits;
  wgrad64_plan(BKP, P, sm_count, &kb_per_split, &splits);
  const long per_split = (long)taps.ntaps * 64 * 64;
  if (ws) DMC_REQUIRE(ws_floats >= splits * per_split, "wgrad: workspace %ld floats < %ld needed", ws_floats,
                      splits * per_split);
  kern<<<(unsigned)splits, 128, S::TOTAL, stream>>>(mGh, mGl, mXh, mXl, taps, dW, P, (int)kb_per_split,
                                                    oihw_taps, g_lo_on, ws);
  if ((rc = dmc_check_launch("wgrad64_kernel"))) return rc;
  if (ws) {
    const int SL = splits >= 8 ? 8 : (splits >= 4 ? 4 : (splits >= 2 ? 2 : 1));
    wgrad_reduce_kernel<<<dim3((unsigned)cdiv(64L * 64, 32), (unsigned)taps.ntaps), 32 * SL, 0, stream>>>(
        ws, (int)splits, taps, 64, 64, oihw_taps, dW);
    return dmc_check_launch("wgrad_reduce_kernel");
  }
  return DMC_OK;
}

static int fill_taps(TapTable& tt, int ntaps, const int* shift, const int* phase, const int* bsel) {
  if (ntaps < 1 || ntaps > MAX_TAPS) return -1;
  tt.ntaps = ntaps;
  for (int i = 0; i < MAX_TAPS; ++i) {
    tt.shift[i] = i < ntaps ? shift[i] : 0;
    tt.phase[i] = i < ntaps ? phase[i] : 0;
    tt.bsel[i] = i < ntaps ? bsel[i] : 0;
  }
  return 0;
}

static int g_sm_count = 0;
static int sm_count() {
  if (!g_sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (g_sm_count <= 0) g_sm_count = 148;
  }
  return g_sm_count;
}

}  // namespace dmc

using namespace dmc;

static int tap_gemm_impl(const void* A_hi, const void* A_lo, int a_phases, long a_rows, int K,
                         const void* B_hi, const void* B_lo, int b_slices, int N, float* D, long M,
                         int ldD, int Hp, int Wp, int ntaps, const int* shift, const int* phase,
                         const int* bsel, double* stats, const BwFuse& bw, const ActFuse& act, int ring,
                         void* stream, int lda = 0, int stats_ld = 0) {
  DMC_REQUIRE(ring >= 1 && (Hp == 0 || (ring < (Hp & 0xffff) && ring < Wp)), "tap_gemm: ring=%d", ring);
  // fused BN backward: Y / act_hi / gb are addressed like D (row pitch ldD); with ldD > N every pointer,
  // mean / invstd and stats included, is the caller's column sub-range of a wider map
  DMC_REQUIRE(bw.Y == nullptr || (stats && bw.mean && bw.invstd && (ldD == N || stats_ld >= N)),
              "tap_gemm: fused BN backward needs stats, mean, invstd and ldD == N (or an explicit stats_ld)");
  DMC_REQUIRE(lda == 0 || (lda >= K && lda % 8 == 0 && a_phases == 1), "tap_gemm: lda=%d", lda);
  DMC_REQUIRE(stats_ld == 0 || stats_ld >= N, "tap_gemm: stats_ld=%d", stats_ld);
  DMC_REQUIRE(act.bias == nullptr || (bw.Y == nullptr && Hp > 0 && Wp > 0 && ldD == N),
              "tap_gemm: fused activation needs a frame geometry, ldD == N and no BN-backward fusion");
  DMC_REQUIRE(K > 0 && K % 64 == 0, "tap_gemm: K=%d must be a positive multiple of 64", K);
  DMC_REQUIRE(N > 0 && N % 32 == 0, "tap_gemm: N=%d must be a multiple of 32", N);
  DMC_REQUIRE(ldD % 4 == 0 && ldD >= N, "tap_gemm: ldD=%d", ldD);
  DMC_REQUIRE(M > 0 && a_rows > 0 && a_rows < (1L << 31), "tap_gemm: bad rows");
  TapTable tt;
  DMC_REQUIRE(fill_taps(tt, ntaps, shift, phase, bsel) == 0, "tap_gemm: ntaps=%d", ntaps);
  for (int i = 0; i < ntaps; ++i)
    DMC_REQUIRE(phase[i] >= 0 && phase[i] < a_phases && bsel[i] >= 0 && bsel[i] < b_slices,
                "tap_gemm: tap %d out of range", i);
  // 128-wide tiles whenever N exceeds one: a partial last tile (N = 192, 320, 576, 832: the inception maps of
  // I3D) reads zero-filled weight rows and skips the stores, which costs less than 64-wide tiles whose MMAs
  // are bounded by shared-memory operand reads (tensor pipe 35-48 % vs 75-80 %)
  const int BN = (N % 128 == 0 || (N > 128 && N % 64 == 0)) ? 128 : (N % 64 == 0 ? 64 : 32);
  CUtensorMap mAh, mAl, mBh, mBl;
  int rc;
  if ((rc = make_map_3d(&mAh, A_hi, K, a_rows, a_phases, 64, 128, lda))) return rc;
  const int a_lo_on = A_lo != nullptr;     // A_lo == NULL: A is used at bf16 precision (two MMAs per k-step)
  if ((rc = make_map_3d(&mAl, a_lo_on ? A_lo : A_hi, K, a_rows, a_phases, 64, 128, lda))) return rc;
  if ((rc = make_map_3d(&mBh, B_hi, K, N, b_slices, 64, BN))) return rc;
  if ((rc = make_map_3d(&mBl, B_lo, K, N, b_slices, 64, BN))) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int sms = sm_count();
  if (BN == 128)
    return launch_tap_gemm_ws<128, 3>(mAh, mAl, mBh, mBl, tt, D, M, N, ldD, K, Hp, Wp, sms, stats, bw, a_lo_on, act, ring, st,
                                      nullptr, stats_ld);
  if (BN == 64) {
    // K = 64 layers are bounded by the L2 -> shared traffic of the A operand: row-window schedule
    static const bool rw_off = getenv("DMC_NO_ROW_WINDOW") != nullptr;
    RwTable rw;
    bool single_phase = a_phases == 1;
    for (int i = 0; i < ntaps; ++i) single_phase = single_phase && phase[i] == 0;
    if (!rw_off && K == 64 && single_phase && ntaps >= 2 && build_row_windows(tt, rw)) {
      CUtensorMap wAh, wAl;
      if ((rc = make_map_3d(&wAh, A_hi, K, a_rows, 1, 64, RW_ROWS, lda))) return rc;
      if ((rc = make_map_3d(&wAl, a_lo_on ? A_lo : A_hi, K, a_rows, 1, 64, RW_ROWS, lda))) return rc;
      // eight epilogue warps where the epilogue reads global memory (fused BN backward, second gradient
      // source, residual): DMC_EPILOGUE_WARPS=4 / 8 overrides for A/B timing
      static const char* epw_env = getenv("DMC_EPILOGUE_WARPS");
      static const int epw = epw_env ? atoi(epw_env) : EPILOGUE_WARPS_DEFAULT;
      if (epw == 8)
        return launch_tap_gemm_ws<64, 2, true, 8>(wAh, wAl, mBh, mBl, tt, D, M, N, ldD, K, Hp, Wp, sms, stats, bw,
                                                  a_lo_on, act, ring, st, &rw, stats_ld);
      return launch_tap_gemm_ws<64, 2, true>(wAh, wAl, mBh, mBl, tt, D, M, N, ldD, K, Hp, Wp, sms, stats, bw,
                                             a_lo_on, act, ring, st, &rw, stats_ld);
    }
    return launch_tap_gemm_ws<64, 4>(mAh, mAl, mBh, mBl, tt, D, M, N, ldD, K, Hp, Wp, sms, stats, bw, a_lo_on, act, ring, st,
                                     nullptr, stats_ld);
  }
  return launch_tap_gemm_ws<32, 4>(mAh, mAl, mBh, mBl, tt, D, M, N, ldD, K, Hp, Wp, sms, stats, bw, a_lo_on, act, ring, st,
                                   nullptr, stats_ld);
}

// D[M][ldD] (cols n<N) = sum_t A[phase_t][q + shift_t][0:K] . B[bsel_t][n][0:K]
// stats (nullable, zeroed by the caller): double [2][N], += per-column sum and sum of squares of D
// (the BatchNorm batch statistics of a convolution output, fused into the epilogue).
// bw_Y != null switches the epilogue to the fused BatchNorm-backward reduction (see BwFuse):
// D receives dz = (D + bw_gb) * [bw_act_hi > 0] and stats[0]/[1] += sum dz / sum dz * xhat
// (bw_act_hi == NULL: no ReLU mask -- the BatchNorm being differentiated follows no ReLU).
// A_lo == NULL: A is taken at bf16 precision (A_hi only, two MMAs per k-step instead of three).
extern "C" int dmc_tc_tap_gemm(const void* A_hi, const void* A_lo, int a_phases, long a_rows, int K,
                               const void* B_hi, const void* B_lo, int b_slices, int N, float* D,
                               long M, int ldD, int Hp, int Wp, int ntaps, const int* shift,
                               const int* phase, const int* bsel, double* stats, const float* bw_Y,
                               const void* bw_act_hi, const float* bw_gb, const float* bw_mean,
                               const float* bw_invstd, void* stream) {
  BwFuse bw;
  bw.Y = bw_Y; bw.act_hi = (const bf16*)bw_act_hi; bw.gb = bw_gb; bw.mean = bw_mean; bw.invstd = bw_invstd;
  bw.slope = 0.f;
  ActFuse act;
  act.bias = nullptr; act.mask = nullptr; act.slope = 1.f; act.frame_rows = 1u;
  act.res_f32 = nullptr; act.res_hi = nullptr; act.res_lo = nullptr; act.out_hi = nullptr; act.out_lo = nullptr;
  return tap_gemm_impl(A_hi, A_lo, a_phases, a_rows, K, B_hi, B_lo, b_slices, N, D, M, ldD, Hp, Wp, ntaps,
                       shift, phase, bsel, stats, bw, act, 1, stream);
}

// dmc_tc_tap_gemm on a layout whose zero ring is `ring` rows / columns wide (pixel (h, w) at
// (h + ring, w + ring); dilated convolutions of ContextNetwork, code/dmcnet/model.py:31-71, need
// ring >= dilation), with a LeakyReLU slope in the fused BatchNorm-backward epilogue:
// dz = (D + bw_gb) * (bw_act_hi > 0 ? 1 : bw_slope).
extern "C" int dmc_tc_tap_gemm_ring(const void* A_hi, const void* A_lo, int a_phases, long a_rows, int K,
                                    const void* B_hi, const void* B_lo, int b_slices, int N, float* D,
                                    long M, int ldD, int Hp, int Wp, int ring, int ntaps, const int* shift,
                                    const int* phase, const int* bsel, double* stats, const float* bw_Y,
                                    const void* bw_act_hi, const float* bw_gb, const float* bw_mean,
                                    const float* bw_invstd, float bw_slope, void* stream) {
  BwFuse bw;
  bw.Y = bw_Y; bw.act_hi = (const bf16*)bw_act_hi; bw.gb = bw_gb; bw.mean = bw_mean; bw.invstd = bw_invstd;
  bw.slope = bw_slope;
  ActFuse act;
  act.bias = nullptr; act.mask = nullptr; act.slope = 1.f; act.frame_rows = 1u;
  act.res_f32 = nullptr; act.res_hi = nullptr; act.res_lo = nullptr; act.out_hi = nullptr; act.out_lo = nullptr;
  return tap_gemm_impl(A_hi, A_lo, a_phases, a_rows, K, B_hi, B_lo, b_slices, N, D, M, ldD, Hp, Wp, ntaps,
                       shift, phase, bsel, stats, bw, act, ring, stream);
}

// dmc_tc_tap_gemm on COLUMN SUB-RANGES of wider maps (the inception blocks of I3D,
// code/dmcnet_I3D/network/i3d.py:391-432: every branch reads a slice of one concatenated map and writes a
// slice of the next).  A_hi / A_lo point at the first column of the slice, lda = row pitch of that map
// (elements; 0 = K); D and -- with the fused BatchNorm backward -- bw_Y / bw_act_hi / bw_gb / bw_mean /
// bw_invstd are the caller's slices with row pitch ldD; stats is double [2][stats_ld] (0 = N), again offset
// to the slice.  bw_Y == NULL and bw_gb != NULL: D = result + bw_gb (plain sum of two data gradients).
// Hp may carry a padded temporal extent (dmc_pack_hp, common.cuh): rows of the t-ring are masked as well.
extern "C" int dmc_tc_tap_gemm_ex(const void* A_hi, const void* A_lo, int lda, long a_rows, int K,
                                  const void* B_hi, const void* B_lo, int b_slices, int N, float* D, long M,
                                  int ldD, int Hp, int Wp, int ntaps, const int* shift, const int* bsel,
                                  double* stats, int stats_ld, const float* bw_Y, const void* bw_act_hi,
                                  const float* bw_gb, const float* bw_mean, const float* bw_invstd,
                                  void* stream) {
  BwFuse bw;
  bw.Y = bw_Y; bw.act_hi = (const bf16*)bw_act_hi; bw.gb = bw_gb; bw.mean = bw_mean; bw.invstd = bw_invstd;
  bw.slope = 0.f;
  ActFuse act;
  act.bias = nullptr; act.mask = nullptr; act.slope = 1.f; act.frame_rows = 1u;
  act.res_f32 = nullptr; act.res_hi = nullptr; act.res_lo = nullptr; act.out_hi = nullptr; act.out_lo = nullptr;
  int phase[MAX_TAPS] = {0};
  DMC_REQUIRE(ntaps >= 1 && ntaps <= MAX_TAPS, "tap_gemm_ex: ntaps=%d", ntaps);
  return tap_gemm_impl(A_hi, A_lo, 1, a_rows, K, B_hi, B_lo, b_slices, N, D, M, ldD, Hp, Wp, ntaps, shift,
                       phase, bsel, stats, bw, act, 1, stream, lda, stats_ld);
}

// The forward GEMM of a discriminator block (code/dmcnet_GAN/model.py:254-279): same contraction, the
// epilogue applies D = mask[frame][n] * LeakyReLU_slope(D + bias[n]) (ActFuse) before the zero-ring
// mask, the store and the column statistics.  bias [N]; mask [frames][N] or NULL.
extern "C" int dmc_tc_tap_gemm_act(const void* A_hi, const void* A_lo, int a_phases, long a_rows, int K,
                                   const void* B_hi, const void* B_lo, int b_slices, int N, float* D,
                                   long M, int ldD, int Hp, int Wp, int ntaps, const int* shift,
                                   const int* phase, const int* bsel, double* stats, const float* bias,
                                   const float* mask, float slope, void* stream) {
  DMC_REQUIRE(bias != nullptr, "tap_gemm_act: bias is required");
  BwFuse bw;
  bw.Y = nullptr; bw.act_hi = nullptr; bw.gb = nullptr; bw.mean = nullptr; bw.invstd = nullptr;
  bw.slope = 0.f;
  ActFuse act;
  act.bias = bias; act.mask = mask; act.slope = slope; act.frame_rows = (unsigned)(Hp * Wp);
  act.res_f32 = nullptr; act.res_hi = nullptr; act.res_lo = nullptr; act.out_hi = nullptr; act.out_lo = nullptr;
  return tap_gemm_impl(A_hi, A_lo, a_phases, a_rows, K, B_hi, B_lo, b_slices, N, D, M, ldD, Hp, Wp, ntaps,
                       shift, phase, bsel, stats, bw, act, 1, stream);
}

// Inference GEMM with BatchNorm folded into the operands (test.py scoring path, code/dmcnet/test.py:139-151):
// out = act( A * B' + bias [+ residual] ) where B' = W * bn_scale (dmc_weight_fold_prep), bias = bn_shift,
// act = LeakyReLU(slope) (0: ReLU, 1: none), residual = res_f32 [M][N] or res_hi/res_lo (bf16 planes) or
// none.  The result goes to out_hi / out_lo (bf16 planes, the operand of the next GEMM) when given,
// else to D (fp32).  Zero ring of width `ring` preserved.
extern "C" int dmc_tc_tap_gemm_fold(const void* A_hi, const void* A_lo, int a_phases, long a_rows, int K,
                                    const void* B_hi, const void* B_lo, int b_slices, int N, float* D, long M,
                                    int Hp, int Wp, int ring, int ntaps, const int* shift, const int* phase,
                                    const int* bsel, const float* bias, float slope, const float* res_f32,
                                    const void* res_hi, const void* res_lo, void* out_hi, void* out_lo,
                                    void* stream) {
  DMC_REQUIRE(bias != nullptr, "tap_gemm_fold: bias is required");
  DMC_REQUIRE((out_hi == nullptr) == (out_lo == nullptr) && (out_hi || D), "tap_gemm_fold: output");
  DMC_REQUIRE((res_hi == nullptr) == (res_lo == nullptr) && !(res_f32 && res_hi), "tap_gemm_fold: residual");
  BwFuse bw;
  bw.Y = nullptr; bw.act_hi = nullptr; bw.gb = nullptr; bw.mean = nullptr; bw.invstd = nullptr;
  bw.slope = 0.f;
  ActFuse act;
  act.bias = bias; act.mask = nullptr; act.slope = slope; act.frame_rows = (unsigned)(Hp * Wp);
  act.res_f32 = res_f32; act.res_hi = (const bf16*)res_hi; act.res_lo = (const bf16*)res_lo;
  act.out_hi = (bf16*)out_hi; act.out_lo = (bf16*)out_lo;
  return tap_gemm_impl(A_hi, A_lo, a_phases, a_rows, K, B_hi, B_lo, b_slices, N, D, M, N, Hp, Wp, ntaps,
                       shift, phase, bsel, nullptr, bw, act, ring, stream);
}

// dW[bsel_t][Cout][Cin] += sum_q G[q][Cout] * X[phase_t][q + shift_t][Cin]   (caller zeroes dW).
// oihw_taps > 0 writes the OIHW gradient directly instead: dW[co][ci][bsel_t], oihw_taps per filter.
// G_lo == NULL: dY is taken at bf16 precision (G_hi only).
// workspace != NULL (>= dmc_tc_wgrad_workspace() floats): split-K partial tiles go to the workspace
// with plain stores and a second kernel sums them in a fixed order (deterministic, no atomics).
static int wgrad_impl(const void* G_hi, const void* G_lo, long P, int Cout, const void* X_hi,
                      const void* X_lo, int x_phases, int Cin, float* dW, int ntaps, const int* shift,
                      const int* phase, const int* bsel, int oihw_taps, float* workspace,
                      long workspace_floats, void* stream, int ldg, int ldx) {
  DMC_REQUIRE(Cout % 64 == 0 && Cin % 64 == 0, "wgrad: Cout=%d Cin=%d must be multiples of 64", Cout,
              Cin);
  DMC_REQUIRE(P > 0 && P < (1L << 31), "wgrad: bad P");
  DMC_REQUIRE((ldg == 0 || (ldg >= Cout && ldg % 8 == 0)) && (ldx == 0 || (ldx >= Cin && ldx % 8 == 0 && x_phases == 1)),
              "wgrad: ldg=%d ldx=%d", ldg, ldx);
  TapTable tt;
  DMC_REQUIRE(fill_taps(tt, ntaps, shift, phase, bsel) == 0, "wgrad: ntaps=%d", ntaps);
  for (int i = 0; i < ntaps; ++i)
    DMC_REQUIRE(phase[i] >= 0 && phase[i] < x_phases, "wgrad: tap %d phase out of range", i);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // taps are grouped so that consecutive taps share a CTA: they must read the same phase-independent
  // dY tile (always true) -- the activation tile is per tap.
  if (Cout == 64 && Cin == 64 && ntaps <= 10)      // layer1: two taps per M = 128 MMA
    return launch_wgrad64<32, 2>(G_hi, G_lo, X_hi, X_lo, x_phases, tt, dW, P, sm_count(), oihw_taps,
                                 workspace, workspace_floats, st, ldg, ldx);
  if (Cin % 128 == 0)       // wide tiles: one tap per CTA, two CTAs per SM (32-pixel k-blocks)
    return launch_wgrad<128, 1, 32, 3>(G_hi, G_lo, X_hi, X_lo, x_phases, tt, dW, Cout, Cin, P,
                                       sm_count(), oihw_taps, workspace, workspace_floats, st, ldg, ldx);
  return launch_wgrad<64, 5, 64, 2>(G_hi, G_lo, X_hi, X_lo, x_phases, tt, dW, Cout, Cin, P, sm_count(),
                                    oihw_taps, workspace, workspace_floats, st, ldg, ldx);
}

extern "C" int dmc_tc_wgrad(const void* G_hi, const void* G_lo, long P, int Cout, const void* X_hi,
                            const void* X_lo, int x_phases, int Cin, float* dW, int ntaps,
                            const int* shift, const int* phase, const int* bsel, int oihw_taps,
                            float* workspace, long workspace_floats, void* stream) {
  return wgrad_impl(G_hi, G_lo, P, Cout, X_hi, X_lo, x_phases, Cin, dW, ntaps, shift, phase, bsel, oihw_taps,
                    workspace, workspace_floats, stream, 0, 0);
}

// dmc_tc_wgrad on column sub-ranges: G = Cout columns of a [P][ldg] map, X = Cin columns of a [P][ldx] map
// (pointers at the first column of each slice), one phase, up to 27 taps (3 x 3 x 3 kernels of I3D).
extern "C" int dmc_tc_wgrad_ex(const void* G_hi, const void* G_lo, int ldg, long P, int Cout, const void* X_hi,
                               const void* X_lo, int ldx, int Cin, float* dW, int ntaps, const int* shift,
                               const int* bsel, float* workspace, long workspace_floats, void* stream) {
  int phase[MAX_TAPS] = {0};
  DMC_REQUIRE(ntaps >= 1 && ntaps <= MAX_TAPS, "wgrad_ex: ntaps=%d", ntaps);
  return wgrad_impl(G_hi, G_lo, P, Cout, X_hi, X_lo, 1, Cin, dW, ntaps, shift, phase, bsel, 0, workspace,
                    workspace_floats, stream, ldg, ldx);
}

// Floats of split-K workspace dmc_tc_wgrad needs for this shape (0 is never returned; passing
// workspace == NULL selects the atomic-accumulation epilogue instead).
extern "C" long dmc_tc_wgrad_workspace(long P, int Cout, int Cin, int ntaps) {
  long kbs, splits;
  if (Cout == 64 && Cin == 64 && ntaps <= 10)
    wgrad64_plan(32, P, sm_count(), &kbs, &splits);
  else if (Cin % 128 == 0)
    wgrad_plan<1>(128, 32, Cout, Cin, ntaps, P, sm_count(), &kbs, &splits);
  else
    wgrad_plan<5>(64, 64, Cout, Cin, ntaps, P, sm_count(), &kbs, &splits);
  return splits * (long)ntaps * Cout * Cin;
}
