// Shared device helpers for the DMC-Net B200 kernels (sm_100a only).
// PTX wrappers for mbarrier / TMA / tcgen05, the bf16 hi/lo split used by the
// tensor-core path, and the padded pixel-major geometry helpers.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define DMC_OK 0
#define DMC_ERR_INVALID_ARG (-1)
#define DMC_ERR_UNSUPPORTED (-2)
#define DMC_ERR_CUDA (-3)

void dmc_set_error(const char* fmt, ...);
int dmc_check_launch(const char* what);
// Tiled TMA descriptor over an fp32 tensor of `rank` dims (dims[0] contiguous), no swizzle,
// out-of-bounds elements read as zero.  Implemented in gemm_tc.cu.
int dmc_make_f32_map(CUtensorMap* m, const void* base, int rank, const unsigned long long* dims,
                     const unsigned long long* strides_elems /*[rank-1], dims 1..*/,
                     const unsigned int* box);

#define DMC_REQUIRE(cond, ...)                                  \
  do {                                                          \
    if (!(cond)) {                                              \
      dmc_set_error(__VA_ARGS__);                               \
      return DMC_ERR_INVALID_ARG;                               \
    }                                                           \
  } while (0)

namespace dmc {

typedef __nv_bfloat16 bf16;

__host__ __device__ inline long cdiv(long a, long b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- bf16 split
// x ~= hi + lo with |x - hi - lo| <= 2^-17 |x|.  Products use hi*hi + hi*lo +
// lo*hi on the tensor cores (fp32 accumulate), i.e. ~2^-16 relative error.
__device__ __forceinline__ void split_bf16(float x, bf16& hi, bf16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ float join_bf16(bf16 hi, bf16 lo) {
  return __bfloat162float(hi) + __bfloat162float(lo);
}

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------- smem / mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
// Bounded spin: a protocol bug traps (kernel error) instead of hanging the box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 28)) __trap();
  }
}

// ---------------------------------------------------------------- TMA (tiled, 3-D)
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
// One lane of a converged warp (elect.sync).  Unlike `lane == 0` the compiler knows the region behind it is
// executed by a single thread and keeps descriptors / barrier addresses in uniform registers instead of
// wrapping every tcgen05.mma in a per-lane uniformisation loop.
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, %1;\n\t"
      "@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred)
      : "r"(0xffffffffu));
  return pred != 0;
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all previously issued MMAs of this thread retire.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (lane = row).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory operand descriptor, SWIZZLE_128B (layout_type 2, version 1).
//   K-major : rows of 128 B (64 bf16 along K), 8-row atoms; SBO = bytes between atoms.
//   MN-major: rows of 128 B (64 bf16 along M/N), one row per k; SBO = bytes between
//             8-k atoms, LBO = bytes between 64-element M/N groups.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes,
                                                    uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  // The base-offset field (bits 49-51) stays 0 even when the start address is only
  // 128-byte (row) aligned: measured on B200, the MMA unit applies the 128B swizzle to the
  // ABSOLUTE shared-memory address bits, exactly as TMA wrote them, so a descriptor may
  // point at any row of a 1024-byte-aligned tile (setting the field corrupts the result).
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
// Instruction descriptor: bf16 x bf16 -> fp32, M=128, N=n, a/b major bits.
__host__ __device__ inline uint32_t umma_idesc_bf16(int n, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                       // c_format = F32
  d |= 1u << 7;                       // a_format = BF16
  d |= 1u << 10;                      // b_format = BF16
  d |= (uint32_t)(a_mn_major & 1) << 15;
  d |= (uint32_t)(b_mn_major & 1) << 16;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(128 >> 4) << 24;
  return d;
}

// ---------------------------------------------------------------- padded pixel-major geometry
// Activations of the classifier live as [frames][Hp][Wp][C]; q is the flat padded pixel index.
// The zero ring is SHARED: row 0 and column 0 of every frame are zero, and the bottom / right
// neighbours of the last row / column are the next frame's row 0 / the next row's column 0
// (reads past the last frame are TMA out-of-bounds zero fill).  Hp = H + 1 + DMC_PAD_HI,
// Wp = W + 1 + DMC_PAD_HI; DMC_PAD_HI = 1 would restore a private ring on every side
// (81 instead of 64 rows per 7x7 frame, 256 instead of 225 per 14x14 frame).
#define DMC_PAD_HI 0
__host__ __device__ inline int dmc_padded(int h) { return h + 1 + DMC_PAD_HI; }
__host__ __device__ inline int dmc_unpadded(int hp) { return hp - 1 - DMC_PAD_HI; }
// 3-D maps (I3D, code/dmcnet_I3D/network/i3d.py): [clips][T+1][H+1][W+1][C] with the same shared zero ring
// on the low side of EVERY dimension.  The kernels below keep their (Hp, Wp) signature: the padded
// temporal extent travels in the high half of Hp (Hp_arg = Hp | Tp << 16 | t_hi << 30, dmc_pack_hp), and a
// row is interior when its (t, h, w) are all >= 1 -- and, with t_hi = 1, its frame is not the LAST of the
// Tp either (the I3D stem map keeps one zero frame behind each clip as well, csrc/i3d.cu).  Tp = 0 (every
// 2-D caller) costs one uniform branch.
__host__ __device__ inline int dmc_pack_hp(int Hp, int Tp, int t_hi = 0) { return Hp | (Tp << 16) | (t_hi << 30); }
__host__ __device__ inline bool interior(long q, int Hp, int Wp) {
  // q < 2^31 for every tensor this library builds: 32-bit unsigned division is ~5x cheaper
  const unsigned uq = (unsigned)q, uw = (unsigned)Wp, uh = (unsigned)Hp & 0xffffu;
  const unsigned ut = ((unsigned)Hp >> 16) & 0x3fffu, thi = ((unsigned)Hp >> 30) & 1u;
  const unsigned row = uq / uw;
  const unsigned wp = uq - row * uw;
  const unsigned fr = row / uh;
  const unsigned hp = row - fr * uh;
  bool ok = wp >= 1u && wp + DMC_PAD_HI < uw && hp >= 1u && hp + DMC_PAD_HI < uh;
  if (ut) {
    const unsigned tp = fr % ut;
    ok = ok && tp >= 1u && tp + thi < ut;
  }
  return ok;
}

// Same layout with a zero ring of R rows / columns (dilated convolutions: R >= dilation), pixel (h, w)
// at (h + R, w + R); R = 1 is `interior` above.
__host__ __device__ inline bool interior_r(long q, int Hp, int Wp, int R) {
  const unsigned uq = (unsigned)q, uw = (unsigned)Wp, uh = (unsigned)Hp & 0xffffu;
  const unsigned ut = ((unsigned)Hp >> 16) & 0x3fffu, thi = ((unsigned)Hp >> 30) & 1u;
  const unsigned row = uq / uw;
  const unsigned wp = uq - row * uw;
  const unsigned fr = row / uh;
  const unsigned hp = row - fr * uh;
  bool ok = wp >= (unsigned)R && hp >= (unsigned)R;
  if (ut) {
    const unsigned tp = fr % ut;
    ok = ok && tp >= (unsigned)R && tp + thi < ut;
  }
  return ok;
}

// i -> (i / d, i % d) for 32-bit i; shift/mask when d is a power of two (channel counts)
struct FastDiv {
  unsigned d, shift, pow2;
  __host__ __device__ explicit FastDiv(unsigned d_) : d(d_), shift(0), pow2((d_ & (d_ - 1)) == 0) {
    while ((1u << shift) < d_) ++shift;
  }
  __host__ __device__ inline unsigned div(unsigned i) const { return pow2 ? (i >> shift) : (i / d); }
  __host__ __device__ inline unsigned mod(unsigned i) const { return pow2 ? (i & (d - 1)) : (i % d); }
};

}  // namespace dmc
