// Input stage of the train / test step: the CoviarDataSet sample arithmetic
// (code/dmcnet/dataset.py:215-263) on the device, so a batch crosses PCIe as the uint8
// [N][H][W][7] stacks the augmentation produces (7 B/pixel) instead of three fp32 tensors
// (28 B/pixel).
//   unpack_normalize_u8  : split the interleaved stack (channels 0-1 optical flow, 2-3 motion
//                          vectors, 4-6 residual; dataset.py:210,222-224) into the planar fp32
//                          tensors Model.forward takes and normalise them,
//                          (v/255 - 0.5)/mean(std) for flow and mv, (v/255 - 0.5)/std_c for the
//                          residual (dataset.py:251-263).
//   flow_block_mean_u8   : the "blocky" flow target of --flow_ds_factor (dataset.py:226-246 with
//                          upsample_interp = False): mean over factor x factor blocks (zero padded
//                          at the bottom/right edge, as skimage.measure.block_reduce), repeated back
//                          to full size, then normalised like the flow above.
// Both are HBM streams (7 B read + 28 B written per pixel).  The arithmetic is bit-exact with the
// reference: the same IEEE fp32 operations in the same order (two divisions and a subtraction per
// value; the block mean is an exact integer sum divided in double and rounded once to fp32, which is
// what numpy's float64 mean followed by .float() gives).
#include "common.cuh"

namespace dmc {

__device__ __forceinline__ float normalize_u8(float v, float divisor) {
  return __fdiv_rn(__fsub_rn(__fdiv_rn(v, 255.0f), 0.5f), divisor);
}

// One thread = 4 consecutive pixels = 28 bytes = 7 aligned 32-bit words in, 7 float4 out.
__global__ void __launch_bounds__(256)
unpack_normalize_u8_kernel(const unsigned int* __restrict__ frames, long groups, int hw4,
                           float div_motion, float div_r0, float div_r1, float div_r2,
                           float* __restrict__ flow, float* __restrict__ mv, float* __restrict__ res) {
  const float divs[7] = {div_motion, div_motion, div_motion, div_motion, div_r0, div_r1, div_r2};
  for (long g = blockIdx.x * (long)blockDim.x + threadIdx.x; g < groups;
       g += (long)gridDim.x * blockDim.x) {
    unsigned int w[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) w[k] = __ldg(frames + g * 7 + k);
    const long n = g / hw4;              // frame
    const long p4 = g - n * hw4;         // float4 index inside a plane
    float out[7][4];
#pragma unroll
    for (int b = 0; b < 28; ++b) {
      const unsigned int byte = (w[b >> 2] >> (8 * (b & 3))) & 0xffu;
      out[b % 7][b / 7] = normalize_u8((float)byte, divs[b % 7]);
    }
#pragma unroll
    for (int c = 0; c < 7; ++c) {
      float* base;
      if (c < 2) {
        if (flow == nullptr) continue;
        base = flow + ((n * 2 + c) * hw4 + p4) * 4;
      } else if (c < 4) {
        base = mv + ((n * 2 + (c - 2)) * hw4 + p4) * 4;
      } else {
        base = res + ((n * 3 + (c - 4)) * hw4 + p4) * 4;
      }
      *reinterpret_cast<float4*>(base) = make_float4(out[c][0], out[c][1], out[c][2], out[c][3]);
    }
  }
}

constexpr int kMaxBlockCols = 256;

// One CTA = one row of blocks of one frame: `rows` image rows x W pixels x 2 flow channels.
__global__ void __launch_bounds__(256)
flow_block_mean_u8_kernel(const unsigned char* __restrict__ frames, int H, int W, int f,
                          float div_motion, float* __restrict__ flow) {
  __shared__ int s_sum[2 * kMaxBlockCols];
  __shared__ float s_val[2 * kMaxBlockCols];
  const int by = blockIdx.x, n = blockIdx.y;
  const int nbx = (W + f - 1) / f;
  const int y0 = by * f;
  const int rows = min(f, H - y0);
  for (int i = threadIdx.x; i < 2 * nbx; i += blockDim.x) s_sum[i] = 0;
  __syncthreads();
  // work item = (image row r, block column bx, channel c): an exact integer sum of <= f bytes
  const int items = rows * nbx * 2;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int c = it & 1;
    const int bx = (it >> 1) % nbx;
    const int r = (it >> 1) / nbx;
    const int x0 = bx * f, xe = min(x0 + f, W);
    const unsigned char* p = frames + (((long)n * H + y0 + r) * W + x0) * 7 + c;
    int s = 0;
    for (int x = x0; x < xe; ++x, p += 7) s += (int)__ldg(p);
    atomicAdd(&s_sum[c * nbx + bx], s);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * nbx; i += blockDim.x) {
    // zero padded blocks still divide by f*f (block_reduce pads with cval = 0)
    const float mean = (float)((double)s_sum[i] / (double)(f * f));
    s_val[i] = normalize_u8(mean, div_motion);
  }
  __syncthreads();
  const int per_c = rows * W;
  for (int idx = threadIdx.x; idx < 2 * per_c; idx += blockDim.x) {
    const int c = idx / per_c;
    const int rem = idx - c * per_c;
    const int r = rem / W;
    const int x = rem - r * W;
    flow[(((long)n * 2 + c) * H + y0 + r) * W + x] = s_val[c * nbx + x / f];
  }
}

// ---- the same two stages with GroupRandomHorizontalFlip folded in (code/dmcnet/transforms.py:47-58).
// A flipped frame is mirrored left-right and its x components (channel 0 = flow x, channel 2 = mv x)
// become 256 - v: the reference computes (v - 128) * (-1) + 128 in int32, so v = 0 gives 256 -- a value
// a uint8 stack cannot hold, which is why the flip has to happen here and not on the host.
// Arithmetic: a value has only 257 possible inputs (0..255, and 256 for a negated x component), so each
// CTA first tabulates normalize_u8 for the four divisors (motion, residual r/g/b) in shared memory --
// the same function on the same fp32 input, hence the same bits -- and the per-pixel work is 28 LDS
// instead of 56 IEEE divisions (the plain kernel above is ALU-bound on those divisions, DESIGN.md
// section 10).
__global__ void __launch_bounds__(256)
unpack_normalize_flip_u8_kernel(const unsigned int* __restrict__ frames,
                                const unsigned char* __restrict__ flip, long groups, int hw4, int w4,
                                float div_motion, float div_r0, float div_r1, float div_r2,
                                float* __restrict__ flow, float* __restrict__ mv, float* __restrict__ res) {
  __shared__ float lut[4][257];
  for (int i = threadIdx.x; i < 4 * 257; i += blockDim.x) {
    const int k = i / 257, v = i - k * 257;
    const float d = k == 0 ? div_motion : (k == 1 ? div_r0 : (k == 2 ? div_r1 : div_r2));
    lut[k][v] = normalize_u8((float)v, d);
  }
  __syncthreads();
  for (long g = blockIdx.x * (long)blockDim.x + threadIdx.x; g < groups;
       g += (long)gridDim.x * blockDim.x) {
    const long n = g / hw4;
    const long p4 = g - n * hw4;
    const bool f = flip[n] != 0;
    long src = g;
    if (f) {                               // the mirrored 4 pixels are another aligned group of the row
      const long row = p4 / w4, cg = p4 - row * w4;
      src = n * hw4 + row * w4 + (w4 - 1 - cg);
    }
    unsigned int w[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) w[k] = __ldg(frames + src * 7 + k);
    float out[7][4];
#pragma unroll
    for (int b = 0; b < 28; ++b) {
      int v = (int)((w[b >> 2] >> (8 * (b & 3))) & 0xffu);
      const int c = b % 7, p = b / 7;
      if (f && (c == 0 || c == 2)) v = 256 - v;
      const float r = lut[c < 4 ? 0 : c - 3][v];
      if (f) out[c][3 - p] = r; else out[c][p] = r;
    }
#pragma unroll
    for (int c = 0; c < 7; ++c) {
      float* base;
      if (c < 2) {
        if (flow == nullptr) continue;
        base = flow + ((n * 2 + c) * hw4 + p4) * 4;
      } else if (c < 4) {
        base = mv + ((n * 2 + (c - 2)) * hw4 + p4) * 4;
      } else {
        base = res + ((n * 3 + (c - 4)) * hw4 + p4) * 4;
      }
      *reinterpret_cast<float4*>(base) = make_float4(out[c][0], out[c][1], out[c][2], out[c][3]);
    }
  }
}

// Block means of a flipped frame: the flip precedes block_reduce in the reference
// (dataset.py:215 then :226-246), so block bx of the output is block nbx-1-bx of the stored frame
// (W must be a multiple of f for the blocks to coincide) and channel 0 averages 256 - v.
__global__ void __launch_bounds__(256)
flow_block_mean_flip_u8_kernel(const unsigned char* __restrict__ frames,
                               const unsigned char* __restrict__ flip, int H, int W, int f,
                               float div_motion, float* __restrict__ flow) {
  __shared__ int s_sum[2 * kMaxBlockCols];
  __shared__ float s_val[2 * kMaxBlockCols];
  const int by = blockIdx.x, n = blockIdx.y;
  const bool fl = flip[n] != 0;
  const int nbx = W / f;                      // host guarantees W % f == 0
  const int y0 = by * f;
  const int rows = min(f, H - y0);
  for (int i = threadIdx.x; i < 2 * nbx; i += blockDim.x) s_sum[i] = 0;
  __syncthreads();
  const int items = rows * nbx * 2;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int c = it & 1;
    const int bx = (it >> 1) % nbx;
    const int r = (it >> 1) / nbx;
    const unsigned char* p = frames + (((long)n * H + y0 + r) * W + bx * f) * 7 + c;
    int s = 0;
    for (int x = 0; x < f; ++x, p += 7) s += (int)__ldg(p);
    atomicAdd(&s_sum[c * nbx + bx], s);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * nbx; i += blockDim.x) {
    const int c = i / nbx, bx = i - c * nbx;
    int s = s_sum[c * nbx + (fl ? nbx - 1 - bx : bx)];
    if (fl && c == 0) s = 256 * rows * f - s;            // sum of (256 - v) over the valid pixels
    const float mean = (float)((double)s / (double)(f * f));
    s_val[i] = normalize_u8(mean, div_motion);
  }
  __syncthreads();
  const int per_c = rows * W;
  for (int idx = threadIdx.x; idx < 2 * per_c; idx += blockDim.x) {
    const int c = idx / per_c;
    const int rem = idx - c * per_c;
    const int r = rem / W;
    const int x = rem - r * W;
    flow[(((long)n * 2 + c) * H + y0 + r) * W + x] = s_val[c * nbx + x / f];
  }
}

// ---- crop + bilinear resize of the uint8 stacks (GroupMultiScaleCrop / GroupScale + GroupCenterCrop /
// GroupOverSample of code/dmcnet/transforms.py:36-44, :60-114, :122-140: numpy crop, then
// cv2.resize(..., INTER_LINEAR) per channel).  OpenCV's 8-bit linear resize is fixed point: 11-bit
// coefficients, horizontal pass D = S[x0]*a0 + S[x1]*a1, vertical pass
// ((b0*(D0>>4))>>16) + ((b1*(D1>>4))>>16) + 2) >> 2.  The host builds the index / coefficient tables
// (float arithmetic as OpenCV, input_stage.resize_tables; source indices already include the crop
// offset), so the kernel is pure integer arithmetic and bit-exact by construction.
// tab: int32 [ntab][4*Wo + 4*Ho] = {x0[Wo], x1[Wo], a0[Wo], a1[Wo], y0[Ho], y1[Ho], b0[Ho], b1[Ho]};
// frame n uses table n / frames_per_tab (one crop per clip).
__global__ void __launch_bounds__(256)
crop_resize_u8_kernel(const unsigned char* __restrict__ src, int Hs, int Ws,
                      const int* __restrict__ tab, int frames_per_tab,
                      unsigned char* __restrict__ out, int Ho, int Wo, long total) {
  const int tstride = 4 * (Wo + Ho);
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total;
       idx += (long)gridDim.x * blockDim.x) {
    const long n = idx / ((long)Ho * Wo);
    const int r = (int)(idx - n * (long)Ho * Wo);
    const int dy = r / Wo, dx = r - dy * Wo;
    const int* tx = tab + (n / frames_per_tab) * tstride;
    const int* ty = tx + 4 * Wo;
    const int x0 = __ldg(tx + dx), x1 = __ldg(tx + Wo + dx);
    const int a0 = __ldg(tx + 2 * Wo + dx), a1 = __ldg(tx + 3 * Wo + dx);
    const int y0 = __ldg(ty + dy), y1 = __ldg(ty + Ho + dy);
    const int b0 = __ldg(ty + 2 * Ho + dy), b1 = __ldg(ty + 3 * Ho + dy);
    const unsigned char* r0 = src + ((n * Hs + y0) * Ws) * 7;
    const unsigned char* r1 = src + ((n * Hs + y1) * Ws) * 7;
    unsigned char* o = out + idx * 7;
#pragma unroll
    for (int c = 0; c < 7; ++c) {
      const int d0 = (int)__ldg(r0 + x0 * 7 + c) * a0 + (int)__ldg(r0 + x1 * 7 + c) * a1;
      const int d1 = (int)__ldg(r1 + x0 * 7 + c) * a0 + (int)__ldg(r1 + x1 * 7 + c) * a1;
      o[c] = (unsigned char)((((b0 * (d0 >> 4)) >> 16) + ((b1 * (d1 >> 4)) >> 16) + 2) >> 2);
    }
  }
}

}  // namespace dmc

using namespace dmc;
#define ST_(s) reinterpret_cast<cudaStream_t>(s)

// frames: uint8 [N][H][W][7] (flow x,y | mv x,y | residual r,g,b), 4-byte aligned, H*W a multiple of 4.
// flow / mv: fp32 [N][2][H][W], res: fp32 [N][3][H][W], contiguous and 16-byte aligned; flow may be
// NULL (then only mv and res are produced -- use dmc_flow_block_mean_u8 for the flow target).
// div_motion = mean(std) (flow and mv), div_r* = std of the three residual channels
// (code/dmcnet/dataset.py:109-112, :251-263).
extern "C" int dmc_unpack_normalize_u8(const unsigned char* frames, int N, int H, int W,
                                       float div_motion, float div_r0, float div_r1, float div_r2,
                                       float* flow, float* mv, float* res, void* stream) {
  DMC_REQUIRE(N > 0 && H > 0 && W > 0, "unpack_normalize_u8: bad shape");
  DMC_REQUIRE(((long)H * W) % 4 == 0, "unpack_normalize_u8: H*W must be a multiple of 4");
  DMC_REQUIRE(frames && mv && res, "unpack_normalize_u8: null pointer");
  DMC_REQUIRE((reinterpret_cast<size_t>(frames) & 3) == 0, "unpack_normalize_u8: frames must be 4-byte aligned");
  DMC_REQUIRE(((reinterpret_cast<size_t>(flow) | reinterpret_cast<size_t>(mv) |
                reinterpret_cast<size_t>(res)) & 15) == 0,
              "unpack_normalize_u8: outputs must be 16-byte aligned");
  DMC_REQUIRE(div_motion != 0.f && div_r0 != 0.f && div_r1 != 0.f && div_r2 != 0.f,
              "unpack_normalize_u8: zero divisor");
  const int hw4 = (int)(((long)H * W) / 4);
  const long groups = (long)N * hw4;
  long blocks = cdiv(groups, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  unpack_normalize_u8_kernel<<<(int)blocks, 256, 0, ST_(stream)>>>(
      reinterpret_cast<const unsigned int*>(frames), groups, hw4, div_motion, div_r0, div_r1, div_r2,
      flow, mv, res);
  return dmc_check_launch("unpack_normalize_u8_kernel");
}

// flow: fp32 [N][2][H][W] = normalised factor x factor block means of channels 0-1 of `frames`
// (uint8 [N][H][W][7]), each mean repeated over its block (code/dmcnet/dataset.py:226-246,
// upsample_interp = False).  factor >= 1; ceil(W / factor) <= 256.
extern "C" int dmc_flow_block_mean_u8(const unsigned char* frames, int N, int H, int W, int factor,
                                      float div_motion, float* flow, void* stream) {
  DMC_REQUIRE(N > 0 && H > 0 && W > 0 && factor >= 1, "flow_block_mean_u8: bad shape");
  DMC_REQUIRE(frames && flow, "flow_block_mean_u8: null pointer");
  DMC_REQUIRE(div_motion != 0.f, "flow_block_mean_u8: zero divisor");
  const int nbx = (W + factor - 1) / factor, nby = (H + factor - 1) / factor;
  DMC_REQUIRE(nbx <= kMaxBlockCols, "flow_block_mean_u8: more than 256 blocks per row");
  DMC_REQUIRE(N <= 65535, "flow_block_mean_u8: more than 65535 frames per call");
  DMC_REQUIRE((long)factor * factor * 255 < 2147483647L, "flow_block_mean_u8: factor too large");
  flow_block_mean_u8_kernel<<<dim3(nby, N), 256, 0, ST_(stream)>>>(frames, H, W, factor, div_motion, flow);
  return dmc_check_launch("flow_block_mean_u8_kernel");
}

// dmc_unpack_normalize_u8 with GroupRandomHorizontalFlip (code/dmcnet/transforms.py:47-58) folded in:
// flip[n] != 0 mirrors frame n left-right and replaces its x components (channels 0 and 2) by 256 - v
// before the normalisation.  flip: uint8 [N] on the device.  W must be a multiple of 4.
extern "C" int dmc_unpack_normalize_flip_u8(const unsigned char* frames, const unsigned char* flip,
                                            int N, int H, int W, float div_motion, float div_r0,
                                            float div_r1, float div_r2, float* flow, float* mv,
                                            float* res, void* stream) {
  DMC_REQUIRE(N > 0 && H > 0 && W > 0, "unpack_normalize_flip_u8: bad shape");
  DMC_REQUIRE(W % 4 == 0, "unpack_normalize_flip_u8: W must be a multiple of 4");
  DMC_REQUIRE(frames && flip && mv && res, "unpack_normalize_flip_u8: null pointer");
  DMC_REQUIRE((reinterpret_cast<size_t>(frames) & 3) == 0, "unpack_normalize_flip_u8: frames must be 4-byte aligned");
  DMC_REQUIRE(((reinterpret_cast<size_t>(flow) | reinterpret_cast<size_t>(mv) |
                reinterpret_cast<size_t>(res)) & 15) == 0,
              "unpack_normalize_flip_u8: outputs must be 16-byte aligned");
  DMC_REQUIRE(div_motion != 0.f && div_r0 != 0.f && div_r1 != 0.f && div_r2 != 0.f,
              "unpack_normalize_flip_u8: zero divisor");
  const int w4 = W / 4, hw4 = H * w4;
  const long groups = (long)N * hw4;
  long blocks = cdiv(groups, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  unpack_normalize_flip_u8_kernel<<<(int)blocks, 256, 0, ST_(stream)>>>(
      reinterpret_cast<const unsigned int*>(frames), flip, groups, hw4, w4, div_motion, div_r0, div_r1,
      div_r2, flow, mv, res);
  return dmc_check_launch("unpack_normalize_flip_u8_kernel");
}

// dmc_flow_block_mean_u8 for frames that may be flipped (flip: uint8 [N] on the device): the flip
// precedes the block mean in the reference, so W must be a multiple of factor.
extern "C" int dmc_flow_block_mean_flip_u8(const unsigned char* frames, const unsigned char* flip,
                                           int N, int H, int W, int factor, float div_motion,
                                           float* flow, void* stream) {
  DMC_REQUIRE(N > 0 && H > 0 && W > 0 && factor >= 1, "flow_block_mean_flip_u8: bad shape");
  DMC_REQUIRE(frames && flip && flow, "flow_block_mean_flip_u8: null pointer");
  DMC_REQUIRE(div_motion != 0.f, "flow_block_mean_flip_u8: zero divisor");
  DMC_REQUIRE(W % factor == 0, "flow_block_mean_flip_u8: W must be a multiple of factor");
  const int nbx = W / factor, nby = (H + factor - 1) / factor;
  DMC_REQUIRE(nbx <= kMaxBlockCols, "flow_block_mean_flip_u8: more than 256 blocks per row");
  DMC_REQUIRE(N <= 65535, "flow_block_mean_flip_u8: more than 65535 frames per call");
  DMC_REQUIRE((long)factor * factor * 256 < 2147483647L, "flow_block_mean_flip_u8: factor too large");
  flow_block_mean_flip_u8_kernel<<<dim3(nby, N), 256, 0, ST_(stream)>>>(frames, flip, H, W, factor,
                                                                         div_motion, flow);
  return dmc_check_launch("flow_block_mean_flip_u8_kernel");
}

// out: uint8 [N][Ho][Wo][7] = crop + cv2.resize(INTER_LINEAR) of src uint8 [N][Hs][Ws][7]
// (code/dmcnet/transforms.py:122-140; also GroupScale / GroupCenterCrop / GroupOverSample windows).
// tab: device int32 [ntab][4*Wo + 4*Ho] built by the host (see the kernel comment); frame n uses table
// n / frames_per_tab.  Every index in the tables must lie inside the source frame.
extern "C" int dmc_crop_resize_u8(const unsigned char* src, int N, int Hs, int Ws, const int* tab,
                                  int frames_per_tab, unsigned char* out, int Ho, int Wo, void* stream) {
  DMC_REQUIRE(N > 0 && Hs > 0 && Ws > 0 && Ho > 0 && Wo > 0 && frames_per_tab > 0,
              "crop_resize_u8: bad shape");
  DMC_REQUIRE(src && tab && out, "crop_resize_u8: null pointer");
  const long total = (long)N * Ho * Wo;
  long blocks = cdiv(total, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  crop_resize_u8_kernel<<<(int)blocks, 256, 0, ST_(stream)>>>(src, Hs, Ws, tab, frames_per_tab, out, Ho,
                                                               Wo, total);
  return dmc_check_launch("crop_resize_u8_kernel");
}
