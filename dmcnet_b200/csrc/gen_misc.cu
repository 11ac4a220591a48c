// Resolution helpers of the generator for --gen_flow_ds_factor f != 0 (code/dmcnet/model.py:326-327,
// :335-337, :347-348): the motion vectors and the residual are average-pooled f x f before the
// estimator, and the low-resolution output is TILED f x f times (Tensor.repeat, not an upsampling)
// back to the frame size.  Planar fp32 tensors, one thread per output element; HBM-bound streams.
#include "common.cuh"

namespace dmc {

__global__ void __launch_bounds__(256)
avgpool_planar_kernel(const float* __restrict__ in, int planes, int H, int W, int f, float* __restrict__ out) {
  const int Ho = H / f, Wo = W / f;
  const long total = (long)planes * Ho * Wo;
  const float inv = 1.0f / (float)(f * f);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wo);
    const int y = (int)((i / Wo) % Ho);
    const long p = i / ((long)Wo * Ho);
    const float* src = in + (p * H + (long)y * f) * W + (long)x * f;
    float s = 0.f;
    for (int a = 0; a < f; ++a)
      for (int b = 0; b < f; ++b) s += src[(long)a * W + b];
    out[i] = s * inv;
  }
}

// out[p][y][x] = in[p][y % h][x % w]
__global__ void __launch_bounds__(256)
tile_repeat_kernel(const float* __restrict__ in, int planes, int h, int w, int f, float* __restrict__ out) {
  const int H = h * f, W = w * f;
  const long total = (long)planes * H * W;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const long p = i / ((long)W * H);
    out[i] = in[(p * h + y % h) * w + x % w];
  }
}

// d_in[n][c][y][x] (+)= sum_{a,b < f} d_out[n][c][y + a*h][x + b*w]   (frame strides in elements)
__global__ void __launch_bounds__(256)
tile_sum_kernel(const float* __restrict__ d_out, long do_ns, int C, int h, int w, int f, int N,
                float* __restrict__ d_in, long di_ns, int accumulate) {
  const int H = h * f, W = w * f;
  const long total = (long)N * C * h * w;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const int y = (int)((i / w) % h);
    const int c = (int)((i / ((long)w * h)) % C);
    const int n = (int)(i / ((long)w * h * C));
    const float* src = d_out + (long)n * do_ns + (long)c * H * W;
    float s = 0.f;
    for (int a = 0; a < f; ++a)
      for (int b = 0; b < f; ++b) s += src[(long)(y + a * h) * W + x + b * w];
    float* dst = d_in + (long)n * di_ns + ((long)c * h + y) * w + x;
    *dst = accumulate ? *dst + s : s;
  }
}

static inline unsigned mgrid(long n) {
  long g = cdiv(n, 256);
  if (g > 148L * 16) g = 148L * 16;
  return (unsigned)(g < 1 ? 1 : g);
}

}  // namespace dmc

using namespace dmc;

// nn.AvgPool2d(f, stride=f) on `planes` contiguous [H][W] planes (code/dmcnet/model.py:326-327, :335-337).
extern "C" int dmc_avgpool_planar(const float* in, int planes, int H, int W, int f, float* out, void* stream) {
  DMC_REQUIRE(in && out && planes > 0 && f >= 1 && H % f == 0 && W % f == 0, "avgpool_planar: H=%d W=%d f=%d", H, W, f);
  avgpool_planar_kernel<<<mgrid((long)planes * (H / f) * (W / f)), 256, 0, (cudaStream_t)stream>>>(in, planes, H, W,
                                                                                                 f, out);
  return dmc_check_launch("avgpool_planar_kernel");
}

// Tensor.repeat(1, 1, f, f) on `planes` contiguous [h][w] planes -> [h*f][w*f] (code/dmcnet/model.py:347-348).
extern "C" int dmc_tile_repeat(const float* in, int planes, int h, int w, int f, float* out, void* stream) {
  DMC_REQUIRE(in && out && planes > 0 && f >= 1 && h > 0 && w > 0, "tile_repeat: bad arguments");
  tile_repeat_kernel<<<mgrid((long)planes * h * w * f * f), 256, 0, (cudaStream_t)stream>>>(in, planes, h, w, f, out);
  return dmc_check_launch("tile_repeat_kernel");
}

// Backward of dmc_tile_repeat: d_in[n][c] (+)= sum of the f*f tiles of d_out[n][c]; d_out [N][..][h*f][w*f]
// with frame stride do_ns (its first C planes are read), d_in [N][..][h][w] with frame stride di_ns.
extern "C" int dmc_tile_sum(const float* d_out, long do_ns, int C, int h, int w, int f, int N, float* d_in,
                            long di_ns, int accumulate, void* stream) {
  DMC_REQUIRE(d_out && d_in && N > 0 && C > 0 && f >= 1 && h > 0 && w > 0, "tile_sum: bad arguments");
  tile_sum_kernel<<<mgrid((long)N * C * h * w), 256, 0, (cudaStream_t)stream>>>(d_out, do_ns, C, h, w, f, N, d_in,
                                                                              di_ns, accumulate);
  return dmc_check_launch("tile_sum_kernel");
}
