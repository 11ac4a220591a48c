// Fused multi-tensor Adam over the flat parameter bucket.
//
// The reference builds one torch.optim.Adam param-group PER TENSOR
// (code/dmcnet/train.py:121-142, code/dmcnet_GAN/train.py:122-153) with
// lr = args.lr * decay * lr_mult and weight_decay = wd * decay_mult
// (adjust_learning_rate, train.py:398-408), eps = 1e-3.  Here all tensors of
// one optimizer live in a flat fp32 bucket; a chunk table maps 1024-element
// chunks to tensors and a per-tensor table carries (lr, weight_decay).  The
// update follows torch 2.x `_single_tensor_adam` exactly:
//   g += wd*p; m = lerp(m, g, 1-b1); v = b2*v + (1-b2)*g*g;
//   p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// The step count t lives on the device (one counter per optimizer) so the whole
// train step can be replayed from a CUDA graph.
#include "common.cuh"

namespace dmc {

struct AdamChunk {
  int offset;   // element offset into the flat bucket (multiple of 4)
  int count;    // <= 1024
  int tensor;   // row of the hyper-parameter table
  int pad;
};

__global__ void adam_tick_kernel(int* __restrict__ step) { *step += 1; }

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
            float* __restrict__ v, const AdamChunk* __restrict__ chunks,
            const float* __restrict__ hyper /*[T][2] lr, wd*/, const int* __restrict__ step, float b1,
            float b2, float eps, float grad_scale) {
  const AdamChunk ch = chunks[blockIdx.x];
  __shared__ float s_step_size, s_inv_sqrt_bc2, s_wd;
  if (threadIdx.x == 0) {
    const double t = (double)(*step);
    const double bc1 = 1.0 - pow((double)b1, t);
    const double bc2 = 1.0 - pow((double)b2, t);
    s_step_size = (float)((double)hyper[2 * ch.tensor] / bc1);
    s_inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    s_wd = hyper[2 * ch.tensor + 1];
  }
  __syncthreads();
  const float step_size = s_step_size, isb2 = s_inv_sqrt_bc2, wd = s_wd;
  for (int i = threadIdx.x * 4; i < ch.count; i += 1024) {
    const long o = (long)ch.offset + i;
    float pv[4], gv[4], mv[4], vv[4];
    const int nvalid = ch.count - i >= 4 ? 4 : ch.count - i;
    if (nvalid == 4) {
      *reinterpret_cast<float4*>(pv) = *reinterpret_cast<const float4*>(p + o);
      *reinterpret_cast<float4*>(gv) = *reinterpret_cast<const float4*>(g + o);
      *reinterpret_cast<float4*>(mv) = *reinterpret_cast<const float4*>(m + o);
      *reinterpret_cast<float4*>(vv) = *reinterpret_cast<const float4*>(v + o);
    } else {
      for (int k = 0; k < nvalid; ++k) { pv[k] = p[o + k]; gv[k] = g[o + k]; mv[k] = m[o + k]; vv[k] = v[o + k]; }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < nvalid) {
        float gk = gv[k] * grad_scale;
        gk = fmaf(wd, pv[k], gk);
        mv[k] = mv[k] + (gk - mv[k]) * (1.f - b1);
        vv[k] = b2 * vv[k] + (1.f - b2) * gk * gk;
        const float denom = sqrtf(vv[k]) * isb2 + eps;
        pv[k] = pv[k] - step_size * (mv[k] / denom);
      }
    }
    if (nvalid == 4) {
      *reinterpret_cast<float4*>(p + o) = *reinterpret_cast<float4*>(pv);
      *reinterpret_cast<float4*>(m + o) = *reinterpret_cast<float4*>(mv);
      *reinterpret_cast<float4*>(v + o) = *reinterpret_cast<float4*>(vv);
    } else {
      for (int k = 0; k < nvalid; ++k) { p[o + k] = pv[k]; m[o + k] = mv[k]; v[o + k] = vv[k]; }
    }
  }
}

}  // namespace dmc

using namespace dmc;

// One optimizer step over `nchunks` chunks.  `step` is a device int that this
// call increments first (t starts at 1, as torch).  grad_scale multiplies the
// (all-reduced) gradient before use (1 for sum-of-pre-scaled shards).
extern "C" int dmc_adam_step(float* p, const float* g, float* m, float* v, const void* chunks,
                             int nchunks, const float* hyper, int* step, float beta1, float beta2,
                             float eps, float grad_scale, void* stream) {
  if (nchunks <= 0) return DMC_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  adam_tick_kernel<<<1, 1, 0, st>>>(step);
  dmc_check_launch("adam_tick_kernel");
  adam_kernel<<<nchunks, 256, 0, st>>>(p, g, m, v, reinterpret_cast<const AdamChunk*>(chunks), hyper,
                                       step, beta1, beta2, eps, grad_scale);
  return dmc_check_launch("adam_kernel");
}
