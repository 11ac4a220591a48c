// Loss heads and small dense layers.
//   linear_fwd / bwd        : classifier fc (512 -> num_class) and the
//                             discriminator's adv_layer (25088 -> 2).
//   ce_head                 : TSN segment consensus (mean of the logits over the
//                             S segments of a clip, code/dmcnet/train.py:239-240),
//                             cross-entropy (train.py:241), its gradient w.r.t. the
//                             per-frame logits, and top-1 / top-5 hits
//                             (train.py:411-424) in one launch.  With S = 1 and
//                             C = 2 it is the adversarial CE (GAN/train.py:274,346).
//   mse_head                : flow-reconstruction MSE (train.py:245) and its
//                             gradient in one streaming pass.
// All losses are reported as SUMS over the local shard plus a caller-provided
// global normaliser, so that a sum all-reduce reproduces the global mean.
#include "common.cuh"

namespace dmc {

__global__ void __launch_bounds__(128)
linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                  const float* __restrict__ b, int K, int N, float* __restrict__ out) {
  const int m = blockIdx.x, j = blockIdx.y;
  const float* xp = x + (long)m * K;
  const float* wp = w + (long)j * K;
  float s = 0.f;
  for (int k = threadIdx.x; k < K; k += 128) s = fmaf(xp[k], wp[k], s);
  __shared__ float red[4];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) out[(long)m * N + j] = red[0] + red[1] + red[2] + red[3] + (b ? b[j] : 0.f);
}

// dx[m][k] = sum_j dy[m][j] * w[j][k]
__global__ void linear_bwd_data_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                       int M, int K, int N, float* __restrict__ dx) {
  const long total = (long)M * K;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const long m = i / K;
    float s = 0.f;
    for (int j = 0; j < N; ++j) s = fmaf(dy[m * N + j], w[(long)j * K + k], s);
    dx[i] = s;
  }
}

// dw[j][k] = sum_m dy[m][j] * x[m][k];  db[j] = sum_m dy[m][j]
__global__ void linear_bwd_weight_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                         int M, int K, int N, float* __restrict__ dw,
                                         float* __restrict__ db) {
  const long total = (long)N * K;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const int j = (int)(i / K);
    float s = 0.f, sb = 0.f;
    for (int m = 0; m < M; ++m) {
      const float d = dy[(long)m * N + j];
      s = fmaf(d, x[(long)m * K + k], s);
      sb += d;
    }
    dw[i] = s;
    if (k == 0 && db) db[j] = sb;
  }
}

// out_stats: [0] loss sum over clips, [1] top-1 hits, [2] top-5 hits  (fp32, overwritten)
// consensus[b][c] written when non-null.  dlogits[n][c] = gscale * (softmax - onehot) / S.
__global__ void __launch_bounds__(256)
ce_head_kernel(const float* __restrict__ logits, int B, int S, int C,
               const long long* __restrict__ target, float gscale, float* __restrict__ consensus,
               float* __restrict__ dlogits, float* __restrict__ out_stats) {
  __shared__ float s_loss[256], s_t1[256], s_t5[256];
  float loss = 0.f, t1 = 0.f, t5 = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float* lp = logits + (long)b * S * C;
    const int tgt = (int)target[b];
    const float invS = 1.f / (float)S;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) {
      float v = 0.f;
      for (int s = 0; s < S; ++s) v += lp[s * C + c];
      v *= invS;
      mx = fmaxf(mx, v);
    }
    float den = 0.f, vt = 0.f;
    for (int c = 0; c < C; ++c) {
      float v = 0.f;
      for (int s = 0; s < S; ++s) v += lp[s * C + c];
      v *= invS;
      if (consensus) consensus[(long)b * C + c] = v;
      den += expf(v - mx);
      if (c == tgt) vt = v;
    }
    const float lse = mx + logf(den);
    loss += lse - vt;
    int rank = 0;
    for (int c = 0; c < C; ++c) {
      float v = 0.f;
      for (int s = 0; s < S; ++s) v += lp[s * C + c];
      v *= invS;
      // torch.topk order: larger value first, lower index first among equals
      if (v > vt || (v == vt && c < tgt)) ++rank;
      if (dlogits) {
        const float p = expf(v - lse);
        const float g = gscale * invS * (p - (c == tgt ? 1.f : 0.f));
        for (int s = 0; s < S; ++s) dlogits[((long)b * S + s) * C + c] = g;
      }
    }
    t1 += rank < 1 ? 1.f : 0.f;
    t5 += rank < 5 ? 1.f : 0.f;
  }
  s_loss[threadIdx.x] = loss;
  s_t1[threadIdx.x] = t1;
  s_t5[threadIdx.x] = t5;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b1 = 0, b5 = 0;
    for (int i = 0; i < blockDim.x; ++i) { a += s_loss[i]; b1 += s_t1[i]; b5 += s_t5[i]; }
    out_stats[0] = (float)a;
    out_stats[1] = (float)b1;
    out_stats[2] = (float)b5;
  }
}

// loss_sum += sum (gen-flow)^2 (double);  dgen = gscale * (gen - flow)  (gscale = 2*lr_mse/numel)
__global__ void __launch_bounds__(256)
mse_head_kernel(const float* __restrict__ gen, const float* __restrict__ flow, long n4, float gscale,
                float* __restrict__ dgen, long frame4, long dgen_ns4, double* __restrict__ loss_sum) {
  float s = 0.f;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(gen)[i];
    const float4 b = reinterpret_cast<const float4*>(flow)[i];
    const float4 d = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
    s += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
    if (dgen) {      // dgen may be a channel slice of a wider buffer: per-frame stride dgen_ns4
      const long o = frame4 == dgen_ns4 ? i : (i / frame4) * dgen_ns4 + (i % frame4);
      reinterpret_cast<float4*>(dgen)[o] =
          make_float4(gscale * d.x, gscale * d.y, gscale * d.z, gscale * d.w);
    }
  }
  __shared__ double red[8];
  double sd = warp_sum_d((double)s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sd;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) sd += red[i];
    atomicAdd(loss_sum, sd);
  }
}

// Flow-reconstruction criterion chosen by --loss-mse (code/dmcnet/train.py:166-172):
// KIND 0 MSELoss, 1 SmoothL1Loss (beta = 1), 2 L1Loss, all with mean reduction.
//   value  f(d): d^2 | (|d| < 1 ? d^2/2 : |d| - 1/2) | |d|
//   slope f'(d): 2d  | clamp(d, -1, 1)               | sign(d)  (0 at d = 0, as ATen)
// loss_sum += sum f(gen-flow) (double);  dgen = gscale * f'(gen-flow), gscale = lr_mse / numel_global.
template <int KIND>
__device__ __forceinline__ void flow_loss_term(float d, float& val, float& slope) {
  if (KIND == 0) {
    val = d * d;
    slope = 2.f * d;
  } else if (KIND == 1) {
    const float a = fabsf(d);
    val = a < 1.f ? 0.5f * d * d : a - 0.5f;
    slope = fminf(fmaxf(d, -1.f), 1.f);
  } else {
    val = fabsf(d);
    slope = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
  }
}

template <int KIND>
__global__ void __launch_bounds__(256)
flow_loss_head_kernel(const float* __restrict__ gen, const float* __restrict__ flow, long n4,
                      float gscale, float* __restrict__ dgen, long frame4, long dgen_ns4,
                      double* __restrict__ loss_sum) {
  float s = 0.f;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(gen)[i];
    const float4 b = reinterpret_cast<const float4*>(flow)[i];
    float4 v, g;
    flow_loss_term<KIND>(a.x - b.x, v.x, g.x);
    flow_loss_term<KIND>(a.y - b.y, v.y, g.y);
    flow_loss_term<KIND>(a.z - b.z, v.z, g.z);
    flow_loss_term<KIND>(a.w - b.w, v.w, g.w);
    s += (v.x + v.y) + (v.z + v.w);
    if (dgen) {      // dgen may be a channel slice of a wider buffer: per-frame stride dgen_ns4
      const long o = frame4 == dgen_ns4 ? i : (i / frame4) * dgen_ns4 + (i % frame4);
      reinterpret_cast<float4*>(dgen)[o] =
          make_float4(gscale * g.x, gscale * g.y, gscale * g.z, gscale * g.w);
    }
  }
  __shared__ double red[8];
  double sd = warp_sum_d((double)s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sd;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) sd += red[i];
    atomicAdd(loss_sum, sd);
  }
}

// Attention-weighted flow criterion, criterion(att * gen, att * flow) of code/dmcnet/train.py:244-247
// (--att 1): with u = att * (gen - flow), loss += crit(u), d/dgen = att * crit'(u),
// d/datt = (gen - flow) * crit'(u).  gen, flow, att, datt are dense [numel]; dgen may be strided.
template <int KIND>
__global__ void __launch_bounds__(256)
att_flow_loss_head_kernel(const float* __restrict__ gen, const float* __restrict__ flow,
                          const float* __restrict__ att, long n4, float gscale, float* __restrict__ dgen,
                          long frame4, long dgen_ns4, float* __restrict__ datt, double* __restrict__ loss_sum) {
  float s = 0.f;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(gen)[i];
    const float4 b = reinterpret_cast<const float4*>(flow)[i];
    const float4 w = reinterpret_cast<const float4*>(att)[i];
    const float d[4] = {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w};
    const float ww[4] = {w.x, w.y, w.z, w.w};
    float gg[4], ga[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v, g;
      flow_loss_term<KIND>(ww[k] * d[k], v, g);
      s += v;
      gg[k] = gscale * ww[k] * g;
      ga[k] = gscale * d[k] * g;
    }
    if (dgen) {
      const long o = frame4 == dgen_ns4 ? i : (i / frame4) * dgen_ns4 + (i % frame4);
      reinterpret_cast<float4*>(dgen)[o] = make_float4(gg[0], gg[1], gg[2], gg[3]);
      reinterpret_cast<float4*>(datt)[i] = make_float4(ga[0], ga[1], ga[2], ga[3]);
    }
  }
  __shared__ double red[8];
  double sd = warp_sum_d((double)s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sd;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) sd += red[i];
    atomicAdd(loss_sum, sd);
  }
}

}  // namespace dmc

using namespace dmc;
#define ST_(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int dmc_linear_fwd(const float* x, const float* w, const float* b, int M, int K, int N,
                              float* out, void* stream) {
  linear_fwd_kernel<<<dim3(M, N), 128, 0, ST_(stream)>>>(x, w, b, K, N, out);
  return dmc_check_launch("linear_fwd_kernel");
}

extern "C" int dmc_linear_bwd(const float* dy, const float* x, const float* w, int M, int K, int N,
                              float* dx, float* dw, float* db, void* stream) {
  if (dx) {
    long blocks = cdiv((long)M * K, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    linear_bwd_data_kernel<<<(int)blocks, 256, 0, ST_(stream)>>>(dy, w, M, K, N, dx);
    int rc = dmc_check_launch("linear_bwd_data_kernel");
    if (rc) return rc;
  }
  if (dw) {
    long blocks = cdiv((long)N * K, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    linear_bwd_weight_kernel<<<(int)blocks, 256, 0, ST_(stream)>>>(dy, x, M, K, N, dw, db);
    return dmc_check_launch("linear_bwd_weight_kernel");
  }
  return DMC_OK;
}

extern "C" int dmc_ce_head(const float* logits, int B, int S, int C, const long long* target,
                           float gscale, float* consensus, float* dlogits, float* out_stats,
                           void* stream) {
  DMC_REQUIRE(B > 0 && S > 0 && C > 0, "ce_head: bad shape");
  ce_head_kernel<<<1, 256, 0, ST_(stream)>>>(logits, B, S, C, target, gscale, consensus, dlogits,
                                             out_stats);
  return dmc_check_launch("ce_head_kernel");
}

// dgen[n][0:frame_elems] (per-frame stride dgen_ns elements) = gscale * (gen - flow)
extern "C" int dmc_mse_head(const float* gen, const float* flow, long numel, float gscale,
                            float* dgen, long frame_elems, long dgen_ns, double* loss_sum,
                            void* stream) {
  DMC_REQUIRE(numel % 4 == 0 && frame_elems % 4 == 0 && dgen_ns % 4 == 0 && frame_elems > 0,
              "mse_head: sizes must be multiples of 4");
  if (cudaMemsetAsync(loss_sum, 0, sizeof(double), ST_(stream)) != cudaSuccess)
    return dmc_check_launch("mse_head memset");
  long blocks = cdiv(numel / 4, 256 * 4);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  mse_head_kernel<<<(int)blocks, 256, 0, ST_(stream)>>>(gen, flow, numel / 4, gscale, dgen,
                                                        frame_elems / 4, dgen_ns / 4, loss_sum);
  return dmc_check_launch("mse_head_kernel");
}

// Flow-reconstruction loss of --loss-mse (code/dmcnet/train.py:166-172, :245): kind 0 = MSELoss,
// 1 = SmoothL1Loss, 2 = L1Loss.  loss_sum (double, overwritten) = sum of the per-element loss;
// dgen[n][0:frame_elems] (per-frame stride dgen_ns elements; may be NULL) = gscale * dloss/dgen with
// gscale = weight / (global element count), i.e. the mean reduction is the caller's normaliser.
extern "C" int dmc_flow_loss_head(int kind, const float* gen, const float* flow, long numel,
                                  float gscale, float* dgen, long frame_elems, long dgen_ns,
                                  double* loss_sum, void* stream) {
  DMC_REQUIRE(kind >= 0 && kind <= 2, "flow_loss_head: kind must be 0 (MSE), 1 (SmoothL1) or 2 (L1)");
  DMC_REQUIRE(numel % 4 == 0 && frame_elems % 4 == 0 && dgen_ns % 4 == 0 && frame_elems > 0,
              "flow_loss_head: sizes must be multiples of 4");
  if (cudaMemsetAsync(loss_sum, 0, sizeof(double), ST_(stream)) != cudaSuccess)
    return dmc_check_launch("flow_loss_head memset");
  long blocks = cdiv(numel / 4, 256 * 4);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  const long n4 = numel / 4, f4 = frame_elems / 4, d4 = dgen_ns / 4;
  if (kind == 0)
    flow_loss_head_kernel<0><<<(int)blocks, 256, 0, ST_(stream)>>>(gen, flow, n4, gscale, dgen, f4, d4, loss_sum);
  else if (kind == 1)
    flow_loss_head_kernel<1><<<(int)blocks, 256, 0, ST_(stream)>>>(gen, flow, n4, gscale, dgen, f4, d4, loss_sum);
  else
    flow_loss_head_kernel<2><<<(int)blocks, 256, 0, ST_(stream)>>>(gen, flow, n4, gscale, dgen, f4, d4, loss_sum);
  return dmc_check_launch("flow_loss_head_kernel");
}

// --att 1 (code/dmcnet/train.py:244-247, GAN :349-352): criterion(att * gen, att * flow), same kinds.
// dgen (strided like dmc_flow_loss_head) = gscale * att * crit'(u), datt [numel] = gscale * (gen - flow) *
// crit'(u), u = att * (gen - flow); both NULL for evaluation.
extern "C" int dmc_att_flow_loss_head(int kind, const float* gen, const float* flow, const float* att,
                                      long numel, float gscale, float* dgen, long frame_elems, long dgen_ns,
                                      float* datt, double* loss_sum, void* stream) {
  DMC_REQUIRE(kind >= 0 && kind <= 2, "att_flow_loss_head: kind must be 0 (MSE), 1 (SmoothL1) or 2 (L1)");
  DMC_REQUIRE(numel % 4 == 0 && frame_elems % 4 == 0 && dgen_ns % 4 == 0 && frame_elems > 0,
              "att_flow_loss_head: sizes must be multiples of 4");
  DMC_REQUIRE(att != nullptr && (dgen == nullptr) == (datt == nullptr), "att_flow_loss_head: att / dgen / datt");
  if (cudaMemsetAsync(loss_sum, 0, sizeof(double), ST_(stream)) != cudaSuccess)
    return dmc_check_launch("att_flow_loss_head memset");
  long blocks = cdiv(numel / 4, 256 * 4);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  const long n4 = numel / 4, f4 = frame_elems / 4, d4 = dgen_ns / 4;
  if (kind == 0)
    att_flow_loss_head_kernel<0><<<(int)blocks, 256, 0, ST_(stream)>>>(gen, flow, att, n4, gscale, dgen, f4, d4, datt, loss_sum);
  else if (kind == 1)
    att_flow_loss_head_kernel<1><<<(int)blocks, 256, 0, ST_(stream)>>>(gen, flow, att, n4, gscale, dgen, f4, d4, datt, loss_sum);
  else
    att_flow_loss_head_kernel<2><<<(int)blocks, 256, 0, ST_(stream)>>>(gen, flow, att, n4, gscale, dgen, f4, d4, datt, loss_sum);
  return dmc_check_launch("att_flow_loss_head_kernel");
}
