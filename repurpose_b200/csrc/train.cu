// First pieces of the training step (BASELINE configs[4]) — memory-bound kernels only:
//   focal_loss_grad   d(sum of the masked sigmoid focal loss)/d(logits)   models/losses.py:5-53 (autograd there),
//                     models/MMCTransformer.py:159-179, scaled by 1/batch_size as main.py:326 does
//   layernorm512_bwd  dx / dgamma / dbeta of nn.LayerNorm(512) (biased variance, eps 1e-5)
//   adam_step         torch.optim.Adam with L2 weight decay folded into the gradient (main.py:190-191),
//                     fp32 master weights, optional bf16 copy for the tensor-core GEMMs
// The GEMM / attention backward kernels these feed are not written yet.
#include <cuda_bf16.h>
#include <math.h>

#include "ptx.cuh"
#include "host_util.h"
#include "kernels.h"

namespace rp {

namespace {

// ---- focal loss gradient -----------------------------------------------------------------------------
// loss_i = alpha_t * ce * (1 - p_t)^gamma,  ce = BCE-with-logits,  p_t = p t + (1 - p)(1 - t)
// d ce / dx = p - t,  d p_t / dx = p (1 - p)(2 t - 1)
// d loss / dx = alpha_t [ (p - t)(1 - p_t)^gamma - gamma ce (1 - p_t)^(gamma - 1) p (1 - p)(2 t - 1) ]
__global__ void focal_loss_grad_kernel(const float* __restrict__ logits, const float* __restrict__ targets,
                                       const uint8_t* __restrict__ mask, int64_t n, float alpha, float gamma,
                                       float scale, float* __restrict__ dlogits) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float g = 0.0f;
  if (mask[i] != 0) {
    const float x = logits[i], t = targets[i];
    const float p = 1.0f / (1.0f + expf(-x));
    const float ce = fmaxf(x, 0.0f) - x * t + log1pf(expf(-fabsf(x)));  // the stable form ATen uses
    const float pt = p * t + (1.0f - p) * (1.0f - t);
    const float om = 1.0f - pt;
    const float mod = gamma == 2.0f ? om * om : powf(om, gamma);
    const float dmod = gamma == 2.0f ? 2.0f * om : (om > 0.0f ? gamma * powf(om, gamma - 1.0f) : 0.0f);
    float d = (p - t) * mod - ce * dmod * p * (1.0f - p) * (2.0f * t - 1.0f);
    if (alpha >= 0.0f) d *= alpha * t + (1.0f - alpha) * (1.0f - t);
    g = d * scale;
  }
  dlogits[i] = g;
}

// ---- LayerNorm(512) backward ---------------------------------------------------------------------------
// One warp per row (16 values per lane), rows strided over the grid; per-lane register accumulators for
// dgamma / dbeta, reduced over the 8 warps of a block through shared memory and written as one partial row per
// block; a second kernel adds the partials in a fixed order (deterministic, no atomics).
constexpr int LNB_WARPS = 8;

__global__ void __launch_bounds__(LNB_WARPS * 32)
layernorm512_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ gamma,
                        int64_t M, float eps, float* __restrict__ dx, float* __restrict__ part_g,
                        float* __restrict__ part_b) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc_g[16], acc_b[16], gm[16];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
    gm[4 * i] = g4.x; gm[4 * i + 1] = g4.y; gm[4 * i + 2] = g4.z; gm[4 * i + 3] = g4.w;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) acc_g[i] = acc_b[i] = 0.0f;
  for (int64_t row = int64_t(blockIdx.x) * LNB_WARPS + warp; row < M; row += int64_t(gridDim.x) * LNB_WARPS) {
    float xv[16], dv[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // lane owns columns 4 (32 i + lane) .. + 3: coalesced 512-byte requests
      const float4 a = __ldg(reinterpret_cast<const float4*>(x + row * 512) + i * 32 + lane);
      const float4 d = __ldg(reinterpret_cast<const float4*>(dy + row * 512) + i * 32 + lane);
      xv[4 * i] = a.x; xv[4 * i + 1] = a.y; xv[4 * i + 2] = a.z; xv[4 * i + 3] = a.w;
      dv[4 * i] = d.x; dv[4 * i + 1] = d.y; dv[4 * i + 2] = d.z; dv[4 * i + 3] = d.w;
    }
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += xv[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / 512.0f);
    float v = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      xv[i] -= mean;
      v = fmaf(xv[i], xv[i], v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = rsqrtf(v * (1.0f / 512.0f) + eps);
    float s1 = 0.0f, s2 = 0.0f;  // sum(dy gamma), sum(dy gamma xhat)
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      xv[i] *= rstd;  // xhat
      acc_b[i] += dv[i];
      acc_g[i] = fmaf(dv[i], xv[i], acc_g[i]);
      dv[i] *= gm[i];
      s1 += dv[i];
      s2 = fmaf(dv[i], xv[i], s2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    s1 *= (1.0f / 512.0f);
    s2 *= (1.0f / 512.0f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 o4;
      o4.x = rstd * (dv[4 * i] - s1 - xv[4 * i] * s2);
      o4.y = rstd * (dv[4 * i + 1] - s1 - xv[4 * i + 1] * s2);
      o4.z = rstd * (dv[4 * i + 2] - s1 - xv[4 * i + 2] * s2);
      o4.w = rstd * (dv[4 * i + 3] - s1 - xv[4 * i + 3] * s2);
      reinterpret_cast<float4*>(dx + row * 512)[i * 32 + lane] = o4;
    }
  }
  __shared__ float sg[LNB_WARPS][512], sb[LNB_WARPS][512];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int col = 4 * (32 * (i >> 2) + lane) + (i & 3);
    sg[warp][col] = acc_g[i];
    sb[warp][col] = acc_b[i];
  }
  __syncthreads();
  for (int col = threadIdx.x; col < 512; col += blockDim.x) {
    float g = 0.0f, b = 0.0f;
#pragma unroll
    for (int w = 0; w < LNB_WARPS; ++w) {
      g += sg[w][col];
      b += sb[w][col];
    }
    part_g[int64_t(blockIdx.x) * 512 + col] = g;
    part_b[int64_t(blockIdx.x) * 512 + col] = b;
  }
}

__global__ void layernorm512_bwd_reduce_kernel(const float* __restrict__ part_g, const float* __restrict__ part_b,
                                               int parts, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= 512) return;
  float g = 0.0f, b = 0.0f;
  for (int p = 0; p < parts; ++p) {
    g += part_g[int64_t(p) * 512 + col];
    b += part_b[int64_t(p) * 512 + col];
  }
  dgamma[col] = g;
  dbeta[col] = b;
}

// ---- Adam (torch.optim.Adam, weight_decay as L2 added to the gradient, no amsgrad) ---------------------------
__global__ void adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, int64_t n, float lr, float beta1, float beta2, float eps,
                                 float weight_decay, float bias_c1, float bias_c2_sqrt,
                                 __nv_bfloat16* __restrict__ p_bf16) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float w = p[i];
  const float gr = fmaf(weight_decay, w, g[i]);
  const float mi = fmaf(beta1, m[i], (1.0f - beta1) * gr);
  const float vi = fmaf(beta2, v[i], (1.0f - beta2) * gr * gr);
  m[i] = mi;
  v[i] = vi;
  // torch: denom = sqrt(v) / sqrt(1 - beta2^t) + eps ; p -= (lr / (1 - beta1^t)) * m / denom
  const float denom = sqrtf(vi) / bias_c2_sqrt + eps;
  const float out = w - (lr / bias_c1) * (mi / denom);
  p[i] = out;
  if (p_bf16 != nullptr) p_bf16[i] = __float2bfloat16_rn(out);
}

}  // namespace

int launch_focal_loss_grad(const float* logits, const float* targets, const uint8_t* mask, int64_t n, float alpha,
                           float gamma, float scale, float* dlogits, cudaStream_t stream) {
  RP_CHECK(n > 0, "focal_loss_grad: empty");
  focal_loss_grad_kernel<<<unsigned((n + 255) / 256), 256, 0, stream>>>(logits, targets, mask, n, alpha, gamma, scale,
                                                                        dlogits);
  count_launch();
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

int64_t layernorm512_bwd_scratch_floats() { return int64_t(2) * 4 * num_sms() * 512; }

int launch_layernorm512_bwd(const float* x, const float* dy, const float* gamma, int64_t M, float eps, float* dx,
                            float* dgamma, float* dbeta, float* scratch, cudaStream_t stream) {
  RP_CHECK(M > 0, "layernorm512_bwd: empty");
  const int sms = num_sms();
  int grid = 4 * sms;
  if (int64_t(grid) * LNB_WARPS > M) grid = int((M + LNB_WARPS - 1) / LNB_WARPS);
  float* part_g = scratch;
  float* part_b = scratch + int64_t(4) * sms * 512;
  layernorm512_bwd_kernel<<<grid, LNB_WARPS * 32, 0, stream>>>(x, dy, gamma, M, eps, dx, part_g, part_b);
  layernorm512_bwd_reduce_kernel<<<4, 128, 0, stream>>>(part_g, part_b, grid, dgamma, dbeta);
  count_launch(2);
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

int launch_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                     float eps, float weight_decay, int step, void* p_bf16, cudaStream_t stream) {
  RP_CHECK(n > 0 && step >= 1, "adam_step: n and step must be positive");
  const float c1 = 1.0f - powf(beta1, float(step));
  const float c2s = sqrtf(1.0f - powf(beta2, float(step)));
  adam_step_kernel<<<unsigned((n + 255) / 256), 256, 0, stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, c1,
                                                                  c2s, reinterpret_cast<__nv_bfloat16*>(p_bf16));
  count_launch();
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

}  // namespace rp
