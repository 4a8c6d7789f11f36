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

__global__ void __launch_bounds__(LNB_WARPS * 32, 2)
layernorm512_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ gamma,
                        int64_t M, float eps, float* __restrict__ dx, float* __restrict__ part_g,
                        float* __restrict__ part_b, int accumulate, __nv_bfloat16* __restrict__ dx_bf16,
                        float* __restrict__ part_c, const DropKey drop) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc_g[16], acc_b[16], acc_c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc_g[i] = acc_b[i] = acc_c[i] = 0.0f;
  for (int64_t row = int64_t(blockIdx.x) * LNB_WARPS + warp; row < M; row += int64_t(gridDim.x) * LNB_WARPS) {
    float xv[16], dv[16], old[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // lane owns columns 4 (32 i + lane) .. + 3: coalesced 512-byte requests
      const float4 a = __ldcs(reinterpret_cast<const float4*>(x + row * 512) + i * 32 + lane);
      const float4 d = __ldcs(reinterpret_cast<const float4*>(dy + row * 512) + i * 32 + lane);
      xv[4 * i] = a.x; xv[4 * i + 1] = a.y; xv[4 * i + 2] = a.z; xv[4 * i + 3] = a.w;
      dv[4 * i] = d.x; dv[4 * i + 1] = d.y; dv[4 * i + 2] = d.z; dv[4 * i + 3] = d.w;
    }
    if (accumulate) {  // the residual stream's gradient so far, requested together with x and dy
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 o = *(reinterpret_cast<const float4*>(dx + row * 512) + i * 32 + lane);
        old[4 * i] = o.x; old[4 * i + 1] = o.y; old[4 * i + 2] = o.z; old[4 * i + 3] = o.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) old[i] = 0.0f;
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
    for (int i = 0; i < 4; ++i) {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
      const float gm[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = 4 * i + e;
        xv[k] *= rstd;  // xhat
        acc_b[k] += dv[k];
        acc_g[k] = fmaf(dv[k], xv[k], acc_g[k]);
        dv[k] *= gm[e];
        s1 += dv[k];
        s2 = fmaf(dv[k], xv[k], s2);
      }
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
      o4.x = old[4 * i] + rstd * (dv[4 * i] - s1 - xv[4 * i] * s2);
      o4.y = old[4 * i + 1] + rstd * (dv[4 * i + 1] - s1 - xv[4 * i + 1] * s2);
      o4.z = old[4 * i + 2] + rstd * (dv[4 * i + 2] - s1 - xv[4 * i + 2] * s2);
      o4.w = old[4 * i + 3] + rstd * (dv[4 * i + 3] - s1 - xv[4 * i + 3] * s2);
      *(reinterpret_cast<float4*>(dx + row * 512) + i * 32 + lane) = o4;
      if (drop.thr != 0u) {  // dropout1 / dropout2 backward: the branch's Linear sees kept ? dx * scale : 0
        const uint32_t k = (uint32_t(row) * 512u + uint32_t(4 * (32 * i + lane))) >> 1;
        drop_pair(k, drop, o4.x, o4.y);
        drop_pair(k + 1u, drop, o4.z, o4.w);
      }
      acc_c[4 * i] += o4.x; acc_c[4 * i + 1] += o4.y; acc_c[4 * i + 2] += o4.z; acc_c[4 * i + 3] += o4.w;
      if (dx_bf16 != nullptr) {
        uint2 pk;
        pk.x = pack_bf16x2(o4.x, o4.y);
        pk.y = pack_bf16x2(o4.z, o4.w);
        reinterpret_cast<uint2*>(dx_bf16 + row * 512)[i * 32 + lane] = pk;
      }
    }
  }
  // per-block partials of dgamma, dbeta and (optional) the column sums of the final dx: one pass through shared
  // memory per quantity (16 KB), summed over the 8 warps in a fixed order
  __shared__ float sred[LNB_WARPS][512];
  auto block_reduce = [&](const float (&acc)[16], float* part) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) sred[warp][4 * (32 * (i >> 2) + lane) + (i & 3)] = acc[i];
    __syncthreads();
    for (int col = threadIdx.x; col < 512; col += blockDim.x) {
      float t = 0.0f;
#pragma unroll
      for (int w = 0; w < LNB_WARPS; ++w) t += sred[w][col];
      part[int64_t(blockIdx.x) * 512 + col] = t;
    }
  };
  block_reduce(acc_g, part_g);
  block_reduce(acc_b, part_b);
  if (part_c != nullptr) block_reduce(acc_c, part_c);
}

__global__ void layernorm512_bwd_reduce_kernel(const float* __restrict__ part_g, const float* __restrict__ part_b,
                                               const float* __restrict__ part_c, int parts, float* __restrict__ dgamma,
                                               float* __restrict__ dbeta, float* __restrict__ dcol) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= 512) return;
  float g = 0.0f, b = 0.0f, c = 0.0f;
  for (int p = 0; p < parts; ++p) {
    g += part_g[int64_t(p) * 512 + col];
    b += part_b[int64_t(p) * 512 + col];
    if (part_c != nullptr) c += part_c[int64_t(p) * 512 + col];
  }
  dgamma[col] = g;
  dbeta[col] = b;
  if (dcol != nullptr) dcol[col] = c;
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


// ---- split-K partial sums -> gradient (fixed summation order: deterministic) ---------------------------------
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  float4 acc = *reinterpret_cast<const float4*>(part + i);
  for (int s = 1; s < splits; ++s) {
    const float4 v = *reinterpret_cast<const float4*>(part + int64_t(s) * n + i);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  *reinterpret_cast<float4*>(out + i) = acc;
}

// ---- column sums of a bf16 [M, N] matrix (bias gradients), two fixed-order stages ----------------------------
// stage 1: block (bx, by) sums rows [by * rows_per, ...) of the 256 columns bx*256.. (thread = column pair... one column
// per thread, coalesced 512-byte row segments); stage 2 adds the per-block partials in order.
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, int64_t M, int N, int64_t rows_per, float* __restrict__ part) {
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col >= N) return;
  const int64_t r0 = int64_t(blockIdx.y) * rows_per;
  const int64_t r1 = r0 + rows_per < M ? r0 + rows_per : M;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int64_t r = r0;
  for (; r + 3 < r1; r += 4) {
    a0 += __bfloat162float(x[r * N + col]);
    a1 += __bfloat162float(x[(r + 1) * N + col]);
    a2 += __bfloat162float(x[(r + 2) * N + col]);
    a3 += __bfloat162float(x[(r + 3) * N + col]);
  }
  for (; r < r1; ++r) a0 += __bfloat162float(x[r * N + col]);
  part[int64_t(blockIdx.y) * N + col] = (a0 + a1) + (a2 + a3);
}
__global__ void colsum_reduce_kernel(const float* __restrict__ part, int parts, int N, float* __restrict__ out) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= N) return;
  float a = 0.f;
  for (int p = 0; p < parts; ++p) a += part[int64_t(p) * N + col];
  out[col] = a;
}

// ---- ReLU backward: dy = act > 0 ? dy : 0 (bf16 in place; fp32 gradient against an fp32 activation) ----------
// two bf16 gradients scaled by `scale` (dropout backward; scale == 1: untouched bits)
__device__ __forceinline__ uint32_t scale_bf16x2(uint32_t dv, float scale) {
  if (scale == 1.0f) return dv;
  return pack_bf16x2(__uint_as_float(dv << 16) * scale, __uint_as_float(dv & 0xffff0000u) * scale);
}
__global__ void relu_bwd_bf16_kernel(__nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ act, int64_t n8,
                                     float scale) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  uint4 d = reinterpret_cast<const uint4*>(dy)[i];
  const uint4 a = __ldcs(reinterpret_cast<const uint4*>(act) + i);
  // bf16 > 0: sign bit clear and not zero
  auto mask2 = [scale](uint32_t dv, uint32_t av) {
    const uint32_t lo = (av & 0x8000u) == 0u && (av & 0x7fffu) != 0u ? 0xffffu : 0u;
    const uint32_t hi = (av & 0x80000000u) == 0u && (av & 0x7fff0000u) != 0u ? 0xffff0000u : 0u;
    return scale_bf16x2(dv & (lo | hi), scale);
  };
  d.x = mask2(d.x, a.x); d.y = mask2(d.y, a.y); d.z = mask2(d.z, a.z); d.w = mask2(d.w, a.w);
  reinterpret_cast<uint4*>(dy)[i] = d;
}
// ReLU backward fused with the bias gradient of the Linear in front of the ReLU: dy = act > 0 ? dy : 0 in place and
// per-block partial column sums of the masked dy (block = a strip of rows, thread = groups of 8 consecutive columns)
__global__ void __launch_bounds__(256)
relu_bwd_colsum_bf16_kernel(__nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ act, int64_t M, int N,
                            int64_t rows_per, float* __restrict__ part, float scale) {
  const int64_t r0 = int64_t(blockIdx.x) * rows_per;
  const int64_t r1 = r0 + rows_per < M ? r0 + rows_per : M;
  const int groups = N >> 3;
  for (int cg = threadIdx.x; cg < groups; cg += blockDim.x) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    auto one = [&](int64_t r) {
      uint4* dp = reinterpret_cast<uint4*>(dy + r * N) + cg;
      uint4 d = *dp;
      const uint4 a = __ldcs(reinterpret_cast<const uint4*>(act + r * N) + cg);
      uint32_t dw[4] = {d.x, d.y, d.z, d.w};
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t lo = (aw[i] & 0x8000u) == 0u && (aw[i] & 0x7fffu) != 0u ? 0xffffu : 0u;
        const uint32_t hi = (aw[i] & 0x80000000u) == 0u && (aw[i] & 0x7fff0000u) != 0u ? 0xffff0000u : 0u;
        dw[i] = scale_bf16x2(dw[i] & (lo | hi), scale);
        acc[2 * i] += __uint_as_float(dw[i] << 16);
        acc[2 * i + 1] += __uint_as_float(dw[i] & 0xffff0000u);
      }
      *dp = make_uint4(dw[0], dw[1], dw[2], dw[3]);
    };
    int64_t r = r0;
    for (; r + 3 < r1; r += 4) {
      one(r); one(r + 1); one(r + 2); one(r + 3);
    }
    for (; r < r1; ++r) one(r);
#pragma unroll
    for (int i = 0; i < 8; ++i) part[int64_t(blockIdx.x) * N + 8 * cg + i] = acc[i];
  }
}

__global__ void relu_bwd_f32_kernel(float* __restrict__ dy, const float* __restrict__ act, int64_t n4, float scale) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 d = reinterpret_cast<const float4*>(dy)[i];
  const float4 a = __ldcs(reinterpret_cast<const float4*>(act) + i);
  d.x = a.x > 0.f ? d.x * scale : 0.f; d.y = a.y > 0.f ? d.y * scale : 0.f;
  d.z = a.z > 0.f ? d.z * scale : 0.f; d.w = a.w > 0.f ? d.w * scale : 0.f;
  reinterpret_cast<float4*>(dy)[i] = d;
}

// ---- last head layer backward: logits[m] = a2[m,:] . w + b  (cls_head.7, models/MMCTransformer.py:71-80) ----------
// da2[m, c] = dlogit[m] * w[c] where a2[m, c] > 0 (the ReLU in front of the layer), bf16; per-block partials of
// dw[c] = sum_m dlogit[m] a2[m, c] and db = sum_m dlogit[m] (reduced in order by colsum_reduce_kernel).
__global__ void __launch_bounds__(256)
head_out_bwd_kernel(const float* __restrict__ dlogit, const __nv_bfloat16* __restrict__ a2, const float* __restrict__ w,
                    int64_t M, int64_t rows_per, __nv_bfloat16* __restrict__ da2, float* __restrict__ part, float scale) {
  const int c = threadIdx.x;  // 256 columns
  const float wc = w[c] * scale;  // (scale: backward of the Dropout between the ReLU and this layer; dw uses the stored a2)
  const int64_t r0 = int64_t(blockIdx.x) * rows_per;
  const int64_t r1 = r0 + rows_per < M ? r0 + rows_per : M;
  float dw = 0.f, db = 0.f;
  for (int64_t r = r0; r < r1; ++r) {
    const float g = dlogit[r];
    const float a = __bfloat162float(a2[r * 256 + c]);
    da2[r * 256 + c] = __float2bfloat16_rn(a > 0.f ? g * wc : 0.f);
    dw = fmaf(g, a, dw);
    db += g;
  }
  part[int64_t(blockIdx.x) * 257 + c] = dw;
  if (c == 0) part[int64_t(blockIdx.x) * 257 + 256] = db;
}

// ---- attention backward pre-process: Dsum[b, h, t] = sum_d dO[b,t,h,d] * O[b,t,h,d] (fp32) ----------------------
// one warp per token row of 512 bf16 (lane owns 16 consecutive columns = a quarter of head lane / 4)
__global__ void __launch_bounds__(256)
attn_bwd_dsum_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o, int B, int T, int H,
                     float* __restrict__ dsum) {
  const int lane = threadIdx.x & 31;
  const int64_t row = int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= int64_t(B) * T) return;
  const int cols = H * 64;
  for (int c0 = lane * 16; c0 < cols; c0 += 512) {
    float acc = 0.f;
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const uint4 a = *reinterpret_cast<const uint4*>(o + row * cols + c0 + 8 * v);
      const uint4 g = *reinterpret_cast<const uint4*>(d_o + row * cols + c0 + 8 * v);
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc = fmaf(__uint_as_float(aw[i] << 16), __uint_as_float(gw[i] << 16), acc);
        acc = fmaf(__uint_as_float(aw[i] & 0xffff0000u), __uint_as_float(gw[i] & 0xffff0000u), acc);
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if ((lane & 3) == 0) {
      const int head = c0 >> 6;
      const int64_t b = row / T, t = row - b * T;
      dsum[(b * H + head) * T + t] = acc;
    }
  }
}

// ---- dropout masks (dropout.cuh) -------------------------------------------------------------------------------------
// keep bits of an attention-dropout site: thread = one 32-bit word = 32 consecutive keys of one (batch, head, query) row
__global__ void __launch_bounds__(256)
attn_dropout_bits_kernel(uint32_t* __restrict__ bits, int64_t n_words, const DropKey drop) {
  const int64_t w = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  const uint32_t k0 = uint32_t(w) * 16u;
  uint32_t word = 0u;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint32_t h = drop_hash(k0 + uint32_t(i), drop.a, drop.b);
    word |= ((h & 0xffffu) >= drop.thr ? 1u : 0u) << (2 * i);
    word |= ((h >> 16) >= drop.thr ? 1u : 0u) << (2 * i + 1);
  }
  bits[w] = word;
}
__global__ void dropout_mask_u8_kernel(uint8_t* __restrict__ keep, int64_t n, const DropKey drop) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;  // pair index
  if (2 * k >= n) return;
  const uint32_t h = drop_hash(uint32_t(k), drop.a, drop.b);
  keep[2 * k] = (h & 0xffffu) >= drop.thr ? 1 : 0;
  if (2 * k + 1 < n) keep[2 * k + 1] = (h >> 16) >= drop.thr ? 1 : 0;
}
}  // namespace

int launch_attn_dropout_bits(uint32_t* bits, int64_t n_words, const DropKey& drop, cudaStream_t stream) {
  RP_CHECK(bits != nullptr && n_words > 0 && n_words * 16 < (int64_t(1) << 32),
           "attn_dropout_bits: 0 < 32 * n_words < 2^33 elements per site");
  attn_dropout_bits_kernel<<<unsigned((n_words + 255) / 256), 256, 0, stream>>>(bits, n_words, drop);
  count_launch();
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}
int launch_dropout_mask_u8(uint8_t* keep, int64_t n, const DropKey& drop, cudaStream_t stream) {
  RP_CHECK(keep != nullptr && n > 0 && n < (int64_t(1) << 32), "dropout_mask: 0 < n < 2^32");
  const int64_t pairs = (n + 1) / 2;
  dropout_mask_u8_kernel<<<unsigned((pairs + 255) / 256), 256, 0, stream>>>(keep, n, drop);
  count_launch();
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

int launch_focal_loss_grad(const float* logits, const float* targets, const uint8_t* mask, int64_t n, float alpha,
                           float gamma, float scale, float* dlogits, cudaStream_t stream) {
  RP_CHECK(n > 0, "focal_loss_grad: empty");
  focal_loss_grad_kernel<<<unsigned((n + 255) / 256), 256, 0, stream>>>(logits, targets, mask, n, alpha, gamma, scale,
                                                                        dlogits);
  count_launch();
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

int64_t layernorm512_bwd_scratch_floats() { return int64_t(3) * 2 * num_sms() * 512; }

int launch_layernorm512_bwd(const float* x, const float* dy, const float* gamma, int64_t M, float eps, float* dx,
                            float* dgamma, float* dbeta, float* scratch, cudaStream_t stream, bool accumulate,
                            void* dx_bf16, float* dx_colsum, const DropKey* drop) {
  RP_CHECK(M > 0, "layernorm512_bwd: empty");
  RP_CHECK(drop == nullptr || M * 512 < (int64_t(1) << 32), "layernorm512_bwd: dropout sites hold fewer than 2^32 elements");
  const int sms = num_sms();
  int grid = 2 * sms;  // two resident blocks per SM walk the rows
  if (int64_t(grid) * LNB_WARPS > M) grid = int((M + LNB_WARPS - 1) / LNB_WARPS);
  float* part_g = scratch;
  float* part_b = scratch + int64_t(2) * sms * 512;
  float* part_c = dx_colsum != nullptr ? scratch + int64_t(4) * sms * 512 : nullptr;
  layernorm512_bwd_kernel<<<grid, LNB_WARPS * 32, 0, stream>>>(x, dy, gamma, M, eps, dx, part_g, part_b,
                                                               accumulate ? 1 : 0,
                                                               reinterpret_cast<__nv_bfloat16*>(dx_bf16), part_c,
                                                               drop != nullptr ? *drop : DropKey{});
  layernorm512_bwd_reduce_kernel<<<4, 128, 0, stream>>>(part_g, part_b, part_c, grid, dgamma, dbeta, dx_colsum);
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


int launch_splitk_reduce(const float* part, int splits, int64_t n, float* out, cudaStream_t stream) {
  RP_CHECK(splits >= 1 && n > 0 && n % 4 == 0, "splitk_reduce: n must be a positive multiple of 4");
  splitk_reduce_kernel<<<unsigned((n / 4 + 255) / 256), 256, 0, stream>>>(part, splits, n, out);
  count_launch();
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

// one scratch size for every two-stage reduction of the backward pass (LayerNorm 2 x 4 SMs x 512, column sums
// <= 4 SMs x 256 per column block, head_out_bwd 4 SMs x 257 + 257)
int64_t train_scratch_floats() { return int64_t(4) * num_sms() * 1024 + 1024; }  // >= 3 x 2 SMs x 512 of the LayerNorm backward

int launch_colsum_bf16(const void* x, int64_t M, int N, float* out, float* scratch, cudaStream_t stream) {
  RP_CHECK(M > 0 && N > 0, "colsum: empty");
  const int bx = (N + 255) / 256;
  int by = 4 * num_sms() / bx;
  if (by < 1) by = 1;
  if (by > M) by = int(M);
  const int64_t rows_per = (M + by - 1) / by;
  by = int((M + rows_per - 1) / rows_per);
  colsum_bf16_kernel<<<dim3(bx, by), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), M, N, rows_per, scratch);
  colsum_reduce_kernel<<<(N + 127) / 128, 128, 0, stream>>>(scratch, by, N, out);
  count_launch(2);
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

int launch_relu_bwd(void* dy, const void* act, int64_t n, bool f32, float scale, cudaStream_t stream) {
  RP_CHECK(n > 0 && n % 8 == 0, "relu_bwd: n must be a positive multiple of 8");
  if (f32)
    relu_bwd_f32_kernel<<<unsigned((n / 4 + 255) / 256), 256, 0, stream>>>(static_cast<float*>(dy),
                                                                           static_cast<const float*>(act), n / 4, scale);
  else
    relu_bwd_bf16_kernel<<<unsigned((n / 8 + 255) / 256), 256, 0, stream>>>(
        static_cast<__nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(act), n / 8, scale);
  count_launch();
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

int launch_relu_bwd_colsum(void* dy, const void* act, int64_t M, int N, float scale, float* colsum, float* scratch,
                           cudaStream_t stream) {
  RP_CHECK(M > 0 && N > 0 && N % 8 == 0, "relu_bwd_colsum: N must be a positive multiple of 8");
  // scratch holds blocks x N partials: train_scratch_floats() >= 4 SMs x 1024 -> blocks <= that / N
  int64_t blocks = train_scratch_floats() / N;
  if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
  if (blocks > M) blocks = M;
  RP_CHECK(blocks >= 1, "relu_bwd_colsum: N too large for the scratch buffer");
  const int64_t rows_per = (M + blocks - 1) / blocks;
  blocks = (M + rows_per - 1) / rows_per;
  relu_bwd_colsum_bf16_kernel<<<unsigned(blocks), 256, 0, stream>>>(static_cast<__nv_bfloat16*>(dy),
                                                                    static_cast<const __nv_bfloat16*>(act), M, N, rows_per,
                                                                    scratch, scale);
  colsum_reduce_kernel<<<(N + 127) / 128, 128, 0, stream>>>(scratch, int(blocks), N, colsum);
  count_launch(2);
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

int launch_head_out_bwd(const float* dlogit, const void* a2, const float* w, int64_t M, float scale, void* da2, float* dw,
                        float* db, float* scratch, cudaStream_t stream) {
  RP_CHECK(M > 0, "head_out_bwd: empty");
  int blocks = 4 * num_sms();
  if (blocks > M) blocks = int(M);
  const int64_t rows_per = (M + blocks - 1) / blocks;
  blocks = int((M + rows_per - 1) / rows_per);
  head_out_bwd_kernel<<<blocks, 256, 0, stream>>>(dlogit, static_cast<const __nv_bfloat16*>(a2), w, M, rows_per,
                                                  static_cast<__nv_bfloat16*>(da2), scratch, scale);
  // the 257 partial columns: dw[0..256) and db
  colsum_reduce_kernel<<<3, 128, 0, stream>>>(scratch, blocks, 257, scratch + int64_t(blocks) * 257);
  RP_CUDA_CHECK(cudaMemcpyAsync(dw, scratch + int64_t(blocks) * 257, 256 * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  RP_CUDA_CHECK(cudaMemcpyAsync(db, scratch + int64_t(blocks) * 257 + 256, sizeof(float), cudaMemcpyDeviceToDevice, stream));
  count_launch(2);
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

int launch_attn_bwd_dsum(const void* o, const void* d_o, int B, int T, int H, float* dsum, cudaStream_t stream) {
  RP_CHECK(B > 0 && T > 0 && H > 0 && H % 8 == 0, "attn_bwd_dsum: H must be a positive multiple of 8");
  const int64_t rows = int64_t(B) * T;
  attn_bwd_dsum_kernel<<<unsigned((rows + 7) / 8), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(o),
                                                                     static_cast<const __nv_bfloat16*>(d_o), B, T, H, dsum);
  count_launch();
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

}  // namespace rp
