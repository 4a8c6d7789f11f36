// AtIoU on the device: precision of predicted segments at tIoU thresholds,
// averaged over videos and thresholds, computed from the gathered fixed-slot segment lists so that a
// 10K-video evaluation returns one scalar.  Restates calculate_tiou (utils/metrics.py:82-111) and the
// accumulation of inference.py:45-55 in float64 with the reference's operation order, so the result is
// bit-identical to the Python evaluation of the same (fp32) segments.
#include "ptx.cuh"
#include "host_util.h"
#include "kernels.h"

namespace rp {

namespace {

// one warp per video: lane p handles predicted segments p, p+32, ...
__global__ void __launch_bounds__(256)
tiou_precision_kernel(const float* __restrict__ slots, int n_videos, int K, const double* __restrict__ gt,
                      const int32_t* __restrict__ gt_counts, int Gmax, const double* __restrict__ thr,
                      int n_thr, double* __restrict__ prec) {
  const int lane = threadIdx.x & 31;
  const int vid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (vid >= n_videos) return;
  const float* row = slots + int64_t(vid) * (1 + 4 * K);
  int n_pred = int(row[0]);
  n_pred = n_pred < 0 ? 0 : (n_pred > K ? K : n_pred);
  const int n_gt = gt_counts[vid];
  const double* g = gt + int64_t(vid) * Gmax * 2;
  int hits[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int pi = lane; pi < n_pred; pi += 32) {
    const double p0 = double(row[1 + 4 * pi]), p1 = double(row[2 + 4 * pi]);
    double best = 0.0;  // max(..., default=0) over the reference segments
    bool any = false;
    for (int gi = 0; gi < n_gt; ++gi) {
      const double g0 = g[2 * gi], g1 = g[2 * gi + 1];
      const double start_max = p0 > g0 ? p0 : g0;
      const double end_min = p1 < g1 ? p1 : g1;
      const double d = end_min - start_max;
      const double inter = d > 0.0 ? d : 0.0;
      const double uni = __dsub_rn(__dadd_rn(p1 - p0, g1 - g0), inter);
      const double iou = uni != 0.0 ? inter / uni : 0.0;
      if (!any || iou > best) best = iou;
      any = true;
    }
    for (int t = 0; t < n_thr && t < 8; ++t) hits[t] += (best >= thr[t]) ? 1 : 0;
  }
  for (int t = 0; t < n_thr && t < 8; ++t) {
    int h = hits[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    if (lane == 0) prec[int64_t(vid) * n_thr + t] = n_pred > 0 ? double(h) / double(n_pred) : 0.0;
  }
}

// Python's sum() over floats (CPython >= 3.12: Neumaier compensated summation, video order).
struct PySum {
  double f = 0.0, c = 0.0;
  __device__ void add(double x) {
    const double t = __dadd_rn(f, x);
    if (fabs(f) >= fabs(x)) c = __dadd_rn(c, __dadd_rn(__dsub_rn(f, t), x));
    else c = __dadd_rn(c, __dadd_rn(__dsub_rn(x, t), f));
    f = t;
  }
  __device__ double result() const { return (c != 0.0 && isfinite(c)) ? __dadd_rn(f, c) : f; }
};

// out[t] = sum_v(prec[v,t]) / n_videos ; out[n_thr] = sum_t(out[t]) / n_thr   (inference.py:49-53)
__global__ void atiou_reduce_kernel(const double* __restrict__ prec, int n_videos, int n_thr,
                                    double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  PySum total;
  for (int t = 0; t < n_thr; ++t) {
    PySum s;
    for (int v = 0; v < n_videos; ++v) s.add(prec[int64_t(v) * n_thr + t]);
    const double m = s.result() / double(n_videos);
    out[t] = m;
    total.add(m);
  }
  out[n_thr] = total.result() / double(n_thr);
}

}  // namespace

int launch_atiou(const float* slots, int n_videos, int K, const double* gt, const int32_t* gt_counts,
                 int Gmax, const double* thresholds, int n_thr, double* per_video, double* out,
                 cudaStream_t stream) {
  RP_CHECK(n_videos > 0 && K >= 1 && Gmax >= 1, "atiou: empty problem");
  RP_CHECK(n_thr >= 1 && n_thr <= 8, "atiou: 1..8 thresholds supported");
  tiou_precision_kernel<<<(n_videos + 7) / 8, 256, 0, stream>>>(slots, n_videos, K, gt, gt_counts, Gmax,
                                                                 thresholds, n_thr, per_video);
  count_launch();
  atiou_reduce_kernel<<<1, 32, 0, stream>>>(per_video, n_videos, n_thr, out);
  count_launch();
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

}  // namespace rp

// ---------------------------------------------------------------------------------------------
// Masked sigmoid focal loss, summed (forward value of MMCTransformer.losses, models/MMCTransformer.py:
// 159-179 with sigmoid_focal_loss of models/losses.py:5-53): fp32 element math in the reference's
// operation order, deterministic two-stage reduction (per-block partials in float64, then one block
// adds them in index order), so the same inputs always give the same bits.
namespace rp {
namespace {

constexpr int FL_BLOCKS = 296;   // two CTAs per SM
constexpr int FL_THREADS = 256;

__device__ __forceinline__ float focal_term(float x, float t, float alpha, float gamma) {
  const float p = 1.0f / (1.0f + expf(-x));  // torch.sigmoid
  // binary_cross_entropy_with_logits: max(x, 0) - x t + log1p(exp(-|x|))
  const float ce = __fadd_rn(__fsub_rn(fmaxf(x, 0.0f), __fmul_rn(x, t)), log1pf(expf(-fabsf(x))));
  const float p_t = __fadd_rn(__fmul_rn(p, t), __fmul_rn(1.0f - p, 1.0f - t));
  const float om = 1.0f - p_t;
  const float mod = gamma == 2.0f ? __fmul_rn(om, om) : powf(om, gamma);
  float loss = __fmul_rn(ce, mod);
  if (alpha >= 0.0f) {
    const float alpha_t = __fadd_rn(__fmul_rn(alpha, t), __fmul_rn(1.0f - alpha, 1.0f - t));
    loss = __fmul_rn(alpha_t, loss);
  }
  return loss;
}

__global__ void __launch_bounds__(FL_THREADS)
focal_partial_kernel(const float* __restrict__ logits, const float* __restrict__ targets,
                     const uint8_t* __restrict__ mask, int64_t n, float alpha, float gamma,
                     double* __restrict__ partial) {
  __shared__ double red[FL_THREADS / 32];
  double acc = 0.0;
  for (int64_t i = blockIdx.x * int64_t(FL_THREADS) + threadIdx.x; i < n; i += int64_t(gridDim.x) * FL_THREADS) {
    if (mask[i]) acc += double(focal_term(logits[i], targets[i], alpha, gamma));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < FL_THREADS / 32; ++w) s += red[w];
    partial[blockIdx.x] = s;
  }
}

__global__ void focal_final_kernel(const double* __restrict__ partial, int n_partial, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n_partial; ++i) s += partial[i];
    out[0] = float(s);
  }
}

}  // namespace

int launch_focal_loss_sum(const float* logits, const float* targets, const uint8_t* mask, int64_t n, float alpha,
                          float gamma, double* scratch, float* out, cudaStream_t stream) {
  RP_CHECK(n > 0, "focal_loss: empty input");
  focal_partial_kernel<<<FL_BLOCKS, FL_THREADS, 0, stream>>>(logits, targets, mask, n, alpha, gamma, scratch);
  focal_final_kernel<<<1, 32, 0, stream>>>(scratch, FL_BLOCKS, out);
  count_launch(2);
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

}  // namespace rp
