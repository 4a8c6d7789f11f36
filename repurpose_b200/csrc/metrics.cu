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
