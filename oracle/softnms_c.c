/* ORACLE (test infrastructure, see oracle/__init__.py): plain-C restatement of the reference's
 * Gaussian Soft-NMS, soft_nms_intervals_cpu (models/softnms.py:3-38), in scalar IEEE float32 with
 * the reference's operation order.  Differs from the NumPy restatement
 * only in the exp implementation (glibc expf vs NumPy's SIMD exp), which is what makes it a useful
 * second witness for the CUDA kernel's expf.  Built with -ffp-contract=off (no FMA contraction).
 *
 * Returns the number of kept rows; keep[] receives original indices in selection order, kscores[]
 * their decayed scores.  Inputs are not modified. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

int rp_oracle_soft_nms(const float* scores, const float* segs, int n, float sigma, float thresh,
                       int max_seg_num, long long* keep, float* kscores) {
  if (n <= 0) return 0;
  float* sc = (float*)malloc(sizeof(float) * n);
  float* begin = (float*)malloc(sizeof(float) * n);
  float* end = (float*)malloc(sizeof(float) * n);
  float* len0 = (float*)malloc(sizeof(float) * n);
  long long* orig = (long long*)malloc(sizeof(long long) * n);
  for (int j = 0; j < n; ++j) {
    sc[j] = scores[j];
    begin[j] = segs[2 * j];
    end[j] = segs[2 * j + 1];
    len0[j] = end[j] - begin[j];   /* models/softnms.py:13 — by position, never permuted */
    orig[j] = j;
  }
  const int m = max_seg_num < n ? max_seg_num : n;
  int cnt = 0;
  for (int i = 0; i < n; ++i) {
    const float tscore = sc[i];    /* :18 pre-swap */
    if (i != n - 1) {
      int mp = i + 1;              /* :21-22 first maximum of the tail */
      for (int j = i + 2; j < n; ++j)
        if (sc[j] > sc[mp]) mp = j;
      if (tscore < sc[mp]) {       /* :23-25 */
        float t;
        long long o;
        t = begin[i]; begin[i] = begin[mp]; begin[mp] = t;
        t = end[i]; end[i] = end[mp]; end[mp] = t;
        t = sc[i]; sc[i] = sc[mp]; sc[mp] = t;
        o = orig[i]; orig[i] = orig[mp]; orig[mp] = o;
      }
    }
    if (tscore > thresh) {         /* :26-29 */
      ++cnt;
      if (cnt >= m) break;
    }
    for (int j = i + 1; j < n; ++j) {   /* :30-36 */
      float ov = fminf(end[i], end[j]) - fmaxf(begin[i], begin[j]);
      ov = ov > 0.0f ? ov : 0.0f;
      const float total = (len0[i] + len0[j]) - ov;
      const float r = ov / total;
      const float w = expf(-(r * r) / sigma);
      sc[j] = w * sc[j];
    }
  }
  int k = 0;
  for (int j = 0; j < n && k < m; ++j)  /* :37 */
    if (sc[j] > thresh) {
      keep[k] = orig[j];
      kscores[k] = sc[j];
      ++k;
    }
  free(sc); free(begin); free(end); free(len0); free(orig);
  return k;
}
