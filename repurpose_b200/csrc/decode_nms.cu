// Per-video segment decode + Gaussian Soft-NMS, one CTA per video, all candidate state in shared
// memory, no host round trip.
//
// decode  : restates MMCTransformer.inference_single_video (models/MMCTransformer.py:181-229):
//           p = sigmoid(logit) * mask ; keep p > pre_nms_thresh ; sort descending ; first
//           pre_nms_topk ; seg = (t - off_l, t + off_r) ; keep duration_thresh < dur < duration_thresh_max
// Soft-NMS: restates soft_nms_intervals_cpu (models/softnms.py:3-38) including its quirks
//           (SURVEY.md Appendix B): pre-swap tscore drives the counter, `lengths` is indexed by
//           position and never permuted, break happens before the decay of that round, final keep
//           = rows with score > thresh in permuted order, first max_segments of them.
// All Soft-NMS arithmetic is fp32 with explicit round-to-nearest intrinsics in NumPy's operation
// order (no FMA contraction).
#include <math.h>

#include "ptx.cuh"
#include "host_util.h"
#include "kernels.h"

namespace rp {

namespace {

constexpr int NT = 1024;          // threads per CTA
constexpr int MAX_T = 8192;       // decode: max feature steps per video
constexpr int MAX_TOPK = 4096;    // decode: max pre-NMS candidates
constexpr int MAX_NMS_N = 8192;   // standalone Soft-NMS: max candidates

__device__ __forceinline__ uint32_t float_orderable(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int o) {
  return __shfl_xor_sync(0xffffffffu, v, o);
}

// block-wide max of 64-bit keys; result broadcast to every thread.  `red` = 32 x u64 scratch.
__device__ __forceinline__ unsigned long long block_max_u64(unsigned long long v,
                                                            unsigned long long* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = shfl_xor_u64(v, o);
    v = t > v ? t : v;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // protect `red` from the previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  unsigned long long r = red[lane];  // NT/32 == 32 warps
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = shfl_xor_u64(r, o);
    r = t > r ? t : r;
  }
  return r;
}

// Ordered compaction support: exclusive prefix sum over per-thread counts (1024 threads).
// Returns this thread's exclusive offset; *total receives the block total.  `red` = 33 ints.
__device__ __forceinline__ int block_excl_scan(int v, int* red, int* total) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) red[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = red[lane];
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    red[lane] = winc - w;  // exclusive warp offsets
    if (lane == 31) red[32] = winc;
  }
  __syncthreads();
  *total = red[32];
  return red[warp] + inc - v;
}

struct NmsState {
  float* begin;
  float* end;
  float* sc;
  float* len0;
  int* oidx;
};

// Soft-NMS over N candidates already resident in shared memory.  On return *out_count rows have
// been written to keep_idx / keep_score (original candidate index, decayed score) in selection
// order.  Every thread of the CTA must call this.
__device__ void soft_nms_block(const NmsState s, int N, int max_seg, float sigma, float thresh,
                               int Kcap, int32_t* __restrict__ keep_idx,
                               float* __restrict__ keep_score, int32_t* __restrict__ out_count,
                               unsigned long long* red64, int* red32) {
  const int tid = threadIdx.x;
  int M = max_seg < N ? max_seg : N;
  if (M > Kcap) M = Kcap;
  if (M <= 0 || N <= 0) {
    if (tid == 0) *out_count = 0;
    return;
  }
  for (int j = tid; j < N; j += NT) {
    s.len0[j] = __fsub_rn(s.end[j], s.begin[j]);
    s.oidx[j] = j;
  }
  __syncthreads();

  // running best over the tail [i+1, N): key = (orderable score << 32) | ~index  (first index wins ties)
  unsigned long long best = 0ull;
  for (int j = 1 + tid; j < N; j += NT) {
    const unsigned long long key =
        (static_cast<unsigned long long>(float_orderable(s.sc[j])) << 32) | (0xffffffffu - uint32_t(j));
    best = key > best ? key : best;
  }
  int cnt = 0;
  for (int i = 0; i < N; ++i) {
    const float tscore = s.sc[i];  // pre-swap value, read by every thread (smem broadcast)
    if (i != N - 1) {
      const unsigned long long top = block_max_u64(best, red64);  // syncs inside
      const int mp = int(0xffffffffu - uint32_t(top & 0xffffffffull));
      const float maxscore = s.sc[mp];
      __syncthreads();  // every thread has read sc[i] / sc[mp] before the swap
      if (tscore < maxscore && tid == 0) {
        float t;
        t = s.begin[i]; s.begin[i] = s.begin[mp]; s.begin[mp] = t;
        t = s.end[i]; s.end[i] = s.end[mp]; s.end[mp] = t;
        t = s.sc[i]; s.sc[i] = s.sc[mp]; s.sc[mp] = t;
        const int o = s.oidx[i]; s.oidx[i] = s.oidx[mp]; s.oidx[mp] = o;
      }
      __syncthreads();
    }
    if (tscore > thresh) {
      ++cnt;
      if (cnt >= M) break;  // uniform: every thread evaluates the same condition
    }
    // decay the tail against row i, and collect the arg-max of [i+2, N) for the next round
    const float bi = s.begin[i], ei = s.end[i], li = s.len0[i];
    best = 0ull;
    for (int j = i + 1 + tid; j < N; j += NT) {
      const float ov = fmaxf(__fsub_rn(fminf(ei, s.end[j]), fmaxf(bi, s.begin[j])), 0.0f);
      const float total = __fsub_rn(__fadd_rn(li, s.len0[j]), ov);
      const float r = __fdiv_rn(ov, total);
      const float w = expf(__fdiv_rn(-__fmul_rn(r, r), sigma));
      const float ns = __fmul_rn(w, s.sc[j]);
      s.sc[j] = ns;
      if (j >= i + 2) {
        const unsigned long long key =
            (static_cast<unsigned long long>(float_orderable(ns)) << 32) | (0xffffffffu - uint32_t(j));
        best = key > best ? key : best;
      }
    }
    __syncthreads();
  }
  __syncthreads();

  // keep = rows with score > thresh, permuted order, first M
  const int per = (N + NT - 1) / NT;
  const int j0 = tid * per;
  int local = 0;
  for (int j = j0; j < j0 + per && j < N; ++j) local += (s.sc[j] > thresh) ? 1 : 0;
  int total;
  int pos = block_excl_scan(local, red32, &total);
  for (int j = j0; j < j0 + per && j < N; ++j) {
    if (s.sc[j] > thresh) {
      if (pos < M) {
        keep_idx[pos] = s.oidx[j];
        keep_score[pos] = s.sc[j];
      }
      ++pos;
    }
  }
  if (tid == 0) *out_count = total < M ? total : M;
}

// ------------------------------------------------------------------------------------------
// standalone Soft-NMS on caller-provided candidates
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1)
soft_nms_kernel(const float* __restrict__ scores, const float* __restrict__ segs,
                const int32_t* __restrict__ n, const int32_t* __restrict__ max_seg, int Nmax,
                float sigma, float thresh, int Kcap, int32_t* __restrict__ keep,
                float* __restrict__ kscores, int32_t* __restrict__ counts) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ unsigned long long red64[32];
  __shared__ int red32[33];
  const int b = blockIdx.x;
  int N = n[b];
  N = N < 0 ? 0 : (N > Nmax ? Nmax : N);
  NmsState s;
  s.begin = reinterpret_cast<float*>(smem_raw);
  s.end = s.begin + Nmax;
  s.sc = s.end + Nmax;
  s.len0 = s.sc + Nmax;
  s.oidx = reinterpret_cast<int*>(s.len0 + Nmax);
  for (int j = threadIdx.x; j < N; j += NT) {
    const float2 se = *reinterpret_cast<const float2*>(segs + (int64_t(b) * Nmax + j) * 2);
    s.begin[j] = se.x;
    s.end[j] = se.y;
    s.sc[j] = scores[int64_t(b) * Nmax + j];
  }
  __syncthreads();
  soft_nms_block(s, N, max_seg[b], sigma, thresh, Kcap, keep + int64_t(b) * Kcap,
                 kscores + int64_t(b) * Kcap, counts + b, red64, red32);
}

// ------------------------------------------------------------------------------------------
// decode + Soft-NMS
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1)
decode_nms_kernel(const float* __restrict__ logits, const float* __restrict__ offsets,
                  const int32_t* __restrict__ lens, const int32_t* __restrict__ max_seg, int T,
                  const DecodeCfg cfg, int Kcap, float* __restrict__ out_segs,
                  float* __restrict__ out_scores, float* __restrict__ out_dscores,
                  int32_t* __restrict__ out_labels, int32_t* __restrict__ out_counts,
                  int32_t* __restrict__ out_ncand, float* __restrict__ cand_segs,
                  float* __restrict__ cand_scores, int32_t* __restrict__ cand_labels, int sort_cap,
                  int cand_cap) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ unsigned long long red64[32];
  __shared__ int red32[33];
  __shared__ int s_ncand;
  __shared__ int32_t s_keep_idx[64];
  __shared__ float s_keep_score[64];
  __shared__ int32_t s_count;

  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
  NmsState s;
  s.begin = reinterpret_cast<float*>(smem_raw + size_t(sort_cap) * 8);
  s.end = s.begin + cand_cap;
  s.sc = s.end + cand_cap;
  s.len0 = s.sc + cand_cap;
  s.oidx = reinterpret_cast<int*>(s.len0 + cand_cap);
  float* prob0 = reinterpret_cast<float*>(s.oidx + cand_cap);  // original probability per candidate
  int* label = reinterpret_cast<int*>(prob0 + cand_cap);       // centre step t per candidate

  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  int len = lens[b];
  len = len < 0 ? 0 : (len > T ? T : len);
  if (tid == 0) s_ncand = 0;
  __syncthreads();

  // 1. probabilities, threshold, unordered compaction of (p, t) sort keys
  const float* lg = logits + int64_t(b) * T;
  for (int t = tid; t < T; t += NT) {
    const float sig = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-lg[t])));
    const float pr = (t < len) ? sig : 0.0f;  // sigmoid * mask
    if (pr > cfg.pre_nms_thresh) {
      const int slot = atomicAdd(&s_ncand, 1);
      keys[slot] = (static_cast<unsigned long long>(float_orderable(pr)) << 32) |
                   (0xffffffffu - uint32_t(t));
    }
  }
  __syncthreads();
  const int n_pass = s_ncand;
  int P2 = 1;
  while (P2 < n_pass) P2 <<= 1;
  for (int i = n_pass + tid; i < P2; i += NT) keys[i] = 0ull;
  __syncthreads();

  // 2. bitonic sort, descending by (p, then ascending t)
  for (int k = 2; k <= P2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P2; i += NT) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], c = keys[ixj];
          const bool desc = (i & k) == 0;
          if (desc ? (a < c) : (a > c)) {
            keys[i] = c;
            keys[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }

  // 3. top-k, segments, duration filter, ordered compaction into the candidate arrays
  int ntop = n_pass < cfg.pre_nms_topk ? n_pass : cfg.pre_nms_topk;
  if (ntop > cand_cap) ntop = cand_cap;
  const int per = (ntop + NT - 1) / NT;
  const int i0 = tid * per;
  const float* off = offsets + int64_t(b) * T * 2;
  int local = 0;
  for (int i = i0; i < i0 + per && i < ntop; ++i) {
    const int t = int(0xffffffffu - uint32_t(keys[i] & 0xffffffffull));
    const float2 o = *reinterpret_cast<const float2*>(off + 2 * t);
    const float sl = __fsub_rn(float(t), o.x);
    const float sr = __fadd_rn(float(t), o.y);
    const float dur = __fsub_rn(sr, sl);
    local += (dur > cfg.duration_thresh && dur < cfg.duration_thresh_max) ? 1 : 0;
  }
  int N;
  int pos = block_excl_scan(local, red32, &N);
  for (int i = i0; i < i0 + per && i < ntop; ++i) {
    const unsigned long long key = keys[i];
    const int t = int(0xffffffffu - uint32_t(key & 0xffffffffull));
    const float2 o = *reinterpret_cast<const float2*>(off + 2 * t);
    const float sl = __fsub_rn(float(t), o.x);
    const float sr = __fadd_rn(float(t), o.y);
    const float dur = __fsub_rn(sr, sl);
    if (dur > cfg.duration_thresh && dur < cfg.duration_thresh_max) {
      const float sig = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-lg[t])));
      s.begin[pos] = sl;
      s.end[pos] = sr;
      s.sc[pos] = sig;
      prob0[pos] = sig;
      label[pos] = t;
      ++pos;
    }
  }
  __syncthreads();
  if (tid == 0) out_ncand[b] = N;
  if (cand_segs != nullptr) {  // optional: the pre-NMS candidate list (inference_single_video output)
    for (int i = tid; i < N; i += NT) {
      const int64_t o = int64_t(b) * cand_cap + i;
      cand_segs[2 * o] = s.begin[i];
      cand_segs[2 * o + 1] = s.end[i];
      cand_scores[o] = prob0[i];
      cand_labels[o] = label[i];
    }
    __syncthreads();
  }

  // 4. Soft-NMS
  int kc = Kcap < 64 ? Kcap : 64;
  soft_nms_block(s, N, max_seg[b], cfg.nms_sigma, cfg.min_score, kc, s_keep_idx, s_keep_score,
                 &s_count, red64, red32);
  __syncthreads();
  const int count = s_count;
  if (tid < Kcap) {
    const int64_t o = int64_t(b) * Kcap + tid;
    if (tid < count) {
      const int ci = s_keep_idx[tid];
      // begin/end were permuted by the NMS; recompute the segment from the candidate's centre
      const int t = label[ci];
      const float2 of = *reinterpret_cast<const float2*>(off + 2 * t);
      out_segs[2 * o] = __fsub_rn(float(t), of.x);
      out_segs[2 * o + 1] = __fadd_rn(float(t), of.y);
      out_scores[o] = prob0[ci];
      out_dscores[o] = s_keep_score[tid];
      out_labels[o] = t;
    } else {
      out_segs[2 * o] = 0.f;
      out_segs[2 * o + 1] = 0.f;
      out_scores[o] = 0.f;
      out_dscores[o] = 0.f;
      out_labels[o] = -1;
    }
  }
  if (tid == 0) out_counts[b] = count;
}

}  // namespace

int launch_soft_nms(const float* scores, const float* segs, const int32_t* n, const int32_t* max_seg,
                    int B, int Nmax, float sigma, float thresh, int Kcap, int32_t* keep,
                    float* kscores, int32_t* counts, cudaStream_t stream) {
  RP_CHECK(B > 0, "soft_nms: empty batch");
  RP_CHECK(Nmax >= 1 && Nmax <= MAX_NMS_N, "soft_nms: Nmax=%d out of range [1,%d]", Nmax, MAX_NMS_N);
  RP_CHECK(Kcap >= 1, "soft_nms: Kcap must be >= 1");
  const size_t smem = size_t(Nmax) * 5 * 4;
  static size_t configured_on[kMaxDevices];
  size_t& configured = configured_on[current_device()];
  if (smem > configured) {
    RP_CUDA_CHECK(cudaFuncSetAttribute(soft_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       int(smem)));
    configured = smem;
  }
  soft_nms_kernel<<<B, NT, smem, stream>>>(scores, segs, n, max_seg, Nmax, sigma, thresh, Kcap, keep,
                                           kscores, counts);
  count_launch();
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

int launch_decode_nms(const float* logits, const float* offsets, const int32_t* lens,
                      const int32_t* max_seg, int B, int T, const DecodeCfg& cfg, int Kcap,
                      float* segs, float* scores, float* dscores, int32_t* labels, int32_t* counts,
                      int32_t* ncand, float* cand_segs, float* cand_scores, int32_t* cand_labels,
                      cudaStream_t stream) {
  RP_CHECK(B > 0 && T > 0, "decode_nms: empty batch");
  RP_CHECK(T <= MAX_T, "decode_nms: T=%d exceeds %d", T, MAX_T);
  RP_CHECK(cfg.pre_nms_topk >= 1 && cfg.pre_nms_topk <= MAX_TOPK,
           "decode_nms: pre_nms_topk=%d out of range [1,%d]", cfg.pre_nms_topk, MAX_TOPK);
  RP_CHECK(Kcap >= 1 && Kcap <= 64, "decode_nms: Kcap=%d out of range [1,64]", Kcap);
  RP_CHECK((cand_segs == nullptr) == (cand_scores == nullptr) && (cand_segs == nullptr) == (cand_labels == nullptr),
           "decode_nms: candidate outputs must be all set or all null");
  int sort_cap = 1;
  while (sort_cap < T) sort_cap <<= 1;
  const int cand_cap = cfg.pre_nms_topk < T ? cfg.pre_nms_topk : T;
  const size_t smem = size_t(sort_cap) * 8 + size_t(cand_cap) * 7 * 4;
  static size_t configured_on[kMaxDevices];
  size_t& configured = configured_on[current_device()];
  if (smem > configured) {
    RP_CUDA_CHECK(cudaFuncSetAttribute(decode_nms_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    configured = smem;
  }
  decode_nms_kernel<<<B, NT, smem, stream>>>(logits, offsets, lens, max_seg, T, cfg, Kcap, segs,
                                             scores, dscores, labels, counts, ncand, cand_segs,
                                             cand_scores, cand_labels, sort_cap, cand_cap);
  count_launch();
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

}  // namespace rp
