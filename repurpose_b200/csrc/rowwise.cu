// Memory-bound row-wise kernels of the hot path: modality concat + bf16 cast, the LayerNorm family
// (one warp per 512-wide row, row held in registers, 128-bit loads/stores), and the final
// 256->1 / 256->2 head projections.  Reference: models/MMCTransformer.py:118 (cat), :124/:127
// (input_norm, positional_encoding), :58/:141 (encoder_norm), :63-68 (feature_map LN+ReLU),
// :71-93 (head LayerNorm and last Linear layers); nn.TransformerEncoderLayer norm1/norm2.
#include <cuda_bf16.h>

#include "ptx.cuh"
#include "host_util.h"
#include "kernels.h"

namespace rp {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ uint2 pack4_bf16(float a, float b, float c, float d) {
  uint2 r;
  r.x = pack_bf16x2(a, b);
  r.y = pack_bf16x2(c, d);
  return r;
}

// ------------------------------------------------------------------------------------------
// concat + cast: out[m, :] = bf16(cat(vis[m], aud[m], txt[m]))
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
concat_cast_kernel(const float* __restrict__ vis, const float* __restrict__ aud,
                   const float* __restrict__ txt, int Cv, int Ca, int Ct,
                   __nv_bfloat16* __restrict__ out, int64_t M) {
  const int C = Cv + Ca + Ct;
  const int groups_per_row = C >> 3;  // 8 elements (32 B in, 16 B out) per thread-iteration
  const int64_t total = M * groups_per_row;
  for (int64_t g = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; g < total;
       g += int64_t(gridDim.x) * blockDim.x) {
    const int64_t m = g / groups_per_row;
    const int c = int(g - m * groups_per_row) << 3;
    const float* src;
    if (c < Cv) src = vis + m * Cv + c;
    else if (c < Cv + Ca) src = aud + m * Ca + (c - Cv);
    else src = txt + m * Ct + (c - Cv - Ca);
    const float4 a = __ldcs(reinterpret_cast<const float4*>(src));
    const float4 b = __ldcs(reinterpret_cast<const float4*>(src) + 1);
    uint4 o;
    o.x = pack_bf16x2(a.x, a.y);
    o.y = pack_bf16x2(a.z, a.w);
    o.z = pack_bf16x2(b.x, b.y);
    o.w = pack_bf16x2(b.z, b.w);
    *reinterpret_cast<uint4*>(out + m * C + c) = o;
  }
}

template <typename IN>
__global__ void __launch_bounds__(256)
ragged_concat_cast_kernel(const IN* __restrict__ vis, const IN* __restrict__ aud,
                          const IN* __restrict__ txt, int Cv, int Ca, int Ct,
                          const int32_t* __restrict__ row_off, const int32_t* __restrict__ txt_off,
                          const int32_t* __restrict__ txt_lens, const int32_t* __restrict__ lens, int B,
                          int T, __nv_bfloat16* __restrict__ out) {
  const int C = Cv + Ca + Ct;
  const int groups_per_row = C >> 3;
  const int64_t total = int64_t(B) * T * groups_per_row;
  for (int64_t g = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; g < total;
       g += int64_t(gridDim.x) * blockDim.x) {
    const int64_t m = g / groups_per_row;         // padded row index b*T + t
    const int c = int(g - m * groups_per_row) << 3;
    const int b = int(m / T);
    const int t = int(m - int64_t(b) * T);
    const IN* src = nullptr;
    if (t < lens[b]) {
      if (c < Cv) src = vis + (int64_t(row_off[b]) + t) * Cv + c;
      else if (c < Cv + Ca) src = aud + (int64_t(row_off[b]) + t) * Ca + (c - Cv);
      else if (t < txt_lens[b]) src = txt + (int64_t(txt_off[b]) + t) * Ct + (c - Cv - Ca);
    }
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (src != nullptr) {
      if (sizeof(IN) == 4) {
        const float4 a = __ldcs(reinterpret_cast<const float4*>(src));
        const float4 bq = __ldcs(reinterpret_cast<const float4*>(src) + 1);
        o.x = pack_bf16x2(a.x, a.y);
        o.y = pack_bf16x2(a.z, a.w);
        o.z = pack_bf16x2(bq.x, bq.y);
        o.w = pack_bf16x2(bq.z, bq.w);
      } else {
        o = __ldcs(reinterpret_cast<const uint4*>(src));  // features already stored as bf16
      }
    }
    *reinterpret_cast<uint4*>(out + m * C + c) = o;
  }
}

__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t n8) {
  for (int64_t g = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; g < n8;
       g += int64_t(gridDim.x) * blockDim.x) {
    const float4 a = __ldcs(reinterpret_cast<const float4*>(in) + 2 * g);
    const float4 b = __ldcs(reinterpret_cast<const float4*>(in) + 2 * g + 1);
    uint4 o;
    o.x = pack_bf16x2(a.x, a.y);
    o.y = pack_bf16x2(a.z, a.w);
    o.z = pack_bf16x2(b.x, b.y);
    o.w = pack_bf16x2(b.z, b.w);
    reinterpret_cast<uint4*>(out)[g] = o;
  }
}

// ------------------------------------------------------------------------------------------
// LayerNorm over 512 columns.  Lane l owns columns {128*i + 4*l .. +3}, i = 0..3, so every
// warp-wide access is a contiguous 512-byte (fp32) or 256-byte (bf16) segment.
// Two-pass statistics in registers (mean, then centred sum of squares) like ATen's CPU kernel.
// ------------------------------------------------------------------------------------------
struct Row {
  float v[16];
};

__device__ __forceinline__ void row_load(Row& r, const float* __restrict__ p, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = *reinterpret_cast<const float4*>(p + 128 * i + 4 * lane);
    r.v[4 * i] = t.x; r.v[4 * i + 1] = t.y; r.v[4 * i + 2] = t.z; r.v[4 * i + 3] = t.w;
  }
}
__device__ __forceinline__ void row_store_f32(const Row& r, float* __restrict__ p, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(p + 128 * i + 4 * lane) =
        make_float4(r.v[4 * i], r.v[4 * i + 1], r.v[4 * i + 2], r.v[4 * i + 3]);
}
__device__ __forceinline__ void row_store_bf16(const Row& r, __nv_bfloat16* __restrict__ p, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<uint2*>(p + 128 * i + 4 * lane) =
        pack4_bf16(r.v[4 * i], r.v[4 * i + 1], r.v[4 * i + 2], r.v[4 * i + 3]);
}
__device__ __forceinline__ void row_stats(const Row& in, float& mean, float& rstd, float eps) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += in.v[i];
  mean = warp_sum(s) * (1.0f / 512.0f);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float d = in.v[i] - mean;
    ss += d * d;
  }
  const float var = warp_sum(ss) * (1.0f / 512.0f);
  rstd = rsqrtf(var + eps);
}
__device__ __forceinline__ void row_apply(Row& out, const Row& in, float mean, float rstd,
                                          const float* __restrict__ g, const float* __restrict__ b, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(g + 128 * i + 4 * lane));
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(b + 128 * i + 4 * lane));
    out.v[4 * i + 0] = (in.v[4 * i + 0] - mean) * rstd * g4.x + b4.x;
    out.v[4 * i + 1] = (in.v[4 * i + 1] - mean) * rstd * g4.y + b4.y;
    out.v[4 * i + 2] = (in.v[4 * i + 2] - mean) * rstd * g4.z + b4.z;
    out.v[4 * i + 3] = (in.v[4 * i + 3] - mean) * rstd * g4.w + b4.w;
  }
}
__device__ __forceinline__ void row_norm(Row& out, const Row& in, const float* __restrict__ g,
                                         const float* __restrict__ b, float eps, int lane) {
  float mean, rstd;
  row_stats(in, mean, rstd, eps);
  row_apply(out, in, mean, rstd, g, b, lane);
}

template <int MODE>
__global__ void __launch_bounds__(256, 2)
layernorm512_kernel(const LnArgs a) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = int64_t(gridDim.x) * (blockDim.x >> 5);
  // Rows are visited from the END of the buffer: the producer GEMM wrote its last tiles most
  // recently, so those rows are still L2-resident, and this kernel in turn finishes on row 0,
  // where the consumer GEMM starts.
  // Two rows per warp and pass: both rows' loads are issued before either row's reduction chain starts.
  auto finish_row = [&](const Row& x, int64_t row) {
    Row y;
    row_norm(y, x, a.g0, a.b0, a.eps, lane);
    if constexpr (MODE == 0) {
      row_store_bf16(y, reinterpret_cast<__nv_bfloat16*>(a.y_bf16) + row * 512, lane);
    } else if constexpr (MODE == 3) {
      row_store_f32(y, a.out_f32 + row * 512, lane);
    } else if constexpr (MODE == 1) {
      Row pe;
      row_load(pe, a.pe + int64_t(row % a.T) * 512, lane);
#pragma unroll
      for (int i = 0; i < 16; ++i) y.v[i] += pe.v[i];
      row_store_f32(y, a.out_f32 + row * 512, lane);
      Row u;
      row_norm(u, y, a.g1, a.b1, a.eps, lane);
      row_store_bf16(u, reinterpret_cast<__nv_bfloat16*>(a.y_bf16) + row * 512, lane);
    } else {  // MODE 2
#pragma unroll
      for (int i = 0; i < 16; ++i) y.v[i] = fmaxf(y.v[i], 0.0f);
      if (a.drop.thr != 0u) {  // train mode: feature_map[3] = nn.Dropout on relu(LN(.)); the heads see the dropped row
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t k = (uint32_t(row) * 512u + uint32_t(128 * i + 4 * lane)) >> 1;
          drop_pair(k, a.drop, y.v[4 * i], y.v[4 * i + 1]);
          drop_pair(k + 1u, a.drop, y.v[4 * i + 2], y.v[4 * i + 3]);
        }
      }
      row_store_f32(y, a.out_f32 + row * 512, lane);
      // both head LayerNorms normalise the same row: one set of statistics, two affine maps
      float mean, rstd;
      row_stats(y, mean, rstd, a.eps);
      Row u;
      row_apply(u, y, mean, rstd, a.g1, a.b1, lane);
      row_store_bf16(u, reinterpret_cast<__nv_bfloat16*>(a.y_bf16) + row * 512, lane);
      row_apply(u, y, mean, rstd, a.g2, a.b2, lane);
      row_store_bf16(u, reinterpret_cast<__nv_bfloat16*>(a.y2_bf16) + row * 512, lane);
    }
  };
  constexpr int ROWS = MODE == 1 ? 2 : 1;  // measured: two rows in flight help mode 1 only (mode 0: 35 -> 41 us with two)
  for (int64_t it = blockIdx.x * int64_t(blockDim.x >> 5) + (threadIdx.x >> 5); it < a.M;
       it += ROWS * warps_total) {
    const int64_t row_a = a.M - 1 - it;
    const int64_t row_b = row_a - warps_total;
    Row xa, xb;
    row_load(xa, a.x + row_a * 512, lane);
    if (ROWS == 2 && row_b >= 0) row_load(xb, a.x + row_b * 512, lane);
    finish_row(xa, row_a);
    if (ROWS == 2 && row_b >= 0) finish_row(xb, row_b);
  }
}

// ------------------------------------------------------------------------------------------
// head outputs: eight lanes per row (four rows per warp and pass, two passes in flight), 256-wide dot
// products against the 3 weight rows held in registers.  Lane l of a row group reads the 16-byte pieces
// {8 j + l} of the row, so every group access is one contiguous 128-byte line; the reduction is three
// shuffle steps per value instead of a full warp tree per row.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
head_out_kernel(const __nv_bfloat16* __restrict__ ac, const __nv_bfloat16* __restrict__ ar,
                const float* __restrict__ wc, const float* __restrict__ bc,
                const float* __restrict__ wr, const float* __restrict__ br,
                float* __restrict__ logits, float* __restrict__ offsets, int64_t M) {
  const int lane = threadIdx.x & 31;
  const int sub = lane & 7;   // position inside the row group
  const int grp = lane >> 3;  // row of the warp's four
  float w0[32], w1[32], w2[32];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const int col = 8 * (8 * j + sub) + 4 * v;
      const float4 a = __ldg(reinterpret_cast<const float4*>(wc + col));
      const float4 b = __ldg(reinterpret_cast<const float4*>(wr + col));
      const float4 c = __ldg(reinterpret_cast<const float4*>(wr + 256 + col));
      w0[8 * j + 4 * v] = a.x; w0[8 * j + 4 * v + 1] = a.y; w0[8 * j + 4 * v + 2] = a.z; w0[8 * j + 4 * v + 3] = a.w;
      w1[8 * j + 4 * v] = b.x; w1[8 * j + 4 * v + 1] = b.y; w1[8 * j + 4 * v + 2] = b.z; w1[8 * j + 4 * v + 3] = b.w;
      w2[8 * j + 4 * v] = c.x; w2[8 * j + 4 * v + 1] = c.y; w2[8 * j + 4 * v + 2] = c.z; w2[8 * j + 4 * v + 3] = c.w;
    }
  const float bias_c = bc[0], bias_r0 = br[0], bias_r1 = br[1];
  const int64_t warps_total = int64_t(gridDim.x) * (blockDim.x >> 5);
  const int64_t warp_id = blockIdx.x * int64_t(blockDim.x >> 5) + (threadIdx.x >> 5);
  pdl_wait();
  for (int64_t row0 = warp_id * 8; row0 < M; row0 += warps_total * 8) {
    uint4 c4[2][4], r4[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t row = row0 + 4 * u + grp;
      if (row < M) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          c4[u][j] = __ldcs(reinterpret_cast<const uint4*>(ac + row * 256) + 8 * j + sub);
          r4[u][j] = __ldcs(reinterpret_cast<const uint4*>(ar + row * 256) + 8 * j + sub);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) c4[u][j] = r4[u][j] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t cw[4] = {c4[u][j].x, c4[u][j].y, c4[u][j].z, c4[u][j].w};
        const uint32_t rw[4] = {r4[u][j].x, r4[u][j].y, r4[u][j].z, r4[u][j].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float c_lo = __uint_as_float(cw[i] << 16), c_hi = __uint_as_float(cw[i] & 0xffff0000u);
          const float r_lo = __uint_as_float(rw[i] << 16), r_hi = __uint_as_float(rw[i] & 0xffff0000u);
          s0 += c_lo * w0[8 * j + 2 * i] + c_hi * w0[8 * j + 2 * i + 1];
          s1 += r_lo * w1[8 * j + 2 * i] + r_hi * w1[8 * j + 2 * i + 1];
          s2 += r_lo * w2[8 * j + 2 * i] + r_hi * w2[8 * j + 2 * i + 1];
        }
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      const int64_t row = row0 + 4 * u + grp;
      if (sub == 0 && row < M) {
        logits[row] = s0 + bias_c;
        *reinterpret_cast<float2*>(offsets + 2 * row) = make_float2(fmaxf(s1 + bias_r0, 0.0f), fmaxf(s2 + bias_r1, 0.0f));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// RowMap of a padded batch (kernels.h): one block; flags per 256-row block, then an ordered compaction by warp 0
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
row_map_kernel(const int32_t* __restrict__ lens, int B, int T, int32_t* __restrict__ blocks, int32_t* __restrict__ count,
               int32_t* __restrict__ order) {
  // order[r] = the batch element with the r-th longest video (ties by index): the attention kernel lays its CTAs out
  // longest video first, so the short CTAs of a ragged batch end the launch instead of a long one starting last
  if (order != nullptr) {
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
      const int li = lens[i];
      int rank = 0;
      for (int j = 0; j < B; ++j) {
        const int lj = lens[j];
        rank += (lj > li || (lj == li && j < i)) ? 1 : 0;
      }
      order[rank] = i;
    }
  }
  const int64_t M = int64_t(B) * T;
  const int nblk = int((M + 255) / 256);
  // pass 1: flag[m] kept in `blocks` itself (0 / 1)
  for (int m = threadIdx.x; m < nblk; m += blockDim.x) {
    const int64_t r0 = int64_t(m) * 256, r1 = r0 + 256 < M ? r0 + 256 : M;
    int valid = 0;
    for (int64_t b = r0 / T; b <= (r1 - 1) / T && !valid; ++b) {
      int len = lens[b];
      len = len < 0 ? 0 : (len > T ? T : len);
      int lim = (len + 127) & ~127;
      if (lim > T) lim = T;
      const int64_t t0 = r0 > b * T ? r0 - b * T : 0;  // first step of video b inside this block
      valid = t0 < lim;
    }
    blocks[m] = valid;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    int out = 0;
    for (int m0 = 0; m0 < nblk; m0 += 32) {
      const int m = m0 + threadIdx.x;
      const int f = m < nblk ? blocks[m] : 0;
      const unsigned bal = __ballot_sync(0xffffffffu, f != 0);
      __syncwarp();
      // in-place ordered compaction is safe: the write index never exceeds the read index
      if (f) blocks[out + __popc(bal & ((1u << threadIdx.x) - 1u))] = m;
      out += __popc(bal);
      __syncwarp();
    }
    if (threadIdx.x == 0) *count = out;
  }
}

__global__ void __launch_bounds__(256)
zero_padded_rows_kernel(float* __restrict__ logits, float* __restrict__ offsets, float* __restrict__ feats,
                        const int32_t* __restrict__ lens, int B, int T, int D) {
  // one warp per padded row, rows enumerated per video from its length on
  const int lane = threadIdx.x & 31;
  const int64_t warp = blockIdx.x * int64_t(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t warps = int64_t(gridDim.x) * (blockDim.x >> 5);
  for (int64_t row = warp; row < int64_t(B) * T; row += warps) {
    const int b = int(row / T), t = int(row - int64_t(b) * T);
    if (t < lens[b]) continue;
    if (lane == 0) {
      logits[row] = 0.f;
      offsets[2 * row] = 0.f;
      offsets[2 * row + 1] = 0.f;
    }
    for (int c = 4 * lane; c < D; c += 128) *reinterpret_cast<float4*>(feats + row * D + c) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

inline int grid_for(int64_t work_items, int per_block) {
  int64_t blocks = (work_items + per_block - 1) / per_block;
  const int64_t cap = int64_t(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return int(blocks);
}

}  // namespace

int launch_concat_cast(const float* vis, const float* aud, const float* txt, int Cv, int Ca, int Ct,
                       void* out_bf16, int64_t M, cudaStream_t stream) {
  RP_CHECK(M > 0, "concat_cast: empty");
  RP_CHECK(Cv % 8 == 0 && Ca % 8 == 0 && Ct % 8 == 0, "concat_cast: dims must be multiples of 8");
  const int64_t groups = M * ((Cv + Ca + Ct) / 8);
  concat_cast_kernel<<<grid_for(groups, 256), 256, 0, stream>>>(
      vis, aud, txt, Cv, Ca, Ct, reinterpret_cast<__nv_bfloat16*>(out_bf16), M);
  count_launch();
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

int launch_ragged_concat_cast(const void* vis, const void* aud, const void* txt, bool in_bf16, int Cv, int Ca,
                              int Ct, const int32_t* row_off, const int32_t* txt_off, const int32_t* txt_lens,
                              const int32_t* lens, int B, int T, void* out_bf16, cudaStream_t stream) {
  RP_CHECK(B > 0 && T > 0, "ragged_concat_cast: empty");
  RP_CHECK(Cv % 8 == 0 && Ca % 8 == 0 && Ct % 8 == 0, "ragged_concat_cast: dims must be multiples of 8");
  const int64_t groups = int64_t(B) * T * ((Cv + Ca + Ct) / 8);
  auto* out = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  if (in_bf16) {
    using T16 = __nv_bfloat16;
    ragged_concat_cast_kernel<T16><<<grid_for(groups, 256), 256, 0, stream>>>(
        static_cast<const T16*>(vis), static_cast<const T16*>(aud), static_cast<const T16*>(txt), Cv, Ca, Ct,
        row_off, txt_off, txt_lens, lens, B, T, out);
  } else {
    ragged_concat_cast_kernel<float><<<grid_for(groups, 256), 256, 0, stream>>>(
        static_cast<const float*>(vis), static_cast<const float*>(aud), static_cast<const float*>(txt), Cv, Ca, Ct,
        row_off, txt_off, txt_lens, lens, B, T, out);
  }
  count_launch();
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

// valid-step count of every video from its key-padding mask, plus a flag that says whether all masks are
// left-aligned (mask[b, t] == t < lens[b]): the attention and decode kernels take lengths, the reference
// takes the mask itself (models/MMCTransformer.py:132-138), so anything else must be refused.
__global__ void mask_lens_kernel(const uint8_t* __restrict__ mask, int T, int32_t* __restrict__ lens,
                                 int32_t* __restrict__ not_aligned) {
  const uint8_t* row = mask + int64_t(blockIdx.x) * T;
  int cnt = 0, last = -1;
  for (int t = threadIdx.x; t < T; t += blockDim.x)
    if (row[t] != 0) {
      ++cnt;
      last = t;  // ascending within a thread
    }
  __shared__ int s_cnt[32], s_last[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
  }
  if ((threadIdx.x & 31) == 0) {
    s_cnt[threadIdx.x >> 5] = cnt;
    s_last[threadIdx.x >> 5] = last;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int c = 0, l = -1;
    for (int w = 0; w < int(blockDim.x >> 5); ++w) {
      c += s_cnt[w];
      l = max(l, s_last[w]);
    }
    lens[blockIdx.x] = c;
    if (l + 1 != c) atomicOr(not_aligned, 1);  // a hole or a leading gap: the last valid step is not count - 1
  }
}

int launch_mask_lens(const uint8_t* mask, int B, int T, int32_t* lens, int32_t* not_aligned, cudaStream_t stream) {
  RP_CHECK(B > 0 && T > 0, "mask_lens: empty");
  RP_CUDA_CHECK(cudaMemsetAsync(not_aligned, 0, sizeof(int32_t), stream));
  mask_lens_kernel<<<B, 256, 0, stream>>>(mask, T, lens, not_aligned);
  count_launch();
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

int launch_row_map(const int32_t* lens, int B, int T, int32_t* blocks, int32_t* count, int32_t* order, cudaStream_t stream) {
  RP_CHECK(B > 0 && T > 0 && lens && blocks && count, "row_map: bad arguments");
  row_map_kernel<<<1, 256, 0, stream>>>(lens, B, T, blocks, count, order);
  count_launch();
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

int launch_zero_padded_rows(float* logits, float* offsets, float* feats, const int32_t* lens, int B, int T, int D,
                            cudaStream_t stream) {
  RP_CHECK(B > 0 && T > 0 && D % 4 == 0, "zero_padded_rows: bad arguments");
  zero_padded_rows_kernel<<<grid_for(int64_t(B) * T, 8), 256, 0, stream>>>(logits, offsets, feats, lens, B, T, D);
  count_launch();
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

int launch_cast_bf16(const float* in, void* out_bf16, int64_t n, cudaStream_t stream) {
  RP_CHECK(n > 0 && n % 8 == 0, "cast_bf16: n must be a positive multiple of 8");
  cast_bf16_kernel<<<grid_for(n / 8, 256), 256, 0, stream>>>(
      in, reinterpret_cast<__nv_bfloat16*>(out_bf16), n / 8);
  count_launch();
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

int launch_layernorm512(int mode, const LnArgs& a, cudaStream_t stream) {
  RP_CHECK(a.M > 0, "layernorm: empty");
  const int grid = grid_for(a.M, 8);
  switch (mode) {
    case 0: RP_CUDA_CHECK(launch_pdl(layernorm512_kernel<0>, dim3(grid), dim3(256), 0, stream, a)); break;
    case 1:
      RP_CHECK(a.pe != nullptr && a.T > 0, "layernorm mode 1 needs pe and T");
      RP_CUDA_CHECK(launch_pdl(layernorm512_kernel<1>, dim3(grid), dim3(256), 0, stream, a));
      break;
    case 2: RP_CUDA_CHECK(launch_pdl(layernorm512_kernel<2>, dim3(grid), dim3(256), 0, stream, a)); break;
    case 3: RP_CUDA_CHECK(launch_pdl(layernorm512_kernel<3>, dim3(grid), dim3(256), 0, stream, a)); break;
    default: set_last_error("layernorm: unknown mode %d", mode); return RP_ERR_INVALID;
  }
  count_launch();
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

int launch_head_out(const void* a_cls_bf16, const void* a_reg_bf16, const float* w_cls,
                    const float* b_cls, const float* w_reg, const float* b_reg, float* logits,
                    float* offsets, int64_t M, cudaStream_t stream) {
  RP_CHECK(M > 0, "head_out: empty");
  RP_CHECK(reinterpret_cast<uintptr_t>(offsets) % 8 == 0 &&
               (reinterpret_cast<uintptr_t>(w_cls) | reinterpret_cast<uintptr_t>(w_reg)) % 16 == 0,
           "head_out: offsets must be 8-byte and the weights 16-byte aligned");
  // persistent: one block per SM (167 registers), every warp walks the rows with the weights loaded once
  int grid = grid_for(M, 64);
  if (grid > num_sms()) grid = num_sms();
  RP_CUDA_CHECK(launch_pdl(head_out_kernel, dim3(grid), dim3(256), 0, stream,
                           reinterpret_cast<const __nv_bfloat16*>(a_cls_bf16),
                           reinterpret_cast<const __nv_bfloat16*>(a_reg_bf16), w_cls, b_cls, w_reg, b_reg, logits,
                           offsets, M));
  count_launch();
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

}  // namespace rp
