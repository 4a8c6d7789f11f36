// Attention backward for sm_100a (training step): head dim 64, bf16 operands, fp32 accumulation in
// TMEM, key-padding mask from kv_lens — the autograd of the forward in fmha.cu, i.e. of nn.MultiheadAttention inside
// nn.TransformerEncoderLayer (reference models/MMCTransformer.py:41-55, 132-138; main.py:326-333 loss.backward()).
//
// With S2 = Q' K^T (Q' = q log2(e)/8 as the forward stores it), P = 2^(S2 - lse2) and Dsum = rowsum(dO o O):
//     dSt = P o (dO V^T - Dsum)            gradient w.r.t. the true logits q.k/8
//     dV  = P^T dO        dK = ln2 dSt^T Q'  (= dSt^T q / 8)        dQ = dSt K / 8   (w.r.t. the UNSCALED q)
// With attention-weight dropout (DROP kernels; O = (keep o P) V / (1 - p) in the forward): dV = (keep o P / (1 - p))^T dO,
// dSt = P o (keep o (dO V^T) / (1 - p) - Dsum); Dsum = rowsum(dO o O) holds unchanged with the dropped O.
// P is recomputed from the log-sum-exp the forward wrote (FmhaArgs::lse); nothing of size T x T is stored.
//
// Default: fmha_bwd_fused_kernel (further down) computes dK, dV and dQ in ONE kernel, five contractions per tile pair, dQ summed
// over the key tiles in L2 by TMA reduce adds.  Deterministic mode (rp_set_attn_bwd_deterministic / RP_FMHA_BWD_FUSED=0):
// two kernels, no atomics:
//   fmha_bwd_dq_kernel    one CTA per (batch, head, 128 queries), two CTAs per SM, walks the key tiles (64 keys):
//                         S = Q K_j^T and dP = dO V_j^T (SS MMAs) -> softmax warps (thread <-> query row) -> dSt as bf16
//                         A operand in TMEM -> dQ += dSt K_j (K_j as MN-major B operand, like V in the forward)
//   fmha_bwd_dkdv_kernel  one CTA per (batch, head, 128 keys), walks the query tiles (64 queries):
//                         S^T = K Q_i^T and dP^T = V dO_i^T (double-buffered in TMEM) -> softmax warps (thread <-> key
//                         row; lse / Dsum of the tile's 64 queries staged in shared memory) -> P^T, dSt^T as bf16 A
//                         operands in TMEM -> dV += P^T dO_i, dK += dSt^T Q_i (Q_i / dO_i as MN-major B operands)
// S and dP are computed twice (once per kernel): 7 MMAs per tile pair instead of 5, in exchange for no atomics on dQ.
// Roles per CTA: dQ kernel (192 threads, two CTAs per SM) warps 0-3 softmax (TMEM lane quarter = warp), warp 4 TMA producer,
// warp 5 MMA issuer; dK/dV kernel (320 threads, one CTA per SM: its TMEM budget is the whole 512 columns) warps 0-7 softmax
// — warps w and w + 4 share lane quarter w and each walks 32 of the tile's 64 query columns, so two instruction streams per
// SM sub-partition overlap —, warp 8 TMA producer, warp 9 MMA issuer.
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include "ptx.cuh"
#include "host_util.h"
#include "kernels.h"

namespace rp {

namespace {

constexpr int HD = 64;
constexpr float LN2 = 0.6931471805599453f;

struct BwdParams {
  int B, H, T;
  const int32_t* kv_lens;
  const float* lse;    // [B, H, T] log2-domain log-sum-exp from the forward
  const float* dsum;   // [B, H, T] rowsum(dO o O)
  // DROP kernels: keep bits of the forward's attention-weight dropout (fmha.cu), row (b*H + h)*T + q, drop_ld words per row
  const uint32_t* drop_bits;
  int64_t drop_ld;
  float drop_scale;
};

// bf16 rows of 128 bytes -> swizzled (SWIZZLE_128B) staging tile for a TMA store: row r, 16-byte chunk c
__device__ __forceinline__ uint32_t swz(uint32_t tile, int r, int c) {
  return tile + uint32_t(r) * 128u + (uint32_t(c ^ (r & 7)) << 4);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// 64 fp32 accumulator columns of this thread's TMEM lane -> bf16 -> one swizzled 128-byte row
__device__ __forceinline__ void store_acc_row(uint32_t taddr, uint32_t tile, int row) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t o[32];
    tmem_ld32(taddr + 32 * half, o);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 4; ++i)
      st_shared_v4(swz(tile, row, 4 * half + i),
                   pack_bf16x2(__uint_as_float(o[8 * i + 0]), __uint_as_float(o[8 * i + 1])),
                   pack_bf16x2(__uint_as_float(o[8 * i + 2]), __uint_as_float(o[8 * i + 3])),
                   pack_bf16x2(__uint_as_float(o[8 * i + 4]), __uint_as_float(o[8 * i + 5])),
                   pack_bf16x2(__uint_as_float(o[8 * i + 6]), __uint_as_float(o[8 * i + 7])));
  }
}

// ============================================================================================================
// dQ
// ============================================================================================================
namespace dq {
// K/V ring: a TMA tile takes ~4000 cycles from issue to complete_tx under load, a key tile ~1200 cycles per SM: three stages
// left the MMA warp waiting for operands (ncu: 41 % of the samples in mbarrier polls); four is what two CTAs per SM can hold
constexpr int QT = 128, KT = 64, STAGES = 4;
constexpr int Q_BYTES = QT * HD * 2, KV_BYTES = KT * HD * 2;
constexpr int SMEM_Q = 0, SMEM_DO = Q_BYTES, SMEM_K = 2 * Q_BYTES, SMEM_V = SMEM_K + STAGES * KV_BYTES;
constexpr int SMEM_BAR = SMEM_V + STAGES * KV_BYTES;
constexpr int SMEM_TOTAL = SMEM_BAR + 256 + 1024;
// dSt (the bf16 A operand of dQ += dSt K) is double-buffered: a single buffer chains consecutive key tiles through the
// completion of the previous tile's dQ MMAs (store -> ds_ready -> issue -> MMAs -> dq_done -> next store)
constexpr int TM_S = 0, TM_DP = 64, TM_DS = 128, TM_DQ = 160, TM_DS1 = 224, TMEM_COLS = 256;
}  // namespace dq

template <bool DROP>
__global__ void __launch_bounds__(192, 2)
fmha_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                   const __grid_constant__ CUtensorMap tmdQ, const BwdParams p) {
  using namespace dq;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t bar = base + SMEM_BAR;
  static_assert(STAGES <= 4, "barrier map");
  const uint32_t qdo_full = bar, sdp_full = bar + 72, sdp_free = bar + 80;
  auto ds_ready = [&](int b_) { return bar + 88u + 8u * b_; };
  auto dq_done = [&](int b_) { return bar + 104u + 8u * b_; };
  auto kv_full = [&](int s) { return bar + 8u + 8u * s; };
  auto kv_empty = [&](int s) { return bar + 40u + 8u * s; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SMEM_BAR + 128);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqb = (p.T + QT - 1) / QT;
  const int qb = int(blockIdx.x) % nqb, bh = int(blockIdx.x) / nqb;
  const int head = bh % p.H, b = bh / p.H;
  const int q_start = qb * QT;
  int kv_len = p.kv_lens != nullptr ? p.kv_lens[b] : p.T;
  kv_len = kv_len < 0 ? 0 : (kv_len > p.T ? p.T : kv_len);
  const int n_kv = (kv_len + KT - 1) / KT;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmdQ);
    mbar_init(qdo_full, 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(sdp_free, 4);
    for (int b_ = 0; b_ < 2; ++b_) {
      mbar_init(ds_ready(b_), 4);
      mbar_init(dq_done(b_), 1);
    }
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc<TMEM_COLS>(base + SMEM_BAR + 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();
  const int col = head * HD;

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer
    if (n_kv > 0) {
      if (elect_one()) {
        mbar_expect_tx(qdo_full, 2 * Q_BYTES);
        tma_load_3d(base + SMEM_Q, &tmQ, qdo_full, col, q_start, b);
        tma_load_3d(base + SMEM_DO, &tmdO, qdo_full, col, q_start, b);
      }
      __syncwarp();
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % STAGES;
        mbar_wait(kv_empty(st), (uint32_t(j / STAGES) & 1u) ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(kv_full(st), 2 * KV_BYTES);
          tma_load_3d(base + SMEM_K + st * KV_BYTES, &tmK, kv_full(st), col, j * KT, b);
          tma_load_3d(base + SMEM_V + st * KV_BYTES, &tmV, kv_full(st), col, j * KT, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer
    if (n_kv > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(QT, KT, false, false);
      constexpr uint32_t idesc_q = make_idesc_bf16(QT, HD, false, true);  // K_j as MN-major B
      auto issue_dq = [&](int jj) {
        const int st = jj % STAGES;
        mbar_wait(ds_ready(jj & 1), uint32_t(jj >> 1) & 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t db = make_smem_desc_sw128(base + SMEM_K + st * KV_BYTES, 1024, 1024);
          const uint32_t t_ds = tmem + uint32_t((jj & 1) ? TM_DS1 : TM_DS);
#pragma unroll
          for (int k = 0; k < KT / 16; ++k)
            mma_ts(tmem + TM_DQ, t_ds + 8 * k, db + uint64_t(128 * k), idesc_q, (jj > 0 || k > 0) ? 1u : 0u);
          tc_commit(dq_done(jj & 1));
          tc_commit(kv_empty(st));
        }
        __syncwarp();
      };
      mbar_wait(qdo_full, 0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % STAGES;
        mbar_wait(kv_full(st), uint32_t(j / STAGES) & 1u);
        if (j > 0) mbar_wait(sdp_free, uint32_t(j - 1) & 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t dq_ = make_smem_desc_sw128(base + SMEM_Q, 1024, 16);
          const uint64_t ddo = make_smem_desc_sw128(base + SMEM_DO, 1024, 16);
          const uint64_t dk = make_smem_desc_sw128(base + SMEM_K + st * KV_BYTES, 1024, 16);
          const uint64_t dv = make_smem_desc_sw128(base + SMEM_V + st * KV_BYTES, 1024, 16);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            mma_ss(tmem + TM_S, dq_ + uint64_t(2 * k), dk + uint64_t(2 * k), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            mma_ss(tmem + TM_DP, ddo + uint64_t(2 * k), dv + uint64_t(2 * k), idesc_s, k > 0 ? 1u : 0u);
          tc_commit(sdp_full);
        }
        __syncwarp();
        if (j > 0) issue_dq(j - 1);
      }
      issue_dq(n_kv - 1);
    }
  } else {
    // ------------------------------------------------------------------ softmax warps: thread <-> query row
    const int row = warp * 32 + lane;
    const int qrow = q_start + row;
    const bool valid = qrow < p.T;
    const int64_t stat = (int64_t(b) * p.H + head) * p.T + qrow;
    const float lse_r = valid ? p.lse[stat] : INFINITY;
    const float dsum_r = valid ? p.dsum[stat] : 0.0f;
    const uint32_t t_lane = tmem + (uint32_t(warp * 32) << 16);
    const uint2* brow = nullptr;  // DROP: this query row's keep bits, two words per 64-key tile
    if constexpr (DROP)
      brow = reinterpret_cast<const uint2*>(p.drop_bits + ((int64_t(b) * p.H + head) * p.T + (valid ? qrow : p.T - 1)) * p.drop_ld);
    for (int j = 0; j < n_kv; ++j) {
      uint2 wj = make_uint2(~0u, ~0u);
      if constexpr (DROP) wj = __ldg(brow + j);  // (requested before the barrier wait)
      mbar_wait(sdp_full, uint32_t(j) & 1u);
      tc_fence_after();
      const int nv = kv_len - j * KT;  // valid keys in this tile (>= KT: all)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t s[32], dp[32], pk[16];
        tmem_ld32(t_lane + TM_S + 32 * half, s);
        tmem_ld32(t_lane + TM_DP + 32 * half, dp);
        tmem_ld_wait();
        if (half == 1) {  // S and dP of this tile are in registers: the next tile's MMAs may overwrite them
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(sdp_free);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float d2[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = 32 * half + 2 * i + e;
            const float pr = c < nv ? ex2_approx(__uint_as_float(s[2 * i + e]) - lse_r) : 0.0f;
            float dpv = __uint_as_float(dp[2 * i + e]);
            if constexpr (DROP) dpv = ((half ? wj.y : wj.x) & (1u << (2 * i + e))) ? dpv * p.drop_scale : 0.0f;
            d2[e] = pr * (dpv - dsum_r) * 0.125f;
          }
          pk[i] = pack_bf16x2(d2[0], d2[1]);
        }
        if (half == 0 && j >= 2) {  // dQ += dSt_{j-2} K_{j-2} must have read this dSt buffer
          mbar_wait(dq_done(j & 1), (uint32_t(j >> 1) + 1u) & 1u);
          tc_fence_after();
        }
        tmem_st16(t_lane + uint32_t((j & 1) ? TM_DS1 : TM_DS) + 16 * half, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_ready(j & 1));
    }
    // ---- epilogue: dQ -> bf16 -> swizzled smem (the Q tile's slot) -> TMA store
    const uint32_t tile = base + SMEM_Q;
    if (n_kv > 0) {
      mbar_wait(dq_done((n_kv - 1) & 1), uint32_t((n_kv - 1) >> 1) & 1u);  // (MMAs retire in order: the last commit covers all)
      tc_fence_after();
      store_acc_row(t_lane + TM_DQ, tile, row);
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c) st_shared_v4(swz(tile, row, c), 0u, 0u, 0u, 0u);
    }
    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (threadIdx.x == 0) {
      tma_store_3d(&tmdQ, tile, col, q_start, b);
      tma_store_commit();
      tma_store_wait_all<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem);
  }
}

// ============================================================================================================
// dK, dV
// ============================================================================================================
namespace dkv {
// Q/dO ring: eight stages (128 KB; one CTA per SM) keep the ~4000-cycle TMA latency off the tile loop
constexpr int KT = 128, QT = 64, STAGES = 8;
constexpr int KV_BYTES = KT * HD * 2, Q_BYTES = QT * HD * 2;
constexpr int SMEM_K = 0, SMEM_V = KV_BYTES, SMEM_RING = 2 * KV_BYTES;  // stage s: Q_i at +2 s Q_BYTES, dO_i after it
constexpr int SMEM_STAT = SMEM_RING + STAGES * 2 * Q_BYTES;             // [2 buffers][lse 64 | dsum 64] floats
constexpr int SMEM_DROP = SMEM_STAT + 2 * 128 * 4;                      // [2 buffers][64 queries][4 words] keep bits
constexpr int SMEM_BAR = SMEM_DROP + 2 * 64 * 4 * 4;
constexpr int SMEM_TOTAL = SMEM_BAR + 256 + 1024;
// P^T / dSt^T (bf16 A operands) double-buffered for the same reason as dSt in the dQ kernel: all 512 columns are used
constexpr int TM_ST = 0, TM_DPT = 128, TM_PT = 256, TM_DST = 288, TM_DV = 320, TM_DK = 384, TM_PT1 = 448, TM_DST1 = 480,
              TMEM_COLS = 512;
}  // namespace dkv

constexpr int DKV_THREADS = 320, DKV_PRODUCER = 8, DKV_MMA = 9;
template <bool DROP>
__global__ void __launch_bounds__(DKV_THREADS, 1)
fmha_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                     const __grid_constant__ CUtensorMap tmdK, const __grid_constant__ CUtensorMap tmdV,
                     const BwdParams p) {
  using namespace dkv;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t bar = base + SMEM_BAR;
  const uint32_t kv_full = bar;
  static_assert(STAGES <= 8, "barrier map");
  auto p_ready = [&](int b_) { return bar + 168u + 8u * b_; };
  auto p_free = [&](int b_) { return bar + 184u + 8u * b_; };
  auto q_full = [&](int s) { return bar + 8u + 8u * s; };
  auto q_empty = [&](int s) { return bar + 72u + 8u * s; };
  auto st_full = [&](int f) { return bar + 136u + 8u * f; };
  auto st_free = [&](int f) { return bar + 152u + 8u * f; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SMEM_BAR + 208);
  float* stat = reinterpret_cast<float*>(smem + SMEM_STAT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (p.T + KT - 1) / KT;
  const int kb = int(blockIdx.x) % nkb, bh = int(blockIdx.x) / nkb;
  const int head = bh % p.H, b = bh / p.H;
  const int k_start = kb * KT;
  const int col = head * HD;
  int kv_len = p.kv_lens != nullptr ? p.kv_lens[b] : p.T;
  kv_len = kv_len < 0 ? 0 : (kv_len > p.T ? p.T : kv_len);
  const int n_q = (p.T + QT - 1) / QT;

  if (k_start >= kv_len) {
    // every key of this tile is padding: its K / V rows receive no gradient
    pdl_wait();
    for (int i = threadIdx.x; i < KV_BYTES / 16; i += blockDim.x) st_shared_v4(base + SMEM_K + 16 * i, 0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
      tma_store_3d(&tmdK, base + SMEM_K, col, k_start, b);
      tma_store_3d(&tmdV, base + SMEM_K, col, k_start, b);
      tma_store_commit();
      tma_store_wait_all<0>();
    }
    return;
  }

  if (warp == DKV_PRODUCER && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmdK);
    tma_prefetch_desc(&tmdV);
    mbar_init(kv_full, 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(q_full(s), 1);
      mbar_init(q_empty(s), 1);
    }
    for (int f = 0; f < 2; ++f) {
      mbar_init(st_full(f), 1);
      mbar_init(st_free(f), 8);
    }
    for (int b_ = 0; b_ < 2; ++b_) {
      mbar_init(p_ready(b_), 8);
      mbar_init(p_free(b_), 1);
    }
    fence_mbar_init();
  }
  if (warp == DKV_MMA) tmem_alloc<TMEM_COLS>(base + SMEM_BAR + 208);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();

  if (warp == DKV_PRODUCER) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_expect_tx(kv_full, 2 * KV_BYTES);
      tma_load_3d(base + SMEM_K, &tmK, kv_full, col, k_start, b);
      tma_load_3d(base + SMEM_V, &tmV, kv_full, col, k_start, b);
    }
    __syncwarp();
    for (int i = 0; i < n_q; ++i) {
      const int st = i % STAGES;
      mbar_wait(q_empty(st), (uint32_t(i / STAGES) & 1u) ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(q_full(st), 2 * Q_BYTES);
        tma_load_3d(base + SMEM_RING + st * 2 * Q_BYTES, &tmQ, q_full(st), col, i * QT, b);
        tma_load_3d(base + SMEM_RING + st * 2 * Q_BYTES + Q_BYTES, &tmdO, q_full(st), col, i * QT, b);
      }
      __syncwarp();
    }
  } else if (warp == DKV_MMA) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_t = make_idesc_bf16(KT, QT, false, false);  // S^T / dP^T: [128 keys, 64 queries]
    constexpr uint32_t idesc_o = make_idesc_bf16(KT, HD, false, true);   // dV / dK: B = dO_i / Q_i, MN-major
    auto issue_dkdv = [&](int ii) {
      const int st = ii % STAGES;
      mbar_wait(p_ready(ii & 1), uint32_t(ii >> 1) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dq_mn = make_smem_desc_sw128(base + SMEM_RING + st * 2 * Q_BYTES, 1024, 1024);
        const uint64_t ddo_mn = make_smem_desc_sw128(base + SMEM_RING + st * 2 * Q_BYTES + Q_BYTES, 1024, 1024);
        const uint32_t t_pt = tmem + uint32_t((ii & 1) ? TM_PT1 : TM_PT), t_dst = tmem + uint32_t((ii & 1) ? TM_DST1 : TM_DST);
#pragma unroll
        for (int k = 0; k < QT / 16; ++k)
          mma_ts(tmem + TM_DV, t_pt + 8 * k, ddo_mn + uint64_t(128 * k), idesc_o, (ii > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < QT / 16; ++k)
          mma_ts(tmem + TM_DK, t_dst + 8 * k, dq_mn + uint64_t(128 * k), idesc_o, (ii > 0 || k > 0) ? 1u : 0u);
        tc_commit(p_free(ii & 1));
        tc_commit(q_empty(st));
      }
      __syncwarp();
    };
    mbar_wait(kv_full, 0);
    for (int i = 0; i < n_q; ++i) {
      const int st = i % STAGES, bf = i & 1;
      mbar_wait(q_full(st), uint32_t(i / STAGES) & 1u);
      if (i >= 2) mbar_wait(st_free(bf), (uint32_t(i >> 1) + 1u) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dk = make_smem_desc_sw128(base + SMEM_K, 1024, 16);
        const uint64_t dv = make_smem_desc_sw128(base + SMEM_V, 1024, 16);
        const uint64_t dq_ = make_smem_desc_sw128(base + SMEM_RING + st * 2 * Q_BYTES, 1024, 16);
        const uint64_t ddo = make_smem_desc_sw128(base + SMEM_RING + st * 2 * Q_BYTES + Q_BYTES, 1024, 16);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          mma_ss(tmem + TM_ST + 64 * bf, dk + uint64_t(2 * k), dq_ + uint64_t(2 * k), idesc_t, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          mma_ss(tmem + TM_DPT + 64 * bf, dv + uint64_t(2 * k), ddo + uint64_t(2 * k), idesc_t, k > 0 ? 1u : 0u);
        tc_commit(st_full(bf));
      }
      __syncwarp();
      if (i > 0) issue_dkdv(i - 1);
    }
    issue_dkdv(n_q - 1);
  } else {
    // ------------------------------------------------------------------ softmax warps: thread <-> (key row, half of the
    // tile's query columns)
    const int wq = warp & 3;    // TMEM lane quarter
    const int half = warp >> 2; // query columns 32 half .. 32 half + 31 of every tile
    const int row = wq * 32 + lane;
    const bool kvalid = k_start + row < kv_len;
    const uint32_t t_lane = tmem + (uint32_t(wq * 32) << 16);
    const int tid = threadIdx.x;  // 0..255; threads 0..127 also fetch the tiles' statistics / keep bits
    const int64_t stat_base = (int64_t(b) * p.H + head) * p.T;
    // lse (threads 0-63) / Dsum (threads 64-127) of the tile's queries, fetched one tile ahead
    auto fetch_stat = [&](int i) {
      const int qi = i * QT + (tid & 63);
      if (tid < 64) return qi < p.T ? p.lse[stat_base + qi] : INFINITY;
      return qi < p.T ? p.dsum[stat_base + qi] : 0.0f;
    };
    // DROP: keep bits of this tile's 128 keys for the 64 queries of a query tile: thread tid fetches words
    // 4 kb + 2 (tid & 1), +1 of query tid >> 1 (one tile ahead, like the statistics); warp w then reads word w of each
    // query (lane quarter w) as a broadcast and tests bit `lane` — the transposed walk over the mask the forward used
    uint32_t* dbits = reinterpret_cast<uint32_t*>(smem + SMEM_DROP);
    auto fetch_bits = [&](int i) {
      const int qi = i * QT + (tid >> 1);
      if (qi >= p.T) return make_uint2(0u, 0u);
      return __ldg(reinterpret_cast<const uint2*>(p.drop_bits + (stat_base + qi) * p.drop_ld + 4 * kb + 2 * (tid & 1)));
    };
    const bool fetcher = tid < 128;
    if (fetcher) {
      stat[tid] = fetch_stat(0);
      if constexpr (DROP) reinterpret_cast<uint2*>(dbits)[tid] = fetch_bits(0);
    }
    for (int i = 0; i < n_q; ++i) {
      const int bf = i & 1;
      float nxt = 0.0f;
      uint2 nxt_bits = make_uint2(0u, 0u);
      if (fetcher && i + 1 < n_q) {
        nxt = fetch_stat(i + 1);
        if constexpr (DROP) nxt_bits = fetch_bits(i + 1);
      }
      named_bar_sync(1, 256);  // buffer bf is complete (written at the end of the previous iteration)
      const float4* lse4 = reinterpret_cast<const float4*>(stat + 128 * bf);
      const float4* ds4 = reinterpret_cast<const float4*>(stat + 128 * bf + 64);
      mbar_wait(st_full(bf), uint32_t(i >> 1) & 1u);
      tc_fence_after();
      {
        uint32_t s[32], dp[32], pp[16], pd[16];
        tmem_ld32(t_lane + TM_ST + 64 * bf + 32 * half, s);
        tmem_ld32(t_lane + TM_DPT + 64 * bf + 32 * half, dp);
        tmem_ld_wait();
        {  // this warp's part of S^T / dP^T is in registers: MMAs two tiles ahead may overwrite the buffer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(st_free(bf));
        }
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 l4 = lse4[8 * half + g];
          const float4 d4 = ds4[8 * half + g];
          const float lq[4] = {l4.x, l4.y, l4.z, l4.w}, dq_[4] = {d4.x, d4.y, d4.z, d4.w};
          float pr[4], dsv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            pr[e] = kvalid ? ex2_approx(__uint_as_float(s[4 * g + e]) - lq[e]) : 0.0f;
            float dpv = __uint_as_float(dp[4 * g + e]);
            if constexpr (DROP) {
              const bool keep = (dbits[(bf * 64 + 32 * half + 4 * g + e) * 4 + wq] >> lane) & 1u;
              dpv = keep ? dpv * p.drop_scale : 0.0f;
              dsv[e] = pr[e] * (dpv - dq_[e]) * LN2;
              pr[e] = keep ? pr[e] * p.drop_scale : 0.0f;  // P_d^T feeds dV
            } else {
              dsv[e] = pr[e] * (dpv - dq_[e]) * LN2;
            }
          }
          pp[2 * g] = pack_bf16x2(pr[0], pr[1]);
          pp[2 * g + 1] = pack_bf16x2(pr[2], pr[3]);
          pd[2 * g] = pack_bf16x2(dsv[0], dsv[1]);
          pd[2 * g + 1] = pack_bf16x2(dsv[2], dsv[3]);
        }
        if (i >= 2) {  // the dV / dK MMAs of tile i - 2 must have read this P^T / dSt^T buffer
          mbar_wait(p_free(bf), (uint32_t(i >> 1) + 1u) & 1u);
          tc_fence_after();
        }
        tmem_st16(t_lane + uint32_t(bf ? TM_PT1 : TM_PT) + 16 * half, pp);
        tmem_st16(t_lane + uint32_t(bf ? TM_DST1 : TM_DST) + 16 * half, pd);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready(bf));
      if (fetcher) {
        stat[128 * (bf ^ 1) + tid] = nxt;
        if constexpr (DROP) reinterpret_cast<uint2*>(dbits)[128 * (bf ^ 1) + tid] = nxt_bits;
      }
    }
    // ---- epilogue: dV (warps 0-3), dK (warps 4-7) -> bf16 -> swizzled smem (the K / V tiles' slots) -> TMA stores
    mbar_wait(p_free((n_q - 1) & 1), uint32_t((n_q - 1) >> 1) & 1u);  // (MMAs retire in order: the last commit covers all)
    tc_fence_after();
    if (half == 0) store_acc_row(t_lane + TM_DV, base + SMEM_K, row);
    else store_acc_row(t_lane + TM_DK, base + SMEM_V, row);
    fence_proxy_async_smem();
    named_bar_sync(1, 256);
    if (threadIdx.x == 0) {
      tma_store_3d(&tmdV, base + SMEM_K, col, k_start, b);
      tma_store_3d(&tmdK, base + SMEM_V, col, k_start, b);
      tma_store_commit();
      tma_store_wait_all<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == DKV_MMA) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem);
  }
}


// ============================================================================================================
// dK, dV and dQ in one kernel (five contractions per tile pair instead of seven; dQ accumulated in L2)
// ============================================================================================================
// The dK/dV kernel above, extended by the third product of the tile pair: dQ_i += dSt_i K_j.  What changes:
//   * dSt^T leaves the softmax threads through SHARED memory (bf16 [128 keys][64 queries], 128-byte rows, SWIZZLE_128B)
//     instead of TMEM: the one tile is the K-major A operand of dK += dSt^T Q_i and — read transposed, as an MN-major A
//     operand — of dQ_i(partial) = dSt K_j (K_j as MN-major B, like V in the forward).
//   * The dQ MMA runs with M = 128 although a tile has 64 queries: its descriptor starts 16 KB BELOW the dSt tile with
//     LBO = 16 KB, so rows 0-63 of A are whatever precedes the tile (never read back) and rows 64-127 are the tile: the 64
//     real accumulator rows land in TMEM lanes 64-127, the lane quarters of warps 10 and 11 of the auxiliary warpgroup.
//     Those two otherwise idle warps drain them (tcgen05.ld -> fp32 staging tile in shared memory) and one thread hands the
//     16 KB tile to the TMA reduce path: cp.reduce.async.bulk .add.f32 into the fp32 accumulator of the (batch, head)'s
//     dQ — the adds happen in L2 (measured tools/ubench/tma_reduce.cu: 5.4 TB/s = 870 cycles per tile and SM when the 15
//     key-tile CTAs of a head add into the same 29 tiles, against ~1600 cycles per iteration here).
//   * The accumulator is laid out per tile as [16 column chunks][64 queries][4 floats] (what the drain warps can write
//     without bank conflicts); attn_bwd_dq_convert_kernel turns it into the bf16 dQ rows and applies 1 / (8 ln 2) (the
//     shared dSt tile carries the ln 2 factor of dK).
// The fp32 adds of the 15 key tiles arrive in no fixed order: dQ is reproducible to fp32 rounding of a 15-term sum, not
// bit for bit.  RP_FMHA_BWD_FUSED=0 selects the two deterministic kernels above.
namespace fus {
constexpr int KT = 128, QT = 64, STAGES = 6;
constexpr int KV_BYTES = KT * HD * 2, Q_BYTES = QT * HD * 2, DQ_BYTES = QT * HD * 4;
constexpr int SMEM_K = 0, SMEM_V = KV_BYTES, SMEM_RING = 2 * KV_BYTES;  // stage s: Q_i at +2 s Q_BYTES, dO_i after it
constexpr int SMEM_DST = SMEM_RING + STAGES * 2 * Q_BYTES;              // [2 buffers] dSt^T tiles (16 KB each)
constexpr int SMEM_DQ = SMEM_DST + 2 * KV_BYTES;                        // [2 buffers] fp32 dQ partial tiles (16 KB each)
constexpr int SMEM_STAT = SMEM_DQ + 2 * DQ_BYTES;                       // [8 warps][2 buffers][lse 32 | dsum 32] floats
constexpr int SMEM_DROP = SMEM_STAT + 8 * 128 * 4;                      // [8 warps][2 buffers][32 queries] keep-bit words
constexpr int SMEM_BAR = SMEM_DROP + 8 * 64 * 4;
constexpr int SMEM_TOTAL = SMEM_BAR + 256 + 1024;
static_assert(SMEM_DST % 1024 == 0 && SMEM_DQ % 1024 == 0 && SMEM_DST >= KV_BYTES, "swizzled tiles are 1024-byte aligned");
static_assert(SMEM_TOTAL <= 227 * 1024, "shared memory budget");
constexpr int TM_ST = 0, TM_DPT = 128, TM_PT = 256, TM_PT1 = 288, TM_DV = 320, TM_DK = 384, TM_DQ = 448, TMEM_COLS = 512;
constexpr int THREADS = 384, PRODUCER = 8, MMA = 9, DRAIN0 = 10;  // warps 10, 11 drain dQ (TMEM lane quarters 2, 3)
}  // namespace fus

template <bool DROP>
__global__ void __launch_bounds__(fus::THREADS, 1)
fmha_bwd_fused_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                      const __grid_constant__ CUtensorMap tmdK, const __grid_constant__ CUtensorMap tmdV,
                      const BwdParams p, float* __restrict__ dq_acc) {
  using namespace fus;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t bar = base + SMEM_BAR;
  const uint32_t kv_full = bar;
  static_assert(STAGES <= 8, "barrier map");
  auto q_full = [&](int s) { return bar + 8u + 8u * s; };
  auto q_empty = [&](int s) { return bar + 72u + 8u * s; };
  auto st_full = [&](int f) { return bar + 136u + 8u * f; };
  auto st_free = [&](int f) { return bar + 152u + 8u * f; };
  auto p_ready = [&](int b_) { return bar + 168u + 8u * b_; };
  auto p_free = [&](int b_) { return bar + 184u + 8u * b_; };
  const uint32_t dq_full = bar + 200u, dq_free = bar + 208u;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SMEM_BAR + 224);
  float* stat = reinterpret_cast<float*>(smem + SMEM_STAT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (p.T + KT - 1) / KT;
  const int kb = int(blockIdx.x) % nkb, bh = int(blockIdx.x) / nkb;
  const int head = bh % p.H, b = bh / p.H;
  const int k_start = kb * KT;
  const int col = head * HD;
  int kv_len = p.kv_lens != nullptr ? p.kv_lens[b] : p.T;
  kv_len = kv_len < 0 ? 0 : (kv_len > p.T ? p.T : kv_len);
  const int n_q = (p.T + QT - 1) / QT;

  if (k_start >= kv_len) {
    // every key of this tile is padding: its K / V rows receive no gradient and it adds nothing to dQ
    pdl_wait();
    for (int i = threadIdx.x; i < KV_BYTES / 16; i += blockDim.x) st_shared_v4(base + SMEM_K + 16 * i, 0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
      tma_store_3d(&tmdK, base + SMEM_K, col, k_start, b);
      tma_store_3d(&tmdV, base + SMEM_K, col, k_start, b);
      tma_store_commit();
      tma_store_wait_all<0>();
    }
    return;
  }

  if (warp == PRODUCER && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmdK);
    tma_prefetch_desc(&tmdV);
    mbar_init(kv_full, 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(q_full(s), 1);
      mbar_init(q_empty(s), 1);
    }
    for (int f = 0; f < 2; ++f) {
      mbar_init(st_full(f), 1);
      mbar_init(st_free(f), 8);
      mbar_init(p_ready(f), 8);
      mbar_init(p_free(f), 1);
    }
    mbar_init(dq_full, 1);
    mbar_init(dq_free, 2);
    fence_mbar_init();
  }
  if (warp == MMA) tmem_alloc<TMEM_COLS>(base + SMEM_BAR + 224);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();

  if (warp == PRODUCER) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_expect_tx(kv_full, 2 * KV_BYTES);
      tma_load_3d(base + SMEM_K, &tmK, kv_full, col, k_start, b);
      tma_load_3d(base + SMEM_V, &tmV, kv_full, col, k_start, b);
    }
    __syncwarp();
    for (int i = 0; i < n_q; ++i) {
      const int st = i % STAGES;
      mbar_wait(q_empty(st), (uint32_t(i / STAGES) & 1u) ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(q_full(st), 2 * Q_BYTES);
        tma_load_3d(base + SMEM_RING + st * 2 * Q_BYTES, &tmQ, q_full(st), col, i * QT, b);
        tma_load_3d(base + SMEM_RING + st * 2 * Q_BYTES + Q_BYTES, &tmdO, q_full(st), col, i * QT, b);
      }
      __syncwarp();
    }
  } else if (warp == MMA) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_t = make_idesc_bf16(KT, QT, false, false);   // S^T / dP^T: [128 keys, 64 queries]
    constexpr uint32_t idesc_o = make_idesc_bf16(KT, HD, false, true);    // dV / dK: B = dO_i / Q_i, MN-major
    constexpr uint32_t idesc_q = make_idesc_bf16(128, HD, true, true);    // dQ: A = dSt (MN-major, rows 64-127), B = K_j
    auto issue_grads = [&](int ii) {
      const int st = ii % STAGES, bf = ii & 1;
      mbar_wait(p_ready(bf), uint32_t(ii >> 1) & 1u);
      if (ii > 0) mbar_wait(dq_free, uint32_t(ii - 1) & 1u);  // the drain warps have read the previous tile's dQ partial
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dq_mn = make_smem_desc_sw128(base + SMEM_RING + st * 2 * Q_BYTES, 1024, 1024);
        const uint64_t ddo_mn = make_smem_desc_sw128(base + SMEM_RING + st * 2 * Q_BYTES + Q_BYTES, 1024, 1024);
        const uint32_t dst_tile = base + SMEM_DST + bf * KV_BYTES;
        const uint64_t dst_k = make_smem_desc_sw128(dst_tile, 1024, 16);                 // [128 keys][64 q], K-major A
        const uint64_t dst_mn = make_smem_desc_sw128(dst_tile - KV_BYTES, 1024, KV_BYTES);  // transposed: tile = rows 64-127
        const uint64_t dk_mn = make_smem_desc_sw128(base + SMEM_K, 1024, 1024);
        const uint32_t t_pt = tmem + uint32_t(bf ? TM_PT1 : TM_PT);
#pragma unroll
        for (int k = 0; k < QT / 16; ++k)
          mma_ts(tmem + TM_DV, t_pt + 8 * k, ddo_mn + uint64_t(128 * k), idesc_o, (ii > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < QT / 16; ++k)
          mma_ss(tmem + TM_DK, dst_k + uint64_t(2 * k), dq_mn + uint64_t(128 * k), idesc_o, (ii > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < KT / 16; ++k)
          mma_ss(tmem + TM_DQ, dst_mn + uint64_t(128 * k), dk_mn + uint64_t(128 * k), idesc_q, k > 0 ? 1u : 0u);
        tc_commit(dq_full);
        tc_commit(p_free(bf));
        tc_commit(q_empty(st));
      }
      __syncwarp();
    };
    mbar_wait(kv_full, 0);
    for (int i = 0; i < n_q; ++i) {
      const int st = i % STAGES, bf = i & 1;
      mbar_wait(q_full(st), uint32_t(i / STAGES) & 1u);
      if (i >= 2) mbar_wait(st_free(bf), (uint32_t(i >> 1) + 1u) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dk = make_smem_desc_sw128(base + SMEM_K, 1024, 16);
        const uint64_t dv = make_smem_desc_sw128(base + SMEM_V, 1024, 16);
        const uint64_t dq_ = make_smem_desc_sw128(base + SMEM_RING + st * 2 * Q_BYTES, 1024, 16);
        const uint64_t ddo = make_smem_desc_sw128(base + SMEM_RING + st * 2 * Q_BYTES + Q_BYTES, 1024, 16);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          mma_ss(tmem + TM_ST + 64 * bf, dk + uint64_t(2 * k), dq_ + uint64_t(2 * k), idesc_t, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          mma_ss(tmem + TM_DPT + 64 * bf, dv + uint64_t(2 * k), ddo + uint64_t(2 * k), idesc_t, k > 0 ? 1u : 0u);
        tc_commit(st_full(bf));
      }
      __syncwarp();
      // the gradient MMAs of tile i - 2, not i - 1: S^T / dP^T of tile i then sit in the tensor-core queue AHEAD of them and
      // are complete long before the softmax warps finish tile i - 1 (issued behind the gradient MMAs of tile i - 1 they
      // arrived when the softmax warps already waited: 17 % of their samples, profiles/r02_notes.md 8)
      if (i > 1) issue_grads(i - 2);
    }
    if (n_q > 1) issue_grads(n_q - 2);
    issue_grads(n_q - 1);
  } else if (warp >= DRAIN0) {
    // ------------------------------------------------------------------ dQ drain: TMEM lanes 64-127 -> fp32 staging tile
    // -> TMA reduce (add.f32) into this (batch, head)'s accumulator tile i
    const int row = (warp - DRAIN0) * 32 + lane;  // query row of the tile
    const uint32_t t_lane = tmem + (uint32_t(64 + (warp - DRAIN0) * 32) << 16) + TM_DQ;
    const bool leader = warp == DRAIN0 && lane == 0;
    float* acc = dq_acc + int64_t(bh) * n_q * (QT * HD);
    for (int i = 0; i < n_q; ++i) {
      const uint32_t stage = base + SMEM_DQ + (i & 1) * DQ_BYTES;
      uint32_t a[32], c[32];
      mbar_wait(dq_full, uint32_t(i) & 1u);
      tc_fence_after();
      tmem_ld32(t_lane, a);
      tmem_ld32(t_lane + 32, c);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free);
      named_bar_sync(2, 64);  // the reduce that last read this staging buffer has finished reading (leader, below)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        st_shared_v4(stage + uint32_t(j * 64 + row) * 16u, a[4 * j], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]);
        st_shared_v4(stage + uint32_t((8 + j) * 64 + row) * 16u, c[4 * j], c[4 * j + 1], c[4 * j + 2], c[4 * j + 3]);
      }
      fence_proxy_async_smem();
      named_bar_sync(2, 64);
      if (leader) {
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(acc + int64_t(i) * (QT * HD)),
                     "r"(stage), "n"(DQ_BYTES)
                     : "memory");
        tma_store_commit();
        tma_store_wait_read<1>();  // every reduce but this one has read its staging buffer: the other buffer is free
      }
    }
    if (leader) tma_store_wait_all<0>();
  } else {
    // ------------------------------------------------------------------ softmax warps: thread <-> (key row, half of the
    // tile's query columns)
    const int wq = warp & 3;    // TMEM lane quarter
    const int half = warp >> 2; // query columns 32 half .. 32 half + 31 of every tile
    const int row = wq * 32 + lane;
    const bool kvalid = k_start + row < kv_len;
    const uint32_t t_lane = tmem + (uint32_t(wq * 32) << 16);
    const int64_t stat_base = (int64_t(b) * p.H + head) * p.T;
    // No barrier ties the eight warps together: each one stages the statistics (and keep bits) of ITS 32 query columns in
    // its own shared-memory rows, one tile ahead (lane l fetches column l; every lane then reads them as broadcasts)
    float* wst = stat + warp * 128;  // [2 buffers][lse 32 | Dsum 32]
    // DROP: word 4 kb + wq of a query's keep bits holds this warp's 32 keys; bit `lane` is this thread's key
    uint32_t* wbits = reinterpret_cast<uint32_t*>(smem + SMEM_DROP) + warp * 64;  // [2 buffers][32 queries]
    float l_n = INFINITY, d_n = 0.0f;
    uint32_t b_n = 0u;
    auto fetch = [&](int i) {
      const int qi = i * QT + 32 * half + lane;
      const bool in = qi < p.T;
      l_n = in ? p.lse[stat_base + qi] : INFINITY;
      d_n = in ? p.dsum[stat_base + qi] : 0.0f;
      if constexpr (DROP) b_n = in ? __ldg(p.drop_bits + (stat_base + qi) * p.drop_ld + 4 * kb + wq) : 0u;
    };
    auto stage = [&](int bf) {
      wst[64 * bf + lane] = l_n;
      wst[64 * bf + 32 + lane] = d_n;
      if constexpr (DROP) wbits[32 * bf + lane] = b_n;
      __syncwarp();
    };
    fetch(0);
    stage(0);
    for (int i = 0; i < n_q; ++i) {
      const int bf = i & 1;
      if (i + 1 < n_q) fetch(i + 1);  // in flight over the whole tile
      // the tile's statistics are read BEFORE the barrier wait and the TMEM loads (which are compiler barriers): their
      // shared-memory latency then overlaps both instead of following them
      const float4* lse4 = reinterpret_cast<const float4*>(wst + 64 * bf);
      const float4* ds4 = reinterpret_cast<const float4*>(wst + 64 * bf + 32);
      float4 lpre[8], dpre[8];
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        lpre[g] = lse4[g];
        dpre[g] = ds4[g];
      }
      mbar_wait(st_full(bf), uint32_t(i >> 1) & 1u);
      tc_fence_after();
      {
        uint32_t s[32], dp[32], pp[16], pd[16];
        tmem_ld32(t_lane + TM_ST + 64 * bf + 32 * half, s);
        tmem_ld32(t_lane + TM_DPT + 64 * bf + 32 * half, dp);
        tmem_ld_wait();
        {  // this warp's part of S^T / dP^T is in registers: MMAs two tiles ahead may overwrite the buffer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(st_free(bf));
        }
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 l4 = lpre[g];
          const float4 d4 = dpre[g];
          const float lq[4] = {l4.x, l4.y, l4.z, l4.w}, dq_[4] = {d4.x, d4.y, d4.z, d4.w};
          float pr[4], dsv[4];
          uint32_t kw[4] = {0u, 0u, 0u, 0u};  // DROP: the keep-bit words of these four query columns (one 16-byte load)
          if constexpr (DROP) {
            const uint4 w4 = reinterpret_cast<const uint4*>(wbits + 32 * bf)[g];
            kw[0] = w4.x; kw[1] = w4.y; kw[2] = w4.z; kw[3] = w4.w;
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            pr[e] = kvalid ? ex2_approx(__uint_as_float(s[4 * g + e]) - lq[e]) : 0.0f;
            float dpv = __uint_as_float(dp[4 * g + e]);
            if constexpr (DROP) {
              const bool keep = (kw[e] >> lane) & 1u;
              dpv = keep ? dpv * p.drop_scale : 0.0f;
              dsv[e] = pr[e] * (dpv - dq_[e]) * LN2;
              pr[e] = keep ? pr[e] * p.drop_scale : 0.0f;  // P_d^T feeds dV
            } else {
              dsv[e] = pr[e] * (dpv - dq_[e]) * LN2;
            }
          }
          pp[2 * g] = pack_bf16x2(pr[0], pr[1]);
          pp[2 * g + 1] = pack_bf16x2(pr[2], pr[3]);
          pd[2 * g] = pack_bf16x2(dsv[0], dsv[1]);
          pd[2 * g + 1] = pack_bf16x2(dsv[2], dsv[3]);
        }
        if (i >= 2) {  // the dV / dK / dQ MMAs of tile i - 2 must have read this P^T buffer and this dSt^T tile
          mbar_wait(p_free(bf), (uint32_t(i >> 1) + 1u) & 1u);
          tc_fence_after();
        }
        tmem_st16(t_lane + uint32_t(bf ? TM_PT1 : TM_PT) + 16 * half, pp);
        const uint32_t dst_tile = base + SMEM_DST + bf * KV_BYTES;  // key row `row`, query columns 32 half .. +31 = 64 bytes
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st_shared_v4(swz(dst_tile, row, 4 * half + j), pd[4 * j], pd[4 * j + 1], pd[4 * j + 2], pd[4 * j + 3]);
      }
      tmem_st_wait();
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready(bf));
      if (i + 1 < n_q) stage(bf ^ 1);
    }
    // ---- epilogue: dV (warps 0-3), dK (warps 4-7) -> bf16 -> swizzled smem (the K / V tiles' slots) -> TMA stores
    mbar_wait(p_free((n_q - 1) & 1), uint32_t((n_q - 1) >> 1) & 1u);  // (MMAs retire in order: the last commit covers all)
    tc_fence_after();
    if (half == 0) store_acc_row(t_lane + TM_DV, base + SMEM_K, row);
    else store_acc_row(t_lane + TM_DK, base + SMEM_V, row);
    fence_proxy_async_smem();
    named_bar_sync(1, 256);
    if (threadIdx.x == 0) {
      tma_store_3d(&tmdV, base + SMEM_K, col, k_start, b);
      tma_store_3d(&tmdK, base + SMEM_V, col, k_start, b);
      tma_store_commit();
      tma_store_wait_all<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == fus::MMA) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem);
  }
}

// fp32 dQ accumulator (per (batch, head): [query tiles][16 column chunks][64 queries][4 floats]) -> bf16 rows of dq, scaled
// by 1 / (8 ln 2); thread <-> (query row of the tile, four column chunks)
__global__ void __launch_bounds__(256)
attn_bwd_dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dq, int64_t ld_dq, int T, int H, int n_q) {
  pdl_wait();
  const int i = int(blockIdx.x) % n_q, bh = int(blockIdx.x) / n_q;
  const int head = bh % H, b = bh / H;
  const int row = threadIdx.x & 63, cg = threadIdx.x >> 6;
  const int q = i * 64 + row;
  if (q >= T) return;
  const float4* tile = reinterpret_cast<const float4*>(acc) + (int64_t(bh) * n_q + i) * 1024;
  constexpr float SC = 0.125f / LN2;
  uint32_t w[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 v = tile[(4 * cg + j) * 64 + row];
    w[2 * j] = pack_bf16x2(v.x * SC, v.y * SC);
    w[2 * j + 1] = pack_bf16x2(v.z * SC, v.w * SC);
  }
  uint4* out = reinterpret_cast<uint4*>(dq + (int64_t(b) * T + q) * ld_dq + head * 64 + 16 * cg);
  out[0] = make_uint4(w[0], w[1], w[2], w[3]);
  out[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

}  // namespace

int launch_attn_bwd_dsum(const void* o, const void* d_o, int B, int T, int H, float* dsum, cudaStream_t stream);

static int g_bwd_deterministic = -1;
void set_fmha_bwd_deterministic(int on) { g_bwd_deterministic = on < 0 ? -1 : (on != 0 ? 1 : 0); }

int launch_fmha_bwd(const FmhaBwdArgs& a, cudaStream_t stream) {
  RP_CHECK(a.B > 0 && a.H > 0 && a.T > 0, "fmha_bwd: empty problem");
  RP_CHECK(a.q && a.k && a.v && a.o && a.d_o && a.lse && a.dsum && a.dq && a.dk && a.dv, "fmha_bwd: null argument");
  RP_CHECK(a.ld_qkv % 8 == 0 && a.ld_o % 8 == 0 && a.ld_dqkv % 8 == 0, "fmha_bwd: pitches must be multiples of 8 elements");
  RP_CHECK(a.H % 8 == 0 && a.ld_o == int64_t(a.H) * HD, "fmha_bwd: o / dO must be dense [B, T, H*64] with H %% 8 == 0");
  const uint64_t cols = uint64_t(a.H) * HD;
  const CUtensorMapDataType bf = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  int rc;
  if ((rc = launch_attn_bwd_dsum(a.o, a.d_o, a.B, a.T, a.H, a.dsum, stream))) return rc;
  BwdParams p{a.B, a.H, a.T, a.kv_lens, a.lse, a.dsum, a.drop_bits, a.drop_ld, a.drop_scale};
  const bool drop = a.drop_bits != nullptr;
  RP_CHECK(!drop || (a.drop_ld % 4 == 0 && a.drop_ld * 32 >= ((int64_t(a.T) + 127) / 128) * 128 &&
                     reinterpret_cast<uintptr_t>(a.drop_bits) % 16 == 0),
           "fmha_bwd: keep-bit rows must be 16-byte aligned and cover every 128-key tile");
  const uint64_t T = a.T, B = a.B;
  static const bool fused_default = !(getenv("RP_FMHA_BWD_FUSED") && atoi(getenv("RP_FMHA_BWD_FUSED")) == 0);
  const bool fused = g_bwd_deterministic < 0 ? fused_default : g_bwd_deterministic == 0;
  if (fused) {
    // per-device fp32 dQ accumulator, grown on demand and kept (an attention backward of the same shape follows every step)
    static float* acc_on[kMaxDevices];
    static size_t acc_bytes_on[kMaxDevices];
    const int dev = current_device();
    const int n_q = (a.T + fus::QT - 1) / fus::QT;
    const size_t need = size_t(a.B) * a.H * n_q * fus::DQ_BYTES;
    if (acc_bytes_on[dev] < need) {
      if (acc_on[dev] != nullptr) {
        RP_CUDA_CHECK(cudaStreamSynchronize(stream));
        RP_CUDA_CHECK(cudaFree(acc_on[dev]));
        acc_on[dev] = nullptr;
        acc_bytes_on[dev] = 0;
      }
      RP_CUDA_CHECK(cudaMalloc(&acc_on[dev], need));
      acc_bytes_on[dev] = need;
    }
    float* acc = acc_on[dev];
    RP_CUDA_CHECK(cudaMemsetAsync(acc, 0, need, stream));
    CUtensorMap tmQ, tmK, tmV, tmdO, tmdK, tmdV;
    if ((rc = make_tmap_3d(&tmQ, bf, a.q, cols, T, B, a.ld_qkv * 2, T * a.ld_qkv * 2, HD, fus::QT))) return rc;
    if ((rc = make_tmap_3d(&tmK, bf, a.k, cols, T, B, a.ld_qkv * 2, T * a.ld_qkv * 2, HD, fus::KT))) return rc;
    if ((rc = make_tmap_3d(&tmV, bf, a.v, cols, T, B, a.ld_qkv * 2, T * a.ld_qkv * 2, HD, fus::KT))) return rc;
    if ((rc = make_tmap_3d(&tmdO, bf, a.d_o, cols, T, B, a.ld_o * 2, T * a.ld_o * 2, HD, fus::QT))) return rc;
    if ((rc = make_tmap_3d(&tmdK, bf, a.dk, cols, T, B, a.ld_dqkv * 2, T * a.ld_dqkv * 2, HD, fus::KT))) return rc;
    if ((rc = make_tmap_3d(&tmdV, bf, a.dv, cols, T, B, a.ld_dqkv * 2, T * a.ld_dqkv * 2, HD, fus::KT))) return rc;
    static bool configured_on[kMaxDevices];
    bool& configured = configured_on[dev];
    if (!configured) {
      RP_CUDA_CHECK(cudaFuncSetAttribute(fmha_bwd_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fus::SMEM_TOTAL));
      RP_CUDA_CHECK(cudaFuncSetAttribute(fmha_bwd_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fus::SMEM_TOTAL));
      configured = true;
    }
    const unsigned grid = unsigned((a.T + fus::KT - 1) / fus::KT) * unsigned(a.H) * unsigned(a.B);
    if (drop)
      RP_CUDA_CHECK(launch_pdl(fmha_bwd_fused_kernel<true>, dim3(grid), dim3(fus::THREADS), fus::SMEM_TOTAL, stream, tmQ, tmK, tmV,
                               tmdO, tmdK, tmdV, p, acc));
    else
      RP_CUDA_CHECK(launch_pdl(fmha_bwd_fused_kernel<false>, dim3(grid), dim3(fus::THREADS), fus::SMEM_TOTAL, stream, tmQ, tmK, tmV,
                               tmdO, tmdK, tmdV, p, acc));
    count_launch();
    RP_CUDA_CHECK(launch_pdl(attn_bwd_dq_convert_kernel, dim3(unsigned(n_q) * unsigned(a.H) * unsigned(a.B)), dim3(256), 0, stream,
                             static_cast<const float*>(acc), static_cast<__nv_bfloat16*>(a.dq), a.ld_dqkv, a.T, a.H, n_q));
    count_launch();
    RP_CUDA_CHECK(cudaGetLastError());
    return RP_OK;
  }
  {
    CUtensorMap tmQ, tmK, tmV, tmdO, tmdQ;
    if ((rc = make_tmap_3d(&tmQ, bf, a.q, cols, T, B, a.ld_qkv * 2, T * a.ld_qkv * 2, HD, dq::QT))) return rc;
    if ((rc = make_tmap_3d(&tmK, bf, a.k, cols, T, B, a.ld_qkv * 2, T * a.ld_qkv * 2, HD, dq::KT))) return rc;
    if ((rc = make_tmap_3d(&tmV, bf, a.v, cols, T, B, a.ld_qkv * 2, T * a.ld_qkv * 2, HD, dq::KT))) return rc;
    if ((rc = make_tmap_3d(&tmdO, bf, a.d_o, cols, T, B, a.ld_o * 2, T * a.ld_o * 2, HD, dq::QT))) return rc;
    if ((rc = make_tmap_3d(&tmdQ, bf, a.dq, cols, T, B, a.ld_dqkv * 2, T * a.ld_dqkv * 2, HD, dq::QT))) return rc;
    static bool configured_on[kMaxDevices];
    bool& configured = configured_on[current_device()];
    if (!configured) {
      RP_CUDA_CHECK(cudaFuncSetAttribute(fmha_bwd_dq_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dq::SMEM_TOTAL));
      RP_CUDA_CHECK(cudaFuncSetAttribute(fmha_bwd_dq_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dq::SMEM_TOTAL));
      configured = true;
    }
    const unsigned grid = unsigned((a.T + dq::QT - 1) / dq::QT) * unsigned(a.H) * unsigned(a.B);
    if (drop)
      RP_CUDA_CHECK(launch_pdl(fmha_bwd_dq_kernel<true>, dim3(grid), dim3(192), dq::SMEM_TOTAL, stream, tmQ, tmK, tmV, tmdO, tmdQ, p));
    else
      RP_CUDA_CHECK(launch_pdl(fmha_bwd_dq_kernel<false>, dim3(grid), dim3(192), dq::SMEM_TOTAL, stream, tmQ, tmK, tmV, tmdO, tmdQ, p));
    count_launch();
  }
  {
    CUtensorMap tmQ, tmK, tmV, tmdO, tmdK, tmdV;
    if ((rc = make_tmap_3d(&tmQ, bf, a.q, cols, T, B, a.ld_qkv * 2, T * a.ld_qkv * 2, HD, dkv::QT))) return rc;
    if ((rc = make_tmap_3d(&tmK, bf, a.k, cols, T, B, a.ld_qkv * 2, T * a.ld_qkv * 2, HD, dkv::KT))) return rc;
    if ((rc = make_tmap_3d(&tmV, bf, a.v, cols, T, B, a.ld_qkv * 2, T * a.ld_qkv * 2, HD, dkv::KT))) return rc;
    if ((rc = make_tmap_3d(&tmdO, bf, a.d_o, cols, T, B, a.ld_o * 2, T * a.ld_o * 2, HD, dkv::QT))) return rc;
    if ((rc = make_tmap_3d(&tmdK, bf, a.dk, cols, T, B, a.ld_dqkv * 2, T * a.ld_dqkv * 2, HD, dkv::KT))) return rc;
    if ((rc = make_tmap_3d(&tmdV, bf, a.dv, cols, T, B, a.ld_dqkv * 2, T * a.ld_dqkv * 2, HD, dkv::KT))) return rc;
    static bool configured_on[kMaxDevices];
    bool& configured = configured_on[current_device()];
    if (!configured) {
      RP_CUDA_CHECK(cudaFuncSetAttribute(fmha_bwd_dkdv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dkv::SMEM_TOTAL));
      RP_CUDA_CHECK(cudaFuncSetAttribute(fmha_bwd_dkdv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dkv::SMEM_TOTAL));
      configured = true;
    }
    const unsigned grid = unsigned((a.T + dkv::KT - 1) / dkv::KT) * unsigned(a.H) * unsigned(a.B);
    if (drop)
      RP_CUDA_CHECK(launch_pdl(fmha_bwd_dkdv_kernel<true>, dim3(grid), dim3(DKV_THREADS), dkv::SMEM_TOTAL, stream, tmQ, tmK, tmV,
                               tmdO, tmdK, tmdV, p));
    else
      RP_CUDA_CHECK(launch_pdl(fmha_bwd_dkdv_kernel<false>, dim3(grid), dim3(DKV_THREADS), dkv::SMEM_TOTAL, stream, tmQ, tmK, tmV,
                               tmdO, tmdK, tmdV, p));
    count_launch();
  }
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

}  // namespace rp
