// Fused multi-head attention forward for sm_100a, head dim 64, bf16 operands, fp32 softmax.
//
// One CTA = one (batch, head, pair of 128-row query tiles).  Roles (320 threads):
//   warps 0-3 : softmax warpgroup for query tile 0   (thread i <-> TMEM lane i <-> query row i)
//   warps 4-7 : softmax warpgroup for query tile 1
//   warp  8   : TMA producer (Q once, K/V tiles through a 3-stage ring, SWIZZLE_128B)
//   warp  9   : tcgen05.mma issuer + TMEM owner
// TMEM (512 columns): S0 [0,128)  S1 [128,256)  O0 [256,320)  O1 [320,384).
// P (bf16) is written back over the first 64 columns of its own S tile and consumed by the PV MMA
// straight from TMEM (A operand in TMEM); V is consumed MN-major from the same swizzled smem tile
// TMA delivers.  Issue order  PV0_j, QK0_{j+1}, PV1_j, QK1_{j+1}  keeps the tensor pipe busy with
// one query tile while the other one is in softmax.
//
// Semantics follow the reference:
//   mask_mode 0  nn.MultiheadAttention with key_padding_mask (models/MMCTransformer.py:132-138):
//                keys >= kv_lens[b] receive -inf; every query row (padded or not) is computed.
//   mask_mode 1  models/transformer.py:52-81 MultiHeadAttention.forward: masked_fill(mask==0, -1e9).
// Q must arrive pre-scaled by log2(e)/sqrt(64) so that S is already in the exp2 domain.
#include <math.h>

#include "ptx.cuh"
#include "host_util.h"
#include "kernels.h"

namespace rp {

namespace {

constexpr int QT = 128;
constexpr int KT = 128;
constexpr int HD = 64;
constexpr int KV_STAGES = 3;
constexpr int TILE_BYTES = QT * HD * 2;  // 16 KB (Q, K and V tiles are all 128 x 64 bf16)
constexpr int SMEM_Q_OFF = 0;
constexpr int SMEM_K_OFF = 2 * TILE_BYTES;
constexpr int SMEM_V_OFF = SMEM_K_OFF + KV_STAGES * TILE_BYTES;
constexpr int SMEM_BAR_OFF = SMEM_V_OFF + KV_STAGES * TILE_BYTES;  // 131072
constexpr int SMEM_TOTAL = SMEM_BAR_OFF + 256 + 1024;
constexpr int NUM_THREADS = 320;
constexpr int TMEM_COLS = 512;
constexpr int TM_S = 0;     // + q*128
constexpr int TM_O = 256;   // + q*64
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units; stale max keeps p <= 2^8
constexpr float MASK_FILL_LOG2 = -1.0e9f * 1.4426950408889634f;

struct FmhaParams {
  int B, H, Tq, Tk;
  const int32_t* kv_lens;
  const uint8_t* mask;
  int64_t mask_b_stride, mask_q_stride;
};

template <int MASK_MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
fmha_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                const FmhaParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);

  const uint32_t bar_base = base + SMEM_BAR_OFF;
  auto q_full = [&](int q) { return bar_base + 8u * q; };
  auto k_full = [&](int s) { return bar_base + 16u + 8u * s; };
  auto k_empty = [&](int s) { return bar_base + 40u + 8u * s; };
  auto v_full = [&](int s) { return bar_base + 64u + 8u * s; };
  auto v_empty = [&](int s) { return bar_base + 88u + 8u * s; };
  auto s_full = [&](int q) { return bar_base + 112u + 8u * q; };
  auto p_ready = [&](int q) { return bar_base + 128u + 8u * q; };
  auto o_done = [&](int q) { return bar_base + 144u + 8u * q; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SMEM_BAR_OFF + 160);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int pair = blockIdx.x;
  const int head = blockIdx.y;
  const int b = blockIdx.z;

  const int q_start0 = pair * (2 * QT);
  const bool q1_active = (q_start0 + QT) < p.Tq;
  const int nq = q1_active ? 2 : 1;
  int kv_len = p.Tk;
  if (MASK_MODE == 0 && p.kv_lens != nullptr) {
    kv_len = p.kv_lens[b];
    kv_len = kv_len < 0 ? 0 : (kv_len > p.Tk ? p.Tk : kv_len);
  }
  const int n_kv = (kv_len + KT - 1) / KT;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    for (int q = 0; q < 2; ++q) {
      mbar_init(q_full(q), 1);
      mbar_init(s_full(q), 1);
      mbar_init(p_ready(q), 128);
      mbar_init(o_done(q), 1);
    }
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(k_full(s), 1);
      mbar_init(k_empty(s), 1);
      mbar_init(v_full(s), 1);
      mbar_init(v_empty(s), 1);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<TMEM_COLS>(base + SMEM_BAR_OFF + 160);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0 && n_kv > 0) {
      const int col = head * HD;
      mbar_expect_tx(q_full(0), TILE_BYTES);
      tma_load_3d(base + SMEM_Q_OFF, &tmQ, q_full(0), col, q_start0, b);
      if (q1_active) {
        mbar_expect_tx(q_full(1), TILE_BYTES);
        tma_load_3d(base + SMEM_Q_OFF + TILE_BYTES, &tmQ, q_full(1), col, q_start0 + QT, b);
      }
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(k_empty(st), ph ^ 1u);
        mbar_expect_tx(k_full(st), TILE_BYTES);
        tma_load_3d(base + SMEM_K_OFF + st * TILE_BYTES, &tmK, k_full(st), col, j * KT, b);
        mbar_wait(v_empty(st), ph ^ 1u);
        mbar_expect_tx(v_full(st), TILE_BYTES);
        tma_load_3d(base + SMEM_V_OFF + st * TILE_BYTES, &tmV, v_full(st), col, j * KT, b);
        if (++st == KV_STAGES) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 9) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0 && n_kv > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(QT, KT, false, false);
      constexpr uint32_t idesc_o = make_idesc_bf16(QT, HD, false, true);  // V is MN-major
      auto issue_qk = [&](int q, int st) {
        const uint32_t a_addr = base + SMEM_Q_OFF + q * TILE_BYTES;
        const uint32_t b_addr = base + SMEM_K_OFF + st * TILE_BYTES;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          mma_ss(tmem_base + TM_S + q * 128, make_smem_desc_sw128(a_addr + k * 32, 1024, 16),
                 make_smem_desc_sw128(b_addr + k * 32, 1024, 16), idesc_s, k > 0 ? 1u : 0u);
        }
      };
      auto issue_pv = [&](int q, int st, bool acc) {
        const uint32_t b_addr = base + SMEM_V_OFF + st * TILE_BYTES;
#pragma unroll
        for (int k = 0; k < KT / 16; ++k) {
          // A: P tile in TMEM, 16 keys = 8 packed columns per step; B: 16 key rows of 128 B each
          mma_ts(tmem_base + TM_O + q * 64, tmem_base + TM_S + q * 128 + k * 8,
                 make_smem_desc_sw128(b_addr + k * 2048, 1024, 1024), idesc_o,
                 (acc || k > 0) ? 1u : 0u);
        }
      };

      mbar_wait(q_full(0), 0);
      mbar_wait(k_full(0), 0);
      tc_fence_after();
      issue_qk(0, 0);
      tc_commit(s_full(0));
      if (nq == 2) {
        mbar_wait(q_full(1), 0);
        tc_fence_after();
        issue_qk(1, 0);
        tc_commit(s_full(1));
      }
      tc_commit(k_empty(0));

      for (int j = 0; j < n_kv; ++j) {
        const int st = j % KV_STAGES;
        const uint32_t ph = uint32_t(j / KV_STAGES) & 1u;
        const int j1 = j + 1;
        const int st1 = j1 % KV_STAGES;
        const uint32_t ph1 = uint32_t(j1 / KV_STAGES) & 1u;
        for (int q = 0; q < nq; ++q) {
          mbar_wait(p_ready(q), uint32_t(j) & 1u);
          if (q == 0) mbar_wait(v_full(st), ph);
          tc_fence_after();
          issue_pv(q, st, j > 0);
          if (q == nq - 1) tc_commit(v_empty(st));
          if (j1 < n_kv) {
            if (q == 0) {
              mbar_wait(k_full(st1), ph1);
              tc_fence_after();
            }
            issue_qk(q, st1);
            tc_commit(s_full(q));
            if (q == nq - 1) tc_commit(k_empty(st1));
          } else {
            tc_commit(o_done(q));
          }
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax warpgroups
    const int q = warp >> 2;
    const int wl = warp & 3;
    const int row_in_tile = wl * 32 + lane;
    const int q_start = q_start0 + q * QT;
    const uint32_t stage_smem = base + SMEM_Q_OFF + q * TILE_BYTES;  // reused for the O tile
    if (q < nq) {
      const uint32_t lane_off = uint32_t(wl * 32) << 16;
      const uint32_t t_s = tmem_base + lane_off + TM_S + q * 128;
      const uint32_t t_o = tmem_base + lane_off + TM_O + q * 64;
      float m = -INFINITY;
      float l = 0.0f;
      const uint8_t* mrow = nullptr;
      if (MASK_MODE == 1) {
        const int qrow = q_start + row_in_tile;
        if (qrow < p.Tq)
          mrow = p.mask + int64_t(b) * p.mask_b_stride + int64_t(qrow) * p.mask_q_stride;
      }

      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(s_full(q), uint32_t(j) & 1u);
        tc_fence_after();
        uint32_t sr[128];
        tmem_ld32(t_s + 0, sr + 0);
        tmem_ld32(t_s + 32, sr + 32);
        tmem_ld32(t_s + 64, sr + 64);
        tmem_ld32(t_s + 96, sr + 96);
        tmem_ld_wait();
        float* s = reinterpret_cast<float*>(sr);

        const int nv = kv_len - j * KT;  // valid keys in this tile
        if (nv < KT) {
#pragma unroll
          for (int c = 0; c < KT; ++c)
            if (c >= nv) s[c] = -INFINITY;
        }
        if (MASK_MODE == 1 && mrow != nullptr) {
          const uint8_t* mp = mrow + j * KT;
#pragma unroll
          for (int c = 0; c < KT; ++c) {
            if (c < nv && mp[c] == 0) s[c] = MASK_FILL_LOG2;
          }
        }

        float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
        for (int c = 4; c < KT; c += 4) {
          mx0 = fmaxf(mx0, s[c]);
          mx1 = fmaxf(mx1, s[c + 1]);
          mx2 = fmaxf(mx2, s[c + 2]);
          mx3 = fmaxf(mx3, s[c + 3]);
        }
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        const float m_new = fmaxf(m, mx);
        if (j == 0) {
          m = m_new;
        } else {
          const bool need = m_new > m + RESCALE_THRESHOLD;
          if (__any_sync(0xffffffffu, need)) {
            // QK_j was issued after PV_{j-1} on the same in-order pipe and its commit has been
            // observed, so O is quiescent here.
            const float alpha = ex2_approx(m - m_new);
            m = m_new;
            l *= alpha;
#pragma unroll
            for (int oc = 0; oc < 4; ++oc) {
              uint32_t o[16];
              tmem_ld16(t_o + oc * 16, o);
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 16; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
              tmem_st16(t_o + oc * 16, o);
            }
          }
        }

        float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          uint32_t pk[16];
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            const int i0 = qd * 32 + 2 * c;
            const float p0 = ex2_approx(s[i0 + 0] - m);
            const float p1 = ex2_approx(s[i0 + 1] - m);
            const float p2 = ex2_approx(s[i0 + 2] - m);
            const float p3 = ex2_approx(s[i0 + 3] - m);
            sum0 += p0; sum1 += p1; sum2 += p2; sum3 += p3;
            pk[c] = pack_bf16x2(p0, p1);
            pk[c + 1] = pack_bf16x2(p2, p3);
          }
          tmem_st16(t_s + qd * 16, pk);
        }
        l += (sum0 + sum1) + (sum2 + sum3);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_ready(q));
      }

      // ---- epilogue: O / l -> bf16 -> swizzled smem (the Q tile's slot) -> TMA store
      uint32_t o[64];
      float inv_l = 0.0f;
      if (n_kv > 0) {
        mbar_wait(o_done(q), 0);
        tc_fence_after();
        tmem_ld32(t_o, o);
        tmem_ld32(t_o + 32, o + 32);
        tmem_ld_wait();
        inv_l = 1.0f / l;
      } else {
#pragma unroll
        for (int c = 0; c < 64; ++c) o[c] = 0u;
      }
      const uint32_t row_addr = stage_smem + uint32_t(row_in_tile) * 128u;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t p0 = pack_bf16x2(__uint_as_float(o[8 * i + 0]) * inv_l, __uint_as_float(o[8 * i + 1]) * inv_l);
        const uint32_t p1 = pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv_l, __uint_as_float(o[8 * i + 3]) * inv_l);
        const uint32_t p2 = pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv_l, __uint_as_float(o[8 * i + 5]) * inv_l);
        const uint32_t p3 = pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv_l, __uint_as_float(o[8 * i + 7]) * inv_l);
        const uint32_t dst = row_addr + (uint32_t(i ^ (row_in_tile & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(p0), "r"(p1), "r"(p2),
                     "r"(p3)
                     : "memory");
      }
      fence_proxy_async_smem();
      if (q == 0) named_bar_sync(1, 128); else named_bar_sync(2, 128);
      if (wl == 0 && lane == 0) {
        tma_store_3d(&tmO, stage_smem, head * HD, q_start, b);
        tma_store_commit();
        tma_store_wait_all<0>();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

template <int MASK_MODE>
int launch_mode(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                const CUtensorMap& tmO, const FmhaParams& p, dim3 grid, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    RP_CUDA_CHECK(cudaFuncSetAttribute(fmha_fwd_kernel<MASK_MODE>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    configured = true;
  }
  fmha_fwd_kernel<MASK_MODE><<<grid, NUM_THREADS, SMEM_TOTAL, stream>>>(tmQ, tmK, tmV, tmO, p);
  count_launch();
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

}  // namespace

int launch_fmha(const FmhaArgs& a, cudaStream_t stream) {
  RP_CHECK(a.B > 0 && a.H > 0 && a.Tq > 0 && a.Tk > 0, "fmha: empty problem");
  RP_CHECK(a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 && a.ldo % 8 == 0 &&
               a.bsq % 8 == 0 && a.bsk % 8 == 0 && a.bsv % 8 == 0 && a.bso % 8 == 0,
           "fmha: pitches must be multiples of 8 elements");
  RP_CHECK((reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.k) |
            reinterpret_cast<uintptr_t>(a.v) | reinterpret_cast<uintptr_t>(a.o)) % 16 == 0,
           "fmha: pointers must be 16-byte aligned");
  RP_CHECK(a.mask_mode == 0 || (a.mask_mode == 1 && a.mask != nullptr), "fmha: bad mask arguments");
  RP_CHECK(a.B <= 65535 && a.H <= 65535, "fmha: grid too large");

  const uint64_t cols = uint64_t(a.H) * HD;
  CUtensorMap tmQ, tmK, tmV, tmO;
  int rc;
  const CUtensorMapDataType bf = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  if ((rc = make_tmap_3d(&tmQ, bf, a.q, cols, a.Tq, a.B, a.ldq * 2, a.bsq * 2, HD, QT))) return rc;
  if ((rc = make_tmap_3d(&tmK, bf, a.k, cols, a.Tk, a.B, a.ldk * 2, a.bsk * 2, HD, KT))) return rc;
  if ((rc = make_tmap_3d(&tmV, bf, a.v, cols, a.Tk, a.B, a.ldv * 2, a.bsv * 2, HD, KT))) return rc;
  if ((rc = make_tmap_3d(&tmO, bf, a.o, cols, a.Tq, a.B, a.ldo * 2, a.bso * 2, HD, QT))) return rc;

  FmhaParams p{a.B, a.H, a.Tq, a.Tk, a.kv_lens, a.mask, a.mask_b_stride, a.mask_q_stride};
  dim3 grid((a.Tq + 2 * QT - 1) / (2 * QT), a.H, a.B);
  if (a.mask_mode == 0) return launch_mode<0>(tmQ, tmK, tmV, tmO, p, grid, stream);
  return launch_mode<1>(tmQ, tmK, tmV, tmO, p, grid, stream);
}

}  // namespace rp
