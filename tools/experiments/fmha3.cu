// Fused multi-head attention forward for sm_100a, head dim 64, bf16 operands, fp32 softmax —
// "independent column-half streams" variant.
//
// One CTA = one (batch, head, 128-row query tile), two CTAs per SM, 12 warps:
//   warps 0-7  : softmax; warp w works on the key-column half h = w >> 2 of every key tile and owns TMEM
//                lanes 32 (w & 3) .. +31 (thread <-> lane <-> query row)
//   warp 8     : TMA producer (Q once, K / V tiles through rings, SWIZZLE_128B)
//   warp 9, 10 : tcgen05.mma issuers, one per half         (warp 11 only completes the warpgroup)
// The two halves of a score row are two completely independent online-softmax streams: each has its
// own running reference m_h, row sum l_h and its OWN output accumulator O_h in TMEM, so the threads
// of a row never exchange anything until the very end, where
//   O = (O_0 2^(m_0 - m) + O_1 2^(m_1 - m)) / (l_0 2^(m_0 - m) + l_1 2^(m_1 - m)),  m = max(m_0, m_1).
// TMEM (256 columns): S_0 | S_1 (64 each) | O_0 | O_1 (64 each).  P_h (bf16, 32 columns) overwrites the
// first half of S_h once the thread holds its 64 scores in registers; the issuer of half h issues
//   P_h V_h (A operand from TMEM)  and then  S_h = Q K_h^T of the NEXT key tile
// back to back, and the tensor pipe executes them in that order, so S_h is only overwritten after P_h
// has been consumed.  A stream is therefore strictly sequential (MMA -> softmax -> MMA ...), and the SM
// hides that latency with the other streams: 2 CTAs x 2 halves = 4 independent softmax warps per
// scheduler that drift apart instead of running in phase.
//
// Softmax arithmetic, masking semantics and the lazy rescale (threshold 2^8) are those of fmha2.cu.
#include <math.h>
#include <stdlib.h>

#include "ptx.cuh"
#include "host_util.h"
#include "kernels.h"

namespace rp {

namespace {

constexpr int QT = 128;
constexpr int KT = 128;
constexpr int HD = 64;
constexpr int TILE_BYTES = QT * HD * 2;  // 16 KB (Q, K and V tiles are all 128 x 64 bf16)

#ifndef RP_FMHA3_KSTAGES
#define RP_FMHA3_KSTAGES 3
#endif
#ifndef RP_FMHA3_VSTAGES
#define RP_FMHA3_VSTAGES 2
#endif
constexpr int K_STAGES = RP_FMHA3_KSTAGES;
constexpr int V_STAGES = RP_FMHA3_VSTAGES;
constexpr int SMEM_Q_OFF = 0;
constexpr int SMEM_K_OFF = TILE_BYTES;
constexpr int SMEM_V_OFF = SMEM_K_OFF + K_STAGES * TILE_BYTES;
constexpr int SMEM_X_OFF = SMEM_V_OFF + V_STAGES * TILE_BYTES;  // [2 halves][2 (m,l)][128 rows] f32
constexpr int SMEM_BAR_OFF = SMEM_X_OFF + 2 * 2 * 128 * 4;
constexpr int SMEM_TOTAL = SMEM_BAR_OFF + 256 + 1024;
constexpr int SOFTMAX_WARPS = 8;
constexpr int PRODUCER_WARP = 8;
constexpr int MMA_WARP0 = 9;  // + h
constexpr int NUM_THREADS = 12 * 32;
constexpr int TMEM_COLS = 256;
constexpr int TM_S = 0;    // + 64 h   (P_h aliases columns [64h, 64h + 32))
constexpr int TM_O = 128;  // + 64 h
// register re-allocation after the role split: 384 threads x 80 at launch -> 8 x 32 x 104 + 4 x 32 x 32
constexpr int SOFTMAX_REGS = 104;
constexpr int AUX_REGS = 32;
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units; stale max keeps p <= 2^8
constexpr float MASK_FILL_LOG2 = -1.0e9f * 1.4426950408889634f;

struct FmhaParams {
  int B, H, Tq, Tk;
  const int32_t* kv_lens;
  const uint8_t* mask;
  int64_t mask_b_stride, mask_q_stride;
};

// ---- packed f32x2 helpers (sm_100 FFMA2/FADD2) ---------------------------------------------------
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b,
                                                   unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// exp2 of two fp32 values on the FMA/ALU pipes (Cody-Waite split + degree-3 minimax polynomial)
__device__ __forceinline__ void exp2_emulated2(unsigned long long x2, float& p0, float& p1) {
  float x0, x1;
  unpack2(x2, x0, x1);
  x0 = fmaxf(x0, -126.0f);
  x1 = fmaxf(x1, -126.0f);
  const unsigned long long xc = pack2(x0, x1);
  const unsigned long long t = add2(xc, pack2(12582912.0f, 12582912.0f));  // low mantissa bits hold rint(x)
  const unsigned long long n = add2(t, pack2(-12582912.0f, -12582912.0f));
  const unsigned long long f = fma2(n, pack2(-1.0f, -1.0f), xc);
  unsigned long long p = fma2(pack2(0.055171653628349304f, 0.055171653628349304f), f,
                              pack2(0.2426111251115799f, 0.2426111251115799f));
  p = fma2(p, f, pack2(0.6932609677314758f, 0.6932609677314758f));
  p = fma2(p, f, pack2(0.9999280571937561f, 0.9999280571937561f));
  float t0, t1, q0, q1;
  unpack2(t, t0, t1);
  unpack2(p, q0, q1);
  p0 = __uint_as_float((__float_as_uint(t0) << 23) + __float_as_uint(q0));
  p1 = __uint_as_float((__float_as_uint(t1) << 23) + __float_as_uint(q1));
}

template <int MASK_MODE, int EMU_PAIRS>
__global__ void __launch_bounds__(NUM_THREADS, 2)
fmha3_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                 const FmhaParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);

  const uint32_t bar_base = pin_u32(base + SMEM_BAR_OFF);
  auto q_full = [&]() { return bar_base; };
  auto k_full = [&](int s) { return bar_base + 8u + 8u * s; };     // 4 slots each
  auto k_empty = [&](int s) { return bar_base + 40u + 8u * s; };
  auto v_full = [&](int s) { return bar_base + 72u + 8u * s; };
  auto v_empty = [&](int s) { return bar_base + 104u + 8u * s; };
  auto s_full = [&](int h) { return bar_base + 136u + 8u * h; };
  auto p_ready = [&](int h) { return bar_base + 152u + 8u * h; };
  auto pv_done = [&](int h) { return bar_base + 168u + 8u * h; };
  constexpr int TMEM_SLOT_OFF = 192;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SMEM_BAR_OFF + TMEM_SLOT_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int q_start = blockIdx.x * QT;
  int kv_len = p.Tk;
  if (MASK_MODE == 0 && p.kv_lens != nullptr) {
    kv_len = p.kv_lens[b];
    kv_len = kv_len < 0 ? 0 : (kv_len > p.Tk ? p.Tk : kv_len);
  }
  const int n_kv = (kv_len + KT - 1) / KT;

  if (warp == PRODUCER_WARP && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    mbar_init(q_full(), 1);
    for (int s = 0; s < 4; ++s) {
      mbar_init(k_full(s), 1);
      mbar_init(k_empty(s), 2);  // both halves' QK^T have read the K tile
      mbar_init(v_full(s), 1);
      mbar_init(v_empty(s), 2);  // both halves' P V have read the V tile
    }
    for (int h = 0; h < 2; ++h) {
      mbar_init(s_full(h), 1);
      mbar_init(p_ready(h), 4);
      mbar_init(pv_done(h), 1);
    }
    fence_mbar_init();
  }
  if (warp == MMA_WARP0) tmem_alloc<TMEM_COLS>(base + SMEM_BAR_OFF + TMEM_SLOT_OFF);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // the set-up above overlapped the previous kernel's tail

  if (warp >= SOFTMAX_WARPS) setmaxnreg_dec<AUX_REGS>();
  if (warp == PRODUCER_WARP) {
    // ---------------------------------------------------------------- TMA producer
    if (n_kv > 0) {
      const int col = head * HD;
      if (elect_one()) {
        mbar_expect_tx(q_full(), TILE_BYTES);
        tma_load_3d(base + SMEM_Q_OFF, &tmQ, q_full(), col, q_start, b);
      }
      __syncwarp();
      auto load_k = [&](int i) {
        const int st = i % K_STAGES;
        mbar_wait(k_empty(st), (uint32_t(i / K_STAGES) & 1u) ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(k_full(st), TILE_BYTES);
          tma_load_3d(base + SMEM_K_OFF + st * TILE_BYTES, &tmK, k_full(st), col, i * KT, b);
        }
        __syncwarp();
      };
      auto load_v = [&](int i) {
        const int st = i % V_STAGES;
        mbar_wait(v_empty(st), (uint32_t(i / V_STAGES) & 1u) ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(v_full(st), TILE_BYTES);
          tma_load_3d(base + SMEM_V_OFF + st * TILE_BYTES, &tmV, v_full(st), col, i * KT, b);
        }
        __syncwarp();
      };
      // K runs ahead of V: K_{i+1} is consumed right after P V_i is issued
      load_k(0);
      for (int i = 0; i < n_kv; ++i) {
        if (i + 1 < n_kv) load_k(i + 1);
        load_v(i);
      }
    }
  } else if (warp == MMA_WARP0 || warp == MMA_WARP0 + 1) {
    // ---------------------------------------------------------------- MMA issuer of half h
    // Converged warp, one elected lane per issue group (descriptors stay in uniform registers).
    const int h = warp - MMA_WARP0;
    if (n_kv > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(QT, KT / 2, false, false);
      constexpr uint32_t idesc_o = make_idesc_bf16(QT, HD, false, true);  // V is MN-major
      // S_h = Q K[64h .. 64h+64)^T
      auto issue_qk = [&](int st) {
        if (elect_one()) {
          const uint64_t da = make_smem_desc_sw128(base + SMEM_Q_OFF, 1024, 16);
          const uint64_t db =
              make_smem_desc_sw128(base + SMEM_K_OFF + st * TILE_BYTES + h * (TILE_BYTES / 2), 1024, 16);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            mma_ss(tmem_base + TM_S + h * 64, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc_s, k > 0 ? 1u : 0u);
          tc_commit(s_full(h));
          tc_commit(k_empty(st));
        }
        __syncwarp();
      };
      // O_h += P_h V[64h .. 64h+64)
      auto issue_pv = [&](int st, bool acc, bool last) {
        if (elect_one()) {
          // A: P_h in TMEM, 16 keys = 8 packed columns per step; B: 16 key rows of 128 B each
          const uint64_t db = make_smem_desc_sw128(base + SMEM_V_OFF + st * TILE_BYTES, 1024, 1024);
#pragma unroll
          for (int k = 0; k < KT / 32; ++k)
            mma_ts(tmem_base + TM_O + h * 64, tmem_base + TM_S + h * 64 + k * 8, db + uint64_t(128 * (4 * h + k)),
                   idesc_o, (acc || k > 0) ? 1u : 0u);
          tc_commit(v_empty(st));
          if (last) tc_commit(pv_done(h));
        }
        __syncwarp();
      };
      mbar_wait(q_full(), 0);
      mbar_wait(k_full(0), 0);
      tc_fence_after();
      issue_qk(0);
      for (int j = 0; j < n_kv; ++j) {
        const int vst = j % V_STAGES;
        mbar_wait(v_full(vst), uint32_t(j / V_STAGES) & 1u);
        mbar_wait(p_ready(h), uint32_t(j) & 1u);
        tc_fence_after();
        issue_pv(vst, j > 0, j == n_kv - 1);
        if (j + 1 < n_kv) {
          const int kst = (j + 1) % K_STAGES;
          mbar_wait(k_full(kst), uint32_t((j + 1) / K_STAGES) & 1u);
          tc_fence_after();
          issue_qk(kst);  // executes after P V_j: S_h may then be overwritten
        }
      }
    }
  } else if (warp < SOFTMAX_WARPS) {
    setmaxnreg_inc<SOFTMAX_REGS>();
    // ---------------------------------------------------------------- softmax streams
    const int h = warp >> 2;
    const int wl = warp & 3;
    const int row_in_tile = wl * 32 + lane;
    const uint32_t stage_smem = base + SMEM_Q_OFF;  // reused for the O tile
    const uint32_t lane_off = uint32_t(wl * 32) << 16;
    const uint32_t t_s = pin_u32(tmem_base + lane_off + TM_S + h * 64);
    const uint32_t t_o = t_s + (TM_O - TM_S);
    const bool lane0 = pin_u32(lane == 0 ? 1u : 0u) != 0u;
    float m = -INFINITY;  // running reference of this stream (log2 domain)
    unsigned long long lsumA = pack2(0.f, 0.f), lsumB = pack2(0.f, 0.f);
    const uint8_t* mrow = nullptr;
    if (MASK_MODE == 1) {
      const int qrow = q_start + row_in_tile;
      if (qrow < p.Tq) mrow = p.mask + int64_t(b) * p.mask_b_stride + int64_t(qrow) * p.mask_q_stride;
    }

    for (int j = 0; j < n_kv; ++j) {
      mbar_wait_spin(s_full(h), uint32_t(j) & 1u);  // S_h of tile j is in TMEM (and P V_{j-1} has completed)
      tc_fence_after();
      uint32_t xs[64];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld16(t_s + 16 * c, xs + 16 * c);
      tmem_ld_wait();

      const int nv = kv_len - j * KT - h * 64;  // valid keys among my 64 columns
      if (nv < 64) {
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (c >= nv) xs[c] = 0xff800000u;  // -inf
      }
      if (MASK_MODE == 1 && mrow != nullptr) {
        const uint8_t* mp = mrow + j * KT + h * 64;
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (c < nv && mp[c] == 0) xs[c] = __float_as_uint(MASK_FILL_LOG2);
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int c = 0; c < 64; c += 4) {
        mx0 = max3(mx0, __uint_as_float(xs[c]), __uint_as_float(xs[c + 1]));
        mx1 = max3(mx1, __uint_as_float(xs[c + 2]), __uint_as_float(xs[c + 3]));
      }
      const float mnew = fmaxf(m, fmaxf(mx0, mx1));
      const bool need = mnew > m + RESCALE_THRESHOLD;  // (true whenever m is still -inf and the tile has a valid key)
      if (__any_sync(0xffffffffu, need)) {
        // O_h is quiescent here: s_full(j) was committed after P V_{j-1} by the same issuer thread
        const float alpha = need ? ex2_approx(m - mnew) : 1.0f;  // 0 when m was -inf
        if (need) m = mnew;
        if (j > 0) {
          const unsigned long long a2 = pack2(alpha, alpha);
          lsumA = fma2(lsumA, a2, pack2(0.f, 0.f));
          lsumB = fma2(lsumB, a2, pack2(0.f, 0.f));
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
      const float mref = m == -INFINITY ? 0.0f : m;  // a stream that has not seen a valid key yet: p = 2^-inf = 0
      const unsigned long long negm2 = pack2(-mref, -mref);
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const unsigned long long x2 = add2(pack2(__uint_as_float(xs[2 * c]), __uint_as_float(xs[2 * c + 1])), negm2);
        float p0, p1;
        if ((c & 3) < EMU_PAIRS) {
          exp2_emulated2(x2, p0, p1);
        } else {
          float x0, x1;
          unpack2(x2, x0, x1);
          p0 = ex2_approx(x0);
          p1 = ex2_approx(x1);
        }
        xs[2 * c] = __float_as_uint(p0);
        xs[2 * c + 1] = __float_as_uint(p1);
      }
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float p0 = __uint_as_float(xs[2 * c]), p1 = __uint_as_float(xs[2 * c + 1]);
        if (c & 1) lsumB = add2(lsumB, pack2(p0, p1));
        else lsumA = add2(lsumA, pack2(p0, p1));
        pk[c] = pack_bf16x2(p0, p1);
      }
      tmem_st16(t_s, pk);  // P_h over the first 32 columns of S_h (the scores live in registers now)
      tmem_st16(t_s + 16, pk + 16);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane0) mbar_arrive(p_ready(h));
    }

    // ---- epilogue: merge the two streams of each row, normalise, bf16 -> swizzled smem -> TMA store
    float* xch = reinterpret_cast<float*>(smem + SMEM_X_OFF);  // [h][0: m, 1: l][row]
    float a0, a1, b0, b1;
    unpack2(lsumA, a0, a1);
    unpack2(lsumB, b0, b1);
    const float l_mine = (a0 + a1) + (b0 + b1);
    xch[(h * 2 + 0) * 128 + row_in_tile] = m;
    xch[(h * 2 + 1) * 128 + row_in_tile] = l_mine;
    if (n_kv > 0) {
      mbar_wait(pv_done(0), 0);
      mbar_wait(pv_done(1), 0);
      tc_fence_after();
    }
    named_bar_sync(1, 256);
    uint32_t outp[16];
    if (n_kv > 0) {
      const float m0 = xch[0 * 128 + row_in_tile], l0 = xch[1 * 128 + row_in_tile];
      const float m1 = xch[2 * 128 + row_in_tile], l1 = xch[3 * 128 + row_in_tile];
      const float mm = fmaxf(m0, m1);                 // finite: kv_len > 0 puts a valid key into half 0 of tile 0
      const float s0 = ex2_approx(m0 - mm), s1 = m1 == -INFINITY ? 0.0f : ex2_approx(m1 - mm);
      const float inv = 1.0f / (l0 * s0 + l1 * s1);
      const float w0 = s0 * inv, w1 = s1 * inv;
      // this thread produces output columns [32h, 32h + 32) of its row from both accumulators
      uint32_t o0[32], o1[32];
      tmem_ld32(tmem_base + lane_off + TM_O + 0 + 32 * h, o0);
      tmem_ld32(tmem_base + lane_off + TM_O + 64 + 32 * h, o1);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        // an accumulator whose stream never ran a valid tile holds P = 0 products only (finite zeros)
        const float v0 = __uint_as_float(o0[2 * c]) * w0 + __uint_as_float(o1[2 * c]) * w1;
        const float v1 = __uint_as_float(o0[2 * c + 1]) * w0 + __uint_as_float(o1[2 * c + 1]) * w1;
        outp[c] = pack_bf16x2(v0, v1);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 16; ++c) outp[c] = 0u;
    }
    const uint32_t row_addr = stage_smem + uint32_t(row_in_tile) * 128u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t dst = row_addr + (uint32_t((4 * h + i) ^ (row_in_tile & 7)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(outp[4 * i]), "r"(outp[4 * i + 1]),
                   "r"(outp[4 * i + 2]), "r"(outp[4 * i + 3])
                   : "memory");
    }
    fence_proxy_async_smem();
    named_bar_sync(2, 256);
    if (warp == 0 && lane == 0) {
      tma_store_3d(&tmO, stage_smem, head * HD, q_start, b);
      tma_store_commit();
      tma_store_wait_all<0>();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP0) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

template <int MASK_MODE, int EMU_PAIRS>
int launch_variant(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                   const CUtensorMap& tmO, const FmhaParams& p, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    RP_CUDA_CHECK(cudaFuncSetAttribute(fmha3_fwd_kernel<MASK_MODE, EMU_PAIRS>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    configured = true;
  }
  dim3 grid((p.Tq + QT - 1) / QT, p.H, p.B);
  RP_CUDA_CHECK(launch_pdl(fmha3_fwd_kernel<MASK_MODE, EMU_PAIRS>, grid, dim3(NUM_THREADS), SMEM_TOTAL, stream, tmQ,
                           tmK, tmV, tmO, p));
  count_launch();
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

}  // namespace

int launch_fmha3(const FmhaArgs& a, cudaStream_t stream) {
  const uint64_t cols = uint64_t(a.H) * HD;
  CUtensorMap tmQ, tmK, tmV, tmO;
  int rc;
  const CUtensorMapDataType bf = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  if ((rc = make_tmap_3d(&tmQ, bf, a.q, cols, a.Tq, a.B, a.ldq * 2, a.bsq * 2, HD, QT))) return rc;
  if ((rc = make_tmap_3d(&tmK, bf, a.k, cols, a.Tk, a.B, a.ldk * 2, a.bsk * 2, HD, KT))) return rc;
  if ((rc = make_tmap_3d(&tmV, bf, a.v, cols, a.Tk, a.B, a.ldv * 2, a.bsv * 2, HD, KT))) return rc;
  if ((rc = make_tmap_3d(&tmO, bf, a.o, cols, a.Tq, a.B, a.ldo * 2, a.bso * 2, HD, QT))) return rc;
  FmhaParams p{a.B, a.H, a.Tq, a.Tk, a.kv_lens, a.mask, a.mask_b_stride, a.mask_q_stride};
  static const int emu = getenv("RP_FMHA_EMU") ? atoi(getenv("RP_FMHA_EMU")) : 1;
  if (a.mask_mode == 1) return launch_variant<1, 1>(tmQ, tmK, tmV, tmO, p, stream);
  switch (emu) {
    case 0: return launch_variant<0, 0>(tmQ, tmK, tmV, tmO, p, stream);
    case 2: return launch_variant<0, 2>(tmQ, tmK, tmV, tmO, p, stream);
    default: return launch_variant<0, 1>(tmQ, tmK, tmV, tmO, p, stream);
  }
}

}  // namespace rp
