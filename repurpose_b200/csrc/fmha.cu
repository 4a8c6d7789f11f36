// Fused multi-head attention forward for sm_100a, head dim 64, bf16 operands, fp32 softmax:
// one score row per thread, software-pipelined softmax.
//
// One CTA = one (batch, head, NQ 128-row query tiles).  Roles ((4 NQ + 4) warps):
//   warps 0 .. 4NQ-1 : softmax; warp w owns TMEM lanes 32 (w & 3) .. +31 of query tile w >> 2 and
//                      thread <-> lane <-> query row: the whole 128-key score row of a key tile
//                      lives in one thread (no cross-thread max / sum exchange, no barriers)
//   warp 4NQ         : TMA producer (Q once, K / V tiles through rings, SWIZZLE_128B)
//   warp 4NQ + 1     : tcgen05.mma issuer for S = Q K^T + TMEM owner
//   warp 4NQ + 2     : tcgen05.mma issuer for O += P V      (warp 4NQ+3 only completes the warpgroup)
// TMEM per query tile: S 128 columns | P 64 (bf16 pairs) | O 64; NQ = 1 -> 256 columns, two CTAs per SM.
//
// Pipeline.  Both MMAs of a key tile are issued as two 64-key halves with their own barriers:
//   S_h = Q K_h^T (SS-mode, N = 64)  -> s_full[h];   O += P_h V_h (A = P from TMEM) -> pv_done[h].
// A softmax thread keeps the 128 scores of tile j in registers and walks them in eight 16-score
// chunks.  Step s of an iteration does
//   E(s)   exp2 of chunk s (MUFU, with EMU_PAIRS of every 4 pairs evaluated on the FMA pipe),
//   D(s-1) row-sum + bf16 pack of chunk s-1, tcgen05.st of its 8 P columns, and tcgen05.ld of chunk
//          s-1 of the NEXT tile's scores into the registers that just became free,
//   M(s-2) running max of the next tile's chunk s-2 (masking of a partial last tile happens once, when
//          the whole tile is in registers),
// so the next tile's row max is known when the iteration ends and S is single-buffered in TMEM but
// double-buffered through registers.  Half h of S / P is handed back to the MMA warps in the middle /
// at the end of the iteration, which gives QK^T_{j+2,h} and P V_{j,h} more than half an iteration to
// complete before the softmax threads need their result; barrier probes (mbarrier.test_wait) are
// issued a step before their answer is needed.  The loop always prefetches "a next tile" (a dummy
// re-run of QK^T after the last one), so the iteration body is one basic block: ptxas front-loads the
// exp2 work and interleaves the rest, which measured faster than forcing a per-chunk schedule with
// opaque branches (profiles/r01_notes.md).
//
// Softmax arithmetic (per score): packed f32x2 subtract of the running reference, exp2 either on the
// MUFU (ex2.approx) or — for EMU_PAIRS out of every 4 pairs — by a Cody-Waite split plus a degree-3
// polynomial on the FMA pipe (max relative error 7.5e-5, far below bf16 resolution), packed f32x2
// row sum, bf16x2 pack.  The reference is only refreshed when the row max grows by more than 2^8
// (lazy rescale), so the O correction in TMEM is rare.
//
// Semantics follow the reference:
//   mask_mode 0  nn.MultiheadAttention with key_padding_mask (models/MMCTransformer.py:132-138):
//                keys >= kv_lens[b] receive -inf; every query row (padded or not) is computed.
//   mask_mode 1  models/transformer.py:52-81 MultiHeadAttention.forward: masked_fill(mask==0, -1e9).
// Q must arrive pre-scaled by log2(e)/sqrt(64) so that S is already in the exp2 domain.
#include <math.h>
#include <stdlib.h>
#include <type_traits>

#include "ptx.cuh"
#include "host_util.h"
#include "kernels.h"

namespace rp {

namespace {

constexpr int QT = 128;
constexpr int KT = 128;
constexpr int HD = 64;
constexpr int TILE_BYTES = QT * HD * 2;  // 16 KB (Q, K and V tiles are all 128 x 64 bf16)

// K ring depth of the two-CTAs-per-SM kernel.  K_{i+1} is requested when Q K^T_{i-1} retires and needed about 1.5
// iterations (~3900 cycles) later; a TMA tile that misses L2 takes ~4000 cycles under load, so two stages leave no
// slack.  The third stage only fits next to a second CTA when the alignment slack of the dynamic segment shrinks to
// 256 bytes.  Measured (profiles/r02_notes.md 1.5): +0.2 .. 0.4 % on the kernel alone at T = 1801 / 1792 / 8192.
#ifndef RP_FMHA_KSTAGES
#define RP_FMHA_KSTAGES 3
#endif
template <int NQ>
struct Cfg {
  static constexpr int K_STAGES = NQ == 2 ? 3 : RP_FMHA_KSTAGES;
  static constexpr int V_STAGES = 3;  // V_j is needed 1.5 iterations after its slot frees with 2 stages: not enough for the TMA latency
  static constexpr int SMEM_Q_OFF = 0;
  static constexpr int SMEM_K_OFF = NQ * TILE_BYTES;
  static constexpr int SMEM_V_OFF = SMEM_K_OFF + K_STAGES * TILE_BYTES;
  static constexpr int SMEM_BAR_OFF = SMEM_V_OFF + V_STAGES * TILE_BYTES;
  // barriers + TMEM slot: 256 bytes; alignment slack of the (1024-aligned, checked at run time) dynamic segment
  static constexpr int SMEM_SLACK = (NQ == 1 && K_STAGES == 3) ? 256 : 1024;
  static constexpr int SMEM_TOTAL = SMEM_BAR_OFF + 256 + SMEM_SLACK;
  static constexpr int SOFTMAX_WARPS = 4 * NQ;
  static constexpr int PRODUCER_WARP = 4 * NQ;
  static constexpr int MMA_WARP = 4 * NQ + 1;
  static constexpr int PV_WARP = 4 * NQ + 2;
  static constexpr int NUM_THREADS = (4 * NQ + 4) * 32;  // softmax warpgroups + one auxiliary warpgroup
  static constexpr int MIN_CTAS = NQ == 2 ? 1 : 2;
  static constexpr int TMEM_COLS = 256 * NQ;
  static constexpr int TM_S = 0;         // + q*128
  static constexpr int TM_P = 128 * NQ;  // + q*64
  static constexpr int TM_O = 192 * NQ;  // + q*64
  // register re-allocation after the role split (setmaxnreg is a warpgroup-wide operation):
  // NQ = 1: 256 threads x 128 at launch -> 4 x 32 x 216 + 4 x 32 x 40;  NQ = 2: 384 x 168 -> 8 x 32 x 232 + 4 x 32 x 40
  static constexpr int SOFTMAX_REGS = NQ == 2 ? 232 : 216;
  static constexpr int AUX_REGS = 40;
};
#ifndef RP_FMHA_TAIL_DEFER
#define RP_FMHA_TAIL_DEFER 96
#endif
#ifndef RP_FMHA_CHUNK
#define RP_FMHA_CHUNK 16
#endif
constexpr int CH = RP_FMHA_CHUNK;  // scores per pipeline step (16 or 32)
constexpr int NCH = KT / CH;
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units; stale max keeps p <= 2^8
constexpr float MASK_FILL_LOG2 = -1.0e9f * 1.4426950408889634f;

#ifndef RP_TRACE_Z
#define RP_TRACE_Z 1
#endif
#ifdef RP_FMHA_TRACE
__device__ unsigned long long g_fmha_trace[8 * 512];  // [role][event] = clock64
#define TRACE(role, idx)                                                              \
  do {                                                                                \
    if (blockIdx.x == 2 + 15 * (3 + 8 * RP_TRACE_Z) && (idx) < 512)                    \
      g_fmha_trace[(role) * 512 + (idx)] = clock64();                                 \
  } while (0)
#else
#define TRACE(role, idx) do {} while (0)
#endif

struct FmhaParams {
  int B, H, Tq, Tk;
  const int32_t* kv_lens;
  const uint8_t* mask;
  int64_t mask_b_stride, mask_q_stride;
  float* lse;  // optional [B, H, Tq]: log2-domain log-sum-exp of every score row (what the backward pass recomputes P from)
  int skip_padded_queries;  // query tiles at or beyond round_up(kv_len, 128) are padding: exit at once
  // DROP kernels (training forward): keep bits of the attention weights, row (b*H + h)*Tq + q, drop_ld words per row
  const uint32_t* drop_bits;
  int64_t drop_ld;
  float drop_scale;
  const int32_t* batch_order;  // optional [B]: block index -> batch element (longest first)
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

// exp2 of two fp32 values on the FMA/ALU pipes: x = n + f, n = rint(x), f in [-0.5, 0.5];
// 2^f by a degree-3 minimax polynomial; 2^n by adding n to the exponent field.
__device__ __forceinline__ void exp2_emulated2(unsigned long long x2, float& p0, float& p1) {
  float x0, x1;
  unpack2(x2, x0, x1);
  x0 = fmaxf(x0, -126.0f);  // keeps the exponent arithmetic in range; 2^-126 rounds to 0 in bf16 sums
  x1 = fmaxf(x1, -126.0f);
  const unsigned long long xc = pack2(x0, x1);
  const unsigned long long magic = pack2(12582912.0f, 12582912.0f);  // 1.5 * 2^23
  const unsigned long long nmagic = pack2(-12582912.0f, -12582912.0f);
  const unsigned long long t = add2(xc, magic);  // low mantissa bits now hold rint(x)
  const unsigned long long n = add2(t, nmagic);
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

// DROP: nn.MultiheadAttention's dropout on the attention weights (train mode): the row sum l runs over every weight, the
// P tile handed to P V holds the kept ones only, and the epilogue scales by 1 / (1 - p) — out = (keep o softmax(S)) V / (1 - p)
template <int MASK_MODE, int EMU_PAIRS, int NQ, bool DROP = false>
__global__ void __launch_bounds__(Cfg<NQ>::NUM_THREADS, Cfg<NQ>::MIN_CTAS)
fmha_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                 const FmhaParams p) {
  using C = Cfg<NQ>;
  constexpr int K_STAGES = C::K_STAGES, V_STAGES = C::V_STAGES;
  constexpr int SMEM_Q_OFF = C::SMEM_Q_OFF, SMEM_K_OFF = C::SMEM_K_OFF, SMEM_V_OFF = C::SMEM_V_OFF;
  constexpr int SMEM_BAR_OFF = C::SMEM_BAR_OFF;
  constexpr int TM_S = C::TM_S, TM_P = C::TM_P, TM_O = C::TM_O, TMEM_COLS = C::TMEM_COLS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  if (base - raw_addr > uint32_t(C::SMEM_SLACK)) __trap();  // the budget leaves SMEM_SLACK bytes of alignment slack

  const uint32_t bar_base = pin_u32(base + SMEM_BAR_OFF);
  auto q_full = [&](int q) { return bar_base + 8u * q; };
  auto k_full = [&](int s) { return bar_base + 16u + 8u * s; };
  auto k_empty = [&](int s) { return bar_base + 40u + 8u * s; };
  auto v_full = [&](int s) { return bar_base + 64u + 8u * s; };
  auto v_empty = [&](int s) { return bar_base + 88u + 8u * s; };
  auto s_full = [&](int q, int h) { return bar_base + 112u + 8u * (2 * q + h); };
  auto s_free = [&](int q, int h) { return bar_base + 144u + 8u * (2 * q + h); };
  auto p_ready = [&](int q, int h) { return bar_base + 176u + 8u * (2 * q + h); };
  auto pv_done = [&](int q, int h) { return bar_base + 208u + 8u * (2 * q + h); };
  constexpr int TMEM_SLOT_OFF = 240;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SMEM_BAR_OFF + TMEM_SLOT_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // 1-D grid over (batch, head, query block), (batch, head) major.  A (batch, head)'s LAST query block is the only one that can
  // be partial, i.e. a short CTA; the short CTAs of the last RP_FMHA_TAIL_DEFER (batch, head) pairs are moved to the end of the
  // launch, where they fill its ragged end.  Only those: a deferred CTA re-reads its K / V, which is still in L2 for the
  // pairs that ran last but long evicted for the early ones (all 256 deferred: +116 MB of DRAM reads per launch at
  // B = 32, T = 1801 — profiles/r02_notes.md 9).
  const int nqb = (p.Tq + NQ * QT - 1) / (NQ * QT);
  const int n_bh = p.H * p.B;
  const int n_def = nqb > 1 ? (n_bh < RP_FMHA_TAIL_DEFER ? n_bh : RP_FMHA_TAIL_DEFER) : 0;
  const int n_first = (n_bh - n_def) * nqb;       // in order, tails included
  const int n_second = n_def * (nqb - 1);         // the last pairs' full blocks
  int qb, bh;
  {
    int x = int(blockIdx.x);
    if (x < n_first) {
      qb = x % nqb;
      bh = x / nqb;
    } else if (x - n_first < n_second) {
      x -= n_first;
      qb = x % (nqb - 1);
      bh = (n_bh - n_def) + x / (nqb - 1);
    } else {
      qb = nqb - 1;
      bh = (n_bh - n_def) + (x - n_first - n_second);
    }
  }
  const int head = bh % p.H;
  const int b = p.batch_order != nullptr ? p.batch_order[bh / p.H] : bh / p.H;

  const int q_start0 = qb * (NQ * QT);
  const bool q1_active = NQ == 2 && (q_start0 + QT) < p.Tq;
  const int nq = q1_active ? 2 : 1;
  int kv_len = p.Tk;
  if (MASK_MODE == 0 && p.kv_lens != nullptr) {
    kv_len = p.kv_lens[b];
    kv_len = kv_len < 0 ? 0 : (kv_len > p.Tk ? p.Tk : kv_len);
  }
  const int n_kv = (kv_len + KT - 1) / KT;
  // self-attention over a padded batch: a query tile that starts beyond the video's last (rounded-up) key tile holds
  // padding rows only — nobody reads its output (FmhaArgs::skip_padded_queries; block-uniform, before any barrier)
  if (MASK_MODE == 0 && p.skip_padded_queries && q_start0 >= n_kv * KT) return;
  const int nv_last = kv_len - (n_kv - 1) * KT;  // valid keys of the last tile (1 .. KT)
  const int nch_last = (nv_last + 15) >> 4;      // its 16-key chunks: the only ones multiplied and exponentiated
  // softmax warps of query tile q with at least one row below Tq (the others exit at once)
  auto warps_of = [&](int q) {
    const int r = p.Tq - (q_start0 + q * QT);
    return r <= 0 ? 0 : (r >= QT ? 4 : (r + 31) >> 5);
  };

  if (warp == C::PRODUCER_WARP && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    for (int q = 0; q < 2; ++q) {
      mbar_init(q_full(q), 1);
      const int nw = warps_of(q) > 0 ? warps_of(q) : 1;
      for (int h = 0; h < 2; ++h) {
        mbar_init(s_full(q, h), 1);
        mbar_init(s_free(q, h), nw);
        mbar_init(p_ready(q, h), nw);
        mbar_init(pv_done(q, h), 1);
      }
    }
    for (int s = 0; s < 3; ++s) {
      mbar_init(k_full(s), 1);
      mbar_init(k_empty(s), 1);
      mbar_init(v_full(s), 1);
      mbar_init(v_empty(s), 1);
    }
    fence_mbar_init();
  }
  if (warp == C::MMA_WARP) tmem_alloc<TMEM_COLS>(base + SMEM_BAR_OFF + TMEM_SLOT_OFF);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // the set-up above overlapped the previous kernel's tail

  if (warp >= C::SOFTMAX_WARPS) setmaxnreg_dec<C::AUX_REGS>();
  if (warp == C::PRODUCER_WARP) {
    // ---------------------------------------------------------------- TMA producer
    // K runs one tile ahead of V: K_{i+1} is needed (for S_{i+1}) long before V_i's ring slot frees.
    if (n_kv > 0) {
      const int col = head * HD;
      if (elect_one()) {
        mbar_expect_tx(q_full(0), TILE_BYTES);
        tma_load_3d(base + SMEM_Q_OFF, &tmQ, q_full(0), col, q_start0, b);
        if (q1_active) {
          mbar_expect_tx(q_full(1), TILE_BYTES);
          tma_load_3d(base + SMEM_Q_OFF + TILE_BYTES, &tmQ, q_full(1), col, q_start0 + QT, b);
        }
      }
      __syncwarp();
      auto load_k = [&](int i) {
        const int st = i % K_STAGES;
        mbar_wait(k_empty(st), (uint32_t(i / K_STAGES) & 1u) ^ 1u);
        if (lane == 0) TRACE(7, 8 * i + 4);
        if (elect_one()) {
          mbar_expect_tx(k_full(st), TILE_BYTES);
          tma_load_3d(base + SMEM_K_OFF + st * TILE_BYTES, &tmK, k_full(st), col, i * KT, b);
        }
        __syncwarp();
      };
      auto load_v = [&](int i) {
        const int st = i % V_STAGES;
        mbar_wait(v_empty(st), (uint32_t(i / V_STAGES) & 1u) ^ 1u);
        if (lane == 0) TRACE(7, 8 * i + 5);
        if (elect_one()) {
          mbar_expect_tx(v_full(st), TILE_BYTES);
          tma_load_3d(base + SMEM_V_OFF + st * TILE_BYTES, &tmV, v_full(st), col, i * KT, b);
        }
        __syncwarp();
      };
      load_k(0);
      for (int i = 0; i < n_kv; ++i) {
        if (i + 1 < n_kv) load_k(i + 1);
        load_v(i);
      }
    }
  } else if (warp == C::MMA_WARP) {
    // ---------------------------------------------------------------- QK^T issuer (and TMEM owner)
    // Converged warp, one elected lane per issue group (descriptors stay in uniform registers).
    // S = Q K^T and O += P V are issued from two different warps so that neither sits behind the
    // other's barrier waits: the softmax hands S_h and P_h back at the same moment and needs both
    // results half an iteration later.
    if (n_kv > 0) {
      // S_q[:, 64h .. 64h+ncols) = Q_q K[64h .. 64h+ncols)^T; ncols = 0 (a half of the last tile without
      // valid keys): nothing to multiply, the commit alone completes the barrier phase
      auto issue_qk = [&](int q, int st, int h, int ncols, uint32_t commit_bar, uint32_t commit_bar2) {
        if (elect_one()) {
          if (ncols > 0) {
            const uint32_t idesc_s = make_idesc_bf16(QT, 0, false, false) | (uint32_t(ncols >> 3) << 17);
            const uint64_t da = make_smem_desc_sw128(base + SMEM_Q_OFF + q * TILE_BYTES, 1024, 16);
            const uint64_t db =
                make_smem_desc_sw128(base + SMEM_K_OFF + st * TILE_BYTES + h * (TILE_BYTES / 2), 1024, 16);
#pragma unroll
            for (int k = 0; k < HD / 16; ++k)
              mma_ss(tmem_base + TM_S + q * 128 + h * 64, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc_s,
                     k > 0 ? 1u : 0u);
          }
          tc_commit(commit_bar);
          if (commit_bar2 != 0) tc_commit(commit_bar2);
        }
        __syncwarp();
      };
      for (int q = 0; q < nq; ++q) mbar_wait(q_full(q), 0);
      mbar_wait(k_full(0), 0);
      tc_fence_after();
      auto ncols_of = [&](int i, int h) {
        if (i + 1 < n_kv) return KT / 2;
        const int c = nch_last * 16 - (KT / 2) * h;
        return c < 0 ? 0 : (c > KT / 2 ? KT / 2 : c);
      };
      for (int h = 0; h < 2; ++h)
        for (int q = 0; q < nq; ++q)
          issue_qk(q, 0, h, ncols_of(0, h), s_full(q, h), (h == 1 && q == nq - 1) ? k_empty(0) : 0u);
      // S_i may be issued once the softmax warps have read S_{i-1} (s_free phase i-1).  The last tile is
      // only multiplied over its valid 16-key chunks.
      for (int i = 1; i < n_kv; ++i) {
        const int st = i % K_STAGES;
        mbar_wait(k_full(st), uint32_t(i / K_STAGES) & 1u);
        for (int h = 0; h < 2; ++h)
          for (int q = 0; q < nq; ++q) {
            mbar_wait(s_free(q, h), uint32_t(i - 1) & 1u);
            tc_fence_after();
            if (lane == 0) TRACE(2 + h, i);
            issue_qk(q, st, h, ncols_of(i, h), s_full(q, h), (h == 1 && q == nq - 1) ? k_empty(st) : 0u);
          }
      }
    }
  } else if (warp == C::PV_WARP) {
    // ---------------------------------------------------------------- P V issuer
    if (n_kv > 0) {
      constexpr uint32_t idesc_o = make_idesc_bf16(QT, HD, false, true);  // V is MN-major
      // O_q += P_q[:, 64h .. 64h+64) V[64h .. 64h+64)
      auto issue_pv = [&](int q, int st, int h, bool acc, int ksteps, uint32_t commit_bar, uint32_t commit_bar2) {
        if (elect_one()) {
          // A: P in TMEM, 16 keys = 8 packed columns per step; B: 16 key rows of 128 B each
          const uint64_t db = make_smem_desc_sw128(base + SMEM_V_OFF + st * TILE_BYTES, 1024, 1024);
#pragma unroll
          for (int k = 0; k < KT / 32; ++k)
            if (k < ksteps)
              mma_ts(tmem_base + TM_O + q * 64, tmem_base + TM_P + q * 64 + h * 32 + k * 8,
                     db + uint64_t(128 * (4 * h + k)), idesc_o, (acc || k > 0) ? 1u : 0u);
          tc_commit(commit_bar);
          if (commit_bar2 != 0) tc_commit(commit_bar2);
        }
        __syncwarp();
      };
      for (int j = 0; j < n_kv; ++j) {
        const int vst = j % V_STAGES;
        mbar_wait(v_full(vst), uint32_t(j / V_STAGES) & 1u);
        for (int h = 0; h < 2; ++h)
          for (int q = 0; q < nq; ++q) {
            mbar_wait(p_ready(q, h), uint32_t(j) & 1u);
            tc_fence_after();
            if (lane == 0) TRACE(h, j);
            const int c = j + 1 < n_kv ? 4 : nch_last - 4 * h;  // 16-key steps of this half (last tile: valid ones)
            issue_pv(q, vst, h, j > 0 || h > 0, c < 0 ? 0 : (c > 4 ? 4 : c), pv_done(q, h),
                     (h == 1 && q == nq - 1) ? v_empty(vst) : 0u);
          }
      }
    }
  } else if (warp < C::SOFTMAX_WARPS) {
    setmaxnreg_inc<C::SOFTMAX_REGS>();
    // ---------------------------------------------------------------- softmax warps
    const int q = warp >> 2;
    const int wl = warp & 3;
    const int row_in_tile = wl * 32 + lane;
    const int q_start = q_start0 + q * QT;
    const uint32_t stage_smem = base + SMEM_Q_OFF + q * TILE_BYTES;  // reused for the O tile
    if (wl < warps_of(q)) {
      const uint32_t lane_off = uint32_t(wl * 32) << 16;
      const uint32_t t_s = pin_u32(tmem_base + lane_off + TM_S + q * 128);
      const uint32_t t_p = t_s + (TM_P - TM_S) - q * 64;
      const uint32_t t_o = t_s + (TM_O - TM_S) - q * 64;
      const bool tracer = wl == 0 && lane == 0 && q == 0;
      float m = -INFINITY;
      unsigned long long lsumA = pack2(0.f, 0.f), lsumB = pack2(0.f, 0.f);
      const uint8_t* mrow = nullptr;
      if (MASK_MODE == 1) {
        const int qrow = q_start + row_in_tile;
        if (qrow < p.Tq) mrow = p.mask + int64_t(b) * p.mask_b_stride + int64_t(qrow) * p.mask_q_stride;
      }
      uint32_t xs[128];  // scores of the current tile -> probabilities -> scores of the next tile
      const bool lane0 = pin_u32(lane == 0 ? 1u : 0u) != 0u;
      // DROP: the 128 keep bits of this row for the current key tile (4 words), fetched one tile ahead
      const uint4* brow = nullptr;
      uint4 wcur = make_uint4(~0u, ~0u, ~0u, ~0u);
      if constexpr (DROP) {
        const int qrow = q_start + row_in_tile < p.Tq ? q_start + row_in_tile : p.Tq - 1;
        brow = reinterpret_cast<const uint4*>(p.drop_bits + ((int64_t(b) * p.H + head) * p.Tq + qrow) * p.drop_ld);
        if (n_kv > 0) wcur = __ldg(brow);
      }
      // the bf16 pair (p0, p1) = keys 2 * pair, 2 * pair + 1 of the tile with the dropped weights zeroed
      auto pack_kept = [&](int pair, float p0, float p1) {
        if constexpr (DROP) {
          const uint32_t w = pair < 16 ? wcur.x : (pair < 32 ? wcur.y : (pair < 48 ? wcur.z : wcur.w));
          const int bit = (2 * pair) & 31;
          p0 = (w & (1u << bit)) ? p0 : 0.0f;
          p1 = (w & (2u << bit)) ? p1 : 0.0f;
        }
        return pack_bf16x2(p0, p1);
      };

      // running max of chunk c (CH scores)
      auto max_chunk = [&](int c, float& mx0, float& mx1) {
#pragma unroll
        for (int i = 0; i < CH; i += 4) {
          mx0 = max3(mx0, __uint_as_float(xs[CH * c + i]), __uint_as_float(xs[CH * c + i + 1]));
          mx1 = max3(mx1, __uint_as_float(xs[CH * c + i + 2]), __uint_as_float(xs[CH * c + i + 3]));
        }
      };
      // Masking is needed for the last (partial) key tile only — and for every tile with an explicit
      // mask: it runs once the whole tile is in registers and recomputes the row max from scratch.
      auto tile_needs_mask = [&](int tile) { return MASK_MODE == 1 || kv_len - tile * KT < KT; };
      auto mask_tile = [&](int tile, float& mx0, float& mx1) {
        const int nv = kv_len - tile * KT;  // valid keys in this tile
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= nv) xs[i] = 0xff800000u;  // -inf
        if (MASK_MODE == 1 && mrow != nullptr) {
          const uint8_t* mp = mrow + tile * KT;
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (i < nv && mp[i] == 0) xs[i] = __float_as_uint(MASK_FILL_LOG2);
        }
        mx0 = mx1 = -INFINITY;
#pragma unroll
        for (int c = 0; c < NCH; ++c) max_chunk(c, mx0, mx1);
      };

      if (n_kv > 0) {
        // ---- prologue: S_0 -> registers, hand both halves back, row max
        mbar_wait(s_full(q, 0), 0);
        mbar_wait(s_full(q, 1), 0);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 8; ++c) tmem_ld16(t_s + 16 * c, xs + 16 * c);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(s_free(q, 0));
          mbar_arrive(s_free(q, 1));
        }
        float mx0 = -INFINITY, mx1 = -INFINITY;
        if (tile_needs_mask(0)) {
          mask_tile(0, mx0, mx1);
        } else {
#pragma unroll
          for (int c = 0; c < NCH; ++c) max_chunk(c, mx0, mx1);
        }
        m = fmaxf(mx0, mx1);
      }

      // ---- tiles 0 .. n_kv-2: each one prefetches the scores of its successor
      for (int j = 0; j + 1 < n_kv; ++j) {
        const uint32_t par = uint32_t(j) & 1u, par_next = par ^ 1u;
        const unsigned long long negm2 = pack2(-m, -m);
        float mx0 = -INFINITY, mx1 = -INFINITY;
        uint32_t probe_pv = 0, probe_s = 0;
        uint4 wnext = wcur;
        if constexpr (DROP) wnext = __ldg(brow + (j + 1));
        if (tracer) TRACE(4, 2 * j);
#pragma unroll
        for (int s = 0; s <= NCH; ++s) {
          // Half h of P may be overwritten once P V_{j-1,h} has completed, and half h of the next
          // tile's scores (a dummy after the last tile) read once S_{j+1,h} is in TMEM: waited for
          // ahead of the step that first touches them.
          if (s == 1 || s == NCH / 2 + 1) {
            // (probed one step ago: the answer is normally "complete" and already in a register)
            if (!probe_pv) mbar_wait_spin(pv_done(q, s > 1), par_next);
            if (!probe_s) mbar_wait_spin(s_full(q, s > 1), par_next);
            tc_fence_after();
          }
          {
            if (tracer) TRACE(5, 16 * j + s);
            if (s == 0 || s == NCH / 2) {
              probe_pv = j > 0 ? mbar_test_wait(pv_done(q, s > 0), par_next) : 1u;
              probe_s = mbar_test_wait(s_full(q, s > 0), par_next);
            }
            // ---- E(s): exp2 of chunk s, in place
            if (s < NCH) {
#pragma unroll
              for (int i = 0; i < CH / 2; ++i) {
                const int c0 = CH * s + 2 * i;
                const unsigned long long x2 =
                    add2(pack2(__uint_as_float(xs[c0]), __uint_as_float(xs[c0 + 1])), negm2);
                float p0, p1;
                if ((i & 3) < EMU_PAIRS) {
                  exp2_emulated2(x2, p0, p1);
                } else {
                  float x0, x1;
                  unpack2(x2, x0, x1);
                  p0 = ex2_approx(x0);
                  p1 = ex2_approx(x1);
                }
                xs[c0] = __float_as_uint(p0);
                xs[c0 + 1] = __float_as_uint(p1);
              }
            }
            // ---- M(s-2): the next tile's chunk s-2 has landed
            if (s >= 2) {
              tmem_ld_wait();
              max_chunk(s - 2, mx0, mx1);
            }
            if (s == NCH / 2 + 1) {
              // first halves are complete: P columns [0,32) stored (step NCH/2) and the next tile's
              // score columns [0,64) read (loads of steps 1..NCH/2, waited for above)
              tmem_st_wait();
              tc_fence_before();
              __syncwarp();
              if (lane0) {
                mbar_arrive(p_ready(q, 0));
                mbar_arrive(s_free(q, 0));
              }
            }
            // ---- D(s-1): consume chunk s-1 (row sum, bf16 pack, P store), refill its registers
            if (s >= 1) {
              const int c = s - 1;
              uint32_t pk[CH / 2];
#pragma unroll
              for (int i = 0; i < CH / 2; ++i) {
                const float p0 = __uint_as_float(xs[CH * c + 2 * i]), p1 = __uint_as_float(xs[CH * c + 2 * i + 1]);
                if (i & 1) lsumB = add2(lsumB, pack2(p0, p1));
                else lsumA = add2(lsumA, pack2(p0, p1));
                pk[i] = pack_kept((CH / 2) * c + i, p0, p1);
              }
              if (CH == 16) {
                tmem_st8(t_p + (CH / 2) * c, pk);
                tmem_ld16(t_s + CH * c, xs + CH * c);
              } else {
                tmem_st16(t_p + (CH / 2) * c, pk);
                tmem_ld32(t_s + CH * c, xs + CH * c);
              }
            }
          }
        }
        if (tracer) TRACE(4, 2 * j + 1);
        // ---- tail: second halves
        tmem_st_wait();
        tmem_ld_wait();
        max_chunk(NCH - 1, mx0, mx1);
        tc_fence_before();
        __syncwarp();
        if (lane0) {
          mbar_arrive(p_ready(q, 1));
          mbar_arrive(s_free(q, 1));
        }
        if constexpr (DROP) wcur = wnext;
        {
          if (tile_needs_mask(j + 1)) mask_tile(j + 1, mx0, mx1);
          const float mnext = fmaxf(m, fmaxf(mx0, mx1));
          const bool need = mnext > m + RESCALE_THRESHOLD;
          if (__any_sync(0xffffffffu, need)) {
            // rare: the row max grew by more than 2^8 — O and l move to the new reference once
            // P V_j (just handed over) has completed
            mbar_wait(pv_done(q, 0), par);
            mbar_wait(pv_done(q, 1), par);
            tc_fence_after();
            const float alpha = need ? ex2_approx(m - mnext) : 1.0f;
            if (need) m = mnext;
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
            tmem_st_wait();
            tc_fence_before();
          }
        }
      }

      // ---- last tile: only its valid 16-key chunks, nothing to prefetch (a T = 1801 sequence has 9 keys
      // in its 15th tile)
      if (n_kv > 0) {
        const int j = n_kv - 1;
        const uint32_t par_next = (uint32_t(j) & 1u) ^ 1u;
        const unsigned long long negm2 = pack2(-m, -m);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          if (c == 0 || c == NCH / 2) {  // P half may be overwritten once P V_{j-1} of that half is done
            if (j > 0) mbar_wait(pv_done(q, c > 0), par_next);
            tc_fence_after();
          }
          if (c < nch_last) {
            uint32_t pk[CH / 2];
#pragma unroll
            for (int i = 0; i < CH / 2; ++i) {
              const int c0 = CH * c + 2 * i;
              const unsigned long long x2 =
                  add2(pack2(__uint_as_float(xs[c0]), __uint_as_float(xs[c0 + 1])), negm2);
              float p0, p1;
              if ((i & 3) < EMU_PAIRS) {
                exp2_emulated2(x2, p0, p1);
              } else {
                float x0, x1;
                unpack2(x2, x0, x1);
                p0 = ex2_approx(x0);
                p1 = ex2_approx(x1);
              }
              if (i & 1) lsumB = add2(lsumB, pack2(p0, p1));
              else lsumA = add2(lsumA, pack2(p0, p1));
              pk[i] = pack_kept((CH / 2) * c + i, p0, p1);
            }
            if (CH == 16) tmem_st8(t_p + (CH / 2) * c, pk);
            else tmem_st16(t_p + (CH / 2) * c, pk);
          }
          if (c == NCH / 2 - 1 || c == NCH - 1) {
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane0) mbar_arrive(p_ready(q, c >= NCH / 2));
          }
        }
      }

      // ---- epilogue: O / l -> bf16 -> swizzled smem (the Q tile's slot) -> TMA store
      const uint32_t row_addr = stage_smem + uint32_t(row_in_tile) * 128u;
      float inv_l = 0.0f;
      if (n_kv == 0 && p.lse != nullptr && q_start + row_in_tile < p.Tq)
        p.lse[(int64_t(b) * p.H + head) * p.Tq + q_start + row_in_tile] = INFINITY;  // no keys: P = 0 in the backward pass
      if (n_kv > 0) {
        float a0, a1, b0, b1;
        unpack2(lsumA, a0, a1);
        unpack2(lsumB, b0, b1);
        const float l_row = (a0 + a1) + (b0 + b1);
        inv_l = 1.0f / l_row;
        if constexpr (DROP) inv_l *= p.drop_scale;
        if (p.lse != nullptr && q_start + row_in_tile < p.Tq)
          p.lse[(int64_t(b) * p.H + head) * p.Tq + q_start + row_in_tile] = m + log2f(l_row);
        mbar_wait(pv_done(q, 0), uint32_t(n_kv - 1) & 1u);
        mbar_wait(pv_done(q, 1), uint32_t(n_kv - 1) & 1u);
        tc_fence_after();
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t o[32];
        if (n_kv > 0) {
          tmem_ld32(t_o + 32 * half, o);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c) o[c] = 0u;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t p0 = pack_bf16x2(__uint_as_float(o[8 * i + 0]) * inv_l, __uint_as_float(o[8 * i + 1]) * inv_l);
          const uint32_t p1 = pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv_l, __uint_as_float(o[8 * i + 3]) * inv_l);
          const uint32_t p2 = pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv_l, __uint_as_float(o[8 * i + 5]) * inv_l);
          const uint32_t p3 = pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv_l, __uint_as_float(o[8 * i + 7]) * inv_l);
          const uint32_t dst = row_addr + (uint32_t((4 * half + i) ^ (row_in_tile & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(p0), "r"(p1), "r"(p2),
                       "r"(p3)
                       : "memory");
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + q, 32 * warps_of(q));
      if (wl == 0 && lane == 0) {
        tma_store_3d(&tmO, stage_smem, head * HD, q_start, b);
        tma_store_commit();
        tma_store_wait_all<0>();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == C::MMA_WARP) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}



template <int MASK_MODE, int EMU_PAIRS, int NQ, bool DROP = false>
int launch_variant(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                   const CUtensorMap& tmO, const FmhaParams& p, cudaStream_t stream) {
  using C = Cfg<NQ>;
  static bool configured_on[kMaxDevices];
  bool& configured = configured_on[current_device()];
  if (!configured) {
    RP_CUDA_CHECK(cudaFuncSetAttribute(fmha_fwd_kernel<MASK_MODE, EMU_PAIRS, NQ, DROP>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_TOTAL));
    configured = true;
  }
  dim3 grid(unsigned((p.Tq + NQ * QT - 1) / (NQ * QT)) * unsigned(p.H) * unsigned(p.B));
  RP_CUDA_CHECK(launch_pdl(fmha_fwd_kernel<MASK_MODE, EMU_PAIRS, NQ, DROP>, grid, dim3(C::NUM_THREADS), C::SMEM_TOTAL,
                           stream, tmQ, tmK, tmV, tmO, p));
  count_launch();
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

}  // namespace

#ifdef RP_FMHA_TRACE
extern "C" int rp_debug_fmha_trace(unsigned long long* host_out) {
  return int(cudaMemcpyFromSymbol(host_out, g_fmha_trace, sizeof(unsigned long long) * 8 * 512));
}

#endif

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

  FmhaParams p{a.B, a.H, a.Tq, a.Tk, a.kv_lens, a.mask, a.mask_b_stride, a.mask_q_stride, a.lse,
               (a.skip_padded_queries && a.kv_lens != nullptr && a.Tq == a.Tk) ? 1 : 0,
               a.drop_bits, a.drop_ld, a.drop_scale, a.batch_order};
  if (a.drop_bits != nullptr) {
    RP_CHECK(a.mask_mode == 0, "fmha: attention-weight dropout is built for the key-padding mode");
    RP_CHECK(a.drop_ld % 4 == 0 && a.drop_ld * 32 >= ((int64_t(a.Tk) + KT - 1) / KT) * KT &&
                 reinterpret_cast<uintptr_t>(a.drop_bits) % 16 == 0,
             "fmha: keep-bit rows must be 16-byte aligned and cover every 128-key tile");
    return launch_variant<0, 1, 1, true>(tmQ, tmK, tmV, tmO, p, stream);
  }
  // One build of the kernel ships (NQ = 1: two independent CTAs per SM; one exp2 pair in four on the FMA pipe).
  // The alternatives that were built, verified and measured — NQ = 2, 0 or 2 emulated pairs, a 64-key-tile
  // three-CTA variant and a two-warpgroup ping-pong kernel — live under tools/experiments/ with their numbers.
  if (a.mask_mode == 1) return launch_variant<1, 1, 1>(tmQ, tmK, tmV, tmO, p, stream);
  return launch_variant<0, 1, 1>(tmQ, tmK, tmV, tmO, p, stream);
}

}  // namespace rp
