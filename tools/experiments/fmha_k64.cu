// NOT BUILT — kept for the record (round 2).  64-key-tile variant of the attention kernel: 160 TMEM columns, 128
// registers per softmax thread and 73 KB of shared memory per CTA, i.e. THREE CTAs (three softmax warps per
// scheduler) per SM.  Parity-green on every test_fmha_* case at the first run; measured on B200, B=32, H=8, T=1801:
// 688 TFLOP/s vs 683 for the two-CTA kernel on the same box (667 vs 668 on another): issue-slot utilisation
// rises from 55 % to 64 % but the instruction count rises by the same 16 % (twice as many iterations, each with
// its barrier probes / hand-overs / rescale check), MUFU stays at 53 %, and 3840 CTAs on 444 slots quantise to
// 9 waves for 8.65.  ncu: profiles/r02_prof_fmha_k64_summary.txt.  It needs the helpers of
// repurpose_b200/csrc/fmha.cu (Cfg-independent part) and ptx.cuh's tmem_alloc2 to compile.
// =================================================================================================
// 64-key-tile variant: the same row-per-thread pipeline with KT = 64, so that a CTA needs 160 TMEM
// columns (S 64 | O 64 in a 128-column allocation, P 32 in a second one), 128 registers per softmax
// thread and 73 KB of shared memory — THREE CTAs per SM, i.e. three softmax warps per scheduler
// instead of two.  The traces of the two-CTA kernel (profiles/r01_notes.md) and of the ping-pong kernel
// below (profiles/r02_notes.md) show in-order warps that are latency-bound in every phase; a third
// independent warp per scheduler is what fills the MUFU and issue slots they leave idle.
// =================================================================================================
namespace k64 {
constexpr int KT = 64;
constexpr int K_STAGES = 3;
constexpr int V_STAGES = 4;
constexpr int SMEM_BAR_OFF = TILE_BYTES + (K_STAGES + V_STAGES) * KT * HD * 2;
constexpr int SMEM_TOTAL = SMEM_BAR_OFF + 512 + 1024;
constexpr int SOFTMAX_WARPS = 4, PRODUCER_WARP = 4, MMA_WARP = 5, PV_WARP = 6;
constexpr int NUM_THREADS = 256;
// 3 CTAs x 256 threads x 80 registers at launch -> 128 x 128 (softmax) + 128 x 32 (auxiliary) per CTA
constexpr int SOFTMAX_REGS = 128;
constexpr int AUX_REGS = 32;
}  // namespace k64

template <int MASK_MODE, int EMU_PAIRS>
__global__ void __launch_bounds__(k64::NUM_THREADS, 3)
fmha_k64_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                 const FmhaParams p) {
  constexpr int NQ = 1;
  constexpr int KT = k64::KT, NCH = KT / CH, KV_BYTES = KT * HD * 2;
  constexpr int K_STAGES = k64::K_STAGES, V_STAGES = k64::V_STAGES;
  constexpr int SMEM_Q_OFF = 0, SMEM_K_OFF = TILE_BYTES, SMEM_V_OFF = SMEM_K_OFF + K_STAGES * KV_BYTES;
  constexpr int SMEM_BAR_OFF = k64::SMEM_BAR_OFF;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);

  const uint32_t bar_base = pin_u32(base + SMEM_BAR_OFF);
  auto q_full = [&](int q) { return bar_base + 8u * q; };
  auto k_full = [&](int s) { return bar_base + 16u + 8u * s; };
  auto k_empty = [&](int s) { return bar_base + 48u + 8u * s; };
  auto v_full = [&](int s) { return bar_base + 80u + 8u * s; };
  auto v_empty = [&](int s) { return bar_base + 112u + 8u * s; };
  auto s_full = [&](int q, int h) { return bar_base + 144u + 8u * (2 * q + h); };
  auto s_free = [&](int q, int h) { return bar_base + 176u + 8u * (2 * q + h); };
  auto p_ready = [&](int q, int h) { return bar_base + 208u + 8u * (2 * q + h); };
  auto pv_done = [&](int q, int h) { return bar_base + 240u + 8u * (2 * q + h); };
  constexpr int TMEM_SLOT_OFF = 288;  // two words: the 128-column allocation (S | O) and the 32-column one (P)
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SMEM_BAR_OFF + TMEM_SLOT_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int b = blockIdx.z;

  const int q_start0 = blockIdx.x * (NQ * QT);
  const bool q1_active = NQ == 2 && (q_start0 + QT) < p.Tq;
  const int nq = q1_active ? 2 : 1;
  int kv_len = p.Tk;
  if (MASK_MODE == 0 && p.kv_lens != nullptr) {
    kv_len = p.kv_lens[b];
    kv_len = kv_len < 0 ? 0 : (kv_len > p.Tk ? p.Tk : kv_len);
  }
  const int n_kv = (kv_len + KT - 1) / KT;

  if (warp == k64::PRODUCER_WARP && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    for (int q = 0; q < 2; ++q) {
      mbar_init(q_full(q), 1);
      for (int h = 0; h < 2; ++h) {
        mbar_init(s_full(q, h), 1);
        mbar_init(s_free(q, h), 4);
        mbar_init(p_ready(q, h), 4);
        mbar_init(pv_done(q, h), 1);
      }
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(k_full(s), 1);
      mbar_init(k_empty(s), 1);
      mbar_init(v_full(s), 1);
      mbar_init(v_empty(s), 1);
    }
    fence_mbar_init();
  }
  if (warp == k64::MMA_WARP) tmem_alloc2<128, 32>(base + SMEM_BAR_OFF + TMEM_SLOT_OFF, base + SMEM_BAR_OFF + TMEM_SLOT_OFF + 4);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_a = tmem_slot[0];  // S [0, 64) | O [64, 128)
  const uint32_t tmem_b = tmem_slot[1];  // P (bf16 pairs) [0, 32)
  pdl_wait();  // the set-up above overlapped the previous kernel's tail

  if (warp >= k64::SOFTMAX_WARPS) setmaxnreg_dec<k64::AUX_REGS>();
  if (warp == k64::PRODUCER_WARP) {
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
          mbar_expect_tx(k_full(st), KV_BYTES);
          tma_load_3d(base + SMEM_K_OFF + st * KV_BYTES, &tmK, k_full(st), col, i * KT, b);
        }
        __syncwarp();
      };
      auto load_v = [&](int i) {
        const int st = i % V_STAGES;
        mbar_wait(v_empty(st), (uint32_t(i / V_STAGES) & 1u) ^ 1u);
        if (lane == 0) TRACE(7, 8 * i + 5);
        if (elect_one()) {
          mbar_expect_tx(v_full(st), KV_BYTES);
          tma_load_3d(base + SMEM_V_OFF + st * KV_BYTES, &tmV, v_full(st), col, i * KT, b);
        }
        __syncwarp();
      };
      load_k(0);
      for (int i = 0; i < n_kv; ++i) {
        if (i + 1 < n_kv) load_k(i + 1);
        load_v(i);
      }
    }
  } else if (warp == k64::MMA_WARP) {
    // ---------------------------------------------------------------- QK^T issuer (and TMEM owner)
    // Converged warp, one elected lane per issue group (descriptors stay in uniform registers).
    // S = Q K^T and O += P V are issued from two different warps so that neither sits behind the
    // other's barrier waits: the softmax hands S_h and P_h back at the same moment and needs both
    // results half an iteration later.
    if (n_kv > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(QT, KT / 2, false, false);
      // S_q[:, 64h .. 64h+64) = Q_q K[64h .. 64h+64)^T
      auto issue_qk = [&](int q, int st, int h, uint32_t commit_bar, uint32_t commit_bar2) {
        if (elect_one()) {
          const uint64_t da = make_smem_desc_sw128(base + SMEM_Q_OFF + q * TILE_BYTES, 1024, 16);
          const uint64_t db =
              make_smem_desc_sw128(base + SMEM_K_OFF + st * KV_BYTES + h * (KV_BYTES / 2), 1024, 16);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            mma_ss(tmem_a + h * (KT / 2), da + uint64_t(2 * k), db + uint64_t(2 * k), idesc_s,
                   k > 0 ? 1u : 0u);
          tc_commit(commit_bar);
          if (commit_bar2 != 0) tc_commit(commit_bar2);
        }
        __syncwarp();
      };
      for (int q = 0; q < nq; ++q) mbar_wait(q_full(q), 0);
      mbar_wait(k_full(0), 0);
      tc_fence_after();
      for (int h = 0; h < 2; ++h)
        for (int q = 0; q < nq; ++q)
          issue_qk(q, 0, h, s_full(q, h), (h == 1 && q == nq - 1) ? k_empty(0) : 0u);
      // The softmax loop is branch-free: it always prefetches "the next tile".  After the last real
      // tile that is a dummy S (tile n_kv-1 once more, from the K stage that is still resident; no
      // ring bookkeeping), whose scores are never used.  S_i may be issued once the softmax warps
      // have read S_{i-1} (s_free phase i-1).
      for (int i = 1; i <= n_kv; ++i) {
        const bool real = i < n_kv;
        const int st = (real ? i : n_kv - 1) % K_STAGES;
        if (real) mbar_wait(k_full(st), uint32_t(i / K_STAGES) & 1u);
        for (int h = 0; h < 2; ++h)
          for (int q = 0; q < nq; ++q) {
            mbar_wait(s_free(q, h), uint32_t(i - 1) & 1u);
            tc_fence_after();
            if (lane == 0) TRACE(2 + h, i);
            issue_qk(q, st, h, s_full(q, h), (real && h == 1 && q == nq - 1) ? k_empty(st) : 0u);
          }
      }
    }
  } else if (warp == k64::PV_WARP) {
    // ---------------------------------------------------------------- P V issuer
    if (n_kv > 0) {
      constexpr uint32_t idesc_o = make_idesc_bf16(QT, HD, false, true);  // V is MN-major
      // O_q += P_q[:, 64h .. 64h+64) V[64h .. 64h+64)
      auto issue_pv = [&](int q, int st, int h, bool acc, uint32_t commit_bar, uint32_t commit_bar2) {
        if (elect_one()) {
          // A: P in TMEM, 16 keys = 8 packed columns per step; B: 16 key rows of 128 B each
          const uint64_t db = make_smem_desc_sw128(base + SMEM_V_OFF + st * KV_BYTES, 1024, 1024);
#pragma unroll
          for (int k = 0; k < KT / 32; ++k)
            mma_ts(tmem_a + 64, tmem_b + h * (KT / 4) + k * 8,
                   db + uint64_t(128 * ((KT / 32) * h + k)), idesc_o, (acc || k > 0) ? 1u : 0u);
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
            issue_pv(q, vst, h, j > 0 || h > 0, pv_done(q, h), (h == 1 && q == nq - 1) ? v_empty(vst) : 0u);
          }
      }
    }
  } else if (warp < k64::SOFTMAX_WARPS) {
    setmaxnreg_inc<k64::SOFTMAX_REGS>();
    // ---------------------------------------------------------------- softmax warps
    const int q = warp >> 2;
    const int wl = warp & 3;
    const int row_in_tile = wl * 32 + lane;
    const int q_start = q_start0 + q * QT;
    const uint32_t stage_smem = base + SMEM_Q_OFF + q * TILE_BYTES;  // reused for the O tile
    if (q < nq) {
      const uint32_t lane_off = uint32_t(wl * 32) << 16;
      const uint32_t t_s = pin_u32(tmem_a + lane_off);
      const uint32_t t_p = pin_u32(tmem_b + lane_off);
      const uint32_t t_o = t_s + 64;
      const bool tracer = wl == 0 && lane == 0 && q == 0;
      float m = -INFINITY;
      unsigned long long lsumA = pack2(0.f, 0.f), lsumB = pack2(0.f, 0.f);
      const uint8_t* mrow = nullptr;
      if (MASK_MODE == 1) {
        const int qrow = q_start + row_in_tile;
        if (qrow < p.Tq) mrow = p.mask + int64_t(b) * p.mask_b_stride + int64_t(qrow) * p.mask_q_stride;
      }
      uint32_t xs[KT];  // scores of the current tile -> probabilities -> scores of the next tile
      const bool lane0 = pin_u32(lane == 0 ? 1u : 0u) != 0u;

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
        for (int i = 0; i < KT; ++i)
          if (i >= nv) xs[i] = 0xff800000u;  // -inf
        if (MASK_MODE == 1 && mrow != nullptr) {
          const uint8_t* mp = mrow + tile * KT;
#pragma unroll
          for (int i = 0; i < KT; ++i)
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
        for (int c = 0; c < NCH; ++c) tmem_ld16(t_s + 16 * c, xs + 16 * c);
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

      for (int j = 0; j < n_kv; ++j) {
        const bool has_next = j + 1 < n_kv;
        const uint32_t par = uint32_t(j) & 1u, par_next = par ^ 1u;
        const unsigned long long negm2 = pack2(-m, -m);
        float mx0 = -INFINITY, mx1 = -INFINITY;
        uint32_t probe_pv = 0, probe_s = 0;
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
                pk[i] = pack_bf16x2(p0, p1);
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
        if (has_next) {
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

      // ---- epilogue: O / l -> bf16 -> swizzled smem (the Q tile's slot) -> TMA store
      const uint32_t row_addr = stage_smem + uint32_t(row_in_tile) * 128u;
      float inv_l = 0.0f;
      if (n_kv > 0) {
        float a0, a1, b0, b1;
        unpack2(lsumA, a0, a1);
        unpack2(lsumB, b0, b1);
        inv_l = 1.0f / ((a0 + a1) + (b0 + b1));
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
      named_bar_sync(1 + q, 128);
      if (wl == 0 && lane == 0) {
        tma_store_3d(&tmO, stage_smem, head * HD, q_start, b);
        tma_store_commit();
        tma_store_wait_all<0>();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == k64::MMA_WARP) {
    tc_fence_after();
    tmem_dealloc<128>(tmem_a);
    tmem_dealloc<32>(tmem_b);
  }
}



template <int MASK_MODE, int EMU_PAIRS>
int launch_k64(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
               const CUtensorMap& tmO, const FmhaParams& p, cudaStream_t stream) {
  static bool configured_on[kMaxDevices];
  bool& configured = configured_on[current_device()];
  if (!configured) {
    RP_CUDA_CHECK(cudaFuncSetAttribute(fmha_k64_kernel<MASK_MODE, EMU_PAIRS>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, k64::SMEM_TOTAL));
    configured = true;
  }
  dim3 grid((p.Tq + QT - 1) / QT, p.H, p.B);
  RP_CUDA_CHECK(launch_pdl(fmha_k64_kernel<MASK_MODE, EMU_PAIRS>, grid, dim3(k64::NUM_THREADS), k64::SMEM_TOTAL,
                           stream, tmQ, tmK, tmV, tmO, p));
  count_launch();
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

