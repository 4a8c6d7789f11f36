// NOT BUILT — kept for the record (round 2).  "Ping-pong" attention kernel: one CTA per 256 query rows, two softmax
// warpgroups that hand the MUFU to each other with a named-barrier token per scheduler, phases
// A (FMA-pipe work) / X (MUFU) / D (sum, pack, P store, next scores) separated by branches ptxas cannot fold.
// Parity-green at the first run; measured 523 TFLOP/s vs 668 for the two-CTA kernel.  The in-kernel trace
// (profiles/r02_notes.md) shows why: the token is acquired without waiting (~50 cycles) — the MUFU was never the
// contended resource — while every phase of an in-order warp is latency-bound on its own (A 670, X 1125 for 96
// MUFU = 11.7 cycles each from ONE warp, D 1650 cycles), so serialising the phases costs more than the hand-over
// buys, and a single CTA per SM leaves the 3500-cycle prologue of every CTA exposed.
// =================================================================================================
// Ping-pong kernel: one CTA = one (batch, head, 256 query rows), one CTA per SM.
//
// Two softmax warpgroups (query tile 0 / query tile 1) share the K / V stages (staged once per 256 rows)
// and — this is the point — hand the MUFU to each other.  An iteration of a softmax warp is split into
//   A  (FMA / ALU pipes)  subtract the running reference, exp2 of the EMU_PAIRS-in-4 emulated pairs
//   X  (MUFU)             the remaining ex2.approx, back to back, between bar.sync / bar.arrive of a
//                         named-barrier token shared by the two warps of one SM sub-partition
//                         (warp w of tile 0 and warp w of tile 1 sit on the same scheduler)
//   D  (everything else)  row sum, bf16 pack, P store, load of the next tile's scores, running max
// so that while one warp of a scheduler occupies the MUFU (768 cycles per 128 x 128 tile with one pair
// in four emulated) the other one issues its D and A work; two co-resident independent CTAs (the
// kernel above) drift into the same phase and leave the MUFU idle ~40 % of the time
// (profiles/r01_notes.md).
//
// The last key tile is peeled: only its ceil(valid keys / 16) sixteen-key chunks are multiplied,
// exponentiated and fed to P V (a T = 1801 sequence has 9 keys in its 15th tile), and there is no dummy
// "next tile".  Softmax warps whose 32 query rows all lie beyond Tq exit at once; the last query block
// of every (batch, head) — the one that may be partial — is scheduled at the end of the grid.
// =================================================================================================
#ifdef RP_FMHA_TRACE
#ifndef RP_TRACE_BLOCK
#define RP_TRACE_BLOCK 700
#endif
__device__ unsigned long long g_pp_trace[8 * 512];  // [role][event] = clock64
#define PTRACE(role, idx)                                                     \
  do {                                                                        \
    if (blockIdx.x == RP_TRACE_BLOCK && (idx) < 512) g_pp_trace[(role) * 512 + (idx)] = clock64(); \
  } while (0)
#else
#define PTRACE(role, idx) do {} while (0)
#endif
namespace pp {
constexpr int K_STAGES = 4;
constexpr int V_STAGES = 4;
constexpr int SMEM_Q_OFF = 0;
constexpr int SMEM_K_OFF = 2 * TILE_BYTES;
constexpr int SMEM_V_OFF = SMEM_K_OFF + K_STAGES * TILE_BYTES;
constexpr int SMEM_BAR_OFF = SMEM_V_OFF + V_STAGES * TILE_BYTES;
constexpr int SMEM_TOTAL = SMEM_BAR_OFF + 1024 + 1024;
constexpr int NUM_THREADS = 384;
constexpr int PRODUCER_WARP = 8, QK_WARP = 9, PV_WARP = 10;
constexpr int TMEM_COLS = 512;
constexpr int TM_S = 0;    // + q * 128
constexpr int TM_P = 256;  // + q * 64
constexpr int TM_O = 384;  // + q * 64
constexpr int SOFTMAX_REGS = 232;  // 384 x 168 at launch -> 256 x 232 + 128 x 40
constexpr int AUX_REGS = 40;
constexpr int BAR_TOKEN0 = 2;   // named barriers 2 .. 9: MUFU token of scheduler wl, owner q: 2 + 2 wl + q
constexpr int BAR_EPI0 = 10;    // 10, 11: epilogue of query tile q
}  // namespace pp

template <int MASK_MODE, int EMU_PAIRS, int PP>
__global__ void __launch_bounds__(pp::NUM_THREADS, 1)
fmha_pp_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
               const FmhaParams p) {
  using namespace pp;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);

  const uint32_t bar_base = pin_u32(base + SMEM_BAR_OFF);
  auto q_full = [&](int q) { return bar_base + 8u * q; };
  auto k_full = [&](int s) { return bar_base + 8u * (2 + s); };
  auto k_empty = [&](int s) { return bar_base + 8u * (6 + s); };
  auto v_full = [&](int s) { return bar_base + 8u * (10 + s); };
  auto v_empty = [&](int s) { return bar_base + 8u * (14 + s); };
  auto s_full = [&](int q, int h) { return bar_base + 8u * (18 + 2 * q + h); };
  auto s_free = [&](int q, int h) { return bar_base + 8u * (22 + 2 * q + h); };
  auto p_ready = [&](int q, int h) { return bar_base + 8u * (26 + 2 * q + h); };
  auto pv_done = [&](int q, int h) { return bar_base + 8u * (30 + 2 * q + h); };
  constexpr int TMEM_SLOT_OFF = 512;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SMEM_BAR_OFF + TMEM_SLOT_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // grid order: all query blocks but the last of every (batch, head) first, the last ones at the end
  const int nqb = (p.Tq + 2 * QT - 1) / (2 * QT);
  const int n_lead = (nqb - 1) * p.H * p.B;
  int qb, bh;
  if (int(blockIdx.x) < n_lead) {
    qb = int(blockIdx.x) % (nqb - 1);
    bh = int(blockIdx.x) / (nqb - 1);
  } else {
    qb = nqb - 1;
    bh = int(blockIdx.x) - n_lead;
  }
  const int head = bh % p.H;
  const int b = bh / p.H;

  const int q_start0 = qb * (2 * QT);
  const int rows_left = p.Tq - q_start0;  // > 0
  const int nq = rows_left > QT ? 2 : 1;
  // softmax warps with at least one query row below Tq, per query tile
  auto warps_of = [&](int q) {
    const int r = rows_left - q * QT;
    return r <= 0 ? 0 : (r >= QT ? 4 : (r + 31) >> 5);
  };
  int kv_len = p.Tk;
  if (MASK_MODE == 0 && p.kv_lens != nullptr) {
    kv_len = p.kv_lens[b];
    kv_len = kv_len < 0 ? 0 : (kv_len > p.Tk ? p.Tk : kv_len);
  }
  const int n_kv = (kv_len + KT - 1) / KT;
  const int nv_last = kv_len - (n_kv - 1) * KT;  // valid keys of the last tile (1 .. 128)
  const int nch_last = (nv_last + 15) >> 4;      // its 16-key chunks (1 .. 8)

  if (warp == PRODUCER_WARP && lane == 0) {
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
    for (int s = 0; s < 4; ++s) {
      mbar_init(k_full(s), 1);
      mbar_init(k_empty(s), 1);
      mbar_init(v_full(s), 1);
      mbar_init(v_empty(s), 1);
    }
    fence_mbar_init();
  }
  if (warp == QK_WARP) tmem_alloc<TMEM_COLS>(base + SMEM_BAR_OFF + TMEM_SLOT_OFF);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // the set-up above overlapped the previous kernel's tail

  if (warp >= 8) setmaxnreg_dec<AUX_REGS>();
  if (warp == PRODUCER_WARP) {
    // ---------------------------------------------------------------- TMA producer
    if (n_kv > 0) {
      const int col = head * HD;
      if (elect_one()) {
        mbar_expect_tx(q_full(0), TILE_BYTES);
        tma_load_3d(base + SMEM_Q_OFF, &tmQ, q_full(0), col, q_start0, b);
        if (nq == 2) {
          mbar_expect_tx(q_full(1), TILE_BYTES);
          tma_load_3d(base + SMEM_Q_OFF + TILE_BYTES, &tmQ, q_full(1), col, q_start0 + QT, b);
        }
      }
      __syncwarp();
      auto load_k = [&](int i) {
        const int st = i % K_STAGES;
        mbar_wait(k_empty(st), (uint32_t(i / K_STAGES) & 1u) ^ 1u);
        if (lane == 0) PTRACE(4, 2 * i);
        if (elect_one()) {
          mbar_expect_tx(k_full(st), TILE_BYTES);
          tma_load_3d(base + SMEM_K_OFF + st * TILE_BYTES, &tmK, k_full(st), col, i * KT, b);
        }
        __syncwarp();
      };
      auto load_v = [&](int i) {
        const int st = i % V_STAGES;
        mbar_wait(v_empty(st), (uint32_t(i / V_STAGES) & 1u) ^ 1u);
        if (lane == 0) PTRACE(4, 2 * i + 1);
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
  } else if (warp == QK_WARP) {
    // ---------------------------------------------------------------- S = Q K^T issuer (TMEM owner)
    // Tile-major in q: the two softmax warpgroups run half an iteration apart, so the hand-backs of
    // one query tile (half 0, half 1) arrive together and those of the other one later.
    if (n_kv > 0) {
      // S_q[:, 64h .. 64h + ncols) = Q_q K[64h .. 64h + ncols)^T;  ncols = 0: nothing to multiply,
      // the commit alone completes the barrier phase
      auto issue_qk = [&](int q, int st, int h, int ncols, uint32_t commit_bar, uint32_t commit_bar2) {
        if (elect_one()) {
          if (ncols > 0) {
            const uint32_t idesc = make_idesc_bf16(QT, 0, false, false) | (uint32_t(ncols >> 3) << 17);
            const uint64_t da = make_smem_desc_sw128(base + SMEM_Q_OFF + q * TILE_BYTES, 1024, 16);
            const uint64_t db =
                make_smem_desc_sw128(base + SMEM_K_OFF + st * TILE_BYTES + h * (TILE_BYTES / 2), 1024, 16);
#pragma unroll
            for (int k = 0; k < HD / 16; ++k)
              mma_ss(tmem_base + TM_S + q * 128 + h * 64, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc,
                     k > 0 ? 1u : 0u);
          }
          tc_commit(commit_bar);
          if (commit_bar2 != 0) tc_commit(commit_bar2);
        }
        __syncwarp();
      };
      auto ncols_of = [&](int i, int h) {
        if (i + 1 < n_kv) return 64;
        const int c = nch_last * 16 - 64 * h;
        return c < 0 ? 0 : (c > 64 ? 64 : c);
      };
      for (int q = 0; q < nq; ++q) mbar_wait(q_full(q), 0);
      for (int i = 0; i < n_kv; ++i) {
        const int st = i % K_STAGES;
        mbar_wait(k_full(st), uint32_t(i / K_STAGES) & 1u);
        if (lane == 0) PTRACE(5, i);
        for (int q = 0; q < nq; ++q)
          for (int h = 0; h < 2; ++h) {
            if (i > 0) mbar_wait(s_free(q, h), uint32_t(i - 1) & 1u);  // S_{i-1} half h is in registers
            tc_fence_after();
            if (lane == 0) PTRACE(2, 4 * i + 2 * q + h);
            issue_qk(q, st, h, ncols_of(i, h), s_full(q, h), (h == 1 && q == nq - 1) ? k_empty(st) : 0u);
          }
      }
    }
  } else if (warp == PV_WARP) {
    // ---------------------------------------------------------------- O += P V issuer
    if (n_kv > 0) {
      constexpr uint32_t idesc_o = make_idesc_bf16(QT, HD, false, true);  // V is MN-major
      // O_q += P_q[:, 64h .. 64h + 16 ksteps) V[64h .. 64h + 16 ksteps)
      auto issue_pv = [&](int q, int st, int h, bool acc, int ksteps, uint32_t commit_bar, uint32_t commit_bar2) {
        if (elect_one()) {
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
      auto ksteps_of = [&](int j, int h) {
        if (j + 1 < n_kv) return 4;
        const int c = nch_last - 4 * h;
        return c < 0 ? 0 : (c > 4 ? 4 : c);
      };
      for (int j = 0; j < n_kv; ++j) {
        const int vst = j % V_STAGES;
        mbar_wait(v_full(vst), uint32_t(j / V_STAGES) & 1u);
        for (int q = 0; q < nq; ++q)
          for (int h = 0; h < 2; ++h) {
            mbar_wait(p_ready(q, h), uint32_t(j) & 1u);
            tc_fence_after();
            if (lane == 0) PTRACE(3, 4 * j + 2 * q + h);
            issue_pv(q, vst, h, j > 0 || h > 0, ksteps_of(j, h), pv_done(q, h),
                     (h == 1 && q == nq - 1) ? v_empty(vst) : 0u);
          }
      }
    }
  } else if (warp < 8) {
    setmaxnreg_inc<SOFTMAX_REGS>();
    // ---------------------------------------------------------------- softmax warps
    const int q = warp >> 2;
    const int wl = warp & 3;
    const int row_in_tile = wl * 32 + lane;
    const int q_start = q_start0 + q * QT;
    const uint32_t stage_smem = base + SMEM_Q_OFF + q * TILE_BYTES;  // reused for the O tile
    if (wl < warps_of(q)) {
      const uint32_t lane_off = uint32_t(wl * 32) << 16;
      const uint32_t t_s = pin_u32(tmem_base + lane_off + TM_S + q * 128);
      const uint32_t t_p = pin_u32(tmem_base + lane_off + TM_P + q * 64);
      const uint32_t t_o = pin_u32(tmem_base + lane_off + TM_O + q * 64);
      // MUFU token: only when the warp of the other query tile on this scheduler exists
      const bool use_tok = PP != 0 && n_kv > 0 && wl < warps_of(1);
      const int tok_mine = BAR_TOKEN0 + 2 * wl + q;
      const int tok_other = BAR_TOKEN0 + 2 * wl + (q ^ 1);
      if (use_tok && q == 1) named_bar_arrive(tok_other, 64);  // tile 0 goes first
      float m = -INFINITY;
      unsigned long long lsumA = pack2(0.f, 0.f), lsumB = pack2(0.f, 0.f);
      const uint8_t* mrow = nullptr;
      if (MASK_MODE == 1) {
        const int qrow = q_start + row_in_tile;
        if (qrow < p.Tq) mrow = p.mask + int64_t(b) * p.mask_b_stride + int64_t(qrow) * p.mask_q_stride;
      }
      uint32_t xs[128];  // scores of the current tile -> probabilities -> scores of the next tile
      const bool lane0 = pin_u32(lane == 0 ? 1u : 0u) != 0u;
      const bool opaque_true = p.H > 0, opaque_true_a = p.B > 0;
      const bool tracer = wl == 0 && lane == 0;
      int trace_j = 0;

      auto max_chunk = [&](int c, float& mx0, float& mx1) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          mx0 = max3(mx0, __uint_as_float(xs[16 * c + i]), __uint_as_float(xs[16 * c + i + 1]));
          mx1 = max3(mx1, __uint_as_float(xs[16 * c + i + 2]), __uint_as_float(xs[16 * c + i + 3]));
        }
      };
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
        for (int c = 0; c < 8; ++c) max_chunk(c, mx0, mx1);
      };
      // A + X phases over chunks [0, nch): xs <- exp2(xs - m)
      auto exp_phase = [&](auto tail_tag, int nch) {
        constexpr bool TAIL = decltype(tail_tag)::value;
        const unsigned long long negm2 = pack2(-m, -m);
        // (the two phases sit behind branches ptxas cannot fold: without them it hoists MUFU instructions
        // above the bar.sync and sinks the emulation arithmetic below it, into the token's hold time)
        if (opaque_true_a) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            if (!TAIL || c < nch) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int c0 = 16 * c + 2 * i;
                const unsigned long long x2 =
                    add2(pack2(__uint_as_float(xs[c0]), __uint_as_float(xs[c0 + 1])), negm2);
                float p0, p1;
                if ((i & 3) < EMU_PAIRS) exp2_emulated2(x2, p0, p1);
                else unpack2(x2, p0, p1);
                xs[c0] = __float_as_uint(p0);
                xs[c0 + 1] = __float_as_uint(p1);
              }
            }
          }
        }
        if (tracer) PTRACE(q, 8 * trace_j + 1);
        if (use_tok) named_bar_sync(tok_mine, 64);
        if (tracer) PTRACE(q, 8 * trace_j + 2);
        if (opaque_true) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            if (!TAIL || c < nch) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                if ((i & 3) >= EMU_PAIRS) {
                  const int c0 = 16 * c + 2 * i;
                  xs[c0] = __float_as_uint(ex2_approx(__uint_as_float(xs[c0])));
                  xs[c0 + 1] = __float_as_uint(ex2_approx(__uint_as_float(xs[c0 + 1])));
                }
              }
            }
          }
        }
      };
      // row sum + bf16 pack + P store of chunk c
      auto consume_chunk = [&](int c) {
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float p0 = __uint_as_float(xs[16 * c + 2 * i]), p1 = __uint_as_float(xs[16 * c + 2 * i + 1]);
          if (i & 1) lsumB = add2(lsumB, pack2(p0, p1));
          else lsumA = add2(lsumA, pack2(p0, p1));
          pk[i] = pack_bf16x2(p0, p1);
        }
        tmem_st8(t_p + 8 * c, pk);
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
        if (lane0) {
          mbar_arrive(s_free(q, 0));
          mbar_arrive(s_free(q, 1));
        }
        float mx0 = -INFINITY, mx1 = -INFINITY;
        if (tile_needs_mask(0)) {
          mask_tile(0, mx0, mx1);
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) max_chunk(c, mx0, mx1);
        }
        m = fmaxf(mx0, mx1);
      }

      // ---- full tiles 0 .. n_kv-2 (each one prefetches the scores of its successor)
      for (int j = 0; j + 1 < n_kv; ++j) {
        const uint32_t par = uint32_t(j) & 1u, par_next = par ^ 1u;
        // probes for the first halves: P V_{j-1,0} done (P half 0 may be overwritten), S_{j+1,0} in TMEM
        trace_j = j;
        if (tracer) PTRACE(q, 8 * j);
        const uint32_t probe_pv0 = j > 0 ? mbar_test_wait(pv_done(q, 0), par_next) : 1u;
        const uint32_t probe_s0 = mbar_test_wait(s_full(q, 0), par_next);
        exp_phase(std::false_type{}, 8);
        const uint32_t probe_pv1 = j > 0 ? mbar_test_wait(pv_done(q, 1), par_next) : 1u;
        const uint32_t probe_s1 = mbar_test_wait(s_full(q, 1), par_next);
        if (tracer) PTRACE(q, 8 * j + 3);
        if (use_tok) named_bar_arrive(tok_other, 64);
        float mx0 = -INFINITY, mx1 = -INFINITY;
        // ---- D, first half
        if (!probe_pv0) mbar_wait_spin(pv_done(q, 0), par_next);
        if (!probe_s0) mbar_wait_spin(s_full(q, 0), par_next);
        tc_fence_after();
        if (tracer) PTRACE(q, 8 * j + 4);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          consume_chunk(c);
          tmem_ld16(t_s + 16 * c, xs + 16 * c);
        }
        tmem_ld_wait();
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane0) {
          mbar_arrive(p_ready(q, 0));
          mbar_arrive(s_free(q, 0));
        }
        if (tracer) PTRACE(q, 8 * j + 5);
#pragma unroll
        for (int c = 0; c < 4; ++c) max_chunk(c, mx0, mx1);
        // ---- D, second half
        if (!probe_pv1) mbar_wait_spin(pv_done(q, 1), par_next);
        if (!probe_s1) mbar_wait_spin(s_full(q, 1), par_next);
        tc_fence_after();
        if (tracer) PTRACE(q, 8 * j + 6);
#pragma unroll
        for (int c = 4; c < 8; ++c) {
          consume_chunk(c);
          tmem_ld16(t_s + 16 * c, xs + 16 * c);
        }
        tmem_ld_wait();
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane0) {
          mbar_arrive(p_ready(q, 1));
          mbar_arrive(s_free(q, 1));
        }
        if (tracer) PTRACE(q, 8 * j + 7);
#pragma unroll
        for (int c = 4; c < 8; ++c) max_chunk(c, mx0, mx1);
        // ---- the next tile is in registers: mask it if needed, move the reference if it grew a lot
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

      // ---- last tile: only its valid 16-key chunks, nothing to prefetch
      if (n_kv > 0) {
        const int j = n_kv - 1;
        const uint32_t par_next = (uint32_t(j) & 1u) ^ 1u;
        exp_phase(std::true_type{}, nch_last);
        if (use_tok && !(q == 1)) named_bar_arrive(tok_other, 64);  // tile 1's last release has no taker
        if (j > 0) mbar_wait(pv_done(q, 0), par_next);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < nch_last) consume_chunk(c);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane0) mbar_arrive(p_ready(q, 0));
        if (j > 0) mbar_wait(pv_done(q, 1), par_next);
        tc_fence_after();
#pragma unroll
        for (int c = 4; c < 8; ++c)
          if (c < nch_last) consume_chunk(c);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane0) mbar_arrive(p_ready(q, 1));
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
      named_bar_sync(BAR_EPI0 + q, 32 * warps_of(q));
      if (wl == 0 && lane == 0) {
        tma_store_3d(&tmO, stage_smem, head * HD, q_start, b);
        tma_store_commit();
        tma_store_wait_all<0>();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == QK_WARP) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

template <int MASK_MODE, int EMU_PAIRS, int PP>
int launch_pp(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
              const CUtensorMap& tmO, const FmhaParams& p, cudaStream_t stream) {
  static bool configured_on[kMaxDevices];
  bool& configured = configured_on[current_device()];
  if (!configured) {
    RP_CUDA_CHECK(cudaFuncSetAttribute(fmha_pp_kernel<MASK_MODE, EMU_PAIRS, PP>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, pp::SMEM_TOTAL));
    configured = true;
  }
  const int nqb = (p.Tq + 2 * QT - 1) / (2 * QT);
  dim3 grid(unsigned(nqb) * unsigned(p.H) * unsigned(p.B));
  RP_CUDA_CHECK(launch_pdl(fmha_pp_kernel<MASK_MODE, EMU_PAIRS, PP>, grid, dim3(pp::NUM_THREADS), pp::SMEM_TOTAL,
                           stream, tmQ, tmK, tmV, tmO, p));
  count_launch();
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

