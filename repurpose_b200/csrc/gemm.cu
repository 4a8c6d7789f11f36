// Persistent warp-specialised bf16 GEMM for sm_100a:  D[M,N] = A[M,K] * W[N,K]^T (+ epilogue).
//
// Two instantiations of one kernel:
//   CG = 2 (default): CTA PAIRS (`tcgen05.mma.cta_group::2`, thread-block cluster of 2).  A pair computes
//           a 256x256 tile: each CTA stages its own 128 rows of A and its own 128 rows (half of N) of W
//           and holds the accumulator rows of its 128 A rows in its own TMEM.  Per CTA and k-block only
//           32 KB enter shared memory (vs 48 KB) and the tensor core of each SM reads half of W from the
//           peer — this is what relieves the shared-memory port that bounds the 1-CTA version.
//   CG = 1: single-CTA 128x256 tile (kept for A/B comparison and small problems).
//
//   warp 0      : TMA producer   (SWIZZLE_128B boxes into a multi-stage ring)
//   warp 1      : tcgen05.mma issuer (converged warp, one elected lane; only the leader CTA of a pair
//                 issues), fp32 accumulators in TMEM, two 256-column accumulator buffers so the epilogue
//                 of tile i overlaps the main loop of tile i+1; also owns TMEM alloc/dealloc
//   warps 2..5  : epilogue: tcgen05.ld -> bias / ReLU / residual -> swizzled smem slab -> TMA store.
//                 Each warp owns 32 accumulator rows and a private ring of 4 KB slabs.
//
// Residual epilogue (h += acc + bias, in place): the residual chunk is TMA-LOADED into the slab a few
// chunks ahead (mbarrier per slab), updated in shared memory, and TMA-stored from the same slab.
// Both directions are therefore full-line bulk copies with several in flight per warp, which is what
// a ~1.5 us DRAM round trip needs (profiles/r01_notes.md).
//
// This kernel implements the Linear layers of the reference hot path
// (models/MMCTransformer.py:121 input_projection, :135-138 in_proj/out_proj/linear1/linear2 inside
// nn.TransformerEncoderLayer, :144 feature_map[0], :147-149 head Linear layers).
#include <stdlib.h>

#include "ptx.cuh"
#include "host_util.h"
#include "kernels.h"

namespace rp {

namespace {

constexpr int BM = 128;                // accumulator rows per CTA
constexpr int BN = 256;                // accumulator columns per tile
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;   // 16 KB
constexpr int SLAB_BYTES = 32 * 128;   // 32 rows x 128 B: one TMA box of the output / residual
constexpr int TMEM_COLS = 512;
constexpr int MAX_SLABS = 8;  // barriers reserved per epilogue warp

// Shared-memory budget per (epilogue kind, CTA-group size): the residual epilogue trades operand
// stages for a deeper slab ring (loads and stores both live there).
// RS = slabs per epilogue warp of the residual epilogues: RS - 2 residual chunks of 4 KB are in flight per warp
// (4 with four epilogue warps, 3 with eight; 6 and 8 were measured and lost operand stages for nothing).
// EW = epilogue warps (4, or 8 for the LayerNorm-fused epilogue of memory-bound problems): with 8, two warps share a
// TMEM lane quarter and each walks half of the tile's columns, so two dependent chunk chains per SM sub-partition
// overlap instead of one (profiles/r02_notes.md: the 4-warp epilogue is latency-, not bandwidth-bound).
template <int EPI, int CG, int RS = 4, int EW = 4>
struct Cfg {
  static constexpr bool RESID = (EPI == EPI_BIAS_RESID_F32 || EPI == EPI_BIAS_RESID_LN);
  static constexpr bool LNF = (EPI == EPI_BIAS_RESID_LN);   // LayerNorm of the updated rows fused in (cluster of 4)
  static constexpr int B_ROWS = BN / CG;               // rows of W staged by each CTA
  static constexpr int B_BYTES = B_ROWS * BK * 2;      // 32 KB (CG=1) / 16 KB (CG=2)
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SLABS = RESID ? RS : 2;          // per epilogue warp
  static constexpr int LOOKAHEAD = SLABS - 2;          // residual chunks in flight per warp
  static constexpr int COL_SPLIT = EW / 4;             // warps per TMEM lane quarter = column parts of a tile
  static constexpr int NUM_THREADS = 64 + 32 * EW;
  static constexpr int BAR_BYTES = EW == 8 ? 1024 : 512;
  // LNF row statistics: EW = 4: [2 buffers][128 rows] (mean, M2) from the partner pair; EW = 8: [2][4 column quarters][128]
  static constexpr int STATS_BYTES = !LNF ? 0 : (EW == 8 ? 2 * 4 * 128 * 8 : 2 * 128 * 8);
  static constexpr int SLACK_BYTES = LNF ? 512 : 1024;  // alignment slack (the dynamic segment is declared 1024-aligned; checked at run time)
  static constexpr int STAGES = (232448 - BAR_BYTES - STATS_BYTES - SLACK_BYTES - EW * SLABS * SLAB_BYTES) / STAGE_BYTES;  // CG1: 4/3, CG2: 6/5
  static constexpr int SMEM_A_OFF = 0;
  static constexpr int SMEM_B_OFF = STAGES * A_BYTES;
  static constexpr int SMEM_D_OFF = SMEM_B_OFF + STAGES * B_BYTES;
  static constexpr int SMEM_BAR_OFF = SMEM_D_OFF + EW * SLABS * SLAB_BYTES;
  static constexpr int SMEM_STATS_OFF = SMEM_BAR_OFF + BAR_BYTES;
  static constexpr int SMEM_TOTAL = SMEM_STATS_OFF + STATS_BYTES + SLACK_BYTES;
  static_assert(EW == 4 || (EW == 8 && LNF && CG == 2), "8 epilogue warps: LayerNorm-fused epilogue only");
  static_assert(STAGES >= 3 && STAGES <= 8, "stage count");
  static_assert(SMEM_TOTAL <= 232448, "shared memory budget");
};

struct GemmArgs {
  int M, N, K;
  const float* bias;
  GemmLnFusion ln;
  // EPI_HEAD_DOT: the head's last Linear folded into the epilogue
  const float* head_w = nullptr;  // [nj, 256]
  const float* head_b = nullptr;  // [nj]
  float* head_out = nullptr;      // [M, nj]
  int head_nj = 0, head_relu = 0;
  // RowMap (kernels.h): when set, only the listed 256-row blocks are computed (tile index -> row_blocks[tile / n_tiles])
  const int32_t* row_blocks = nullptr;
  const int32_t* row_count = nullptr;
  int splits = 1;        // split-K: split s reduces k-blocks [s * kb_per_split, ...) and writes rows [s*M, (s+1)*M) of D
  int kb_per_split = 0;  // (0 = all)
  DropKey drop;          // train-mode forward: nn.Dropout on (acc + bias [ReLU]) before the residual add (thr = 0: off)
};

// TR selects the operand layouts (the backward GEMMs of the training step):
//   0  A [M,K] and W [N,K] both K-major (forward: y = x W^T)
//   1  A K-major, B MN-major: B is stored [K,N] (dgrad: dX[M,Kin] = dY[M,Nout] W[Nout,Kin], W as nn.Linear keeps it)
//   3  A and B MN-major: A stored [K,M], B stored [K,N] (wgrad: dW[Nout,Kin] = dY[tok,Nout]^T X[tok,Kin], reduction
//      over the tokens; rows beyond K are zero-filled by TMA)
// An MN-major operand tile is staged as 64-column blocks of BK rows x 128 B (one TMA box each); the UMMA descriptor
// walks them with LBO = BK * 128 (stride between 64-wide MN blocks) and SBO = 1024 (8 K indices).
template <int EPI, int CG, int RS, int TR, int EW>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmR,
                 const __grid_constant__ CUtensorMap tmU, const GemmArgs g) {
  using C = Cfg<EPI, CG, RS, EW>;
  constexpr int STAGES = C::STAGES;
  constexpr int SLABS = C::SLABS;
  constexpr bool OUT_F32 = (EPI == EPI_BIAS_F32 || C::RESID);
  constexpr bool LNF = C::LNF;
  constexpr bool A_MN = (TR & 2) != 0, B_MN = (TR & 1) != 0;
  static_assert(TR == 0 || (CG == 2 && !C::RESID), "transposed operands: CTA pairs, plain epilogues");
  constexpr int CPC = OUT_F32 ? 32 : 64;  // output columns per 128-byte slab row
  constexpr int CHUNKS = BN / CPC;        // slab-sized chunks per tile and warp

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  if (LNF && base - raw_addr > 512u) __trap();  // the LNF budget only leaves 512 bytes of alignment slack

  // barrier map (bytes from bar_base): full[8] @0, empty[8] @64, tfull[2] @128, tempty[2] @144,
  // tmem slot @160, resid[4][8] @192, stats[2] @448
  const uint32_t bar_base = base + C::SMEM_BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 64u + 8u * s; };
  auto tfull_bar = [&](int b) { return bar_base + 128u + 8u * b; };
  auto tempty_bar = [&](int b) { return bar_base + 144u + 8u * b; };
  auto resid_bar = [&](int ew, int s) { return bar_base + 192u + 8u * (ew * MAX_SLABS + s); };
  auto stats_bar = [&](int b) { return bar_base + (EW == 8 ? 704u : 448u) + 8u * b; };  // LNF: partner pair's row statistics have landed
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + C::SMEM_BAR_OFF + 160);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cluster_rank = CG == 2 ? int(cluster_ctarank()) : 0;  // LNF: clusters of 4 = two CTA pairs
  const int cta_rank = cluster_rank & 1;                          // 0 = leader of the pair
  const uint16_t pair_mask = uint16_t(3u << (cluster_rank & 2));
  const int group = blockIdx.x / CG;                           // tile-scheduler slot (pair index)
  const int num_groups = gridDim.x / CG;

  int m_tiles = (g.M + BM * CG - 1) / (BM * CG);  // (replaced by the RowMap's block count once it may be read)
  const int n_tiles = (g.N + BN - 1) / BN;
  int tiles_per_split = m_tiles * n_tiles;
  int total_tiles = tiles_per_split * g.splits;
  const int k_blocks = (g.K + BK - 1) / BK;
  const int kbps = g.kb_per_split > 0 ? g.kb_per_split : k_blocks;
  // tile index -> (split, m block, n block, k-block range)
  auto decode = [&](int tile, int& sp, int& m_blk, int& n_blk, int& kb0, int& kb1) {
    sp = tile / tiles_per_split;
    const int t2 = tile - sp * tiles_per_split;
    m_blk = t2 / n_tiles;
    n_blk = t2 - m_blk * n_tiles;
    if (g.row_blocks != nullptr) m_blk = __ldg(g.row_blocks + m_blk);
    kb0 = sp * kbps;
    kb1 = kb0 + kbps < k_blocks ? kb0 + kbps : k_blocks;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD);
    if (C::RESID) tma_prefetch_desc(&tmR);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), EW * CG);
    }
    for (int i = 0; i < EW * MAX_SLABS; ++i) mbar_init(bar_base + 192u + 8u * i, 1);
    // EW = 4: the partner pair's 128 threads arrive; EW = 8: this CTA's 256 epilogue threads and the partner's 256
    for (int b = 0; b < 2; ++b) mbar_init(stats_bar(b), EW == 8 ? 512 : 128);
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (CG == 2) tmem_alloc_cg2<TMEM_COLS>(base + C::SMEM_BAR_OFF + 160);
    else tmem_alloc<TMEM_COLS>(base + C::SMEM_BAR_OFF + 160);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();  // peer barriers initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlapped the previous kernel's tail; its results are needed from here on
  if (g.row_blocks != nullptr) {
    m_tiles = __ldg(g.row_count);
    tiles_per_split = m_tiles * n_tiles;
    total_tiles = tiles_per_split * g.splits;
  }

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // (the whole warp walks the loop; one elected lane issues)
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = group; tile < total_tiles; tile += num_groups) {
      int sp, m_blk, n_blk, kb0, kb1;
      decode(tile, sp, m_blk, n_blk, kb0, kb1);
      const int a_row = (m_blk * CG + cta_rank) * BM;
      const int b_row = n_blk * BN + cta_rank * C::B_ROWS;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (elect_one()) {
          const uint32_t sa = base + C::SMEM_A_OFF + stage * A_BYTES;
          const uint32_t sb = base + C::SMEM_B_OFF + stage * C::B_BYTES;
          if constexpr (CG == 2) {
            // both CTAs' bytes are credited to the leader's barrier, which the leader arms
            if (cta_rank == 0) mbar_expect_tx(full_bar(stage), 2 * C::STAGE_BYTES);
            if constexpr (A_MN) {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)
                tma_load_2d_cg2(sa + j * (BK * 128), &tmA, full_bar(stage), a_row + 64 * j, kb * BK);
            } else {
              tma_load_2d_cg2(sa, &tmA, full_bar(stage), kb * BK, a_row);
            }
            if constexpr (B_MN) {
#pragma unroll
              for (int j = 0; j < C::B_ROWS / 64; ++j)
                tma_load_2d_cg2(sb + j * (BK * 128), &tmB, full_bar(stage), b_row + 64 * j, kb * BK);
            } else {
              tma_load_2d_cg2(sb, &tmB, full_bar(stage), kb * BK, b_row);
            }
          } else {
            mbar_expect_tx(full_bar(stage), C::STAGE_BYTES);
            tma_load_2d(sa, &tmA, full_bar(stage), kb * BK, a_row);
            tma_load_2d(sb, &tmB, full_bar(stage), kb * BK, b_row);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    // Converged control flow keeps the descriptors in uniform registers; a divergent
    // single-thread loop costs ~100 cycles per tcgen05.mma issue.
    if (cta_rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM * CG, BN, A_MN, B_MN);
      constexpr uint32_t A_LBO = A_MN ? BK * 128 : 16, B_LBO = B_MN ? BK * 128 : 16;
      constexpr uint64_t A_KSTEP = A_MN ? 128 : 2, B_KSTEP = B_MN ? 128 : 2;  // 16 K indices, in 16-byte units
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = group; tile < total_tiles; tile += num_groups, ++it) {
        int sp, m_blk, n_blk, kb0, kb1;
        decode(tile, sp, m_blk, n_blk, kb0, kb1);
        const int buf = it & 1;
        const uint32_t use_parity = (uint32_t(it) >> 1) & 1u;
        mbar_wait(tempty_bar(buf), use_parity ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(buf * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = make_smem_desc_sw128(base + C::SMEM_A_OFF + stage * A_BYTES, 1024, A_LBO);
            const uint64_t db = make_smem_desc_sw128(base + C::SMEM_B_OFF + stage * C::B_BYTES, 1024, B_LBO);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {  // K-major: +32 bytes along K per step; MN-major: +16 rows of 128 B
              const uint32_t acc = (kb > kb0 || k > 0) ? 1u : 0u;
              if constexpr (CG == 2)
                mma_ss_cg2(d_tmem, da + A_KSTEP * uint64_t(k), db + B_KSTEP * uint64_t(k), idesc, acc);
              else
                mma_ss(d_tmem, da + A_KSTEP * uint64_t(k), db + B_KSTEP * uint64_t(k), idesc, acc);
            }
            if constexpr (CG == 2) {
              tc_commit_cg2(empty_bar(stage), pair_mask);  // frees the slot in BOTH CTAs once these MMAs retire
              if (kb == kb1 - 1) tc_commit_cg2(tfull_bar(buf), pair_mask);
            } else {
              tc_commit(empty_bar(stage));
              if (kb == kb1 - 1) tc_commit(tfull_bar(buf));
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;   // TMEM lane quarter this warp may access
    const int ew = warp - 2;  // slab ring / barrier set of this warp
    const int ch = ew >> 2;   // column part of the tile this warp walks (EW = 8: two warps per lane quarter)
    const uint32_t slab0 = base + C::SMEM_D_OFF + ew * SLABS * SLAB_BYTES;
    const int my_tiles = group < total_tiles ? (total_tiles - group + num_groups - 1) / num_groups : 0;
    // Slab uses per tile and warp: its residual/output chunks, then (LNF) its chunks of normalised bf16 rows.
    constexpr int WCH = CHUNKS / C::COL_SPLIT;           // fp32 chunks per warp and tile
    constexpr int WLN = (BN / 64) / C::COL_SPLIT;        // LayerNorm output chunks per warp and tile
    constexpr int UPT = WCH + (LNF ? WLN : 0);
    static_assert(!LNF || CHUNKS == 8, "LNF: 8 fp32 chunks of 32 columns per 256-column tile");
    const int total_chunks = my_tiles * UPT;
    auto tile_row0 = [&](int m_blk) { return (m_blk * CG + cta_rank) * BM + q * 32; };

    // slab use gc (per-warp counter) -> if it is a residual chunk, issue its TMA load into slab gc % SLABS
    auto issue_resid_load = [&](int gc) {
      const int cl = gc % UPT;
      if (cl >= WCH) return;  // a LayerNorm output chunk: nothing to load
      const int c = ch * WCH + cl;
      const int tile = group + (gc / UPT) * num_groups;
      int sp, m_blk, n_blk, kb0, kb1;
      decode(tile, sp, m_blk, n_blk, kb0, kb1);  // (residual epilogues never run split-K: sp = 0)
      const int s = gc % SLABS;
      mbar_expect_tx(resid_bar(ew, s), SLAB_BYTES);
      tma_load_2d(slab0 + s * SLAB_BYTES, &tmR, resid_bar(ew, s), n_blk * BN + c * CPC, tile_row0(m_blk));
    };
    if constexpr (C::RESID) {
      // (an L2 prefetch of the next tiles' residual rows was tried and measured slower: out-proj 61 -> 71 us, +35 %
      // DRAM reads — profiles/r02_notes.md)
      if (lane == 0)
        for (int gc = 0; gc < C::LOOKAHEAD && gc < total_chunks; ++gc) issue_resid_load(gc);
    }

    int it = 0;
    int gc = 0;  // chunks processed so far by this warp
    uint32_t resid_phase = 0;  // bit s = parity of the next residual load to complete in slab s
    for (int tile = group; tile < total_tiles; tile += num_groups, ++it) {
      int sp, m_blk, n_blk, kb0, kb1;
      decode(tile, sp, m_blk, n_blk, kb0, kb1);
      const int buf = it & 1;
      const uint32_t use_parity = (uint32_t(it) >> 1) & 1u;
      const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(buf * BN);
      const int row0 = tile_row0(m_blk) + sp * g.M;  // split-K partials are stacked along the rows of D
      // LNF: row statistics of the updated h over this tile's 256 columns, accumulated about a shift (the row's
      // first value) so that a large common offset of the residual stream does not cancel in E[x^2] - mean^2
      float st_sum = 0.f, st_sq = 0.f, st_shift = 0.f;
      float hd0 = 0.f, hd1 = 0.f;  // EPI_HEAD_DOT: this row's dot products with the last layer's weight rows
      mbar_wait(tfull_bar(buf), use_parity);
      tc_fence_after();

#pragma unroll 1
      for (int c = ch * WCH; c < (ch + 1) * WCH; ++c, ++gc) {
        const int s = gc % SLABS;
        const uint32_t slab = slab0 + s * SLAB_BYTES;
        // bias of this chunk: requested before the barrier waits below so that the global-load
        // latency overlaps them (fp32 epilogues: 32 columns per chunk)
        float4 bpre[8];
        if constexpr (OUT_F32) {
          if (g.bias != nullptr) {
            const float4* bp = reinterpret_cast<const float4*>(g.bias + n_blk * BN + c * CPC);
#pragma unroll
            for (int i = 0; i < 8; ++i) bpre[i] = __ldg(bp + i);
          }
        }
        if constexpr (C::RESID) {
          // keep LOOKAHEAD residual loads in flight: chunk gc+LOOKAHEAD lands in the slab chunk gc-2
          // was stored from, so all but the most recent store must have finished reading smem
          if (lane == 0 && gc + C::LOOKAHEAD < total_chunks) {
            tma_store_wait_read<1>();
            issue_resid_load(gc + C::LOOKAHEAD);
          }
          mbar_wait(resid_bar(ew, s), (resid_phase >> s) & 1u);
          resid_phase ^= 1u << s;
        } else if constexpr (EPI != EPI_HEAD_DOT) {
          // the TMA store that last read this slab (SLABS chunks ago) must have drained it
          if (lane == 0) tma_store_wait_read<SLABS - 1>();
          __syncwarp();
        }
#pragma unroll
        for (int half = 0; half < CPC / 32; ++half) {
          uint32_t v[32];
          tmem_ld32(t_row + uint32_t(c * CPC + half * 32), v);
          tmem_ld_wait();
          const int col0 = n_blk * BN + c * CPC + half * 32;
          if (g.bias != nullptr) {
            const float4* bp = reinterpret_cast<const float4*>(g.bias + col0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b4 = OUT_F32 ? bpre[i] : __ldg(bp + i);
              v[4 * i + 0] = __float_as_uint(__uint_as_float(v[4 * i + 0]) + b4.x);
              v[4 * i + 1] = __float_as_uint(__uint_as_float(v[4 * i + 1]) + b4.y);
              v[4 * i + 2] = __float_as_uint(__uint_as_float(v[4 * i + 2]) + b4.z);
              v[4 * i + 3] = __float_as_uint(__uint_as_float(v[4 * i + 3]) + b4.w);
            }
          }
          if constexpr (EPI == EPI_BIAS_RELU_BF16 || EPI == EPI_HEAD_DOT) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              v[i] = __float_as_uint(fmaxf(__uint_as_float(v[i]), 0.0f));
          }
          if constexpr (EPI == EPI_BIAS_RELU_BF16 || EPI == EPI_BIAS_RESID_F32) {
            if (g.drop.thr != 0u) {  // (uniform branch on a kernel parameter; inference launches never take it)
              const uint32_t k0 = (uint32_t(row0 + lane) * uint32_t(g.N) + uint32_t(col0)) >> 1;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                float a0 = __uint_as_float(v[2 * i]), a1 = __uint_as_float(v[2 * i + 1]);
                drop_pair(k0 + uint32_t(i), g.drop, a0, a1);
                v[2 * i] = __float_as_uint(a0);
                v[2 * i + 1] = __float_as_uint(a1);
              }
            }
          }
          const uint32_t row_addr = slab + uint32_t(lane) * 128u;
          if constexpr (OUT_F32) {
#pragma unroll
            if constexpr (C::RESID) {
              // slab currently holds the residual: read the whole row first (eight independent loads in
              // flight instead of a load -> add -> store chain per 16 bytes), then update in place
              uint32_t r[32];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const uint32_t src = row_addr + (uint32_t(i ^ (lane & 7)) << 4);
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(r[4 * i]), "=r"(r[4 * i + 1]), "=r"(r[4 * i + 2]), "=r"(r[4 * i + 3])
                             : "r"(src));
              }
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(r[i]));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint32_t dst = row_addr + (uint32_t(i ^ (lane & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v[4 * i]),
                           "r"(v[4 * i + 1]), "r"(v[4 * i + 2]), "r"(v[4 * i + 3])
                           : "memory");
            }
            if constexpr (LNF) {
              // keep the updated row in TMEM for the normalisation pass and accumulate its statistics
              tmem_st32(t_row + uint32_t(c * CPC + half * 32), v);
              if (c == ch * WCH && half == 0) st_shift = __uint_as_float(v[0]);
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float d = __uint_as_float(v[i]) - st_shift;
                st_sum += d;
                st_sq = fmaf(d, d, st_sq);
              }
            }
          } else if constexpr (EPI == EPI_HEAD_DOT) {
            // the row's 32 activations of this half against the matching slice of the last layer's weight rows
            // (every lane reads the same addresses: one broadcast request per float4)
            const float4* w0 = reinterpret_cast<const float4*>(g.head_w + col0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 w = __ldg(w0 + i);
              hd0 = fmaf(__uint_as_float(v[4 * i]), w.x, hd0);
              hd0 = fmaf(__uint_as_float(v[4 * i + 1]), w.y, hd0);
              hd0 = fmaf(__uint_as_float(v[4 * i + 2]), w.z, hd0);
              hd0 = fmaf(__uint_as_float(v[4 * i + 3]), w.w, hd0);
            }
            if (g.head_nj == 2) {
              const float4* w1 = reinterpret_cast<const float4*>(g.head_w + BN + col0);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 w = __ldg(w1 + i);
                hd1 = fmaf(__uint_as_float(v[4 * i]), w.x, hd1);
                hd1 = fmaf(__uint_as_float(v[4 * i + 1]), w.y, hd1);
                hd1 = fmaf(__uint_as_float(v[4 * i + 2]), w.z, hd1);
                hd1 = fmaf(__uint_as_float(v[4 * i + 3]), w.w, hd1);
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t p0 = pack_bf16x2(__uint_as_float(v[8 * i + 0]), __uint_as_float(v[8 * i + 1]));
              const uint32_t p1 = pack_bf16x2(__uint_as_float(v[8 * i + 2]), __uint_as_float(v[8 * i + 3]));
              const uint32_t p2 = pack_bf16x2(__uint_as_float(v[8 * i + 4]), __uint_as_float(v[8 * i + 5]));
              const uint32_t p3 = pack_bf16x2(__uint_as_float(v[8 * i + 6]), __uint_as_float(v[8 * i + 7]));
              const int chunk16 = half * 4 + i;
              const uint32_t dst = row_addr + (uint32_t(chunk16 ^ (lane & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(p0), "r"(p1),
                           "r"(p2), "r"(p3)
                           : "memory");
            }
          }
        }
        if constexpr (EPI != EPI_HEAD_DOT) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmD, slab, n_blk * BN + c * CPC, row0);
            tma_store_commit();
          }
        }
      }
      if constexpr (EPI == EPI_HEAD_DOT) {
        const int row = row0 + lane;
        if (row < g.M) {
          float o0 = hd0 + g.head_b[0];
          if (g.head_relu) o0 = fmaxf(o0, 0.0f);
          if (g.head_nj == 2) {
            float o1 = hd1 + g.head_b[1];
            if (g.head_relu) o1 = fmaxf(o1, 0.0f);
            *reinterpret_cast<float2*>(g.head_out + int64_t(row) * 2) = make_float2(o0, o1);
          } else {
            g.head_out[row] = o0;
          }
        }
      }
      if constexpr (LNF) {
        // ---- fused LayerNorm: swap (sum, sumsq) with the CTA of the partner pair that holds the other
        // 256 columns of the same rows, then normalise the rows kept in TMEM and store them as bf16
        tmem_st_wait();
        const int my_row = q * 32 + lane;
        const uint32_t partner = uint32_t(cluster_rank ^ 2);
        constexpr int WCOLS = BN / C::COL_SPLIT;  // columns behind this thread's statistics
        // this part's mean and centred sum of squares; parts are merged with the pairwise update of Chan et al. in
        // a fixed order over the column quarters, so every CTA and warp computes bit-identical statistics
        const float my_mean = st_shift + st_sum * (1.0f / float(WCOLS));
        const float my_m2 = fmaxf(st_sq - st_sum * st_sum * (1.0f / float(WCOLS)), 0.f);
        float mean, m2;
        if constexpr (EW == 8) {
          // [buffer][global column quarter][row]: written here and in the partner pair's CTA of the same rows
          const int gq_ = ((cluster_rank >> 1) & 1) * 2 + ch;
          const uint32_t slot = base + C::SMEM_STATS_OFF + uint32_t((buf * 4 + gq_) * 128 + my_row) * 8u;
          asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(slot), "f"(my_mean), "f"(my_m2) : "memory");
          mbar_arrive(stats_bar(buf));
          st_cluster_f32x2(mapa_shared(slot, partner), my_mean, my_m2);
          mbar_arrive_cluster(mapa_shared(stats_bar(buf), partner));
          for (uint32_t polls = 0; !mbar_try_wait_cluster(stats_bar(buf), use_parity);)
            if (++polls > (1u << 26)) __trap();  // a protocol bug must trap, not hang the GPU (ptx.cuh bounded-wait policy)
          float2 pq[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];"
                         : "=f"(pq[j].x), "=f"(pq[j].y)
                         : "r"(base + C::SMEM_STATS_OFF + uint32_t((buf * 4 + j) * 128 + my_row) * 8u)
                         : "memory");
          const float d01 = pq[0].x - pq[1].x, d23 = pq[2].x - pq[3].x;
          const float mA = 0.5f * (pq[0].x + pq[1].x), mB = 0.5f * (pq[2].x + pq[3].x);
          const float sA = pq[0].y + pq[1].y + d01 * d01 * (0.5f * float(WCOLS));
          const float sB = pq[2].y + pq[3].y + d23 * d23 * (0.5f * float(WCOLS));
          const float dAB = mA - mB;
          mean = 0.5f * (mA + mB);
          m2 = sA + sB + dAB * dAB * (0.5f * float(2 * WCOLS));
        } else {
          st_cluster_f32x2(mapa_shared(base + C::SMEM_STATS_OFF + uint32_t(buf * 128 + my_row) * 8u, partner), my_mean, my_m2);
          mbar_arrive_cluster(mapa_shared(stats_bar(buf), partner));
          for (uint32_t polls = 0; !mbar_try_wait_cluster(stats_bar(buf), use_parity);)
            if (++polls > (1u << 26)) __trap();  // a protocol bug must trap, not hang the GPU (ptx.cuh bounded-wait policy)
          float2 other;
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];"
                       : "=f"(other.x), "=f"(other.y)
                       : "r"(base + C::SMEM_STATS_OFF + uint32_t(buf * 128 + my_row) * 8u)
                       : "memory");
          mean = 0.5f * (my_mean + other.x);
          const float dm = my_mean - other.x;
          m2 = my_m2 + other.y + dm * dm * (0.5f * float(BN));
        }
        const float rstd = rsqrtf(m2 * (1.0f / float(2 * BN)) + g.ln.eps);
        const float nmr = -mean * rstd;
#pragma unroll 1
        for (int c2 = ch * WLN; c2 < (ch + 1) * WLN; ++c2, ++gc) {
          const int s = gc % SLABS;
          const uint32_t slab = slab0 + s * SLAB_BYTES;
          if (lane == 0) {
            if (gc + C::LOOKAHEAD < total_chunks) {
              tma_store_wait_read<1>();
              issue_resid_load(gc + C::LOOKAHEAD);
            } else {
              tma_store_wait_read<SLABS - 1>();
            }
          }
          __syncwarp();
          const uint32_t row_addr = slab + uint32_t(lane) * 128u;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t v[32];
            tmem_ld32(t_row + uint32_t(c2 * 64 + half * 32), v);
            const int col0 = n_blk * BN + c2 * 64 + half * 32;
            const float4* gp = reinterpret_cast<const float4*>(g.ln.ln_gamma + col0);
            const float4* bp = reinterpret_cast<const float4*>(g.ln.ln_beta + col0);
            float4 gq[8], bq[8];  // requested while the TMEM load is in flight
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              gq[i] = __ldg(gp + i);
              bq[i] = __ldg(bp + i);
            }
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 g0 = gq[2 * i], g1 = gq[2 * i + 1];
              const float4 b0 = bq[2 * i], b1 = bq[2 * i + 1];
              auto nrm = [&](uint32_t x, float ga, float be) { return fmaf(fmaf(__uint_as_float(x), rstd, nmr), ga, be); };
              const uint32_t p0 = pack_bf16x2(nrm(v[8 * i + 0], g0.x, b0.x), nrm(v[8 * i + 1], g0.y, b0.y));
              const uint32_t p1 = pack_bf16x2(nrm(v[8 * i + 2], g0.z, b0.z), nrm(v[8 * i + 3], g0.w, b0.w));
              const uint32_t p2 = pack_bf16x2(nrm(v[8 * i + 4], g1.x, b1.x), nrm(v[8 * i + 5], g1.y, b1.y));
              const uint32_t p3 = pack_bf16x2(nrm(v[8 * i + 6], g1.z, b1.z), nrm(v[8 * i + 7], g1.w, b1.w));
              const uint32_t dst = row_addr + (uint32_t((half * 4 + i) ^ (lane & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(p0), "r"(p1), "r"(p2), "r"(p3)
                           : "memory");
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmU, slab, n_blk * BN + c2 * 64, row0);
            tma_store_commit();
          }
        }
      }
      // all TMEM reads of this accumulator buffer are complete -> hand it back to the MMA warp
      // (of the leader CTA; both CTAs' epilogue warps report there)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_leader(tempty_bar(buf));
        else mbar_arrive(tempty_bar(buf));
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();  // the peer may still read my smem / signal my barriers
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_cg2<TMEM_COLS>(tmem_base);
    else tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

template <int EPI, int CG, int RS = 4, int TR = 0, int EW = 4>
int launch_one(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD,
               const CUtensorMap& tmR, const CUtensorMap& tmU, const GemmArgs& g, int groups,
               cudaStream_t stream) {
  using C = Cfg<EPI, CG, RS, EW>;
  constexpr int CLUSTER = C::LNF ? 4 : CG;  // LNF: two CTA pairs per cluster share full 512-column rows
  static bool configured_on[kMaxDevices];
  static int max_clusters_on[kMaxDevices];
  bool& configured = configured_on[current_device()];
  int& max_clusters = max_clusters_on[current_device()];
  if (!configured) {
    RP_CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_kernel<EPI, CG, RS, TR, EW>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_TOTAL));
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(unsigned(groups * CG));
  cfg.blockDim = dim3(C::NUM_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (C::LNF) {
    // Clusters of four do not tile all 148 SMs (GPC sizes): size the persistent grid to what can be
    // co-resident (33 clusters = 132 SMs on B200).  Pairs 2c / 2c+1 walk tiles (m, 0) / (m, 1) in step.
    if (max_clusters == 0) {
      cfg.gridDim = dim3(unsigned(num_sms() / 4 * 4));
      RP_CUDA_CHECK(cudaOccupancyMaxActiveClusters(&max_clusters, gemm_bf16_kernel<EPI, CG, RS, TR, EW>, &cfg));
      RP_CHECK(max_clusters > 0, "gemm: no cluster of 4 CTAs fits");
    }
    const int m_tiles = (g.M + BM * CG - 1) / (BM * CG);
    const int clusters = m_tiles < max_clusters ? m_tiles : max_clusters;
    cfg.gridDim = dim3(unsigned(clusters * 4));
  }
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  RP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<EPI, CG, RS, TR, EW>, tmA, tmB, tmD, tmR, tmU, g));
  count_launch();
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

template <int CG>
int launch_cg(int epilogue, const void* A, int64_t lda, const void* W, int64_t ldw, void* D, int64_t ldd,
              const float* bias, const float* resid, int64_t ldr, int M, int N, int K,
              const GemmLnFusion& ln, cudaStream_t stream, const RowMap* rows, const DropKey* drop = nullptr) {
  const bool resid_epi = (epilogue == EPI_BIAS_RESID_F32 || epilogue == EPI_BIAS_RESID_LN);
  const bool out_f32 = (epilogue == EPI_BIAS_F32 || resid_epi);
  CUtensorMap tmA, tmB, tmD, tmR, tmU;
  int rc;
  if ((rc = make_tmap_2d(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, A, K, M, lda * 2, BK, BM))) return rc;
  if ((rc = make_tmap_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, W, K, N, ldw * 2, BK, BN / CG))) return rc;
  if (out_f32)
    rc = make_tmap_2d(&tmD, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, D, N, M, ldd * 4, 32, 32);
  else
    rc = make_tmap_2d(&tmD, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, D, N, M, ldd * 2, 64, 32);
  if (rc) return rc;
  tmR = tmD;
  if (resid_epi && (rc = make_tmap_2d(&tmR, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, resid, N, M, ldr * 4, 32, 32)))
    return rc;
  tmU = tmD;
  if (epilogue == EPI_BIAS_RESID_LN &&
      (rc = make_tmap_2d(&tmU, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, ln.u_out, N, M, ln.ld_u * 2, 64, 32)))
    return rc;

  GemmArgs g{M, N, K, bias, ln};
  if (drop != nullptr) g.drop = *drop;
  if (CG == 2 && rows != nullptr) {  // (256-row blocks are the CTA pair's tile height)
    g.row_blocks = rows->blocks;
    g.row_count = rows->count;
  }
  const int total_tiles = ((M + BM * CG - 1) / (BM * CG)) * (N / BN);
  const int sms = num_sms();
  if (sms <= 0) return RP_ERR_NO_DEVICE;
  const int max_groups = sms / CG;
  const int groups = total_tiles < max_groups ? total_tiles : max_groups;
  switch (epilogue) {
    case EPI_BIAS_BF16: return launch_one<EPI_BIAS_BF16, CG>(tmA, tmB, tmD, tmR, tmU, g, groups, stream);
    case EPI_BIAS_RELU_BF16: return launch_one<EPI_BIAS_RELU_BF16, CG>(tmA, tmB, tmD, tmR, tmU, g, groups, stream);
    case EPI_BIAS_F32: return launch_one<EPI_BIAS_F32, CG>(tmA, tmB, tmD, tmR, tmU, g, groups, stream);
    case EPI_BIAS_RESID_F32: return launch_one<EPI_BIAS_RESID_F32, CG>(tmA, tmB, tmD, tmR, tmU, g, groups, stream);
    case EPI_BIAS_RESID_LN:
      if constexpr (CG == 2) {
        // memory-bound instances (K <= 512: out-proj): eight epilogue warps with three slabs each and three operand
        // stages (92 -> 84 us); compute-bound ones (FF2) keep four warps, four slabs and five stages
        if (K <= 512) return launch_one<EPI_BIAS_RESID_LN, 2, 3, 0, 8>(tmA, tmB, tmD, tmR, tmU, g, groups, stream);
        return launch_one<EPI_BIAS_RESID_LN, 2, 4>(tmA, tmB, tmD, tmR, tmU, g, groups, stream);
      }
      set_last_error("gemm: the LayerNorm-fused epilogue needs CTA pairs");
      return RP_ERR_INVALID;
    default: set_last_error("gemm: unknown epilogue %d", epilogue); return RP_ERR_INVALID;
  }
}

// Backward GEMMs (operand layouts TR = 1 / 3, see the kernel): plain fp32 / bf16 epilogues, no bias, CTA pairs,
// optional split-K whose partial results are stacked along the rows of D ([splits * M, N]).
int launch_tr(int tr, bool out_f32, const void* A, int64_t lda, const void* B, int64_t ldb, void* D, int64_t ldd,
              int M, int N, int K, int splits, cudaStream_t stream) {
  CUtensorMap tmA, tmB, tmD;
  int rc;
  const CUtensorMapDataType bf = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  if (tr & 2) rc = make_tmap_2d(&tmA, bf, A, M, K, lda * 2, 64, BK);   // stored [K, M]
  else rc = make_tmap_2d(&tmA, bf, A, K, M, lda * 2, BK, BM);          // stored [M, K]
  if (rc) return rc;
  if ((rc = make_tmap_2d(&tmB, bf, B, N, K, ldb * 2, 64, BK))) return rc;  // stored [K, N]
  if (out_f32)
    rc = make_tmap_2d(&tmD, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, D, N, uint64_t(M) * splits, ldd * 4, 32, 32);
  else
    rc = make_tmap_2d(&tmD, bf, D, N, uint64_t(M) * splits, ldd * 2, 64, 32);
  if (rc) return rc;
  GemmArgs g{M, N, K, nullptr, GemmLnFusion{}};
  const int k_blocks = (K + BK - 1) / BK;
  g.splits = splits;
  g.kb_per_split = (k_blocks + splits - 1) / splits;
  const int total_tiles = ((M + BM * 2 - 1) / (BM * 2)) * ((N + BN - 1) / BN) * splits;
  const int sms = num_sms();
  if (sms <= 0) return RP_ERR_NO_DEVICE;
  const int groups = total_tiles < sms / 2 ? total_tiles : sms / 2;
  if (tr == 1) {
    if (out_f32) return launch_one<EPI_BIAS_F32, 2, 4, 1>(tmA, tmB, tmD, tmD, tmD, g, groups, stream);
    return launch_one<EPI_BIAS_BF16, 2, 4, 1>(tmA, tmB, tmD, tmD, tmD, g, groups, stream);
  }
  if (out_f32) return launch_one<EPI_BIAS_F32, 2, 4, 3>(tmA, tmB, tmD, tmD, tmD, g, groups, stream);
  return launch_one<EPI_BIAS_BF16, 2, 4, 3>(tmA, tmB, tmD, tmD, tmD, g, groups, stream);
}

}  // namespace

int launch_gemm_bwd(int kind, bool out_f32, const void* A, int64_t lda, const void* B, int64_t ldb, void* D,
                    int64_t ldd, int M, int N, int K, int splits, cudaStream_t stream) {
  RP_CHECK(kind == 1 || kind == 3, "gemm_bwd: kind must be 1 (dgrad) or 3 (wgrad)");
  RP_CHECK(M > BM && N > 0 && K > 0, "gemm_bwd: needs more than %d output rows (M=%d N=%d K=%d)", BM, M, N, K);
  RP_CHECK(N % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && (ldd * (out_f32 ? 4 : 2)) % 16 == 0,
           "gemm_bwd: N and the pitches must keep rows 16-byte aligned");
  RP_CHECK(kind == 3 || K % BK == 0, "gemm_bwd: dgrad needs K %% %d == 0", BK);
  RP_CHECK(splits >= 1 && splits <= (K + BK - 1) / BK, "gemm_bwd: bad split count %d", splits);
  {
    // every split must own at least one k-block (an empty one would never complete its accumulator barrier)
    const int kb = (K + BK - 1) / BK, per = (kb + splits - 1) / splits;
    RP_CHECK((splits - 1) * per < kb, "gemm_bwd: %d splits of %d k-blocks leave the last split empty (use %d)", splits,
             per, (kb + per - 1) / per);
  }
  RP_CHECK(splits == 1 || M % (2 * BM) == 0, "gemm_bwd: split-K needs M %% %d == 0", 2 * BM);
  RP_CHECK((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(D)) % 16 == 0,
           "gemm_bwd: pointers must be 16-byte aligned");
  return launch_tr(kind, out_f32, A, lda, B, ldb, D, ldd, M, N, K, splits, stream);
}

int launch_gemm_head_dot(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const float* w7,
                         const float* b7, int nj, bool final_relu, float* out, int M, int K, cudaStream_t stream,
                         const RowMap* rows) {
  RP_CHECK(A && W && bias && w7 && b7 && out, "gemm_head_dot: null argument");
  RP_CHECK(M > 0 && K > 0 && K % BK == 0 && lda % 8 == 0 && ldw % 8 == 0, "gemm_head_dot: K %% 64 and pitches %% 8 required");
  RP_CHECK(nj == 1 || nj == 2, "gemm_head_dot: nj must be 1 or 2");
  RP_CHECK((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(bias) |
            reinterpret_cast<uintptr_t>(w7)) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 8 == 0,
           "gemm_head_dot: pointers must be 16-byte aligned (out: 8)");
  const int N = BN;  // the head's hidden width: one 256-column tile, so an epilogue thread sees its whole row
  CUtensorMap tmA, tmB;
  int rc;
  const bool pair = M > BM;
  if ((rc = make_tmap_2d(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, A, K, M, lda * 2, BK, BM))) return rc;
  if ((rc = make_tmap_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, W, K, N, ldw * 2, BK, pair ? BN / 2 : BN))) return rc;
  GemmArgs g{M, N, K, bias, GemmLnFusion{}};
  g.head_w = w7; g.head_b = b7; g.head_out = out; g.head_nj = nj; g.head_relu = final_relu ? 1 : 0;
  const int sms = num_sms();
  if (sms <= 0) return RP_ERR_NO_DEVICE;
  if (pair) {
    if (rows != nullptr) {
      g.row_blocks = rows->blocks;
      g.row_count = rows->count;
    }
    const int tiles = (M + 2 * BM - 1) / (2 * BM);
    return launch_one<EPI_HEAD_DOT, 2>(tmA, tmB, tmA, tmA, tmA, g, tiles < sms / 2 ? tiles : sms / 2, stream);
  }
  return launch_one<EPI_HEAD_DOT, 1>(tmA, tmB, tmA, tmA, tmA, g, 1, stream);
}

namespace {
int launch_gemm_checked(int epilogue, const void* A, int64_t lda, const void* W, int64_t ldw, void* D,
                        int64_t ldd, const float* bias, const float* resid, int64_t ldr, int M, int N, int K,
                        const GemmLnFusion& ln, cudaStream_t stream, const RowMap* rows, const DropKey* drop) {
  RP_CHECK(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  RP_CHECK(N % BN == 0, "gemm: N=%d must be a multiple of %d", N, BN);
  RP_CHECK(K % BK == 0, "gemm: K=%d must be a multiple of %d", K, BK);
  RP_CHECK(lda % 8 == 0 && ldw % 8 == 0, "gemm: lda/ldw must be multiples of 8 elements");
  const bool resid_epi = (epilogue == EPI_BIAS_RESID_F32 || epilogue == EPI_BIAS_RESID_LN);
  const bool out_f32 = (epilogue == EPI_BIAS_F32 || resid_epi);
  RP_CHECK((ldd * (out_f32 ? 4 : 2)) % 16 == 0, "gemm: output pitch must be 16-byte aligned");
  RP_CHECK(!resid_epi || (resid != nullptr && ldr % 4 == 0),
           "gemm: residual epilogue needs a 16-byte aligned residual");
  RP_CHECK(epilogue != EPI_BIAS_RESID_LN ||
               (N == 2 * BN && M > BM && ln.ln_gamma && ln.ln_beta && ln.u_out && ln.ld_u % 8 == 0),
           "gemm: the LayerNorm-fused epilogue needs N = 512, M > 128, gamma/beta and a bf16 output");
  RP_CHECK((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) |
            reinterpret_cast<uintptr_t>(D) | reinterpret_cast<uintptr_t>(bias) |
            reinterpret_cast<uintptr_t>(resid)) % 16 == 0,
           "gemm: pointers must be 16-byte aligned");
  // CTA pairs (cta_group::2) need two 128-row blocks; a problem of at most 128 rows runs the single-CTA kernel
  if (M <= BM && epilogue != EPI_BIAS_RESID_LN)
    return launch_cg<1>(epilogue, A, lda, W, ldw, D, ldd, bias, resid, ldr, M, N, K, ln, stream, nullptr, drop);
  return launch_cg<2>(epilogue, A, lda, W, ldw, D, ldd, bias, resid, ldr, M, N, K, ln, stream, rows, drop);
}
}  // namespace

int launch_gemm_ln(int epilogue, const void* A, int64_t lda, const void* W, int64_t ldw, void* D,
                   int64_t ldd, const float* bias, const float* resid, int64_t ldr, int M, int N, int K,
                   const GemmLnFusion& ln, cudaStream_t stream, const RowMap* rows) {
  return launch_gemm_checked(epilogue, A, lda, W, ldw, D, ldd, bias, resid, ldr, M, N, K, ln, stream, rows, nullptr);
}

int launch_gemm_dropout(int epilogue, const void* A, int64_t lda, const void* W, int64_t ldw, void* D,
                        int64_t ldd, const float* bias, const float* resid, int64_t ldr, int M, int N, int K,
                        const DropKey& drop, cudaStream_t stream) {
  RP_CHECK(epilogue == EPI_BIAS_RELU_BF16 || epilogue == EPI_BIAS_RESID_F32,
           "gemm_dropout: dropout follows the ReLU (epilogue 1) or precedes the residual add (epilogue 3)");
  RP_CHECK(int64_t(M) * N < (int64_t(1) << 32), "gemm_dropout: M * N must stay below 2^32 elements");
  return launch_gemm_checked(epilogue, A, lda, W, ldw, D, ldd, bias, resid, ldr, M, N, K, GemmLnFusion{}, stream,
                             nullptr, &drop);
}

int launch_gemm(int epilogue, const void* A, int64_t lda, const void* W, int64_t ldw, void* D,
                int64_t ldd, const float* bias, const float* resid, int64_t ldr, int M, int N, int K,
                cudaStream_t stream, const RowMap* rows) {
  return launch_gemm_ln(epilogue, A, lda, W, ldw, D, ldd, bias, resid, ldr, M, N, K, GemmLnFusion{},
                        stream, rows);
}

}  // namespace rp
