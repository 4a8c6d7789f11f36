// Internal launcher declarations (C++ side of the C-ABI in include/repurpose_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dropout.cuh"

namespace rp {

// ---- GEMM: D[M,N] = A[M,K] (bf16, K-major) * W[N,K]^T (bf16, K-major) + epilogue --------------
enum GemmEpilogue : int {
  EPI_BIAS_BF16 = 0,       // D(bf16) = acc + bias
  EPI_BIAS_RELU_BF16 = 1,  // D(bf16) = relu(acc + bias)
  EPI_BIAS_F32 = 2,        // D(f32)  = acc + bias
  EPI_BIAS_RESID_F32 = 3,  // D(f32)  = acc + bias + R   (R may alias D)
  EPI_BIAS_RESID_LN = 4,   // same, plus U(bf16) = LayerNorm(D row; gamma, beta) — N must be 512 (GemmLnFusion::u_out)
  EPI_HEAD_DOT = 5,        // out[M, nj] = (relu)(relu(acc + bias) . w7[j] + b7[j]): the last two layers of a head in one
                           // kernel, N must be 256 (launch_gemm_head_dot); nothing of size [M, 256] is written
};
// Valid 256-row blocks of a padded [B, T] token matrix (launch_row_map): GEMMs given a RowMap only compute those blocks,
// i.e. skip blocks that hold nothing but padding rows of short videos.
struct RowMap {
  const int32_t* blocks = nullptr;  // [count] ascending block indices (device)
  const int32_t* count = nullptr;   // device scalar
};
int launch_gemm(int epilogue, const void* A, int64_t lda, const void* W, int64_t ldw, void* D,
                int64_t ldd, const float* bias, const float* resid, int64_t ldr, int M, int N, int K,
                cudaStream_t stream, const RowMap* rows = nullptr);
// the same with nn.Dropout applied to (acc + bias [ReLU]) before the residual is added (dropout.cuh; element index =
// row * N + column): the train-mode forward of out_proj / linear1 / linear2 / the head layers
int launch_gemm_dropout(int epilogue, const void* A, int64_t lda, const void* W, int64_t ldw, void* D,
                        int64_t ldd, const float* bias, const float* resid, int64_t ldr, int M, int N, int K,
                        const DropKey& drop, cudaStream_t stream);

// cls_head / reg_head tail (models/MMCTransformer.py:71-93): Linear(256,256) + ReLU + Linear(256, nj) (+ final ReLU for
// reg_head) fused: out[m, j] = act(sum_c relu(A[m,:] . W[c,:] + bias[c]) * w7[j, c] + b7[j]), nj = 1 or 2, fp32 out.
int launch_gemm_head_dot(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const float* w7,
                         const float* b7, int nj, bool final_relu, float* out, int M, int K, cudaStream_t stream,
                         const RowMap* rows = nullptr);

// Arguments of the LayerNorm-fused residual epilogue:
struct GemmLnFusion {
  float eps = 1e-5f;
  // EPI_BIAS_RESID_LN: the LayerNorm that follows the residual update is computed by the GEMM itself.
  // Clusters of four CTAs (two CTA pairs) cover the full 512-column rows of a 256-row block; the pairs
  // swap per-row (sum, sum of squares) through distributed shared memory, so no extra pass over h.
  const float* ln_gamma = nullptr;   // [512]
  const float* ln_beta = nullptr;    // [512]
  void* u_out = nullptr;             // bf16 [M, 512]
  int64_t ld_u = 0;
};
int launch_gemm_ln(int epilogue, const void* A, int64_t lda, const void* W, int64_t ldw, void* D,
                   int64_t ldd, const float* bias, const float* resid, int64_t ldr, int M, int N, int K,
                   const GemmLnFusion& ln, cudaStream_t stream, const RowMap* rows = nullptr);

// Backward GEMMs of the training step, bf16 operands, fp32 accumulation, no bias:
//   kind 1 (dgrad)  D[M,N] = A[M,K] * B[K,N]      A row-major [M,K] (dY), B row-major [K,N] (an nn.Linear weight
//                                                  [out=K, in=N] as stored) -> dX
//   kind 3 (wgrad)  D[M,N] = A[K,M]^T * B[K,N]    A row-major [K,M] (dY over K tokens), B row-major [K,N] (X) -> dW;
//                                                  K need not be a multiple of 64 (TMA zero-fills the tail)
// out_f32: D fp32, else bf16.  splits > 1: split-K, D must hold [splits * M, N]; partial s in rows [s*M, (s+1)*M).
int launch_gemm_bwd(int kind, bool out_f32, const void* A, int64_t lda, const void* B, int64_t ldb, void* D,
                    int64_t ldd, int M, int N, int K, int splits, cudaStream_t stream);

// ---- fused multi-head attention (d_k = 64) -----------------------------------------------------
// q/k/v: bf16, row pitch ld* elements, batch pitch bs* elements; head h occupies columns
// [h*64, h*64+64) of each row.  q is expected pre-scaled by log2(e)/sqrt(d_k).
// mask_mode 0: key-padding from kv_lens[b] (keys >= len get -inf, reference nn.MultiheadAttention
//              semantics).  kv_lens == nullptr means all Tk keys valid.
// mask_mode 1: explicit uint8 mask[b, (q), k] (0 = masked_fill(-1e9), models/transformer.py:70-72);
//              mask_q_stride == 0 broadcasts over queries.
struct FmhaArgs {
  const void* q; const void* k; const void* v; void* o;
  int64_t ldq, ldk, ldv, ldo;
  int64_t bsq, bsk, bsv, bso;
  int B, H, Tq, Tk;
  const int32_t* kv_lens;
  int mask_mode;
  const uint8_t* mask; int64_t mask_b_stride, mask_q_stride;
  float* lse = nullptr;  // optional output [B, H, Tq] fp32: log2-domain log-sum-exp per score row (training forward)
  // mask_mode 0 only: query tiles that start at or beyond round_up(kv_lens[b], 128) are padding (self-attention over a
  // padded batch): their CTAs exit at once and their output rows are left untouched
  bool skip_padded_queries = false;
  // training forward with attention-weight dropout (mask_mode 0 only): keep bits [B*H*Tq rows, drop_ld words], bit
  // (k & 31) of word k >> 5 = key k of that row kept (launch_attn_dropout_bits); the row sum runs over all weights,
  // P V over the kept ones, the output is scaled by drop_scale = 1 / (1 - p)
  const uint32_t* drop_bits = nullptr;
  int64_t drop_ld = 0;
  float drop_scale = 1.0f;
  // optional [B] permutation of the batch indices (device): CTAs are laid out over batch_order[0], batch_order[1], ... —
  // longest videos first (launch_row_map), so that the short CTAs of a ragged batch fill the end of the launch
  const int32_t* batch_order = nullptr;
};
int launch_fmha(const FmhaArgs& a, cudaStream_t stream);
// keep bits of one attention-dropout site: n_words words, word w covers elements 32 w .. 32 w + 31 of the site
int launch_attn_dropout_bits(uint32_t* bits, int64_t n_words, const DropKey& drop, cudaStream_t stream);
// keep[i] = 1 / 0 for the first n elements of an element-wise site (what the tests feed to the autograd reference)
int launch_dropout_mask_u8(uint8_t* keep, int64_t n, const DropKey& drop, cudaStream_t stream);

// ---- memory-bound kernels ------------------------------------------------------------------------
// concat(vis, aud, txt) fp32 -> bf16 [M, Cv+Ca+Ct]
int launch_concat_cast(const float* vis, const float* aud, const float* txt, int Cv, int Ca, int Ct,
                       void* out_bf16, int64_t M, cudaStream_t stream);
// GPU collate: per-video features stored back to back (no padding) -> padded bf16
// [B,T,Cv+Ca+Ct]; rows t >= lens[b] (and text rows t >= txt_lens[b]) are zero, exactly what
// dataset/RepurposeClip.py:450-485 pads with.
int launch_ragged_concat_cast(const void* vis, const void* aud, const void* txt, bool in_bf16, int Cv, int Ca,
                              int Ct, const int32_t* row_off, const int32_t* txt_off, const int32_t* txt_lens,
                              const int32_t* lens, int B, int T, void* out_bf16, cudaStream_t stream);
// lens[b] = number of non-zero bytes of mask[b, 0..T); *not_aligned = 1 if some mask is not of the form
// t < lens[b] (both device pointers)
int launch_mask_lens(const uint8_t* mask, int B, int T, int32_t* lens, int32_t* not_aligned, cudaStream_t stream);
// blocks / count of the RowMap of a padded [B, T] batch: block m (rows 256 m .. 256 m + 255 of the flattened [B*T]
// matrix) is valid iff it holds a row (b, t) with t < min(T, round_up(lens[b], 128)) — every row an attention key tile
// of a valid query can touch is computed, whole-padding blocks are not
int launch_row_map(const int32_t* lens, int B, int T, int32_t* blocks, int32_t* count, int32_t* order, cudaStream_t stream);
// zero the rows t >= lens[b] of the three forward outputs (padded steps carry no information; with block skipping
// they would otherwise hold whatever the buffers held before)
int launch_zero_padded_rows(float* logits, float* offsets, float* feats, const int32_t* lens, int B, int T, int D,
                            cudaStream_t stream);
// generic fp32 -> bf16 cast of a contiguous buffer (n % 8 == 0)
int launch_cast_bf16(const float* in, void* out_bf16, int64_t n, cudaStream_t stream);

// LayerNorm over rows of 512 (eps 1e-5, biased variance), one warp per row.
//   mode 0: y(bf16) = LN(x; g0,b0)
//   mode 1: h(f32)  = LN(x; g0,b0) + pe[row % T];  y(bf16) = LN(h; g1,b1)
//   mode 2: f(f32)  = relu(LN(x; g0,b0));  y(bf16) = LN(f; g1,b1);  y2(bf16) = LN(f; g2,b2)
//   mode 3: y(f32)  = LN(x; g0,b0)                (used by tests / MHA module)
struct LnArgs {
  const float* x; int64_t M; int T;
  const float* g0; const float* b0; const float* g1; const float* b1; const float* g2; const float* b2;
  const float* pe;
  float* out_f32; void* y_bf16; void* y2_bf16;
  float eps;
  DropKey drop;  // mode 2 only: feature_map's Dropout on f (before the head LayerNorms), element index = row * 512 + column
};
int launch_layernorm512(int mode, const LnArgs& a, cudaStream_t stream);

// final head projections: logits[M] = a_c[M,256].w_c + b_c ; offs[M,2] = relu(a_r[M,256].W_r^T + b_r)
int launch_head_out(const void* a_cls_bf16, const void* a_reg_bf16, const float* w_cls,
                    const float* b_cls, const float* w_reg, const float* b_reg, float* logits,
                    float* offsets, int64_t M, cudaStream_t stream);

// ---- decode + Soft-NMS (one CTA per video) -------------------------------------------------------
struct DecodeCfg {
  int pre_nms_topk;
  float pre_nms_thresh, duration_thresh, duration_thresh_max;
  float nms_sigma, min_score;
};
// logits [B,T], offsets [B,T,2], lens [B] (valid steps), max_seg [B].
// outputs (Kcap slots per video): segs [B,Kcap,2] f32, scores [B,Kcap] f32 (original probability),
// dscores [B,Kcap] f32 (decayed), labels [B,Kcap] i32, counts [B] i32, ncand [B] i32.
// Optional (all or none): cand_segs [B,C,2], cand_scores [B,C], cand_labels [B,C] with
// C = min(pre_nms_topk, T): the pre-NMS candidate list in descending score order.
int launch_decode_nms(const float* logits, const float* offsets, const int32_t* lens,
                      const int32_t* max_seg, int B, int T, const DecodeCfg& cfg, int Kcap,
                      float* segs, float* scores, float* dscores, int32_t* labels, int32_t* counts,
                      int32_t* ncand, float* cand_segs, float* cand_scores, int32_t* cand_labels,
                      cudaStream_t stream);
// Soft-NMS alone on caller candidates: scores [B,Nmax], segs [B,Nmax,2], n [B], max_seg [B]
// -> keep [B,Kcap] i32 (original indices, selection order), kscores [B,Kcap] (decayed), counts [B].
int launch_soft_nms(const float* scores, const float* segs, const int32_t* n, const int32_t* max_seg,
                    int B, int Nmax, float sigma, float thresh, int Kcap, int32_t* keep,
                    float* kscores, int32_t* counts, cudaStream_t stream);

// ---- AtIoU on the device (utils/metrics.py:82-111 + inference.py:45-55), float64, bit-exact ----------
// slots [n,1+4K] f32 (scheduler layout), gt [n,Gmax,2] f64, gt_counts [n], thresholds [n_thr] f64 ->
// per_video [n,n_thr] f64 precision, out [n_thr+1] f64 (per-threshold means, then their mean = AtIoU).
int launch_atiou(const float* slots, int n_videos, int K, const double* gt, const int32_t* gt_counts,
                 int Gmax, const double* thresholds, int n_thr, double* per_video, double* out,
                 cudaStream_t stream);

// masked sigmoid focal loss, summed (forward of MMCTransformer.losses); scratch: >= 296 doubles
int launch_focal_loss_sum(const float* logits, const float* targets, const uint8_t* mask, int64_t n, float alpha,
                          float gamma, double* scratch, float* out, cudaStream_t stream);

// ---- first pieces of the training step: see train.cu ------------------------------------
// dlogits[i] = mask[i] ? scale * d focal(logits[i], targets[i]) / d logits[i] : 0
int launch_focal_loss_grad(const float* logits, const float* targets, const uint8_t* mask, int64_t n, float alpha,
                           float gamma, float scale, float* dlogits, cudaStream_t stream);
// LayerNorm over rows of 512: dx [M,512], dgamma [512], dbeta [512] from x, dy, gamma; scratch: fp32,
// layernorm512_bwd_scratch_floats() values
int64_t layernorm512_bwd_scratch_floats();
// accumulate: dx += (the residual stream's gradient); dx_bf16 (optional): bf16 copy of the final dx; dx_colsum
// (optional): column sums [512] of the final dx = the bias gradient of the Linear whose output gradient dx is
// drop (optional): dx is the output gradient of a residual branch that ends in nn.Dropout (dropout1 / dropout2): dx_bf16
// and dx_colsum then hold dropout's backward of dx (kept ? dx * scale : 0) — the dY of that branch's Linear — while dx
// itself stays the residual stream's gradient
int launch_layernorm512_bwd(const float* x, const float* dy, const float* gamma, int64_t M, float eps, float* dx,
                            float* dgamma, float* dbeta, float* scratch, cudaStream_t stream, bool accumulate = false,
                            void* dx_bf16 = nullptr, float* dx_colsum = nullptr, const DropKey* drop = nullptr);
// one torch.optim.Adam step (L2 weight decay added to the gradient) on a flat fp32 buffer; p_bf16 (optional): bf16 copy
int launch_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                     float eps, float weight_decay, int step, void* p_bf16, cudaStream_t stream);


// ---- the rest of the training step's backward pass (train.cu, fmha_bwd.cu, gemm.cu) ---------------------------
// out[i] = sum_s part[s * n + i] in order (split-K partials of launch_gemm_bwd -> gradient)
int launch_splitk_reduce(const float* part, int splits, int64_t n, float* out, cudaStream_t stream);
// out[N] = column sums of bf16 x[M, N] (bias gradients); scratch: train_scratch_floats() fp32 values
int64_t train_scratch_floats();
int launch_colsum_bf16(const void* x, int64_t M, int N, float* out, float* scratch, cudaStream_t stream);
// dy = act > 0 ? dy * scale : 0, in place; f32: both fp32, else both bf16; n % 8 == 0.  scale = 1 for a plain ReLU;
// 1 / (1 - p) when `act` is dropout(relu(.)) as the train-mode forward stores it (kept and positive <=> act > 0)
int launch_relu_bwd(void* dy, const void* act, int64_t n, bool f32, float scale, cudaStream_t stream);
// the same for bf16 [M, N] fused with the bias gradient: colsum[N] = column sums of the masked dy
int launch_relu_bwd_colsum(void* dy, const void* act, int64_t M, int N, float colscale, float* colsum, float* scratch,
                           cudaStream_t stream);
// last cls_head layer: da2[M,256] bf16 = dlogit[m] w[c] * scale where a2 > 0, dw[256], db[1]
int launch_head_out_bwd(const float* dlogit, const void* a2, const float* w, int64_t M, float scale, void* da2, float* dw,
                        float* db, float* scratch, cudaStream_t stream);
// attention backward: q/k/v [B,T,ld_qkv] (head h in columns h*64..), o / dO dense [B,T,H*64], lse from the forward,
// dsum scratch [B,H,T]; writes dq (w.r.t. the unscaled q), dk, dv with row pitch ld_dqkv
struct FmhaBwdArgs {
  const void* q; const void* k; const void* v; const void* o; const void* d_o;
  const float* lse; float* dsum;
  void* dq; void* dk; void* dv;
  int64_t ld_qkv, ld_o, ld_dqkv;
  int B, H, T;
  const int32_t* kv_lens;
  // attention-weight dropout of the forward (FmhaArgs::drop_bits): P_d = keep o P * scale feeds dV, dP = keep o dP_d * scale
  const uint32_t* drop_bits = nullptr;
  int64_t drop_ld = 0;
  float drop_scale = 1.0f;
};
int launch_fmha_bwd(const FmhaBwdArgs& a, cudaStream_t stream);
// 1: the two deterministic kernels (dQ, then dK/dV; S and dP recomputed); 0: the fused kernel whose dQ is summed over the
// key tiles by fp32 adds in L2 (no fixed order); -1: the default (RP_FMHA_BWD_FUSED, fused unless it is 0)
void set_fmha_bwd_deterministic(int on);

}  // namespace rp
