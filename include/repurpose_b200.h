/* repurpose_b200 — C ABI of the B200 (sm_100a) implementation of the Repurpose inference hot path.
 *
 * The reference (YosubShin/Repurpose) is 100 % Python and has no FFI of its own; its "operator
 * interface" for this path is three Python callables plus the state-dict schema.  Each entry point
 * below names the reference call it replaces (paths relative to the reference repo):
 *
 *   rp_forward        <- MMCTransformer.forward            models/MMCTransformer.py:109-151
 *   rp_decode_nms     <- inference_single_video + the per-video loop of inference_
 *                                                          models/MMCTransformer.py:181-229, 248-273
 *   rp_soft_nms       <- soft_nms_intervals_cpu            models/softnms.py:3-38
 *   rp_fmha (+ rp_gemm_bf16 / rp_cast_bf16)
 *                     <- MultiHeadAttention.forward        models/transformer.py:52-81
 *   rp_atiou          <- calculate_tiou + averaging      utils/metrics.py:82-111, inference.py:45-55
 *   rp_load_weight    <- model.load_state_dict(ckpt['model'])  inference.py:33-34 (same key names)
 *
 * Conventions: every pointer is a DEVICE pointer unless stated otherwise; sizes are explicit; the
 * last argument is a cudaStream_t passed as void*; the return value is 0 on success, non-zero on
 * failure (rp_last_error() returns a thread-local message).  Nothing throws or aborts across the
 * ABI, nothing is allocated behind the caller's back except inside rp_create (weights) — activations
 * live in a caller-provided workspace.  There is no CPU fallback: without an sm_100 device the
 * compute entry points return RP_ERR_NO_DEVICE / a CUDA error.
 */
#ifndef REPURPOSE_B200_H_
#define REPURPOSE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RP_ABI_VERSION 1

enum {
  RP_OK = 0,
  RP_ERR_INVALID = 1,
  RP_ERR_CUDA = 2,
  RP_ERR_NO_DEVICE = 3,
  RP_ERR_WORKSPACE = 4
};

typedef struct rp_handle rp_handle;

/* models/MMCTransformer.py:26 constructor arguments that shape the graph
 * (text_num_layers / cross_num_layers are accepted and ignored by the reference). */
typedef struct rp_model_cfg {
  int32_t vis_dim, aud_dim, text_dim; /* configs/Repurpose.yaml:22-32 : 512 / 2048 / 384 */
  int32_t d_model;                    /* 512 (the kernels are specialised for 512)          */
  int32_t num_layers;                 /* self_num_layers, 16                                 */
  int32_t num_heads;                  /* 8  (head dim must be 64)                            */
  int32_t d_ff;                       /* 2048                                                */
  int32_t head_hidden;                /* 256 (models/MMCTransformer.py:60)                   */
  int32_t max_len;                    /* rows of positional_encoding.pe the caller will load */
} rp_model_cfg;

/* configs/Repurpose.yaml:52-61 test_cfg (max_seg_per_min is applied by the host: it only feeds
 * max_seg[b] = ceil((duration // 60) * max_seg_per_min), models/MMCTransformer.py:255-257). */
typedef struct rp_decode_cfg {
  int32_t pre_nms_topk;
  float pre_nms_thresh;
  float duration_thresh;
  float duration_thresh_max;
  float nms_sigma;
  float min_score;
} rp_decode_cfg;

int32_t rp_abi_version(void);
const char* rp_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py: gpu_launches) */
int64_t rp_launch_count(void);

/* ---- model handle ------------------------------------------------------------------------- */
int32_t rp_create(const rp_model_cfg* cfg, rp_handle** out);
void rp_destroy(rp_handle* h);
/* name = reference state-dict key (e.g. "multimodal_encoder.layers.3.self_attn.in_proj_weight");
 * src = fp32 device tensor with the reference shape, numel elements.  Matrices are repacked to bf16
 * (the query rows of in_proj are pre-scaled by log2(e)/sqrt(64)); vectors stay fp32. */
int32_t rp_load_weight(rp_handle* h, const char* name, const float* src, int64_t numel, void* stream);
/* 0 when every tensor of the schema has been loaded, else RP_ERR_INVALID (+ message naming one) */
int32_t rp_weights_complete(const rp_handle* h);
int64_t rp_workspace_bytes(const rp_handle* h, int32_t B, int32_t T);
/* vis [B,T,vis_dim] aud [B,T,aud_dim] txt [B,T,text_dim] fp32; lens [B] int32 valid steps
 * (masks[b,0,t] = t < lens[b]); outputs logits [B,T] (== [B,T,1]), offsets [B,T,2], feats [B,T,512]. */
int32_t rp_forward(rp_handle* h, const float* vis, const float* aud, const float* txt,
                   const int32_t* lens, int32_t B, int32_t T, float* out_logits, float* out_offsets,
                   float* out_feats, void* workspace, int64_t workspace_bytes, void* stream);

/* Same forward from UNPADDED per-video features (GPU collate): vis / aud hold the videos'
 * rows back to back ([sum(lens), dim]; video b starts at row row_off[b]), txt likewise from txt_off[b]
 * with txt_lens[b] rows available (text may be shorter than visual: dataset/RepurposeClip.py:975-980).
 * Padding rows are materialised as zeros on the device, as preprocessing() does on the host
 * (dataset/RepurposeClip.py:450-485).  Outputs are padded [B,T,...] exactly like rp_forward. */
int32_t rp_forward_ragged(rp_handle* h, const float* vis, const float* aud, const float* txt,
                          const int32_t* row_off, const int32_t* txt_off, const int32_t* txt_lens,
                          const int32_t* lens, int32_t B, int32_t T, float* out_logits,
                          float* out_offsets, float* out_feats, void* workspace,
                          int64_t workspace_bytes, void* stream);

/* Same as rp_forward_ragged for feature rows stored as bf16 (feature files converted once, offline):
 * the first thing the forward does with fp32 features is round them to bf16 (concat + cast feeding the
 * bf16 input projection), so results are bit-identical to rp_forward_ragged on the fp32 originals
 * while the host->device copy — the end-to-end bound of this path — moves half the bytes
 * (replaces the fp32 `.npy` load of dataset/RepurposeClip.py:415-446 for pre-converted
 * features). vis/aud/txt: bf16 [sum T, C]. */
int32_t rp_forward_ragged_bf16(rp_handle* h, const void* vis, const void* aud, const void* txt,
                          const int32_t* row_off, const int32_t* txt_off, const int32_t* txt_lens,
                          const int32_t* lens, int32_t B, int32_t T, float* out_logits,
                          float* out_offsets, float* out_feats, void* workspace,
                          int64_t workspace_bytes, void* stream);

/* Optional in-situ profiler: between begin and end every kernel rp_forward launches is bracketed
 * by CUDA events on the caller's stream; end() returns the summed device time (ms) and launch count
 * per kernel class.  Arrays must hold RP_NUM_TAGS entries. */
enum {
  RP_TAG_CAST = 0, RP_TAG_GEMM_IN = 1, RP_TAG_LAYERNORM = 2, RP_TAG_GEMM_QKV = 3, RP_TAG_FMHA = 4,
  RP_TAG_GEMM_OUT = 5, RP_TAG_GEMM_FF1 = 6, RP_TAG_GEMM_FF2 = 7, RP_TAG_GEMM_FMAP = 8,
  RP_TAG_GEMM_HEAD = 9, RP_TAG_HEAD_OUT = 10, RP_NUM_TAGS = 11
};
/* rp_forward* skip work that only touches padding (default on; RP_SKIP_PADDING=0 in the environment turns it off at
 * rp_create): 256-row blocks of the padded [B, T] token matrix without a step t < round_up(lens[b], 128) are left out of
 * every GEMM, attention query tiles beyond that limit exit at once, and the padded steps (t >= lens[b]) of logits /
 * offsets / feats are returned as zeros.  The reference computes padded steps too and returns finite values nobody
 * reads (models/MMCTransformer.py:132-138 masks them as keys; callers mask them as outputs); valid steps are
 * bit-identical either way. */
int32_t rp_set_skip_padding(rp_handle* h, int32_t on);
int32_t rp_profile_begin(rp_handle* h);
int32_t rp_profile_end(rp_handle* h, float* ms_by_tag, int32_t* launches_by_tag);

/* ---- decode + Soft-NMS ------------------------------------------------------------------------ */
/* logits [B,T], offsets [B,T,2], lens [B], max_seg [B] ->
 * segs [B,Kcap,2] f32, scores [B,Kcap] f32 (probability of the kept candidate), dscores [B,Kcap]
 * f32 (its decayed Soft-NMS score), labels [B,Kcap] i32 (centre step), counts [B] i32,
 * ncand [B] i32 (candidates that entered Soft-NMS).  Optional (all NULL or all set) pre-NMS
 * candidate list = the return value of inference_single_video: cand_segs [B,C,2], cand_scores [B,C],
 * cand_labels [B,C], C = min(pre_nms_topk, T).  Requires T <= 8192, topk <= 4096, Kcap <= 64. */
int32_t rp_decode_nms(const float* logits, const float* offsets, const int32_t* lens,
                      const int32_t* max_seg, int32_t B, int32_t T, const rp_decode_cfg* cfg,
                      int32_t Kcap, float* segs, float* scores, float* dscores, int32_t* labels,
                      int32_t* counts, int32_t* ncand, float* cand_segs, float* cand_scores,
                      int32_t* cand_labels, void* stream);
/* scores [B,Nmax] (descending, as produced by decode), segs [B,Nmax,2], n [B], max_seg [B] ->
 * keep [B,Kcap] i32 original indices in selection order, kscores [B,Kcap] decayed scores,
 * counts [B].  Nmax <= 8192. */
int32_t rp_soft_nms(const float* scores, const float* segs, const int32_t* n, const int32_t* max_seg,
                    int32_t B, int32_t Nmax, float sigma, float thresh, int32_t Kcap, int32_t* keep,
                    float* kscores, int32_t* counts, void* stream);

/* AtIoU on the device: calculate_tiou (utils/metrics.py:82-111) per video and the
 * averaging of inference.py:45-55, in float64 with the reference's operation order.
 * slots [n,1+4K] f32 = [count,(start,end,score,label)*K] per video (the all-gathered layout),
 * gt [n,Gmax,2] f64, gt_counts [n] i32, thresholds [n_thr<=8] f64 -> per_video [n,n_thr] f64,
 * out [n_thr+1] f64: mean precision per threshold, then their mean (the reported "average tIoU"). */
int32_t rp_atiou(const float* slots, int32_t n_videos, int32_t K, const double* gt,
                 const int32_t* gt_counts, int32_t Gmax, const double* thresholds, int32_t n_thr,
                 double* per_video, double* out, void* stream);

/* Forward value of MMCTransformer.losses (models/MMCTransformer.py:159-179): sum over valid steps of
 * sigmoid_focal_loss(logits, labels; alpha, gamma) (models/losses.py:5-53; the reference uses alpha = 0.7,
 * gamma = 2).  logits / targets [n] f32, mask [n] u8 (non-zero = valid step), scratch >= 296 doubles,
 * out [1] f32.  Deterministic.  No gradient: the training step is outside this path. */
int32_t rp_focal_loss_sum(const float* logits, const float* targets, const uint8_t* mask, int64_t n, float alpha,
                          float gamma, double* scratch, float* out, void* stream);

/* ---- first pieces of the training step (reference main.py:294-409) ------------------------
 * rp_focal_loss_grad      <- autograd of sigmoid_focal_loss(...).sum() / batch_size   models/losses.py:5-53,
 *                            models/MMCTransformer.py:159-179, main.py:326-333: dlogits[i] = mask[i] ? scale * dL_i/dx_i : 0
 * rp_layernorm512_bwd     <- autograd of nn.LayerNorm(512) (models/MMCTransformer.py:46-58 norm1/norm2/encoder_norm...):
 *                            dx [M,512], dgamma [512], dbeta [512]; scratch: rp_layernorm512_bwd_scratch_bytes()
 * rp_adam_step            <- optim.Adam(model.parameters(), lr, weight_decay).step()   main.py:190-191, :368;
 *                            flat fp32 buffers, step >= 1, optional bf16 copy of the updated parameters */
int32_t rp_focal_loss_grad(const float* logits, const float* targets, const uint8_t* mask, int64_t n, float alpha,
                           float gamma, float scale, float* dlogits, void* stream);
int64_t rp_layernorm512_bwd_scratch_bytes(void);
int32_t rp_layernorm512_bwd(const float* x, const float* dy, const float* gamma, int64_t M, float eps, float* dx,
                            float* dgamma, float* dbeta, void* scratch, int64_t scratch_bytes, void* stream);
int32_t rp_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                     float beta1, float beta2, float eps, float weight_decay, int32_t step, void* param_bf16,
                     void* stream);

/* ---- backward pass of the training step (reference main.py:326-333 `final_loss.backward()`,
 *      i.e. torch.autograd over models/MMCTransformer.py:109-151).  bf16 tensor-core GEMMs with fp32 accumulation,
 *      fp32 gradients, every reduction in a fixed order (no atomics).
 * rp_train_scratch_bytes   size of the scratch buffer the two-stage reductions below need
 * rp_layernorm512_bwd_acc  autograd of nn.LayerNorm(512) on a pre-LN branch: dh_inout (+)= dx (accumulate != 0: +=);
 *                          dh_bf16 (optional) = bf16 copy of the resulting dh (the operand of the next dgrad / wgrad
 *                          GEMMs); dh_colsum (optional) = its column sums [512] = the bias gradient of the Linear
 *                          whose output gradient dh is
 * rp_gemm_bwd              autograd of nn.Linear: kind 1 (dgrad) D[M,N] = A[M,K] B[K,N] with B the weight [out=K, in=N]
 *                          as stored; kind 3 (wgrad) D[M,N] = A[K,M]^T B[K,N] with A = dY [tokens, out], B = X [tokens, in]
 *                          (K = tokens, any count); out_f32 != 0: fp32 D, else bf16; splits > 1: split-K partials
 *                          stacked in D [splits * M, N] (M % 256 == 0), to be summed by rp_splitk_reduce
 * rp_colsum_bf16           bias gradient: out[N] = column sums of bf16 x[M, N]
 * rp_relu_bwd              autograd of nn.ReLU: dy = act > 0 ? dy : 0 in place (bf16/bf16 or fp32/fp32)
 * rp_head_out_bwd          autograd of cls_head.7 (Linear 256 -> 1) and the ReLU in front of it
 *                          (models/MMCTransformer.py:71-80): da2 [M,256] bf16, dw [256], db [1]
 * rp_fmha_train            rp_fmha (key-padding mode) that also writes the log2-domain log-sum-exp [B,H,T]
 * rp_fmha_bwd              autograd of the attention inside nn.MultiheadAttention (models/MMCTransformer.py:41-55,
 *                          132-138): dq (w.r.t. the unscaled q), dk, dv; dsum = scratch [B,H,T] fp32 */
/* out = in with the first n_scaled elements multiplied by scale, as bf16 and / or fp32: the attention kernels take
 * q pre-scaled by log2(e)/sqrt(64), so the training forward's in_proj operand is in_proj_weight / in_proj_bias with the
 * q rows scaled (rp_load_weight does the same fold for inference) */
int32_t rp_cast_scaled(const float* in, int64_t n, int64_t n_scaled, float scale, void* out_bf16, float* out_f32,
                       void* stream);
int64_t rp_train_scratch_bytes(void);
int32_t rp_layernorm512_bwd_acc(const float* x, const float* dy, const float* gamma, int64_t M, float eps,
                                int32_t accumulate, float* dh_inout, void* dh_bf16, float* dh_colsum, float* dgamma,
                                float* dbeta, void* scratch, int64_t scratch_bytes, void* stream);
int32_t rp_gemm_bwd(int32_t kind, int32_t out_f32, const void* A, int64_t lda, const void* B, int64_t ldb, void* D,
                    int64_t ldd, int32_t M, int32_t N, int32_t K, int32_t splits, void* stream);
int32_t rp_splitk_reduce(const float* partials, int32_t splits, int64_t n, float* out, void* stream);
int32_t rp_colsum_bf16(const void* x, int64_t M, int32_t N, float* out, void* scratch, int64_t scratch_bytes,
                       void* stream);
int32_t rp_relu_bwd(void* dy, const void* act, int64_t n, int32_t is_f32, void* stream);
/* rp_relu_bwd followed by rp_colsum_bf16 in one pass: the ReLU mask and the bias gradient of the Linear in front of it */
int32_t rp_relu_bwd_colsum(void* dy_bf16, const void* act_bf16, int64_t M, int32_t N, float* colsum, void* scratch,
                           int64_t scratch_bytes, void* stream);
int32_t rp_head_out_bwd(const float* dlogits, const void* a2_bf16, const float* w, int64_t M, void* da2_bf16,
                        float* dw, float* db, void* scratch, int64_t scratch_bytes, void* stream);
int32_t rp_fmha_train(const void* q, const void* k, const void* v, void* o, int64_t ld_qkv, int64_t ld_o, int32_t B,
                      int32_t H, int32_t T, const int32_t* kv_lens, float* lse, void* stream);
int32_t rp_fmha_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                    float* dsum, void* dq, void* dk, void* dv, int64_t ld_qkv, int64_t ld_o, int64_t ld_dqkv,
                    int32_t B, int32_t H, int32_t T, const int32_t* kv_lens, void* stream);
/* rp_fmha_bwd / rp_fmha_bwd_dropout run ONE kernel for dK, dV and dQ by default; its dQ is summed over the key tiles by
 * fp32 adds in L2 whose order is not fixed, so dq is reproducible to the rounding of a T/128-term fp32 sum, not bit for bit
 * (torch's own attention backward makes the same trade unless torch.use_deterministic_algorithms is set).  on = 1 selects
 * the two deterministic kernels (about 20 % slower), on = 0 the fused one, on < 0 the default (environment
 * RP_FMHA_BWD_FUSED, fused unless it is 0).  Process-wide.  The fused path keeps one fp32 dQ accumulator per device
 * (B * H * ceil(T/64) * 16 KB, allocated on first use and grown on demand): calls on the SAME device must be ordered on
 * one stream (as a training step's are); the deterministic kernels have no such state. */
int32_t rp_set_attn_bwd_deterministic(int32_t on);

/* ---- train-mode dropout (the nn.Dropout(0.1) sites of the reference graph under model.train(), main.py:285:
 * nn.TransformerEncoderLayer's dropout1 / dropout / dropout2 and the attention-weight dropout of its
 * nn.MultiheadAttention, models/MMCTransformer.py:41-49; feature_map[3], cls_head[3]/[6], reg_head[3]/[6], :63-93).
 * A site's mask is a pure function of (key_a, key_b, element index): pair k = element >> 1,
 * h = fmix32(k * 0x9E3779B1 + key_a) ^ key_b, element 2k kept iff (h & 0xffff) >= round(p * 65536), element 2k+1 iff
 * (h >> 16) >= the same threshold; kept values are scaled by 1 / (1 - p).  The forward epilogues and the backward
 * kernels evaluate the same function, so nothing is stored for the element-wise sites; the attention-weight mask is
 * materialised as one keep bit per (query, key).  torch's Philox stream is not reproduced (same distribution,
 * different draws).
 *   rp_dropout_mask_u8            keep[i] in {0,1} for the first n elements of a site (tests feed it to autograd)
 *   rp_attn_dropout_bits          keep bits of an attention site: word w = elements 32 w .. 32 w + 31 (bit = element & 31);
 *                                 rows of bits_ld words (bits_ld % 4 == 0, 32 * bits_ld >= round_up(T, 128)) per (b, h, query)
 *   rp_gemm_bf16_dropout          rp_gemm_bf16 epilogue 1 (Dropout after the ReLU) or 3 (Dropout before the residual add);
 *                                 element index = row * N + column
 *   rp_layernorm512_dropout       rp_layernorm512 mode 2 with feature_map's Dropout on f (the head LayerNorms see the
 *                                 dropped row); element index = row * 512 + column
 *   rp_layernorm512_bwd_acc_dropout  rp_layernorm512_bwd_acc where dh is the output gradient of a residual branch ending in
 *                                 Dropout: dh_bf16 / dh_colsum hold kept ? dh / (1 - p) : 0, dh_inout stays the stream's gradient
 *   rp_relu_bwd_scaled, rp_relu_bwd_colsum_scaled, rp_head_out_bwd_scaled
 *                                 ReLU backward through a stored dropout(relu(.)) activation: act > 0 ? dy * scale : 0
 *   rp_fmha_train_dropout / rp_fmha_bwd_dropout
 *                                 out = (keep o softmax(S)) V / (1 - p) and its autograd */
typedef struct rp_dropout {
  uint32_t key_a, key_b;
  float p; /* 0 = off */
} rp_dropout;
int32_t rp_dropout_mask_u8(const rp_dropout* drop, int64_t n, uint8_t* keep, void* stream);
int32_t rp_attn_dropout_bits(const rp_dropout* drop, int64_t n_words, uint32_t* keep_bits, void* stream);
int32_t rp_gemm_bf16_dropout(int32_t epilogue, const void* A, int64_t lda, const void* W, int64_t ldw, void* D,
                             int64_t ldd, const float* bias, const float* resid, int64_t ldr, int32_t M, int32_t N,
                             int32_t K, const rp_dropout* drop, void* stream);
int32_t rp_layernorm512_dropout(int32_t mode, const float* x, int64_t M, int32_t T, const float* g0,
                                const float* b0, const float* g1, const float* b1, const float* g2,
                                const float* b2, const float* pe, float* out_f32, void* y_bf16, void* y2_bf16,
                                const rp_dropout* drop, void* stream);
int32_t rp_layernorm512_bwd_acc_dropout(const float* x, const float* dy, const float* gamma, int64_t M, float eps,
                                        int32_t accumulate, float* dh_inout, void* dh_bf16, float* dh_colsum,
                                        float* dgamma, float* dbeta, void* scratch, int64_t scratch_bytes,
                                        const rp_dropout* drop, void* stream);
int32_t rp_relu_bwd_scaled(void* dy, const void* act, int64_t n, int32_t is_f32, float scale, void* stream);
int32_t rp_relu_bwd_colsum_scaled(void* dy_bf16, const void* act_bf16, int64_t M, int32_t N, float scale, float* colsum,
                                  void* scratch, int64_t scratch_bytes, void* stream);
int32_t rp_head_out_bwd_scaled(const float* dlogits, const void* a2_bf16, const float* w, int64_t M, float scale,
                               void* da2_bf16, float* dw, float* db, void* scratch, int64_t scratch_bytes, void* stream);
int32_t rp_fmha_train_dropout(const void* q, const void* k, const void* v, void* o, int64_t ld_qkv, int64_t ld_o,
                              int32_t B, int32_t H, int32_t T, const int32_t* kv_lens, float* lse,
                              const uint32_t* keep_bits, int64_t bits_ld, float p, void* stream);
int32_t rp_fmha_bwd_dropout(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                            float* dsum, void* dq, void* dk, void* dv, int64_t ld_qkv, int64_t ld_o, int64_t ld_dqkv,
                            int32_t B, int32_t H, int32_t T, const int32_t* kv_lens, const uint32_t* keep_bits,
                            int64_t bits_ld, float p, void* stream);

/* ---- building blocks (also what the unit tests drive) ---------------------------------------- */
/* D[M,N] = A[M,K] * W[N,K]^T + bias (+ReLU | +residual); A, W bf16 row-major with pitches lda/ldw
 * (elements); epilogue: 0 bf16 out, 1 bf16 out + ReLU, 2 f32 out, 3 f32 out + residual (may alias D).
 * N % 256 == 0, K % 64 == 0. */
int32_t rp_gemm_bf16(int32_t epilogue, const void* A, int64_t lda, const void* W, int64_t ldw, void* D,
                     int64_t ldd, const float* bias, const float* resid, int64_t ldr, int32_t M,
                     int32_t N, int32_t K, void* stream);
/* h[M,512] (f32, in place) += A[M,K] W[512,K]^T + bias, and u[M,512] (bf16) = LayerNorm(h; gamma, beta,
 * eps) of the updated rows, in one kernel: the residual update of nn.TransformerEncoderLayer followed by
 * the pre-LN of the next sub-block (models/MMCTransformer.py:135-138: x = x + sa(norm1(x)); x = x +
 * ff(norm2(x))).  M > 128. */
int32_t rp_gemm_resid_ln(const void* A, int64_t lda, const void* W, int64_t ldw, float* h, int64_t ldh,
                         const float* bias, const float* gamma, const float* beta, float eps, void* u_bf16,
                         int64_t ldu, int32_t M, int32_t K, void* stream);
/* The tail of cls_head / reg_head (models/MMCTransformer.py:71-93: Linear(256,256), ReLU, [Dropout], Linear(256, nj), and
 * reg_head's final ReLU) in one kernel: out[m, j] = act(sum_c relu(A[m,:] . W[c,:] + bias[c]) * w_last[j, c] + b_last[j]);
 * A bf16 [M,K] (pitch lda), W bf16 [256,K], bias [256], w_last fp32 [nj,256], nj = 1 or 2, out fp32 [M,nj]. */
int32_t rp_gemm_head_dot(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const float* w_last,
                         const float* b_last, int32_t nj, int32_t final_relu, float* out, int32_t M, int32_t K,
                         void* stream);
/* softmax(q k^T + mask) v per head of 64; q must be pre-scaled by log2(e)/8.  bf16 in/out; ld* row
 * pitch and bs* batch pitch in elements.  mask_mode 0: keys >= kv_lens[b] are -inf (kv_lens may be
 * NULL); mask_mode 1: uint8 mask, 0 => masked_fill(-1e9), mask_q_stride 0 broadcasts over queries. */
int32_t rp_fmha(const void* q, const void* k, const void* v, void* o, int64_t ldq, int64_t ldk,
                int64_t ldv, int64_t ldo, int64_t bsq, int64_t bsk, int64_t bsv, int64_t bso,
                int32_t B, int32_t H, int32_t Tq, int32_t Tk, const int32_t* kv_lens, int32_t mask_mode,
                const uint8_t* mask, int64_t mask_b_stride, int64_t mask_q_stride, void* stream);
int32_t rp_concat_cast(const float* vis, const float* aud, const float* txt, int32_t Cv, int32_t Ca,
                       int32_t Ct, void* out_bf16, int64_t M, void* stream);
/* masks [B, T] (one byte per step, non-zero = valid; the bool mask of dataset/RepurposeClip.py:528-531 viewed
 * as bytes) -> lens [B] int32 = valid steps per video, *not_aligned (device int32) = 1 if some mask is not
 * left-aligned (mask[b][t] != (t < lens[b])).  The reference hands the mask itself to nn.MultiheadAttention
 * (models/MMCTransformer.py:132-138); rp_forward / rp_decode_nms take lengths, so the host mirror refuses
 * anything this flag marks. */
int32_t rp_mask_lens(const uint8_t* mask, int32_t B, int32_t T, int32_t* lens, int32_t* not_aligned, void* stream);
int32_t rp_cast_bf16(const float* in, void* out_bf16, int64_t n, void* stream);
/* LayerNorm over rows of 512, eps 1e-5.  mode 0: y=bf16(LN(x)); 1: h=LN(x)+pe[row%T] (f32 out),
 * y=bf16(LN1(h)); 2: f=relu(LN(x)) (f32 out), y=bf16(LN1(f)), y2=bf16(LN2(f)); 3: f32 out = LN(x). */
int32_t rp_layernorm512(int32_t mode, const float* x, int64_t M, int32_t T, const float* g0,
                        const float* b0, const float* g1, const float* b1, const float* g2,
                        const float* b2, const float* pe, float* out_f32, void* y_bf16, void* y2_bf16,
                        void* stream);
int32_t rp_head_out(const void* a_cls_bf16, const void* a_reg_bf16, const float* w_cls,
                    const float* b_cls, const float* w_reg, const float* b_reg, float* logits,
                    float* offsets, int64_t M, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* REPURPOSE_B200_H_ */
