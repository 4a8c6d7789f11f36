// C-ABI (include/repurpose_b200.h): model handle, reference-state-dict weight repacking, and the
// forward-pass launch sequence that restates MMCTransformer.forward
// (reference models/MMCTransformer.py:109-151; arithmetic in SURVEY.md Appendix A).
#include <cuda_bf16.h>

#include <stdlib.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/repurpose_b200.h"
#include "host_util.h"
#include "kernels.h"
#include "ptx.cuh"

using namespace rp;

namespace {

constexpr float kQScale = 1.4426950408889634f / 8.0f;  // log2(e) / sqrt(64), folded into Wq, bq

enum WeightKind { W_MAT_BF16, W_VEC_F32, W_QKV_MAT, W_QKV_VEC, W_PE };

struct WeightSlot {
  WeightKind kind;
  int64_t numel;       // expected element count (W_PE: elements per position row * max_len)
  void* dev = nullptr; // bf16 for matrices, f32 otherwise
  bool loaded = false;
};

struct LayerW {
  __nv_bfloat16 *w_qkv, *w_out, *w_ff1, *w_ff2;
  float *b_qkv, *b_out, *b_ff1, *b_ff2;
  float *n1_g, *n1_b, *n2_g, *n2_b;
};

}  // namespace

struct rp_handle {
  rp_model_cfg cfg;
  std::map<std::string, WeightSlot> slots;
  std::vector<LayerW> layers;
  // non-layer weights
  __nv_bfloat16 *w_in, *w_fm, *w_c1, *w_c4, *w_r1, *w_r4;
  float *b_in, *in_g, *in_b, *enc_g, *enc_b, *b_fm, *fm_g, *fm_b;
  float *c0_g, *c0_b, *b_c1, *b_c4, *w_c7, *b_c7;
  float *r0_g, *r0_b, *b_r1, *b_r4, *w_r7, *b_r7;
  float* pe;
  int64_t pe_rows_loaded = 0;
  // skip 256-row blocks / query tiles that hold nothing but padding (rp_set_skip_padding; default RP_SKIP_PADDING or 1)
  bool skip_padding = !(getenv("RP_SKIP_PADDING") && atoi(getenv("RP_SKIP_PADDING")) == 0);
  // optional per-kernel-class CUDA-event profiler (rp_profile_begin / rp_profile_end)
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_pool;
  size_t prof_used = 0;
  struct ProfRec { int tag; cudaEvent_t e0, e1; };
  std::vector<ProfRec> prof_recs;
  cudaEvent_t prof_event() {
    if (prof_used == prof_pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      prof_pool.push_back(e);
    }
    return prof_pool[prof_used++];
  }
};

namespace {

__global__ void repack_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                   int64_t n, int64_t n_scaled, float scale) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n;
       i += int64_t(gridDim.x) * blockDim.x) {
    const float v = src[i];
    dst[i] = __float2bfloat16_rn(i < n_scaled ? v * scale : v);
  }
}
__global__ void repack_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n,
                                  int64_t n_scaled, float scale) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n;
       i += int64_t(gridDim.x) * blockDim.x) {
    const float v = src[i];
    dst[i] = i < n_scaled ? v * scale : v;
  }
}

template <typename T>
int add_slot(rp_handle* h, const std::string& name, WeightKind kind, int64_t numel, T** field) {
  WeightSlot s;
  s.kind = kind;
  s.numel = numel;
  const size_t bytes = size_t(numel) * ((kind == W_MAT_BF16 || kind == W_QKV_MAT) ? 2 : 4);
  RP_CUDA_CHECK(cudaMalloc(&s.dev, bytes));
  *field = reinterpret_cast<T*>(s.dev);
  h->slots[name] = s;
  return RP_OK;
}

#define ADD(name, kind, numel, field)                                         \
  do {                                                                        \
    int _rc = add_slot(h, (name), (kind), (numel), &(field));                 \
    if (_rc) return _rc;                                                      \
  } while (0)

int build_slots(rp_handle* h) {
  const rp_model_cfg& c = h->cfg;
  const int64_t D = c.d_model, F = c.d_ff, Hh = c.head_hidden;
  const int64_t Cin = int64_t(c.vis_dim) + c.aud_dim + c.text_dim;
  ADD("input_projection.weight", W_MAT_BF16, D * Cin, h->w_in);
  ADD("input_projection.bias", W_VEC_F32, D, h->b_in);
  ADD("input_norm.weight", W_VEC_F32, D, h->in_g);
  ADD("input_norm.bias", W_VEC_F32, D, h->in_b);
  ADD("positional_encoding.pe", W_PE, int64_t(c.max_len) * D, h->pe);
  h->layers.resize(c.num_layers);
  for (int l = 0; l < c.num_layers; ++l) {
    const std::string p = "multimodal_encoder.layers." + std::to_string(l) + ".";
    LayerW& L = h->layers[l];
    ADD(p + "self_attn.in_proj_weight", W_QKV_MAT, 3 * D * D, L.w_qkv);
    ADD(p + "self_attn.in_proj_bias", W_QKV_VEC, 3 * D, L.b_qkv);
    ADD(p + "self_attn.out_proj.weight", W_MAT_BF16, D * D, L.w_out);
    ADD(p + "self_attn.out_proj.bias", W_VEC_F32, D, L.b_out);
    ADD(p + "linear1.weight", W_MAT_BF16, F * D, L.w_ff1);
    ADD(p + "linear1.bias", W_VEC_F32, F, L.b_ff1);
    ADD(p + "linear2.weight", W_MAT_BF16, D * F, L.w_ff2);
    ADD(p + "linear2.bias", W_VEC_F32, D, L.b_ff2);
    ADD(p + "norm1.weight", W_VEC_F32, D, L.n1_g);
    ADD(p + "norm1.bias", W_VEC_F32, D, L.n1_b);
    ADD(p + "norm2.weight", W_VEC_F32, D, L.n2_g);
    ADD(p + "norm2.bias", W_VEC_F32, D, L.n2_b);
  }
  ADD("encoder_norm.weight", W_VEC_F32, D, h->enc_g);
  ADD("encoder_norm.bias", W_VEC_F32, D, h->enc_b);
  ADD("feature_map.0.weight", W_MAT_BF16, D * D, h->w_fm);
  ADD("feature_map.0.bias", W_VEC_F32, D, h->b_fm);
  ADD("feature_map.1.weight", W_VEC_F32, D, h->fm_g);
  ADD("feature_map.1.bias", W_VEC_F32, D, h->fm_b);
  ADD("cls_head.0.weight", W_VEC_F32, D, h->c0_g);
  ADD("cls_head.0.bias", W_VEC_F32, D, h->c0_b);
  ADD("cls_head.1.weight", W_MAT_BF16, Hh * D, h->w_c1);
  ADD("cls_head.1.bias", W_VEC_F32, Hh, h->b_c1);
  ADD("cls_head.4.weight", W_MAT_BF16, Hh * Hh, h->w_c4);
  ADD("cls_head.4.bias", W_VEC_F32, Hh, h->b_c4);
  ADD("cls_head.7.weight", W_VEC_F32, Hh, h->w_c7);
  ADD("cls_head.7.bias", W_VEC_F32, 1, h->b_c7);
  ADD("reg_head.0.weight", W_VEC_F32, D, h->r0_g);
  ADD("reg_head.0.bias", W_VEC_F32, D, h->r0_b);
  ADD("reg_head.1.weight", W_MAT_BF16, Hh * D, h->w_r1);
  ADD("reg_head.1.bias", W_VEC_F32, Hh, h->b_r1);
  ADD("reg_head.4.weight", W_MAT_BF16, Hh * Hh, h->w_r4);
  ADD("reg_head.4.bias", W_VEC_F32, Hh, h->b_r4);
  ADD("reg_head.7.weight", W_VEC_F32, 2 * Hh, h->w_r7);
  ADD("reg_head.7.bias", W_VEC_F32, 2, h->b_r7);
  return RP_OK;
}

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

struct Workspace {
  float* h;             // [M,512] fp32 residual stream
  __nv_bfloat16* u;     // [M,512]
  __nv_bfloat16* attn;  // [M,512]
  __nv_bfloat16* qkv;   // [M,1536]   } contiguous: also holds xcat [M,Cin] during the input stage
  __nv_bfloat16* ffn;   // [M,d_ff]   }
  int32_t* row_blocks;  // RowMap: [ceil(M/256)] valid 256-row blocks + their count (padding-only blocks are skipped)
  int32_t* row_count;
  int32_t* batch_order;  // [B] batch elements by decreasing length (the attention kernel's CTA layout)
  int64_t bytes;
};

Workspace carve(const rp_model_cfg& c, int64_t M, int64_t B, void* base) {
  Workspace w;
  const int64_t D = c.d_model;
  const int64_t Cin = int64_t(c.vis_dim) + c.aud_dim + c.text_dim;
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    const int64_t o = off;
    off += align_up(bytes, 1024);
    return o;
  };
  const int64_t o_h = take(M * D * 4);
  const int64_t o_u = take(M * D * 2);
  const int64_t o_a = take(M * D * 2);
  int64_t wide = M * (3 * D + c.d_ff) * 2;
  if (wide < M * Cin * 2) wide = M * Cin * 2;
  const int64_t o_q = take(wide);
  const int64_t o_m = take(((M + 255) / 256 + 1) * 4);
  const int64_t o_b = take(B * 4);
  uint8_t* b = reinterpret_cast<uint8_t*>(base);
  w.h = reinterpret_cast<float*>(b + o_h);
  w.u = reinterpret_cast<__nv_bfloat16*>(b + o_u);
  w.attn = reinterpret_cast<__nv_bfloat16*>(b + o_a);
  w.qkv = reinterpret_cast<__nv_bfloat16*>(b + o_q);
  w.ffn = w.qkv + M * 3 * D;
  w.row_count = reinterpret_cast<int32_t*>(b + o_m);
  w.row_blocks = w.row_count + 1;
  w.batch_order = reinterpret_cast<int32_t*>(b + o_b);
  w.bytes = off;
  return w;
}

}  // namespace

extern "C" {

int32_t rp_abi_version(void) { return RP_ABI_VERSION; }
const char* rp_last_error(void) { return rp::last_error(); }
int64_t rp_launch_count(void) { return rp::launch_count(); }

int32_t rp_create(const rp_model_cfg* cfg, rp_handle** out) {
  RP_CHECK(cfg != nullptr && out != nullptr, "rp_create: null argument");
  RP_CHECK(cfg->d_model == 512, "rp_create: d_model=%d unsupported (kernels are specialised for 512)",
           cfg->d_model);
  RP_CHECK(cfg->num_heads > 0 && cfg->d_model / cfg->num_heads == 64 &&
               cfg->d_model % cfg->num_heads == 0,
           "rp_create: head dim must be 64 (d_model=%d, heads=%d)", cfg->d_model, cfg->num_heads);
  RP_CHECK(cfg->head_hidden == 256, "rp_create: head_hidden must be 256");
  RP_CHECK(cfg->d_ff > 0 && cfg->d_ff % 256 == 0, "rp_create: d_ff must be a multiple of 256");
  RP_CHECK(cfg->vis_dim % 8 == 0 && cfg->aud_dim % 8 == 0 && cfg->text_dim % 8 == 0 &&
               (cfg->vis_dim + cfg->aud_dim + cfg->text_dim) % 64 == 0,
           "rp_create: modality dims must be multiples of 8 and sum to a multiple of 64");
  RP_CHECK(cfg->num_layers >= 1 && cfg->max_len >= 1, "rp_create: bad num_layers / max_len");
  if (num_sms() <= 0) return RP_ERR_NO_DEVICE;
  rp_handle* h = new rp_handle();
  h->cfg = *cfg;
  int rc = build_slots(h);
  if (rc) {
    rp_destroy(h);
    return rc;
  }
  *out = h;
  return RP_OK;
}

void rp_destroy(rp_handle* h) {
  if (h == nullptr) return;
  for (auto& kv : h->slots) {
    if (kv.second.dev) cudaFree(kv.second.dev);
  }
  for (cudaEvent_t e : h->prof_pool) cudaEventDestroy(e);
  delete h;
}

int32_t rp_load_weight(rp_handle* h, const char* name, const float* src, int64_t numel, void* stream) {
  RP_CHECK(h != nullptr && name != nullptr && src != nullptr, "rp_load_weight: null argument");
  auto it = h->slots.find(name);
  RP_CHECK(it != h->slots.end(), "rp_load_weight: unexpected key '%s'", name);
  WeightSlot& s = it->second;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t D = h->cfg.d_model;
  if (s.kind == W_PE) {
    RP_CHECK(numel % D == 0 && numel >= D && numel <= s.numel,
             "rp_load_weight: positional_encoding.pe has %lld elements, handle holds %lld",
             (long long)numel, (long long)s.numel);
    h->pe_rows_loaded = numel / D;
  } else {
    RP_CHECK(numel == s.numel, "rp_load_weight: '%s' has %lld elements, expected %lld", name,
             (long long)numel, (long long)s.numel);
  }
  const int threads = 256;
  int64_t blocks64 = (numel + threads - 1) / threads;
  const int blocks = int(blocks64 > 4096 ? 4096 : blocks64);
  switch (s.kind) {
    case W_MAT_BF16:
      repack_bf16_kernel<<<blocks, threads, 0, st>>>(src, reinterpret_cast<__nv_bfloat16*>(s.dev),
                                                     numel, 0, 1.0f);
      break;
    case W_QKV_MAT:  // rows [0, D) are the query projection
      repack_bf16_kernel<<<blocks, threads, 0, st>>>(src, reinterpret_cast<__nv_bfloat16*>(s.dev),
                                                     numel, D * D, kQScale);
      break;
    case W_QKV_VEC:
      repack_f32_kernel<<<blocks, threads, 0, st>>>(src, reinterpret_cast<float*>(s.dev), numel, D,
                                                    kQScale);
      break;
    case W_VEC_F32:
    case W_PE:
      repack_f32_kernel<<<blocks, threads, 0, st>>>(src, reinterpret_cast<float*>(s.dev), numel, 0,
                                                    1.0f);
      break;
  }
  count_launch();
  RP_CUDA_CHECK(cudaGetLastError());
  s.loaded = true;
  return RP_OK;
}

int32_t rp_weights_complete(const rp_handle* h) {
  RP_CHECK(h != nullptr, "rp_weights_complete: null handle");
  for (const auto& kv : h->slots)
    RP_CHECK(kv.second.loaded, "missing weight '%s'", kv.first.c_str());
  return RP_OK;
}

int64_t rp_workspace_bytes(const rp_handle* h, int32_t B, int32_t T) {
  if (h == nullptr || B <= 0 || T <= 0) return -1;
  return carve(h->cfg, int64_t(B) * T, B, nullptr).bytes;
}

// Ragged input description (rp_forward_ragged); all null for the padded entry point.
struct RaggedIn {
  const int32_t* row_off = nullptr;   // [B] first row of video b in vis / aud
  const int32_t* txt_off = nullptr;   // [B] first row of video b in txt
  const int32_t* txt_lens = nullptr;  // [B] rows of text features available for video b
  bool bf16 = false;                  // feature rows are bf16 (pre-converted feature files) instead of fp32
};

static int32_t forward_core(rp_handle* h, const void* vis_v, const void* aud_v, const void* txt_v,
                            const RaggedIn& rg, const int32_t* lens, int32_t B, int32_t T,
                            float* out_logits, float* out_offsets, float* out_feats, void* workspace,
                            int64_t workspace_bytes, void* stream) {
  const float* vis = static_cast<const float*>(vis_v);
  const float* aud = static_cast<const float*>(aud_v);
  const float* txt = static_cast<const float*>(txt_v);
  RP_CHECK(h && vis && aud && txt && lens && out_logits && out_offsets && out_feats && workspace,
           "rp_forward: null argument");
  RP_CHECK(B > 0 && T > 0, "rp_forward: empty batch");
  int rc = rp_weights_complete(h);
  if (rc) return rc;
  RP_CHECK(T <= h->pe_rows_loaded,
           "rp_forward: T=%d exceeds the %lld positional-encoding rows loaded", T,
           (long long)h->pe_rows_loaded);
  const rp_model_cfg& c = h->cfg;
  const int64_t M64 = int64_t(B) * T;
  RP_CHECK(M64 < (int64_t(1) << 31), "rp_forward: B*T too large");
  const int M = int(M64);
  Workspace w = carve(c, M, B, workspace);
  if (w.bytes > workspace_bytes) {
    set_last_error("rp_forward: workspace too small (%lld < %lld)", (long long)workspace_bytes,
                   (long long)w.bytes);
    return RP_ERR_WORKSPACE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int D = c.d_model, F = c.d_ff, Hh = c.head_hidden, H = c.num_heads;
  const int Cin = c.vis_dim + c.aud_dim + c.text_dim;
  const float eps = 1e-5f;
#define RUN(tag, expr)                                         \
  do {                                                         \
    cudaEvent_t _e0 = nullptr, _e1 = nullptr;                  \
    if (h->prof_on) {                                          \
      _e0 = h->prof_event();                                   \
      _e1 = h->prof_event();                                   \
      cudaEventRecord(_e0, st);                                \
    }                                                          \
    rc = (expr);                                               \
    if (rc) return rc;                                         \
    if (h->prof_on) {                                          \
      cudaEventRecord(_e1, st);                                \
      h->prof_recs.push_back({(tag), _e0, _e1});               \
    }                                                          \
  } while (0)

  // Padding-only blocks are skipped (rp_set_skip_padding / RP_SKIP_PADDING, default on): the GEMMs walk the RowMap's 256-row blocks, attention
  // CTAs of padding-only query tiles exit at once, and the padded steps of the three outputs are zeroed at the end.
  // The reference computes padded rows too (finite values nobody reads); valid rows are bit-identical
  // with and without skipping because no tile's arithmetic depends on which other tiles run.
  const bool skip = h->skip_padding && M > 128;
  RowMap rmap{w.row_blocks, w.row_count};
  const RowMap* rows = skip ? &rmap : nullptr;
  // (the same tiny kernel ranks the videos by length for the attention kernel's CTA layout: useful with or without the skip)
  const bool ranked = skip || B > 1;
  if (ranked) {
    rc = launch_row_map(lens, B, T, w.row_blocks, w.row_count, w.batch_order, st);
    if (rc) return rc;
  }
  // (1) concat + cast, input projection (fp32 out), input_norm + PE -> h, layers[0].norm1 -> u
  __nv_bfloat16* xcat = w.qkv;
  if (rg.row_off != nullptr)
    RUN(RP_TAG_CAST, launch_ragged_concat_cast(vis_v, aud_v, txt_v, rg.bf16, c.vis_dim, c.aud_dim, c.text_dim, rg.row_off,
                                               rg.txt_off, rg.txt_lens, lens, B, T, xcat, st));
  else
    RUN(RP_TAG_CAST, launch_concat_cast(vis, aud, txt, c.vis_dim, c.aud_dim, c.text_dim, xcat, M, st));
  RUN(RP_TAG_GEMM_IN, launch_gemm(EPI_BIAS_F32, xcat, Cin, h->w_in, Cin, w.h, D, h->b_in, nullptr, 0, M, D, Cin, st, rows));
  {
    LnArgs a{};
    a.x = w.h; a.M = M; a.T = T; a.eps = eps;
    a.g0 = h->in_g; a.b0 = h->in_b; a.pe = h->pe;
    a.g1 = h->layers[0].n1_g; a.b1 = h->layers[0].n1_b;
    a.out_f32 = w.h; a.y_bf16 = w.u;
    RUN(RP_TAG_LAYERNORM, launch_layernorm512(1, a, st));
  }
  // (2) encoder layers (pre-LN): h += MHA(LN1(h)); h += FFN(LN2(h))
  // RP_LN_IN_GEMM (default 1; 0 = stand-alone LayerNorm kernels): the LayerNorm after each residual update
  // runs inside the GEMM epilogue (needs d_model = 512 and more than one 128-row block).  Removes 32
  // launches / 1.05 ms per step, costs the out-proj / FF2 GEMMs 0.7 ms: +2.4 % (profiles/r01_notes.md).
  static const bool ln_in_gemm_env = !(getenv("RP_LN_IN_GEMM") && atoi(getenv("RP_LN_IN_GEMM")) == 0);
  const bool ln_in_gemm = ln_in_gemm_env && D == 512 && M > 128;
  for (int l = 0; l < c.num_layers; ++l) {
    const LayerW& L = h->layers[l];
    RUN(RP_TAG_GEMM_QKV, launch_gemm(EPI_BIAS_BF16, w.u, D, L.w_qkv, D, w.qkv, 3 * D, L.b_qkv, nullptr, 0, M, 3 * D, D, st, rows));
    FmhaArgs fa{};
    fa.q = w.qkv; fa.k = w.qkv + D; fa.v = w.qkv + 2 * D; fa.o = w.attn;
    fa.ldq = fa.ldk = fa.ldv = 3 * D; fa.ldo = D;
    fa.bsq = fa.bsk = fa.bsv = int64_t(T) * 3 * D; fa.bso = int64_t(T) * D;
    fa.B = B; fa.H = H; fa.Tq = T; fa.Tk = T; fa.kv_lens = lens; fa.mask_mode = 0; fa.skip_padded_queries = skip;
    fa.batch_order = ranked ? w.batch_order : nullptr;
    RUN(RP_TAG_FMHA, launch_fmha(fa, st));
    if (ln_in_gemm) {
      // The LayerNorm that follows each residual update runs inside the GEMM epilogue (clusters of
      // four CTAs own complete 512-column rows): u = LN(h) is written next to the updated h.
      GemmLnFusion f{};
      f.eps = eps; f.u_out = w.u; f.ld_u = D;
      f.ln_gamma = L.n2_g; f.ln_beta = L.n2_b;
      RUN(RP_TAG_GEMM_OUT, launch_gemm_ln(EPI_BIAS_RESID_LN, w.attn, D, L.w_out, D, w.h, D, L.b_out, w.h, D, M, D, D, f, st, rows));
      RUN(RP_TAG_GEMM_FF1, launch_gemm(EPI_BIAS_RELU_BF16, w.u, D, L.w_ff1, D, w.ffn, F, L.b_ff1, nullptr, 0, M, F, D, st, rows));
      if (l + 1 < c.num_layers) { f.ln_gamma = h->layers[l + 1].n1_g; f.ln_beta = h->layers[l + 1].n1_b; }
      else { f.ln_gamma = h->enc_g; f.ln_beta = h->enc_b; }  // encoder_norm feeds feature_map
      RUN(RP_TAG_GEMM_FF2, launch_gemm_ln(EPI_BIAS_RESID_LN, w.ffn, F, L.w_ff2, F, w.h, D, L.b_ff2, w.h, D, M, D, F, f, st, rows));
    } else {
      RUN(RP_TAG_GEMM_OUT, launch_gemm(EPI_BIAS_RESID_F32, w.attn, D, L.w_out, D, w.h, D, L.b_out, w.h, D, M, D, D, st, rows));
      {
        LnArgs a{};
        a.x = w.h; a.M = M; a.T = T; a.eps = eps; a.g0 = L.n2_g; a.b0 = L.n2_b; a.y_bf16 = w.u;
        RUN(RP_TAG_LAYERNORM, launch_layernorm512(0, a, st));
      }
      RUN(RP_TAG_GEMM_FF1, launch_gemm(EPI_BIAS_RELU_BF16, w.u, D, L.w_ff1, D, w.ffn, F, L.b_ff1, nullptr, 0, M, F, D, st, rows));
      RUN(RP_TAG_GEMM_FF2, launch_gemm(EPI_BIAS_RESID_F32, w.ffn, F, L.w_ff2, F, w.h, D, L.b_ff2, w.h, D, M, D, F, st, rows));
      {
        LnArgs a{};
        a.x = w.h; a.M = M; a.T = T; a.eps = eps; a.y_bf16 = w.u;
        if (l + 1 < c.num_layers) { a.g0 = h->layers[l + 1].n1_g; a.b0 = h->layers[l + 1].n1_b; }
        else { a.g0 = h->enc_g; a.b0 = h->enc_b; }  // encoder_norm feeds feature_map
        RUN(RP_TAG_LAYERNORM, launch_layernorm512(0, a, st));
      }
    }
  }
  // (3) feature_map: Linear -> LN -> ReLU = feats (returned); head LayerNorms
  RUN(RP_TAG_GEMM_FMAP, launch_gemm(EPI_BIAS_F32, w.u, D, h->w_fm, D, w.h, D, h->b_fm, nullptr, 0, M, D, D, st, rows));
  {
    LnArgs a{};
    a.x = w.h; a.M = M; a.T = T; a.eps = eps;
    a.g0 = h->fm_g; a.b0 = h->fm_b; a.g1 = h->c0_g; a.b1 = h->c0_b; a.g2 = h->r0_g; a.b2 = h->r0_b;
    a.out_f32 = out_feats; a.y_bf16 = w.u; a.y2_bf16 = w.attn;
    RUN(RP_TAG_LAYERNORM, launch_layernorm512(2, a, st));
  }
  // (4) heads: 512->256 ReLU -> 256->256 ReLU -> {1, 2 (+ReLU)}
  __nv_bfloat16* a1c = w.qkv;
  __nv_bfloat16* a2c = a1c + int64_t(M) * Hh;
  __nv_bfloat16* a1r = a2c + int64_t(M) * Hh;
  __nv_bfloat16* a2r = a1r + int64_t(M) * Hh;
  RUN(RP_TAG_GEMM_HEAD, launch_gemm(EPI_BIAS_RELU_BF16, w.u, D, h->w_c1, D, a1c, Hh, h->b_c1, nullptr, 0, M, Hh, D, st, rows));
  RUN(RP_TAG_GEMM_HEAD, launch_gemm(EPI_BIAS_RELU_BF16, w.attn, D, h->w_r1, D, a1r, Hh, h->b_r1, nullptr, 0, M, Hh, D, st, rows));
  if (Hh == 256) {
    // Linear(256,256) + ReLU + Linear(256, 1 | 2) (+ ReLU) in one kernel per head: the second hidden activation never
    // leaves the epilogue registers; no [M,256] write, no head_out pass
    RUN(RP_TAG_GEMM_HEAD, launch_gemm_head_dot(a1c, Hh, h->w_c4, Hh, h->b_c4, h->w_c7, h->b_c7, 1, false, out_logits, M, Hh, st, rows));
    RUN(RP_TAG_GEMM_HEAD, launch_gemm_head_dot(a1r, Hh, h->w_r4, Hh, h->b_r4, h->w_r7, h->b_r7, 2, true, out_offsets, M, Hh, st, rows));
  } else {
    RUN(RP_TAG_GEMM_HEAD, launch_gemm(EPI_BIAS_RELU_BF16, a1c, Hh, h->w_c4, Hh, a2c, Hh, h->b_c4, nullptr, 0, M, Hh, Hh, st, rows));
    RUN(RP_TAG_GEMM_HEAD, launch_gemm(EPI_BIAS_RELU_BF16, a1r, Hh, h->w_r4, Hh, a2r, Hh, h->b_r4, nullptr, 0, M, Hh, Hh, st, rows));
    RUN(RP_TAG_HEAD_OUT, launch_head_out(a2c, a2r, h->w_c7, h->b_c7, h->w_r7, h->b_r7, out_logits, out_offsets, M, st));
  }
  if (skip) {
    rc = launch_zero_padded_rows(out_logits, out_offsets, out_feats, lens, B, T, D, st);
    if (rc) return rc;
  }
#undef RUN
  return RP_OK;
}

int32_t rp_forward(rp_handle* h, const float* vis, const float* aud, const float* txt,
                   const int32_t* lens, int32_t B, int32_t T, float* out_logits, float* out_offsets,
                   float* out_feats, void* workspace, int64_t workspace_bytes, void* stream) {
  return forward_core(h, vis, aud, txt, RaggedIn{}, lens, B, T, out_logits, out_offsets, out_feats,
                      workspace, workspace_bytes, stream);
}

int32_t rp_forward_ragged(rp_handle* h, const float* vis, const float* aud, const float* txt,
                          const int32_t* row_off, const int32_t* txt_off, const int32_t* txt_lens,
                          const int32_t* lens, int32_t B, int32_t T, float* out_logits,
                          float* out_offsets, float* out_feats, void* workspace,
                          int64_t workspace_bytes, void* stream) {
  RP_CHECK(row_off && txt_off && txt_lens, "rp_forward_ragged: null offsets");
  RaggedIn rg;
  rg.row_off = row_off; rg.txt_off = txt_off; rg.txt_lens = txt_lens;
  return forward_core(h, vis, aud, txt, rg, lens, B, T, out_logits, out_offsets, out_feats, workspace,
                      workspace_bytes, stream);
}

int32_t rp_forward_ragged_bf16(rp_handle* h, const void* vis, const void* aud, const void* txt,
                               const int32_t* row_off, const int32_t* txt_off, const int32_t* txt_lens,
                               const int32_t* lens, int32_t B, int32_t T, float* out_logits,
                               float* out_offsets, float* out_feats, void* workspace,
                               int64_t workspace_bytes, void* stream) {
  RP_CHECK(row_off && txt_off && txt_lens, "rp_forward_ragged_bf16: null offsets");
  RaggedIn rg;
  rg.row_off = row_off; rg.txt_off = txt_off; rg.txt_lens = txt_lens; rg.bf16 = true;
  return forward_core(h, vis, aud, txt, rg, lens, B, T, out_logits, out_offsets, out_feats, workspace,
                      workspace_bytes, stream);
}

int32_t rp_set_skip_padding(rp_handle* h, int32_t on) {
  RP_CHECK(h != nullptr, "rp_set_skip_padding: null handle");
  h->skip_padding = on != 0;
  return RP_OK;
}

int32_t rp_profile_begin(rp_handle* h) {
  RP_CHECK(h != nullptr, "rp_profile_begin: null handle");
  h->prof_on = true;
  h->prof_used = 0;
  h->prof_recs.clear();
  return RP_OK;
}

int32_t rp_profile_end(rp_handle* h, float* ms_by_tag, int32_t* launches_by_tag) {
  RP_CHECK(h != nullptr && ms_by_tag != nullptr && launches_by_tag != nullptr,
           "rp_profile_end: null argument");
  h->prof_on = false;
  for (int i = 0; i < RP_NUM_TAGS; ++i) {
    ms_by_tag[i] = 0.f;
    launches_by_tag[i] = 0;
  }
  for (const auto& r : h->prof_recs) {
    RP_CUDA_CHECK(cudaEventSynchronize(r.e1));
    float ms = 0.f;
    RP_CUDA_CHECK(cudaEventElapsedTime(&ms, r.e0, r.e1));
    ms_by_tag[r.tag] += ms;
    launches_by_tag[r.tag] += 1;
  }
  h->prof_recs.clear();
  h->prof_used = 0;
  return RP_OK;
}

int32_t rp_decode_nms(const float* logits, const float* offsets, const int32_t* lens,
                      const int32_t* max_seg, int32_t B, int32_t T, const rp_decode_cfg* cfg,
                      int32_t Kcap, float* segs, float* scores, float* dscores, int32_t* labels,
                      int32_t* counts, int32_t* ncand, float* cand_segs, float* cand_scores,
                      int32_t* cand_labels, void* stream) {
  RP_CHECK(logits && offsets && lens && max_seg && cfg && segs && scores && dscores && labels &&
               counts && ncand,
           "rp_decode_nms: null argument");
  DecodeCfg d{cfg->pre_nms_topk, cfg->pre_nms_thresh, cfg->duration_thresh, cfg->duration_thresh_max,
              cfg->nms_sigma, cfg->min_score};
  return launch_decode_nms(logits, offsets, lens, max_seg, B, T, d, Kcap, segs, scores, dscores,
                           labels, counts, ncand, cand_segs, cand_scores, cand_labels,
                           reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_soft_nms(const float* scores, const float* segs, const int32_t* n, const int32_t* max_seg,
                    int32_t B, int32_t Nmax, float sigma, float thresh, int32_t Kcap, int32_t* keep,
                    float* kscores, int32_t* counts, void* stream) {
  RP_CHECK(scores && segs && n && max_seg && keep && kscores && counts, "rp_soft_nms: null argument");
  return launch_soft_nms(scores, segs, n, max_seg, B, Nmax, sigma, thresh, Kcap, keep, kscores,
                         counts, reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_atiou(const float* slots, int32_t n_videos, int32_t K, const double* gt,
                 const int32_t* gt_counts, int32_t Gmax, const double* thresholds, int32_t n_thr,
                 double* per_video, double* out, void* stream) {
  RP_CHECK(slots && gt && gt_counts && thresholds && per_video && out, "rp_atiou: null argument");
  return launch_atiou(slots, n_videos, K, gt, gt_counts, Gmax, thresholds, n_thr, per_video, out,
                      reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_focal_loss_sum(const float* logits, const float* targets, const uint8_t* mask, int64_t n, float alpha,
                          float gamma, double* scratch, float* out, void* stream) {
  RP_CHECK(logits && targets && mask && scratch && out, "rp_focal_loss_sum: null argument");
  return launch_focal_loss_sum(logits, targets, mask, n, alpha, gamma, scratch, out,
                               reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_focal_loss_grad(const float* logits, const float* targets, const uint8_t* mask, int64_t n, float alpha,
                           float gamma, float scale, float* dlogits, void* stream) {
  RP_CHECK(logits && targets && mask && dlogits, "rp_focal_loss_grad: null argument");
  return launch_focal_loss_grad(logits, targets, mask, n, alpha, gamma, scale, dlogits,
                                reinterpret_cast<cudaStream_t>(stream));
}

int64_t rp_layernorm512_bwd_scratch_bytes(void) { return layernorm512_bwd_scratch_floats() * int64_t(sizeof(float)); }

int32_t rp_layernorm512_bwd(const float* x, const float* dy, const float* gamma, int64_t M, float eps, float* dx,
                            float* dgamma, float* dbeta, void* scratch, int64_t scratch_bytes, void* stream) {
  RP_CHECK(x && dy && gamma && dx && dgamma && dbeta && scratch, "rp_layernorm512_bwd: null argument");
  RP_CHECK(scratch_bytes >= rp_layernorm512_bwd_scratch_bytes(), "rp_layernorm512_bwd: scratch too small");
  return launch_layernorm512_bwd(x, dy, gamma, M, eps, dx, dgamma, dbeta, reinterpret_cast<float*>(scratch),
                                 reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_cast_scaled(const float* in, int64_t n, int64_t n_scaled, float scale, void* out_bf16, float* out_f32,
                       void* stream) {
  RP_CHECK(in && (out_bf16 || out_f32) && n > 0, "rp_cast_scaled: null argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = int(std::min<int64_t>((n + 255) / 256, 4096));
  if (out_bf16) repack_bf16_kernel<<<grid, 256, 0, st>>>(in, reinterpret_cast<__nv_bfloat16*>(out_bf16), n, n_scaled, scale);
  if (out_f32) repack_f32_kernel<<<grid, 256, 0, st>>>(in, out_f32, n, n_scaled, scale);
  count_launch((out_bf16 ? 1 : 0) + (out_f32 ? 1 : 0));
  RP_CUDA_CHECK(cudaGetLastError());
  return RP_OK;
}

int64_t rp_train_scratch_bytes(void) { return train_scratch_floats() * int64_t(sizeof(float)); }

int32_t rp_layernorm512_bwd_acc(const float* x, const float* dy, const float* gamma, int64_t M, float eps,
                                int32_t accumulate, float* dh_inout, void* dh_bf16, float* dh_colsum, float* dgamma,
                                float* dbeta, void* scratch, int64_t scratch_bytes, void* stream) {
  RP_CHECK(x && dy && gamma && dh_inout && dgamma && dbeta && scratch, "rp_layernorm512_bwd_acc: null argument");
  RP_CHECK(scratch_bytes >= rp_train_scratch_bytes(), "rp_layernorm512_bwd_acc: scratch too small");
  return launch_layernorm512_bwd(x, dy, gamma, M, eps, dh_inout, dgamma, dbeta, reinterpret_cast<float*>(scratch),
                                 reinterpret_cast<cudaStream_t>(stream), accumulate != 0, dh_bf16, dh_colsum);
}

int32_t rp_gemm_bwd(int32_t kind, int32_t out_f32, const void* A, int64_t lda, const void* B, int64_t ldb, void* D,
                    int64_t ldd, int32_t M, int32_t N, int32_t K, int32_t splits, void* stream) {
  RP_CHECK(A && B && D, "rp_gemm_bwd: null argument");
  return launch_gemm_bwd(kind, out_f32 != 0, A, lda, B, ldb, D, ldd, M, N, K, splits,
                         reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_splitk_reduce(const float* partials, int32_t splits, int64_t n, float* out, void* stream) {
  RP_CHECK(partials && out, "rp_splitk_reduce: null argument");
  return launch_splitk_reduce(partials, splits, n, out, reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_colsum_bf16(const void* x, int64_t M, int32_t N, float* out, void* scratch, int64_t scratch_bytes,
                       void* stream) {
  RP_CHECK(x && out && scratch, "rp_colsum_bf16: null argument");
  RP_CHECK(scratch_bytes >= rp_train_scratch_bytes(), "rp_colsum_bf16: scratch too small");
  return launch_colsum_bf16(x, M, N, out, reinterpret_cast<float*>(scratch), reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_relu_bwd(void* dy, const void* act, int64_t n, int32_t is_f32, void* stream) {
  RP_CHECK(dy && act, "rp_relu_bwd: null argument");
  return launch_relu_bwd(dy, act, n, is_f32 != 0, 1.0f, reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_relu_bwd_scaled(void* dy, const void* act, int64_t n, int32_t is_f32, float scale, void* stream) {
  RP_CHECK(dy && act, "rp_relu_bwd_scaled: null argument");
  return launch_relu_bwd(dy, act, n, is_f32 != 0, scale, reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_relu_bwd_colsum(void* dy_bf16, const void* act_bf16, int64_t M, int32_t N, float* colsum, void* scratch,
                           int64_t scratch_bytes, void* stream) {
  RP_CHECK(dy_bf16 && act_bf16 && colsum && scratch, "rp_relu_bwd_colsum: null argument");
  RP_CHECK(scratch_bytes >= rp_train_scratch_bytes(), "rp_relu_bwd_colsum: scratch too small");
  return launch_relu_bwd_colsum(dy_bf16, act_bf16, M, N, 1.0f, colsum, reinterpret_cast<float*>(scratch),
                                reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_relu_bwd_colsum_scaled(void* dy_bf16, const void* act_bf16, int64_t M, int32_t N, float scale, float* colsum,
                                  void* scratch, int64_t scratch_bytes, void* stream) {
  RP_CHECK(dy_bf16 && act_bf16 && colsum && scratch, "rp_relu_bwd_colsum_scaled: null argument");
  RP_CHECK(scratch_bytes >= rp_train_scratch_bytes(), "rp_relu_bwd_colsum_scaled: scratch too small");
  return launch_relu_bwd_colsum(dy_bf16, act_bf16, M, N, scale, colsum, reinterpret_cast<float*>(scratch),
                                reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_head_out_bwd(const float* dlogits, const void* a2_bf16, const float* w, int64_t M, void* da2_bf16,
                        float* dw, float* db, void* scratch, int64_t scratch_bytes, void* stream) {
  RP_CHECK(dlogits && a2_bf16 && w && da2_bf16 && dw && db && scratch, "rp_head_out_bwd: null argument");
  RP_CHECK(scratch_bytes >= rp_train_scratch_bytes(), "rp_head_out_bwd: scratch too small");
  return launch_head_out_bwd(dlogits, a2_bf16, w, M, 1.0f, da2_bf16, dw, db, reinterpret_cast<float*>(scratch),
                             reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_head_out_bwd_scaled(const float* dlogits, const void* a2_bf16, const float* w, int64_t M, float scale,
                               void* da2_bf16, float* dw, float* db, void* scratch, int64_t scratch_bytes, void* stream) {
  RP_CHECK(dlogits && a2_bf16 && w && da2_bf16 && dw && db && scratch, "rp_head_out_bwd_scaled: null argument");
  RP_CHECK(scratch_bytes >= rp_train_scratch_bytes(), "rp_head_out_bwd_scaled: scratch too small");
  return launch_head_out_bwd(dlogits, a2_bf16, w, M, scale, da2_bf16, dw, db, reinterpret_cast<float*>(scratch),
                             reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_fmha_train(const void* q, const void* k, const void* v, void* o, int64_t ld_qkv, int64_t ld_o, int32_t B,
                      int32_t H, int32_t T, const int32_t* kv_lens, float* lse, void* stream) {
  RP_CHECK(q && k && v && o && lse, "rp_fmha_train: null argument");
  FmhaArgs a{};
  a.q = q; a.k = k; a.v = v; a.o = o;
  a.ldq = a.ldk = a.ldv = ld_qkv; a.ldo = ld_o;
  a.bsq = a.bsk = a.bsv = int64_t(T) * ld_qkv; a.bso = int64_t(T) * ld_o;
  a.B = B; a.H = H; a.Tq = T; a.Tk = T; a.kv_lens = kv_lens; a.mask_mode = 0;
  a.lse = lse;
  return launch_fmha(a, reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_fmha_train_dropout(const void* q, const void* k, const void* v, void* o, int64_t ld_qkv, int64_t ld_o,
                              int32_t B, int32_t H, int32_t T, const int32_t* kv_lens, float* lse,
                              const uint32_t* keep_bits, int64_t bits_ld, float p, void* stream) {
  RP_CHECK(q && k && v && o && lse && keep_bits, "rp_fmha_train_dropout: null argument");
  RP_CHECK(p > 0.0f && p < 1.0f, "rp_fmha_train_dropout: 0 < p < 1");
  FmhaArgs a{};
  a.q = q; a.k = k; a.v = v; a.o = o;
  a.ldq = a.ldk = a.ldv = ld_qkv; a.ldo = ld_o;
  a.bsq = a.bsk = a.bsv = int64_t(T) * ld_qkv; a.bso = int64_t(T) * ld_o;
  a.B = B; a.H = H; a.Tq = T; a.Tk = T; a.kv_lens = kv_lens; a.mask_mode = 0;
  a.lse = lse;
  a.drop_bits = keep_bits; a.drop_ld = bits_ld; a.drop_scale = 1.0f / (1.0f - p);
  return launch_fmha(a, reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_fmha_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                    float* dsum, void* dq, void* dk, void* dv, int64_t ld_qkv, int64_t ld_o, int64_t ld_dqkv,
                    int32_t B, int32_t H, int32_t T, const int32_t* kv_lens, void* stream) {
  FmhaBwdArgs a{q, k, v, o, d_o, lse, dsum, dq, dk, dv, ld_qkv, ld_o, ld_dqkv, B, H, T, kv_lens};
  return launch_fmha_bwd(a, reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_set_attn_bwd_deterministic(int32_t on) {
  set_fmha_bwd_deterministic(on);
  return RP_OK;
}

int32_t rp_fmha_bwd_dropout(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                            float* dsum, void* dq, void* dk, void* dv, int64_t ld_qkv, int64_t ld_o, int64_t ld_dqkv,
                            int32_t B, int32_t H, int32_t T, const int32_t* kv_lens, const uint32_t* keep_bits,
                            int64_t bits_ld, float p, void* stream) {
  RP_CHECK(keep_bits != nullptr && p > 0.0f && p < 1.0f, "rp_fmha_bwd_dropout: keep bits and 0 < p < 1 required");
  FmhaBwdArgs a{q, k, v, o, d_o, lse, dsum, dq, dk, dv, ld_qkv, ld_o, ld_dqkv, B, H, T, kv_lens};
  a.drop_bits = keep_bits; a.drop_ld = bits_ld; a.drop_scale = 1.0f / (1.0f - p);
  return launch_fmha_bwd(a, reinterpret_cast<cudaStream_t>(stream));
}

static DropKey drop_key_of(const rp_dropout* d) {
  return (d == nullptr || !(d->p > 0.0f)) ? DropKey{} : make_drop_key(d->key_a, d->key_b, d->p);
}
#define RP_CHECK_DROP(d, what) RP_CHECK((d) == nullptr || ((d)->p >= 0.0f && (d)->p < 1.0f), what ": 0 <= p < 1")

int32_t rp_dropout_mask_u8(const rp_dropout* drop, int64_t n, uint8_t* keep, void* stream) {
  RP_CHECK(drop && keep, "rp_dropout_mask_u8: null argument");
  RP_CHECK_DROP(drop, "rp_dropout_mask_u8");
  return launch_dropout_mask_u8(keep, n, drop_key_of(drop), reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_attn_dropout_bits(const rp_dropout* drop, int64_t n_words, uint32_t* keep_bits, void* stream) {
  RP_CHECK(drop && keep_bits, "rp_attn_dropout_bits: null argument");
  RP_CHECK_DROP(drop, "rp_attn_dropout_bits");
  return launch_attn_dropout_bits(keep_bits, n_words, drop_key_of(drop), reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_layernorm512_bwd_acc_dropout(const float* x, const float* dy, const float* gamma, int64_t M, float eps,
                                        int32_t accumulate, float* dh_inout, void* dh_bf16, float* dh_colsum,
                                        float* dgamma, float* dbeta, void* scratch, int64_t scratch_bytes,
                                        const rp_dropout* drop, void* stream) {
  RP_CHECK(x && dy && gamma && dh_inout && dgamma && dbeta && scratch, "rp_layernorm512_bwd_acc_dropout: null argument");
  RP_CHECK(scratch_bytes >= rp_train_scratch_bytes(), "rp_layernorm512_bwd_acc_dropout: scratch too small");
  RP_CHECK_DROP(drop, "rp_layernorm512_bwd_acc_dropout");
  const DropKey key = drop_key_of(drop);
  return launch_layernorm512_bwd(x, dy, gamma, M, eps, dh_inout, dgamma, dbeta, reinterpret_cast<float*>(scratch),
                                 reinterpret_cast<cudaStream_t>(stream), accumulate != 0, dh_bf16, dh_colsum,
                                 key.thr != 0u ? &key : nullptr);
}

int32_t rp_gemm_bf16_dropout(int32_t epilogue, const void* A, int64_t lda, const void* W, int64_t ldw, void* D,
                             int64_t ldd, const float* bias, const float* resid, int64_t ldr, int32_t M, int32_t N,
                             int32_t K, const rp_dropout* drop, void* stream) {
  RP_CHECK(A && W && D, "rp_gemm_bf16_dropout: null argument");
  RP_CHECK_DROP(drop, "rp_gemm_bf16_dropout");
  return launch_gemm_dropout(epilogue, A, lda, W, ldw, D, ldd, bias, resid, ldr, M, N, K, drop_key_of(drop),
                             reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                     float beta1, float beta2, float eps, float weight_decay, int32_t step, void* param_bf16,
                     void* stream) {
  RP_CHECK(param && grad && exp_avg && exp_avg_sq, "rp_adam_step: null argument");
  return launch_adam_step(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, param_bf16,
                          reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_gemm_bf16(int32_t epilogue, const void* A, int64_t lda, const void* W, int64_t ldw, void* D,
                     int64_t ldd, const float* bias, const float* resid, int64_t ldr, int32_t M,
                     int32_t N, int32_t K, void* stream) {
  RP_CHECK(A && W && D, "rp_gemm_bf16: null argument");
  return launch_gemm(epilogue, A, lda, W, ldw, D, ldd, bias, resid, ldr, M, N, K,
                     reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_fmha(const void* q, const void* k, const void* v, void* o, int64_t ldq, int64_t ldk,
                int64_t ldv, int64_t ldo, int64_t bsq, int64_t bsk, int64_t bsv, int64_t bso,
                int32_t B, int32_t H, int32_t Tq, int32_t Tk, const int32_t* kv_lens, int32_t mask_mode,
                const uint8_t* mask, int64_t mask_b_stride, int64_t mask_q_stride, void* stream) {
  RP_CHECK(q && k && v && o, "rp_fmha: null argument");
  FmhaArgs a{};
  a.q = q; a.k = k; a.v = v; a.o = o;
  a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo;
  a.bsq = bsq; a.bsk = bsk; a.bsv = bsv; a.bso = bso;
  a.B = B; a.H = H; a.Tq = Tq; a.Tk = Tk; a.kv_lens = kv_lens; a.mask_mode = mask_mode;
  a.mask = mask; a.mask_b_stride = mask_b_stride; a.mask_q_stride = mask_q_stride;
  return launch_fmha(a, reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_gemm_head_dot(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const float* w_last,
                         const float* b_last, int32_t nj, int32_t final_relu, float* out, int32_t M, int32_t K,
                         void* stream) {
  return launch_gemm_head_dot(A, lda, W, ldw, bias, w_last, b_last, nj, final_relu != 0, out, M, K,
                              reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_gemm_resid_ln(const void* A, int64_t lda, const void* W, int64_t ldw, float* h, int64_t ldh,
                         const float* bias, const float* gamma, const float* beta, float eps, void* u_bf16,
                         int64_t ldu, int32_t M, int32_t K, void* stream) {
  RP_CHECK(A && W && h && gamma && beta && u_bf16, "rp_gemm_resid_ln: null argument");
  GemmLnFusion f{};
  f.eps = eps; f.ln_gamma = gamma; f.ln_beta = beta; f.u_out = u_bf16; f.ld_u = ldu;
  return launch_gemm_ln(EPI_BIAS_RESID_LN, A, lda, W, ldw, h, ldh, bias, h, ldh, M, 512, K, f,
                        reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_concat_cast(const float* vis, const float* aud, const float* txt, int32_t Cv, int32_t Ca,
                       int32_t Ct, void* out_bf16, int64_t M, void* stream) {
  RP_CHECK(vis && aud && txt && out_bf16, "rp_concat_cast: null argument");
  return launch_concat_cast(vis, aud, txt, Cv, Ca, Ct, out_bf16, M,
                            reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_mask_lens(const uint8_t* mask, int32_t B, int32_t T, int32_t* lens, int32_t* not_aligned, void* stream) {
  RP_CHECK(mask && lens && not_aligned, "rp_mask_lens: null argument");
  return launch_mask_lens(mask, B, T, lens, not_aligned, reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_cast_bf16(const float* in, void* out_bf16, int64_t n, void* stream) {
  RP_CHECK(in && out_bf16, "rp_cast_bf16: null argument");
  return launch_cast_bf16(in, out_bf16, n, reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_layernorm512(int32_t mode, const float* x, int64_t M, int32_t T, const float* g0,
                        const float* b0, const float* g1, const float* b1, const float* g2,
                        const float* b2, const float* pe, float* out_f32, void* y_bf16, void* y2_bf16,
                        void* stream) {
  RP_CHECK(x && g0 && b0, "rp_layernorm512: null argument");
  LnArgs a{};
  a.x = x; a.M = M; a.T = T; a.g0 = g0; a.b0 = b0; a.g1 = g1; a.b1 = b1; a.g2 = g2; a.b2 = b2;
  a.pe = pe; a.out_f32 = out_f32; a.y_bf16 = y_bf16; a.y2_bf16 = y2_bf16; a.eps = 1e-5f;
  return launch_layernorm512(mode, a, reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_layernorm512_dropout(int32_t mode, const float* x, int64_t M, int32_t T, const float* g0,
                                const float* b0, const float* g1, const float* b1, const float* g2,
                                const float* b2, const float* pe, float* out_f32, void* y_bf16, void* y2_bf16,
                                const rp_dropout* drop, void* stream) {
  RP_CHECK(x && g0 && b0, "rp_layernorm512_dropout: null argument");
  RP_CHECK(mode == 2, "rp_layernorm512_dropout: only mode 2 (feature_map) is followed by a Dropout");
  RP_CHECK_DROP(drop, "rp_layernorm512_dropout");
  RP_CHECK(M * 512 < (int64_t(1) << 32), "rp_layernorm512_dropout: dropout sites hold fewer than 2^32 elements");
  LnArgs a{};
  a.x = x; a.M = M; a.T = T; a.g0 = g0; a.b0 = b0; a.g1 = g1; a.b1 = b1; a.g2 = g2; a.b2 = b2;
  a.pe = pe; a.out_f32 = out_f32; a.y_bf16 = y_bf16; a.y2_bf16 = y2_bf16; a.eps = 1e-5f;
  a.drop = drop_key_of(drop);
  return launch_layernorm512(mode, a, reinterpret_cast<cudaStream_t>(stream));
}

int32_t rp_head_out(const void* a_cls_bf16, const void* a_reg_bf16, const float* w_cls,
                    const float* b_cls, const float* w_reg, const float* b_reg, float* logits,
                    float* offsets, int64_t M, void* stream) {
  RP_CHECK(a_cls_bf16 && a_reg_bf16 && w_cls && b_cls && w_reg && b_reg && logits && offsets,
           "rp_head_out: null argument");
  return launch_head_out(a_cls_bf16, a_reg_bf16, w_cls, b_cls, w_reg, b_reg, logits, offsets, M,
                         reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
