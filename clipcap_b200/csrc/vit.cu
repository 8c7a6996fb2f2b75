// Stage 1 — CLIP ViT image tower engine (replaces clip_model.encode_image behind CLIPModel.forward,
// clipcap/encoders/clip.py:112-129).  OpenAI CLIP VisionTransformer semantics:
//   patches = conv1(x) (14x14 stride 14, no bias) == GEMM over gathered non-overlapping patches
//   h = LN_pre(cat(class_emb, patches) + pos);  24 x pre-LN block (QuickGELU MLP, full attention)
//   e = LN_post(h[:,0]) @ proj;  optional e /= ||e||  (clip.py:122-123)
#include <string>

#include <cstdlib>

#include "common.h"

struct cc_vit {
  cc_vit_cfg cfg;
  int max_batch = 0;
  int T = 0;      // tokens per image (1 + grid^2)
  int k_pad = 0;  // 3*p*p rounded up to a multiple of 8 (16-byte rows for TMA)
  cc::Arena arena;
  cc::Stack st;
  __half* cols16 = nullptr;   // [B*grid^2, k_pad]
  float* patches32 = nullptr; // [B*grid^2, w]
  __half* cls16 = nullptr;    // [B, w]   LN_post(h[:,0])
  float* out32 = nullptr;     // [B, out]
  const float *cls = nullptr, *pos = nullptr, *lnpre_g = nullptr, *lnpre_b = nullptr, *lnpost_g = nullptr,
              *lnpost_b = nullptr;
  cc::GemmPlan p_patch, p_proj;
  int launches = 0;
};

namespace cc {
namespace {

int vit_build(cc_vit* m, const cc_tensor* w, int nw) {
  const cc_vit_cfg& c = m->cfg;
  Arena stage;
  const int B = m->max_batch;
  const int g = c.image_size / c.patch;
  const int np = g * g;
  const int kk = 3 * c.patch * c.patch;
  m->T = np + 1;
  m->k_pad = (kk + 7) / 8 * 8;
  const int wd = c.width;
  CC_TRY(m->st.init(m->arena, wd, c.mlp_dim, c.heads, EPI_F16_QUICKGELU, false, c.eps, B * m->T));
  CC_TRY(m->arena.alloc_t(&m->cols16, static_cast<size_t>(B) * np * m->k_pad));
  CC_TRY(m->arena.alloc_t(&m->patches32, static_cast<size_t>(B) * np * wd));
  CC_TRY(m->arena.alloc_t(&m->cls16, static_cast<size_t>(B) * wd));
  CC_TRY(m->arena.alloc_t(&m->out32, static_cast<size_t>(B) * c.out_dim));

  const float *conv, *cls, *pos, *g0, *b0, *g1, *b1, *proj;
  CC_TRY(find_weight(w, nw, "visual.conv1.weight", static_cast<int64_t>(wd) * kk, stage, &conv));
  CC_TRY(find_weight(w, nw, "visual.class_embedding", wd, stage, &cls));
  CC_TRY(find_weight(w, nw, "visual.positional_embedding", static_cast<int64_t>(m->T) * wd, stage, &pos));
  CC_TRY(find_weight(w, nw, "visual.ln_pre.weight", wd, stage, &g0));
  CC_TRY(find_weight(w, nw, "visual.ln_pre.bias", wd, stage, &b0));
  CC_TRY(find_weight(w, nw, "visual.ln_post.weight", wd, stage, &g1));
  CC_TRY(find_weight(w, nw, "visual.ln_post.bias", wd, stage, &b1));
  CC_TRY(find_weight(w, nw, "visual.proj", static_cast<int64_t>(wd) * c.out_dim, stage, &proj));
  const __half *convh, *projh;
  CC_TRY(pack_f16(m->arena, conv, wd, kk, false, m->k_pad, &convh));         // [w, k_pad], zero padded columns
  CC_TRY(pack_f16(m->arena, proj, wd, c.out_dim, true, wd, &projh));          // proj [w,out] -> [out, w]
  CC_TRY(keep_f32(m->arena, cls, wd, &m->cls));
  CC_TRY(keep_f32(m->arena, pos, static_cast<size_t>(m->T) * wd, &m->pos));
  CC_TRY(keep_f32(m->arena, g0, wd, &m->lnpre_g));
  CC_TRY(keep_f32(m->arena, b0, wd, &m->lnpre_b));
  CC_TRY(keep_f32(m->arena, g1, wd, &m->lnpost_g));
  CC_TRY(keep_f32(m->arena, b1, wd, &m->lnpost_b));
  CC_TRY(gemm_plan(&m->p_patch, m->cols16, m->k_pad, B * np, convh, wd, m->k_pad, EPI_F32, nullptr, m->patches32, wd));
  CC_TRY(gemm_plan(&m->p_proj, m->cls16, wd, B, projh, c.out_dim, wd, EPI_F32, nullptr, m->out32, c.out_dim));
  stage.release();

  m->st.layers.resize(c.layers);
  for (int l = 0; l < c.layers; ++l) {
    const std::string p = "visual.transformer.resblocks." + std::to_string(l) + ".";
    LayerW& L = m->st.layers[l];
    const float *a, *b, *cc_, *d_, *wi, *bi, *wo, *bo, *w1, *b1_, *w2, *b2;
    CC_TRY(find_weight(w, nw, p + "ln_1.weight", wd, stage, &a));
    CC_TRY(find_weight(w, nw, p + "ln_1.bias", wd, stage, &b));
    CC_TRY(find_weight(w, nw, p + "ln_2.weight", wd, stage, &cc_));
    CC_TRY(find_weight(w, nw, p + "ln_2.bias", wd, stage, &d_));
    CC_TRY(find_weight(w, nw, p + "attn.in_proj_weight", 3LL * wd * wd, stage, &wi));
    CC_TRY(find_weight(w, nw, p + "attn.in_proj_bias", 3 * wd, stage, &bi));
    CC_TRY(find_weight(w, nw, p + "attn.out_proj.weight", static_cast<int64_t>(wd) * wd, stage, &wo));
    CC_TRY(find_weight(w, nw, p + "attn.out_proj.bias", wd, stage, &bo));
    CC_TRY(find_weight(w, nw, p + "mlp.c_fc.weight", static_cast<int64_t>(c.mlp_dim) * wd, stage, &w1));
    CC_TRY(find_weight(w, nw, p + "mlp.c_fc.bias", c.mlp_dim, stage, &b1_));
    CC_TRY(find_weight(w, nw, p + "mlp.c_proj.weight", static_cast<int64_t>(wd) * c.mlp_dim, stage, &w2));
    CC_TRY(find_weight(w, nw, p + "mlp.c_proj.bias", wd, stage, &b2));
    CC_TRY(keep_f32(m->arena, a, wd, &L.ln1_g));
    CC_TRY(keep_f32(m->arena, b, wd, &L.ln1_b));
    CC_TRY(keep_f32(m->arena, cc_, wd, &L.ln2_g));
    CC_TRY(keep_f32(m->arena, d_, wd, &L.ln2_b));
    CC_TRY(pack_f16(m->arena, wi, 3 * wd, wd, false, wd, &L.wqkv));
    CC_TRY(keep_f32(m->arena, bi, 3 * static_cast<size_t>(wd), &L.bqkv));
    CC_TRY(pack_f16(m->arena, wo, wd, wd, false, wd, &L.wo));
    CC_TRY(keep_f32(m->arena, bo, wd, &L.bo));
    CC_TRY(pack_f16(m->arena, w1, c.mlp_dim, wd, false, wd, &L.w1));
    CC_TRY(keep_f32(m->arena, b1_, c.mlp_dim, &L.b1));
    CC_TRY(pack_f16(m->arena, w2, wd, c.mlp_dim, false, c.mlp_dim, &L.w2));
    CC_TRY(keep_f32(m->arena, b2, wd, &L.b2));
    stage.release();
  }
  {
    // ViT-L/14 @ 224 (257 tokens, head dim 64): head-major QKV + tcgen05 attention. CLIPCAP_B200_NO_TC_ATTN=1 keeps the
    // generic path (A/B measurements).
    const char* e = getenv("CLIPCAP_B200_NO_TC_ATTN");
    const bool no_tc = e != nullptr && e[0] == '1';
    if (!no_tc && m->T == kVitAttnTokens && m->st.hd == 64) m->st.heads_S = m->T;
  }
  {
    // Only the class token leaves the tower: the last block's out-proj / MLP run on the B class rows alone.
    // CLIPCAP_B200_FULL_LAST_LAYER=1 computes every row like the reference (A/B measurements).
    const char* e = getenv("CLIPCAP_B200_FULL_LAST_LAYER");
    if (!(e != nullptr && e[0] == '1')) m->st.cls_last_S = m->T;
  }
  CC_TRY(m->st.plan());
  return CC_OK;
}

}  // namespace
}  // namespace cc

extern "C" {

int cc_vit_create(cc_vit** h, const cc_vit_cfg* cfg, const cc_tensor* weights, int n_weights, int max_batch) {
  using namespace cc;
  CC_REQUIRE(h != nullptr && cfg != nullptr && weights != nullptr, CC_EINVAL, "cc_vit_create: null argument");
  *h = nullptr;
  CC_TRY(check_device_sm100());
  CC_REQUIRE(max_batch > 0, CC_EINVAL, "cc_vit_create: max_batch %d", max_batch);
  CC_REQUIRE(cfg->patch > 0 && cfg->image_size > 0 && cfg->image_size % cfg->patch == 0, CC_ESHAPE,
             "vit: image %d not a multiple of patch %d", cfg->image_size, cfg->patch);
  CC_REQUIRE(cfg->width > 0 && cfg->width % 8 == 0 && cfg->mlp_dim % 8 == 0 && cfg->out_dim > 0 && cfg->layers > 0 &&
                 cfg->heads > 0,
             CC_ESHAPE, "vit: width=%d mlp=%d out=%d layers=%d heads=%d", cfg->width, cfg->mlp_dim, cfg->out_dim,
             cfg->layers, cfg->heads);
  cc_vit* m = new cc_vit();
  m->cfg = *cfg;
  if (m->cfg.eps <= 0.f) m->cfg.eps = 1e-5f;
  m->max_batch = max_batch;
  const int st = vit_build(m, weights, n_weights);
  if (st != CC_OK) {
    delete m;
    return st;
  }
  *h = m;
  return CC_OK;
}

int cc_vit_forward(cc_vit* m, const void* pixels, int pix_dtype, int B, int normalize, void* out, int out_dtype,
                   void* stream) {
  using namespace cc;
  CC_REQUIRE(m != nullptr && pixels != nullptr && out != nullptr, CC_EINVAL, "cc_vit_forward: null argument");
  CC_REQUIRE(B > 0 && B <= m->max_batch, CC_ESHAPE, "cc_vit_forward: batch %d outside 1..%d", B, m->max_batch);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const cc_vit_cfg& c = m->cfg;
  const int np = m->T - 1;
  CC_TRY(vit_im2col_run(pixels, pix_dtype, m->cols16, B, c.image_size, c.patch, m->k_pad, s));
  CC_TRY(gemm_run(m->p_patch, B * np, s));
  CC_TRY(vit_embed_lnpre_run(m->patches32, m->cls, m->pos, m->lnpre_g, m->lnpre_b, m->st.h, B, m->T, c.width, c.eps, s));
  m->st.launches = 0;
  const int full_layers = m->st.cls_last_S > 0 ? c.layers - 1 : c.layers;
  for (int l = 0; l < full_layers; ++l) CC_TRY(m->st.layer_full(l, B, m->T, nullptr, 0, s));
  if (full_layers < c.layers) CC_TRY(m->st.layer_cls_only(c.layers - 1, B, m->T, s));
  // LN_post on the class token of every image (row b*T of h), then the output projection
  CC_TRY(layernorm_run(m->st.h, static_cast<int64_t>(m->T) * c.width, m->lnpost_g, m->lnpost_b, m->cls16, c.width, B,
                       c.width, c.eps, s));
  CC_TRY(gemm_run(m->p_proj, B, s));
  m->launches = m->st.launches + 5;
  if (normalize) {
    CC_TRY(l2_normalize_run(m->out32, B, c.out_dim, s));
    m->launches += 1;
  }
  CC_TRY(convert_from_f32_run(m->out32, c.out_dim, out, out_dtype, B, c.out_dim, s));
  m->launches += 1;
  return CC_OK;
}

int cc_vit_last_launches(cc_vit* m) { return m ? m->launches : 0; }

void cc_vit_destroy(cc_vit* m) { delete m; }

}  // extern "C"
