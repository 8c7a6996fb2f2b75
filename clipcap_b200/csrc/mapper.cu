// Stage 2 — prefix mapper engine.
// TransformerMapper.forward (clipcap/model/mapper.py:122-130), TransformerMapperWindowed.forward (mapper.py:148-160) and
// the upstream-defined MLP mapper.  The input projection GEMM writes straight into rows 0..P-1 of the
// [B, S, d] fp32 residual stream (the reference's .view + torch.cat), rows P.. are filled from prefix_const.
#include <string>

#include "common.h"

struct cc_mapper {
  cc_mapper_cfg cfg;
  int max_batch = 0;
  int S = 0, Ptot = 0;  // tokens per sample; projected tokens (W*P when windowed)
  cc::Arena arena;
  cc::Stack st;
  __half* emb16 = nullptr;
  const float* prefix_const = nullptr;
  const float* pos_emb = nullptr;
  std::vector<cc::GemmPlan> p_lin;  // one per window slot (1 when not windowed)
  // MLP kind
  __half* hid16 = nullptr;
  float* out32 = nullptr;
  cc::GemmPlan p_m0, p_m2;
  int launches = 0;
};

namespace cc {
namespace {

int mapper_build(cc_mapper* m, const cc_tensor* w, int nw) {
  const cc_mapper_cfg& c = m->cfg;
  Arena stage;  // fp32 staging of host tensors, dropped when create returns
  const int B = m->max_batch;
  if (c.kind == CC_MAPPER_MLP) {
    CC_REQUIRE((static_cast<int64_t>(c.K) * c.d) % 2 == 0, CC_ESHAPE, "mlp mapper: K*d must be even");
    const int hid = c.K * c.d / 2, out = c.K * c.d;
    const float *w0, *b0, *w2, *b2;
    CC_TRY(find_weight(w, nw, "model.0.weight", static_cast<int64_t>(hid) * c.E, stage, &w0));
    CC_TRY(find_weight(w, nw, "model.0.bias", hid, stage, &b0));
    CC_TRY(find_weight(w, nw, "model.2.weight", static_cast<int64_t>(out) * hid, stage, &w2));
    CC_TRY(find_weight(w, nw, "model.2.bias", out, stage, &b2));
    const __half *w0h, *w2h;
    const float *b0k, *b2k;
    CC_TRY(pack_f16(m->arena, w0, hid, c.E, false, c.E, &w0h));
    CC_TRY(pack_f16(m->arena, w2, out, hid, false, hid, &w2h));
    CC_TRY(keep_f32(m->arena, b0, hid, &b0k));
    CC_TRY(keep_f32(m->arena, b2, out, &b2k));
    CC_TRY(m->arena.alloc_t(&m->emb16, static_cast<size_t>(B) * c.E));
    CC_TRY(m->arena.alloc_t(&m->hid16, static_cast<size_t>(B) * hid));
    CC_TRY(m->arena.alloc_t(&m->out32, static_cast<size_t>(B) * out));
    CC_TRY(gemm_plan(&m->p_m0, m->emb16, c.E, B, w0h, hid, c.E, EPI_F16_TANH, b0k, m->hid16, hid));
    CC_TRY(gemm_plan(&m->p_m2, m->hid16, hid, B, w2h, out, hid, EPI_F32, b2k, m->out32, out));
    return CC_OK;
  }

  const int W = c.kind == CC_MAPPER_WINDOWED ? c.W : 1;
  m->Ptot = W * c.P;
  m->S = m->Ptot + c.K;
  const int d = c.d;
  CC_TRY(m->st.init(m->arena, d, 2 * d, c.H, EPI_F16_RELU, false, c.eps, B * m->S));
  CC_TRY(m->arena.alloc_t(&m->emb16, static_cast<size_t>(B) * W * c.E));

  const float *lw, *lb, *pc;
  CC_TRY(find_weight(w, nw, "linear.weight", static_cast<int64_t>(c.P) * d * c.E, stage, &lw));
  CC_TRY(find_weight(w, nw, "linear.bias", static_cast<int64_t>(c.P) * d, stage, &lb));
  CC_TRY(find_weight(w, nw, "prefix_const", static_cast<int64_t>(c.K) * d, stage, &pc));
  const __half* lwh;
  const float* lbk;
  CC_TRY(pack_f16(m->arena, lw, c.P * d, c.E, false, c.E, &lwh));
  CC_TRY(keep_f32(m->arena, lb, static_cast<size_t>(c.P) * d, &lbk));
  CC_TRY(keep_f32(m->arena, pc, static_cast<size_t>(c.K) * d, &m->prefix_const));
  if (c.kind == CC_MAPPER_WINDOWED && c.use_pos) {
    const float* pe;
    CC_TRY(find_weight(w, nw, "pos_embeddings", static_cast<int64_t>(m->Ptot) * d, stage, &pe));
    CC_TRY(keep_f32(m->arena, pe, static_cast<size_t>(m->Ptot) * d, &m->pos_emb));
  }
  // Input projection: one row per (sample, window); N = P*d contiguous outputs = P consecutive token rows of h.
  // Non-windowed: row b lands at h[b*S*d]; windowed: row (b, w) lands at h[(b*S + w*P)*d], which is only a constant row
  // stride when W == 1, so the windowed variant issues one GEMM per window (W is small: window_size + 1).
  m->p_lin.resize(W);
  for (int wdw = 0; wdw < W; ++wdw)
    CC_TRY(gemm_plan(&m->p_lin[wdw], m->emb16 + static_cast<size_t>(wdw) * c.E, static_cast<int64_t>(W) * c.E, B, lwh,
                     c.P * d, c.E, EPI_F32, lbk, m->st.h + static_cast<size_t>(wdw) * c.P * d,
                     static_cast<int64_t>(m->S) * d));

  m->st.layers.resize(c.L);
  for (int l = 0; l < c.L; ++l) {
    const std::string p = "transformer.layers." + std::to_string(l) + ".";
    LayerW& lw_ = m->st.layers[l];
    const float *g1, *b1, *g2, *b2, *wq, *wkv, *wp, *bp, *f1w, *f1b, *f2w, *f2b;
    CC_TRY(find_weight(w, nw, p + "norm1.weight", d, stage, &g1));
    CC_TRY(find_weight(w, nw, p + "norm1.bias", d, stage, &b1));
    CC_TRY(find_weight(w, nw, p + "norm2.weight", d, stage, &g2));
    CC_TRY(find_weight(w, nw, p + "norm2.bias", d, stage, &b2));
    CC_TRY(find_weight(w, nw, p + "attn.to_queries.weight", static_cast<int64_t>(d) * d, stage, &wq));
    CC_TRY(find_weight(w, nw, p + "attn.to_keys_values.weight", 2LL * d * d, stage, &wkv));
    CC_TRY(find_weight(w, nw, p + "attn.project.weight", static_cast<int64_t>(d) * d, stage, &wp));
    CC_TRY(find_weight(w, nw, p + "attn.project.bias", d, stage, &bp));
    CC_TRY(find_weight(w, nw, p + "mlp.fc1.weight", 2LL * d * d, stage, &f1w));
    CC_TRY(find_weight(w, nw, p + "mlp.fc1.bias", 2 * d, stage, &f1b));
    CC_TRY(find_weight(w, nw, p + "mlp.fc2.weight", 2LL * d * d, stage, &f2w));
    CC_TRY(find_weight(w, nw, p + "mlp.fc2.bias", d, stage, &f2b));
    CC_TRY(keep_f32(m->arena, g1, d, &lw_.ln1_g));
    CC_TRY(keep_f32(m->arena, b1, d, &lw_.ln1_b));
    CC_TRY(keep_f32(m->arena, g2, d, &lw_.ln2_g));
    CC_TRY(keep_f32(m->arena, b2, d, &lw_.ln2_b));
    // fused QKV weight: rows 0..d-1 = to_queries, rows d..3d-1 = to_keys_values (keys then values, attention.py:24-30)
    __half* wqkv = nullptr;
    CC_TRY(m->arena.alloc_t(&wqkv, 3 * static_cast<size_t>(d) * d));
    CC_TRY(pack_weight_run(wq, d, d, false, wqkv, d, nullptr));
    CC_TRY(pack_weight_run(wkv, 2 * d, d, false, wqkv + static_cast<size_t>(d) * d, d, nullptr));
    CC_CUDA(cudaStreamSynchronize(nullptr));
    lw_.wqkv = wqkv;
    CC_TRY(pack_f16(m->arena, wp, d, d, false, d, &lw_.wo));
    CC_TRY(keep_f32(m->arena, bp, d, &lw_.bo));
    CC_TRY(pack_f16(m->arena, f1w, 2 * d, d, false, d, &lw_.w1));
    CC_TRY(keep_f32(m->arena, f1b, 2 * d, &lw_.b1));
    CC_TRY(pack_f16(m->arena, f2w, d, 2 * d, false, 2 * d, &lw_.w2));
    CC_TRY(keep_f32(m->arena, f2b, d, &lw_.b2));
    stage.release();
  }
  CC_TRY(m->st.plan());
  return CC_OK;
}

}  // namespace
}  // namespace cc

extern "C" {

int cc_mapper_create(cc_mapper** h, const cc_mapper_cfg* cfg, const cc_tensor* weights, int n_weights, int max_batch) {
  using namespace cc;
  CC_REQUIRE(h != nullptr && cfg != nullptr && weights != nullptr, CC_EINVAL, "cc_mapper_create: null argument");
  *h = nullptr;
  CC_TRY(check_device_sm100());
  CC_REQUIRE(max_batch > 0, CC_EINVAL, "cc_mapper_create: max_batch %d", max_batch);
  CC_REQUIRE(cfg->kind >= CC_MAPPER_TRANSFORMER && cfg->kind <= CC_MAPPER_MLP, CC_EINVAL, "mapper kind %d", cfg->kind);
  CC_REQUIRE(cfg->E > 0 && cfg->E % 8 == 0 && cfg->d > 0 && cfg->d % 8 == 0 && cfg->K > 0, CC_ESHAPE,
             "mapper: E=%d d=%d K=%d (E and d must be positive multiples of 8)", cfg->E, cfg->d, cfg->K);
  if (cfg->kind != CC_MAPPER_MLP)
    CC_REQUIRE(cfg->P > 0 && cfg->H > 0 && cfg->L > 0, CC_ESHAPE, "mapper: P=%d H=%d L=%d", cfg->P, cfg->H, cfg->L);
  if (cfg->kind == CC_MAPPER_WINDOWED) CC_REQUIRE(cfg->W >= 1, CC_ESHAPE, "windowed mapper: W=%d", cfg->W);
  cc_mapper* m = new cc_mapper();
  m->cfg = *cfg;
  if (m->cfg.eps <= 0.f) m->cfg.eps = 1e-5f;
  m->max_batch = max_batch;
  const int st = mapper_build(m, weights, n_weights);
  if (st != CC_OK) {
    delete m;
    return st;
  }
  *h = m;
  return CC_OK;
}

int cc_mapper_forward(cc_mapper* m, const void* emb, int emb_dtype, int B, void* prefix, int prefix_dtype,
                      void* stream) {
  using namespace cc;
  CC_REQUIRE(m != nullptr && emb != nullptr && prefix != nullptr, CC_EINVAL, "cc_mapper_forward: null argument");
  CC_REQUIRE(B > 0 && B <= m->max_batch, CC_ESHAPE, "cc_mapper_forward: batch %d outside 1..%d", B, m->max_batch);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const cc_mapper_cfg& c = m->cfg;
  m->launches = 0;
  if (c.kind == CC_MAPPER_MLP) {
    CC_TRY(convert_to_f16_run(emb, emb_dtype, m->emb16, static_cast<int64_t>(B) * c.E, s));
    CC_TRY(gemm_run(m->p_m0, B, s));
    CC_TRY(gemm_run(m->p_m2, B, s));
    CC_TRY(convert_from_f32_run(m->out32, static_cast<int64_t>(c.K) * c.d, prefix, prefix_dtype, B, c.K * c.d, s));
    m->launches = 4;
    return CC_OK;
  }
  const int W = c.kind == CC_MAPPER_WINDOWED ? c.W : 1;
  const int d = c.d, S = m->S;
  CC_TRY(convert_to_f16_run(emb, emb_dtype, m->emb16, static_cast<int64_t>(B) * W * c.E, s));
  // one GEMM per window slot: A rows (b, w) at stride W*E, C rows at stride S*d, offset w*P*d
  for (int wdw = 0; wdw < W; ++wdw) CC_TRY(gemm_run(m->p_lin[wdw], B, s));
  CC_TRY(mapper_fill_const_run(m->st.h, m->prefix_const, m->pos_emb, B, m->Ptot, c.K, d, s));
  m->st.launches = 0;
  for (int l = 0; l < c.L; ++l) CC_TRY(m->st.layer_full(l, B, S, nullptr, 0, s));
  CC_TRY(convert_from_f32_run(m->st.h + static_cast<size_t>(m->Ptot) * d, static_cast<int64_t>(S) * d, prefix,
                              prefix_dtype, B, c.K * d, s));
  m->launches = m->st.launches + 3 + W;
  return CC_OK;
}

int cc_mapper_last_launches(cc_mapper* m) { return m ? m->launches : 0; }

void cc_mapper_destroy(cc_mapper* m) { delete m; }

}  // extern "C"
