// Training step of ClipCapModelPrefixOnly: the forward of clipcap/model/model.py:43-58, the loss of
// training_step (model.py:94-113: logits[:, prefix_length-1:-1] against the caption tokens, cross-entropy with
// ignore_index 0) and its backward down to every transformer_mapper parameter, with the language model frozen
// (model.py:116-123). Dense work runs on the tcgen05 GEMM of gemm.cu:
//   forward   y  = x W^T            A = x   [rows, in],  operand W  [out, in]
//   dgrad     dx = dy W             A = dy  [rows, out], operand W^T [in, out]   (GPT-2's Conv1D weights are stored [in, out]
//                                                                               already: the checkpoint layout IS the operand)
//   wgrad     dW = dy^T x           A = dy^T [out, rows], operand x^T [in, rows] (fp16 transposes, fp32 accumulation and output)
// Activations the backward needs are stashed per layer in fp16 (QKV, attention output, MLP pre-activation / hidden) and
// fp32 (residual stream at the two LayerNorm inputs); LayerNorm outputs and softmax probabilities are recomputed.
// Gradients flow in fp16 between GEMMs under a static loss scale (removed when the fp32 parameter gradients are written).
#include <string>

#include "train.h"

struct cc_train {
  cc_mapper_cfg mc;
  cc_gpt2_cfg gc;
  int max_batch = 0, max_tokens = 0;
  int S = 0, T_max = 0, rows_m = 0, rows_l = 0, rows_sel = 0, v_pad = 0, head_chunk = 0;
  int W = 1, Ptot = 0;      // windows per sample (TransformerMapperWindowed: window_size + 1) and projected tokens W * P
  int* nonfinite = nullptr;   // device counter: non-finite gradient elements of the last step with gradients
  float* pos_tmp = nullptr;  // [Ptot * d] column sums of the projected-token gradients (pos_embeddings gradient / bias fold)
  cc::Arena arena;
  // ---- frozen language model
  struct LmLayer {
    const float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *bqkv, *bo, *b1, *b2;
    const __half *wqkv, *wo, *w1, *w2;          // forward operands [out, in]
    const __half *wqkv_t, *wo_t, *w1_t, *w2_t;  // dgrad operands  [in, out] (Conv1D checkpoint orientation)
    float *hin, *hmid;                          // stash: residual stream at LN1 / LN2 input
    __half *qkv, *att, *pre;                    // stash: packed QKV, attention output, fc1 pre-activation
  };
  std::vector<LmLayer> lm;
  const float *wte32 = nullptr, *wpe32 = nullptr, *lnf_g = nullptr, *lnf_b = nullptr;
  const __half *wte16 = nullptr, *wte_t16 = nullptr;  // [V, d] and [d, v_pad]
  // ---- mapper (trainable): fp16 working copies re-packed from the caller's fp32 parameters every step
  struct MapLayer {
    __half *wqkv, *wqkv_t, *wo, *wo_t, *w1, *w1_t, *w2, *w2_t;
    float *hin, *hmid;
    __half *qkv, *att, *hid;
  };
  std::vector<MapLayer> mp;
  __half* lin16 = nullptr;  // [P*d, E]
  // ---- activations / scratch
  __half *emb16 = nullptr, *ln16 = nullptr, *hid16 = nullptr, *g16 = nullptr, *dbig16 = nullptr, *dqkv16 = nullptr,
         *datt16 = nullptr, *lnf16 = nullptr, *dlogits16 = nullptr, *tr_a = nullptr, *tr_b = nullptr, *dlin16 = nullptr;
  float *hm = nullptr, *h = nullptr, *dh = nullptr, *dln32 = nullptr, *hsel = nullptr, *dhsel = nullptr, *logits = nullptr,
        *dlnf32 = nullptr, *row_loss = nullptr, *loss = nullptr, *ln_scratch = nullptr, *wqkv_grad = nullptr,
        *cs_scratch = nullptr;
  size_t cs_floats = 0;
  int32_t *tokens = nullptr, *targets = nullptr;
  int* n_valid = nullptr;
  int launches = 0;
};

namespace cc {
namespace {

// y[M, N] = A[M, K] W[N, K]^T (+ epilogue): plan + launch. Plans are cheap host work (tensor-map encodes).
int gemm(const __half* a, int64_t lda, int M, const __half* w, int N, int K, int epi, const float* bias, void* out,
         int64_t ldc, cudaStream_t s, int* launches) {
  GemmPlan p;
  CC_TRY(gemm_plan(&p, a, lda, M, w, N, K, epi, bias, out, ldc));
  ++*launches;
  return gemm_run(p, M, s);
}

const cc_tensor* find_tensor(const cc_tensor* t, int n, const std::string& name) {
  for (int i = 0; i < n; ++i)
    if (t[i].name != nullptr && name == t[i].name) return &t[i];
  return nullptr;
}

int dev_f32(const cc_tensor* t, int n, const std::string& name, int64_t numel, const float** out) {
  const cc_tensor* x = find_tensor(t, n, name);
  CC_REQUIRE(x != nullptr, CC_EINVAL, "cc_train_step: tensor '%s' missing", name.c_str());
  int64_t have = 1;
  for (int k = 0; k < x->ndim; ++k) have *= x->shape[k];
  CC_REQUIRE(x->dtype == CC_F32 && have == numel && x->data != nullptr, CC_ESHAPE,
             "cc_train_step: tensor '%s' must be fp32 with %lld elements (got %lld)", name.c_str(), (long long)numel,
             (long long)have);
  *out = static_cast<const float*>(x->data);
  return CC_OK;
}

int train_build(cc_train* t, const cc_tensor* w, int nw) {
  const cc_mapper_cfg& mc = t->mc;
  const cc_gpt2_cfg& gc = t->gc;
  const int d = gc.d, B = t->max_batch;
  t->W = mc.kind == CC_MAPPER_WINDOWED ? mc.W : 1;
  t->Ptot = t->W * mc.P;
  t->S = t->Ptot + mc.K;
  t->T_max = mc.K + t->max_tokens;
  t->rows_m = B * t->S;
  t->rows_l = B * t->T_max;
  t->rows_sel = B * t->max_tokens;
  t->v_pad = (gc.V + 7) / 8 * 8;
  t->head_chunk = t->rows_sel < 1024 ? t->rows_sel : 1024;
  Arena stage;
  Arena& A = t->arena;
  const float *wte, *wpe, *g, *b;
  CC_TRY(find_weight(w, nw, "transformer.wte.weight", static_cast<int64_t>(gc.V) * d, stage, &wte));
  CC_TRY(find_weight(w, nw, "transformer.wpe.weight", static_cast<int64_t>(gc.n_pos) * d, stage, &wpe));
  CC_TRY(find_weight(w, nw, "transformer.ln_f.weight", d, stage, &g));
  CC_TRY(find_weight(w, nw, "transformer.ln_f.bias", d, stage, &b));
  CC_TRY(keep_f32(A, wte, static_cast<size_t>(gc.V) * d, &t->wte32));
  CC_TRY(keep_f32(A, wpe, static_cast<size_t>(gc.n_pos) * d, &t->wpe32));
  CC_TRY(keep_f32(A, g, d, &t->lnf_g));
  CC_TRY(keep_f32(A, b, d, &t->lnf_b));
  CC_TRY(pack_f16(A, wte, gc.V, d, false, d, &t->wte16));
  CC_TRY(pack_f16(A, wte, gc.V, d, true, t->v_pad, &t->wte_t16));
  stage.release();

  const size_t rl = static_cast<size_t>(t->rows_l), rm = static_cast<size_t>(t->rows_m);
  t->lm.resize(gc.L);
  for (int l = 0; l < gc.L; ++l) {
    const std::string p = "transformer.h." + std::to_string(l) + ".";
    cc_train::LmLayer& L = t->lm[l];
    const float *g1, *b1, *g2, *b2, *wa, *ba, *wo, *bo, *wf, *bf, *wp, *bp;
    CC_TRY(find_weight(w, nw, p + "ln_1.weight", d, stage, &g1));
    CC_TRY(find_weight(w, nw, p + "ln_1.bias", d, stage, &b1));
    CC_TRY(find_weight(w, nw, p + "ln_2.weight", d, stage, &g2));
    CC_TRY(find_weight(w, nw, p + "ln_2.bias", d, stage, &b2));
    CC_TRY(find_weight(w, nw, p + "attn.c_attn.weight", 3LL * d * d, stage, &wa));
    CC_TRY(find_weight(w, nw, p + "attn.c_attn.bias", 3 * d, stage, &ba));
    CC_TRY(find_weight(w, nw, p + "attn.c_proj.weight", static_cast<int64_t>(d) * d, stage, &wo));
    CC_TRY(find_weight(w, nw, p + "attn.c_proj.bias", d, stage, &bo));
    CC_TRY(find_weight(w, nw, p + "mlp.c_fc.weight", 4LL * d * d, stage, &wf));
    CC_TRY(find_weight(w, nw, p + "mlp.c_fc.bias", 4 * d, stage, &bf));
    CC_TRY(find_weight(w, nw, p + "mlp.c_proj.weight", 4LL * d * d, stage, &wp));
    CC_TRY(find_weight(w, nw, p + "mlp.c_proj.bias", d, stage, &bp));
    CC_TRY(keep_f32(A, g1, d, &L.ln1_g));
    CC_TRY(keep_f32(A, b1, d, &L.ln1_b));
    CC_TRY(keep_f32(A, g2, d, &L.ln2_g));
    CC_TRY(keep_f32(A, b2, d, &L.ln2_b));
    CC_TRY(keep_f32(A, ba, 3 * static_cast<size_t>(d), &L.bqkv));
    CC_TRY(keep_f32(A, bo, d, &L.bo));
    CC_TRY(keep_f32(A, bf, 4 * static_cast<size_t>(d), &L.b1));
    CC_TRY(keep_f32(A, bp, d, &L.b2));
    // Conv1D [in, out]: transposed once for the forward operand, kept as stored for the dgrad operand
    CC_TRY(pack_f16(A, wa, d, 3 * d, true, d, &L.wqkv));
    CC_TRY(pack_f16(A, wa, d, 3 * d, false, 3 * d, &L.wqkv_t));
    CC_TRY(pack_f16(A, wo, d, d, true, d, &L.wo));
    CC_TRY(pack_f16(A, wo, d, d, false, d, &L.wo_t));
    CC_TRY(pack_f16(A, wf, d, 4 * d, true, d, &L.w1));
    CC_TRY(pack_f16(A, wf, d, 4 * d, false, 4 * d, &L.w1_t));
    CC_TRY(pack_f16(A, wp, 4 * d, d, true, 4 * d, &L.w2));
    CC_TRY(pack_f16(A, wp, 4 * d, d, false, d, &L.w2_t));
    CC_TRY(A.alloc_t(&L.hin, rl * d));
    CC_TRY(A.alloc_t(&L.hmid, rl * d));
    CC_TRY(A.alloc_t(&L.qkv, rl * 3 * d));
    CC_TRY(A.alloc_t(&L.att, rl * d));
    CC_TRY(A.alloc_t(&L.pre, rl * 4 * d));
    stage.release();
  }
  t->mp.resize(mc.L);
  const size_t dd = static_cast<size_t>(d) * d;
  for (int l = 0; l < mc.L; ++l) {
    cc_train::MapLayer& M = t->mp[l];
    CC_TRY(A.alloc_t(&M.wqkv, 3 * dd));
    CC_TRY(A.alloc_t(&M.wqkv_t, 3 * dd));
    CC_TRY(A.alloc_t(&M.wo, dd));
    CC_TRY(A.alloc_t(&M.wo_t, dd));
    CC_TRY(A.alloc_t(&M.w1, 2 * dd));
    CC_TRY(A.alloc_t(&M.w1_t, 2 * dd));
    CC_TRY(A.alloc_t(&M.w2, 2 * dd));
    CC_TRY(A.alloc_t(&M.w2_t, 2 * dd));
    CC_TRY(A.alloc_t(&M.hin, rm * d));
    CC_TRY(A.alloc_t(&M.hmid, rm * d));
    CC_TRY(A.alloc_t(&M.qkv, rm * 3 * d));
    CC_TRY(A.alloc_t(&M.att, rm * d));
    CC_TRY(A.alloc_t(&M.hid, rm * 2 * d));
  }
  CC_TRY(A.alloc_t(&t->lin16, static_cast<size_t>(mc.P) * d * mc.E));
  const size_t rmax = rl > rm ? rl : rm;
  const size_t rows_pad = (rmax + 7) / 8 * 8;
  const size_t b_pad = (static_cast<size_t>(B) * t->W + 7) / 8 * 8;  // rows of the input projection: (sample, window)
  CC_TRY(A.alloc_t(&t->emb16, static_cast<size_t>(B) * t->W * mc.E));
  CC_TRY(A.alloc_t(&t->ln16, rmax * d));
  CC_TRY(A.alloc_t(&t->hid16, rmax * 4 * d));
  CC_TRY(A.alloc_t(&t->g16, rmax * d));
  CC_TRY(A.alloc_t(&t->dbig16, rmax * 4 * d));
  CC_TRY(A.alloc_t(&t->dqkv16, rmax * 3 * d));
  CC_TRY(A.alloc_t(&t->datt16, rmax * d));
  CC_TRY(A.alloc_t(&t->hm, rm * d));
  CC_TRY(A.alloc_t(&t->h, rl * d));
  CC_TRY(A.alloc_t(&t->dh, rmax * d));
  CC_TRY(A.alloc_t(&t->dln32, rmax * d));
  const size_t rs = static_cast<size_t>(t->rows_sel);
  CC_TRY(A.alloc_t(&t->hsel, rs * d));
  CC_TRY(A.alloc_t(&t->dhsel, rs * d));
  CC_TRY(A.alloc_t(&t->lnf16, rs * d));
  CC_TRY(A.alloc_t(&t->dlnf32, rs * d));
  CC_TRY(A.alloc_t(&t->logits, static_cast<size_t>(t->head_chunk) * t->v_pad));
  CC_TRY(A.alloc_t(&t->dlogits16, static_cast<size_t>(t->head_chunk) * t->v_pad));
  CC_TRY(A.alloc_t(&t->row_loss, rs));
  CC_TRY(A.alloc_t(&t->loss, 1));
  CC_TRY(A.alloc_t(&t->n_valid, 1));
  CC_TRY(A.alloc_t(&t->tokens, rs));
  CC_TRY(A.alloc_t(&t->targets, rs));
  CC_TRY(A.alloc_t(&t->ln_scratch, ln_bwd_scratch_floats(d)));
  {
    int widest = 2 * d;
    if (mc.K * d > widest) widest = mc.K * d;
    if (t->Ptot * d > widest) widest = t->Ptot * d;
    t->cs_floats = colsum_scratch_floats(widest);
    CC_TRY(A.alloc_t(&t->cs_scratch, t->cs_floats));
  }
  // transposed operands of the weight-gradient GEMMs: dY^T up to [max(3d, P*d), rows] and X^T up to [max(2d, E), rows]
  size_t a_rows = 3 * static_cast<size_t>(d), b_rows = 2 * static_cast<size_t>(d);
  size_t a_elems = a_rows * rows_pad, b_elems = b_rows * rows_pad;
  const size_t lin_a = static_cast<size_t>(mc.P) * d * b_pad, lin_b = static_cast<size_t>(mc.E) * b_pad;
  if (lin_a > a_elems) a_elems = lin_a;
  if (lin_b > b_elems) b_elems = lin_b;
  CC_TRY(A.alloc_t(&t->tr_a, a_elems));
  CC_TRY(A.alloc_t(&t->tr_b, b_elems));
  CC_TRY(A.alloc_t(&t->dlin16, static_cast<size_t>(B) * t->Ptot * d));
  CC_TRY(A.alloc_t(&t->pos_tmp, static_cast<size_t>(t->Ptot) * d));
  CC_TRY(A.alloc_t(&t->nonfinite, 1));
  CC_CUDA(cudaMemset(t->nonfinite, 0, sizeof(int)));
  CC_TRY(A.alloc_t(&t->wqkv_grad, 3 * dd));
  return CC_OK;
}

struct MapParams {  // the caller's fp32 parameter (or gradient) tensors of one mapper layer
  const float *n1_g, *n1_b, *n2_g, *n2_b, *wq, *wkv, *wp, *bp, *f1w, *f1b, *f2w, *f2b;
};

int map_layer_tensors(const cc_tensor* t, int n, int l, int d, MapParams* o) {
  const std::string p = "transformer.layers." + std::to_string(l) + ".";
  const int64_t dd = static_cast<int64_t>(d) * d;
  CC_TRY(dev_f32(t, n, p + "norm1.weight", d, &o->n1_g));
  CC_TRY(dev_f32(t, n, p + "norm1.bias", d, &o->n1_b));
  CC_TRY(dev_f32(t, n, p + "norm2.weight", d, &o->n2_g));
  CC_TRY(dev_f32(t, n, p + "norm2.bias", d, &o->n2_b));
  CC_TRY(dev_f32(t, n, p + "attn.to_queries.weight", dd, &o->wq));
  CC_TRY(dev_f32(t, n, p + "attn.to_keys_values.weight", 2 * dd, &o->wkv));
  CC_TRY(dev_f32(t, n, p + "attn.project.weight", dd, &o->wp));
  CC_TRY(dev_f32(t, n, p + "attn.project.bias", d, &o->bp));
  CC_TRY(dev_f32(t, n, p + "mlp.fc1.weight", 2 * dd, &o->f1w));
  CC_TRY(dev_f32(t, n, p + "mlp.fc1.bias", 2 * d, &o->f1b));
  CC_TRY(dev_f32(t, n, p + "mlp.fc2.weight", 2 * dd, &o->f2w));
  CC_TRY(dev_f32(t, n, p + "mlp.fc2.bias", d, &o->f2b));
  return CC_OK;
}

}  // namespace
}  // namespace cc

extern "C" {

int cc_train_create(cc_train** h, const cc_mapper_cfg* mcfg, const cc_gpt2_cfg* gcfg, const cc_tensor* lm_weights,
                    int n_weights, int max_batch, int max_tokens) {
  using namespace cc;
  CC_REQUIRE(h != nullptr && mcfg != nullptr && gcfg != nullptr && lm_weights != nullptr, CC_EINVAL,
             "cc_train_create: null argument");
  *h = nullptr;
  CC_TRY(check_device_sm100());
  CC_REQUIRE(mcfg->kind == CC_MAPPER_TRANSFORMER || mcfg->kind == CC_MAPPER_WINDOWED, CC_ESHAPE,
             "cc_train_create: the trainable mappers are TransformerMapper and TransformerMapperWindowed (kind %d)",
             mcfg->kind);
  if (mcfg->kind == CC_MAPPER_WINDOWED) CC_REQUIRE(mcfg->W >= 1, CC_ESHAPE, "cc_train_create: windowed mapper W=%d", mcfg->W);
  CC_REQUIRE(max_batch > 0 && max_tokens > 0, CC_EINVAL, "cc_train_create: max_batch=%d max_tokens=%d", max_batch, max_tokens);
  CC_REQUIRE(mcfg->d == gcfg->d, CC_ESHAPE, "cc_train_create: mapper width %d != LM width %d", mcfg->d, gcfg->d);
  CC_REQUIRE(gcfg->d % 8 == 0 && gcfg->d % gcfg->H == 0 && gcfg->d / gcfg->H == 64, CC_ESHAPE,
             "cc_train_create: LM head dim must be 64 (d=%d H=%d)", gcfg->d, gcfg->H);
  CC_REQUIRE(mcfg->E > 0 && mcfg->E % 8 == 0 && mcfg->P > 0 && mcfg->K > 0 && mcfg->L > 0 && mcfg->H > 0 &&
                 mcfg->d % mcfg->H == 0,
             CC_ESHAPE, "cc_train_create: mapper E=%d P=%d K=%d L=%d H=%d", mcfg->E, mcfg->P, mcfg->K, mcfg->L, mcfg->H);
  const int mhd = mcfg->d / mcfg->H;
  CC_REQUIRE(mhd == 48 || mhd == 64 || mhd == 96 || mhd == 128, CC_ESHAPE, "cc_train_create: mapper head dim %d", mhd);
  CC_REQUIRE(mcfg->K + max_tokens <= gcfg->n_pos, CC_ESHAPE, "cc_train_create: %d + %d positions exceed n_positions %d",
             mcfg->K, max_tokens, gcfg->n_pos);
  cc_train* t = new cc_train();
  t->mc = *mcfg;
  t->gc = *gcfg;
  if (t->mc.eps <= 0.f) t->mc.eps = 1e-5f;
  if (t->gc.eps <= 0.f) t->gc.eps = 1e-5f;
  t->max_batch = max_batch;
  t->max_tokens = max_tokens;
  const int st = train_build(t, lm_weights, n_weights);
  if (st != CC_OK) {
    delete t;
    return st;
  }
  *h = t;
  return CC_OK;
}

int cc_train_step(cc_train* t, const cc_tensor* params, int n_params, const cc_tensor* grads, int n_grads,
                  const void* emb, int emb_dtype, const int32_t* tokens, int B, int Tt, float loss_scale, float* loss,
                  void* stream) {
  using namespace cc;
  CC_REQUIRE(t != nullptr && params != nullptr && emb != nullptr && tokens != nullptr && loss != nullptr, CC_EINVAL,
             "cc_train_step: null argument");
  CC_REQUIRE(B > 0 && B <= t->max_batch && Tt > 0 && Tt <= t->max_tokens, CC_ESHAPE,
             "cc_train_step: B=%d Tt=%d outside the handle (max_batch %d, max_tokens %d)", B, Tt, t->max_batch,
             t->max_tokens);
  const bool backward = grads != nullptr && n_grads > 0;
  if (loss_scale <= 0.f) loss_scale = 1024.f;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const cc_mapper_cfg& mc = t->mc;
  const cc_gpt2_cfg& gc = t->gc;
  const int d = gc.d, S = t->S, K = mc.K, P = mc.P, T = K + Tt;
  const int W = t->W, Ptot = t->Ptot;  // windowed mapper: W embeddings per sample -> W * P projected tokens (mapper.py:148-150)
  const bool use_pos = mc.kind == CC_MAPPER_WINDOWED && mc.use_pos != 0;
  const int rows_m = B * S, rows_l = B * T, rows_sel = B * Tt;
  const int mhd = d / mc.H;
  const float mscale = 1.0f / sqrtf(static_cast<float>(mhd)), lscale = 0.125f;
  const float inv_scale = 1.0f / loss_scale;
  const size_t hbytes_m = static_cast<size_t>(rows_m) * d * sizeof(float);
  const size_t hbytes_l = static_cast<size_t>(rows_l) * d * sizeof(float);
  int* nl = &t->launches;
  *nl = 0;

  // ---------------------------------------------------------------- parameters of this step -> fp16 operands
  const float *lin_w, *lin_b, *prefix_const;
  CC_TRY(dev_f32(params, n_params, "linear.weight", static_cast<int64_t>(P) * d * mc.E, &lin_w));
  CC_TRY(dev_f32(params, n_params, "linear.bias", static_cast<int64_t>(P) * d, &lin_b));
  CC_TRY(dev_f32(params, n_params, "prefix_const", static_cast<int64_t>(K) * d, &prefix_const));
  const float* pos_emb = nullptr;
  if (use_pos) CC_TRY(dev_f32(params, n_params, "pos_embeddings", static_cast<int64_t>(Ptot) * d, &pos_emb));
  CC_TRY(pack_weight_run(lin_w, P * d, mc.E, false, t->lin16, mc.E, s));
  std::vector<MapParams> mp(mc.L), mg(mc.L);
  for (int l = 0; l < mc.L; ++l) {
    CC_TRY(map_layer_tensors(params, n_params, l, d, &mp[l]));
    cc_train::MapLayer& M = t->mp[l];
    const size_t dd = static_cast<size_t>(d) * d;
    CC_TRY(pack_weight_run(mp[l].wq, d, d, false, M.wqkv, d, s));
    CC_TRY(pack_weight_run(mp[l].wkv, 2 * d, d, false, M.wqkv + dd, d, s));
    CC_TRY(pack_weight_run(mp[l].wp, d, d, false, M.wo, d, s));
    CC_TRY(pack_weight_run(mp[l].f1w, 2 * d, d, false, M.w1, d, s));
    CC_TRY(pack_weight_run(mp[l].f2w, d, 2 * d, false, M.w2, 2 * d, s));
    *nl += 5;
    if (backward) {
      // dgrad operands: W^T, turned from the fp16 forward operands ([3d, d] -> [d, 3d], ...)
      CC_TRY(transpose16_run(M.wqkv, d, 3 * d, d, M.wqkv_t, 3 * d, s));
      CC_TRY(transpose16_run(M.wo, d, d, d, M.wo_t, d, s));
      CC_TRY(transpose16_run(M.w1, d, 2 * d, d, M.w1_t, 2 * d, s));
      CC_TRY(transpose16_run(M.w2, 2 * d, d, 2 * d, M.w2_t, d, s));
      *nl += 4;
    }
  }

  // ---------------------------------------------------------------- mapper forward (mapper.py:122-130)
  // x = linear(emb).view(B, W * P, d) [+ pos_embeddings]; cat prefix_const (mapper.py:123-126 / 148-156). Window w of
  // sample b lands at rows w*P .. of the sample's block: one GEMM per window (row stride S*d), as in cc_mapper_forward.
  CC_TRY(convert_to_f16_run(emb, emb_dtype, t->emb16, static_cast<int64_t>(B) * W * mc.E, s));
  for (int wdw = 0; wdw < W; ++wdw)
    CC_TRY(gemm(t->emb16 + static_cast<size_t>(wdw) * mc.E, static_cast<int64_t>(W) * mc.E, B, t->lin16, P * d, mc.E, EPI_F32,
                lin_b, t->hm + static_cast<size_t>(wdw) * P * d, static_cast<int64_t>(S) * d, s, nl));
  CC_TRY(mapper_fill_const_run(t->hm, prefix_const, pos_emb, B, Ptot, K, d, s));
  *nl += 2;
  for (int l = 0; l < mc.L; ++l) {
    cc_train::MapLayer& M = t->mp[l];
    const MapParams& w = mp[l];
    CC_CUDA(cudaMemcpyAsync(M.hin, t->hm, hbytes_m, cudaMemcpyDeviceToDevice, s));
    CC_TRY(layernorm_run(t->hm, d, w.n1_g, w.n1_b, t->ln16, d, rows_m, d, mc.eps, s));
    CC_TRY(gemm(t->ln16, d, rows_m, M.wqkv, 3 * d, d, EPI_F16_NONE, nullptr, M.qkv, 3 * d, s, nl));
    CC_TRY(attention_run(M.qkv, M.qkv + d, M.qkv + 2 * d, 3 * d, M.att, d, B, S, mc.H, mhd, false, mscale, s));
    CC_TRY(gemm(M.att, d, rows_m, M.wo, d, d, EPI_RESID_F32, w.bp, t->hm, d, s, nl));
    CC_CUDA(cudaMemcpyAsync(M.hmid, t->hm, hbytes_m, cudaMemcpyDeviceToDevice, s));
    CC_TRY(layernorm_run(t->hm, d, w.n2_g, w.n2_b, t->ln16, d, rows_m, d, mc.eps, s));
    CC_TRY(gemm(t->ln16, d, rows_m, M.w1, 2 * d, d, EPI_F16_RELU, w.f1b, M.hid, 2 * d, s, nl));
    CC_TRY(gemm(M.hid, 2 * d, rows_m, M.w2, d, 2 * d, EPI_RESID_F32, w.f2b, t->hm, d, s, nl));
    *nl += 5;
  }

  // ---------------------------------------------------------------- LM forward over [prefix, tokens] (model.py:45-56)
  CC_CUDA(cudaMemcpyAsync(t->tokens, tokens, sizeof(int32_t) * rows_sel, cudaMemcpyDefault, s));
  CC_TRY(train_embed_run(t->tokens, B, Tt, K, t->hm + static_cast<size_t>(Ptot) * d, static_cast<int64_t>(S) * d, t->wte32,
                         t->wpe32, t->h, t->targets, d, gc.V, s));
  *nl += 1;
  for (int l = 0; l < gc.L; ++l) {
    cc_train::LmLayer& L = t->lm[l];
    if (backward) CC_CUDA(cudaMemcpyAsync(L.hin, t->h, hbytes_l, cudaMemcpyDeviceToDevice, s));
    CC_TRY(layernorm_run(t->h, d, L.ln1_g, L.ln1_b, t->ln16, d, rows_l, d, gc.eps, s));
    CC_TRY(gemm(t->ln16, d, rows_l, L.wqkv, 3 * d, d, EPI_F16_NONE, L.bqkv, L.qkv, 3 * d, s, nl));
    CC_TRY(attention_run(L.qkv, L.qkv + d, L.qkv + 2 * d, 3 * d, L.att, d, B, T, gc.H, 64, true, lscale, s));
    CC_TRY(gemm(L.att, d, rows_l, L.wo, d, d, EPI_RESID_F32, L.bo, t->h, d, s, nl));
    if (backward) CC_CUDA(cudaMemcpyAsync(L.hmid, t->h, hbytes_l, cudaMemcpyDeviceToDevice, s));
    CC_TRY(layernorm_run(t->h, d, L.ln2_g, L.ln2_b, t->ln16, d, rows_l, d, gc.eps, s));
    CC_TRY(gemm(t->ln16, d, rows_l, L.w1, 4 * d, d, EPI_F16_NONE, L.b1, L.pre, 4 * d, s, nl));
    CC_TRY(gelu_new_fwd_run(L.pre, t->hid16, static_cast<int64_t>(rows_l) * 4 * d, s));
    CC_TRY(gemm(t->hid16, 4 * d, rows_l, L.w2, d, 4 * d, EPI_RESID_F32, L.b2, t->h, d, s, nl));
    *nl += 4;
  }

  // ---------------------------------------------------------------- head + loss (model.py:109-110)
  // logits[:, K-1 : T-1] predict tokens[:, 0 : Tt]: gather those rows, ln_f, tied head in row chunks, cross-entropy
  CC_CUDA(cudaMemcpy2DAsync(t->hsel, static_cast<size_t>(Tt) * d * sizeof(float), t->h + static_cast<size_t>(K - 1) * d,
                            static_cast<size_t>(T) * d * sizeof(float), static_cast<size_t>(Tt) * d * sizeof(float), B,
                            cudaMemcpyDeviceToDevice, s));
  CC_TRY(layernorm_run(t->hsel, d, t->lnf_g, t->lnf_b, t->lnf16, d, rows_sel, d, gc.eps, s));
  CC_TRY(count_valid_run(t->targets, rows_sel, t->n_valid, s));
  *nl += 2;
  for (int r0 = 0; r0 < rows_sel; r0 += t->head_chunk) {
    const int m = rows_sel - r0 < t->head_chunk ? rows_sel - r0 : t->head_chunk;
    CC_TRY(gemm(t->lnf16 + static_cast<size_t>(r0) * d, d, m, t->wte16, gc.V, d, EPI_F32, nullptr, t->logits, t->v_pad, s, nl));
    CC_TRY(ce_loss_run(t->logits, t->v_pad, gc.V, t->targets + r0, m, t->n_valid, loss_scale, t->row_loss + r0,
                       t->dlogits16, t->v_pad, s));
    *nl += 1;
    if (backward)  // d ln_f output = dlogits wte
      CC_TRY(gemm(t->dlogits16, t->v_pad, m, t->wte_t16, d, t->v_pad, EPI_F32, nullptr,
                  t->dlnf32 + static_cast<size_t>(r0) * d, d, s, nl));
  }
  CC_TRY(loss_reduce_run(t->row_loss, rows_sel, t->n_valid, t->loss, s));
  CC_CUDA(cudaMemcpyAsync(loss, t->loss, sizeof(float), cudaMemcpyDeviceToDevice, s));
  *nl += 1;
  if (!backward) return CC_OK;

  // ---------------------------------------------------------------- LM backward (activation gradients only)
  CC_CUDA(cudaMemsetAsync(t->dhsel, 0, static_cast<size_t>(rows_sel) * d * sizeof(float), s));
  CC_TRY(layernorm_bwd_run(t->dlnf32, d, t->hsel, d, t->lnf_g, t->dhsel, d, rows_sel, d, gc.eps, nullptr, nullptr, 1.f,
                           nullptr, s));
  CC_CUDA(cudaMemsetAsync(t->dh, 0, hbytes_l, s));
  CC_CUDA(cudaMemcpy2DAsync(t->dh + static_cast<size_t>(K - 1) * d, static_cast<size_t>(T) * d * sizeof(float), t->dhsel,
                            static_cast<size_t>(Tt) * d * sizeof(float), static_cast<size_t>(Tt) * d * sizeof(float), B,
                            cudaMemcpyDeviceToDevice, s));
  *nl += 1;
  const int64_t n_l = static_cast<int64_t>(rows_l) * d;
  for (int l = gc.L - 1; l >= 0; --l) {
    cc_train::LmLayer& L = t->lm[l];
    // MLP branch: h_out = h_mid + c_proj(gelu_new(c_fc(LN2(h_mid))))
    CC_TRY(convert_to_f16_run(t->dh, CC_F32, t->g16, n_l, s));
    CC_TRY(gemm(t->g16, d, rows_l, L.w2_t, 4 * d, d, EPI_F16_NONE, nullptr, t->dbig16, 4 * d, s, nl));
    CC_TRY(gelu_new_bwd_run(t->dbig16, L.pre, static_cast<int64_t>(rows_l) * 4 * d, s));
    CC_TRY(gemm(t->dbig16, 4 * d, rows_l, L.w1_t, d, 4 * d, EPI_F32, nullptr, t->dln32, d, s, nl));
    CC_TRY(layernorm_bwd_run(t->dln32, d, L.hmid, d, L.ln2_g, t->dh, d, rows_l, d, gc.eps, nullptr, nullptr, 1.f, nullptr, s));
    // attention branch: h_mid = h_in + c_proj(attn(c_attn(LN1(h_in))))
    CC_TRY(convert_to_f16_run(t->dh, CC_F32, t->g16, n_l, s));
    CC_TRY(gemm(t->g16, d, rows_l, L.wo_t, d, d, EPI_F16_NONE, nullptr, t->datt16, d, s, nl));
    CC_TRY(attention_bwd_run(L.qkv, L.qkv + d, L.qkv + 2 * d, 3 * d, t->datt16, d, t->dqkv16, t->dqkv16 + d,
                             t->dqkv16 + 2 * d, 3 * d, B, T, gc.H, 64, true, lscale, s));
    CC_TRY(gemm(t->dqkv16, 3 * d, rows_l, L.wqkv_t, d, 3 * d, EPI_F32, nullptr, t->dln32, d, s, nl));
    CC_TRY(layernorm_bwd_run(t->dln32, d, L.hin, d, L.ln1_g, t->dh, d, rows_l, d, gc.eps, nullptr, nullptr, 1.f, nullptr, s));
    *nl += 6;
  }

  // ---------------------------------------------------------------- mapper backward (activation + parameter gradients)
  // d prefix = dh[:, :K]  ->  rows W*P.. of the mapper stream; the projected-token rows start at zero
  float* dhm = t->hm;  // the mapper's stream buffer is free now (its layers are stashed): reuse it for the gradient
  CC_CUDA(cudaMemsetAsync(dhm, 0, hbytes_m, s));
  CC_CUDA(cudaMemcpy2DAsync(dhm + static_cast<size_t>(Ptot) * d, static_cast<size_t>(S) * d * sizeof(float), t->dh,
                            static_cast<size_t>(T) * d * sizeof(float), static_cast<size_t>(K) * d * sizeof(float), B,
                            cudaMemcpyDeviceToDevice, s));
  float* gl_w;
  float* gl_b;
  float* g_pc;
  {
    const float *a, *b2, *c;
    CC_TRY(dev_f32(grads, n_grads, "linear.weight", static_cast<int64_t>(P) * d * mc.E, &a));
    CC_TRY(dev_f32(grads, n_grads, "linear.bias", static_cast<int64_t>(P) * d, &b2));
    CC_TRY(dev_f32(grads, n_grads, "prefix_const", static_cast<int64_t>(K) * d, &c));
    gl_w = const_cast<float*>(a);
    gl_b = const_cast<float*>(b2);
    g_pc = const_cast<float*>(c);
  }
  for (int l = 0; l < mc.L; ++l) CC_TRY(map_layer_tensors(grads, n_grads, l, d, &mg[l]));
  const int64_t n_m = static_cast<int64_t>(rows_m) * d;
  const int mp_rows = (rows_m + 7) / 8 * 8;  // K dimension of the weight-gradient GEMMs
  const int64_t dd = static_cast<int64_t>(d) * d;
  for (int l = mc.L - 1; l >= 0; --l) {
    cc_train::MapLayer& M = t->mp[l];
    const MapParams& w = mp[l];
    const MapParams& g = mg[l];
    // ---- MLP branch
    CC_TRY(convert_to_f16_run(dhm, CC_F32, t->g16, n_m, s));
    CC_TRY(colsum_f32_run(dhm, d, rows_m, d, inv_scale, const_cast<float*>(g.f2b), t->cs_scratch, t->cs_floats, s));
    CC_TRY(transpose16_run(t->g16, d, rows_m, d, t->tr_a, mp_rows, s));
    CC_TRY(transpose16_run(M.hid, 2 * d, rows_m, 2 * d, t->tr_b, mp_rows, s));
    CC_TRY(gemm(t->tr_a, mp_rows, d, t->tr_b, 2 * d, mp_rows, EPI_F32, nullptr, const_cast<float*>(g.f2w), 2 * d, s, nl));
    CC_TRY(scale_f32_run(const_cast<float*>(g.f2w), 2 * dd, inv_scale, s));
    CC_TRY(gemm(t->g16, d, rows_m, M.w2_t, 2 * d, d, EPI_F16_NONE, nullptr, t->dbig16, 2 * d, s, nl));
    CC_TRY(relu_bwd_run(t->dbig16, M.hid, static_cast<int64_t>(rows_m) * 2 * d, s));
    CC_TRY(colsum_f16_run(t->dbig16, 2 * d, rows_m, 2 * d, inv_scale, const_cast<float*>(g.f1b), t->cs_scratch, t->cs_floats, s));
    CC_TRY(layernorm_run(M.hmid, d, w.n2_g, w.n2_b, t->ln16, d, rows_m, d, mc.eps, s));  // LN2 output, recomputed
    CC_TRY(transpose16_run(t->dbig16, 2 * d, rows_m, 2 * d, t->tr_a, mp_rows, s));
    CC_TRY(transpose16_run(t->ln16, d, rows_m, d, t->tr_b, mp_rows, s));
    CC_TRY(gemm(t->tr_a, mp_rows, 2 * d, t->tr_b, d, mp_rows, EPI_F32, nullptr, const_cast<float*>(g.f1w), d, s, nl));
    CC_TRY(scale_f32_run(const_cast<float*>(g.f1w), 2 * dd, inv_scale, s));
    CC_TRY(gemm(t->dbig16, 2 * d, rows_m, M.w1_t, d, 2 * d, EPI_F32, nullptr, t->dln32, d, s, nl));
    CC_TRY(layernorm_bwd_run(t->dln32, d, M.hmid, d, w.n2_g, dhm, d, rows_m, d, mc.eps, const_cast<float*>(g.n2_g),
                             const_cast<float*>(g.n2_b), inv_scale, t->ln_scratch, s));
    // ---- attention branch
    CC_TRY(convert_to_f16_run(dhm, CC_F32, t->g16, n_m, s));
    CC_TRY(colsum_f32_run(dhm, d, rows_m, d, inv_scale, const_cast<float*>(g.bp), t->cs_scratch, t->cs_floats, s));
    CC_TRY(transpose16_run(t->g16, d, rows_m, d, t->tr_a, mp_rows, s));
    CC_TRY(transpose16_run(M.att, d, rows_m, d, t->tr_b, mp_rows, s));
    CC_TRY(gemm(t->tr_a, mp_rows, d, t->tr_b, d, mp_rows, EPI_F32, nullptr, const_cast<float*>(g.wp), d, s, nl));
    CC_TRY(scale_f32_run(const_cast<float*>(g.wp), dd, inv_scale, s));
    CC_TRY(gemm(t->g16, d, rows_m, M.wo_t, d, d, EPI_F16_NONE, nullptr, t->datt16, d, s, nl));
    CC_TRY(attention_bwd_run(M.qkv, M.qkv + d, M.qkv + 2 * d, 3 * d, t->datt16, d, t->dqkv16, t->dqkv16 + d,
                             t->dqkv16 + 2 * d, 3 * d, B, S, mc.H, mhd, false, mscale, s));
    CC_TRY(layernorm_run(M.hin, d, w.n1_g, w.n1_b, t->ln16, d, rows_m, d, mc.eps, s));  // LN1 output, recomputed
    CC_TRY(transpose16_run(t->dqkv16, 3 * d, rows_m, 3 * d, t->tr_a, mp_rows, s));
    CC_TRY(transpose16_run(t->ln16, d, rows_m, d, t->tr_b, mp_rows, s));
    // [3d, d] gradient of the fused QKV weight: rows 0..d-1 -> to_queries, rows d..3d-1 -> to_keys_values
    CC_TRY(gemm(t->tr_a, mp_rows, 3 * d, t->tr_b, d, mp_rows, EPI_F32, nullptr, t->wqkv_grad, d, s, nl));
    CC_TRY(scale_f32_run(t->wqkv_grad, 3 * dd, inv_scale, s));
    CC_CUDA(cudaMemcpyAsync(const_cast<float*>(g.wq), t->wqkv_grad, sizeof(float) * dd, cudaMemcpyDeviceToDevice, s));
    CC_CUDA(cudaMemcpyAsync(const_cast<float*>(g.wkv), t->wqkv_grad + dd, sizeof(float) * 2 * dd, cudaMemcpyDeviceToDevice, s));
    CC_TRY(gemm(t->dqkv16, 3 * d, rows_m, M.wqkv_t, d, 3 * d, EPI_F32, nullptr, t->dln32, d, s, nl));
    CC_TRY(layernorm_bwd_run(t->dln32, d, M.hin, d, w.n1_g, dhm, d, rows_m, d, mc.eps, const_cast<float*>(g.n1_g),
                             const_cast<float*>(g.n1_b), inv_scale, t->ln_scratch, s));
    *nl += 26;
  }
  // ---- inputs of the transformer: x = cat(linear(emb).view(B, W*P, d) [+ pos_embeddings], prefix_const)
  //      (mapper.py:123-126 / 148-156). The W*P*d projected-token gradients of a sample are contiguous, so [B, W*P*d]
  //      (row stride S*d) read as [(B*W), P*d] is the gradient of the linear layer's output in (sample, window) order —
  //      the order of the embeddings.
  CC_TRY(colsum_f32_run(dhm + static_cast<size_t>(Ptot) * d, static_cast<int64_t>(S) * d, B, K * d, inv_scale, g_pc, t->cs_scratch, t->cs_floats, s));
  float* g_pos = t->pos_tmp;  // sum over the batch per (window, token, channel): IS the pos_embeddings gradient
  if (use_pos) {
    const float* gp;
    CC_TRY(dev_f32(grads, n_grads, "pos_embeddings", static_cast<int64_t>(Ptot) * d, &gp));
    g_pos = const_cast<float*>(gp);
  }
  CC_TRY(colsum_f32_run(dhm, static_cast<int64_t>(S) * d, B, Ptot * d, inv_scale, g_pos, t->cs_scratch, t->cs_floats, s));
  if (W == 1) {
    if (g_pos != gl_b) CC_CUDA(cudaMemcpyAsync(gl_b, g_pos, sizeof(float) * P * d, cudaMemcpyDeviceToDevice, s));
  } else {  // the bias is shared by the windows: fold them
    CC_TRY(colsum_f32_run(g_pos, static_cast<int64_t>(P) * d, W, P * d, 1.f, gl_b, t->cs_scratch, t->cs_floats, s));
  }
  CC_TRY(convert_from_f32_run(dhm, static_cast<int64_t>(S) * d, t->dlin16, CC_F16, B, Ptot * d, s));
  const int bw = B * W;
  const int bp_rows = (bw + 7) / 8 * 8;
  CC_TRY(transpose16_run(t->dlin16, static_cast<int64_t>(P) * d, bw, P * d, t->tr_a, bp_rows, s));
  CC_TRY(transpose16_run(t->emb16, mc.E, bw, mc.E, t->tr_b, bp_rows, s));
  CC_TRY(gemm(t->tr_a, bp_rows, P * d, t->tr_b, mc.E, bp_rows, EPI_F32, nullptr, gl_w, mc.E, s, nl));
  CC_TRY(scale_f32_run(gl_w, static_cast<int64_t>(P) * d * mc.E, inv_scale, s));
  *nl += 9;
  // overflow check: every gradient tensor handed back is scanned once (75 M floats at the benchmark mapper: ~0.1 ms)
  CC_CUDA(cudaMemsetAsync(t->nonfinite, 0, sizeof(int), s));
  for (int i = 0; i < n_grads; ++i) {
    int64_t numel = 1;
    for (int k = 0; k < grads[i].ndim; ++k) numel *= grads[i].shape[k];
    CC_TRY(count_nonfinite_run(static_cast<const float*>(grads[i].data), numel, t->nonfinite, s));
  }
  *nl += n_grads;
  return CC_OK;
}

int cc_train_last_nonfinite(cc_train* t, void* stream) {
  if (t == nullptr) return -1;
  int host = 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (cudaMemcpyAsync(&host, t->nonfinite, sizeof(int), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
      cudaStreamSynchronize(s) != cudaSuccess)
    return -1;
  return host;
}

int cc_train_last_launches(cc_train* t) { return t ? t->launches : 0; }

void cc_train_destroy(cc_train* t) { delete t; }

int cc_op_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                float weight_decay, int step, void* stream) {
  using namespace cc;
  CC_REQUIRE(p != nullptr && g != nullptr && m != nullptr && v != nullptr, CC_EINVAL, "cc_op_adamw: null argument");
  CC_TRY(check_device_sm100());
  return adamw_run(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
