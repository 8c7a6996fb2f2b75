// Pre-LN transformer block stack shared by the three dense stages (see common.h) + create-time weight helpers.
#include "common.h"

namespace cc {

int keep_f32(Arena& arena, const float* src_dev, size_t n, const float** out) {
  float* p = nullptr;
  CC_TRY(arena.alloc_t(&p, n));
  CC_CUDA(cudaMemcpy(p, src_dev, n * sizeof(float), cudaMemcpyDeviceToDevice));
  *out = p;
  return CC_OK;
}

int pack_f16(Arena& arena, const float* src_dev, int rows, int cols, bool transpose, int64_t ld_out, const __half** out) {
  const int out_rows = transpose ? cols : rows;
  __half* p = nullptr;
  CC_TRY(arena.alloc_t(&p, static_cast<size_t>(out_rows) * ld_out));
  CC_TRY(pack_weight_run(src_dev, rows, cols, transpose, p, ld_out, nullptr));
  CC_CUDA(cudaStreamSynchronize(nullptr));
  *out = p;
  return CC_OK;
}

int Stack::init(Arena& arena, int d_, int dff_, int H_, int act_epi_, bool causal_, float eps_, int max_rows_,
                int dec_rows_) {
  CC_REQUIRE(d_ > 0 && H_ > 0 && d_ % H_ == 0, CC_ESHAPE, "stack: width %d not divisible by %d heads", d_, H_);
  CC_REQUIRE(d_ % 8 == 0 && dff_ % 8 == 0, CC_ESHAPE, "stack: widths must be multiples of 8 (d=%d dff=%d)", d_, dff_);
  d = d_;
  dff = dff_;
  H = H_;
  hd = d_ / H_;
  CC_REQUIRE(hd == 48 || hd == 64 || hd == 96 || hd == 128, CC_ESHAPE,
             "stack: head dim %d not supported (48, 64, 96, 128)", hd);
  act_epi = act_epi_;
  causal = causal_;
  eps = eps_;
  scale = 1.0f / sqrtf(static_cast<float>(hd));
  max_rows = max_rows_;
  const size_t r = static_cast<size_t>(max_rows);
  CC_TRY(arena.alloc_t(&h, r * d));
  CC_TRY(arena.alloc_t(&ln16, r * d));
  CC_TRY(arena.alloc_t(&qkv16, r * 3 * d));
  CC_TRY(arena.alloc_t(&att16, r * d));
  CC_TRY(arena.alloc_t(&mlp16, r * dff));
  arena_ = &arena;
  dec_rows = dec_rows_;
  dec_rows_pad = (dec_rows_ + 127) / 128 * 128;
  return CC_OK;
}

int Stack::plan() {
  const size_t L = layers.size();
  p_qkv.resize(L);
  p_o.resize(L);
  p_1.resize(L);
  p_2.resize(L);
  for (size_t l = 0; l < L; ++l) {
    const LayerW& w = layers[l];
    if (heads_S > 0)
      CC_TRY(gemm_plan_heads(&p_qkv[l], ln16, d, max_rows / heads_S, heads_S, H, w.wqkv, d, w.bqkv, qkv16));
    else
      CC_TRY(gemm_plan(&p_qkv[l], ln16, d, max_rows, w.wqkv, 3 * d, d, EPI_F16_NONE, w.bqkv, qkv16, 3 * d));
    CC_TRY(gemm_plan(&p_o[l], att16, d, max_rows, w.wo, d, d, EPI_RESID_F32, w.bo, h, d));
    CC_TRY(gemm_plan(&p_1[l], ln16, d, max_rows, w.w1, dff, d, act_epi, w.b1, mlp16, dff));
    CC_TRY(gemm_plan(&p_2[l], mlp16, dff, max_rows, w.w2, d, dff, EPI_RESID_F32, w.b2, h, d));
  }
  if (cls_last_S > 0 && L > 0 && hd == 64 && !causal) {
    const LayerW& w = layers[L - 1];
    const int nb = max_rows / cls_last_S;
    const int64_t ld = static_cast<int64_t>(cls_last_S) * d;  // class-token rows of h are cls_last_S rows apart
    CC_TRY(gemm_plan(&p_o_cls, att16, d, nb, w.wo, d, d, EPI_RESID_F32, w.bo, h, ld));
    CC_TRY(gemm_plan(&p_2_cls, mlp16, dff, nb, w.w2, d, dff, EPI_RESID_F32, w.b2, h, ld));
  } else {
    cls_last_S = 0;
  }
  if (dec_rows > 0) {
    // `part` holds the split-K partial sums of one projection: sized for the largest split factor the planner may choose
    // (16) so that plans made later for another row-block count or SM budget always fit. Allocated once per stack.
    const int max_blocks = dec_rows_pad / 128;
    if (part == nullptr) CC_TRY(arena_->alloc_t(&part, static_cast<size_t>(kMaxSplitK) * dec_rows_pad * d));
    dec_plans.clear();
    DecPlans* unused = nullptr;
    CC_TRY(dec_plans_for(max_blocks, &unused));  // the capacity plans exist from the start; smaller steps add theirs
  }
  return CC_OK;
}

int Stack::dec_plans_for(int blocks, DecPlans** out) {
  CC_REQUIRE(blocks >= 1 && blocks <= dec_rows_pad / 128, CC_ESHAPE, "stack: %d decode row blocks (max %d)", blocks,
             dec_rows_pad / 128);
  DecPlans& dp = dec_plans[std::make_pair(blocks, num_sms())];  // tile shapes / split factors depend on the SM budget
  if (dp.o.empty()) {
    const size_t L = layers.size();
    const int rows = blocks * 128;
    int bn_o, sp_o, bn_2, sp_2;
    gemm_pick_split(rows, d, d, &bn_o, &sp_o);
    gemm_pick_split(rows, d, dff, &bn_2, &sp_2);
    CC_REQUIRE(sp_o <= kMaxSplitK && sp_2 <= kMaxSplitK, CC_EINVAL, "stack: split-K factor %d / %d exceeds %d", sp_o, sp_2,
               kMaxSplitK);
    std::vector<GemmPlan> po(L), p2(L);
    const int plan_rows = rows < dec_rows ? rows : dec_rows;
    for (size_t l = 0; l < L; ++l) {
      const LayerW& w = layers[l];
      CC_TRY(gemm_plan_partial(&po[l], att16, d, plan_rows, w.wo, d, d, part, rows, sp_o, bn_o));
      CC_TRY(gemm_plan_partial(&p2[l], mlp16, dff, plan_rows, w.w2, d, dff, part, rows, sp_2, bn_2));
    }
    dp.o.swap(po);
    dp.p2.swap(p2);
  }
  *out = &dp;
  return CC_OK;
}

bool Stack::skinny_step(int nseq, int row0) const {
  return nseq <= kSkinnyMaxRows && row0 == 0 && pend_splits == 0 && act_epi == EPI_F16_GELU_NEW && d % 32 == 0 &&
         d <= 2048 && dff % 32 == 0 && skinny_enabled();
}

int Stack::ln_decode(const float* g, const float* b, __half* y, int nseq, cudaStream_t s, int row0) {
  const int sp = pend_splits;
  const float* bias = pend_bias;
  const int64_t stride = pend_stride;
  pend_splits = 0;
  pend_bias = nullptr;
  launches += 1;
  const size_t off = static_cast<size_t>(row0) * d;
  return layernorm_reduce_run(h + off, d, sp > 0 ? part + off : nullptr, sp, stride, bias, g,
                              b, y + off, d, nseq, d, eps, s);
}

int Stack::layer_full(int l, int B, int S, KvCache* kv, int slot_stride, cudaStream_t s) {
  const int rows = B * S;
  CC_REQUIRE(rows <= max_rows, CC_ESHAPE, "stack: %d x %d rows exceed the %d the handle was created for", B, S, max_rows);
  const LayerW& w = layers[l];
  CC_TRY(layernorm_run(h, d, w.ln1_g, w.ln1_b, ln16, d, rows, d, eps, s));
  CC_TRY(gemm_run(p_qkv[l], rows, s));
  bool kv_done = false;
  if (heads_S > 0) {
    CC_REQUIRE(S == heads_S, CC_ESHAPE, "stack: head-major QKV planned for %d tokens per image, got %d", heads_S, S);
    CC_TRY(vit_attention_heads_run(qkv16, att16, d, B, S, H, scale, s));
  } else if (kv != nullptr && hd == 64 && small_attention_fits(S, hd, 3 * d, d, qkv16, qkv16 + d, qkv16 + 2 * d, att16)) {
    // prefill: the short-sequence kernel holds every K / V tile in the cache's own row layout and stores it there itself
    CC_TRY(small_attention_run(qkv16, qkv16 + d, qkv16 + 2 * d, 3 * d, att16, d, B, S, H, hd, causal, scale, s,
                               kv->k + l * kv->layer_elems, kv->v + l * kv->layer_elems, kv->t_max, slot_stride));
    kv_done = true;
  } else {
    CC_TRY(attention_run(qkv16, qkv16 + d, qkv16 + 2 * d, 3 * d, att16, d, B, S, H, hd, causal, scale, s));
  }
  launches += 3;
  if (kv != nullptr && !kv_done) {
    CC_TRY(kv_scatter_run(qkv16, kv->k + l * kv->layer_elems, kv->v + l * kv->layer_elems, B, S, H, kv->t_max, 0,
                          slot_stride, s));
    launches += 1;
  }
  CC_TRY(gemm_run(p_o[l], rows, s));
  CC_TRY(layernorm_run(h, d, w.ln2_g, w.ln2_b, ln16, d, rows, d, eps, s));
  CC_TRY(gemm_run(p_1[l], rows, s));
  CC_TRY(gemm_run(p_2[l], rows, s));
  launches += 4;
  return CC_OK;
}

int Stack::layer_cls_only(int l, int B, int S, cudaStream_t s) {
  CC_REQUIRE(cls_last_S == S && l == static_cast<int>(layers.size()) - 1, CC_EINVAL,
             "stack: class-token-only pass is planned for the last layer with %d tokens per sequence", cls_last_S);
  const int rows = B * S;
  CC_REQUIRE(rows <= max_rows, CC_ESHAPE, "stack: %d x %d rows exceed the %d the handle was created for", B, S, max_rows);
  const LayerW& w = layers[l];
  CC_TRY(layernorm_run(h, d, w.ln1_g, w.ln1_b, ln16, d, rows, d, eps, s));
  CC_TRY(gemm_run(p_qkv[l], rows, s));  // K and V of every token (and Q, of which only the class rows are read)
  if (heads_S > 0) {  // head-major planes [(b*3 + which)*H + h][S][64]
    const long long plane = static_cast<long long>(S) * 64;
    CC_TRY(cls_attention_run(qkv16, 3LL * H * plane, static_cast<long long>(H) * plane, plane, 64, att16, d, B, S, H, scale, s));
  } else {  // packed rows [b*S + t][3d]
    CC_TRY(cls_attention_run(qkv16, static_cast<long long>(S) * 3 * d, d, 64, 3LL * d, att16, d, B, S, H, scale, s));
  }
  CC_TRY(gemm_run(p_o_cls, B, s));  // h[b*S] += att_cls Wo^T + bo
  CC_TRY(layernorm_run(h, static_cast<int64_t>(S) * d, w.ln2_g, w.ln2_b, ln16, d, B, d, eps, s));
  CC_TRY(gemm_run(p_1[l], B, s));
  CC_TRY(gemm_run(p_2_cls, B, s));  // h[b*S] += act(...) W2^T + b2
  launches += 7;
  return CC_OK;
}

int Stack::layer_last_row(int l, int B, int S, KvCache* kv, int slot_stride, cudaStream_t s) {
  const int rows = B * S;
  CC_REQUIRE(rows <= max_rows, CC_ESHAPE, "stack: %d x %d rows exceed the %d the handle was created for", B, S, max_rows);
  CC_REQUIRE(hd == 64 && causal && heads_S == 0, CC_EINVAL, "stack: last-row pass needs the causal packed-QKV stack, head dim 64");
  const LayerW& w = layers[l];
  CC_TRY(layernorm_run(h, d, w.ln1_g, w.ln1_b, ln16, d, rows, d, eps, s));
  CC_TRY(gemm_run(p_qkv[l], rows, s));
  if (kv != nullptr)
    CC_TRY(kv_scatter_run(qkv16, kv->k + l * kv->layer_elems, kv->v + l * kv->layer_elems, B, S, H, kv->t_max, 0,
                          slot_stride, s));
  // query = last position: every key is visible, so the causal mask is a no-op
  CC_TRY(cls_attention_run(qkv16, static_cast<long long>(S) * 3 * d, d, 64, 3LL * d, att16, d, B, S, H, scale, s, S - 1));
  float* hl = h + static_cast<size_t>(S - 1) * d;  // row S-1 of every sequence, S rows apart
  const int64_t ld = static_cast<int64_t>(S) * d;
  // The two residual GEMMs write rows S apart: their tensor maps depend on S and are encoded once per (layer, S)
  LastRowPlans& lp = last_row_plans[std::make_pair(l, S)];
  if (lp.rows == 0) {
    const int nb = max_rows / S;
    CC_TRY(gemm_plan(&lp.o, att16, d, nb, w.wo, d, d, EPI_RESID_F32, w.bo, hl, ld));
    CC_TRY(gemm_plan(&lp.p2, mlp16, dff, nb, w.w2, d, dff, EPI_RESID_F32, w.b2, hl, ld));
    lp.rows = nb;
  }
  CC_TRY(gemm_run(lp.o, B, s));
  CC_TRY(layernorm_run(hl, ld, w.ln2_g, w.ln2_b, ln16, d, B, d, eps, s));
  CC_TRY(gemm_run(p_1[l], B, s));
  CC_TRY(gemm_run(lp.p2, B, s));
  launches += kv != nullptr ? 8 : 7;
  return CC_OK;
}

int Stack::layer_decode(int l, int nseq, KvCache* kv, const int32_t* anc, int pos, cudaStream_t s, int row0, int beam,
                        int shared_len) {
  CC_REQUIRE(row0 + nseq <= max_rows && row0 + nseq <= kv->slots, CC_ESHAPE, "stack: %d sequences exceed handle capacity",
             row0 + nseq);
  CC_REQUIRE(row0 == 0 || anc == nullptr, CC_EINVAL, "stack: row groups need the ancestry-free (greedy) cache layout");
  CC_REQUIRE(hd == 64, CC_ESHAPE, "decode attention needs head dim 64 (got %d)", hd);
  CC_REQUIRE(row0 + nseq <= dec_rows, CC_ESHAPE, "stack: %d sequences exceed the %d decode rows planned", row0 + nseq,
             dec_rows);
  const LayerW& w = layers[l];
  if (skinny_step(nseq, row0)) {
    // <= 16 rows: five weight-streaming kernels per layer, LayerNorms applied while the rows are staged (skinny.cu)
    const size_t koff = l * kv->layer_elems;
    CC_TRY(skinny_gemm_run(h, w.ln1_g, w.ln1_b, eps, nullptr, d, nseq, w.wqkv, 3 * d, d, EPI_F16_NONE, w.bqkv, qkv16, 3 * d, s));
    CC_TRY(decode_attention_run(qkv16, kv->k + koff, kv->v + koff, anc, att16, nseq, H, kv->t_max, pos, scale, s, beam,
                                shared_len));
    CC_TRY(skinny_gemm_run(nullptr, nullptr, nullptr, eps, att16, d, nseq, w.wo, d, d, EPI_RESID_F32, w.bo, h, d, s));
    CC_TRY(skinny_gemm_run(h, w.ln2_g, w.ln2_b, eps, nullptr, d, nseq, w.w1, dff, d, EPI_F16_GELU_NEW, w.b1, mlp16, dff, s));
    CC_TRY(skinny_gemm_run(nullptr, nullptr, nullptr, eps, mlp16, dff, nseq, w.w2, d, dff, EPI_RESID_F32, w.b2, h, d, s));
    launches += 5;
    return CC_OK;
  }
  DecPlans* dp = nullptr;  // plans for this step's row-block count (row groups use the capacity plans: they share `part`)
  CC_TRY(dec_plans_for(row0 == 0 ? (nseq + 127) / 128 : dec_rows_pad / 128, &dp));
  const size_t cache_off = l * kv->layer_elems + static_cast<size_t>(row0) * H * kv->t_max * 64;  // slot == row
  CC_TRY(ln_decode(w.ln1_g, w.ln1_b, ln16, nseq, s, row0));  // absorbs the previous layer's fc2 partial sums
  CC_TRY(gemm_run(p_qkv[l], nseq, s, row0));
  CC_TRY(decode_attention_run(qkv16 + static_cast<size_t>(row0) * 3 * d, kv->k + cache_off, kv->v + cache_off, anc,
                              att16 + static_cast<size_t>(row0) * d, nseq, H, kv->t_max, pos, scale, s, beam, shared_len));
  CC_TRY(gemm_run(dp->o[l], nseq, s, row0));
  pend_splits = dp->o[l].splits;
  pend_stride = static_cast<int64_t>(dp->o[l].split_rows) * d;
  pend_bias = w.bo;
  CC_TRY(ln_decode(w.ln2_g, w.ln2_b, ln16, nseq, s, row0));
  CC_TRY(gemm_run(p_1[l], nseq, s, row0));
  CC_TRY(gemm_run(dp->p2[l], nseq, s, row0));
  pend_splits = dp->p2[l].splits;
  pend_stride = static_cast<int64_t>(dp->p2[l].split_rows) * d;
  pend_bias = w.b2;
  launches += 5;
  return CC_OK;
}

}  // namespace cc
