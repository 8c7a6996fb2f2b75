// Stage 3 — GPT-2 engine and the decode loops.
// Replaces model.language_model(inputs_embeds=...) / get_input_embeddings() as used by clipcap/inference/base.py:76,81,117
// (HF GPT2LMHeadModel arithmetic) and generate_beam (base.py:55-132).
//
// The reference re-runs the whole LM over the growing sequence each step (no KV cache, base.py:81,117-118). Here the
// prefix is prefilled once into an fp16 KV cache [layer][slot][head][t_max][64] and every further token is a one-row
// decode pass; under causal attention the two are the same function. A whole cc_generate call (prefill + all decode
// steps + token selection) is captured into one CUDA graph per (mode, B, Tp, beam, entry_length) and replayed.
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>

#include "common.h"
#include "decode.h"

struct cc_gpt2 {
  cc_gpt2_cfg cfg;
  int max_seqs = 0, max_len = 0, t_max = 0, max_rows = 0;
  int v_ld = 0;  // row stride of the internal logits buffer (V rounded up to 8)
  cc::Arena arena;
  cc::Stack st;
  cc::KvCache kv;
  const float *wte32 = nullptr, *wpe32 = nullptr, *lnf_g = nullptr, *lnf_b = nullptr;
  const __half* wte16 = nullptr;  // [V, d]: already the [N, K] operand of the tied LM head
  __half* lnf16 = nullptr;        // [max_rows, d]
  float* logits = nullptr;        // [max_seqs, v_ld]  (beam mode)
  unsigned long long* keys = nullptr;
  int32_t *g_tokens = nullptr, *g_lengths = nullptr, *g_stopped = nullptr;  // greedy state / final outputs
  float* g_scores = nullptr;
  unsigned long long* d_seed = nullptr;  // sampling modes: Philox seed of the current call
  int32_t* d_history = nullptr;          // sampling modes: text-prefix tokens (device copy), kMaxHistory entries
  float* cand_val = nullptr;
  int32_t* cand_idx = nullptr;
  cc::BeamState beam{};
  int max_entry = 0;
  cc::GemmPlan p_head_keys, p_head_logits;
  cc::GemmPlan p_logits_out;           // cc_gpt2_logits: head writing into the caller's buffer `logits_plan_out`
  const float* logits_plan_out = nullptr;
  cudaStream_t cap_stream = nullptr;
  // Greedy decode can run as independent row groups on parallel streams (parallel branches of the captured graph).
  // Measured on B200 at B=256 it loses (1 group 31.9 ms, 2 groups 36.2 ms, 4 groups 45.5 ms per generate call: the
  // branches' GEMM CTAs compete for whole SMs and every group re-streams the weights), so the default is one group;
  // CLIPCAP_B200_DECODE_GROUPS=n keeps the experiment reproducible.
  static constexpr int kMaxGroups = 4;
  int decode_groups = 1;
  cudaStream_t grp_stream[kMaxGroups] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[kMaxGroups] = {nullptr, nullptr, nullptr, nullptr};
  // (mode, B, Tp, beam, entry_length, stop_token, temperature bits, SM budget, capture stream, phase)
  using Key = std::tuple<int, int, int, int, int, int, uint32_t, int, uintptr_t, int, int>;
  struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    unsigned long long last_use = 0;
    int launches = 0;  // kernels inside the graph
  };
  std::map<Key, GraphEntry> graphs;  // bounded: least recently used entry evicted beyond kMaxGraphs
  static constexpr size_t kMaxGraphs = 16;
  unsigned long long use_clock = 0;
  bool use_graphs = true;
  int launches = 0;
  int phase_launches = 0;  // kernels of the last prefill phase (cc_generate_prefill), added to the decode phase's count
  int prefill_defer = 0;   // two-phase generate: trailing prefill blocks that run at the head of the decode phase

  ~cc_gpt2() {
    for (auto& kv_ : graphs) cudaGraphExecDestroy(kv_.second.exec);
    if (cap_stream) cudaStreamDestroy(cap_stream);
    for (int i = 0; i < kMaxGroups; ++i) {
      if (grp_stream[i]) cudaStreamDestroy(grp_stream[i]);
      if (ev_join[i]) cudaEventDestroy(ev_join[i]);
    }
    if (ev_fork) cudaEventDestroy(ev_fork);
  }
};

namespace cc {
namespace {

constexpr int kMaxEntry = 128;   // generated tokens per call the state buffers are sized for
constexpr int kMaxHistory = 512;  // text-prefix tokens a sampling call may carry

int gpt2_build(cc_gpt2* m, const cc_tensor* w, int nw) {
  const cc_gpt2_cfg& c = m->cfg;
  Arena stage;
  const int d = c.d;
  m->t_max = (m->max_len + 7) / 8 * 8;
  m->max_rows = m->max_seqs * m->max_len;
  m->v_ld = (c.V + 7) / 8 * 8;
  m->max_entry = kMaxEntry;
  CC_TRY(m->st.init(m->arena, d, 4 * d, c.H, EPI_F16_GELU_NEW, true, c.eps, m->max_rows, m->max_seqs));
  CC_REQUIRE(m->st.hd == 64, CC_ESHAPE, "gpt2: head dim %d (n_embd %d / n_head %d) must be 64", m->st.hd, d, c.H);

  const float *wte, *wpe, *g, *b;
  CC_TRY(find_weight(w, nw, "transformer.wte.weight", static_cast<int64_t>(c.V) * d, stage, &wte));
  CC_TRY(find_weight(w, nw, "transformer.wpe.weight", static_cast<int64_t>(c.n_pos) * d, stage, &wpe));
  CC_TRY(find_weight(w, nw, "transformer.ln_f.weight", d, stage, &g));
  CC_TRY(find_weight(w, nw, "transformer.ln_f.bias", d, stage, &b));
  CC_TRY(keep_f32(m->arena, wte, static_cast<size_t>(c.V) * d, &m->wte32));
  CC_TRY(keep_f32(m->arena, wpe, static_cast<size_t>(c.n_pos) * d, &m->wpe32));
  CC_TRY(keep_f32(m->arena, g, d, &m->lnf_g));
  CC_TRY(keep_f32(m->arena, b, d, &m->lnf_b));
  CC_TRY(pack_f16(m->arena, wte, c.V, d, false, d, &m->wte16));
  stage.release();

  m->st.layers.resize(c.L);
  for (int l = 0; l < c.L; ++l) {
    const std::string p = "transformer.h." + std::to_string(l) + ".";
    LayerW& L = m->st.layers[l];
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
    CC_TRY(keep_f32(m->arena, g1, d, &L.ln1_g));
    CC_TRY(keep_f32(m->arena, b1, d, &L.ln1_b));
    CC_TRY(keep_f32(m->arena, g2, d, &L.ln2_g));
    CC_TRY(keep_f32(m->arena, b2, d, &L.ln2_b));
    // HF Conv1D keeps [in, out]; the GEMM wants [out, in] (K-major), so transpose once here.
    CC_TRY(pack_f16(m->arena, wa, d, 3 * d, true, d, &L.wqkv));
    CC_TRY(keep_f32(m->arena, ba, 3 * static_cast<size_t>(d), &L.bqkv));
    CC_TRY(pack_f16(m->arena, wo, d, d, true, d, &L.wo));
    CC_TRY(keep_f32(m->arena, bo, d, &L.bo));
    CC_TRY(pack_f16(m->arena, wf, d, 4 * d, true, d, &L.w1));
    CC_TRY(keep_f32(m->arena, bf, 4 * static_cast<size_t>(d), &L.b1));
    CC_TRY(pack_f16(m->arena, wp, 4 * d, d, true, 4 * d, &L.w2));
    CC_TRY(keep_f32(m->arena, bp, d, &L.b2));
    stage.release();
  }
  CC_TRY(m->st.plan());

  // KV cache
  m->kv.slots = m->max_seqs;
  m->kv.t_max = m->t_max;
  m->kv.layer_elems = static_cast<size_t>(m->max_seqs) * c.H * m->t_max * 64;
  CC_TRY(m->arena.alloc_t(&m->kv.k, m->kv.layer_elems * c.L));
  CC_TRY(m->arena.alloc_t(&m->kv.v, m->kv.layer_elems * c.L));
  CC_CUDA(cudaMemset(m->kv.k, 0, m->kv.layer_elems * c.L * sizeof(__half)));
  CC_CUDA(cudaMemset(m->kv.v, 0, m->kv.layer_elems * c.L * sizeof(__half)));

  const size_t ns = static_cast<size_t>(m->max_seqs);
  CC_TRY(m->arena.alloc_t(&m->lnf16, static_cast<size_t>(m->max_rows) * d));
  CC_TRY(m->arena.alloc_t(&m->logits, ns * m->v_ld));
  CC_TRY(m->arena.alloc_t(&m->keys, ns));
  CC_TRY(m->arena.alloc_t(&m->g_tokens, ns * m->max_entry));
  CC_TRY(m->arena.alloc_t(&m->g_lengths, ns));
  CC_TRY(m->arena.alloc_t(&m->g_stopped, ns));
  CC_TRY(m->arena.alloc_t(&m->g_scores, ns));
  CC_TRY(m->arena.alloc_t(&m->d_seed, 1));
  CC_TRY(m->arena.alloc_t(&m->d_history, kMaxHistory));
  CC_TRY(m->arena.alloc_t(&m->cand_val, ns * kMaxBeam));
  CC_TRY(m->arena.alloc_t(&m->cand_idx, ns * kMaxBeam));
  CC_TRY(m->arena.alloc_t(&m->beam.scores, ns));
  CC_TRY(m->arena.alloc_t(&m->beam.seq_len, ns));
  CC_TRY(m->arena.alloc_t(&m->beam.stopped, ns));
  for (int i = 0; i < 2; ++i) {
    CC_TRY(m->arena.alloc_t(&m->beam.tokens[i], ns * m->max_entry));
    CC_TRY(m->arena.alloc_t(&m->beam.anc[i], ns * m->t_max));
    CC_CUDA(cudaMemset(m->beam.tokens[i], 0, ns * m->max_entry * sizeof(int32_t)));
    CC_CUDA(cudaMemset(m->beam.anc[i], 0, ns * m->t_max * sizeof(int32_t)));
  }
  CC_CUDA(cudaMemset(m->keys, 0, ns * sizeof(unsigned long long)));
  CC_CUDA(cudaMemset(m->g_scores, 0, ns * sizeof(float)));
  CC_TRY(gemm_plan(&m->p_head_keys, m->lnf16, d, m->max_seqs, m->wte16, c.V, d, EPI_ARGMAX, nullptr, m->keys, 1));
  CC_TRY(gemm_plan(&m->p_head_logits, m->lnf16, d, m->max_seqs, m->wte16, c.V, d, EPI_F32, nullptr, m->logits, m->v_ld));
  CC_CUDA(cudaStreamCreateWithFlags(&m->cap_stream, cudaStreamNonBlocking));
  {
    const char* e = getenv("CLIPCAP_B200_DECODE_GROUPS");
    if (e != nullptr) m->decode_groups = atoi(e);
    if (m->decode_groups < 1) m->decode_groups = 1;
    if (m->decode_groups > cc_gpt2::kMaxGroups) m->decode_groups = cc_gpt2::kMaxGroups;
    CC_CUDA(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
    for (int i = 1; i < m->decode_groups; ++i) {
      CC_CUDA(cudaStreamCreateWithFlags(&m->grp_stream[i], cudaStreamNonBlocking));
      CC_CUDA(cudaEventCreateWithFlags(&m->ev_join[i], cudaEventDisableTiming));
    }
  }
  const char* ng = getenv("CLIPCAP_B200_NO_GRAPH");
  m->use_graphs = !(ng != nullptr && ng[0] == '1');
  return CC_OK;
}

// ln_f + tied LM head of a decode step: fused-argmax keys (greedy) or fp32 logits (beam / sampling). Up to 16 rows run as
// one skinny kernel (LayerNorm applied while the rows are staged), otherwise LayerNorm(+reduce) + the tcgen05 head.
int head_decode(cc_gpt2* m, int nseq, bool keys, cudaStream_t s, int row0 = 0) {
  Stack& st = m->st;
  const cc_gpt2_cfg& c = m->cfg;
  if (st.skinny_step(nseq, row0))
    return skinny_gemm_run(st.h, m->lnf_g, m->lnf_b, c.eps, nullptr, c.d, nseq, m->wte16, c.V, c.d,
                           keys ? EPI_ARGMAX : EPI_F32, nullptr, keys ? static_cast<void*>(m->keys) : static_cast<void*>(m->logits),
                           keys ? 1 : m->v_ld, s);
  CC_TRY(st.ln_decode(m->lnf_g, m->lnf_b, m->lnf16, nseq, s, row0));  // absorbs the last layer's fc2 partial sums
  return gemm_run(keys ? m->p_head_keys : m->p_head_logits, nseq, s, row0);
}

inline float inv_temp_of(const cc_gen_cfg& g) { return 1.0f / (g.temperature > 0.f ? g.temperature : 1.0f); }
inline int nseq_of(int B, int beam) { return B * beam; }

// Everything of one generate call after the prefix rows have been written into st.h; results land in g_tokens /
// g_lengths / g_scores. Safe to capture into a graph: touches only handle-owned memory.
enum GenPhase { PHASE_ALL = 0, PHASE_PREFILL = 1, PHASE_DECODE = 2 };

int enqueue_generate(cc_gpt2* m, int B, int Tp, const cc_gen_cfg& g, cudaStream_t s, int phase = PHASE_ALL) {
  const cc_gpt2_cfg& c = m->cfg;
  const int d = c.d, EL = g.entry_length;
  const bool is_beam = g.mode == CC_GEN_BEAM;
  const bool is_sample = g.mode == CC_GEN_NUCLEUS || g.mode == CC_GEN_SAMPLE;
  const int beam = is_beam ? g.beam : 1;
  // one sampling step on the logits of the rows' last position (sample.cu)
  auto sample_step = [&](int step) -> int {
    const float lps = g.desired_sentence_length != 0
                          ? g.sentence_length_factor / static_cast<float>(g.desired_sentence_length)
                          : 0.f;
    return sample_run(m->logits, m->v_ld, c.V, g.mode, inv_temp_of(g), g.top_p, g.top_k, g.repetition_penalty, lps,
                      g.stop_token, m->d_history, g.mode == CC_GEN_SAMPLE ? g.n_history : 0, m->g_tokens, EL, step,
                      m->g_stopped, m->g_lengths, m->d_seed, nseq_of(B, beam), s);
  };
  const int nseq = B * beam;
  const float inv_temp = 1.0f / (g.temperature > 0.f ? g.temperature : 1.0f);
  Stack& st = m->st;
  st.launches = 0;
  st.pend_splits = 0;
  int extra = 0;
  int cur = 0;  // ping-pong index of the beam token / ancestry tables
  // ---- prefill: all Tp prefix positions at once; K,V go to slot img*beam
  // The first token needs only the last prefix position of the last block (LM head on row Tp-1): that block runs its
  // out-proj / MLP for those B rows alone (K, V of every position still go to the cache).
  // Two-phase generate (cc_generate_prefill / cc_generate_decode on different SM partitions): the last `defer` blocks of
  // the prefill, the LM head and the first token move to the head of the decode phase (cc_gpt2_set_prefill_defer) —
  // the caller's way of shifting work between its two streams; the kernels and their order are the same.
  static const bool full_last = [] {
    const char* e = getenv("CLIPCAP_B200_FULL_LAST_LAYER");
    return e != nullptr && e[0] == '1';
  }();
  const bool last_row = !full_last && Tp > 1;
  const int defer = phase == PHASE_ALL ? 0 : m->prefill_defer;
  const int l_split = c.L - defer;  // blocks [0, l_split) belong to the prefill phase
  auto prefill_blocks = [&](int l0, int l1) -> int {
    for (int l = l0; l < l1; ++l) {
      if (l == c.L - 1 && last_row) CC_TRY(st.layer_last_row(l, B, Tp, &m->kv, beam, s));
      else CC_TRY(st.layer_full(l, B, Tp, &m->kv, beam, s));
    }
    return CC_OK;
  };
  auto first_token = [&]() -> int {
    CC_TRY(layernorm_run(st.h + static_cast<size_t>(Tp - 1) * d, static_cast<int64_t>(Tp) * d, m->lnf_g, m->lnf_b,
                         m->lnf16, d, B, d, c.eps, s));
    extra += 1;
    if (is_sample) {
      CC_TRY(gen_reset_run(m->g_stopped, m->g_lengths, m->keys, m->g_scores, B, s));
      CC_TRY(gemm_run(m->p_head_logits, B, s));
      CC_TRY(sample_step(0));
    } else if (!is_beam) {
      CC_TRY(gen_reset_run(m->g_stopped, m->g_lengths, m->keys, m->g_scores, B, s));
      CC_TRY(gemm_run(m->p_head_keys, B, s));
      CC_TRY(greedy_select_run(m->keys, m->g_tokens, EL, 0, m->g_stopped, m->g_lengths, g.stop_token, B, s));
    } else {
      CC_TRY(gemm_run(m->p_head_logits, B, s));
      CC_TRY(row_topk_run(m->logits, m->v_ld, c.V, inv_temp, beam, nullptr, m->cand_val, m->cand_idx, B, s));
      CC_TRY(beam_init_run(m->cand_val, m->cand_idx, m->beam, beam, EL, m->t_max, Tp, g.stop_token, B, s));
    }
    extra += 3;
    return CC_OK;
  };
  if (phase != PHASE_DECODE) {
    CC_TRY(prefill_blocks(0, l_split));
    if (defer == 0) CC_TRY(first_token());
  }
  if (phase == PHASE_PREFILL) {
    m->launches = st.launches + extra;
    return CC_OK;
  }
  if (phase == PHASE_DECODE && defer > 0) {
    CC_TRY(prefill_blocks(l_split, c.L));
    CC_TRY(first_token());
  }
  if (is_sample) {
    for (int step = 1; step < EL; ++step) {
      const int pos = Tp + step - 1;
      CC_TRY(gpt2_embed_tokens_run(m->g_tokens + (step - 1), EL, m->wte32, m->wpe32, st.h, nseq, d, pos, c.V, s));
      for (int l = 0; l < c.L; ++l) CC_TRY(st.layer_decode(l, nseq, &m->kv, nullptr, pos, s));
      CC_TRY(head_decode(m, nseq, false, s));
      CC_TRY(sample_step(step));
      extra += 4;
    }
  } else if (!is_beam) {
    // Row groups: boundaries are multiples of 32 rows (the GEMM epilogue stores whole 32-row groups), group 0 stays on
    // the caller's stream, the others fork from it after the first token and join before the results are read.
    int G = m->decode_groups;
    int rows_per = ((nseq + G - 1) / G + 31) / 32 * 32;
    if (nseq < 64 || EL < 2) {
      G = 1;
      rows_per = nseq;
    }
    G = (nseq + rows_per - 1) / rows_per;
    if (G > 1) CC_CUDA(cudaEventRecord(m->ev_fork, s));
    for (int gi = 0; gi < G; ++gi) {
      const int row0 = gi * rows_per;
      const int n = nseq - row0 < rows_per ? nseq - row0 : rows_per;
      cudaStream_t gs = gi == 0 ? s : m->grp_stream[gi];
      if (gi > 0) CC_CUDA(cudaStreamWaitEvent(gs, m->ev_fork, 0));
      for (int step = 1; step < EL; ++step) {
        const int pos = Tp + step - 1;  // position of the token fed this step
        CC_TRY(gpt2_embed_tokens_run(m->g_tokens + static_cast<size_t>(row0) * EL + (step - 1), EL, m->wte32, m->wpe32,
                                     st.h + static_cast<size_t>(row0) * d, n, d, pos, c.V, gs));
        for (int l = 0; l < c.L; ++l) CC_TRY(st.layer_decode(l, n, &m->kv, nullptr, pos, gs, row0));
        CC_TRY(head_decode(m, n, true, gs, row0));
        CC_TRY(greedy_select_run(m->keys + row0, m->g_tokens + static_cast<size_t>(row0) * EL, EL, step,
                                 m->g_stopped + row0, m->g_lengths + row0, g.stop_token, n, gs));
        extra += 3;
      }
      if (gi > 0) {
        CC_CUDA(cudaEventRecord(m->ev_join[gi], gs));
        CC_CUDA(cudaStreamWaitEvent(s, m->ev_join[gi], 0));
      }
    }
  } else {
    for (int step = 1; step < EL; ++step) {
      const int pos = Tp + step - 1;  // position of the token fed this step
      CC_TRY(gpt2_embed_tokens_run(m->beam.tokens[cur] + (step - 1), EL, m->wte32, m->wpe32, st.h, nseq, d, pos, c.V, s));
      for (int l = 0; l < c.L; ++l) CC_TRY(st.layer_decode(l, nseq, &m->kv, m->beam.anc[cur], pos, s, 0, beam, Tp));
      CC_TRY(head_decode(m, nseq, false, s));
      CC_TRY(row_topk_run(m->logits, m->v_ld, c.V, inv_temp, beam, m->beam.stopped, m->cand_val, m->cand_idx, nseq, s));
      CC_TRY(beam_step_run(m->cand_val, m->cand_idx, m->beam, cur, beam, c.V, EL, m->t_max, step, pos, g.stop_token, B,
                           s));
      cur ^= 1;
      extra += 4;
    }
  }
  if (is_beam) {
    CC_TRY(beam_final_run(m->beam, cur, beam, EL, m->g_tokens, m->g_lengths, m->g_scores, B, s));
    extra += 1;
  }
  m->launches = st.launches + extra;
  return CC_OK;
}

}  // namespace
}  // namespace cc

extern "C" {

int cc_gpt2_create(cc_gpt2** h, const cc_gpt2_cfg* cfg, const cc_tensor* weights, int n_weights, int max_seqs,
                   int max_len) {
  using namespace cc;
  CC_REQUIRE(h != nullptr && cfg != nullptr && weights != nullptr, CC_EINVAL, "cc_gpt2_create: null argument");
  *h = nullptr;
  CC_TRY(check_device_sm100());
  CC_REQUIRE(max_seqs > 0 && max_len > 0, CC_EINVAL, "cc_gpt2_create: max_seqs=%d max_len=%d", max_seqs, max_len);
  CC_REQUIRE(cfg->d > 0 && cfg->d % 8 == 0 && cfg->L > 0 && cfg->H > 0 && cfg->V > 0 && cfg->n_pos > 0, CC_ESHAPE,
             "gpt2: d=%d L=%d H=%d V=%d n_pos=%d", cfg->d, cfg->L, cfg->H, cfg->V, cfg->n_pos);
  CC_REQUIRE(max_len <= cfg->n_pos, CC_ESHAPE, "gpt2: max_len %d exceeds n_positions %d", max_len, cfg->n_pos);
  cc_gpt2* m = new cc_gpt2();
  m->cfg = *cfg;
  if (m->cfg.eps <= 0.f) m->cfg.eps = 1e-5f;
  m->max_seqs = max_seqs;
  m->max_len = max_len;
  const int st = gpt2_build(m, weights, n_weights);
  if (st != CC_OK) {
    delete m;
    return st;
  }
  *h = m;
  return CC_OK;
}

int cc_gpt2_logits(cc_gpt2* m, const void* embeds, int dtype, int B, int T, int all_positions, float* logits,
                   void* stream) {
  using namespace cc;
  CC_REQUIRE(m != nullptr && embeds != nullptr && logits != nullptr, CC_EINVAL, "cc_gpt2_logits: null argument");
  CC_REQUIRE(B > 0 && T > 0 && T <= m->max_len && static_cast<long long>(B) * T <= m->max_rows, CC_ESHAPE,
             "cc_gpt2_logits: B=%d T=%d outside the handle (max_seqs %d, max_len %d)", B, T, m->max_seqs, m->max_len);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const cc_gpt2_cfg& c = m->cfg;
  const int d = c.d;
  Stack& st = m->st;
  CC_TRY(gpt2_embed_prefix_run(embeds, dtype, m->wpe32, st.h, B, T, d, 0, s));
  for (int l = 0; l < c.L; ++l) CC_TRY(st.layer_full(l, B, T, nullptr, 0, s));
  // The head writes straight into the caller's buffer: its plan is encoded once per output pointer and reused while the
  // caller keeps handing in the same buffer (the tensor map itself does not depend on the row count).
  const int rows = all_positions ? B * T : B;
  if (m->logits_plan_out != logits) {
    CC_TRY(gemm_plan(&m->p_logits_out, m->lnf16, d, m->max_rows, m->wte16, c.V, d, EPI_F32, nullptr, logits, c.V));
    m->logits_plan_out = logits;
  }
  if (all_positions)
    CC_TRY(layernorm_run(st.h, d, m->lnf_g, m->lnf_b, m->lnf16, d, rows, d, c.eps, s));
  else
    CC_TRY(layernorm_run(st.h + static_cast<size_t>(T - 1) * d, static_cast<int64_t>(T) * d, m->lnf_g, m->lnf_b,
                         m->lnf16, d, rows, d, c.eps, s));
  return gemm_run(m->p_logits_out, rows, s);
}

int cc_gpt2_embed(cc_gpt2* m, const int32_t* ids, int n, void* out, int out_dtype, void* stream) {
  using namespace cc;
  CC_REQUIRE(m != nullptr && ids != nullptr && out != nullptr && n > 0, CC_EINVAL, "cc_gpt2_embed: bad argument");
  return gather_rows_run(ids, m->wte32, out, out_dtype, n, m->cfg.d, m->cfg.V, static_cast<cudaStream_t>(stream));
}

namespace cc {
namespace {

int validate_generate(cc_gpt2* m, int B, int Tp, const cc_gen_cfg* g) {
  CC_REQUIRE(g->mode >= CC_GEN_GREEDY && g->mode <= CC_GEN_SAMPLE, CC_EINVAL, "cc_generate: mode %d", g->mode);
  const bool sampling = g->mode == CC_GEN_NUCLEUS || g->mode == CC_GEN_SAMPLE;
  if (sampling) {
    CC_REQUIRE(g->n_history >= 0 && g->n_history <= kMaxHistory && (g->n_history == 0 || g->history != nullptr), CC_EINVAL,
               "cc_generate: %d history tokens (max %d)", g->n_history, kMaxHistory);
    CC_REQUIRE(g->top_k >= 0, CC_EINVAL, "cc_generate: top_k %d", g->top_k);
  }
  const int beam = g->mode == CC_GEN_BEAM ? g->beam : 1;
  CC_REQUIRE(beam >= 1 && beam <= kMaxBeam, CC_ESHAPE, "cc_generate: beam size %d outside 1..%d", beam, kMaxBeam);
  CC_REQUIRE(g->entry_length >= 1 && g->entry_length <= m->max_entry, CC_ESHAPE,
             "cc_generate: entry_length %d outside 1..%d", g->entry_length, m->max_entry);
  CC_REQUIRE(B > 0 && B * beam <= m->max_seqs, CC_ESHAPE, "cc_generate: %d sequences x %d beams exceed max_seqs %d", B,
             beam, m->max_seqs);
  CC_REQUIRE(Tp > 0 && Tp + g->entry_length - 1 <= m->max_len, CC_ESHAPE,
             "cc_generate: prefix %d + %d generated positions exceed max_len %d", Tp, g->entry_length - 1, m->max_len);
  CC_REQUIRE(beam <= m->cfg.V, CC_ESHAPE, "cc_generate: beam %d > vocabulary %d", beam, m->cfg.V);
  return CC_OK;
}

// Runs one phase of a generate call on stream s: replayed from the handle's graph cache (greedy / beam) or enqueued
// directly (sampling calls carry many free parameters and are not cached).
int run_phase(cc_gpt2* m, int B, int Tp, const cc_gen_cfg* g, cudaStream_t s, int phase) {
  const bool sampling = g->mode == CC_GEN_NUCLEUS || g->mode == CC_GEN_SAMPLE;
  const int beam = g->mode == CC_GEN_BEAM ? g->beam : 1;
  const int EL = g->entry_length;
  if (!m->use_graphs || sampling) {
    const int before = phase == PHASE_DECODE ? m->launches : 0;
    CC_TRY(enqueue_generate(m, B, Tp, *g, s, phase));
    m->launches += before;
    return CC_OK;
  }
  uint32_t tbits;
  const float temp = g->temperature > 0.f ? g->temperature : 1.0f;
  memcpy(&tbits, &temp, 4);
  // Kernel nodes run in the context of the stream they were captured on. A caller inside an SM partition (a green
  // context stream, cc_set_sm_budget) must get nodes bound to ITS partition, so the capture happens on the caller's own
  // stream and the graph is cached per (SM budget, stream); the legacy default stream cannot capture and uses the
  // handle's private stream (whole device).
  const bool own = s != nullptr && s != cudaStreamLegacy && s != cudaStreamPerThread && sm_budget() > 0;
  cudaStream_t cs = own ? s : m->cap_stream;
  const cc_gpt2::Key key{g->mode, B, Tp, beam, EL, g->stop_token, tbits, sm_budget(),
                         own ? reinterpret_cast<uintptr_t>(s) : 0, phase, phase == PHASE_ALL ? 0 : m->prefill_defer};
  auto it = m->graphs.find(key);
  if (it == m->graphs.end()) {
    if (m->graphs.size() >= cc_gpt2::kMaxGraphs) {  // evict the least recently used graph (ragged batches, prompt lengths ...)
      auto victim = m->graphs.begin();
      for (auto jt = m->graphs.begin(); jt != m->graphs.end(); ++jt)
        if (jt->second.last_use < victim->second.last_use) victim = jt;
      CC_CUDA(cudaStreamSynchronize(s));  // an evicted graph may still be running on this stream
      cudaGraphExecDestroy(victim->second.exec);
      m->graphs.erase(victim);
    }
    cudaGraph_t graph = nullptr;
    CC_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    const int st = enqueue_generate(m, B, Tp, *g, cs, phase);
    const cudaError_t e = cudaStreamEndCapture(cs, &graph);
    if (st != CC_OK) {
      if (graph) cudaGraphDestroy(graph);
      (void)cudaGetLastError();
      return st;
    }
    CC_REQUIRE(e == cudaSuccess && graph != nullptr, CC_ECUDA, "cc_generate: graph capture failed: %s",
               cudaGetErrorString(e));
    cudaGraphExec_t exec = nullptr;
    const cudaError_t e2 = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    CC_REQUIRE(e2 == cudaSuccess, CC_ECUDA, "cc_generate: graph instantiate failed: %s", cudaGetErrorString(e2));
    cc_gpt2::GraphEntry ent;
    ent.exec = exec;
    ent.launches = m->launches;
    it = m->graphs.emplace(key, ent).first;
  }
  it->second.last_use = ++m->use_clock;
  CC_CUDA(cudaGraphLaunch(it->second.exec, s));
  m->launches = (phase == PHASE_DECODE ? m->phase_launches : 0) + it->second.launches;
  if (phase == PHASE_PREFILL) m->phase_launches = it->second.launches;
  return CC_OK;
}

int generate_prefill(cc_gpt2* m, const void* prefix, int dtype, int B, int Tp, const cc_gen_cfg* g, cudaStream_t s,
                     int phase) {
  const bool sampling = g->mode == CC_GEN_NUCLEUS || g->mode == CC_GEN_SAMPLE;
  // h = inputs_embeds + wpe[0..Tp-1]   (modeling_gpt2.py:579-585); reads the caller's buffer, so it stays outside the graph
  CC_TRY(gpt2_embed_prefix_run(prefix, dtype, m->wpe32, m->st.h, B, Tp, m->cfg.d, 0, s));
  if (sampling) {
    // per-call inputs of the sampling kernel live in handle-owned device memory
    const unsigned long long seed = g->seed;
    CC_CUDA(cudaMemcpyAsync(m->d_seed, &seed, sizeof(seed), cudaMemcpyHostToDevice, s));
    if (g->n_history > 0)
      CC_CUDA(cudaMemcpyAsync(m->d_history, g->history, sizeof(int32_t) * g->n_history, cudaMemcpyHostToDevice, s));
  }
  return run_phase(m, B, Tp, g, s, phase);
}

int generate_results(cc_gpt2* m, int B, int EL, int32_t* tokens, int32_t* lengths, float* scores, cudaStream_t s) {
  CC_CUDA(cudaMemcpyAsync(tokens, m->g_tokens, static_cast<size_t>(B) * EL * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
  CC_CUDA(cudaMemcpyAsync(lengths, m->g_lengths, static_cast<size_t>(B) * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
  if (scores != nullptr)
    CC_CUDA(cudaMemcpyAsync(scores, m->g_scores, static_cast<size_t>(B) * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return CC_OK;
}

}  // namespace
}  // namespace cc

int cc_generate(cc_gpt2* m, const void* prefix, int dtype, int B, int Tp, const cc_gen_cfg* g, int32_t* tokens,
                int32_t* lengths, float* scores, void* stream) {
  using namespace cc;
  CC_REQUIRE(m != nullptr && prefix != nullptr && g != nullptr && tokens != nullptr && lengths != nullptr, CC_EINVAL,
             "cc_generate: null argument");
  CC_TRY(validate_generate(m, B, Tp, g));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CC_TRY(generate_prefill(m, prefix, dtype, B, Tp, g, s, PHASE_ALL));
  return generate_results(m, B, g->entry_length, tokens, lengths, scores, s);
}

int cc_generate_prefill(cc_gpt2* m, const void* prefix, int dtype, int B, int Tp, const cc_gen_cfg* g, void* stream) {
  using namespace cc;
  CC_REQUIRE(m != nullptr && prefix != nullptr && g != nullptr, CC_EINVAL, "cc_generate_prefill: null argument");
  CC_TRY(validate_generate(m, B, Tp, g));
  return generate_prefill(m, prefix, dtype, B, Tp, g, static_cast<cudaStream_t>(stream), PHASE_PREFILL);
}

int cc_generate_decode(cc_gpt2* m, int B, int Tp, const cc_gen_cfg* g, int32_t* tokens, int32_t* lengths, float* scores,
                       void* stream) {
  using namespace cc;
  CC_REQUIRE(m != nullptr && g != nullptr && tokens != nullptr && lengths != nullptr, CC_EINVAL,
             "cc_generate_decode: null argument");
  CC_TRY(validate_generate(m, B, Tp, g));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CC_TRY(run_phase(m, B, Tp, g, s, PHASE_DECODE));
  return generate_results(m, B, g->entry_length, tokens, lengths, scores, s);
}

int cc_gpt2_set_prefill_defer(cc_gpt2* m, int blocks) {
  using namespace cc;
  CC_REQUIRE(m != nullptr, CC_EINVAL, "cc_gpt2_set_prefill_defer: null handle");
  CC_REQUIRE(blocks >= 0 && blocks < m->cfg.L, CC_ESHAPE, "cc_gpt2_set_prefill_defer: %d of %d blocks", blocks, m->cfg.L);
  m->prefill_defer = blocks;
  return CC_OK;
}

int cc_gpt2_last_launches(cc_gpt2* m) { return m ? m->launches : 0; }

void cc_gpt2_destroy(cc_gpt2* m) { delete m; }

}  // extern "C"
