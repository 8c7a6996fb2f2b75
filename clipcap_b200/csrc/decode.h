// Device-side state and launchers of the decode loops (decode.cu).
#pragma once
#include "common.h"

namespace cc {

constexpr int kMaxBeam = 8;

// Beam-search state for n_img images x beam rows (row = img * beam + j), mirroring the tensors of
// clipcap/inference/base.py:67-71: scores, seq_lengths, has_stopped, tokens — plus the KV-cache ancestry table that
// replaces the reference's `embeds = embeds[next_tokens_source]` re-gather (base.py:112).
struct BeamState {
  float* scores;       // [rows]   sum of log-probs ("scores")
  float* seq_len;      // [rows]   fp32 like the reference's torch.ones(beam_size)
  int32_t* stopped;    // [rows]
  int32_t* tokens[2];  // ping-pong [rows, entry_len]
  int32_t* anc[2];     // ping-pong [rows, t_max]: cache slot holding position t of this row's history
};

int gen_reset_run(int32_t* stopped, int32_t* lengths, unsigned long long* keys, float* scores, int n, cudaStream_t s);
int greedy_select_run(unsigned long long* keys, int32_t* tokens, int entry_len, int step, int32_t* stopped,
                      int32_t* lengths, int stop_token, int n, cudaStream_t s);
int row_topk_run(const float* logits, int64_t ldl, int V, float inv_temp, int beam, const int32_t* stopped,
                 float* out_val, int32_t* out_idx, int rows, cudaStream_t s);
int beam_init_run(const float* cand_val, const int32_t* cand_idx, const BeamState& st, int beam, int entry_len,
                  int t_max, int Tp, int stop_token, int n_img, cudaStream_t s);
int beam_step_run(const float* cand_val, const int32_t* cand_idx, const BeamState& st, int in, int beam, int V,
                  int entry_len, int t_max, int step, int pos, int stop_token, int n_img, cudaStream_t s);
// One sampling step for `rows` sequences (sample.cu): NUCLEUS / SAMPLE token selection on fp32 logits.
int sample_run(const float* logits, int64_t ldl, int V, int mode, float inv_temp, float top_p, int top_k,
               float rep_penalty, float len_penalty_scale, int stop_token, const int32_t* prefix_hist, int n_prefix,
               int32_t* tokens, int entry_len, int step, int32_t* stopped, int32_t* lengths,
               const unsigned long long* seed, int rows, cudaStream_t s);
int beam_final_run(const BeamState& st, int cur, int beam, int entry_len, int32_t* tokens, int32_t* lengths,
                   float* scores, int n_img, cudaStream_t s);

}  // namespace cc
