// Kernels of the training step (train_kernels.cu) — backward of the blocks in common.h `Stack`, the LM-head
// cross-entropy, and the optimiser update. Composed by train.cu (cc_train_step).
#pragma once
#include "common.h"

namespace cc {

// dst[c][r] = src[r][c] for r < rows, c < cols; dst columns rows .. ld_dst-1 are zeroed (ld_dst % 8 == 0): the
// MN-major -> K-major turn that makes a weight gradient dW = dY^T X a TN GEMM over the token dimension.
int transpose16_run(const __half* src, int64_t ld, int rows, int cols, __half* dst, int64_t ld_dst, cudaStream_t s);
// out[c] = alpha * sum_r x[r * ld + c]   (bias / prefix_const gradients; two stages in a fixed summation order).
// `scratch` holds the per-slice partial sums: colsum_scratch_floats(cols) floats are always enough, fewer just mean fewer
// slices (at least `cols`).
size_t colsum_scratch_floats(int max_cols);
int colsum_f32_run(const float* x, int64_t ld, int rows, int cols, float alpha, float* out, float* scratch,
                   size_t scratch_floats, cudaStream_t s);
int colsum_f16_run(const __half* x, int64_t ld, int rows, int cols, float alpha, float* out, float* scratch,
                   size_t scratch_floats, cudaStream_t s);
// x[i] *= alpha
int scale_f32_run(float* x, int64_t n, float alpha, cudaStream_t s);

// gelu_new (transformers/activations.py:59-66) applied to a stored fp16 pre-activation, and its derivative:
// dhid[i] *= gelu_new'(pre[i]);  ReLU: dhid[i] = hid[i] > 0 ? dhid[i] : 0   (mapper.py:10 act=relu)
int gelu_new_fwd_run(const __half* pre, __half* out, int64_t n, cudaStream_t s);
int gelu_new_bwd_run(__half* dhid, const __half* pre, int64_t n, cudaStream_t s);
int relu_bwd_run(__half* dhid, const __half* hid, int64_t n, cudaStream_t s);

// LayerNorm backward for y = LN(x) * gamma + beta, rows of d (same two-pass fp32 statistics as the forward kernel):
//   dx_accum[r,:] += rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma
// dgamma / dbeta (nullable): alpha * sum_r dy * xhat, alpha * sum_r dy; `scratch` must hold ln_bwd_scratch_floats(d).
size_t ln_bwd_scratch_floats(int d);
int layernorm_bwd_run(const float* dy, int64_t dy_ld, const float* x, int64_t x_ld, const float* gamma, float* dx_accum,
                      int64_t dx_ld, int rows, int d, float eps, float* dgamma, float* dbeta, float alpha,
                      float* scratch, cudaStream_t s);

// Attention backward for one packed projection matrix (head h at column h*hd of each of q, k, v, row stride ld) and
// the gradient of the attention output d_o [B*S, ldo]: writes dq, dk, dv (same layout, row stride ldd). Softmax
// probabilities are recomputed. One CTA per (batch, head); S is bounded by shared memory (CC_ESHAPE beyond).
int attention_bwd_run(const __half* q, const __half* k, const __half* v, int64_t ld, const __half* d_o, int64_t ldo,
                      __half* dq, __half* dk, __half* dv, int64_t ldd, int B, int S, int H, int hd, bool causal,
                      float scale, cudaStream_t s);

// Cross-entropy with ignore_index = 0 over logits [rows, ld] (model.py:109-110): n_valid = #targets != 0,
// row_loss[r] = lse(logits[r]) - logits[r, t] (0 when ignored), dlogits[r, :] = coef * (softmax - onehot) in fp16 with
// coef = loss_scale / n_valid (zeros when ignored; columns V .. ld_d-1 zeroed).
int count_valid_run(const int32_t* targets, int n, int* n_valid, cudaStream_t s);
int ce_loss_run(const float* logits, int64_t ld, int V, const int32_t* targets, int rows, const int* n_valid,
                float loss_scale, float* row_loss, __half* dlogits, int64_t ld_d, cudaStream_t s);
// loss = sum(row_loss) / n_valid (NaN when nothing is valid, like F.cross_entropy)
int loss_reduce_run(const float* row_loss, int n, const int* n_valid, float* loss, cudaStream_t s);

// Teacher-forced LM input (model.py:45-49 + GPT-2 wpe): h[b, t] = (t < K ? prefix[b, t] : wte[max(tok[b, t-K], 0)]) +
// wpe[t]; targets[b, j] = max(tok[b, j], 0). prefix rows have stride prefix_ld between samples.
int train_embed_run(const int32_t* tokens, int B, int Tt, int K, const float* prefix, int64_t prefix_ld,
                    const float* wte, const float* wpe, float* h, int32_t* targets, int d, int V, cudaStream_t s);

// count += number of non-finite elements of x (overflow check of the gradients produced under the static loss scale)
int count_nonfinite_run(const float* x, int64_t n, int* count, cudaStream_t s);

// torch.optim.AdamW update (decoupled weight decay, bias correction), step >= 1.
int adamw_run(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
              float weight_decay, int step, cudaStream_t s);

}  // namespace cc
