// C-ABI odds and ends: error/version strings and the kernel-level test hooks declared at the bottom of
// include/clipcap_b200.h.  The hooks launch exactly the kernels the engines launch (same plans, same heuristics).
#include "common.h"
#include "decode.h"
#include "train.h"

extern "C" {

const char* cc_last_error(void) { return cc::get_error(); }

const char* cc_version(void) { return "clipcap_b200 0.1.0 sm_100a"; }

void cc_set_sm_budget(int n_sms) { cc::set_sm_budget(n_sms); }

int cc_get_sm_budget(void) { return cc::sm_budget(); }

void cc_prof_enable(int on) { cc::gemm_prof_enable(on != 0); }

void cc_prof_read(double* ms, double* flops, long long* n) {
  if (ms == nullptr || flops == nullptr || n == nullptr) return;
  cc::gemm_prof_read(ms, flops, n);
}

void cc_prof_read_family(int bn, double* ms, double* flops, long long* n) {
  if (ms == nullptr || flops == nullptr || n == nullptr) return;
  cc::gemm_prof_read_family(bn, ms, flops, n);
}

int cc_op_gemm(const void* a, int64_t lda, const void* w, const float* bias, void* out, int64_t ldc, int M, int N, int K,
               int epi, int bn, void* stream) {
  using namespace cc;
  CC_REQUIRE(a != nullptr && w != nullptr && out != nullptr, CC_EINVAL, "cc_op_gemm: null argument");
  CC_TRY(check_device_sm100());
  CC_REQUIRE(bn == 0 || bn == 32 || bn == 64 || bn == 128 || bn == 256 || bn == 512, CC_EINVAL,
             "cc_op_gemm: BLOCK_N %d not in {0,32,64,128,256} or 512 (256 x 256 CTA-pair tile)", bn);
  GemmPlan p;
  CC_TRY(gemm_plan(&p, static_cast<const __half*>(a), lda, M, static_cast<const __half*>(w), N, K, epi, bias, out, ldc));
  p.force_bn = bn;
  return gemm_run(p, M, static_cast<cudaStream_t>(stream));
}

int cc_op_tile_image(const float* image, int size, int tiles_per_axis, int pixels_per_tile, int step, float* tiles,
                     void* stream) {
  using namespace cc;
  CC_REQUIRE(image != nullptr && tiles != nullptr, CC_EINVAL, "cc_op_tile_image: null argument");
  CC_TRY(check_device_sm100());
  return tile_image_run(image, size, tiles_per_axis, pixels_per_tile, step, tiles, static_cast<cudaStream_t>(stream));
}

int cc_op_layernorm(const float* x, int64_t x_ld, const float* gamma, const float* beta, void* y, int64_t y_ld, int rows,
                    int d, float eps, void* stream) {
  using namespace cc;
  CC_REQUIRE(x != nullptr && gamma != nullptr && beta != nullptr && y != nullptr, CC_EINVAL,
             "cc_op_layernorm: null argument");
  CC_TRY(check_device_sm100());
  return layernorm_run(x, x_ld, gamma, beta, static_cast<__half*>(y), y_ld, rows, d, eps,
                       static_cast<cudaStream_t>(stream));
}

int cc_op_attention(const void* q, const void* k, const void* v, int64_t ld, void* o, int64_t ldo, int B, int S, int H,
                    int hd, int causal, float scale, void* stream) {
  using namespace cc;
  CC_REQUIRE(q != nullptr && k != nullptr && v != nullptr && o != nullptr, CC_EINVAL, "cc_op_attention: null argument");
  CC_TRY(check_device_sm100());
  return attention_run(static_cast<const __half*>(q), static_cast<const __half*>(k), static_cast<const __half*>(v), ld,
                       static_cast<__half*>(o), ldo, B, S, H, hd, causal != 0, scale, static_cast<cudaStream_t>(stream));
}

int cc_op_skinny_gemm(const float* x32, const float* gamma, const float* beta, float eps, const void* x16, int64_t ldx, int M,
                      const void* w, int N, int K, int epi, const float* bias, void* out, int64_t ldc, void* stream) {
  using namespace cc;
  CC_REQUIRE(w != nullptr && out != nullptr, CC_EINVAL, "cc_op_skinny_gemm: null argument");
  CC_TRY(check_device_sm100());
  return skinny_gemm_run(x32, gamma, beta, eps, static_cast<const __half*>(x16), ldx, M, static_cast<const __half*>(w), N, K,
                         epi, bias, out, ldc, static_cast<cudaStream_t>(stream));
}

int cc_op_attention_bwd(const void* q, const void* k, const void* v, int64_t ld, const void* d_o, int64_t ldo, void* dq,
                        void* dk, void* dv, int64_t ldd, int B, int S, int H, int hd, int causal, float scale,
                        void* stream) {
  using namespace cc;
  CC_REQUIRE(q != nullptr && k != nullptr && v != nullptr && d_o != nullptr && dq != nullptr && dk != nullptr &&
                 dv != nullptr,
             CC_EINVAL, "cc_op_attention_bwd: null argument");
  CC_TRY(check_device_sm100());
  return attention_bwd_run(static_cast<const __half*>(q), static_cast<const __half*>(k), static_cast<const __half*>(v), ld,
                           static_cast<const __half*>(d_o), ldo, static_cast<__half*>(dq), static_cast<__half*>(dk),
                           static_cast<__half*>(dv), ldd, B, S, H, hd, causal != 0, scale,
                           static_cast<cudaStream_t>(stream));
}

int cc_op_decode_attention(const void* qkv, void* kcache, void* vcache, const int32_t* anc, void* o, int nseq, int H,
                           int t_max, int pos, float scale, void* stream) {
  using namespace cc;
  CC_REQUIRE(qkv != nullptr && kcache != nullptr && vcache != nullptr && o != nullptr, CC_EINVAL,
             "cc_op_decode_attention: null argument");
  CC_TRY(check_device_sm100());
  return decode_attention_run(static_cast<const __half*>(qkv), static_cast<__half*>(kcache),
                              static_cast<__half*>(vcache), anc, static_cast<__half*>(o), nseq, H, t_max, pos, scale,
                              static_cast<cudaStream_t>(stream));
}

int cc_op_decode_attention_beam(const void* qkv, void* kcache, void* vcache, const int32_t* anc, void* o, int nseq, int H,
                                int t_max, int pos, int beam, int shared_len, float scale, void* stream) {
  using namespace cc;
  CC_REQUIRE(qkv != nullptr && kcache != nullptr && vcache != nullptr && o != nullptr && anc != nullptr, CC_EINVAL,
             "cc_op_decode_attention_beam: null argument");
  CC_REQUIRE(beam >= 1 && nseq % beam == 0 && shared_len >= 0 && shared_len <= pos, CC_ESHAPE,
             "cc_op_decode_attention_beam: beam %d, %d rows, shared_len %d, pos %d", beam, nseq, shared_len, pos);
  CC_TRY(check_device_sm100());
  return decode_attention_run(static_cast<const __half*>(qkv), static_cast<__half*>(kcache),
                              static_cast<__half*>(vcache), anc, static_cast<__half*>(o), nseq, H, t_max, pos, scale,
                              static_cast<cudaStream_t>(stream), beam, shared_len);
}

int cc_op_sample(const float* logits, int rows, int V, const cc_gen_cfg* g, int step, int32_t* tokens, int32_t* stopped,
                 int32_t* lengths, void* stream) {
  using namespace cc;
  CC_REQUIRE(logits != nullptr && g != nullptr && tokens != nullptr && stopped != nullptr && lengths != nullptr, CC_EINVAL,
             "cc_op_sample: null argument");
  CC_TRY(check_device_sm100());
  CC_REQUIRE(step >= 0 && step < g->entry_length, CC_EINVAL, "cc_op_sample: step %d outside entry_length %d", step,
             g->entry_length);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static unsigned long long* d_seed = nullptr;  // test hook: one scratch scalar per process
  if (d_seed == nullptr) CC_CUDA(cudaMalloc(&d_seed, sizeof(unsigned long long)));
  const unsigned long long seed = g->seed;
  CC_CUDA(cudaMemcpyAsync(d_seed, &seed, sizeof(seed), cudaMemcpyHostToDevice, s));
  const float inv_temp = 1.0f / (g->temperature > 0.f ? g->temperature : 1.0f);
  const float lps = g->desired_sentence_length != 0
                        ? g->sentence_length_factor / static_cast<float>(g->desired_sentence_length)
                        : 0.f;
  return sample_run(logits, V, V, g->mode, inv_temp, g->top_p, g->top_k, g->repetition_penalty, lps, g->stop_token,
                    g->history, g->mode == CC_GEN_SAMPLE ? g->n_history : 0, tokens, g->entry_length, step, stopped, lengths,
                    d_seed, rows, s);
}

}  // extern "C"
