// Shared host-side declarations for libclipcap_b200: status/error plumbing, device buffers, and the launchers of every
// kernel family (GEMM, LayerNorm, attention, element-wise glue). Engines (vit.cu, mapper.cu, gpt2.cu) compose these.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/clipcap_b200.h"

namespace cc {

void set_error(const char* fmt, ...);
const char* get_error();

#define CC_CUDA(expr)                                                                                  \
  do {                                                                                                 \
    cudaError_t e__ = (expr);                                                                          \
    if (e__ != cudaSuccess) {                                                                          \
      cc::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__));            \
      return CC_ECUDA;                                                                                 \
    }                                                                                                  \
  } while (0)

#define CC_TRY(expr)              \
  do {                            \
    int s__ = (expr);             \
    if (s__ != CC_OK) return s__; \
  } while (0)

#define CC_REQUIRE(cond, code, ...) \
  do {                              \
    if (!(cond)) {                  \
      cc::set_error(__VA_ARGS__);   \
      return (code);                \
    }                               \
  } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device (per-context) property of a kernel: opt in once per
// (call site, device) — a process may drive several GPUs through handles created on different devices.
#define CC_OPT_IN_SMEM(kern, bytes)                                                                          \
  do {                                                                                                       \
    static std::atomic<unsigned long long> done__{0};                                                        \
    int dev__ = 0;                                                                                           \
    CC_CUDA(cudaGetDevice(&dev__));                                                                          \
    const unsigned long long bit__ = 1ull << (dev__ & 63);                                                   \
    if (!(done__.load() & bit__)) {                                                                          \
      CC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (bytes)));             \
      done__.fetch_or(bit__);                                                                                \
    }                                                                                                        \
  } while (0)

// ------------------------------------------------------------------ device memory owned by an engine handle
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int alloc(size_t n);
  void release();
  template <class T>
  T* as() const {
    return reinterpret_cast<T*>(p);
  }
};

struct Arena {  // owns every allocation of one engine; freed in one go
  std::vector<void*> ptrs;
  size_t total = 0;
  int alloc(void** out, size_t bytes);
  template <class T>
  int alloc_t(T** out, size_t count) {
    return alloc(reinterpret_cast<void**>(out), count * sizeof(T));
  }
  void release();
  ~Arena() { release(); }
};

// Launch with programmatic dependent launch enabled (the kernel must call pdl_wait(), see ptx.cuh). Disabled with
// CLIPCAP_B200_NO_PDL=1.
bool pdl_enabled();
template <class... KArgs, class... Args>
cudaError_t launch_pdl_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster_x,
                               Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (cluster_x > 1) {  // thread-block cluster (CTA pairs of the cta_group::2 GEMM)
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = cluster_x;
    at[n].val.clusterDim.y = 1;
    at[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <class... KArgs, class... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  return launch_pdl_cluster(kern, grid, block, smem, s, 1, args...);
}

int check_device_sm100();  // CC_EARCH unless the current device is compute capability 10.x
// SMs a launch may count on: the device's SM count, or the caller's SM budget (cc_set_sm_budget) when the stream it
// enqueues on belongs to an SM partition (green context) smaller than the device.
int num_sms();
int device_sms();
void set_sm_budget(int n);  // 0 = whole device
int sm_budget();

// ------------------------------------------------------------------ weights lookup
// Finds `name` among the caller's tensors, checks the element count, and returns a device fp32 pointer. Host pointers
// are staged through the arena.
int find_weight(const cc_tensor* w, int n, const std::string& name, int64_t expect_numel, Arena& arena,
                const float** out);

// ------------------------------------------------------------------ GEMM: C[M,N] = A[M,K] * W[N,K]^T (+ epilogue)
enum Epi {
  EPI_F16_NONE = 0,   // C16 = acc + bias
  EPI_F16_RELU,       // C16 = relu(acc + bias)
  EPI_F16_QUICKGELU,  // C16 = x * sigmoid(1.702 x)
  EPI_F16_GELU_NEW,   // C16 = 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))
  EPI_F16_TANH,       // C16 = tanh(acc + bias)
  EPI_F32,            // C32 = acc + bias
  EPI_RESID_F32,      // C32 += acc + bias       (in-place fp32 residual stream)
  EPI_ARGMAX,         // keys[m] = max over n of pack(acc, n)   (fused greedy LM head; no logits written)
  EPI_PARTIAL_F32,    // split-K: partial[split][m][n] = acc over this split's k-range (summed by layernorm_reduce)
  EPI_F16_HEADS,      // packed QKV projection written head-major: out[((b*3 + which)*H + h)*S + tok][64] = acc + bias
  EPI_F16_GELU_ERF,   // C16 = 0.5 x (1 + erf(x / sqrt 2)): the exact GELU of the Swin MLP (CLAP audio tower)
  EPI_COUNT
};

struct GemmPlan {
  CUtensorMap map_a;     // A: [rows, K] fp16, row stride lda
  CUtensorMap map_b[4];  // W: [N, K] fp16, one box height per BLOCK_N in {32, 64, 128, 256}
  CUtensorMap map_c;     // out: [rows, N] fp16 / fp32 for the TMA store / reduce-add epilogues
  int max_rows = 0, N = 0, K = 0;
  int epi = EPI_F16_NONE;
  const float* bias = nullptr;
  void* out = nullptr;  // half* / float* / unsigned long long* by epilogue
  int64_t ldc = 0;
  int force_bn = 0;  // 0 = heuristic; 512 = the 256 x 256 CTA-pair tile (cta_group::2)
  int splits = 1;      // EPI_PARTIAL_F32: split-K factor and the row pitch of one split inside `out`
  int split_rows = 0;
  int heads_S = 0, heads_H = 0;  // EPI_F16_HEADS: tokens per image, heads (N = 3 * H * 64)
};

// TMA descriptor of a 2-D fp16 row-major [rows, cols] matrix (row stride ld elements), box {64 columns, box_rows rows},
// 128-byte swizzle: the layout tcgen05 K-major operands (and MN-major operands of 64 columns) expect.
int tma_map_f16_sw128(CUtensorMap* m, const __half* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

// Encodes the tensor maps. `a` must stay at this address with >= max_rows rows readable.
int gemm_plan(GemmPlan* p, const __half* a, int64_t lda, int max_rows, const __half* w, int N, int K, int epi,
              const float* bias, void* out, int64_t ldc);
// Split-K plan for skinny problems (decode): partial[s][split_rows][N] fp32 holds split s of A W^T without bias; rows
// M..split_rows of each split are scratch. The reduction over s happens, in order, inside layernorm_reduce_run.
// QKV projection of a [images * S, d] activation into the head-major layout the tcgen05 attention kernel reads:
// out is [images * 3 * H][S][64] fp16 (plane = (image * 3 + {q,k,v}) * H + head), so every (image, head) operand is one
// contiguous S x 64 block. N must be 3 * H * 64.
int gemm_plan_heads(GemmPlan* p, const __half* a, int64_t lda, int max_images, int S, int H, const __half* w, int K,
                    const float* bias, __half* out);
int gemm_plan_partial(GemmPlan* p, const __half* a, int64_t lda, int max_rows, const __half* w, int N, int K,
                      float* partial, int split_rows, int splits, int bn);
constexpr int kMaxSplitK = 16;  // largest split-K factor gemm_pick_split returns
void gemm_pick_split(int M, int N, int K, int* bn_out, int* splits_out);
// row0: first row of A / out covered (a multiple of 32); lets independent row groups of one plan run on different streams.
int gemm_run(const GemmPlan& p, int M, cudaStream_t s, int row0 = 0);
int gemm_pick_bn(int M, int N, int K);
// Live timing of GEMM launches with CUDA events on the launching stream (cc_prof_*; bench.py's roofline figure).
void gemm_prof_enable(bool on);
void gemm_prof_read(double* ms, double* flops, long long* n);
void gemm_prof_read_family(int bn, double* ms, double* flops, long long* n);

// pack(acc, n) used by EPI_ARGMAX: high 32 bits = order-preserving float key, low 32 bits = ~n (ties -> lowest index)
__host__ __device__ inline uint32_t argmax_key_index(unsigned long long key) { return ~static_cast<uint32_t>(key); }

// ------------------------------------------------------------------ skinny GEMM (skinny.cu): M <= 16 rows, weight streaming
// y[M, N] = X[M, K] W[N, K]^T with X = LayerNorm(x32; gamma, beta, eps) (x32 != nullptr, fp32 rows of stride ldx) or the
// fp16 rows x16; epilogues EPI_F16_NONE, EPI_F16_GELU_NEW, EPI_RESID_F32 (out += acc + bias), EPI_F32, EPI_ARGMAX.
// The single-image / small-beam decode step is built from it (Stack::layer_decode, gpt2.cu head).
constexpr int kSkinnyMaxRows = 16;
bool skinny_enabled();  // CLIPCAP_B200_NO_SKINNY=1 keeps the tcgen05 path for every row count
int skinny_gemm_run(const float* x32, const float* gamma, const float* beta, float eps, const __half* x16, int64_t ldx,
                    int M, const __half* w, int N, int K, int epi, const float* bias, void* out, int64_t ldc,
                    cudaStream_t s);

// ------------------------------------------------------------------ LayerNorm (fp32 in, fp16 out), row-strided
int layernorm_run(const float* x, int64_t x_ld, const float* gamma, const float* beta, __half* y, int64_t y_ld, int rows,
                  int d, float eps, cudaStream_t s);

// h[r,:] += bias + sum_s partial[s][r,:] (s ascending, fixed order), written back, then LayerNorm(h[r,:]) -> y fp16.
// partial == nullptr / splits == 0 degenerates to layernorm_run.
int layernorm_reduce_run(float* h, int64_t h_ld, const float* partial, int splits, int64_t split_stride,
                         const float* bias, const float* gamma, const float* beta, __half* y, int64_t y_ld, int rows,
                         int d, float eps, cudaStream_t s);

// ------------------------------------------------------------------ attention
// Full / causal self-attention over packed projections. q,k,v point at the first element of their column block inside
// one [B*S, ld] fp16 matrix (head h at column h*hd); o is [B*S, ldo] fp16. hd in {48, 64, 96, 128}.
int attention_run(const __half* q, const __half* k, const __half* v, int64_t ld, __half* o, int64_t ldo, int B, int S,
                  int H, int hd, bool causal, float scale, cudaStream_t s);

// tcgen05 attention for the ViT-L/14 shape (S = 257 = CLS + 256 patches, head dim 64, non-causal). Returns CC_ESHAPE
// without launching if the shape does not fit; attention_run tries it first.
bool vit_attention_fits(int S, int hd, bool causal, int64_t ld, int64_t ldo, const __half* q, const __half* k,
                        const __half* v);
int vit_attention_run(const __half* q, const __half* k, const __half* v, int64_t ld, __half* o, int64_t ldo, int B, int S,
                      int H, float scale, cudaStream_t s);
// Same kernel on the head-major QKV layout gemm_plan_heads writes ([B*3*H][S][64]): every operand tile is one contiguous
// block, which is what lets the loads run at HBM speed.
int vit_attention_heads_run(const __half* qkvh, __half* o, int64_t ldo, int B, int S, int H, float scale, cudaStream_t s);
constexpr int kVitAttnTokens = 257;
// tcgen05 attention for S <= 128 tokens, head dim 64 / 128, full or causal (mapper, GPT-2 prefill, training forward):
// one (sequence, head) = one 128-row tile (small_attention.cu). attention_run tries it after the ViT kernel.
bool small_attention_fits(int S, int hd, int64_t ld, int64_t ldo, const __half* q, const __half* k, const __half* v,
                          const __half* o);
// kcache / vcache != nullptr (head dim 64): the kernel also writes positions 0 .. S-1 of every (sequence, head) into the
// decode cache [slot = seq * slot_stride][head][t_max][64] (what kv_scatter_run does as a separate pass).
int small_attention_run(const __half* q, const __half* k, const __half* v, int64_t ld, __half* o, int64_t ldo, int B, int S,
                        int H, int hd, bool causal, float scale, cudaStream_t s, __half* kcache = nullptr,
                        __half* vcache = nullptr, int t_max = 0, int slot_stride = 1);

// KV cache: [layer][k|v][slot][head][t_max][64] fp16, the eight 16-byte chunks of a row rotated by its position (chunk c
// of position t at chunk c ^ (t & 7), attention.cu kv_chunk). Decode step: append this step's k,v (from qkv[nseq,3d]) at position
// `pos` of slot `seq`, then attend over positions 0..pos, position t being read from slot anc[seq*t_max + t]
// (anc == nullptr: the sequence's own slot).
// beam > 1 (with anc): rows seq = img * beam + b belong to one image and positions 0 .. shared_len-1 of all of them live
// in one cache slot (the prefix): those rows are loaded once per image.
int decode_attention_run(const __half* qkv, __half* kcache, __half* vcache, const int32_t* anc, __half* o, int nseq,
                         int H, int t_max, int pos, float scale, cudaStream_t s, int beam = 1, int shared_len = 0);
// Prefill: scatter k,v of [nseq*T, 3d] into the cache at positions pos0 .. pos0+T-1 of slot seq*slot_stride.
int kv_scatter_run(const __half* qkv, __half* kcache, __half* vcache, int nseq, int T, int H, int t_max, int pos0,
                   int slot_stride, cudaStream_t s);

// Attention of ONE query token per sequence over all S keys (the ViT class token of the last block: qrow = 0; the last
// prefix position of the last GPT-2 prefill block: qrow = S - 1, for which the causal mask hides nothing): o[b, h*64 ..].
// q/k/v element (b, which, h, t, c) lives at qkv[b*sb + which*sw + h*sh + t*st + c]; head dim 64.
int cls_attention_run(const __half* qkv, long long sb, long long sw, long long sh, long long st, __half* o, int64_t ldo,
                      int B, int S, int H, float scale, cudaStream_t s, int qrow = 0);

// ------------------------------------------------------------------ element-wise glue (elementwise.cu)
int convert_to_f16_run(const void* src, int src_dtype, __half* dst, int64_t n, cudaStream_t s);
int convert_from_f32_run(const float* src, int64_t src_ld, void* dst, int dst_dtype, int rows, int cols, cudaStream_t s);
int convert_to_f32_run(const void* src, int src_dtype, float* dst, int64_t n, cudaStream_t s);
// weights: fp32 [rows, cols] (optionally transposed on the fly) -> fp16 [rows_out, ld_out] zero padded
int pack_weight_run(const float* src, int rows, int cols, bool transpose, __half* dst, int64_t ld_out, cudaStream_t s);
// ViT
int vit_im2col_run(const void* pixels, int dtype, __half* out, int B, int img, int patch, int k_pad, cudaStream_t s);
int vit_embed_lnpre_run(const float* patches, const float* cls, const float* pos, const float* g, const float* b,
                        float* h, int B, int T, int w, float eps, cudaStream_t s);
int l2_normalize_run(float* x, int rows, int cols, cudaStream_t s);
// window tiling of a decoded square image [3, S, S] fp32 -> [n*n, 3, p, p] (tile (ty, tx) starts at pixel (ty, tx) * step)
int tile_image_run(const float* img, int S, int n, int p, int step, float* tiles, cudaStream_t s);
// mapper
int mapper_fill_const_run(float* h, const float* prefix_const, const float* pos_emb, int B, int P, int K, int d,
                          cudaStream_t s);
// GPT-2
int gpt2_embed_prefix_run(const void* embeds, int dtype, const float* wpe, float* h, int B, int T, int d, int pos0,
                          cudaStream_t s);
int gpt2_embed_tokens_run(const int32_t* tokens, int64_t tok_stride, const float* wte, const float* wpe, float* h, int n,
                          int d, int pos, int V, cudaStream_t s);
int gather_rows_run(const int32_t* ids, const float* table, void* out, int out_dtype, int n, int d, int V,
                    cudaStream_t s);

// ------------------------------------------------------------------ create-time weight helpers (stack.cu)
// Copies n fp32 values (device pointer) into the engine's arena so the caller's tensors need not outlive *_create.
int keep_f32(Arena& arena, const float* src_dev, size_t n, const float** out);
// fp32 [rows, cols] -> arena-owned fp16 [rows_out, ld_out] (rows_out = cols if transpose else rows), zero padded.
int pack_f16(Arena& arena, const float* src_dev, int rows, int cols, bool transpose, int64_t ld_out, const __half** out);

// ------------------------------------------------------------------ pre-LN transformer block stack (stack.cu)
// The three stacks of the path share one block shape:
//   h += Wo * attn(LN1(h) * Wqkv + bqkv) + bo ;  h += W2 * act(LN2(h) * W1 + b1) + b2
// ViT (OpenAI CLIP ResidualAttentionBlock: QuickGELU, full attention), the mapper's TransformerLayer
// (clipcap/model/mapper.py:91-110: ReLU, no qkv bias, full attention) and the GPT-2 block (HF modeling_gpt2.py:262-310:
// gelu_new, causal).  h is the fp32 residual stream; every GEMM operand is fp16.
struct LayerW {
  const float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
  const __half *wqkv = nullptr, *wo = nullptr, *w1 = nullptr, *w2 = nullptr;  // [3d,d] [d,d] [dff,d] [d,dff]
  const float *bqkv = nullptr, *bo = nullptr, *b1 = nullptr, *b2 = nullptr;   // nullable
};

struct KvCache {  // GPT-2 only: [layer][slot][head][t_max][64] fp16, K and V separate
  __half* k = nullptr;
  __half* v = nullptr;
  int slots = 0, t_max = 0;
  size_t layer_elems = 0;
};

struct Stack {
  int d = 0, dff = 0, H = 0, hd = 0, act_epi = EPI_F16_NONE, max_rows = 0;
  bool causal = false;
  float eps = 1e-5f, scale = 1.f;
  float* h = nullptr;  // [max_rows, d] fp32 residual stream
  __half *ln16 = nullptr, *qkv16 = nullptr, *att16 = nullptr, *mlp16 = nullptr;
  std::vector<LayerW> layers;
  std::vector<GemmPlan> p_qkv, p_o, p_1, p_2;
  int launches = 0;  // kernels enqueued since the counter was last reset

  // Decode path (GPT-2 only, dec_rows > 0): the two N = d projections run split-K into `part` and the LayerNorm that
  // follows (LN2, the next layer's LN1, or ln_f) folds bias + partial sums into h in a fixed order (deterministic).
  int dec_rows = 0, dec_rows_pad = 0;
  Arena* arena_ = nullptr;  // the owning engine's arena (set by init)
  // Tile shape and split factor depend on how many 128-row blocks a step really has (a handle created for 1280 beam rows
  // must not run a 256-row greedy step with the 1280-row plan), so the split-K plans are kept per row-block count and
  // built on first use; split s of a plan lives at part + s * split_rows * d with split_rows = blocks * 128.
  struct DecPlans {
    std::vector<GemmPlan> o, p2;
  };
  float* part = nullptr;  // fp32 partial sums, sized in plan() for the largest splits * split_rows over all block counts
  std::map<std::pair<int, int>, DecPlans> dec_plans;  // key = (row blocks, SM budget the plan was made for)
  int dec_plans_for(int blocks, DecPlans** out);
  // true when a decode step of nseq rows runs on the skinny kernels (<= 16 rows, no partial sums pending)
  bool skinny_step(int nseq, int row0) const;
  int pend_splits = 0;  // partial sums waiting in `part` for the next LayerNorm (0 = none)
  int64_t pend_stride = 0;  // elements between two splits of the pending partial sums
  const float* pend_bias = nullptr;

  // ViT-L/14 shape only (tokens per image == kVitAttnTokens, head dim 64): QKV is written head-major and attention
  // runs on the tcgen05 kernel. Set before plan().
  int heads_S = 0;
  // Only token 0 of every sequence is consumed after the last layer (ViT: ln_post(x[:, 0]), clip encode_image): with
  // cls_last_S = tokens per sequence set before plan(), layer_cls_only() runs that layer's attention output, out-proj and
  // MLP for the class-token rows alone (K and V still come from every token). Results for those rows are the same
  // function of the same inputs; the other rows of h are left as the previous layer wrote them.
  int cls_last_S = 0;
  GemmPlan p_o_cls, p_2_cls;
  int layer_cls_only(int l, int B, int S, cudaStream_t s);
  // Last block of a causal prefill whose only consumer is the last position (the LM head of the first generated token):
  // K, V of every position still go to the cache, everything after the attention scores runs for row S-1 of each sequence.
  int layer_last_row(int l, int B, int S, KvCache* kv, int slot_stride, cudaStream_t s);
  struct LastRowPlans {
    GemmPlan o, p2;
    int rows = 0;
  };
  std::map<std::pair<int, int>, LastRowPlans> last_row_plans;  // key = (layer, S): encoded on first use, never per call

  int init(Arena& arena, int d_, int dff_, int H_, int act_epi_, bool causal_, float eps_, int max_rows_,
           int dec_rows_ = 0);
  // LayerNorm of the first nseq rows of h for the decode path, absorbing pending split-K partial sums first.
  int ln_decode(const float* g, const float* b, __half* y, int nseq, cudaStream_t s, int row0 = 0);
  int plan();  // after `layers` is filled
  // Full-sequence pass of layer l over B sequences of S rows (rows b*S .. b*S+S-1 of h). If `kv` is given the layer's
  // K,V rows are also scattered into the cache at positions 0..S-1 of slot b*slot_stride.
  int layer_full(int l, int B, int S, KvCache* kv, int slot_stride, cudaStream_t s);
  // One decode step of layer l for nseq single-row sequences at cache position pos.
  // row0 > 0 (greedy only, anc == nullptr): the rows row0 .. row0 + nseq - 1, an independent group on its own stream.
  // beam / shared_len: rows per image and length of their common prefix when `anc` is given (beam search).
  int layer_decode(int l, int nseq, KvCache* kv, const int32_t* anc, int pos, cudaStream_t s, int row0 = 0, int beam = 1,
                   int shared_len = 0);
};

}  // namespace cc
