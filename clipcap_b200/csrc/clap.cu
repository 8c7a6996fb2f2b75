// CLAP audio tower (BASELINE configs[4]): HTSAT-tiny — a Swin transformer over the log-mel spectrogram
// folded into a 256 x 256 image — followed by the 2-layer audio projection, [B, C, T <= 1024, 64] mel -> [B, 512].
// Replaces the arithmetic behind the reference's CLAPModel.forward (clipcap/encoders/clap.py:105-131:
// laion_clap's get_audio_embedding_from_data, not installed and not runnable as committed); the arithmetic reproduced is
// transformers' ClapAudioModelWithProjection (modeling_clap.py), the stand-in SURVEY §8c names — oracle/restate_clap.py.
//
// Data path (eval mode):
//   mel --[BatchNorm over mel bins, bicubic stretch of T to 1024 (align_corners), fold 4 time chunks along frequency,
//         4x4/4 patch gather: ONE kernel]--> cols [B*4096, 16] --tcgen05 GEMM--> tokens [B*4096, 96] fp32 --LayerNorm-->
//   4 stages of Swin blocks (depths 2/2/6/2, 96..768 channels, heads 4..32 => head dim 24, 8x8 windows, odd blocks shifted
//   by 4): LN -> QKV GEMM (q, k, v weights concatenated) -> window attention -> out-proj GEMM (fp32 residual, TMA
//   reduce-add) -> LN -> fc1 GEMM -> exact GELU -> fc2 GEMM (residual); between stages 2x2 patch merging (gather ->
//   LN(4C) -> GEMM 4C -> 2C);  final LN + mean over the 64 tokens -> Linear + ReLU -> Linear.
// Tokens stay in image order the whole time: the cyclic shift and the window partition / reverse of the reference are
// index arithmetic inside the attention kernel (it gathers its window's rows of the QKV matrix and scatters its output
// rows), not data movement. Only `is_longer == False` samples are supported (the fusion branch of the patch embedding
// only acts on clips longer than the 10 s window; BASELINE configs[4] feeds 10 s clips): channel 0 of the input is used.
#include <string>
#include <vector>

#include "common.h"
#include "ptx.cuh"

struct cc_clap {
  cc_clap_cfg cfg;
  int max_batch = 0;
  int grid0 = 0;  // tokens per side after the patch embedding (spec_size / patch)
  cc::Arena arena;
  struct Block {
    const float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *bqkv, *bo, *b1, *b2;
    const __half *wqkv, *wo, *w1, *w2;
    const float* rel_bias;  // [heads][64][64] gathered from relative_position_bias_table
    cc::GemmPlan p_qkv, p_o, p_1, p_2;
  };
  struct Stage {
    int C = 0, heads = 0, res = 0;  // channels, heads, tokens per side
    std::vector<Block> blocks;
    const float *mg = nullptr, *mb = nullptr;  // patch-merging LayerNorm(4C)
    const __half* wm = nullptr;                // reduction [2C, 4C]
    cc::GemmPlan p_merge;
  };
  std::vector<Stage> stages;
  const float *bn_scale = nullptr, *bn_shift = nullptr;  // BatchNorm2d(num_mel_bins) in eval mode as y = x * scale + shift
  const float *pe_b = nullptr, *pe_g = nullptr, *pe_beta = nullptr, *norm_g = nullptr, *norm_b = nullptr, *pb1 = nullptr,
              *pb2 = nullptr;
  const __half *pe_w = nullptr, *pw1 = nullptr, *pw2 = nullptr;
  __half *cols16 = nullptr, *ln16 = nullptr, *qkv16 = nullptr, *att16 = nullptr, *mlp16 = nullptr, *pool16 = nullptr,
         *hid16 = nullptr;
  float *x = nullptr, *merge32 = nullptr, *out32 = nullptr;
  cc::GemmPlan p_embed, p_proj1, p_proj2;
  int launches = 0;
};

namespace cc {
namespace {

// ---------------------------------------------------------------- input: BatchNorm + bicubic stretch + fold + 4x4 patches
__device__ __forceinline__ void cubic_coeffs(float t, float (&c)[4]) {  // PyTorch upsample_bicubic2d, A = -0.75
  const float A = -0.75f;
  const float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
  c[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  c[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
  c[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  c[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}

// one thread per (sample, token): 16 patch values. img[y][x] = stretched[t = (y / F) * (S / ratio ...)]: see reshape_mel2img.
template <typename SRC>
__global__ void clap_patches_kernel(const SRC* __restrict__ mel, long long sample_stride, int T, int F, int S, int patch,
                                    const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
                                    __half* __restrict__ cols, int B) {
  const int g = S / patch;                       // tokens per side
  const long long n = static_cast<long long>(B) * g * g;
  const int ratio = S / F;                       // time chunks stacked along frequency
  const int Tw = S * ratio;                      // stretched time length (1024)
  const int chunk = Tw / ratio;                  // = S
  const float scale = Tw > 1 ? static_cast<float>(T - 1) / static_cast<float>(Tw - 1) : 0.f;  // align_corners = True
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int px = static_cast<int>(i % g), py = static_cast<int>((i / g) % g);
    const long long b = i / (static_cast<long long>(g) * g);
    const SRC* src = mel + b * sample_stride;  // channel 0: [T][F]
    __half* dst = cols + i * (patch * patch);
    for (int ky = 0; ky < patch; ++ky) {
      const int y = py * patch + ky;
      const int f = y % F, r = y / F;
      for (int kx = 0; kx < patch; ++kx) {
        const int t = r * chunk + px * patch + kx;
        float v;
        if (T == Tw) {
          v = static_cast<float>(src[static_cast<long long>(t) * F + f]);
        } else {
          const float real = scale * static_cast<float>(t);
          const int i0 = static_cast<int>(floorf(real));
          float c[4];
          cubic_coeffs(real - static_cast<float>(i0), c);
          v = 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            int idx = i0 - 1 + k;
            idx = idx < 0 ? 0 : (idx > T - 1 ? T - 1 : idx);
            v += c[k] * static_cast<float>(src[static_cast<long long>(idx) * F + f]);
          }
        }
        dst[ky * patch + kx] = __float2half_rn(v * bn_scale[f] + bn_shift[f]);
      }
    }
  }
}

// ---------------------------------------------------------------- window attention (8x8 windows, head dim 24)
// One CTA per (sample, window, head), one thread per query token. The window is defined in the cyclically shifted frame;
// its tokens are gathered from / scattered to image order: shifted (sy, sx) <-> original ((sy + shift) % res, ...).
constexpr int CW_HD = 24;
__global__ void __launch_bounds__(64)
clap_window_attn_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, const float* __restrict__ rel_bias,
                        int res, int ws, int shift, int C, int heads, float scale) {
  __shared__ float ks[64][CW_HD + 1], vs[64][CW_HD + 1];
  const int n_tok = ws * ws;  // 64 (or res * res when the image is a single window)
  const int wpr = res / ws;
  const int head = blockIdx.y;
  const int win = blockIdx.x % (wpr * wpr);
  const long long b = blockIdx.x / (wpr * wpr);
  const int wy = win / wpr, wx = win - wy * wpr;
  const int i = threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  const int ty = i / ws, tx = i - ty * ws;
  const int sy = wy * ws + ty, sx = wx * ws + tx;
  const int oy = (sy + shift) % res, ox = (sx + shift) % res;
  const long long row = b * res * res + static_cast<long long>(oy) * res + ox;
  // region id of the token in the shifted frame (modeling_clap.py:525-550)
  int region = 0;
  if (shift > 0) {
    const int ry = sy < res - ws ? 0 : (sy < res - shift ? 1 : 2);
    const int rx = sx < res - ws ? 0 : (sx < res - shift ? 1 : 2);
    region = ry * 3 + rx;
  }
  __shared__ int regions[64];
  float q[CW_HD];
  if (i < n_tok) {
    const __half* base = qkv + row * (3LL * C) + head * CW_HD;
#pragma unroll
    for (int c = 0; c < CW_HD; c += 2) {
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(base + c));
      const float2 kk = __half22float2(*reinterpret_cast<const __half2*>(base + C + c));
      const float2 vv = __half22float2(*reinterpret_cast<const __half2*>(base + 2 * C + c));
      q[c] = a.x;
      q[c + 1] = a.y;
      ks[i][c] = kk.x;
      ks[i][c + 1] = kk.y;
      vs[i][c] = vv.x;
      vs[i][c + 1] = vv.y;
    }
    regions[i] = region;
  }
  __syncthreads();
  if (i >= n_tok) return;
  float p[64];
  float mx = -INFINITY;
  const float* bias = rel_bias + (static_cast<long long>(head) * 64 + i) * 64;
#pragma unroll
  for (int j = 0; j < 64; ++j) {
    float s = -INFINITY;
    if (j < n_tok) {
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < CW_HD; ++c) d += q[c] * ks[j][c];
      s = d * scale + bias[j] + (regions[j] != region ? -100.f : 0.f);
    }
    p[j] = s;
    mx = fmaxf(mx, s);
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 64; ++j) {
    p[j] = __expf(p[j] - mx);
    sum += p[j];
  }
  const float inv = 1.f / sum;
  float o[CW_HD];
#pragma unroll
  for (int c = 0; c < CW_HD; ++c) o[c] = 0.f;
#pragma unroll
  for (int j = 0; j < 64; ++j) {
    const float pj = p[j];  // exp(-inf) = 0 beyond n_tok
#pragma unroll
    for (int c = 0; c < CW_HD; ++c) o[c] += pj * vs[j][c];
  }
  __half* dst = out + row * C + head * CW_HD;
#pragma unroll
  for (int c = 0; c < CW_HD; c += 2) *reinterpret_cast<__half2*>(dst + c) = __floats2half2_rn(o[c] * inv, o[c + 1] * inv);
}

// ---------------------------------------------------------------- exact GELU, patch merging gather, final LN + mean pool
__global__ void gelu_erf_kernel(__half2* __restrict__ x, long long n2) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n2;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float2 v = __half22float2(x[i]);
    x[i] = __floats2half2_rn(0.5f * v.x * (1.f + erff(v.x * 0.70710678118654752f)),
                             0.5f * v.y * (1.f + erff(v.y * 0.70710678118654752f)));
  }
}

// out[b, y2 * (res/2) + x2, q * C + c] = x[b, (2 y2 + dy_q) * res + 2 x2 + dx_q, c], q = 0..3 <-> (dy, dx) = (0,0), (1,0), (0,1), (1,1)
__global__ void clap_merge_gather_kernel(const float4* __restrict__ x, float4* __restrict__ out, int res, int C4, long long n) {
  pdl_launch_dependents();
  pdl_wait();
  const int half = res / 2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C4);
    long long r = i / C4;
    const int qd = static_cast<int>(r % 4);
    r /= 4;
    const int x2 = static_cast<int>(r % half);
    r /= half;
    const int y2 = static_cast<int>(r % half);
    const long long b = r / half;
    const int dy = qd & 1, dx = qd >> 1;
    out[i] = x[(b * res * res + static_cast<long long>(2 * y2 + dy) * res + (2 * x2 + dx)) * C4 + c];
  }
}

// pooled[b, :] = mean over the `tokens` rows of LayerNorm(x[b, t, :])  (fp32 statistics and mean; one CTA per sample)
__global__ void __launch_bounds__(256)
clap_ln_meanpool_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ bta,
                        __half* __restrict__ pooled, int tokens, int C, float eps) {
  extern __shared__ float acc[];  // [C]
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int c = threadIdx.x; c < C; c += blockDim.x) acc[c] = 0.f;
  __syncthreads();
  for (int t = warp; t < tokens; t += nw) {
    const float* row = x + (static_cast<long long>(b) * tokens + t) * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += row[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float d = row[c] - mean;
      q += d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / C + eps);
    for (int c = lane; c < C; c += 32) atomicAdd(&acc[c], (row[c] - mean) * rstd * g[c] + bta[c]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) pooled[static_cast<long long>(b) * C + c] = __float2half_rn(acc[c] / tokens);
}

// fp32 -> fp32 LayerNorm, one warp per row (the patch-embedding norm, whose output is the residual stream itself)
__global__ void __launch_bounds__(256)
clap_ln_f32_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ bta,
                   float* __restrict__ y, long long rows, int C, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long nwarps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  for (long long r = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += nwarps) {
    const float* row = x + r * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += row[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float d = row[c] - mean;
      q += d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / C + eps);
    for (int c = lane; c < C; c += 32) y[r * C + c] = (row[c] - mean) * rstd * g[c] + bta[c];
  }
}

inline int grid_for(long long n, int threads) {
  const long long blocks = (n + threads - 1) / threads;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  return static_cast<int>(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

// ---------------------------------------------------------------- build
int get(const cc_tensor* w, int nw, const std::string& name, int64_t numel, Arena& stage, const float** out) {
  return find_weight(w, nw, name, numel, stage, out);
}

int clap_build(cc_clap* m, const cc_tensor* w, int nw) {
  const cc_clap_cfg& c = m->cfg;
  Arena& A = m->arena;
  Arena stage;
  const int B = m->max_batch;
  const int F = c.num_mel_bins, S = c.spec_size;
  m->grid0 = S / c.patch;
  const std::string enc = "audio_model.audio_encoder.";
  const int pp = c.patch * c.patch;
  // BatchNorm2d(num_mel_bins) in eval mode -> per-bin scale / shift (host side: 64 values)
  {
    const float *bw, *bb, *rm, *rv;
    CC_TRY(get(w, nw, enc + "batch_norm.weight", F, stage, &bw));
    CC_TRY(get(w, nw, enc + "batch_norm.bias", F, stage, &bb));
    CC_TRY(get(w, nw, enc + "batch_norm.running_mean", F, stage, &rm));
    CC_TRY(get(w, nw, enc + "batch_norm.running_var", F, stage, &rv));
    std::vector<float> hw(F), hb(F), hm(F), hv(F), sc(F), sh(F);
    CC_CUDA(cudaMemcpy(hw.data(), bw, F * 4, cudaMemcpyDeviceToHost));
    CC_CUDA(cudaMemcpy(hb.data(), bb, F * 4, cudaMemcpyDeviceToHost));
    CC_CUDA(cudaMemcpy(hm.data(), rm, F * 4, cudaMemcpyDeviceToHost));
    CC_CUDA(cudaMemcpy(hv.data(), rv, F * 4, cudaMemcpyDeviceToHost));
    for (int f = 0; f < F; ++f) {
      sc[f] = hw[f] / sqrtf(hv[f] + 1e-5f);
      sh[f] = hb[f] - hm[f] * sc[f];
    }
    float *dsc, *dsh;
    CC_TRY(A.alloc_t(&dsc, F));
    CC_TRY(A.alloc_t(&dsh, F));
    CC_CUDA(cudaMemcpy(dsc, sc.data(), F * 4, cudaMemcpyHostToDevice));
    CC_CUDA(cudaMemcpy(dsh, sh.data(), F * 4, cudaMemcpyHostToDevice));
    m->bn_scale = dsc;
    m->bn_shift = dsh;
  }
  // activations (stage 0 is the largest everywhere: rows halve by 4, channels double)
  const size_t rows0 = static_cast<size_t>(B) * m->grid0 * m->grid0;
  const int C0 = c.embed;
  CC_TRY(A.alloc_t(&m->cols16, rows0 * pp));
  CC_TRY(A.alloc_t(&m->x, rows0 * C0));
  CC_TRY(A.alloc_t(&m->merge32, rows0 * C0));      // rows / 4 x 4C
  CC_TRY(A.alloc_t(&m->ln16, rows0 * C0));
  CC_TRY(A.alloc_t(&m->qkv16, rows0 * 3 * C0));
  CC_TRY(A.alloc_t(&m->att16, rows0 * C0));
  CC_TRY(A.alloc_t(&m->mlp16, rows0 * 4 * C0));
  const int n_stages = 4;
  const int Cl = C0 << (n_stages - 1);
  const size_t Bp = (static_cast<size_t>(B) + 31) / 32 * 32;  // GEMM epilogues write whole 32-row groups
  CC_TRY(A.alloc_t(&m->pool16, Bp * Cl));
  CC_TRY(A.alloc_t(&m->hid16, Bp * c.projection_dim));
  CC_TRY(A.alloc_t(&m->out32, Bp * c.projection_dim));

  // patch embedding: conv [C0, 1, p, p] == [C0, p*p] GEMM operand
  {
    const float *pw, *pb, *g, *b;
    CC_TRY(get(w, nw, enc + "patch_embed.proj.weight", static_cast<int64_t>(C0) * pp, stage, &pw));
    CC_TRY(get(w, nw, enc + "patch_embed.proj.bias", C0, stage, &pb));
    CC_TRY(get(w, nw, enc + "patch_embed.norm.weight", C0, stage, &g));
    CC_TRY(get(w, nw, enc + "patch_embed.norm.bias", C0, stage, &b));
    CC_TRY(pack_f16(A, pw, C0, pp, false, pp, &m->pe_w));
    CC_TRY(keep_f32(A, pb, C0, &m->pe_b));
    CC_TRY(keep_f32(A, g, C0, &m->pe_g));
    CC_TRY(keep_f32(A, b, C0, &m->pe_beta));
    // conv output goes to merge32 (scratch), its LayerNorm into the residual stream x
    CC_TRY(gemm_plan(&m->p_embed, m->cols16, pp, static_cast<int>(rows0), m->pe_w, C0, pp, EPI_F32, m->pe_b, m->merge32, C0));
  }
  // relative position index of an 8x8 window (modeling_clap.py:427-438)
  const int ws = c.window, nt = ws * ws;
  std::vector<int> rel_index(static_cast<size_t>(nt) * nt);
  for (int i = 0; i < nt; ++i)
    for (int j = 0; j < nt; ++j) {
      const int dy = i / ws - j / ws + ws - 1, dx = i % ws - j % ws + ws - 1;
      rel_index[static_cast<size_t>(i) * nt + j] = dy * (2 * ws - 1) + dx;
    }
  m->stages.resize(n_stages);
  for (int si = 0; si < n_stages; ++si) {
    cc_clap::Stage& st = m->stages[si];
    st.C = C0 << si;
    st.heads = c.heads[si];
    st.res = m->grid0 >> si;
    CC_REQUIRE(st.C % st.heads == 0 && st.C / st.heads == CW_HD, CC_ESHAPE, "clap: stage %d head dim %d (kernels need %d)", si,
               st.C / st.heads, CW_HD);
    CC_REQUIRE(st.res >= ws && st.res % ws == 0, CC_ESHAPE, "clap: stage %d resolution %d vs window %d", si, st.res, ws);
    const int C = st.C, rows = B * st.res * st.res;
    const int tsz = (2 * ws - 1) * (2 * ws - 1);
    st.blocks.resize(c.depths[si]);
    for (int bi = 0; bi < c.depths[si]; ++bi) {
      cc_clap::Block& bk = st.blocks[bi];
      const std::string p = enc + "layers." + std::to_string(si) + ".blocks." + std::to_string(bi) + ".";
      const float *g1, *b1, *g2, *b2, *wq, *bq, *wk, *bk_, *wv, *bv, *wo, *bo, *w1, *bb1, *w2, *bb2, *tab;
      const int64_t CC = static_cast<int64_t>(C) * C;
      CC_TRY(get(w, nw, p + "layernorm_before.weight", C, stage, &g1));
      CC_TRY(get(w, nw, p + "layernorm_before.bias", C, stage, &b1));
      CC_TRY(get(w, nw, p + "layernorm_after.weight", C, stage, &g2));
      CC_TRY(get(w, nw, p + "layernorm_after.bias", C, stage, &b2));
      CC_TRY(get(w, nw, p + "attention.self.query.weight", CC, stage, &wq));
      CC_TRY(get(w, nw, p + "attention.self.query.bias", C, stage, &bq));
      CC_TRY(get(w, nw, p + "attention.self.key.weight", CC, stage, &wk));
      CC_TRY(get(w, nw, p + "attention.self.key.bias", C, stage, &bk_));
      CC_TRY(get(w, nw, p + "attention.self.value.weight", CC, stage, &wv));
      CC_TRY(get(w, nw, p + "attention.self.value.bias", C, stage, &bv));
      CC_TRY(get(w, nw, p + "attention.self.relative_position_bias_table", static_cast<int64_t>(tsz) * st.heads, stage, &tab));
      CC_TRY(get(w, nw, p + "attention.output.dense.weight", CC, stage, &wo));
      CC_TRY(get(w, nw, p + "attention.output.dense.bias", C, stage, &bo));
      CC_TRY(get(w, nw, p + "intermediate.dense.weight", 4 * CC, stage, &w1));
      CC_TRY(get(w, nw, p + "intermediate.dense.bias", 4 * C, stage, &bb1));
      CC_TRY(get(w, nw, p + "output.dense.weight", 4 * CC, stage, &w2));
      CC_TRY(get(w, nw, p + "output.dense.bias", C, stage, &bb2));
      CC_TRY(keep_f32(A, g1, C, &bk.ln1_g));
      CC_TRY(keep_f32(A, b1, C, &bk.ln1_b));
      CC_TRY(keep_f32(A, g2, C, &bk.ln2_g));
      CC_TRY(keep_f32(A, b2, C, &bk.ln2_b));
      // q, k, v projections as one [3C, C] operand (+ [3C] bias)
      __half* wqkv = nullptr;
      float* bqkv = nullptr;
      CC_TRY(A.alloc_t(&wqkv, 3 * static_cast<size_t>(CC)));
      CC_TRY(A.alloc_t(&bqkv, 3 * static_cast<size_t>(C)));
      CC_TRY(pack_weight_run(wq, C, C, false, wqkv, C, nullptr));
      CC_TRY(pack_weight_run(wk, C, C, false, wqkv + CC, C, nullptr));
      CC_TRY(pack_weight_run(wv, C, C, false, wqkv + 2 * CC, C, nullptr));
      CC_CUDA(cudaMemcpy(bqkv, bq, C * 4, cudaMemcpyDeviceToDevice));
      CC_CUDA(cudaMemcpy(bqkv + C, bk_, C * 4, cudaMemcpyDeviceToDevice));
      CC_CUDA(cudaMemcpy(bqkv + 2 * C, bv, C * 4, cudaMemcpyDeviceToDevice));
      CC_CUDA(cudaStreamSynchronize(nullptr));
      bk.wqkv = wqkv;
      bk.bqkv = bqkv;
      CC_TRY(pack_f16(A, wo, C, C, false, C, &bk.wo));
      CC_TRY(keep_f32(A, bo, C, &bk.bo));
      CC_TRY(pack_f16(A, w1, 4 * C, C, false, C, &bk.w1));
      CC_TRY(keep_f32(A, bb1, 4 * static_cast<size_t>(C), &bk.b1));
      CC_TRY(pack_f16(A, w2, C, 4 * C, false, 4 * C, &bk.w2));
      CC_TRY(keep_f32(A, bb2, C, &bk.b2));
      // relative position bias [heads][64][64] (table [tsz, heads] gathered through the index; host side, 64 KB per block)
      {
        std::vector<float> table(static_cast<size_t>(tsz) * st.heads), bias(static_cast<size_t>(st.heads) * 64 * 64, 0.f);
        CC_CUDA(cudaMemcpy(table.data(), tab, table.size() * 4, cudaMemcpyDeviceToHost));
        for (int h = 0; h < st.heads; ++h)
          for (int i = 0; i < nt; ++i)
            for (int j = 0; j < nt; ++j)
              bias[(static_cast<size_t>(h) * 64 + i) * 64 + j] = table[static_cast<size_t>(rel_index[static_cast<size_t>(i) * nt + j]) * st.heads + h];
        float* d = nullptr;
        CC_TRY(A.alloc_t(&d, bias.size()));
        CC_CUDA(cudaMemcpy(d, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
        bk.rel_bias = d;
      }
      CC_TRY(gemm_plan(&bk.p_qkv, m->ln16, C, rows, bk.wqkv, 3 * C, C, EPI_F16_NONE, bk.bqkv, m->qkv16, 3 * C));
      CC_TRY(gemm_plan(&bk.p_o, m->att16, C, rows, bk.wo, C, C, EPI_RESID_F32, bk.bo, m->x, C));
      CC_TRY(gemm_plan(&bk.p_1, m->ln16, C, rows, bk.w1, 4 * C, C, EPI_F16_NONE, bk.b1, m->mlp16, 4 * C));
      CC_TRY(gemm_plan(&bk.p_2, m->mlp16, 4 * C, rows, bk.w2, C, 4 * C, EPI_RESID_F32, bk.b2, m->x, C));
      stage.release();
    }
    if (si < n_stages - 1) {
      const std::string p = enc + "layers." + std::to_string(si) + ".downsample.";
      const float *g, *b, *wr;
      CC_TRY(get(w, nw, p + "norm.weight", 4 * C, stage, &g));
      CC_TRY(get(w, nw, p + "norm.bias", 4 * C, stage, &b));
      CC_TRY(get(w, nw, p + "reduction.weight", 8LL * C * C, stage, &wr));
      CC_TRY(keep_f32(A, g, 4 * static_cast<size_t>(C), &st.mg));
      CC_TRY(keep_f32(A, b, 4 * static_cast<size_t>(C), &st.mb));
      CC_TRY(pack_f16(A, wr, 2 * C, 4 * C, false, 4 * C, &st.wm));
      // merged rows: LN(4C) of merge32 -> ln16 -> GEMM -> x, which becomes the next stage's [rows / 4, 2C] stream (the
      // gather that read the old x is two kernels upstream on the same dependency chain)
      CC_TRY(gemm_plan(&st.p_merge, m->ln16, 4 * C, rows / 4, st.wm, 2 * C, 4 * C, EPI_F32, nullptr, m->x, 2 * C));
      stage.release();
    }
  }
  {
    const float *g, *b, *w1, *b1, *w2, *b2;
    CC_TRY(get(w, nw, enc + "norm.weight", Cl, stage, &g));
    CC_TRY(get(w, nw, enc + "norm.bias", Cl, stage, &b));
    CC_TRY(get(w, nw, "audio_projection.linear1.weight", static_cast<int64_t>(c.projection_dim) * Cl, stage, &w1));
    CC_TRY(get(w, nw, "audio_projection.linear1.bias", c.projection_dim, stage, &b1));
    CC_TRY(get(w, nw, "audio_projection.linear2.weight", static_cast<int64_t>(c.projection_dim) * c.projection_dim, stage, &w2));
    CC_TRY(get(w, nw, "audio_projection.linear2.bias", c.projection_dim, stage, &b2));
    CC_TRY(keep_f32(A, g, Cl, &m->norm_g));
    CC_TRY(keep_f32(A, b, Cl, &m->norm_b));
    CC_TRY(pack_f16(A, w1, c.projection_dim, Cl, false, Cl, &m->pw1));
    CC_TRY(keep_f32(A, b1, c.projection_dim, &m->pb1));
    CC_TRY(pack_f16(A, w2, c.projection_dim, c.projection_dim, false, c.projection_dim, &m->pw2));
    CC_TRY(keep_f32(A, b2, c.projection_dim, &m->pb2));
    CC_TRY(gemm_plan(&m->p_proj1, m->pool16, Cl, B, m->pw1, c.projection_dim, Cl, EPI_F16_RELU, m->pb1, m->hid16, c.projection_dim));
    CC_TRY(gemm_plan(&m->p_proj2, m->hid16, c.projection_dim, B, m->pw2, c.projection_dim, c.projection_dim, EPI_F32, m->pb2,
                     m->out32, c.projection_dim));
  }
  return CC_OK;
}

}  // namespace
}  // namespace cc

extern "C" {

int cc_clap_create(cc_clap** h, const cc_clap_cfg* cfg, const cc_tensor* weights, int n_weights, int max_batch) {
  using namespace cc;
  CC_REQUIRE(h != nullptr && cfg != nullptr && weights != nullptr, CC_EINVAL, "cc_clap_create: null argument");
  *h = nullptr;
  CC_TRY(check_device_sm100());
  CC_REQUIRE(max_batch > 0, CC_EINVAL, "cc_clap_create: max_batch %d", max_batch);
  CC_REQUIRE(cfg->num_mel_bins > 0 && cfg->spec_size % cfg->num_mel_bins == 0 && cfg->patch > 0 &&
                 cfg->spec_size % cfg->patch == 0 && (cfg->patch * cfg->patch) % 8 == 0 && cfg->embed % 8 == 0 &&
                 cfg->window == 8 && cfg->projection_dim % 8 == 0,
             CC_ESHAPE, "clap: mel bins %d spec %d patch %d embed %d window %d projection %d", cfg->num_mel_bins,
             cfg->spec_size, cfg->patch, cfg->embed, cfg->window, cfg->projection_dim);
  for (int i = 0; i < 4; ++i)
    CC_REQUIRE(cfg->depths[i] > 0 && cfg->heads[i] > 0, CC_ESHAPE, "clap: stage %d depth %d heads %d", i, cfg->depths[i], cfg->heads[i]);
  CC_REQUIRE((cfg->spec_size / cfg->patch) % 64 == 0, CC_ESHAPE, "clap: %d tokens per side must be a multiple of 64",
             cfg->spec_size / cfg->patch);
  cc_clap* m = new cc_clap();
  m->cfg = *cfg;
  if (m->cfg.eps <= 0.f) m->cfg.eps = 1e-5f;
  m->max_batch = max_batch;
  const int st = clap_build(m, weights, n_weights);
  if (st != CC_OK) {
    delete m;
    return st;
  }
  *h = m;
  return CC_OK;
}

int cc_clap_forward(cc_clap* m, const void* mel, int mel_dtype, int B, int channels, int T, int normalize, void* out,
                    int out_dtype, int stop_after_stage, float* dump, void* stream) {
  using namespace cc;
  CC_REQUIRE(m != nullptr && mel != nullptr && out != nullptr, CC_EINVAL, "cc_clap_forward: null argument");
  CC_REQUIRE(B > 0 && B <= m->max_batch, CC_ESHAPE, "cc_clap_forward: batch %d outside 1..%d", B, m->max_batch);
  const cc_clap_cfg& c = m->cfg;
  const int F = c.num_mel_bins, S = c.spec_size, ratio = S / F;
  CC_REQUIRE(channels >= 1 && T >= 1 && T <= S * ratio, CC_ESHAPE, "cc_clap_forward: %d channels, %d frames (max %d)", channels,
             T, S * ratio);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  m->launches = 0;
  const int g0 = m->grid0;
  const long long tokens0 = static_cast<long long>(B) * g0 * g0;
  const long long sample_stride = static_cast<long long>(channels) * T * F;
  if (mel_dtype == CC_F32)
    clap_patches_kernel<float><<<grid_for(tokens0, 128), 128, 0, s>>>(static_cast<const float*>(mel), sample_stride, T, F, S,
                                                                       c.patch, m->bn_scale, m->bn_shift, m->cols16, B);
  else if (mel_dtype == CC_F16)
    clap_patches_kernel<__half><<<grid_for(tokens0, 128), 128, 0, s>>>(static_cast<const __half*>(mel), sample_stride, T, F, S,
                                                                        c.patch, m->bn_scale, m->bn_shift, m->cols16, B);
  else {
    set_error("cc_clap_forward: unknown mel dtype %d", mel_dtype);
    return CC_EINVAL;
  }
  CC_CUDA(cudaGetLastError());
  CC_TRY(gemm_run(m->p_embed, static_cast<int>(tokens0), s));  // conv as GEMM -> merge32 (scratch)
  // patch LayerNorm, fp32 -> the fp32 residual stream
  CC_CUDA(launch_pdl(clap_ln_f32_kernel, dim3(grid_for(tokens0 * 32, 256)), dim3(256), 0, s,
                     static_cast<const float*>(m->merge32), m->pe_g, m->pe_beta, m->x, tokens0, c.embed, c.eps));
  m->launches += 3;
  for (size_t si = 0; si < m->stages.size(); ++si) {
    cc_clap::Stage& st = m->stages[si];
    const int C = st.C, res = st.res;
    const int rows = B * res * res;
    const int ws = res <= c.window ? res : c.window;
    const int windows = B * (res / ws) * (res / ws);
    for (size_t bi = 0; bi < st.blocks.size(); ++bi) {
      cc_clap::Block& bk = st.blocks[bi];
      const int shift = (bi % 2 == 1 && res > c.window) ? c.window / 2 : 0;
      CC_TRY(layernorm_run(m->x, C, bk.ln1_g, bk.ln1_b, m->ln16, C, rows, C, c.eps, s));
      CC_TRY(gemm_run(bk.p_qkv, rows, s));
      CC_CUDA(launch_pdl(clap_window_attn_kernel, dim3(windows, st.heads), dim3(64), 0, s,
                         static_cast<const __half*>(m->qkv16), m->att16, bk.rel_bias, res, ws, shift, C, st.heads,
                         1.0f / sqrtf(static_cast<float>(CW_HD))));
      CC_TRY(gemm_run(bk.p_o, rows, s));
      CC_TRY(layernorm_run(m->x, C, bk.ln2_g, bk.ln2_b, m->ln16, C, rows, C, c.eps, s));
      CC_TRY(gemm_run(bk.p_1, rows, s));
      const long long n2 = static_cast<long long>(rows) * 4 * C / 2;
      CC_CUDA(launch_pdl(gelu_erf_kernel, dim3(grid_for(n2, 256)), dim3(256), 0, s, reinterpret_cast<__half2*>(m->mlp16), n2));
      CC_TRY(gemm_run(bk.p_2, rows, s));
      m->launches += 8;
    }
    if (static_cast<int>(si) == stop_after_stage && dump != nullptr) {
      CC_CUDA(cudaMemcpyAsync(dump, m->x, static_cast<size_t>(rows) * C * sizeof(float), cudaMemcpyDeviceToDevice, s));
      return CC_OK;
    }
    if (si + 1 < m->stages.size()) {
      const long long n = static_cast<long long>(rows) * C / 4;  // float4 elements of the gathered [rows/4, 4C] matrix
      CC_CUDA(launch_pdl(clap_merge_gather_kernel, dim3(grid_for(n, 256)), dim3(256), 0, s,
                         reinterpret_cast<const float4*>(m->x), reinterpret_cast<float4*>(m->merge32), res, C / 4, n));
      CC_TRY(layernorm_run(m->merge32, 4 * C, st.mg, st.mb, m->ln16, 4 * C, rows / 4, 4 * C, c.eps, s));
      CC_TRY(gemm_run(st.p_merge, rows / 4, s));
      m->launches += 3;
    }
  }
  const cc_clap::Stage& last = m->stages.back();
  const int tok = last.res * last.res;
  CC_CUDA(launch_pdl(clap_ln_meanpool_kernel, dim3(B), dim3(256), static_cast<size_t>(last.C) * sizeof(float), s,
                     static_cast<const float*>(m->x), m->norm_g, m->norm_b, m->pool16, tok, last.C, c.eps));
  CC_TRY(gemm_run(m->p_proj1, B, s));
  CC_TRY(gemm_run(m->p_proj2, B, s));
  m->launches += 3;
  if (normalize) {
    CC_TRY(l2_normalize_run(m->out32, B, c.projection_dim, s));
    m->launches += 1;
  }
  CC_TRY(convert_from_f32_run(m->out32, c.projection_dim, out, out_dtype, B, c.projection_dim, s));
  m->launches += 1;
  return CC_OK;
}

int cc_clap_last_launches(cc_clap* m) { return m ? m->launches : 0; }

void cc_clap_destroy(cc_clap* m) { delete m; }

}  // extern "C"
