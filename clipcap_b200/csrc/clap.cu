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
//   reduce-add) -> LN -> fc1 GEMM with the exact (erf) GELU in its epilogue -> fc2 GEMM (residual); between stages 2x2 patch merging (gather ->
//   LN(4C) -> GEMM 4C -> 2C);  final LN + mean over the 64 tokens -> Linear + ReLU -> Linear.
// Tokens stay in image order the whole time: the cyclic shift and the window partition / reverse of the reference are
// index arithmetic inside the attention kernel (it gathers its window's rows of the QKV matrix and scatters its output
// rows), not data movement. Samples flagged is_longer (clips longer than the 10 s window) additionally run the feature
// fusion of the patch embedding: the three local mel views through a 4x12 / (4,12) convolution, then the AFF block that gates
// between the global map and the local one (modeling_clap.py:296-344, :238-245) — three fp32 kernels on the conv output.
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
    const __half* rel_bias16;  // [heads][64][64] fp16, gathered from relative_position_bias_table through the index
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
  // feature fusion of the patch embedding (samples flagged is_longer): 4x12 / (4,12) conv of the three local views and the
  // AFF block with its BatchNorms folded into the 1x1 convolutions. All fp32.
  bool has_fusion = false;
  int aff_hidden = 0, local_w = 0;  // C0 / r; output columns of the local conv per view ((S - 3p) / 3p + 1 = 21)
  const float *mel_w = nullptr, *mel_b = nullptr;                                          // [C0][p * 3p], [C0]
  const float *lw1 = nullptr, *lb1 = nullptr, *lw2 = nullptr, *lb2 = nullptr;              // local attention branch
  const float *gw1 = nullptr, *gb1 = nullptr, *gw2 = nullptr, *gb2 = nullptr;              // global attention branch
  float *local32 = nullptr, *gvec = nullptr;  // per fused sample (slot): [g0][3 * local_w][C0]; partial sums [CF_CHUNKS][C0]
  int* slots = nullptr;                       // [max_batch]: sample index of each slot (uploaded per call)
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

// Value of the folded spectrogram image at (y, x) for one [T][F] mel channel: BatchNorm over the mel bin, bicubic stretch of
// the time axis to ratio * S frames (align_corners), time chunk y / F laid along frequency (reshape_mel2img).
template <typename SRC>
__device__ __forceinline__ float mel_image_value(const SRC* __restrict__ src, int T, int F, int S, int y, int x,
                                                 const float* __restrict__ bn_scale, const float* __restrict__ bn_shift) {
  const int Tw = S * (S / F);
  const int f = y % F, t = (y / F) * S + x;
  float v;
  if (T == Tw) {
    v = static_cast<float>(src[static_cast<long long>(t) * F + f]);
  } else {
    const float scale = Tw > 1 ? static_cast<float>(T - 1) / static_cast<float>(Tw - 1) : 0.f;
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
  return v * bn_scale[f] + bn_shift[f];
}

// one thread per (sample, token): the patch x patch values of the global view (channel 0) as one row of the GEMM operand
template <typename SRC>
__global__ void clap_patches_kernel(const SRC* __restrict__ mel, long long sample_stride, int T, int F, int S, int patch,
                                    const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
                                    __half* __restrict__ cols, int B) {
  const int g = S / patch;  // tokens per side
  const long long n = static_cast<long long>(B) * g * g;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int px = static_cast<int>(i % g), py = static_cast<int>((i / g) % g);
    const SRC* src = mel + (i / (static_cast<long long>(g) * g)) * sample_stride;  // channel 0: [T][F]
    __half* dst = cols + i * (patch * patch);
    for (int ky = 0; ky < patch; ++ky)
      for (int kx = 0; kx < patch; ++kx)
        dst[ky * patch + kx] =
            __float2half_rn(mel_image_value(src, T, F, S, py * patch + ky, px * patch + kx, bn_scale, bn_shift));
  }
}

// ---------------------------------------------------------------- feature fusion of the long clips (ClapAudioPatchEmbed + AFF)
// Every kernel runs over the flagged samples of the batch at once: fused[slot] = index of the slot-th flagged sample.
// local[y][view * LW + xw][co] = bias[co] + sum_{ky,kx} w[co][ky][kx] * image_{view+1}[p y + ky][3p xw + kx]: the three local
// views through the p x 3p / (p, 3p) convolution, laid side by side along time. One CTA per (output row y, view, sample):
// the p image rows the row needs are built once in shared memory, thread = output channel with its filter in registers.
constexpr int CF_P = 4;  // patch size the fusion kernels are written for
template <typename SRC>
__global__ void clap_fusion_local_kernel(const SRC* __restrict__ mel, long long sample_stride, const int* __restrict__ fused,
                                         int T, int F, int S, int LW, int C0, const float* __restrict__ bn_scale,
                                         const float* __restrict__ bn_shift, const float* __restrict__ w,
                                         const float* __restrict__ bias, float* __restrict__ local_all) {
  extern __shared__ float rows_s[];  // [CF_P][LW * 3 * CF_P]
  constexpr int KW = 3 * CF_P, KN = CF_P * KW;
  const int y = blockIdx.x, view = blockIdx.y, slot = blockIdx.z, b = fused[slot];
  const int width = LW * KW;
  const SRC* src = mel + b * sample_stride + static_cast<long long>(view + 1) * T * F;
  for (int i = threadIdx.x; i < CF_P * width; i += blockDim.x)
    rows_s[i] = mel_image_value(src, T, F, S, y * CF_P + i / width, i % width, bn_scale, bn_shift);
  __syncthreads();
  float* local = local_all + (static_cast<long long>(slot) * (S / CF_P) + y) * (3 * LW) * C0;
  for (int co = threadIdx.x; co < C0; co += blockDim.x) {
    float wr[KN];
#pragma unroll
    for (int k = 0; k < KN; ++k) wr[k] = w[co * KN + k];
    const float bco = bias[co];
    for (int xw = 0; xw < LW; ++xw) {
      float acc = bco;
#pragma unroll
      for (int k = 0; k < KN; ++k) acc += wr[k] * rows_s[(k / KW) * width + xw * KW + (k % KW)];
      local[(view * LW + xw) * C0 + co] = acc;
    }
  }
}

// partial[slot][chunk][c] = sum over the chunk's positions of (global + local)[c]  (the mean feeding the global-attention branch)
constexpr int CF_CHUNKS = 32;
__global__ void __launch_bounds__(1024)
clap_fusion_sum_kernel(const float* __restrict__ glob_all, const float* __restrict__ local_all, const int* __restrict__ fused,
                       int g0, int LW3, int C0, float* __restrict__ partial) {
  extern __shared__ float fs[];  // [groups][C0]
  const int groups = blockDim.x / C0;
  const int c = threadIdx.x % C0, grp = threadIdx.x / C0;
  pdl_launch_dependents();
  const int slot = blockIdx.y, b = fused[slot];  // uploaded before the predecessor kernel was launched
  pdl_wait();
  const float* glob = glob_all + static_cast<long long>(b) * g0 * g0 * C0;
  const float* local = local_all + static_cast<long long>(slot) * g0 * LW3 * C0;
  const int per = (g0 * g0 + CF_CHUNKS - 1) / CF_CHUNKS;
  const int p0 = blockIdx.x * per, p1 = min(g0 * g0, p0 + per);
  float acc = 0.f;
  for (int p = p0 + grp; p < p1; p += groups) {
    const int y = p / g0, x = p - y * g0;
    acc += glob[static_cast<long long>(p) * C0 + c] + (x < LW3 ? local[(static_cast<long long>(y) * LW3 + x) * C0 + c] : 0.f);
  }
  fs[grp * C0 + c] = acc;
  __syncthreads();
  if (threadIdx.x < C0) {
    float t = 0.f;
    for (int k = 0; k < groups; ++k) t += fs[k * C0 + threadIdx.x];
    partial[(static_cast<long long>(slot) * CF_CHUNKS + blockIdx.x) * C0 + threadIdx.x] = t;
  }
}

// glob[p][c] <- 2 glob gate + 2 local (1 - gate), gate = sigmoid(local_att(glob + local)[c] + global_att(mean)[c]); thread =
// position. Every CTA first finishes the global-attention branch for its sample (sum of the 32 partials -> 96 -> hid -> 96).
constexpr int CF_MAXHID = 48;
__global__ void __launch_bounds__(128)
clap_fusion_apply_kernel(float* __restrict__ glob_all, const float* __restrict__ local_all, const int* __restrict__ fused,
                         int g0, int LW3, int C0, int hid, const float* __restrict__ w1, const float* __restrict__ b1,
                         const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ gw1,
                         const float* __restrict__ gb1, const float* __restrict__ gw2, const float* __restrict__ gb2,
                         const float* __restrict__ partial) {
  extern __shared__ float ws_[];  // w1 [hid][C0], w2 [C0][hid], b1 [hid], b2 + global branch [C0], mean [C0], ghid [hid]
  const int slot = blockIdx.y, b = fused[slot];
  float* glob = glob_all + static_cast<long long>(b) * g0 * g0 * C0;
  const float* local = local_all + static_cast<long long>(slot) * g0 * LW3 * C0;
  float* w1s = ws_;
  float* w2s = w1s + hid * C0;
  float* b1s = w2s + C0 * hid;
  float* b2s = b1s + hid;
  float* mean = b2s + C0;
  float* ghid = mean + C0;
  for (int i = threadIdx.x; i < hid * C0; i += blockDim.x) {
    w1s[i] = w1[i];
    w2s[i] = w2[i];
  }
  for (int i = threadIdx.x; i < hid; i += blockDim.x) b1s[i] = b1[i];
  pdl_launch_dependents();
  pdl_wait();
  for (int i = threadIdx.x; i < C0; i += blockDim.x) {
    float t = 0.f;
    for (int k = 0; k < CF_CHUNKS; ++k) t += partial[(static_cast<long long>(slot) * CF_CHUNKS + k) * C0 + i];
    mean[i] = t / static_cast<float>(g0 * g0);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < hid; j += blockDim.x) {
    float t = gb1[j];
    for (int k = 0; k < C0; ++k) t += gw1[j * C0 + k] * mean[k];
    ghid[j] = fmaxf(t, 0.f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C0; i += blockDim.x) {
    float t = gb2[i] + b2[i];
    for (int k = 0; k < hid; ++k) t += gw2[i * hid + k] * ghid[k];
    b2s[i] = t;
  }
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g0 * g0) return;
  const int y = p / g0, x = p - y * g0;
  float* gp = glob + static_cast<long long>(p) * C0;
  const float* lp = x < LW3 ? local + (static_cast<long long>(y) * LW3 + x) * C0 : nullptr;
  float hidden[CF_MAXHID];
#pragma unroll
  for (int j = 0; j < CF_MAXHID; ++j) hidden[j] = j < hid ? b1s[j] : 0.f;
  for (int c = 0; c < C0; ++c) {
    const float xa = gp[c] + (lp ? lp[c] : 0.f);
#pragma unroll
    for (int j = 0; j < CF_MAXHID; ++j)
      if (j < hid) hidden[j] += w1s[j * C0 + c] * xa;
  }
#pragma unroll
  for (int j = 0; j < CF_MAXHID; ++j) hidden[j] = fmaxf(hidden[j], 0.f);
  for (int c = 0; c < C0; ++c) {
    float l = b2s[c];
#pragma unroll
    for (int j = 0; j < CF_MAXHID; ++j)
      if (j < hid) l += w2s[c * hid + j] * hidden[j];
    const float gate = 1.f / (1.f + __expf(-l));
    const float xg = gp[c], xl = lp ? lp[c] : 0.f;
    gp[c] = 2.f * xg * gate + 2.f * xl * (1.f - gate);
  }
}

// ---------------------------------------------------------------- window attention (8x8 windows, head dim 24) on mma.sync
constexpr int CW_HD = 24;
// CTA = 16 warps = 4 window slots x 4 heads: a slot's 4 warps own one (window, group of 4 heads) at a time — the 96 q, 96 k
// and 96 v columns of those heads for the window's 64 tokens (36 KB of shared memory), each warp one head: S = Q K^T as
// 4 row strips of 16 x 64 (k = 24 = one m16n8k16 + one m16n8k8 step), softmax on the accumulator registers, O = P V with P
// re-used from those registers as the A operand. Slots synchronise on their own named barrier, so one slot's cp.async
// gather overlaps the others' math. The relative-position bias of the CTA's head group sits in shared memory as fp16
// for the CTA's lifetime; a CTA walks the windows of one head group.
constexpr int CW_LD = 104;       // row pitch of the q / k / v tiles in halves: 208 B, conflict-free ldmatrix
constexpr int CW_BLD = 72;       // row pitch of the bias tile in halves
constexpr int CW_SLOTS = 4;
constexpr int CW_TILE = 64 * CW_LD;  // halves per q / k / v tile
constexpr size_t CW_SMEM = static_cast<size_t>(4 * 64 * CW_BLD + CW_SLOTS * 3 * CW_TILE) * sizeof(__half) + CW_SLOTS * 64 * sizeof(int);

__device__ __forceinline__ void slot_barrier(int slot) { asm volatile("bar.sync %0, 128;" ::"r"(slot + 1) : "memory"); }

__global__ void __launch_bounds__(CW_SLOTS * 128, 1)
clap_window_attn_mma_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, const __half* __restrict__ rel_bias16,
                            int res, int shift, int C, int n_windows, int ctas_per_group, float scale_log2e) {
  extern __shared__ __align__(16) unsigned char cw_smem[];
  __half* bias_s = reinterpret_cast<__half*>(cw_smem);                     // [4 heads][64][CW_BLD]
  __half* tiles = bias_s + 4 * 64 * CW_BLD;                                // [slot][q, k, v][64][CW_LD]
  int* region_s = reinterpret_cast<int*>(tiles + CW_SLOTS * 3 * CW_TILE);  // [slot][64]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slot = warp >> 2, head_l = warp & 3, st = tid & 127;  // st: thread index inside the slot
  const int group = blockIdx.x / ctas_per_group, cta_in_group = blockIdx.x - group * ctas_per_group;
  const int wpr = res >> 3, wps = wpr * wpr;  // windows per row / per sample
  // bias of this head group (constant weights: may be read before the predecessor kernel has finished)
  for (int i = tid; i < 4 * 64 * 8; i += blockDim.x) {  // 16-byte pieces: [head][row][8 pieces]
    const int piece = i & 7, row = (i >> 3) & 63, h = i >> 9;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(rel_bias16 + ((static_cast<long long>(group) * 4 + h) * 64 + row) * 64) + piece);
    *reinterpret_cast<uint4*>(bias_s + (h * 64 + row) * CW_BLD + piece * 8) = v;
  }
  pdl_launch_dependents();
  pdl_wait();
  __syncthreads();
  __half* qs = tiles + slot * 3 * CW_TILE;
  __half* ks = qs + CW_TILE;
  __half* vs = ks + CW_TILE;
  int* regs = region_s + slot * 64;
  const uint32_t qs_u = smem_u32(qs), ks_u = smem_u32(ks), vs_u = smem_u32(vs);
  const int g = lane >> 2, q4 = lane & 3;
  const __half* bias_h = bias_s + head_l * 64 * CW_BLD;
  const float kLog2e = 1.4426950408889634f;

  for (int w = cta_in_group * CW_SLOTS + slot; w < n_windows; w += ctas_per_group * CW_SLOTS) {
    const long long b = w / wps;
    const int win = w - static_cast<int>(b) * wps;
    const int wy = win / wpr, wx = win - wy * wpr;
    const long long sample_row0 = b * res * res;
    slot_barrier(slot);  // the previous window's output tile has been stored
    // ---- gather: 64 tokens x {q, k, v} x 12 pieces of 16 B
#pragma unroll 3
    for (int m = 0; m < 18; ++m) {
      const int k = st + 128 * m;
      const int tok = k / 36, rem = k - tok * 36, which = rem / 12, piece = rem - which * 12;
      int oy = wy * 8 + (tok >> 3) + shift, ox = wx * 8 + (tok & 7) + shift;
      oy -= oy >= res ? res : 0;
      ox -= ox >= res ? res : 0;
      const __half* src = qkv + (sample_row0 + static_cast<long long>(oy) * res + ox) * (3LL * C) + which * C + group * 96 + piece * 8;
      cp_async_16(qs_u + static_cast<uint32_t>((which * CW_TILE + tok * CW_LD + piece * 8) * 2), src, true);
    }
    cp_async_commit();
    if (st < 64) {
      int region = 0;
      if (shift > 0) {
        const int sy = wy * 8 + (st >> 3), sx = wx * 8 + (st & 7);
        const int ry = sy < res - 8 ? 0 : (sy < res - shift ? 1 : 2);
        const int rx = sx < res - 8 ? 0 : (sx < res - shift ? 1 : 2);
        region = ry * 3 + rx;
      }
      regs[st] = region;
    }
    cp_async_wait<0>();
    slot_barrier(slot);
    // ---- this warp's head: the K fragments stay in registers for the four row strips
    const int hc = head_l * CW_HD;  // first column of the head inside the tiles
    uint32_t kf[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int mi = lane >> 3;
      ldmatrix_x4(kf[nt], ks_u + static_cast<uint32_t>(((nt * 8 + (lane & 7)) * CW_LD + hc + (mi < 2 ? mi : 2) * 8) * 2));
    }
#pragma unroll 1
    for (int mt = 0; mt < 4; ++mt) {
      uint32_t qa[4], qb[2];
      {
        const int mi = lane >> 3;
        ldmatrix_x4(qa, qs_u + static_cast<uint32_t>(((mt * 16 + (lane & 7) + (mi & 1) * 8) * CW_LD + hc + (mi >> 1) * 8) * 2));
        ldmatrix_x2(qb, qs_u + static_cast<uint32_t>(((mt * 16 + (lane & 15)) * CW_LD + hc + 16) * 2));
      }
      float sacc[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        sacc[nt][0] = sacc[nt][1] = sacc[nt][2] = sacc[nt][3] = 0.f;
        mma_16816(sacc[nt], qa, kf[nt][0], kf[nt][1]);
        mma_1688(sacc[nt], qb, kf[nt][2]);
      }
      // scores in the log2 domain: (q.k * scale + bias + mask) * log2(e); rows r0 = mt*16 + g and r0 + 8
      const int r0 = mt * 16 + g;
      const int reg_r0 = regs[r0], reg_r1 = regs[r0 + 8];
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float2 b0 = __half22float2(*reinterpret_cast<const __half2*>(bias_h + r0 * CW_BLD + nt * 8 + 2 * q4));
        const float2 b1 = __half22float2(*reinterpret_cast<const __half2*>(bias_h + (r0 + 8) * CW_BLD + nt * 8 + 2 * q4));
        const int2 rc = *reinterpret_cast<const int2*>(regs + nt * 8 + 2 * q4);  // regions of this thread's two columns
        sacc[nt][0] = sacc[nt][0] * scale_log2e + (b0.x + (rc.x != reg_r0 ? -100.f : 0.f)) * kLog2e;
        sacc[nt][1] = sacc[nt][1] * scale_log2e + (b0.y + (rc.y != reg_r0 ? -100.f : 0.f)) * kLog2e;
        sacc[nt][2] = sacc[nt][2] * scale_log2e + (b1.x + (rc.x != reg_r1 ? -100.f : 0.f)) * kLog2e;
        sacc[nt][3] = sacc[nt][3] * scale_log2e + (b1.y + (rc.y != reg_r1 ? -100.f : 0.f)) * kLog2e;
        mx0 = fmaxf(mx0, fmaxf(sacc[nt][0], sacc[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(sacc[nt][2], sacc[nt][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      float sum0 = 0.f, sum1 = 0.f;
      uint32_t pa[4][4];  // P as the A operand of the four 16-token k-steps
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float p0 = fast_exp2(sacc[nt][0] - mx0), p1 = fast_exp2(sacc[nt][1] - mx0);
        const float p2 = fast_exp2(sacc[nt][2] - mx1), p3 = fast_exp2(sacc[nt][3] - mx1);
        sum0 += p0 + p1;
        sum1 += p2 + p3;
        pa[nt >> 1][(nt & 1) * 2] = pack_half2(p0, p1);
        pa[nt >> 1][(nt & 1) * 2 + 1] = pack_half2(p2, p3);
      }
      sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
      sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
      sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
      sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
      float oacc[3][4];
#pragma unroll
      for (int n3 = 0; n3 < 3; ++n3) oacc[n3][0] = oacc[n3][1] = oacc[n3][2] = oacc[n3][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {  // V fragments of 16 tokens: (tokens, hd 0-7 | 8-15) transposed x4, hd 16-23 x2
        const int mi = lane >> 3;
        uint32_t t4[4], t2[2];
        ldmatrix_x4_trans(t4, vs_u + static_cast<uint32_t>(((kk * 16 + (lane & 7) + (mi & 1) * 8) * CW_LD + hc + (mi >> 1) * 8) * 2));
        ldmatrix_x2_trans(t2, vs_u + static_cast<uint32_t>(((kk * 16 + (lane & 15)) * CW_LD + hc + 16) * 2));
        mma_16816(oacc[0], pa[kk], t4[0], t4[1]);
        mma_16816(oacc[1], pa[kk], t4[2], t4[3]);
        mma_16816(oacc[2], pa[kk], t2[0], t2[1]);
      }
      // O over this head's q columns of the strip (its fragments are in registers already; no other warp reads them)
      const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;
      __syncwarp();
#pragma unroll
      for (int n3 = 0; n3 < 3; ++n3) {
        *reinterpret_cast<uint32_t*>(qs + r0 * CW_LD + hc + n3 * 8 + 2 * q4) = pack_half2(oacc[n3][0] * inv0, oacc[n3][1] * inv0);
        *reinterpret_cast<uint32_t*>(qs + (r0 + 8) * CW_LD + hc + n3 * 8 + 2 * q4) = pack_half2(oacc[n3][2] * inv1, oacc[n3][3] * inv1);
      }
    }
    slot_barrier(slot);
    // ---- scatter the 64 x 96 output tile: 12 pieces of 16 B per token
#pragma unroll
    for (int m = 0; m < 6; ++m) {
      const int k = st + 128 * m;
      const int tok = k / 12, piece = k - tok * 12;
      int oy = wy * 8 + (tok >> 3) + shift, ox = wx * 8 + (tok & 7) + shift;
      oy -= oy >= res ? res : 0;
      ox -= ox >= res ? res : 0;
      const uint4 v = *reinterpret_cast<const uint4*>(qs + tok * CW_LD + piece * 8);
      *reinterpret_cast<uint4*>(out + (sample_row0 + static_cast<long long>(oy) * res + ox) * C + group * 96 + piece * 8) = v;
    }
  }
}

// ---------------------------------------------------------------- patch merging gather, final LN + mean pool
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

// pooled[b, :] = mean over the `tokens` rows of LayerNorm(x[b, t, :])  (fp32 statistics and mean; one CTA per sample). Every
// warp sums its tokens (t = warp, warp + 8, ...) into its own row of shared memory, the rows are then added in warp order:
// the result does not depend on scheduling.
constexpr int CP_WARPS = 8;
__global__ void __launch_bounds__(CP_WARPS * 32)
clap_ln_meanpool_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ bta,
                        __half* __restrict__ pooled, int tokens, int C, float eps) {
  extern __shared__ float acc[];  // [CP_WARPS][C]
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* mine = acc + warp * C;
  for (int c = lane; c < C; c += 32) mine[c] = 0.f;
  for (int t = warp; t < tokens; t += CP_WARPS) {
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
    for (int c = lane; c < C; c += 32) mine[c] += (row[c] - mean) * rstd * g[c] + bta[c];  // channel c belongs to this lane
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < CP_WARPS; ++w) t += acc[w * C + c];
    pooled[static_cast<long long>(b) * C + c] = __float2half_rn(t / tokens);
  }
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
  CC_REQUIRE(static_cast<size_t>(CP_WARPS) * Cl * sizeof(float) <= 48 * 1024, CC_ESHAPE, "clap: final width %d too large for the pool kernel", Cl);
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
  // feature fusion (present when the checkpoint was built with enable_fusion): fold every eval-mode BatchNorm into the
  // 1x1 convolution in front of it, on the host (a few thousand values)
  {
    const std::string pe = enc + "patch_embed.";
    bool present = false;
    int64_t hid_numel = 0;
    for (int i = 0; i < nw; ++i)
      if (w[i].name != nullptr && pe + "fusion_model.local_att.0.weight" == w[i].name) {
        present = true;
        hid_numel = 1;
        for (int k = 0; k < w[i].ndim; ++k) hid_numel *= w[i].shape[k];
      }
    if (present) {
      const int hid = static_cast<int>(hid_numel / C0);
      CC_REQUIRE(hid >= 1 && hid <= CF_MAXHID && static_cast<int64_t>(hid) * C0 == hid_numel, CC_ESHAPE,
                 "clap: AFF hidden width %d (max %d)", hid, CF_MAXHID);
      CC_REQUIRE(S >= 3 * c.patch && c.patch == CF_P && C0 <= 1024, CC_ESHAPE,
                 "clap: the fusion kernels take patch %d (got %d), spec_size %d, embed %d", CF_P, c.patch, S, C0);
      m->aff_hidden = hid;
      m->local_w = (S - 3 * c.patch) / (3 * c.patch) + 1;
      auto fetch = [&](const std::string& name, int64_t numel, std::vector<float>& out) -> int {
        const float* d = nullptr;
        CC_TRY(get(w, nw, name, numel, stage, &d));
        out.resize(static_cast<size_t>(numel));
        CC_CUDA(cudaMemcpy(out.data(), d, out.size() * sizeof(float), cudaMemcpyDeviceToHost));
        return CC_OK;
      };
      auto upload = [&](const std::vector<float>& v, const float** out) -> int {
        float* d = nullptr;
        CC_TRY(A.alloc_t(&d, v.size()));
        CC_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
        *out = d;
        return CC_OK;
      };
      // conv [out, in, 1, 1] + BatchNorm(out)  ->  W' = s W, b' = s (b - mean) + beta, s = gamma / sqrt(var + 1e-5)
      auto folded = [&](const std::string& conv, const std::string& bn, int out_ch, int in_ch, const float** wd,
                        const float** bd) -> int {
        std::vector<float> cw, cb, g, b, rm, rv;
        CC_TRY(fetch(conv + "weight", static_cast<int64_t>(out_ch) * in_ch, cw));
        CC_TRY(fetch(conv + "bias", out_ch, cb));
        CC_TRY(fetch(bn + "weight", out_ch, g));
        CC_TRY(fetch(bn + "bias", out_ch, b));
        CC_TRY(fetch(bn + "running_mean", out_ch, rm));
        CC_TRY(fetch(bn + "running_var", out_ch, rv));
        for (int o = 0; o < out_ch; ++o) {
          const float sc = g[o] / sqrtf(rv[o] + 1e-5f);
          for (int i = 0; i < in_ch; ++i) cw[static_cast<size_t>(o) * in_ch + i] *= sc;
          cb[o] = sc * (cb[o] - rm[o]) + b[o];
        }
        CC_TRY(upload(cw, wd));
        CC_TRY(upload(cb, bd));
        return CC_OK;
      };
      const std::string fm = pe + "fusion_model.";
      CC_TRY(folded(fm + "local_att.0.", fm + "local_att.1.", hid, C0, &m->lw1, &m->lb1));
      CC_TRY(folded(fm + "local_att.3.", fm + "local_att.4.", C0, hid, &m->lw2, &m->lb2));
      CC_TRY(folded(fm + "global_att.1.", fm + "global_att.2.", hid, C0, &m->gw1, &m->gb1));
      CC_TRY(folded(fm + "global_att.4.", fm + "global_att.5.", C0, hid, &m->gw2, &m->gb2));
      const float *mw, *mb;
      CC_TRY(get(w, nw, pe + "mel_conv2d.weight", static_cast<int64_t>(C0) * pp * 3, stage, &mw));
      CC_TRY(get(w, nw, pe + "mel_conv2d.bias", C0, stage, &mb));
      CC_TRY(keep_f32(A, mw, static_cast<size_t>(C0) * pp * 3, &m->mel_w));
      CC_TRY(keep_f32(A, mb, C0, &m->mel_b));
      CC_TRY(A.alloc_t(&m->local32, static_cast<size_t>(B) * m->grid0 * 3 * m->local_w * C0));
      CC_TRY(A.alloc_t(&m->gvec, static_cast<size_t>(B) * CF_CHUNKS * C0));
      CC_TRY(A.alloc_t(&m->slots, static_cast<size_t>(B)));
      CC_REQUIRE(3 * m->local_w <= m->grid0, CC_ESHAPE, "clap: %d local columns exceed the %d global ones", 3 * m->local_w, m->grid0);
      m->has_fusion = true;
      stage.release();
    }
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
    CC_REQUIRE(st.heads % 4 == 0, CC_ESHAPE, "clap: stage %d has %d heads (kernels take groups of 4)", si, st.heads);
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
        std::vector<__half> bias16(bias.size());
        for (size_t i = 0; i < bias.size(); ++i) bias16[i] = __float2half(bias[i]);
        __half* d16 = nullptr;
        CC_TRY(A.alloc_t(&d16, bias16.size()));
        CC_CUDA(cudaMemcpy(d16, bias16.data(), bias16.size() * sizeof(__half), cudaMemcpyHostToDevice));
        bk.rel_bias16 = d16;
      }
      CC_TRY(gemm_plan(&bk.p_qkv, m->ln16, C, rows, bk.wqkv, 3 * C, C, EPI_F16_NONE, bk.bqkv, m->qkv16, 3 * C));
      CC_TRY(gemm_plan(&bk.p_o, m->att16, C, rows, bk.wo, C, C, EPI_RESID_F32, bk.bo, m->x, C));
      CC_TRY(gemm_plan(&bk.p_1, m->ln16, C, rows, bk.w1, 4 * C, C, EPI_F16_GELU_ERF, bk.b1, m->mlp16, 4 * C));
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
  // per device: the handle's device is the one current now
  CC_CUDA(cudaFuncSetAttribute(clap_window_attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(CW_SMEM)));
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

int cc_clap_forward(cc_clap* m, const void* mel, int mel_dtype, const unsigned char* is_longer, int B, int channels, int T,
                    int normalize, void* out, int out_dtype, int stop_after_stage, float* dump, void* stream) {
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
  int n_fused = 0;
  if (is_longer != nullptr)
    for (int b = 0; b < B; ++b) n_fused += is_longer[b] != 0;
  if (n_fused > 0) {  // modeling_clap.py:310-338: the flagged samples' global map is fused with their local views
    CC_REQUIRE(m->has_fusion, CC_EINVAL, "cc_clap_forward: %d samples are flagged is_longer but the handle has no fusion weights", n_fused);
    CC_REQUIRE(channels == 4, CC_ESHAPE, "cc_clap_forward: feature fusion needs 4 mel views per sample, got %d", channels);
    const int C0 = c.embed, LW = m->local_w, hid = m->aff_hidden;
    std::vector<int> slots;  // slot -> sample
    for (int b = 0; b < B; ++b)
      if (is_longer[b]) slots.push_back(b);
    // pageable source: the copy is staged before the call returns, so the vector may die with this scope
    CC_CUDA(cudaMemcpyAsync(m->slots, slots.data(), slots.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    const size_t psm = static_cast<size_t>(CF_P) * LW * 3 * CF_P * sizeof(float);
    const dim3 lgrid(g0, 3, n_fused);
    if (mel_dtype == CC_F32)
      clap_fusion_local_kernel<float><<<lgrid, 96, psm, s>>>(static_cast<const float*>(mel), sample_stride, m->slots, T, F, S, LW, C0,
                                                             m->bn_scale, m->bn_shift, m->mel_w, m->mel_b, m->local32);
    else
      clap_fusion_local_kernel<__half><<<lgrid, 96, psm, s>>>(static_cast<const __half*>(mel), sample_stride, m->slots, T, F, S, LW, C0,
                                                              m->bn_scale, m->bn_shift, m->mel_w, m->mel_b, m->local32);
    CC_CUDA(cudaGetLastError());
    const int gthreads = (1024 / C0) * C0 > 0 ? (1024 / C0) * C0 : C0;
    CC_CUDA(launch_pdl(clap_fusion_sum_kernel, dim3(CF_CHUNKS, n_fused), dim3(gthreads), static_cast<size_t>(gthreads) * sizeof(float), s,
                       static_cast<const float*>(m->merge32), static_cast<const float*>(m->local32),
                       static_cast<const int*>(m->slots), g0, 3 * LW, C0, m->gvec));
    CC_CUDA(launch_pdl(clap_fusion_apply_kernel, dim3((g0 * g0 + 127) / 128, n_fused), dim3(128),
                       static_cast<size_t>(2 * hid * C0 + 2 * hid + 2 * C0) * sizeof(float), s, m->merge32,
                       static_cast<const float*>(m->local32), static_cast<const int*>(m->slots), g0, 3 * LW, C0, hid, m->lw1,
                       m->lb1, m->lw2, m->lb2, m->gw1, m->gb1, m->gw2, m->gb2, static_cast<const float*>(m->gvec)));
    m->launches += 3;
  }
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
      {  // one CTA walks the windows of one group of 4 heads; about one CTA per SM over all groups
        const int groups = st.heads / 4;
        int per = num_sms() / groups;
        per = per < 1 ? 1 : per;
        per = per > (windows + CW_SLOTS - 1) / CW_SLOTS ? (windows + CW_SLOTS - 1) / CW_SLOTS : per;
        CC_CUDA(launch_pdl(clap_window_attn_mma_kernel, dim3(groups * per), dim3(CW_SLOTS * 128), CW_SMEM, s,
                           static_cast<const __half*>(m->qkv16), m->att16, bk.rel_bias16, res, shift, C, windows, per,
                           1.4426950408889634f / sqrtf(static_cast<float>(CW_HD))));
      }
      CC_TRY(gemm_run(bk.p_o, rows, s));
      CC_TRY(layernorm_run(m->x, C, bk.ln2_g, bk.ln2_b, m->ln16, C, rows, C, c.eps, s));
      CC_TRY(gemm_run(bk.p_1, rows, s));
      CC_TRY(gemm_run(bk.p_2, rows, s));
      m->launches += 7;
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
  CC_CUDA(launch_pdl(clap_ln_meanpool_kernel, dim3(B), dim3(CP_WARPS * 32), static_cast<size_t>(CP_WARPS) * last.C * sizeof(float), s,
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
