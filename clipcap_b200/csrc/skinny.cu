// Skinny GEMM for the single-image / small-beam decode step:  y[M, N] = LN?(x)[M, K] * W[N, K]^T (+ epilogue), M <= 16.
//
// With at most 16 rows a decode step is pure weight streaming (GPT-2-medium: 707 MB per step against 0.7 GFLOP per row),
// and the 128-row tcgen05 tile spends its time on set-up instead (TMEM allocation, tensor-map fetch, a 320-thread CTA
// per 128 x 64 tile whose 127 other rows are padding): 37.7 us per layer at one row. Here
//   * a CTA owns 16 (or 8) output columns at a time and walks column tiles grid-stride; its 8 warps split K, every lane
//     pulls 16 contiguous bytes of a weight row per step (64 B per row and quarter-warp: whole sectors), with all of a
//     warp's loads of a batch in flight before the first use — and the first batch is issued BEFORE griddepcontrol.wait
//     (weights never depend on the previous kernel);
//   * the M activation rows live in shared memory as fp16 (64-byte row pad: conflict-free 16-byte reads); the LayerNorm
//     that precedes QKV / fc1 / the LM head is applied while they are staged (every CTA normalises the <= 16 rows itself:
//     64 KB of L2 reads instead of a kernel boundary);
//   * the products run on warp MMA (m16n8k16, rows M..15 are zero registers); because every lane holds the same 8
//     consecutive k of its A row and its W row, the k positions inside a 32-wide step are a fixed permutation on both
//     operands and no ldmatrix / transposition is needed;
//   * the 8 partial sums of an output element are added in warp order (deterministic); epilogues: fp16 (+gelu_new), fp32
//     residual update in place, fp32 logits, fused argmax keys (same packing as the tcgen05 EPI_ARGMAX).
#include <algorithm>
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"

namespace cc {
namespace {

constexpr int SK_THREADS = 256;
constexpr int SK_WARPS = 8;
constexpr int SK_PAD = 32;   // halves of padding per activation row (64 bytes)
constexpr int SK_BATCH = 4;   // k-steps (32 halves each) per unit; two units' weight loads are in flight per warp
constexpr int SK_LN_V4 = 16;  // float4 per lane of a LayerNorm row: K <= 2048

__device__ __forceinline__ float sk_gelu_new(float x) {  // same form as the tcgen05 epilogue (gemm.cu act_gelu_new)
  const float u2 = (2.f * 0.7978845608028654f * 1.4426950408889634f) * (x + 0.044715f * x * x * x);
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-u2));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return x * r;
}
__device__ __forceinline__ uint32_t sk_order_key(float x) {
  const uint32_t b = __float_as_uint(x);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float sk_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct SkArgs {
  const float* x32;  // LayerNorm input rows (fp32, stride ldx) or nullptr
  const float* gamma;
  const float* beta;
  const __half* x16;  // plain fp16 input rows (stride ldx) when x32 == nullptr
  long long ldx;
  const __half* w;  // [N, K]
  const float* bias;
  void* out;
  long long ldc;
  int M, N, K;
  float eps;
};

// NT = column tiles of 8 per CTA step (1 or 2), ROWS16 = rows 8..15 exist
template <int NT, int EPI, bool ROWS16>
__global__ void __launch_bounds__(SK_THREADS, 2)
skinny_gemm_kernel(SkArgs a) {
  extern __shared__ __align__(16) uint8_t sk_smem[];
  constexpr int COLS = 8 * NT;
  constexpr int XROWS = ROWS16 ? 16 : 8;
  const int pitch = a.K + SK_PAD;  // halves
  __half* xs = reinterpret_cast<__half*>(sk_smem);
  float* red = reinterpret_cast<float*>(sk_smem + static_cast<size_t>(XROWS) * pitch * 2);  // [SK_WARPS][16][COLS]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int n_tiles = (a.N + COLS - 1) / COLS;
  const int steps = a.K >> 5;  // k-steps of 32 halves; warp w takes steps w, w + 8, ...
  const int my_steps = (steps - warp + SK_WARPS - 1) / SK_WARPS;
  // Work of this CTA = its column tiles (grid-stride) x batches of SK_BATCH k-steps, as one flat sequence of units; the
  // weight loads of unit u + 1 are issued before unit u is consumed (two register buffers).
  const int nb = (((steps + SK_WARPS - 1) / SK_WARPS) + SK_BATCH - 1) / SK_BATCH;  // batches per tile (same for all warps)
  const int my_tiles = static_cast<int>(blockIdx.x) < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const int units = my_tiles * nb;

  uint4 wa[SK_BATCH][NT], wb[SK_BATCH][NT];
  auto load_w = [&](uint4 (&wreg)[SK_BATCH][NT], int u) {
    const int tile = blockIdx.x + (u / nb) * gridDim.x;
    const int first = (u % nb) * SK_BATCH;
#pragma unroll
    for (int i = 0; i < SK_BATCH; ++i) {
      const int st = first + i;
      if (st < my_steps) {
        const int k = ((warp + st * SK_WARPS) << 5) + 8 * t;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          int n = tile * COLS + nt * 8 + g;
          n = n < a.N ? n : a.N - 1;  // clamp: columns past N are computed on the last row and dropped
          wreg[i][nt] = __ldg(reinterpret_cast<const uint4*>(a.w + static_cast<long long>(n) * a.K + k));
        }
      }
    }
  };
  if (units > 0) load_w(wa, 0);  // weights do not depend on the previous kernel
  pdl_launch_dependents();
  pdl_wait();

  // ---- stage the activation rows (LayerNorm applied on the way in: warp per row, the row held in registers)
  if (a.x32 != nullptr) {
    const int nv = a.K >> 2;
    for (int m = warp; m < XROWS; m += SK_WARPS) {
      __half* dst = xs + m * pitch;
      if (m < a.M) {
        const float4* row = reinterpret_cast<const float4*>(a.x32 + m * a.ldx);
        float4 v[SK_LN_V4];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < SK_LN_V4; ++i) {
          const int c = lane + 32 * i;
          v[i] = c < nv ? row[c] : make_float4(0.f, 0.f, 0.f, 0.f);
          sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
        const float mean = sk_warp_sum(sum) / static_cast<float>(a.K);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < SK_LN_V4; ++i) {
          if (lane + 32 * i < nv) {
            const float d0 = v[i].x - mean, d1 = v[i].y - mean, d2 = v[i].z - mean, d3 = v[i].w - mean;
            q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
          }
        }
        const float rstd = rsqrtf(sk_warp_sum(q) / static_cast<float>(a.K) + a.eps);
#pragma unroll
        for (int i = 0; i < SK_LN_V4; ++i) {
          const int c = lane + 32 * i;
          if (c < nv) {
            const float4 gm = __ldg(reinterpret_cast<const float4*>(a.gamma) + c);
            const float4 bt = __ldg(reinterpret_cast<const float4*>(a.beta) + c);
            uint2 o;
            o.x = pack_half2((v[i].x - mean) * rstd * gm.x + bt.x, (v[i].y - mean) * rstd * gm.y + bt.y);
            o.y = pack_half2((v[i].z - mean) * rstd * gm.z + bt.z, (v[i].w - mean) * rstd * gm.w + bt.w);
            *reinterpret_cast<uint2*>(dst + 4 * c) = o;
          }
        }
      } else {
        for (int c = lane; c < nv; c += 32) *reinterpret_cast<uint2*>(dst + 4 * c) = make_uint2(0u, 0u);
      }
    }
  } else {
    const int cpr = a.K >> 3;  // 16-byte chunks per row
    for (int i = threadIdx.x; i < XROWS * cpr; i += SK_THREADS) {
      const int m = i / cpr, c = i - m * cpr;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (m < a.M) v = *reinterpret_cast<const uint4*>(a.x16 + m * a.ldx + 8 * c);
      *reinterpret_cast<uint4*>(xs + m * pitch + 8 * c) = v;
    }
  }
  __syncthreads();

  float acc[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
  unsigned long long best_key = 0ull;  // EPI_ARGMAX: this thread's best (value, column) over the CTA's tiles

  auto consume = [&](const uint4 (&wreg)[SK_BATCH][NT], int u) {
    const int first = (u % nb) * SK_BATCH;
#pragma unroll
    for (int i = 0; i < SK_BATCH; ++i) {
      if (first + i < my_steps) {
        const int k = ((warp + (first + i) * SK_WARPS) << 5) + 8 * t;
        const uint4 xa = *reinterpret_cast<const uint4*>(xs + g * pitch + k);
        uint4 xb = make_uint4(0u, 0u, 0u, 0u);
        if constexpr (ROWS16) xb = *reinterpret_cast<const uint4*>(xs + (g + 8) * pitch + k);
        const uint32_t a1[4] = {xa.x, xb.x, xa.y, xb.y};  // k slots {2t,2t+1 | 2t+8,2t+9} <- halves 0..3 of the lane's 8
        const uint32_t a2[4] = {xa.z, xb.z, xa.w, xb.w};  // halves 4..7
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          mma_16816(acc[nt], a1, wreg[i][nt].x, wreg[i][nt].y);
          mma_16816(acc[nt], a2, wreg[i][nt].z, wreg[i][nt].w);
        }
      }
    }
    if (u % nb != nb - 1) return;  // tile not finished yet
    // ---- cross-warp reduction (fixed order) + epilogue of the finished tile
    const int tile = blockIdx.x + (u / nb) * gridDim.x;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      float* r = red + (warp * 16) * COLS + nt * 8 + 2 * t;
      r[g * COLS] = acc[nt][0];
      r[g * COLS + 1] = acc[nt][1];
      r[(g + 8) * COLS] = acc[nt][2];
      r[(g + 8) * COLS + 1] = acc[nt][3];
      acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
    }
    __syncthreads();
    if (threadIdx.x < 16 * COLS) {
      const int m = threadIdx.x / COLS, nl = threadIdx.x - m * COLS;
      const int n = tile * COLS + nl;
      float v = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < SK_WARPS; ++w8) v += red[(w8 * 16 + m) * COLS + nl];
      const bool ok = m < a.M && n < a.N;
      if (ok && a.bias != nullptr) v += __ldg(a.bias + n);
      if constexpr (EPI == EPI_F16_NONE) {
        if (ok) reinterpret_cast<__half*>(a.out)[m * a.ldc + n] = __float2half_rn(v);
      } else if constexpr (EPI == EPI_F16_GELU_NEW) {
        if (ok) reinterpret_cast<__half*>(a.out)[m * a.ldc + n] = __float2half_rn(sk_gelu_new(v));
      } else if constexpr (EPI == EPI_RESID_F32) {
        if (ok) reinterpret_cast<float*>(a.out)[m * a.ldc + n] += v;
      } else if constexpr (EPI == EPI_F32) {
        if (ok) reinterpret_cast<float*>(a.out)[m * a.ldc + n] = v;
      } else {  // EPI_ARGMAX
        if (ok) {
          const unsigned long long key = (static_cast<unsigned long long>(sk_order_key(v)) << 32) |
                                         static_cast<unsigned long long>(~static_cast<uint32_t>(n));
          best_key = key > best_key ? key : best_key;
        }
      }
    }
    __syncthreads();  // `red` is rewritten by the next tile
  };

  for (int u = 0; u < units; u += 2) {
    if (u + 1 < units) load_w(wb, u + 1);
    consume(wa, u);
    if (u + 1 < units) {
      if (u + 2 < units) load_w(wa, u + 2);
      consume(wb, u + 1);
    }
  }
  if constexpr (EPI == EPI_ARGMAX) {
    // one atomic per (CTA, row): COLS consecutive lanes hold one row's candidates
    if (threadIdx.x < 16 * COLS) {
      const int m = threadIdx.x / COLS, nl = threadIdx.x - m * COLS;
#pragma unroll
      for (int o = COLS / 2; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best_key, o);
        best_key = other > best_key ? other : best_key;
      }
      if (nl == 0 && m < a.M && best_key != 0ull) atomicMax(reinterpret_cast<unsigned long long*>(a.out) + m, best_key);
    }
  }
}

template <int NT, int EPI, bool ROWS16>
int sk_launch(const SkArgs& a, cudaStream_t s) {
  constexpr int COLS = 8 * NT;
  const size_t smem = static_cast<size_t>(ROWS16 ? 16 : 8) * (a.K + SK_PAD) * 2 + static_cast<size_t>(SK_WARPS) * 16 * COLS * 4;
  CC_REQUIRE(smem <= 227 * 1024, CC_ESHAPE, "skinny gemm: K=%d needs %zu bytes of shared memory", a.K, smem);
  auto kern = skinny_gemm_kernel<NT, EPI, ROWS16>;
  CC_OPT_IN_SMEM(kern, 227 * 1024);
  const int n_tiles = (a.N + COLS - 1) / COLS;
  int per_sm = static_cast<int>(std::min<size_t>(4, (200 * 1024) / smem));
  if (EPI == EPI_ARGMAX) per_sm = std::min(per_sm, 2);  // every CTA ends with one atomicMax per row on the same few words
  const int grid = std::min(n_tiles, num_sms() * std::max(1, per_sm));
  CC_CUDA(launch_pdl(kern, dim3(grid), dim3(SK_THREADS), smem, s, a));
  return CC_OK;
}

template <int EPI>
int sk_dispatch(const SkArgs& a, cudaStream_t s) {
  // 16-column tiles unless that leaves SMs without a tile
  const bool wide = (a.N + 15) / 16 >= num_sms();
  if (a.M > 8) return wide ? sk_launch<2, EPI, true>(a, s) : sk_launch<1, EPI, true>(a, s);
  return wide ? sk_launch<2, EPI, false>(a, s) : sk_launch<1, EPI, false>(a, s);
}

}  // namespace

bool skinny_enabled() {
  static const bool on = [] {
    const char* e = getenv("CLIPCAP_B200_NO_SKINNY");
    return !(e != nullptr && e[0] == '1');
  }();
  return on;
}

int skinny_gemm_run(const float* x32, const float* gamma, const float* beta, float eps, const __half* x16, int64_t ldx,
                    int M, const __half* w, int N, int K, int epi, const float* bias, void* out, int64_t ldc,
                    cudaStream_t s) {
  CC_REQUIRE(M >= 1 && M <= kSkinnyMaxRows && N >= 1 && K >= 32 && K % 32 == 0, CC_ESHAPE,
             "skinny gemm: M=%d (max %d) N=%d K=%d (multiple of 32)", M, kSkinnyMaxRows, N, K);
  CC_REQUIRE((x32 != nullptr) != (x16 != nullptr), CC_EINVAL, "skinny gemm: exactly one of the fp32 (LayerNorm) and fp16 inputs");
  CC_REQUIRE(x32 == nullptr || (gamma != nullptr && beta != nullptr && ldx % 4 == 0 && K <= 128 * SK_LN_V4), CC_EINVAL,
             "skinny gemm: LayerNorm input needs gamma, beta, a row stride that is a multiple of 4 and K <= %d", 128 * SK_LN_V4);
  CC_REQUIRE(x16 == nullptr || ldx % 8 == 0, CC_EALIGN, "skinny gemm: fp16 input row stride must be a multiple of 8");
  SkArgs a{x32, gamma, beta, x16, static_cast<long long>(ldx), w, bias, out, static_cast<long long>(ldc), M, N, K, eps};
  switch (epi) {
    case EPI_F16_NONE: return sk_dispatch<EPI_F16_NONE>(a, s);
    case EPI_F16_GELU_NEW: return sk_dispatch<EPI_F16_GELU_NEW>(a, s);
    case EPI_RESID_F32: return sk_dispatch<EPI_RESID_F32>(a, s);
    case EPI_F32: return sk_dispatch<EPI_F32>(a, s);
    case EPI_ARGMAX: return sk_dispatch<EPI_ARGMAX>(a, s);
  }
  set_error("skinny gemm: epilogue %d not supported", epi);
  return CC_EINVAL;
}

}  // namespace cc
