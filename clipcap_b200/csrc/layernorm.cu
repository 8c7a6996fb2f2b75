// LayerNorm over the fp32 residual stream, emitting the fp16 GEMM operand.  HBM-bound: 4 B read + 2 B written per
// element. One warp per row, the row is held in registers (float4 per lane per 128 columns), statistics are two-pass
// fp32 (mean, then centred variance) like torch.nn.LayerNorm / OpenAI CLIP's fp32 LayerNorm.
#include "common.h"
#include "ptx.cuh"

namespace cc {
namespace {

constexpr int LN_WARPS = 4;
constexpr int LN_MAX_V4 = 16;  // up to 16 float4 per lane => d <= 2048

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// REDUCE: the row first absorbs the split-K partial sums of the preceding GEMM (decode path):
//   x[r,:] += bias + partial[0][r,:] + partial[1][r,:] + ...   in that fixed order, written back to x.
template <int NV4, bool REDUCE>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_kernel(float* __restrict__ x, long long x_ld, const float* __restrict__ partial, int splits,
                 long long split_stride, const float* __restrict__ bias, const float* __restrict__ gamma,
                 const float* __restrict__ beta, __half* __restrict__ y, long long y_ld, int rows, int d, float eps) {
  const int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();
  pdl_wait();
  if (row >= rows) return;
  float4* xr = reinterpret_cast<float4*>(x + row * x_ld);
  const int nv = d >> 2;
  float4 v[NV4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      v[i] = xr[c];
      if constexpr (REDUCE) {
        if (bias != nullptr) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + c);
          v[i].x += b.x;
          v[i].y += b.y;
          v[i].z += b.z;
          v[i].w += b.w;
        }
        const float4* pr = reinterpret_cast<const float4*>(partial + row * static_cast<long long>(d)) + c;
        for (int sp = 0; sp < splits; ++sp) {
          const float4 p4 = __ldcg(pr + sp * (split_stride >> 2));
          v[i].x += p4.x;
          v[i].y += p4.y;
          v[i].z += p4.z;
          v[i].w += p4.w;
        }
        xr[c] = v[i];
      }
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const float mean = warp_sum(s) / static_cast<float>(d);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, e = v[i].z - mean, f = v[i].w - mean;
      q += (a * a + b * b) + (e * e + f * f);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(d) + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  uint2* yr = reinterpret_cast<uint2*>(y + row * y_ld);
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      const float4 g = __ldg(g4 + c), b = __ldg(b4 + c);
      uint2 o;
      o.x = pack_half2((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
      o.y = pack_half2((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
      yr[c] = o;
    }
  }
}

// Few rows (decode, last-position LayerNorms): one 128-thread CTA per row so that 256 rows spread over every SM and each
// thread has all of its loads (row, bias, every split's partial sum) in flight at once.
constexpr int LNR_THREADS = 128;

__device__ __forceinline__ float block_sum_128(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5;
  __syncthreads();  // protect `red` from the previous reduction's readers
  if ((threadIdx.x & 31) == 0) red[warp] = v;
  __syncthreads();
  return (red[0] + red[1]) + (red[2] + red[3]);
}

template <int NV>
__global__ void __launch_bounds__(LNR_THREADS)
layernorm_row_kernel(float* __restrict__ x, long long x_ld, const float* __restrict__ partial, int splits,
                     long long split_stride, const float* __restrict__ bias, const float* __restrict__ gamma,
                     const float* __restrict__ beta, __half* __restrict__ y, long long y_ld, int d, float eps) {
  __shared__ float red[4];
  const int row = blockIdx.x;
  const int nv = d >> 2;
  pdl_launch_dependents();
  // parameters are never written by a kernel of the stream: gamma / beta / bias are in registers before the wait, so the
  // chain behind it is row + partial sums -> two block reductions -> store
  float4 g[NV], be[NV], bi[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = threadIdx.x + i * LNR_THREADS;
    const bool in = c < nv;
    g[i] = in ? __ldg(reinterpret_cast<const float4*>(gamma) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    be[i] = in ? __ldg(reinterpret_cast<const float4*>(beta) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    bi[i] = in && splits > 0 && bias != nullptr ? __ldg(reinterpret_cast<const float4*>(bias) + c)
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  pdl_wait();
  float4* xr = reinterpret_cast<float4*>(x + row * x_ld);
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = threadIdx.x + i * LNR_THREADS;
    v[i] = c < nv ? xr[c] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (splits > 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {  // bias first, then the splits in ascending order: the sum is deterministic
      v[i].x += bi[i].x;
      v[i].y += bi[i].y;
      v[i].z += bi[i].z;
      v[i].w += bi[i].w;
    }
    const float4* pr = reinterpret_cast<const float4*>(partial + row * static_cast<long long>(d));
    const long long ss = split_stride >> 2;
#pragma unroll 8
    for (int sp = 0; sp < splits; ++sp) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = threadIdx.x + i * LNR_THREADS;
        if (c < nv) {
          const float4 p4 = __ldcg(pr + sp * ss + c);
          v[i].x += p4.x;
          v[i].y += p4.y;
          v[i].z += p4.z;
          v[i].w += p4.w;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = threadIdx.x + i * LNR_THREADS;
      if (c < nv) xr[c] = v[i];
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);  // padding lanes hold zeros
  const float mean = block_sum_128(s, red) / static_cast<float>(d);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = threadIdx.x + i * LNR_THREADS;
    if (c < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, e = v[i].z - mean, f = v[i].w - mean;
      q += (a * a + b * b) + (e * e + f * f);
    }
  }
  const float rstd = rsqrtf(block_sum_128(q, red) / static_cast<float>(d) + eps);
  uint2* yr = reinterpret_cast<uint2*>(y + row * y_ld);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = threadIdx.x + i * LNR_THREADS;
    if (c < nv) {
      uint2 o;
      o.x = pack_half2((v[i].x - mean) * rstd * g[i].x + be[i].x, (v[i].y - mean) * rstd * g[i].y + be[i].y);
      o.y = pack_half2((v[i].z - mean) * rstd * g[i].z + be[i].z, (v[i].w - mean) * rstd * g[i].w + be[i].w);
      yr[c] = o;
    }
  }
}

}  // namespace

int layernorm_reduce_run(float* h, int64_t h_ld, const float* partial, int splits, int64_t split_stride,
                         const float* bias, const float* gamma, const float* beta, __half* y, int64_t y_ld, int rows,
                         int d, float eps, cudaStream_t s) {
  CC_REQUIRE(d % 4 == 0 && d <= LN_MAX_V4 * 128, CC_ESHAPE, "layernorm: d=%d must be a multiple of 4 and <= %d", d,
             LN_MAX_V4 * 128);
  CC_REQUIRE(h_ld % 4 == 0 && y_ld % 4 == 0 && split_stride % 4 == 0, CC_EALIGN,
             "layernorm: row strides must be multiples of 4");
  if (rows <= 0) return CC_OK;
  const bool red = partial != nullptr && splits > 0;
  if (red || rows <= 2048) {
    const int nvt = (d / 4 + LNR_THREADS - 1) / LNR_THREADS;
#define CC_LNR_CASE(N)                                                                                               \
  if (nvt <= N) {                                                                                                    \
    CC_CUDA(launch_pdl(layernorm_row_kernel<N>, dim3(rows), dim3(LNR_THREADS), 0, s, h, static_cast<long long>(h_ld), \
                       partial, red ? splits : 0, static_cast<long long>(split_stride), bias, gamma, beta, y,        \
                       static_cast<long long>(y_ld), d, eps));                                                       \
    return CC_OK;                                                                                                    \
  }
    CC_LNR_CASE(1)
    CC_LNR_CASE(2)
    CC_LNR_CASE(4)
#undef CC_LNR_CASE
    return CC_ESHAPE;
  }
  const int grid = (rows + LN_WARPS - 1) / LN_WARPS;
  const int nv4 = (d / 4 + 31) / 32;
#define CC_LN_CASE(N)                                                                                              \
  if (nv4 <= N) {                                                                                                  \
    CC_CUDA(launch_pdl(layernorm_kernel<N, false>, dim3(grid), dim3(LN_WARPS * 32), 0, s, h,                        \
                       static_cast<long long>(h_ld), static_cast<const float*>(nullptr), 0, 0LL,                   \
                       static_cast<const float*>(nullptr), gamma, beta, y, static_cast<long long>(y_ld), rows, d,  \
                       eps));                                                                                      \
    return CC_OK;                                                                                                  \
  }
  CC_LN_CASE(2)
  CC_LN_CASE(4)
  CC_LN_CASE(6)
  CC_LN_CASE(8)
  CC_LN_CASE(12)
  CC_LN_CASE(16)
#undef CC_LN_CASE
  return CC_ESHAPE;
}

int layernorm_run(const float* x, int64_t x_ld, const float* gamma, const float* beta, __half* y, int64_t y_ld, int rows,
                  int d, float eps, cudaStream_t s) {
  return layernorm_reduce_run(const_cast<float*>(x), x_ld, nullptr, 0, 0, nullptr, gamma, beta, y, y_ld, rows, d, eps, s);
}

}  // namespace cc
