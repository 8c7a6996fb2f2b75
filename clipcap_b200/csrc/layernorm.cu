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

template <int NV4>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_kernel(const float* __restrict__ x, long long x_ld, const float* __restrict__ gamma,
                 const float* __restrict__ beta, __half* __restrict__ y, long long y_ld, int rows, int d, float eps) {
  const int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + row * x_ld);
  const int nv = d >> 2;
  float4 v[NV4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      v[i] = xr[c];
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

}  // namespace

int layernorm_run(const float* x, int64_t x_ld, const float* gamma, const float* beta, __half* y, int64_t y_ld, int rows,
                  int d, float eps, cudaStream_t s) {
  CC_REQUIRE(d % 4 == 0 && d <= LN_MAX_V4 * 128, CC_ESHAPE, "layernorm: d=%d must be a multiple of 4 and <= %d", d,
             LN_MAX_V4 * 128);
  CC_REQUIRE(x_ld % 4 == 0 && y_ld % 4 == 0, CC_EALIGN, "layernorm: row strides must be multiples of 4");
  if (rows <= 0) return CC_OK;
  const int grid = (rows + LN_WARPS - 1) / LN_WARPS;
  const int nv4 = (d / 4 + 31) / 32;
#define CC_LN_CASE(N)                                                                                           \
  if (nv4 <= N) {                                                                                               \
    layernorm_kernel<N><<<grid, LN_WARPS * 32, 0, s>>>(x, x_ld, gamma, beta, y, y_ld, rows, d, eps);            \
    CC_CUDA(cudaGetLastError());                                                                                \
    return CC_OK;                                                                                               \
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

}  // namespace cc
