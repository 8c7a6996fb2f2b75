// Kernels of the training step (declared in train.h): the backward halves of the pre-LN block (LayerNorm, activation,
// attention), the LM-head cross-entropy with ignore_index 0 (clipcap/model/model.py:109-110), the K-major turns that make
// weight gradients TN GEMMs, bias reductions and the AdamW update. All GEMM-shaped work of the backward pass runs on the
// tcgen05 GEMM (gemm.cu); what is here is HBM-bound element-wise / reduction work plus the small attention backward.
#include <algorithm>
#include <cstdlib>

#include "train.h"
// (common.h first: ptx.cuh uses printf)
#include "ptx.cuh"

namespace cc {
namespace {

inline int grid_for(long long work_items, int threads, int per_sm = 8) {
  const long long blocks = (work_items + threads - 1) / threads;
  const long long cap = static_cast<long long>(num_sms()) * per_sm;
  return static_cast<int>(std::max<long long>(1, std::min(blocks, cap)));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------- transpose (fp16) with zero padding
__global__ void transpose16_kernel(const __half* __restrict__ src, long long ld, int rows, int cols,
                                   __half* __restrict__ dst, long long ld_dst) {
  __shared__ __half tile[32][34];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;  // r: source row (= destination column)
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < rows && c < cols) ? src[static_cast<long long>(r) * ld + c] : __float2half_rn(0.f);
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (c < cols && r < ld_dst) dst[static_cast<long long>(c) * ld_dst + r] = tile[threadIdx.x][j];
  }
}

// ---------------------------------------------------------------- column sums (two stages, fixed order)
template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ldf<__half>(const __half* p) { return __half2float(*p); }

// grid (ceil(cols/32), RS): block (32, 8); slice blockIdx.y sums its rows of 32 columns into part[blockIdx.y][col]
template <typename T>
__global__ void colsum_partial_kernel(const T* __restrict__ x, long long ld, int rows, int cols, int rows_per_slice,
                                      float* __restrict__ part) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int r_begin = blockIdx.y * rows_per_slice;
  const int r_end = min(rows, r_begin + rows_per_slice);
  float acc = 0.f;
  if (c < cols)
    for (int r = r_begin + threadIdx.y; r < r_end; r += 8) acc += ldf<T>(x + static_cast<long long>(r) * ld + c);
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][threadIdx.x];
    part[static_cast<long long>(blockIdx.y) * cols + c] = t;
  }
}
__global__ void colsum_final_kernel(const float* __restrict__ part, int slices, int cols, float alpha,
                                    float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float t = 0.f;
  for (int sidx = 0; sidx < slices; ++sidx) t += part[static_cast<long long>(sidx) * cols + c];
  out[c] = alpha * t;
}

__global__ void scale_f32_kernel(float* __restrict__ x, long long n, float alpha) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += stride) x[i] *= alpha;
}

// ---------------------------------------------------------------- activations
__device__ __forceinline__ float gelu_new_f(float x) {
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  return 0.5f * x * (1.f + tanhf(u));
}
__device__ __forceinline__ float gelu_new_grad(float x) {
  const float x2 = x * x;
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x2);
  const float t = tanhf(u);
  const float du = 0.7978845608028654f * (1.f + 3.f * 0.044715f * x2);
  return 0.5f * (1.f + t) + 0.5f * x * (1.f - t * t) * du;
}
// MODE 0: out = gelu_new(a); 1: a *= gelu_new'(b); 2: a = b > 0 ? a : 0.   Two halves per thread-iteration.
template <int MODE>
__global__ void act_kernel(__half2* __restrict__ a, const __half2* __restrict__ b, __half2* __restrict__ out, long long n2) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n2; i += stride) {
    if constexpr (MODE == 0) {
      const float2 p = __half22float2(b[i]);
      out[i] = __floats2half2_rn(gelu_new_f(p.x), gelu_new_f(p.y));
    } else if constexpr (MODE == 1) {
      const float2 g = __half22float2(a[i]);
      const float2 p = __half22float2(b[i]);
      a[i] = __floats2half2_rn(g.x * gelu_new_grad(p.x), g.y * gelu_new_grad(p.y));
    } else {
      const float2 g = __half22float2(a[i]);
      const float2 hdn = __half22float2(b[i]);
      a[i] = __floats2half2_rn(hdn.x > 0.f ? g.x : 0.f, hdn.y > 0.f ? g.y : 0.f);
    }
  }
}

// ---------------------------------------------------------------- LayerNorm backward
constexpr int LNB_WARPS = 4;
// One warp per row, grid-stride over rows; a lane owns float4 columns lane + 32 i. PARAM: per-lane partial dgamma / dbeta
// are combined across the block's warps and written to scratch[blockIdx.x][2][d]; a column sum finishes them.
template <int NV4, bool PARAM>
__global__ void __launch_bounds__(LNB_WARPS * 32)
layernorm_bwd_kernel(const float* __restrict__ dy, long long dy_ld, const float* __restrict__ x, long long x_ld,
                     const float* __restrict__ gamma, float* __restrict__ dx, long long dx_ld, int rows, int d, float eps,
                     float* __restrict__ scratch) {
  extern __shared__ float lnb_smem[];  // PARAM: [LNB_WARPS][2][d]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = d >> 2;
  const float inv_d = 1.f / static_cast<float>(d);
  float4 dg[PARAM ? NV4 : 1], db[PARAM ? NV4 : 1];
  if constexpr (PARAM) {
#pragma unroll
    for (int i = 0; i < NV4; ++i) dg[i] = db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float4 g4[NV4];
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const int c = lane + i * 32;
    g4[i] = c < nv ? __ldg(reinterpret_cast<const float4*>(gamma) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int row = blockIdx.x * LNB_WARPS + warp; row < rows; row += gridDim.x * LNB_WARPS) {
    const float4* xr = reinterpret_cast<const float4*>(x + row * x_ld);
    const float4* dyr = reinterpret_cast<const float4*>(dy + row * dy_ld);
    float4 xv[NV4], gv[NV4];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + i * 32;
      xv[i] = c < nv ? xr[c] : make_float4(0.f, 0.f, 0.f, 0.f);
      gv[i] = c < nv ? dyr[c] : make_float4(0.f, 0.f, 0.f, 0.f);
      sum += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
    }
    const float mean = warp_sum(sum) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + i * 32;
      if (c < nv) {
        xv[i].x -= mean; xv[i].y -= mean; xv[i].z -= mean; xv[i].w -= mean;
        q += (xv[i].x * xv[i].x + xv[i].y * xv[i].y) + (xv[i].z * xv[i].z + xv[i].w * xv[i].w);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
    float s1 = 0.f, s2 = 0.f;  // sum g, sum g * xhat  with g = dy * gamma
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + i * 32;
      if (c < nv) {
        xv[i].x *= rstd; xv[i].y *= rstd; xv[i].z *= rstd; xv[i].w *= rstd;  // xhat
        if constexpr (PARAM) {
          dg[i].x += gv[i].x * xv[i].x; dg[i].y += gv[i].y * xv[i].y; dg[i].z += gv[i].z * xv[i].z; dg[i].w += gv[i].w * xv[i].w;
          db[i].x += gv[i].x; db[i].y += gv[i].y; db[i].z += gv[i].z; db[i].w += gv[i].w;
        }
        gv[i].x *= g4[i].x; gv[i].y *= g4[i].y; gv[i].z *= g4[i].z; gv[i].w *= g4[i].w;
        s1 += (gv[i].x + gv[i].y) + (gv[i].z + gv[i].w);
        s2 += (gv[i].x * xv[i].x + gv[i].y * xv[i].y) + (gv[i].z * xv[i].z + gv[i].w * xv[i].w);
      }
    }
    const float c1 = warp_sum(s1) * inv_d, c2 = warp_sum(s2) * inv_d;
    float4* dxr = reinterpret_cast<float4*>(dx + row * dx_ld);
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + i * 32;
      if (c < nv) {
        float4 o = dxr[c];
        o.x += rstd * (gv[i].x - c1 - xv[i].x * c2);
        o.y += rstd * (gv[i].y - c1 - xv[i].y * c2);
        o.z += rstd * (gv[i].z - c1 - xv[i].z * c2);
        o.w += rstd * (gv[i].w - c1 - xv[i].w * c2);
        dxr[c] = o;
      }
    }
  }
  if constexpr (PARAM) {
    float4* sg = reinterpret_cast<float4*>(lnb_smem + static_cast<size_t>(warp) * 2 * d);
    float4* sb = reinterpret_cast<float4*>(lnb_smem + static_cast<size_t>(warp) * 2 * d + d);
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + i * 32;
      if (c < nv) {
        sg[c] = dg[i];
        sb[c] = db[i];
      }
    }
    __syncthreads();
    float* out = scratch + static_cast<size_t>(blockIdx.x) * 2 * d;
    for (int c = threadIdx.x; c < 2 * d; c += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < LNB_WARPS; ++w) t += lnb_smem[static_cast<size_t>(w) * 2 * d + c];
      out[c] = t;
    }
  }
}

constexpr int LNB_MAX_BLOCKS = 296;

// ---------------------------------------------------------------- attention backward
// One CTA per (batch, head). Shared memory: Q, K, V, dO as [S][hd + 2] fp16 (row stride of an odd number of 32-bit words:
// conflict-free when a warp reads one column pair of 32 different rows) and P / dS as [S][S + 1] fp32.
constexpr int ATB_THREADS = 256;
constexpr int ATB_MAXJ = 8;  // keys per lane: S <= 256

__device__ __forceinline__ float dot_rows(const __half* a, const __half* b, int hd) {
  float acc = 0.f;
  const __half2* a2 = reinterpret_cast<const __half2*>(a);
  const __half2* b2 = reinterpret_cast<const __half2*>(b);
  for (int c = 0; c < (hd >> 1); ++c) {
    const float2 x = __half22float2(a2[c]);
    const float2 y = __half22float2(b2[c]);
    acc += x.x * y.x + x.y * y.y;
  }
  return acc;
}

__global__ void __launch_bounds__(ATB_THREADS)
attention_bwd_kernel(const __half* __restrict__ q, const __half* __restrict__ k, const __half* __restrict__ v, long long ld,
                     const __half* __restrict__ d_o, long long ldo, __half* __restrict__ dq, __half* __restrict__ dk,
                     __half* __restrict__ dv, long long ldd, int S, int H, int hd, int causal, float scale) {
  extern __shared__ __align__(16) uint8_t atb_smem[];
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int HDP = hd + 2, SP = S + 1, hd2 = hd >> 1;
  __half* Qs = reinterpret_cast<__half*>(atb_smem);
  __half* Ks = Qs + static_cast<size_t>(S) * HDP;
  __half* Vs = Ks + static_cast<size_t>(S) * HDP;
  __half* Os = Vs + static_cast<size_t>(S) * HDP;  // dO
  float* Ps = reinterpret_cast<float*>(Os + static_cast<size_t>(S) * HDP + ((static_cast<size_t>(S) * HDP * 4) & 1));
  const long long row0 = static_cast<long long>(b) * S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = ATB_THREADS / 32;

  for (int idx = threadIdx.x; idx < S * hd2; idx += ATB_THREADS) {
    const int r = idx / hd2, c2 = idx - r * hd2;
    const long long g = (row0 + r) * ld + h * hd + 2 * c2;
    reinterpret_cast<__half2*>(Qs + r * HDP)[c2] = *reinterpret_cast<const __half2*>(q + g);
    reinterpret_cast<__half2*>(Ks + r * HDP)[c2] = *reinterpret_cast<const __half2*>(k + g);
    reinterpret_cast<__half2*>(Vs + r * HDP)[c2] = *reinterpret_cast<const __half2*>(v + g);
    reinterpret_cast<__half2*>(Os + r * HDP)[c2] = *reinterpret_cast<const __half2*>(d_o + (row0 + r) * ldo + h * hd + 2 * c2);
  }
  __syncthreads();

  // P = softmax(scale * Q K^T) (causal: keys j <= i), one warp per query row
  for (int i = warp; i < S; i += nwarps) {
    const int jmax = causal ? i + 1 : S;
    float mx = -INFINITY;
    for (int j = lane; j < jmax; j += 32) {
      const float sc = scale * dot_rows(Qs + i * HDP, Ks + j * HDP, hd);
      Ps[i * SP + j] = sc;
      mx = fmaxf(mx, sc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < jmax; j += 32) {
      const float e = __expf(Ps[i * SP + j] - mx);
      Ps[i * SP + j] = e;
      sum += e;
    }
    const float inv = 1.f / warp_sum(sum);
    for (int j = lane; j < S; j += 32) Ps[i * SP + j] = j < jmax ? Ps[i * SP + j] * inv : 0.f;
  }
  __syncthreads();

  // dV[j] = sum_i P[i][j] dO[i]
  for (int idx = threadIdx.x; idx < S * hd2; idx += ATB_THREADS) {
    const int j = idx / hd2, c2 = idx - j * hd2;
    float ax = 0.f, ay = 0.f;
    for (int i = causal ? j : 0; i < S; ++i) {
      const float p = Ps[i * SP + j];
      const float2 o = __half22float2(reinterpret_cast<const __half2*>(Os + i * HDP)[c2]);
      ax += p * o.x;
      ay += p * o.y;
    }
    *reinterpret_cast<__half2*>(dv + (row0 + j) * ldd + h * hd + 2 * c2) = __floats2half2_rn(ax, ay);
  }
  __syncthreads();

  // dS[i][j] = scale * P[i][j] * (dP[i][j] - sum_j P[i][j] dP[i][j]),  dP[i][j] = dO[i] . V[j]   (overwrites P)
  for (int i = warp; i < S; i += nwarps) {
    const int jmax = causal ? i + 1 : S;
    float dpv[ATB_MAXJ];
    float dsum = 0.f;
#pragma unroll
    for (int t = 0; t < ATB_MAXJ; ++t) {
      const int j = lane + 32 * t;
      dpv[t] = 0.f;
      if (j < jmax) {
        dpv[t] = dot_rows(Os + i * HDP, Vs + j * HDP, hd);
        dsum += Ps[i * SP + j] * dpv[t];
      }
    }
    dsum = warp_sum(dsum);
#pragma unroll
    for (int t = 0; t < ATB_MAXJ; ++t) {
      const int j = lane + 32 * t;
      if (j < jmax) Ps[i * SP + j] = scale * Ps[i * SP + j] * (dpv[t] - dsum);
    }
  }
  __syncthreads();

  // dQ[i] = sum_j dS[i][j] K[j];  dK[j] = sum_i dS[i][j] Q[i]
  for (int idx = threadIdx.x; idx < S * hd2; idx += ATB_THREADS) {
    const int r = idx / hd2, c2 = idx - r * hd2;
    float qx = 0.f, qy = 0.f, kx = 0.f, ky = 0.f;
    const int jmax = causal ? r + 1 : S;
    for (int j = 0; j < jmax; ++j) {
      const float ds = Ps[r * SP + j];
      const float2 kk = __half22float2(reinterpret_cast<const __half2*>(Ks + j * HDP)[c2]);
      qx += ds * kk.x;
      qy += ds * kk.y;
    }
    for (int i = causal ? r : 0; i < S; ++i) {
      const float ds = Ps[i * SP + r];
      const float2 qq = __half22float2(reinterpret_cast<const __half2*>(Qs + i * HDP)[c2]);
      kx += ds * qq.x;
      ky += ds * qq.y;
    }
    *reinterpret_cast<__half2*>(dq + (row0 + r) * ldd + h * hd + 2 * c2) = __floats2half2_rn(qx, qy);
    *reinterpret_cast<__half2*>(dk + (row0 + r) * ldd + h * hd + 2 * c2) = __floats2half2_rn(kx, ky);
  }
}

// ---------------------------------------------------------------- attention backward on warp MMA (mma.m16n8k16)
// Same contract as attention_bwd_kernel for the shapes of the training step (S <= 8 * NTMAX keys): Q, K, V, dO and the two
// S x S matrices P and dS live in shared memory as fp16 with a 16-byte row pad (conflict-free ldmatrix); every product
//   S = Q K^T, dP = dO V^T (A row-major, B stored [n][k]);  dV = P^T dO, dK = dS^T Q (A transposed, B stored [k][n]);
//   dQ = dS K (A row-major, B stored [k][n])
// is a strip of 16 output rows per warp. Softmax statistics and dS are formed in fp32 registers.
template <int HD, int NTMAX>
__global__ void __launch_bounds__(256)
attention_bwd_mma_kernel(const __half* __restrict__ q, const __half* __restrict__ k, const __half* __restrict__ v,
                         long long ld, const __half* __restrict__ d_o, long long ldo, __half* __restrict__ dq,
                         __half* __restrict__ dk, __half* __restrict__ dv, long long ldd, int S, int H, int causal,
                         float scale) {
  extern __shared__ __align__(16) uint8_t atm_smem[];
  constexpr int PH = HD * 2 + 16;  // row pitch (bytes) of the [S][HD] tiles
  const int SPAD = (S + 15) & ~15;
  const int PS = SPAD * 2 + 16;    // row pitch of the [S][S] matrices
  const uint32_t sQ = smem_u32(atm_smem);
  const uint32_t sK = sQ + SPAD * PH, sV = sK + SPAD * PH, sO = sV + SPAD * PH;
  const uint32_t sP = sO + SPAD * PH, sD = sP + SPAD * PS;
  uint8_t* pP = atm_smem + 4 * SPAD * PH;
  uint8_t* pD = pP + SPAD * PS;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const long long row0 = static_cast<long long>(b) * S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3, mi = lane >> 3, rr = lane & 7;

  {  // operand tiles, rows >= S zero-filled
    constexpr int CPR = HD / 8;
    for (int i = threadIdx.x; i < SPAD * CPR; i += blockDim.x) {
      const int r = i / CPR, c = i - r * CPR;
      const bool ok = r < S;
      const long long gq = (row0 + (ok ? r : 0)) * ld + h * HD + c * 8;
      const long long go = (row0 + (ok ? r : 0)) * ldo + h * HD + c * 8;
      const uint32_t so = r * PH + c * 16;
      cp_async_16(sQ + so, q + gq, ok);
      cp_async_16(sK + so, k + gq, ok);
      cp_async_16(sV + so, v + gq, ok);
      cp_async_16(sO + so, d_o + go, ok);
    }
    cp_async_commit();
    cp_async_wait<0>();
  }
  __syncthreads();

  const int strips = SPAD >> 4;
  // ---- phase 1: P and dS, one strip of 16 query rows per warp
  for (int st = warp; st < strips; st += nwarps) {
    const int m0 = st * 16;
    float acc[NTMAX][4];
    // acc = X[m0 strip] Y^T with X, Y in {Q, K} or {dO, V}
    auto strip_xyT = [&](uint32_t sX, uint32_t sY) {
#pragma unroll
      for (int nt = 0; nt < NTMAX; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks) {
        uint32_t a[4];
        ldmatrix_x4(a, sX + (m0 + (mi & 1) * 8 + rr) * PH + (ks * 16 + (mi >> 1) * 8) * 2);
#pragma unroll
        for (int np = 0; np < NTMAX / 2; ++np) {
          if (np * 16 < SPAD) {
            uint32_t bf[4];
            ldmatrix_x4(bf, sY + (np * 16 + (mi >> 1) * 8 + rr) * PH + (ks * 16 + (mi & 1) * 8) * 2);
            mma_16816(acc[np * 2], a, bf[0], bf[1]);
            mma_16816(acc[np * 2 + 1], a, bf[2], bf[3]);
          }
        }
      }
    };
    strip_xyT(sQ, sK);
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < NTMAX; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = nt * 8 + 2 * t + (e & 1), row = m0 + g + (e >> 1) * 8;
        const bool ok = key < S && (!causal || key <= row);
        acc[nt][e] = ok ? acc[nt][e] * scale : -INFINITY;
        mx[e >> 1] = fmaxf(mx[e >> 1], acc[nt][e]);
      }
    }
    float sum[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      if (mx[r] == -INFINITY) mx[r] = 0.f;  // padded query rows: every key masked
    }
#pragma unroll
    for (int nt = 0; nt < NTMAX; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pe = __expf(acc[nt][e] - mx[e >> 1]);
        acc[nt][e] = pe;
        sum[e >> 1] += pe;
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 1);
      sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 2);
      sum[r] = sum[r] > 0.f ? 1.f / sum[r] : 0.f;
    }
#pragma unroll
    for (int nt = 0; nt < NTMAX; ++nt) {
      if (nt * 8 < SPAD) {
        const int c = nt * 8 + 2 * t;
        *reinterpret_cast<uint32_t*>(pP + (m0 + g) * PS + c * 2) = pack_half2(acc[nt][0] * sum[0], acc[nt][1] * sum[0]);
        *reinterpret_cast<uint32_t*>(pP + (m0 + g + 8) * PS + c * 2) = pack_half2(acc[nt][2] * sum[1], acc[nt][3] * sum[1]);
      }
    }
    __syncwarp();
    strip_xyT(sO, sV);  // dP
    float dsum[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < NTMAX; ++nt) {
      if (nt * 8 < SPAD) {
        const int c = nt * 8 + 2 * t;
        const float2 p0 = __half22float2(*reinterpret_cast<const __half2*>(pP + (m0 + g) * PS + c * 2));
        const float2 p1 = __half22float2(*reinterpret_cast<const __half2*>(pP + (m0 + g + 8) * PS + c * 2));
        dsum[0] += p0.x * acc[nt][0] + p0.y * acc[nt][1];
        dsum[1] += p1.x * acc[nt][2] + p1.y * acc[nt][3];
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      dsum[r] += __shfl_xor_sync(0xffffffffu, dsum[r], 1);
      dsum[r] += __shfl_xor_sync(0xffffffffu, dsum[r], 2);
    }
#pragma unroll
    for (int nt = 0; nt < NTMAX; ++nt) {
      if (nt * 8 < SPAD) {
        const int c = nt * 8 + 2 * t;
        const float2 p0 = __half22float2(*reinterpret_cast<const __half2*>(pP + (m0 + g) * PS + c * 2));
        const float2 p1 = __half22float2(*reinterpret_cast<const __half2*>(pP + (m0 + g + 8) * PS + c * 2));
        *reinterpret_cast<uint32_t*>(pD + (m0 + g) * PS + c * 2) =
            pack_half2(scale * p0.x * (acc[nt][0] - dsum[0]), scale * p0.y * (acc[nt][1] - dsum[0]));
        *reinterpret_cast<uint32_t*>(pD + (m0 + g + 8) * PS + c * 2) =
            pack_half2(scale * p1.x * (acc[nt][2] - dsum[1]), scale * p1.y * (acc[nt][3] - dsum[1]));
      }
    }
  }
  __syncthreads();

  // ---- phase 2: the three [S, HD] outputs, one strip of 16 output rows per (warp, product)
  // out[m0 strip] = op(A) B, B stored [k][n] with pitch PH; TRANS_A: A[m][k] = X[k][m] (X = P or dS, pitch PS)
  auto strip_out = [&](uint32_t sX, bool trans_a, uint32_t sB, int m0, int k_lo, int k_hi, __half* out) {
    float acc[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    for (int k0 = k_lo; k0 < k_hi; k0 += 16) {
      uint32_t a[4];
      if (trans_a) ldmatrix_x4_trans(a, sX + (k0 + (mi >> 1) * 8 + rr) * PS + (m0 + (mi & 1) * 8) * 2);
      else ldmatrix_x4(a, sX + (m0 + (mi & 1) * 8 + rr) * PS + (k0 + (mi >> 1) * 8) * 2);
#pragma unroll
      for (int np = 0; np < HD / 16; ++np) {
        uint32_t bf[4];
        ldmatrix_x4_trans(bf, sB + (k0 + (mi & 1) * 8 + rr) * PH + (np * 16 + (mi >> 1) * 8) * 2);
        mma_16816(acc[np * 2], a, bf[0], bf[1]);
        mma_16816(acc[np * 2 + 1], a, bf[2], bf[3]);
      }
    }
    const int ra = m0 + g, rb = m0 + g + 8;
#pragma unroll
    for (int nt = 0; nt < HD / 8; ++nt) {
      const int c = h * HD + nt * 8 + 2 * t;
      if (ra < S) *reinterpret_cast<uint32_t*>(out + (row0 + ra) * ldd + c) = pack_half2(acc[nt][0], acc[nt][1]);
      if (rb < S) *reinterpret_cast<uint32_t*>(out + (row0 + rb) * ldd + c) = pack_half2(acc[nt][2], acc[nt][3]);
    }
  };
  for (int w = warp; w < 3 * strips; w += nwarps) {
    const int which = w / strips, m0 = (w - which * strips) * 16;
    if (which == 0) strip_out(sP, true, sO, m0, causal ? m0 : 0, SPAD, dv);        // dV[j] = sum_{i >= j} P[i][j] dO[i]
    else if (which == 1) strip_out(sD, true, sQ, m0, causal ? m0 : 0, SPAD, dk);   // dK[j] = sum_{i >= j} dS[i][j] Q[i]
    else strip_out(sD, false, sK, m0, 0, causal ? m0 + 16 : SPAD, dq);             // dQ[i] = sum_{j <= i} dS[i][j] K[j]
  }
}

// ---------------------------------------------------------------- cross-entropy (ignore_index 0)
__global__ void count_valid_kernel(const int32_t* __restrict__ targets, int n, int* __restrict__ n_valid) {
  __shared__ int red[32];
  int c = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) c += targets[i] != 0 ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) t += red[w];
    *n_valid = t;
  }
}

constexpr int CE_THREADS = 256;
__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = red[0];
#pragma unroll
  for (int w = 1; w < CE_THREADS / 32; ++w) t = is_max ? fmaxf(t, red[w]) : t + red[w];
  return t;
}

__global__ void __launch_bounds__(CE_THREADS)
ce_loss_kernel(const float* __restrict__ logits, long long ld, int V, const int32_t* __restrict__ targets,
               const int* __restrict__ n_valid, float loss_scale, float* __restrict__ row_loss,
               __half* __restrict__ dlogits, long long ld_d) {
  __shared__ float red[CE_THREADS / 32];
  const int row = blockIdx.x;
  const int t = targets[row];
  __half* drow = dlogits + row * ld_d;
  if (t == 0) {  // ignore_index
    if (threadIdx.x == 0) row_loss[row] = 0.f;
    for (int c = threadIdx.x; c < ld_d; c += CE_THREADS) drow[c] = __float2half_rn(0.f);
    return;
  }
  const float* lr = logits + row * ld;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < V; c += CE_THREADS) mx = fmaxf(mx, lr[c]);
  mx = block_reduce(mx, red, true);
  float sum = 0.f;
  for (int c = threadIdx.x; c < V; c += CE_THREADS) sum += __expf(lr[c] - mx);
  sum = block_reduce(sum, red, false);
  const float lse = mx + __logf(sum);
  if (threadIdx.x == 0) row_loss[row] = lse - lr[t];
  const float coef = loss_scale / static_cast<float>(*n_valid);
  const float inv = coef / sum;
  for (int c = threadIdx.x; c < ld_d; c += CE_THREADS) {
    float g = 0.f;
    if (c < V) g = __expf(lr[c] - mx) * inv - (c == t ? coef : 0.f);
    drow[c] = __float2half_rn(g);
  }
}

__global__ void loss_reduce_kernel(const float* __restrict__ row_loss, int n, const int* __restrict__ n_valid,
                                   float* __restrict__ loss) {
  __shared__ float red[CE_THREADS / 32];
  float sacc = 0.f;
  for (int i = threadIdx.x; i < n; i += CE_THREADS) sacc += row_loss[i];
  sacc = block_reduce(sacc, red, false);
  if (threadIdx.x == 0) *loss = sacc / static_cast<float>(*n_valid);
}

// ---------------------------------------------------------------- teacher-forced LM input
__global__ void train_embed_kernel(const int32_t* __restrict__ tokens, int Tt, int K, const float* __restrict__ prefix,
                                   long long prefix_ld, const float* __restrict__ wte, const float* __restrict__ wpe,
                                   float* __restrict__ h, int32_t* __restrict__ targets, int d, int V) {
  const int T = K + Tt;
  const int b = blockIdx.x / T, t = blockIdx.x % T;
  const float* src;
  if (t < K) {
    src = prefix + b * prefix_ld + static_cast<long long>(t) * d;
  } else {
    int tok = tokens[b * Tt + (t - K)];
    tok = tok < 0 ? 0 : (tok >= V ? V - 1 : tok);
    if (threadIdx.x == 0) targets[b * Tt + (t - K)] = tok;
    src = wte + static_cast<long long>(tok) * d;
  }
  const float* pe = wpe + static_cast<long long>(t) * d;
  float* dst = h + static_cast<long long>(blockIdx.x) * d;
  for (int c = threadIdx.x; c < d; c += blockDim.x) dst[c] = src[c] + pe[c];
}

// ---------------------------------------------------------------- overflow detection for the static loss scale
__global__ void count_nonfinite_kernel(const float* __restrict__ x, long long n, int* __restrict__ count) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  int bad = 0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += stride) bad += !isfinite(x[i]);
  bad = __reduce_add_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0 && bad != 0) atomicAdd(count, bad);
}

// ---------------------------------------------------------------- AdamW
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, long long n, float lr, float beta1, float beta2, float eps,
                             float weight_decay, float bc1, float bc2_sqrt) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += stride) {
    const float gi = g[i];
    // A non-finite gradient element (fp16 overflow of an activation gradient under the static loss scale) must not reach the
    // moments: exp_avg / exp_avg_sq would stay poisoned for good. The element is skipped for this step; the step's
    // non-finite count is reported by cc_train_step (cc_train_last_nonfinite) so a caller can lower its loss scale.
    if (!isfinite(gi)) continue;
    float pi = p[i] * (1.f - lr * weight_decay);
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi;
  }
}

}  // namespace

// ==================================================================== launchers
int transpose16_run(const __half* src, int64_t ld, int rows, int cols, __half* dst, int64_t ld_dst, cudaStream_t s) {
  CC_REQUIRE(rows > 0 && cols > 0 && ld_dst >= rows && ld_dst % 8 == 0, CC_ESHAPE,
             "transpose16: rows=%d cols=%d ld_dst=%lld", rows, cols, (long long)ld_dst);
  dim3 grid(static_cast<unsigned>((ld_dst + 31) / 32), static_cast<unsigned>((cols + 31) / 32));
  transpose16_kernel<<<grid, dim3(32, 8), 0, s>>>(src, ld, rows, cols, dst, ld_dst);
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

namespace {
constexpr int COLSUM_MAX_SLICES = 64;
template <typename T>
int colsum_run(const T* x, int64_t ld, int rows, int cols, float alpha, float* out, float* scratch, size_t scratch_floats,
               cudaStream_t s) {
  CC_REQUIRE(rows > 0 && cols > 0, CC_ESHAPE, "colsum: rows=%d cols=%d", rows, cols);
  CC_REQUIRE(scratch != nullptr && scratch_floats >= static_cast<size_t>(cols), CC_EINVAL,
             "colsum: scratch of %zu floats cannot hold one slice of %d columns", scratch_floats, cols);
  int slices = std::min(COLSUM_MAX_SLICES, (rows + 63) / 64);
  slices = std::min<size_t>(slices, scratch_floats / cols);
  const int per = (rows + slices - 1) / slices;
  slices = (rows + per - 1) / per;
  colsum_partial_kernel<T><<<dim3((cols + 31) / 32, slices), dim3(32, 8), 0, s>>>(x, ld, rows, cols, per, scratch);
  colsum_final_kernel<<<(cols + 255) / 256, 256, 0, s>>>(scratch, slices, cols, alpha, out);
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}
}  // namespace

size_t colsum_scratch_floats(int max_cols) { return static_cast<size_t>(COLSUM_MAX_SLICES) * max_cols; }

int colsum_f32_run(const float* x, int64_t ld, int rows, int cols, float alpha, float* out, float* scratch,
                   size_t scratch_floats, cudaStream_t s) {
  return colsum_run<float>(x, ld, rows, cols, alpha, out, scratch, scratch_floats, s);
}
int colsum_f16_run(const __half* x, int64_t ld, int rows, int cols, float alpha, float* out, float* scratch,
                   size_t scratch_floats, cudaStream_t s) {
  return colsum_run<__half>(x, ld, rows, cols, alpha, out, scratch, scratch_floats, s);
}

int scale_f32_run(float* x, int64_t n, float alpha, cudaStream_t s) {
  if (n <= 0) return CC_OK;
  scale_f32_kernel<<<grid_for(n, 256), 256, 0, s>>>(x, n, alpha);
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

int gelu_new_fwd_run(const __half* pre, __half* out, int64_t n, cudaStream_t s) {
  CC_REQUIRE(n % 2 == 0, CC_ESHAPE, "gelu_new_fwd: n=%lld must be even", (long long)n);
  act_kernel<0><<<grid_for(n / 2, 256), 256, 0, s>>>(nullptr, reinterpret_cast<const __half2*>(pre),
                                                     reinterpret_cast<__half2*>(out), n / 2);
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}
int gelu_new_bwd_run(__half* dhid, const __half* pre, int64_t n, cudaStream_t s) {
  CC_REQUIRE(n % 2 == 0, CC_ESHAPE, "gelu_new_bwd: n=%lld must be even", (long long)n);
  act_kernel<1><<<grid_for(n / 2, 256), 256, 0, s>>>(reinterpret_cast<__half2*>(dhid),
                                                     reinterpret_cast<const __half2*>(pre), nullptr, n / 2);
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}
int relu_bwd_run(__half* dhid, const __half* hid, int64_t n, cudaStream_t s) {
  CC_REQUIRE(n % 2 == 0, CC_ESHAPE, "relu_bwd: n=%lld must be even", (long long)n);
  act_kernel<2><<<grid_for(n / 2, 256), 256, 0, s>>>(reinterpret_cast<__half2*>(dhid),
                                                     reinterpret_cast<const __half2*>(hid), nullptr, n / 2);
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

size_t ln_bwd_scratch_floats(int d) {  // block partials [LNB_MAX_BLOCKS][2d] + result [2d] + column-sum slices
  return static_cast<size_t>(LNB_MAX_BLOCKS) * 2 * d + 2 * static_cast<size_t>(d) + colsum_scratch_floats(2 * d);
}

int layernorm_bwd_run(const float* dy, int64_t dy_ld, const float* x, int64_t x_ld, const float* gamma, float* dx_accum,
                      int64_t dx_ld, int rows, int d, float eps, float* dgamma, float* dbeta, float alpha,
                      float* scratch, cudaStream_t s) {
  CC_REQUIRE(d % 4 == 0 && d <= 16 * 128, CC_ESHAPE, "layernorm_bwd: d=%d must be a multiple of 4 and <= 2048", d);
  CC_REQUIRE(dy_ld % 4 == 0 && x_ld % 4 == 0 && dx_ld % 4 == 0, CC_EALIGN, "layernorm_bwd: strides must be multiples of 4");
  if (rows <= 0) return CC_OK;
  const bool param = dgamma != nullptr && dbeta != nullptr;
  CC_REQUIRE(!param || scratch != nullptr, CC_EINVAL, "layernorm_bwd: parameter gradients need scratch");
  const int grid = std::min((rows + LNB_WARPS - 1) / LNB_WARPS, param ? LNB_MAX_BLOCKS : num_sms() * 8);
  const int nv4 = (d / 4 + 31) / 32;
  const size_t smem = param ? static_cast<size_t>(LNB_WARPS) * 2 * d * sizeof(float) : 0;
#define CC_LNB_CASE(N)                                                                                             \
  if (nv4 <= N) {                                                                                                  \
    if (param) {                                                                                                   \
      CC_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                   static_cast<int>(smem)));                                                       \
      layernorm_bwd_kernel<N, true><<<grid, LNB_WARPS * 32, smem, s>>>(dy, dy_ld, x, x_ld, gamma, dx_accum, dx_ld,  \
                                                                       rows, d, eps, scratch);                     \
    } else {                                                                                                       \
      layernorm_bwd_kernel<N, false><<<grid, LNB_WARPS * 32, 0, s>>>(dy, dy_ld, x, x_ld, gamma, dx_accum, dx_ld,    \
                                                                     rows, d, eps, nullptr);                       \
    }                                                                                                              \
  } else
  CC_LNB_CASE(1)
  CC_LNB_CASE(2)
  CC_LNB_CASE(4)
  CC_LNB_CASE(8)
  CC_LNB_CASE(12)
  CC_LNB_CASE(16) { return CC_ESHAPE; }
#undef CC_LNB_CASE
  CC_CUDA(cudaGetLastError());
  if (param) {
    // scratch is [grid][2][d]: one column sum over the blocks gives dgamma (first d columns) and dbeta (last d)
    float* both = scratch + static_cast<size_t>(LNB_MAX_BLOCKS) * 2 * d;  // [2][d] result, then the column-sum slices
    CC_TRY(colsum_f32_run(scratch, 2 * d, grid, 2 * d, alpha, both, both + 2 * d, colsum_scratch_floats(2 * d), s));
    CC_CUDA(cudaMemcpyAsync(dgamma, both, sizeof(float) * d, cudaMemcpyDeviceToDevice, s));
    CC_CUDA(cudaMemcpyAsync(dbeta, both + d, sizeof(float) * d, cudaMemcpyDeviceToDevice, s));
  }
  return CC_OK;
}

namespace {
template <int HD, int NTMAX>
int launch_attn_bwd_mma(const __half* q, const __half* k, const __half* v, int64_t ld, const __half* d_o, int64_t ldo,
                        __half* dq, __half* dk, __half* dv, int64_t ldd, int B, int S, int H, bool causal, float scale,
                        cudaStream_t s) {
  const int spad = (S + 15) & ~15;
  const size_t smem = 4 * static_cast<size_t>(spad) * (HD * 2 + 16) + 2 * static_cast<size_t>(spad) * (spad * 2 + 16);
  if (smem > 227 * 1024) return CC_ESHAPE;
  auto kern = attention_bwd_mma_kernel<HD, NTMAX>;
  CC_OPT_IN_SMEM(kern, 227 * 1024);
  kern<<<B * H, 256, smem, s>>>(q, k, v, ld, d_o, ldo, dq, dk, dv, ldd, S, H, causal ? 1 : 0, scale);
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}
bool attn_bwd_scalar_forced() {
  static const bool on = [] {
    const char* e = getenv("CLIPCAP_B200_ATTN_BWD_SCALAR");
    return e != nullptr && e[0] == '1';
  }();
  return on;
}
}  // namespace

int attention_bwd_run(const __half* q, const __half* k, const __half* v, int64_t ld, const __half* d_o, int64_t ldo,
                      __half* dq, __half* dk, __half* dv, int64_t ldd, int B, int S, int H, int hd, bool causal,
                      float scale, cudaStream_t s) {
  CC_REQUIRE(B > 0 && S > 0 && H > 0 && hd > 0 && hd % 2 == 0, CC_ESHAPE, "attention_bwd: B=%d S=%d H=%d hd=%d", B, S, H, hd);
  CC_REQUIRE(ld % 2 == 0 && ldo % 2 == 0 && ldd % 2 == 0, CC_EALIGN, "attention_bwd: strides must be even");
  // tensor-core kernel: 16-byte aligned rows, head dim 48/64/96/128, S <= 160 (within shared memory)
  const bool aligned = ld % 8 == 0 && ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(q) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(k) & 15) == 0 && (reinterpret_cast<uintptr_t>(v) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(d_o) & 15) == 0 && hd % 8 == 0;
  if (aligned && !attn_bwd_scalar_forced()) {
    int st = CC_ESHAPE;
#define CC_ATB_CASE(HD, NT)                                                                                          \
  if (st == CC_ESHAPE && hd == HD && S <= 8 * NT)                                                                    \
    st = launch_attn_bwd_mma<HD, NT>(q, k, v, ld, d_o, ldo, dq, dk, dv, ldd, B, S, H, causal, scale, s);
    CC_ATB_CASE(64, 8)
    CC_ATB_CASE(64, 14)
    CC_ATB_CASE(64, 20)
    CC_ATB_CASE(48, 8)
    CC_ATB_CASE(48, 14)
    CC_ATB_CASE(96, 8)
    CC_ATB_CASE(96, 14)
    CC_ATB_CASE(128, 8)
    CC_ATB_CASE(128, 14)
#undef CC_ATB_CASE
    if (st != CC_ESHAPE) return st;
  }
  CC_REQUIRE(S <= 32 * ATB_MAXJ, CC_ESHAPE, "attention_bwd: sequence length %d exceeds %d", S, 32 * ATB_MAXJ);
  const size_t tiles = 4 * static_cast<size_t>(S) * (hd + 2) * sizeof(__half);
  const size_t smem = ((tiles + 15) & ~static_cast<size_t>(15)) + static_cast<size_t>(S) * (S + 1) * sizeof(float);
  CC_REQUIRE(smem <= 227 * 1024, CC_ESHAPE,
             "attention_bwd: S=%d hd=%d needs %zu bytes of shared memory (max 232448)", S, hd, smem);
  static size_t configured = 0;
  if (smem > configured) {
    CC_CUDA(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = 227 * 1024;
  }
  attention_bwd_kernel<<<B * H, ATB_THREADS, smem, s>>>(q, k, v, ld, d_o, ldo, dq, dk, dv, ldd, S, H, hd, causal ? 1 : 0,
                                                       scale);
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

int count_valid_run(const int32_t* targets, int n, int* n_valid, cudaStream_t s) {
  count_valid_kernel<<<1, 1024, 0, s>>>(targets, n, n_valid);
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

int ce_loss_run(const float* logits, int64_t ld, int V, const int32_t* targets, int rows, const int* n_valid,
                float loss_scale, float* row_loss, __half* dlogits, int64_t ld_d, cudaStream_t s) {
  CC_REQUIRE(rows > 0 && V > 0 && ld >= V && ld_d >= V, CC_ESHAPE, "ce_loss: rows=%d V=%d ld=%lld ld_d=%lld", rows, V,
             (long long)ld, (long long)ld_d);
  ce_loss_kernel<<<rows, CE_THREADS, 0, s>>>(logits, ld, V, targets, n_valid, loss_scale, row_loss, dlogits, ld_d);
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

int loss_reduce_run(const float* row_loss, int n, const int* n_valid, float* loss, cudaStream_t s) {
  loss_reduce_kernel<<<1, CE_THREADS, 0, s>>>(row_loss, n, n_valid, loss);
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

int train_embed_run(const int32_t* tokens, int B, int Tt, int K, const float* prefix, int64_t prefix_ld,
                    const float* wte, const float* wpe, float* h, int32_t* targets, int d, int V, cudaStream_t s) {
  train_embed_kernel<<<B * (K + Tt), 256, 0, s>>>(tokens, Tt, K, prefix, prefix_ld, wte, wpe, h, targets, d, V);
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

int count_nonfinite_run(const float* x, int64_t n, int* count, cudaStream_t s) {
  if (n <= 0) return CC_OK;
  count_nonfinite_kernel<<<grid_for(n, 256), 256, 0, s>>>(x, n, count);
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

int adamw_run(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
              float weight_decay, int step, cudaStream_t s) {
  CC_REQUIRE(step >= 1, CC_EINVAL, "adamw: step %d must be >= 1", step);
  if (n <= 0) return CC_OK;
  const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.f - powf(beta2, static_cast<float>(step));
  adamw_kernel<<<grid_for(n, 256), 256, 0, s>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, sqrtf(bc2));
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

}  // namespace cc
