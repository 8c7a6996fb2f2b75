// Element-wise glue kernels of the path: dtype conversion at the boundary, weight packing, ViT patch gathering and
// embedding (+ln_pre), mapper constant rows, GPT-2 input embedding. All HBM-bound, vectorised where the layout allows;
// grids are sized in multiples of the SM count and grid-stride over the data.
#include <algorithm>

#include "common.h"
#include "ptx.cuh"

namespace cc {
namespace {

inline int grid_for(long long work_items, int threads) {
  const long long blocks = (work_items + threads - 1) / threads;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  return static_cast<int>(std::max<long long>(1, std::min(blocks, cap)));
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------- conversions
template <typename SRC>
__global__ void to_f16_kernel(const SRC* __restrict__ src, __half* __restrict__ dst, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += stride)
    dst[i] = __float2half_rn(static_cast<float>(src[i]));
}
__global__ void f32_to_f16_vec_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, long long n4) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
    const float4 v = src[i];
    uint2 o;
    o.x = pack_half2(v.x, v.y);
    o.y = pack_half2(v.z, v.w);
    dst[i] = o;
  }
}
template <typename SRC>
__global__ void to_f32_kernel(const SRC* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += stride)
    dst[i] = static_cast<float>(src[i]);
}
template <typename DST>
__global__ void from_f32_rows_kernel(const float* __restrict__ src, long long src_ld, DST* __restrict__ dst, int rows,
                                     int cols) {
  const long long n = static_cast<long long>(rows) * cols;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += stride) {
    const long long r = i / cols;
    const int c = static_cast<int>(i - r * cols);
    const float v = src[r * src_ld + c];
    if constexpr (sizeof(DST) == 2) dst[i] = __float2half_rn(v);
    else dst[i] = v;
  }
}

// weights: fp32 [rows, cols] -> fp16 [rows_out, ld_out], optionally transposed, zero padded columns.
// 32x32 shared-memory tile so both the read and the write are coalesced in the transposed case.
__global__ void pack_weight_kernel(const float* __restrict__ src, int rows, int cols, int transpose,
                                   __half* __restrict__ dst, long long ld_out) {
  __shared__ float tile[32][33];
  const int out_rows = transpose ? cols : rows;
  const int out_cols = transpose ? rows : cols;
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;  // bx: out col tile, by: out row tile
  if (!transpose) {
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
      const int r = by + j, c = bx + threadIdx.x;
      if (r < out_rows && c < ld_out)
        dst[static_cast<long long>(r) * ld_out + c] =
            c < out_cols ? __float2half_rn(src[static_cast<long long>(r) * cols + c]) : __float2half_rn(0.f);
    }
    return;
  }
  // transposed: out[r][c] = src[c][r]; read src rows (c) along src cols (r)
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int sc = by + threadIdx.x;  // src col = out row
    const int sr = bx + j;            // src row = out col
    tile[j][threadIdx.x] = (sr < rows && sc < cols) ? src[static_cast<long long>(sr) * cols + sc] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = by + j, c = bx + threadIdx.x;
    if (r < out_rows && c < ld_out) dst[static_cast<long long>(r) * ld_out + c] = __float2half_rn(tile[threadIdx.x][j]);
  }
}

// ---------------------------------------------------------------- ViT
// Non-overlapping 14x14 patches: row (b, py, px) of the GEMM operand gathers 3 x 14 x 14 pixels, column order
// (c, ky, kx) = the flattening of conv1.weight [w, 3, 14, 14]; columns >= 588 are zero padding (16-byte rows for TMA).
// PATCH > 0: the patch size as a compile-time constant (the divisions of the index arithmetic become multiplications:
// ViT-L/14 and ViT-B/16 / B/32 take this path), 0: run-time patch size.
template <typename SRC, int PATCH>
__global__ void __launch_bounds__(256)
vit_im2col_kernel(const SRC* __restrict__ px, __half* __restrict__ out, int img, int patch_rt, int k_pad_rt) {
  // one CTA per (image, patch row): 32-bit index arithmetic only, the output rows of the CTA are contiguous
  const int patch = PATCH > 0 ? PATCH : patch_rt;
  const int k_pad = PATCH > 0 ? (3 * PATCH * PATCH + 7) / 8 * 8 : k_pad_rt;  // the host checks that this is the row length
  const int grid_w = img / patch;
  const int pp = patch * patch, kk = 3 * pp;
  const int b = blockIdx.x / grid_w, pyi = blockIdx.x - b * grid_w;
  const SRC* src = px + static_cast<long long>(b) * 3 * img * img;
  __half* dst = out + static_cast<long long>(blockIdx.x) * grid_w * k_pad;
  const int n = grid_w * k_pad;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int pxi = i / k_pad, col = i - pxi * k_pad;
    float v = 0.f;
    if (col < kk) {
      const int c = col / pp, rem = col - c * pp;
      const int ky = rem / patch, kx = rem - ky * patch;
      v = static_cast<float>(src[(c * img + pyi * patch + ky) * img + pxi * patch + kx]);
    }
    dst[i] = __float2half_rn(v);
  }
}

// h[b, 0] = LN(class_emb + pos[0]); h[b, 1+i] = LN(patches[b, i] + pos[1+i]).  One warp per output row.
__global__ void __launch_bounds__(128)
vit_embed_lnpre_kernel(const float* __restrict__ patches, const float* __restrict__ cls, const float* __restrict__ pos,
                       const float* __restrict__ g, const float* __restrict__ bta, float* __restrict__ h, int B, int T,
                       int w, float eps) {
  const long long row = blockIdx.x * 4LL + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= static_cast<long long>(B) * T) return;
  const int t = static_cast<int>(row % T);
  const long long b = row / T;
  const float* src = t == 0 ? cls : patches + (b * (T - 1) + (t - 1)) * w;
  const float* pr = pos + static_cast<long long>(t) * w;
  float* dst = h + row * w;
  float s = 0.f;
  for (int c = lane; c < w; c += 32) {
    const float v = src[c] + pr[c];
    dst[c] = v;
    s += v;
  }
  const float mean = warp_sum_f(s) / w;
  float q = 0.f;
  for (int c = lane; c < w; c += 32) {
    const float dlt = dst[c] - mean;
    q += dlt * dlt;
  }
  const float rstd = rsqrtf(warp_sum_f(q) / w + eps);
  for (int c = lane; c < w; c += 32) dst[c] = (dst[c] - mean) * rstd * g[c] + bta[c];
}

__global__ void l2_normalize_kernel(float* __restrict__ x, int rows, int cols) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* xr = x + static_cast<long long>(row) * cols;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s += xr[c] * xr[c];
  const float inv = 1.f / sqrtf(warp_sum_f(s));
  for (int c = lane; c < cols; c += 32) xr[c] *= inv;
}

// ---------------------------------------------------------------- mapper
// h: [B, S=P+K, d] fp32. Rows P.. get prefix_const (mapper.py:125-126); rows < P get += pos_emb (windowed, :153).
__global__ void mapper_fill_const_kernel(float* __restrict__ h, const float* __restrict__ prefix_const,
                                         const float* __restrict__ pos_emb, int B, int P, int K, int d) {
  const int S = P + K;
  const long long n = static_cast<long long>(B) * S * d;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += stride) {
    const int c = static_cast<int>(i % d);
    const int srow = static_cast<int>((i / d) % S);
    if (srow >= P) h[i] = prefix_const[static_cast<long long>(srow - P) * d + c];
    else if (pos_emb != nullptr) h[i] += pos_emb[static_cast<long long>(srow) * d + c];
  }
}

// ---------------------------------------------------------------- window tiling of a decoded image (encoders/clip.py:60-82)
// tiles[(ty * n + tx), c, i, j] = image[c, ty * step + i, tx * step + j]: the reference's
// tensor.unfold(1, p, step).unfold(2, p, step) on the decoded [3, S, S] image, written tile-major so every tile is an
// ordinary [3, p, p] image for the resize / normalise step that follows. Pure index arithmetic, one pass over the pixels.
__global__ void tile_image_kernel(const float* __restrict__ img, int S, int n, int p, int step, float* __restrict__ tiles) {
  const long long total = static_cast<long long>(n) * n * 3 * p * p;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += stride) {
    const int j = static_cast<int>(i % p);
    long long r = i / p;
    const int ii = static_cast<int>(r % p);
    r /= p;
    const int c = static_cast<int>(r % 3);
    const int tile = static_cast<int>(r / 3);
    const int ty = tile / n, tx = tile - ty * n;
    tiles[i] = img[(static_cast<long long>(c) * S + ty * step + ii) * S + tx * step + j];
  }
}

// ---------------------------------------------------------------- GPT-2 input embedding
template <typename SRC>
__global__ void gpt2_embed_prefix_kernel(const SRC* __restrict__ e, const float* __restrict__ wpe, float* __restrict__ h,
                                         int B, int T, int d, int pos0) {
  const long long n = static_cast<long long>(B) * T * d;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += stride) {
    const int c = static_cast<int>(i % d);
    const int t = static_cast<int>((i / d) % T);
    h[i] = static_cast<float>(e[i]) + wpe[static_cast<long long>(pos0 + t) * d + c];
  }
}
__global__ void gpt2_embed_tokens_kernel(const int32_t* __restrict__ tokens, long long tok_stride,
                                         const float* __restrict__ wte, const float* __restrict__ wpe,
                                         float* __restrict__ h, int n, int d, int pos, int V) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = static_cast<long long>(n) * (d / 4);
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const float4* wpe4 = reinterpret_cast<const float4*>(wpe + static_cast<long long>(pos) * d);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += stride) {
    const int c4 = static_cast<int>(i % (d / 4));
    const long long r = i / (d / 4);
    int tok = tokens[r * tok_stride];
    tok = tok < 0 ? 0 : (tok >= V ? V - 1 : tok);
    const float4 a = reinterpret_cast<const float4*>(wte + static_cast<long long>(tok) * d)[c4];
    const float4 p = wpe4[c4];
    reinterpret_cast<float4*>(h + r * d)[c4] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
  }
}
template <typename DST>
__global__ void gather_rows_kernel(const int32_t* __restrict__ ids, const float* __restrict__ table, DST* __restrict__ out,
                                   int n, int d, int V) {
  const long long total = static_cast<long long>(n) * d;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += stride) {
    const int c = static_cast<int>(i % d);
    const long long r = i / d;
    int tok = ids[r];
    tok = tok < 0 ? 0 : (tok >= V ? V - 1 : tok);
    const float v = table[static_cast<long long>(tok) * d + c];
    if constexpr (sizeof(DST) == 2) out[i] = __float2half_rn(v);
    else out[i] = v;
  }
}

}  // namespace

int convert_to_f16_run(const void* src, int src_dtype, __half* dst, int64_t n, cudaStream_t s) {
  if (n <= 0) return CC_OK;
  if (src_dtype == CC_F32) {
    if ((n % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0)) {
      f32_to_f16_vec_kernel<<<grid_for(n / 4, 256), 256, 0, s>>>(static_cast<const float4*>(src),
                                                               reinterpret_cast<uint2*>(dst), n / 4);
    } else {
      to_f16_kernel<float><<<grid_for(n, 256), 256, 0, s>>>(static_cast<const float*>(src), dst, n);
    }
  } else if (src_dtype == CC_F16) {
    CC_CUDA(cudaMemcpyAsync(dst, src, static_cast<size_t>(n) * 2, cudaMemcpyDeviceToDevice, s));
    return CC_OK;
  } else {
    set_error("convert: unknown dtype %d", src_dtype);
    return CC_EINVAL;
  }
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

int convert_to_f32_run(const void* src, int src_dtype, float* dst, int64_t n, cudaStream_t s) {
  if (n <= 0) return CC_OK;
  if (src_dtype == CC_F32) {
    CC_CUDA(cudaMemcpyAsync(dst, src, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToDevice, s));
    return CC_OK;
  }
  CC_REQUIRE(src_dtype == CC_F16, CC_EINVAL, "convert: unknown dtype %d", src_dtype);
  to_f32_kernel<__half><<<grid_for(n, 256), 256, 0, s>>>(static_cast<const __half*>(src), dst, n);
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

int convert_from_f32_run(const float* src, int64_t src_ld, void* dst, int dst_dtype, int rows, int cols, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return CC_OK;
  const long long n = static_cast<long long>(rows) * cols;
  if (dst_dtype == CC_F32)
    from_f32_rows_kernel<float><<<grid_for(n, 256), 256, 0, s>>>(src, src_ld, static_cast<float*>(dst), rows, cols);
  else if (dst_dtype == CC_F16)
    from_f32_rows_kernel<__half><<<grid_for(n, 256), 256, 0, s>>>(src, src_ld, static_cast<__half*>(dst), rows, cols);
  else {
    set_error("convert: unknown dtype %d", dst_dtype);
    return CC_EINVAL;
  }
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

int pack_weight_run(const float* src, int rows, int cols, bool transpose, __half* dst, int64_t ld_out, cudaStream_t s) {
  const int out_rows = transpose ? cols : rows;
  dim3 grid(static_cast<unsigned>((ld_out + 31) / 32), static_cast<unsigned>((out_rows + 31) / 32));
  pack_weight_kernel<<<grid, dim3(32, 8), 0, s>>>(src, rows, cols, transpose ? 1 : 0, dst, ld_out);
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

int vit_im2col_run(const void* pixels, int dtype, __half* out, int B, int img, int patch, int k_pad, cudaStream_t s) {
  const int gw = img / patch;
  if (dtype != CC_F32 && dtype != CC_F16) {
    set_error("vit: unknown pixel dtype %d", dtype);
    return CC_EINVAL;
  }
#define CC_IM2COL(SRC, P)                                                                                          \
  vit_im2col_kernel<SRC, P><<<B * gw, 256, 0, s>>>(static_cast<const SRC*>(pixels), out, img, patch, k_pad)
#define CC_IM2COL_DT(P)                    \
  do {                                     \
    if (dtype == CC_F32) CC_IM2COL(float, P); \
    else CC_IM2COL(__half, P);             \
  } while (0)
  const bool ct = k_pad == (3 * patch * patch + 7) / 8 * 8;  // the padded row length the constant-patch kernels assume
  if (ct && patch == 14) CC_IM2COL_DT(14);
  else if (ct && patch == 16) CC_IM2COL_DT(16);
  else if (ct && patch == 32) CC_IM2COL_DT(32);
  else CC_IM2COL_DT(0);
#undef CC_IM2COL_DT
#undef CC_IM2COL
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

int vit_embed_lnpre_run(const float* patches, const float* cls, const float* pos, const float* g, const float* b,
                        float* h, int B, int T, int w, float eps, cudaStream_t s) {
  const long long rows = static_cast<long long>(B) * T;
  vit_embed_lnpre_kernel<<<static_cast<unsigned>((rows + 3) / 4), 128, 0, s>>>(patches, cls, pos, g, b, h, B, T, w, eps);
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

int l2_normalize_run(float* x, int rows, int cols, cudaStream_t s) {
  l2_normalize_kernel<<<(rows + 3) / 4, 128, 0, s>>>(x, rows, cols);
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

int mapper_fill_const_run(float* h, const float* prefix_const, const float* pos_emb, int B, int P, int K, int d,
                          cudaStream_t s) {
  const long long n = static_cast<long long>(B) * (P + K) * d;
  mapper_fill_const_kernel<<<grid_for(n, 256), 256, 0, s>>>(h, prefix_const, pos_emb, B, P, K, d);
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

int tile_image_run(const float* img, int S, int n, int p, int step, float* tiles, cudaStream_t s) {
  CC_REQUIRE(S > 0 && n > 0 && p > 0 && step > 0 && (n - 1) * step + p <= S, CC_ESHAPE,
             "tile_image: %d x %d tiles of %d px at step %d do not fit a %d px image", n, n, p, step, S);
  const long long total = static_cast<long long>(n) * n * 3 * p * p;
  tile_image_kernel<<<grid_for(total, 256), 256, 0, s>>>(img, S, n, p, step, tiles);
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

int gpt2_embed_prefix_run(const void* embeds, int dtype, const float* wpe, float* h, int B, int T, int d, int pos0,
                          cudaStream_t s) {
  const long long n = static_cast<long long>(B) * T * d;
  if (dtype == CC_F32)
    gpt2_embed_prefix_kernel<float><<<grid_for(n, 256), 256, 0, s>>>(static_cast<const float*>(embeds), wpe, h, B, T, d, pos0);
  else if (dtype == CC_F16)
    gpt2_embed_prefix_kernel<__half><<<grid_for(n, 256), 256, 0, s>>>(static_cast<const __half*>(embeds), wpe, h, B, T, d, pos0);
  else {
    set_error("gpt2: unknown embeds dtype %d", dtype);
    return CC_EINVAL;
  }
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

int gpt2_embed_tokens_run(const int32_t* tokens, int64_t tok_stride, const float* wte, const float* wpe, float* h, int n,
                          int d, int pos, int V, cudaStream_t s) {
  const long long total = static_cast<long long>(n) * (d / 4);
  CC_CUDA(launch_pdl(gpt2_embed_tokens_kernel, dim3(grid_for(total, 256)), dim3(256), 0, s, tokens,
                     static_cast<long long>(tok_stride), wte, wpe, h, n, d, pos, V));
  return CC_OK;
}

int gather_rows_run(const int32_t* ids, const float* table, void* out, int out_dtype, int n, int d, int V,
                    cudaStream_t s) {
  const long long total = static_cast<long long>(n) * d;
  if (out_dtype == CC_F32)
    gather_rows_kernel<float><<<grid_for(total, 256), 256, 0, s>>>(ids, table, static_cast<float*>(out), n, d, V);
  else if (out_dtype == CC_F16)
    gather_rows_kernel<__half><<<grid_for(total, 256), 256, 0, s>>>(ids, table, static_cast<__half*>(out), n, d, V);
  else {
    set_error("gather: unknown dtype %d", out_dtype);
    return CC_EINVAL;
  }
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

}  // namespace cc
