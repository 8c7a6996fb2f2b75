// Attention kernels.
//
// (1) attn_fwd_kernel: fused softmax(Q K^T * scale [+ causal]) V for the three batched attention sites of the path —
//     ViT self-attention (S=257, hd=64; clip encode_image), the mapper's MultiHeadAttention
//     (clipcap/model/attention.py:17-43, S=P+K, hd=d/H) and GPT-2 prefill (causal, hd=64).  Flash-style: one warp owns
//     16 query rows, K/V stream through double-buffered shared memory in chunks of 64 keys (cp.async), scores never
//     leave registers (online softmax in fp32, exp2 with the scale folded in), P is re-used in registers as the A
//     operand of the P.V product. Tensor work uses warp-level mma.m16n8k16 (fp16 in, fp32 accumulate): the tiles here
//     (<= 257 keys x 64..128) are too small to amortise a TMEM round trip per head.
// (2) decode_attn_kernel: one query row per (sequence, head) against the fp16 KV cache — HBM-bound streaming of
//     K and V rows (128 B each) with 8 lanes per key, plus the in-place cache append of the step's own k,v.
// (3) kv_scatter_kernel: prefill K,V rows -> cache.
#include <algorithm>
#include <cstdlib>

#include "common.h"
#include "decode.h"
#include "ptx.cuh"
#include "mv_attn.cuh"

namespace cc {
namespace {

constexpr int KC = 64;  // keys per shared-memory chunk

template <int HD>
struct AttnSmem {
  static constexpr int kRowBytes = HD * 2 + 16;  // +16 B pad: conflict-free ldmatrix without swizzling
  static constexpr int kChunkBytes = KC * kRowBytes;
  static constexpr int kBytes = 4 * kChunkBytes;  // {K,V} x double buffer
};

template <int HD>
__device__ __forceinline__ void load_kv_chunk(uint32_t sk, uint32_t sv, const __half* kbase, const __half* vbase,
                                              long long ld, int key0, int S) {
  constexpr int CPR = HD / 8;  // 16-byte chunks per row
  for (int i = threadIdx.x; i < KC * CPR; i += blockDim.x) {
    const int r = i / CPR, c = i % CPR;
    const int key = key0 + r;
    const bool ok = key < S;
    const long long off = static_cast<long long>(ok ? key : 0) * ld + c * 8;
    const uint32_t so = r * AttnSmem<HD>::kRowBytes + c * 16;
    cp_async_16(sk + so, kbase + off, ok);
    cp_async_16(sv + so, vbase + off, ok);
  }
}

template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(256)
attn_fwd_kernel(const __half* __restrict__ q, const __half* __restrict__ k, const __half* __restrict__ v, long long ld,
                __half* __restrict__ o, long long ldo, int S, float scale_log2) {
  extern __shared__ __align__(16) uint8_t smem[];
  const uint32_t s0 = smem_u32(smem);
  const int nwarps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int h = blockIdx.y, b = blockIdx.z;
  pdl_launch_dependents();
  pdl_wait();
  const long long row_base = static_cast<long long>(b) * S;
  const __half* qh = q + row_base * ld + h * HD;
  const __half* kh = k + row_base * ld + h * HD;
  const __half* vh = v + row_base * ld + h * HD;

  const int q_lo = blockIdx.x * nwarps * 16;  // first query row of this CTA
  const int r0 = q_lo + warp * 16;            // first query row of this warp
  const bool warp_active = r0 < S;

  // keys this CTA needs: all of them, or (causal) up to its last query row
  int kv_end = S;
  if (CAUSAL) {
    const int q_hi = min(S, q_lo + nwarps * 16);
    kv_end = q_hi;
  }
  const int nchunks = (kv_end + KC - 1) / KC;

  // Q fragments (A operand), straight from global memory
  uint32_t qa[HD / 16][4];
  {
    const int ra = r0 + g, rb = r0 + g + 8;
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
      const int c = ks * 16 + 2 * t;
      qa[ks][0] = (warp_active && ra < S) ? *reinterpret_cast<const uint32_t*>(qh + ra * ld + c) : 0u;
      qa[ks][1] = (warp_active && rb < S) ? *reinterpret_cast<const uint32_t*>(qh + rb * ld + c) : 0u;
      qa[ks][2] = (warp_active && ra < S) ? *reinterpret_cast<const uint32_t*>(qh + ra * ld + c + 8) : 0u;
      qa[ks][3] = (warp_active && rb < S) ? *reinterpret_cast<const uint32_t*>(qh + rb * ld + c + 8) : 0u;
    }
  }

  float oacc[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f;
  float mrow[2] = {-INFINITY, -INFINITY};  // running max (raw scores) for rows g and g+8
  float lrow[2] = {0.f, 0.f};              // running sum of exp

  auto sK = [&](int buf) { return s0 + (buf * 2 + 0) * AttnSmem<HD>::kChunkBytes; };
  auto sV = [&](int buf) { return s0 + (buf * 2 + 1) * AttnSmem<HD>::kChunkBytes; };

  load_kv_chunk<HD>(sK(0), sV(0), kh, vh, ld, 0, S);
  cp_async_commit();

  for (int ch = 0; ch < nchunks; ++ch) {
    const int buf = ch & 1;
    if (ch + 1 < nchunks) {
      load_kv_chunk<HD>(sK(buf ^ 1), sV(buf ^ 1), kh, vh, ld, (ch + 1) * KC, S);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    const int key0 = ch * KC;
    const bool need = warp_active && (!CAUSAL || key0 <= r0 + 15);
    if (need) {
      // ---- S = Q K^T for 16 rows x 64 keys
      float sacc[KC / 8][4];
#pragma unroll
      for (int i = 0; i < KC / 8; ++i) sacc[i][0] = sacc[i][1] = sacc[i][2] = sacc[i][3] = 0.f;
      const uint32_t kb = sK(buf);
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks) {
#pragma unroll
        for (int np = 0; np < KC / 16; ++np) {
          // matrices: (keys +0..7, dims +0..7), (keys +0..7, dims +8..15), (keys +8..15, dims +0..7), (keys +8..15, dims +8..15)
          const int mi = lane >> 3, rr = lane & 7;
          const int key = np * 16 + (mi >> 1) * 8 + rr;
          const int dim = ks * 16 + (mi & 1) * 8;
          uint32_t bf[4];
          ldmatrix_x4(bf, kb + key * AttnSmem<HD>::kRowBytes + dim * 2);
          mma_16816(sacc[np * 2], qa[ks], bf[0], bf[1]);
          mma_16816(sacc[np * 2 + 1], qa[ks], bf[2], bf[3]);
        }
      }
      // ---- mask + online softmax
      float cmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < KC / 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = key0 + nt * 8 + 2 * t + (e & 1);
          const int row = r0 + g + (e >> 1) * 8;
          bool ok = key < S;
          if (CAUSAL) ok = ok && key <= row;
          if (!ok) sacc[nt][e] = -INFINITY;
          cmax[e >> 1] = fmaxf(cmax[e >> 1], sacc[nt][e]);
        }
      }
      float corr[2], mnew_s[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        cmax[r] = fmaxf(cmax[r], __shfl_xor_sync(0xffffffffu, cmax[r], 1));
        cmax[r] = fmaxf(cmax[r], __shfl_xor_sync(0xffffffffu, cmax[r], 2));
        const float mnew = fmaxf(mrow[r], cmax[r]);
        // rows that have seen no valid key yet keep m = -inf; guard the (-inf) - (-inf) case
        mnew_s[r] = (mnew == -INFINITY) ? 0.f : mnew * scale_log2;
        corr[r] = (mrow[r] == -INFINITY) ? 0.f : exp2f(mrow[r] * scale_log2 - mnew_s[r]);
        mrow[r] = mnew;
        lrow[r] *= corr[r];
      }
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) {
        oacc[i][0] *= corr[0];
        oacc[i][1] *= corr[0];
        oacc[i][2] *= corr[1];
        oacc[i][3] *= corr[1];
      }
      float csum[2] = {0.f, 0.f};
#pragma unroll
      for (int nt = 0; nt < KC / 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float p = exp2f(sacc[nt][e] * scale_log2 - mnew_s[e >> 1]);
          sacc[nt][e] = p;
          csum[e >> 1] += p;
        }
      }
      lrow[0] += csum[0];
      lrow[1] += csum[1];

      // ---- O += P V
      const uint32_t vb = sV(buf);
#pragma unroll
      for (int kk = 0; kk < KC / 16; ++kk) {
        uint32_t pa[4];
        pa[0] = pack_half2(sacc[2 * kk][0], sacc[2 * kk][1]);
        pa[1] = pack_half2(sacc[2 * kk][2], sacc[2 * kk][3]);
        pa[2] = pack_half2(sacc[2 * kk + 1][0], sacc[2 * kk + 1][1]);
        pa[3] = pack_half2(sacc[2 * kk + 1][2], sacc[2 * kk + 1][3]);
#pragma unroll
        for (int np = 0; np < HD / 16; ++np) {
          // transposed loads: (keys +0..7, dims d0..), (keys +8..15, dims d0..), (keys +0..7, dims d0+8..), (keys +8..15, dims d0+8..)
          const int mi = lane >> 3, rr = lane & 7;
          const int key = kk * 16 + (mi & 1) * 8 + rr;
          const int dim = np * 16 + (mi >> 1) * 8;
          uint32_t bf[4];
          ldmatrix_x4_trans(bf, vb + key * AttnSmem<HD>::kRowBytes + dim * 2);
          mma_16816(oacc[np * 2], pa, bf[0], bf[1]);
          mma_16816(oacc[np * 2 + 1], pa, bf[2], bf[3]);
        }
      }
    }
    __syncthreads();  // everyone done with `buf` before it is refilled two iterations later
  }

  if (warp_active) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 1);
      lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 2);
    }
    const float inv0 = lrow[0] > 0.f ? 1.f / lrow[0] : 0.f;
    const float inv1 = lrow[1] > 0.f ? 1.f / lrow[1] : 0.f;
    const int ra = r0 + g, rb = r0 + g + 8;
    __half* oh = o + row_base * ldo + h * HD;
#pragma unroll
    for (int nt = 0; nt < HD / 8; ++nt) {
      const int c = nt * 8 + 2 * t;
      if (ra < S) *reinterpret_cast<uint32_t*>(oh + ra * ldo + c) = pack_half2(oacc[nt][0] * inv0, oacc[nt][1] * inv0);
      if (rb < S) *reinterpret_cast<uint32_t*>(oh + rb * ldo + c) = pack_half2(oacc[nt][2] * inv1, oacc[nt][3] * inv1);
    }
  }
}

template <int HD, bool CAUSAL>
int launch_attn(const __half* q, const __half* k, const __half* v, int64_t ld, __half* o, int64_t ldo, int B, int S,
                int H, float scale, cudaStream_t s) {
  auto kern = attn_fwd_kernel<HD, CAUSAL>;
  CC_OPT_IN_SMEM(kern, AttnSmem<HD>::kBytes);
  const int tiles = (S + 15) / 16;
  const int nblk = (tiles + 7) / 8;
  const int nw = (tiles + nblk - 1) / nblk;
  dim3 grid(nblk, H, B);
  CC_CUDA(launch_pdl(kern, grid, dim3(nw * 32), AttnSmem<HD>::kBytes, s, q, k, v, static_cast<long long>(ld), o,
                     static_cast<long long>(ldo), S, scale * 1.4426950408889634f));
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

// ------------------------------------------------------------------ decode attention over the KV cache (hd = 64)
// One warp per (sequence, head). A key/value row is 128 B = 8 lanes x 16 B, so a warp covers 4 keys per load
// instruction; each pass of the loop issues 4 K loads + 4 V loads per lane (16 keys, 4 KB in flight per warp) before
// any arithmetic, and keeps a single online-softmax state, so the dependent chain is ceil(T/16) memory round trips.
// The step's own k,v come straight from the qkv row (and are appended to the cache for later steps).
constexpr int DEC_WARPS = 4;
constexpr int DEC_KEYS = 16;  // keys per loop pass

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __half2* hp = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(hp[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

__global__ void __launch_bounds__(DEC_WARPS * 32, 8)
decode_attn_kernel(const __half* __restrict__ qkv, __half* __restrict__ kcache, __half* __restrict__ vcache,
                   const int32_t* __restrict__ anc, __half* __restrict__ o, int nseq, int H, int t_max, int pos,
                   float scale_log2) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x * DEC_WARPS + warp;
  pdl_launch_dependents();
  pdl_wait();
  if (pair >= nseq * H) return;
  const int seq = pair / H, h = pair % H;
  const int d = H * 64;
  const int kq = lane >> 3;  // key slot 0..3 within a load
  const int c = lane & 7;    // 16-byte chunk (8 dims) of the head row

  const __half* qrow = qkv + static_cast<long long>(seq) * 3 * d + h * 64;
  const __half* knew = qrow + d + c * 8;
  const __half* vnew = qrow + 2 * d + c * 8;
  // append this step's k, v (lanes 0..7 copy k, 8..15 copy v)
  {
    const long long dst = ((static_cast<long long>(seq) * H + h) * t_max + pos) * 64 + kv_chunk(pos, c) * 8;
    if (kq == 0) *reinterpret_cast<uint4*>(kcache + dst) = *reinterpret_cast<const uint4*>(knew);
    else if (kq == 1) *reinterpret_cast<uint4*>(vcache + dst) = *reinterpret_cast<const uint4*>(vnew);
  }

  float qf[8];
  unpack8(*reinterpret_cast<const uint4*>(qrow + c * 8), qf);
  const int T = pos + 1;
  const int32_t* arow = anc ? anc + static_cast<long long>(seq) * t_max : nullptr;
  const long long own = (static_cast<long long>(seq) * H + h) * t_max;

  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  float mx = -INFINITY, lsum = 0.f;

  for (int t0 = 0; t0 < T; t0 += DEC_KEYS) {
    uint4 kk[4], vv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int tt = t0 + 4 * j + kq;
      if (tt < pos) {
        const long long base = arow ? (static_cast<long long>(arow[tt]) * H + h) * t_max : own;
        kk[j] = *reinterpret_cast<const uint4*>(kcache + (base + tt) * 64 + kv_chunk(tt, c) * 8);
        vv[j] = *reinterpret_cast<const uint4*>(vcache + (base + tt) * 64 + kv_chunk(tt, c) * 8);
      } else if (tt == pos) {
        kk[j] = *reinterpret_cast<const uint4*>(knew);
        vv[j] = *reinterpret_cast<const uint4*>(vnew);
      } else {
        kk[j] = make_uint4(0u, 0u, 0u, 0u);
        vv[j] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    float dot[4];
    float bm = -INFINITY;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float kf[8];
      unpack8(kk[j], kf);
      float dd = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) dd += qf[i] * kf[i];
      dd += __shfl_xor_sync(0xffffffffu, dd, 1);
      dd += __shfl_xor_sync(0xffffffffu, dd, 2);
      dd += __shfl_xor_sync(0xffffffffu, dd, 4);
      dot[j] = (t0 + 4 * j + kq < T) ? dd : -INFINITY;
      bm = fmaxf(bm, dot[j]);
    }
    bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 8));
    bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 16));
    const float nm = fmaxf(mx, bm);  // finite: every pass holds at least one valid key
    const float corr = exp2f((mx - nm) * scale_log2);
    mx = nm;
    lsum *= corr;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] *= corr;
    const float nms = nm * scale_log2;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float p = exp2f(dot[j] * scale_log2 - nms);  // exp2(-inf) = 0 for masked keys
      lsum += p;
      float vf[8];
      unpack8(vv[j], vf);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += p * vf[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
  }
  lsum += __shfl_xor_sync(0xffffffffu, lsum, 8);
  lsum += __shfl_xor_sync(0xffffffffu, lsum, 16);
  if (kq == 0) {
    const float inv = 1.f / lsum;
    uint4 out;
    out.x = pack_half2(acc[0] * inv, acc[1] * inv);
    out.y = pack_half2(acc[2] * inv, acc[3] * inv);
    out.z = pack_half2(acc[4] * inv, acc[5] * inv);
    out.w = pack_half2(acc[6] * inv, acc[7] * inv);
    *reinterpret_cast<uint4*>(o + static_cast<long long>(seq) * d + h * 64 + c * 8) = out;
  }
}

// Beam search: the `beam` sequences of one image share their whole prefix — positions 0 .. shared_len-1 live in ONE cache
// slot (the image's prefill slot, decode.cu beam_init / beam_step) — so one CTA takes an (image, head) with one warp per
// beam: the shared K / V rows are pulled into shared memory once with two bulk copies (instead of once per beam from
// L2 / HBM), the few generated positions come through the ancestry table as in decode_attn_kernel. The chunks are walked
// from the newest key down, so the ancestry loads are in flight while the bulk copy lands. Same lane layout and
// online-softmax arithmetic per row as decode_attn_kernel (the summation order over chunks differs).
__global__ void __launch_bounds__(kMaxBeam * 32)
decode_attn_beam_kernel(const __half* __restrict__ qkv, __half* __restrict__ kcache, __half* __restrict__ vcache,
                        const int32_t* __restrict__ anc, __half* __restrict__ o, int beam, int H, int t_max, int pos,
                        int shared_len, float scale_log2) {
  extern __shared__ __align__(128) uint8_t bsm[];
  __shared__ __align__(8) unsigned long long bbar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int img = blockIdx.x / H, h = blockIdx.x % H;
  const uint32_t bar = smem_u32(&bbar);
  const uint32_t kbuf = smem_u32(bsm), vbuf = kbuf + shared_len * 128;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_launch_dependents();
  pdl_wait();
  const int d = H * 64;
  const int seq = img * beam + warp;  // warp == beam index (blockDim.x == beam * 32)
  const int kq = lane >> 3;
  const int c = lane & 7;
  const int32_t* arow = anc + static_cast<long long>(seq) * t_max;
  if (threadIdx.x == 0 && shared_len > 0) {
    const long long src = ((static_cast<long long>(arow[0]) * H + h) * t_max) * 64;  // slot of position 0 = the shared slot
    mbar_arrive_expect_tx(bar, 2u * shared_len * 128u);
    bulk_load(kbuf, kcache + src, shared_len * 128u, bar);
    bulk_load(vbuf, vcache + src, shared_len * 128u, bar);
  }
  const __half* qrow = qkv + static_cast<long long>(seq) * 3 * d + h * 64;
  const uint4 knew = *reinterpret_cast<const uint4*>(qrow + d + c * 8);
  const uint4 vnew = *reinterpret_cast<const uint4*>(qrow + 2 * d + c * 8);
  {  // append this step's k, v to the row's own slot
    const long long dst = ((static_cast<long long>(seq) * H + h) * t_max + pos) * 64 + kv_chunk(pos, c) * 8;
    if (kq == 0) *reinterpret_cast<uint4*>(kcache + dst) = knew;
    else if (kq == 1) *reinterpret_cast<uint4*>(vcache + dst) = vnew;
  }
  float qf[8];
  unpack8(*reinterpret_cast<const uint4*>(qrow + c * 8), qf);
  const int T = pos + 1;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  float mx = -INFINITY, lsum = 0.f;
  bool landed = shared_len == 0;

  for (int t0 = ((T - 1) / DEC_KEYS) * DEC_KEYS; t0 >= 0; t0 -= DEC_KEYS) {
    if (!landed && t0 < shared_len) {  // warp-uniform: first chunk that needs the shared rows
      mbar_wait(bar, 0);
      landed = true;
    }
    uint4 kk[4], vv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int tt = t0 + 4 * j + kq;
      if (tt < shared_len) {
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(kk[j].x), "=r"(kk[j].y), "=r"(kk[j].z), "=r"(kk[j].w)
                     : "r"(kbuf + tt * 128 + kv_chunk(tt, c) * 16));
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(vv[j].x), "=r"(vv[j].y), "=r"(vv[j].z), "=r"(vv[j].w)
                     : "r"(vbuf + tt * 128 + kv_chunk(tt, c) * 16));
      } else if (tt < pos) {
        const long long base = ((static_cast<long long>(arow[tt]) * H + h) * t_max + tt) * 64 + kv_chunk(tt, c) * 8;
        kk[j] = *reinterpret_cast<const uint4*>(kcache + base);
        vv[j] = *reinterpret_cast<const uint4*>(vcache + base);
      } else if (tt == pos) {
        kk[j] = knew;
        vv[j] = vnew;
      } else {
        kk[j] = make_uint4(0u, 0u, 0u, 0u);
        vv[j] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    float dot[4];
    float bm = -INFINITY;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float kf[8];
      unpack8(kk[j], kf);
      float dd = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) dd += qf[i] * kf[i];
      dd += __shfl_xor_sync(0xffffffffu, dd, 1);
      dd += __shfl_xor_sync(0xffffffffu, dd, 2);
      dd += __shfl_xor_sync(0xffffffffu, dd, 4);
      dot[j] = (t0 + 4 * j + kq < T) ? dd : -INFINITY;
      bm = fmaxf(bm, dot[j]);
    }
    bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 8));
    bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 16));
    const float nm = fmaxf(mx, bm);  // finite: every chunk visited holds at least one valid key
    const float corr = exp2f((mx - nm) * scale_log2);
    mx = nm;
    lsum *= corr;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] *= corr;
    const float nms = nm * scale_log2;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float p = exp2f(dot[j] * scale_log2 - nms);
      lsum += p;
      float vf[8];
      unpack8(vv[j], vf);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += p * vf[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
  }
  lsum += __shfl_xor_sync(0xffffffffu, lsum, 8);
  lsum += __shfl_xor_sync(0xffffffffu, lsum, 16);
  if (kq == 0) {
    const float inv = 1.f / lsum;
    uint4 out;
    out.x = pack_half2(acc[0] * inv, acc[1] * inv);
    out.y = pack_half2(acc[2] * inv, acc[3] * inv);
    out.z = pack_half2(acc[4] * inv, acc[5] * inv);
    out.w = pack_half2(acc[6] * inv, acc[7] * inv);
    *reinterpret_cast<uint4*>(o + static_cast<long long>(seq) * d + h * 64 + c * 8) = out;
  }
}

// Greedy decode (no ancestry table): the keys and values of one (sequence, head) are contiguous in the cache, so each
// warp pulls them into shared memory with two bulk copies (all bytes in flight at once, completion on the warp's own
// mbarrier) instead of walking them 16 keys per round trip. Same arithmetic and lane layout as decode_attn_kernel.
__global__ void __launch_bounds__(DEC_WARPS * 32)
decode_attn_bulk_kernel(const __half* __restrict__ qkv, __half* __restrict__ kcache, __half* __restrict__ vcache,
                        __half* __restrict__ o, int nseq, int H, int t_max, int pos, float scale_log2) {
  extern __shared__ __align__(128) uint8_t dsm[];
  __shared__ __align__(8) unsigned long long bars[DEC_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x * DEC_WARPS + warp;
  const uint32_t bar = smem_u32(&bars[warp]);
  const uint32_t row_bytes = 128;
  const uint32_t kbuf = smem_u32(dsm) + warp * 2 * pos * row_bytes;  // [pos] K rows, then [pos] V rows
  const uint32_t vbuf = kbuf + pos * row_bytes;
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  pdl_launch_dependents();
  const bool live = pair < nseq * H;
  const int seq = pair / H, h = pair % H;
  const long long own = (static_cast<long long>(seq) * H + h) * t_max;
  // The cache rows of positions < pos were written by EARLIER decode steps / the prefill (hundreds of kernels back in the
  // stream), never by the predecessor kernel, so they start streaming before griddepcontrol.wait; q and this step's k, v
  // (the predecessor's output) are only touched after it.
  if (live && lane == 0 && pos > 0) {
    mbar_arrive_expect_tx(bar, 2u * pos * row_bytes);
    bulk_load(kbuf, kcache + own * 64, pos * row_bytes, bar);
    bulk_load(vbuf, vcache + own * 64, pos * row_bytes, bar);
  }
  pdl_wait();
  if (!live) return;
  const int d = H * 64;
  const int kq = lane >> 3;
  const int c = lane & 7;
  const __half* qrow = qkv + static_cast<long long>(seq) * 3 * d + h * 64;
  const uint4 knew = *reinterpret_cast<const uint4*>(qrow + d + c * 8);
  const uint4 vnew = *reinterpret_cast<const uint4*>(qrow + 2 * d + c * 8);
  {  // append this step's k, v for later steps
    const long long dst = (own + pos) * 64 + kv_chunk(pos, c) * 8;
    if (kq == 0) *reinterpret_cast<uint4*>(kcache + dst) = knew;
    else if (kq == 1) *reinterpret_cast<uint4*>(vcache + dst) = vnew;
  }
  float qf[8];
  unpack8(*reinterpret_cast<const uint4*>(qrow + c * 8), qf);
  const int T = pos + 1;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  float mx = -INFINITY, lsum = 0.f;
  if (pos > 0) mbar_wait(bar, 0);

  for (int t0 = 0; t0 < T; t0 += DEC_KEYS) {
    uint4 kk[4], vv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int tt = t0 + 4 * j + kq;
      if (tt < pos) {
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(kk[j].x), "=r"(kk[j].y), "=r"(kk[j].z), "=r"(kk[j].w)
                     : "r"(kbuf + tt * row_bytes + kv_chunk(tt, c) * 16));
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(vv[j].x), "=r"(vv[j].y), "=r"(vv[j].z), "=r"(vv[j].w)
                     : "r"(vbuf + tt * row_bytes + kv_chunk(tt, c) * 16));
      } else if (tt == pos) {
        kk[j] = knew;
        vv[j] = vnew;
      } else {
        kk[j] = make_uint4(0u, 0u, 0u, 0u);
        vv[j] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    float dot[4];
    float bm = -INFINITY;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float kf[8];
      unpack8(kk[j], kf);
      float dd = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) dd += qf[i] * kf[i];
      dd += __shfl_xor_sync(0xffffffffu, dd, 1);
      dd += __shfl_xor_sync(0xffffffffu, dd, 2);
      dd += __shfl_xor_sync(0xffffffffu, dd, 4);
      dot[j] = (t0 + 4 * j + kq < T) ? dd : -INFINITY;
      bm = fmaxf(bm, dot[j]);
    }
    bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 8));
    bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 16));
    const float nm = fmaxf(mx, bm);
    const float corr = exp2f((mx - nm) * scale_log2);
    mx = nm;
    lsum *= corr;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] *= corr;
    const float nms = nm * scale_log2;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float p = exp2f(dot[j] * scale_log2 - nms);
      lsum += p;
      float vf[8];
      unpack8(vv[j], vf);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += p * vf[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
  }
  lsum += __shfl_xor_sync(0xffffffffu, lsum, 8);
  lsum += __shfl_xor_sync(0xffffffffu, lsum, 16);
  if (kq == 0) {
    const float inv = 1.f / lsum;
    uint4 out;
    out.x = pack_half2(acc[0] * inv, acc[1] * inv);
    out.y = pack_half2(acc[2] * inv, acc[3] * inv);
    out.z = pack_half2(acc[4] * inv, acc[5] * inv);
    out.w = pack_half2(acc[6] * inv, acc[7] * inv);
    *reinterpret_cast<uint4*>(o + static_cast<long long>(seq) * d + h * 64 + c * 8) = out;
  }
}

// ------------------------------------------------------------------ decode attention on the warp-level tensor-core path
// One warp per (sequence, head): the two contractions are matrix-vector products issued as mma.m16n8k16 with the query /
// probability row in row 0 of the A operand (the other 15 rows are zero: the tensor pipe is idle in these kernels,
// instruction issue is what they run out of — measured 1432 warp instructions per pair in the FMA kernel, 692 with mv2_block):
//   s[1 x T]  = q[1 x 64] K^T      B = K rows as stored ([key][dim] is the col-major operand): ldmatrix.x4 per 8 keys x 32 dims
//   o[1 x 64] = p[1 x T] V         A = p re-packed from the score accumulators in registers, B = V through ldmatrix.trans
// The rotated chunk order of the cache rows (kv_chunk) makes every ldmatrix conflict-free after a plain bulk copy.
struct MvState {
  float acc[8][4];  // o accumulators: tile n holds dims 8 n + 2 (lane & 3) (+1) in [0], [1] on lanes 0..3
  float mx, lsum;
};

__device__ __forceinline__ void mv_init(MvState& st) {
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int i = 0; i < 4; ++i) st.acc[n][i] = 0.f;
  st.mx = -INFINITY;
  st.lsum = 0.f;
}

// A fragments of q: word (lane & 3) of each 16-byte chunk, lanes 0..3 (row 0 of the 16-row operand) only
__device__ __forceinline__ void mv_load_q(const __half* qrow, int lane, uint32_t (&qa)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) qa[i] = lane < 4 ? reinterpret_cast<const uint32_t*>(qrow)[4 * i + lane] : 0u;
}

// One block of up to 8 NT keys (NT = 8: 64, NT = 4: 32 — fewer live registers) with an online-softmax update. kbuf / vbuf: shared-memory rows of the block, row r holding
// cache position t0 + r (which fixes its chunk rotation); only the nvalid (1 .. 8 NT) real rows exist in shared memory — the
// products run over nvalid rounded up to 16 keys, and every operand fetch of a row >= nvalid is pointed at `zero16`, a
// 16-byte chunk of zeros (ldmatrix takes one address per 8 x 8 matrix row), so padded keys contribute exactly 0 to p V and
// no staging row has to be cleared; their scores are masked by select.
template <int NT>
__device__ __forceinline__ void mv_block(const uint32_t (&qa)[8], uint32_t kbuf, uint32_t vbuf, int t0, int nvalid,
                                         uint32_t zero16, float scale_log2, int lane, MvState& st) {
  const int t4 = lane & 3, lrow = lane & 7, lmat = lane >> 3;
  const int nt = ((nvalid + 15) & ~15) >> 3;  // 8-key tiles (even)
  float sc[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    if (j < nt) {
#pragma unroll
      for (int i = 0; i < 4; ++i) sc[j][i] = 0.f;
      const int r = 8 * j + lrow;
      const uint32_t row = kbuf + r * 128;
      const bool real = r < nvalid;
      uint32_t b[4];
      ldmatrix_x4(b, real ? row + (kv_chunk(t0 + r, lmat) << 4) : zero16);  // dims 0 .. 31
      {
        const uint32_t a0[4] = {qa[0], 0u, qa[1], 0u}, a1[4] = {qa[2], 0u, qa[3], 0u};
        mma_16816(sc[j], a0, b[0], b[1]);
        mma_16816(sc[j], a1, b[2], b[3]);
      }
      ldmatrix_x4(b, real ? row + (kv_chunk(t0 + r, lmat + 4) << 4) : zero16);  // dims 32 .. 63
      {
        const uint32_t a2[4] = {qa[4], 0u, qa[5], 0u}, a3[4] = {qa[6], 0u, qa[7], 0u};
        mma_16816(sc[j], a2, b[0], b[1]);
        mma_16816(sc[j], a3, b[2], b[3]);
      }
    }
  }
  // lane t4 of the first quad holds the scores of keys 8 j + 2 t4 (+1); the other quads hold zero rows
  float bm = -INFINITY;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    if (j < nt) {
      const int k0 = 8 * j + 2 * t4;
      sc[j][0] = k0 < nvalid ? sc[j][0] : -INFINITY;
      sc[j][1] = k0 + 1 < nvalid ? sc[j][1] : -INFINITY;
      bm = fmaxf(bm, fmaxf(sc[j][0], sc[j][1]));
    }
  }
  bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 1));
  bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 2));
  const float nm = fmaxf(st.mx, bm);  // finite: the block holds at least one valid key
  const float corr = fast_exp2((st.mx - nm) * scale_log2);  // first block: exp2(-inf) = 0 over zero accumulators
  st.mx = nm;
  st.lsum *= corr;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    st.acc[n][0] *= corr;
    st.acc[n][1] *= corr;
  }
  const float nms = nm * scale_log2;
  uint32_t pa[NT];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    if (j < nt) {
      const float p0 = fast_exp2(sc[j][0] * scale_log2 - nms);  // exp2(-inf) = 0 for masked keys
      const float p1 = fast_exp2(sc[j][1] * scale_log2 - nms);
      st.lsum += p0 + p1;
      pa[j] = pack_half2(p0, p1);
    }
  }
#pragma unroll
  for (int kk = 0; kk < NT / 2; ++kk) {
    if (2 * kk < nt) {
      const uint32_t a[4] = {pa[2 * kk], 0u, pa[2 * kk + 1], 0u};
      const int r = 16 * kk + lrow + 8 * (lmat & 1);
      const uint32_t row = vbuf + r * 128;
      const bool real = r < nvalid;
#pragma unroll
      for (int n2 = 0; n2 < 4; ++n2) {
        uint32_t b[4];
        ldmatrix_x4_trans(b, real ? row + (kv_chunk(t0 + r, 2 * n2 + (lmat >> 1)) << 4) : zero16);
        mma_16816(st.acc[2 * n2], a, b[0], b[1]);
        mma_16816(st.acc[2 * n2 + 1], a, b[2], b[3]);
      }
    }
  }
}

__device__ __forceinline__ void mv_store(MvState& st, __half* orow, int lane) {
  st.lsum += __shfl_xor_sync(0xffffffffu, st.lsum, 1);
  st.lsum += __shfl_xor_sync(0xffffffffu, st.lsum, 2);
  if (lane < 4) {
    const float inv = 1.f / st.lsum;
    uint32_t* o32 = reinterpret_cast<uint32_t*>(orow);
#pragma unroll
    for (int n = 0; n < 8; ++n) o32[4 * n + lane] = pack_half2(st.acc[n][0] * inv, st.acc[n][1] * inv);
  }
}

__device__ __forceinline__ void sts_v4(uint32_t addr, const uint4& x) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x.x), "r"(x.y), "r"(x.z), "r"(x.w) : "memory");
}

// Greedy decode: the K / V rows of a (sequence, head) are contiguous in the cache; each warp pulls them with two bulk
// copies issued before griddepcontrol.wait and adds this step's row behind them.
__global__ void __launch_bounds__(DEC_WARPS * 32)
decode_attn_mma_kernel(const __half* __restrict__ qkv, __half* __restrict__ kcache, __half* __restrict__ vcache,
                       __half* __restrict__ o, int nseq, int H, int t_max, int pos, float scale_log2) {
  extern __shared__ __align__(128) uint8_t dsm[];
  __shared__ __align__(8) unsigned long long bars[DEC_WARPS];
  __shared__ __align__(16) uint32_t zeros[4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x * DEC_WARPS + warp;
  const uint32_t bar = smem_u32(&bars[warp]);
  const int T = pos + 1;
  const uint32_t kbuf = smem_u32(dsm) + warp * 2 * T * 128;  // [T] K rows, then [T] V rows
  const uint32_t vbuf = kbuf + T * 128;
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 4) zeros[threadIdx.x] = 0u;
  __syncthreads();
  pdl_launch_dependents();
  const bool live = pair < nseq * H;
  const int seq = pair / H, h = pair % H;
  const long long own = (static_cast<long long>(seq) * H + h) * t_max;
  // cached rows (written by earlier steps / the prefill, never by the predecessor kernel) stream before griddepcontrol.wait
  if (live && lane == 0 && pos > 0) {
    mbar_arrive_expect_tx(bar, 2u * pos * 128u);
    bulk_load(kbuf, kcache + own * 64, pos * 128u, bar);
    bulk_load(vbuf, vcache + own * 64, pos * 128u, bar);
  }
  pdl_wait();
  if (!live) return;
  const int d = H * 64;
  const __half* qrow = qkv + static_cast<long long>(seq) * 3 * d + h * 64;
  uint32_t qa[8];
  mv2_load_q(qrow, lane, qa);
  if (lane < 16) {  // this step's k (lanes 0-7) and v (lanes 8-15): to the cache for later steps and to the staging row
    const int c = lane & 7;
    const uint4 x = *reinterpret_cast<const uint4*>(qrow + (lane < 8 ? d : 2 * d) + c * 8);
    const int pc = kv_chunk(pos, c);
    *reinterpret_cast<uint4*>((lane < 8 ? kcache : vcache) + (own + pos) * 64 + pc * 8) = x;
    sts_v4((lane < 8 ? kbuf : vbuf) + pos * 128 + pc * 16, x);
  }
  __syncwarp();
  if (pos > 0) mbar_wait(bar, 0);
  Mv2State st;
  mv2_init(st);
  const uint32_t zero16 = smem_u32(zeros);
  for (int kb = 0; kb < T; kb += 64)
    mv2_block<4>(qa, kbuf + kb * 128, vbuf + kb * 128, kb, min(T - kb, 64), zero16, scale_log2, lane, st);
  mv2_store(st, o + static_cast<long long>(seq) * d + h * 64, lane);
}

// Beam search on the same path: one CTA per (image, head), beam + 1 warps. Positions 0 .. shared_len-1 of every beam live
// in ONE cache slot (decode.cu beam_init / beam_step): the CTA stages them once (two bulk copies) and the LAST warp runs
// them for all beams at once — the beams' queries are rows 0 .. beam-1 of the A operand, so the shared part costs one
// pass of tensor-core work per image instead of one per beam (the warp-level MMA rate is what bounds this kernel when
// every beam repeats the prefix). The generated positions of a beam sit wherever the ancestry table says: warp b gathers
// beam b's rows (8 lanes per 128-byte row, all loads in flight before the first store) into a private staging area and
// runs them with its query in row 0. The two partial softmaxes of a beam meet in shared memory (maximum, sum and the
// unnormalised output row of the shared part) and warp b merges and stores. Replaces the FMA kernel above (kept behind
// CLIPCAP_B200_DECODE_ATTN_FMA=1).
constexpr int BEAM_GATHER = 5;  // 16-byte loads per lane and operand in flight: 20 rows per pass
template <int MAX_WARPS, int MIN_CTAS>  // <6, 4>: beams up to 5 with four CTAs per SM (80 registers); <9, 2>: up to kMaxBeam
__global__ void __launch_bounds__(MAX_WARPS * 32, MIN_CTAS)
decode_attn_beam_mma_kernel(const __half* __restrict__ qkv, __half* __restrict__ kcache, __half* __restrict__ vcache,
                            const int32_t* __restrict__ anc, __half* __restrict__ o, int beam, int H, int t_max, int pos,
                            int shared_len, float scale_log2) {
  extern __shared__ __align__(128) uint8_t bsm[];
  __shared__ __align__(8) unsigned long long bbar;
  __shared__ __align__(16) uint32_t zeros[4];
  __shared__ float sh_o[kMaxBeam][64];
  __shared__ float sh_m[kMaxBeam], sh_l[kMaxBeam];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int img = blockIdx.x / H, h = blockIdx.x % H;
  const uint32_t bar = smem_u32(&bbar);
  const int T = pos + 1;
  const int ngen = T - shared_len;  // a beam's own positions (the last one is this step's)
  const int d = H * 64;
  const uint32_t kbuf = smem_u32(bsm), vbuf = kbuf + shared_len * 128;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 4) zeros[threadIdx.x] = 0u;
  __syncthreads();
  pdl_launch_dependents();
  // The ancestry table and cache rows are read after the wait: a version that read them early (they are the product of
  // launches far back in the stream) measured no gain and is not provably ordered behind beam_step.
  pdl_wait();
  const uint32_t zero16 = smem_u32(zeros);
  MvState st;
  mv_init(st);
  if (warp == beam) {
    // ---------------------------------------------------------------- shared positions, all beams: row g = beam g
    if (lane == 0) {
      const int32_t slot = anc[static_cast<long long>(img) * beam * t_max];  // slot of position 0 = the shared slot
      const long long src = ((static_cast<long long>(slot) * H + h) * t_max) * 64;
      mbar_arrive_expect_tx(bar, 2u * shared_len * 128u);
      bulk_load(kbuf, kcache + src, shared_len * 128u, bar);
      bulk_load(vbuf, vcache + src, shared_len * 128u, bar);
    }
    const int g = lane >> 2, t4 = lane & 3;
    uint32_t qa[8];
    const uint32_t* q32 = reinterpret_cast<const uint32_t*>(qkv + static_cast<long long>(img * beam + (g < beam ? g : 0)) * 3 * d + h * 64);
#pragma unroll
    for (int i = 0; i < 8; ++i) qa[i] = g < beam ? q32[4 * i + t4] : 0u;
    mbar_wait(bar, 0);
    for (int kb = 0; kb < shared_len; kb += 32)
      mv_block<4>(qa, kbuf + kb * 128, vbuf + kb * 128, kb, min(shared_len - kb, 32), zero16, scale_log2, lane, st);
    st.lsum += __shfl_xor_sync(0xffffffffu, st.lsum, 1);
    st.lsum += __shfl_xor_sync(0xffffffffu, st.lsum, 2);
    if (g < beam) {
      if (t4 == 0) {
        sh_m[g] = st.mx;
        sh_l[g] = st.lsum;
      }
#pragma unroll
      for (int n = 0; n < 8; ++n)
        *reinterpret_cast<float2*>(&sh_o[g][8 * n + 2 * t4]) = make_float2(st.acc[n][0], st.acc[n][1]);
    }
    asm volatile("bar.sync 1, %0;" ::"r"((beam + 1) * 32) : "memory");
    return;
  }
  // ------------------------------------------------------------------ warp b: beam b's own positions (keys in the M dimension)
  const int seq = img * beam + warp;
  const int32_t* arow = anc + static_cast<long long>(seq) * t_max;
  const uint32_t kown = vbuf + shared_len * 128 + warp * 2 * ngen * 128, vown = kown + ngen * 128;
  const int c = lane & 7;
  for (int r0 = 0; r0 < ngen - 1; r0 += 4 * BEAM_GATHER) {  // rows r0 + 4 j + (lane >> 3), copied verbatim (rotation included)
    uint4 kk[BEAM_GATHER], vv[BEAM_GATHER];
#pragma unroll
    for (int j = 0; j < BEAM_GATHER; ++j) {
      const int r = r0 + 4 * j + (lane >> 3);
      if (r < ngen - 1) {
        const int t = shared_len + r;
        const long long src = ((static_cast<long long>(arow[t]) * H + h) * t_max + t) * 64 + c * 8;
        kk[j] = *reinterpret_cast<const uint4*>(kcache + src);
        vv[j] = *reinterpret_cast<const uint4*>(vcache + src);
      }
    }
#pragma unroll
    for (int j = 0; j < BEAM_GATHER; ++j) {
      const int r = r0 + 4 * j + (lane >> 3);
      if (r < ngen - 1) {
        sts_v4(kown + r * 128 + c * 16, kk[j]);
        sts_v4(vown + r * 128 + c * 16, vv[j]);
      }
    }
  }
  const __half* qrow = qkv + static_cast<long long>(seq) * 3 * d + h * 64;
  uint32_t qa[8];
  mv2_load_q(qrow, lane, qa);
  if (lane < 16) {  // this step's k, v: appended to the row's own slot and placed as the last private row
    const uint4 x = *reinterpret_cast<const uint4*>(qrow + (lane < 8 ? d : 2 * d) + c * 8);
    const int pc = kv_chunk(pos, c);
    *reinterpret_cast<uint4*>((lane < 8 ? kcache : vcache) + ((static_cast<long long>(seq) * H + h) * t_max + pos) * 64 + pc * 8) = x;
    sts_v4((lane < 8 ? kown : vown) + (ngen - 1) * 128 + pc * 16, x);
  }
  __syncwarp();
  Mv2State ps;
  mv2_init(ps);
  for (int kb = 0; kb < ngen; kb += 32)
    mv2_block<2>(qa, kown + kb * 128, vown + kb * 128, shared_len + kb, min(ngen - kb, 32), zero16, scale_log2, lane, ps);
  mv2_finish(ps);
  asm volatile("bar.sync 1, %0;" ::"r"((beam + 1) * 32) : "memory");  // the shared part of every beam is in shared memory
  {
    const int g = lane >> 2, t4 = lane & 3;
    const float ms = sh_m[warp], ls = sh_l[warp];
    const float m = fmaxf(ms, ps.mx);
    const float es = fast_exp2((ms - m) * scale_log2), ep = fast_exp2((ps.mx - m) * scale_log2);
    const float inv = 1.f / (ls * es + ps.lsum * ep);
    float lo, hi;
    mv2_mine(ps, lane, ep, lo, hi);
    lo = (sh_o[warp][16 * t4 + g] * es + lo) * inv;
    hi = (sh_o[warp][16 * t4 + 8 + g] * es + hi) * inv;
    mv2_store_pairs(lo, hi, o + static_cast<long long>(seq) * d + h * 64, lane);
  }
}

// Attention of query row 0 (the class token) of every (sequence, head) over all S keys: one warp per (b, h). A lane owns
// keys lane, lane + 32, ...: it scores them against the query held in registers (whole 128-byte K rows, all loads
// independent), the softmax is one warp reduction, then the lane accumulates p_j * V_j over ITS keys into 64 registers
// (again whole rows, all loads in flight) and the 32 partial vectors are folded through shared memory. HBM-bound: K and V
// are read exactly once.
constexpr int CLS_KEYS_PER_LANE = 9;  // 288 keys per pass (ViT-L/14: 257 tokens = one pass)
__global__ void __launch_bounds__(128)
cls_attn_kernel(const __half* __restrict__ qkv, long long sb, long long sw, long long sh, long long st,
                __half* __restrict__ o, long long ldo, int B, int S, int H, int qrow, float scale_log2) {
  __shared__ float fold[4][32][65];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x * 4 + warp;
  pdl_launch_dependents();
  pdl_wait();
  if (pair >= B * H) return;
  const int b = pair / H, h = pair % H;
  const __half* base = qkv + b * sb + h * sh;
  const __half* qp = base + qrow * st;
  const __half* kp = base + sw;
  const __half* vp = base + 2 * sw;
  float acc[64];
#pragma unroll
  for (int e = 0; e < 64; ++e) acc[e] = 0.f;
  float run_max = -INFINITY, run_sum = 0.f;
  for (int t0 = 0; t0 < S; t0 += 32 * CLS_KEYS_PER_LANE) {
    const int n = min(32 * CLS_KEYS_PER_LANE, S - t0);
    float sc[CLS_KEYS_PER_LANE];
    float mx = -INFINITY;
    {
      float qf[64];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float tmp[8];
        unpack8(*reinterpret_cast<const uint4*>(qp + 8 * c), tmp);
#pragma unroll
        for (int e = 0; e < 8; ++e) qf[8 * c + e] = tmp[e];
      }
#pragma unroll
      for (int i = 0; i < CLS_KEYS_PER_LANE; ++i) {
        const int j = lane + 32 * i;
        sc[i] = -INFINITY;
        if (j < n) {
          const __half* kr = kp + (t0 + j) * st;
          float d = 0.f;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float kf[8];
            unpack8(*reinterpret_cast<const uint4*>(kr + 8 * c), kf);
#pragma unroll
            for (int e = 0; e < 8; ++e) d += qf[8 * c + e] * kf[e];
          }
          sc[i] = d;
          mx = fmaxf(mx, d);
        }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    const float new_max = fmaxf(run_max, mx);
    const float corr = exp2f((run_max - new_max) * scale_log2);  // exp2(-inf) = 0 on the first pass
    run_max = new_max;
    float psum = 0.f;
#pragma unroll
    for (int e = 0; e < 64; ++e) acc[e] *= corr;
#pragma unroll
    for (int i = 0; i < CLS_KEYS_PER_LANE; ++i) {
      const int j = lane + 32 * i;
      if (j < n) {
        const float p = exp2f((sc[i] - new_max) * scale_log2);
        psum += p;
        const __half* vr = vp + (t0 + j) * st;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float vf[8];
          unpack8(*reinterpret_cast<const uint4*>(vr + 8 * c), vf);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[8 * c + e] += p * vf[e];
        }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, off);
    run_sum = run_sum * corr + psum;
  }
  // fold the 32 lanes' partial output vectors: lane c sums columns c and c + 32
#pragma unroll
  for (int e = 0; e < 64; ++e) fold[warp][lane][e] = acc[e];
  __syncwarp();
  float o0 = 0.f, o1 = 0.f;
#pragma unroll 8
  for (int r = 0; r < 32; ++r) {
    o0 += fold[warp][r][lane];
    o1 += fold[warp][r][lane + 32];
  }
  const float inv = 1.f / run_sum;
  __half* orow = o + b * ldo + h * 64;
  orow[lane] = __float2half_rn(o0 * inv);
  orow[lane + 32] = __float2half_rn(o1 * inv);
}

// The same single-query attention on the warp-level tensor-core path (mv_block): one CTA of four warps per (sequence,
// head). The CTA stages the S key / value rows in shared memory with the chunk rotation of kv_chunk applied on the way in
// (8 lanes per 128-byte row, coalesced loads, all of a thread's loads in flight before its first store), warp w walks the
// 64-key blocks w, w + 4, ..., and the four partial softmaxes (maximum, sum, unnormalised row) are merged by warp 0.
// ViT-L/14 class row (S = 257, 4096 pairs): 102 us cold (270 MB at 2.6 TB/s, three CTAs per SM by shared memory) against
// 280 us for cls_attn_kernel, whose lanes each walk whole K / V rows.
constexpr int CLSM_BATCH = 8;  // 16-byte loads per thread in flight
__global__ void __launch_bounds__(128)
cls_attn_mma_kernel(const __half* __restrict__ qkv, long long sb, long long sw, long long sh, long long st,
                    __half* __restrict__ o, long long ldo, int S, int H, int qrow, float scale_log2) {
  extern __shared__ __align__(128) uint8_t csm[];
  __shared__ __align__(16) uint32_t zeros[4];
  __shared__ float sh_o[4][64];
  __shared__ float sh_m[4], sh_l[4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const uint32_t kbuf = smem_u32(csm), vbuf = kbuf + S * 128;
  if (threadIdx.x < 4) zeros[threadIdx.x] = 0u;
  pdl_launch_dependents();
  pdl_wait();
  const __half* base = qkv + b * sb + h * sh;
  {  // stage K and V: chunk c of row t -> row t, chunk c ^ (t & 7)
    const int c = threadIdx.x & 7;
    const int n = 2 * S;  // rows of K then rows of V, 16 per pass of the CTA
    for (int r0 = threadIdx.x >> 3; r0 < n; r0 += 16 * CLSM_BATCH) {
      uint4 x[CLSM_BATCH];
#pragma unroll
      for (int j = 0; j < CLSM_BATCH; ++j) {
        const int r = r0 + 16 * j;
        if (r < n) {
          const int t = r < S ? r : r - S;
          x[j] = *reinterpret_cast<const uint4*>(base + (r < S ? sw : 2 * sw) + t * st + c * 8);
        }
      }
#pragma unroll
      for (int j = 0; j < CLSM_BATCH; ++j) {
        const int r = r0 + 16 * j;
        if (r < n) {
          const int t = r < S ? r : r - S;
          sts_v4((r < S ? kbuf : vbuf) + t * 128 + (kv_chunk(t, c) << 4), x[j]);
        }
      }
    }
  }
  uint32_t qa[8];
  mv2_load_q(base + qrow * st, lane, qa);
  __syncthreads();
  Mv2State ms;
  mv2_init(ms);
  const uint32_t zero16 = smem_u32(zeros);
  bool any = false;
  for (int kb = 64 * warp; kb < S; kb += 256) {
    mv2_block<4>(qa, kbuf + kb * 128, vbuf + kb * 128, kb, min(S - kb, 64), zero16, scale_log2, lane, ms);
    any = true;
  }
  mv2_finish(ms);
  {
    float lo, hi;
    mv2_mine(ms, lane, 1.f, lo, hi);
    sh_o[warp][16 * (lane & 3) + (lane >> 2)] = lo;
    sh_o[warp][16 * (lane & 3) + 8 + (lane >> 2)] = hi;
    if (lane == 0) {
      sh_m[warp] = any ? ms.mx : -INFINITY;
      sh_l[warp] = any ? ms.lsum : 0.f;
    }
  }
  __syncthreads();
  if (warp == 0) {  // lane owns dims 2 lane, 2 lane + 1
    const float m = fmaxf(fmaxf(sh_m[0], sh_m[1]), fmaxf(sh_m[2], sh_m[3]));
    float o0 = 0.f, o1 = 0.f, l = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float e = fast_exp2((sh_m[w] - m) * scale_log2);  // a warp without keys: exp2(-inf) = 0
      const float2 v = *reinterpret_cast<const float2*>(&sh_o[w][2 * lane]);
      o0 += e * v.x;
      o1 += e * v.y;
      l += e * sh_l[w];
    }
    const float inv = 1.f / l;
    *reinterpret_cast<uint32_t*>(o + b * ldo + h * 64 + 2 * lane) = pack_half2(o0 * inv, o1 * inv);
  }
}

__global__ void kv_scatter_kernel(const __half* __restrict__ qkv, __half* __restrict__ kcache,
                                  __half* __restrict__ vcache, int nseq, int T, int H, int t_max, int pos0,
                                  int slot_stride) {
  // one thread per 16-byte chunk of one (seq, t, head) row, k and v
  pdl_launch_dependents();
  pdl_wait();
  const long long n = static_cast<long long>(nseq) * T * H * 8;
  const int d = H * 64;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = i & 7;
    long long r = i >> 3;
    const int h = r % H;
    r /= H;
    const int t = r % T;
    const int seq = r / T;
    const __half* src = qkv + (static_cast<long long>(seq) * T + t) * 3 * d + h * 64 + c * 8;
    const long long dst = ((static_cast<long long>(seq) * slot_stride * H + h) * t_max + pos0 + t) * 64 + kv_chunk(pos0 + t, c) * 8;
    *reinterpret_cast<uint4*>(kcache + dst) = *reinterpret_cast<const uint4*>(src + d);
    *reinterpret_cast<uint4*>(vcache + dst) = *reinterpret_cast<const uint4*>(src + 2 * d);
  }
}

}  // namespace

int attention_run(const __half* q, const __half* k, const __half* v, int64_t ld, __half* o, int64_t ldo, int B, int S,
                  int H, int hd, bool causal, float scale, cudaStream_t s) {
  CC_REQUIRE(B > 0 && S > 0 && H > 0, CC_ESHAPE, "attention: bad shape B=%d S=%d H=%d", B, S, H);
  CC_REQUIRE(ld % 8 == 0 && ldo % 2 == 0, CC_EALIGN, "attention: ld must be a multiple of 8 elements");
  CC_REQUIRE(H <= 65535 && B <= 65535, CC_ESHAPE, "attention: grid too large (H=%d B=%d)", H, B);
  {
    static const bool no_tc = [] {
      const char* e = getenv("CLIPCAP_B200_NO_TC_ATTN");
      return e != nullptr && e[0] == '1';
    }();
    if (!no_tc && vit_attention_fits(S, hd, causal, ld, ldo, q, k, v)) return vit_attention_run(q, k, v, ld, o, ldo, B, S, H, scale, s);
    if (!no_tc && small_attention_fits(S, hd, ld, ldo, q, k, v, o))
      return small_attention_run(q, k, v, ld, o, ldo, B, S, H, hd, causal, scale, s);
  }
#define CC_ATTN_CASE(HD)                                                                         \
  if (hd == HD)                                                                                  \
    return causal ? launch_attn<HD, true>(q, k, v, ld, o, ldo, B, S, H, scale, s)                \
                  : launch_attn<HD, false>(q, k, v, ld, o, ldo, B, S, H, scale, s);
  CC_ATTN_CASE(48)
  CC_ATTN_CASE(64)
  CC_ATTN_CASE(96)
  CC_ATTN_CASE(128)
#undef CC_ATTN_CASE
  set_error("attention: head dim %d not supported (48, 64, 96, 128)", hd);
  return CC_ESHAPE;
}

int decode_attention_run(const __half* qkv, __half* kcache, __half* vcache, const int32_t* anc, __half* o, int nseq,
                         int H, int t_max, int pos, float scale, cudaStream_t s, int beam, int shared_len) {
  CC_REQUIRE(pos >= 0 && pos < t_max, CC_ESHAPE, "decode attention: position %d outside cache (t_max %d)", pos, t_max);
  const int pairs = nseq * H;
  const int grid = (pairs + DEC_WARPS - 1) / DEC_WARPS;
  if (anc != nullptr && beam > 1 && beam <= kMaxBeam && nseq % beam == 0 && shared_len > 0 && shared_len <= pos &&
      static_cast<size_t>(shared_len) * 256 <= 64 * 1024) {
    // beams of an image share the cache rows of positions 0 .. shared_len-1: one CTA per (image, head), warp per beam
    static const bool off = [] {
      const char* e = getenv("CLIPCAP_B200_NO_BEAM_ATTN");
      return e != nullptr && e[0] == '1';
    }();
    static const bool fma = [] {
      const char* e = getenv("CLIPCAP_B200_DECODE_ATTN_FMA");
      return e != nullptr && e[0] == '1';
    }();
    const size_t mma_smem = static_cast<size_t>(2 * shared_len + beam * 2 * (pos + 1 - shared_len)) * 128;
    if (!off && !fma && mma_smem <= 96 * 1024) {
      if (beam <= 5) {
        auto kern = decode_attn_beam_mma_kernel<6, 4>;
        CC_OPT_IN_SMEM(kern, 96 * 1024);
        CC_CUDA(launch_pdl(kern, dim3((nseq / beam) * H), dim3((beam + 1) * 32), mma_smem, s, qkv, kcache, vcache, anc, o,
                           beam, H, t_max, pos, shared_len, scale * 1.4426950408889634f));
      } else {
        auto kern = decode_attn_beam_mma_kernel<kMaxBeam + 1, 2>;
        CC_OPT_IN_SMEM(kern, 96 * 1024);
        CC_CUDA(launch_pdl(kern, dim3((nseq / beam) * H), dim3((beam + 1) * 32), mma_smem, s, qkv, kcache, vcache, anc, o,
                           beam, H, t_max, pos, shared_len, scale * 1.4426950408889634f));
      }
      return CC_OK;
    }
    if (!off) {
      CC_OPT_IN_SMEM(decode_attn_beam_kernel, 64 * 1024);
      CC_CUDA(launch_pdl(decode_attn_beam_kernel, dim3((nseq / beam) * H), dim3(beam * 32),
                         static_cast<size_t>(shared_len) * 256, s, qkv, kcache, vcache, anc, o, beam, H, t_max, pos,
                         shared_len, scale * 1.4426950408889634f));
      return CC_OK;
    }
  }
  // greedy (no ancestry indirection): rows of a (sequence, head) are contiguous — bulk copies into shared memory, staging
  // sized to fit beside two other CTAs. Tensor-core matrix-vector kernel by default; CLIPCAP_B200_DECODE_ATTN_FMA=1 keeps
  // the FMA kernel (fp32 probabilities) for comparison.
  if (anc == nullptr) {
    static const bool fma = [] {
      const char* e = getenv("CLIPCAP_B200_DECODE_ATTN_FMA");
      return e != nullptr && e[0] == '1';
    }();
    const size_t mma_smem = static_cast<size_t>(DEC_WARPS) * 2 * (pos + 1) * 128;
    if (!fma && mma_smem <= 72 * 1024) {
      CC_OPT_IN_SMEM(decode_attn_mma_kernel, 72 * 1024);
      CC_CUDA(launch_pdl(decode_attn_mma_kernel, dim3(grid), dim3(DEC_WARPS * 32), mma_smem, s, qkv, kcache, vcache, o,
                         nseq, H, t_max, pos, scale * 1.4426950408889634f));
      return CC_OK;
    }
    const size_t bulk_smem = static_cast<size_t>(DEC_WARPS) * 2 * pos * 128;  // the pos cached rows of K and V per warp
    if (bulk_smem <= 72 * 1024) {
      CC_OPT_IN_SMEM(decode_attn_bulk_kernel, 72 * 1024);
      CC_CUDA(launch_pdl(decode_attn_bulk_kernel, dim3(grid), dim3(DEC_WARPS * 32), bulk_smem, s, qkv, kcache, vcache, o,
                         nseq, H, t_max, pos, scale * 1.4426950408889634f));
      return CC_OK;
    }
  }
  CC_CUDA(launch_pdl(decode_attn_kernel, dim3(grid), dim3(DEC_WARPS * 32), 0, s, qkv, kcache, vcache, anc, o, nseq, H,
                     t_max, pos, scale * 1.4426950408889634f));
  return CC_OK;
}

int cls_attention_run(const __half* qkv, long long sb, long long sw, long long sh, long long st, __half* o, int64_t ldo,
                      int B, int S, int H, float scale, cudaStream_t s, int qrow) {
  CC_REQUIRE(B > 0 && S > 0 && H > 0 && qrow >= 0 && qrow < S, CC_ESHAPE, "cls attention: B=%d S=%d H=%d qrow=%d", B, S, H, qrow);
  CC_REQUIRE(sb % 8 == 0 && sw % 8 == 0 && sh % 8 == 0 && st % 8 == 0 && ldo % 2 == 0, CC_EALIGN,
             "cls attention: strides must keep 16-byte rows");
  const size_t mma_smem = static_cast<size_t>(S) * 256;
  static const bool fma = [] {
    const char* e = getenv("CLIPCAP_B200_DECODE_ATTN_FMA");
    return e != nullptr && e[0] == '1';
  }();
  if (!fma && mma_smem <= 160 * 1024) {
    CC_OPT_IN_SMEM(cls_attn_mma_kernel, 160 * 1024);
    CC_CUDA(launch_pdl(cls_attn_mma_kernel, dim3(B * H), dim3(128), mma_smem, s, qkv, sb, sw, sh, st, o,
                       static_cast<long long>(ldo), S, H, qrow, scale * 1.4426950408889634f));
    return CC_OK;
  }
  const int pairs = B * H;
  CC_CUDA(launch_pdl(cls_attn_kernel, dim3((pairs + 3) / 4), dim3(128), 0, s, qkv, sb, sw, sh, st, o,
                     static_cast<long long>(ldo), B, S, H, qrow, scale * 1.4426950408889634f));
  return CC_OK;
}

int kv_scatter_run(const __half* qkv, __half* kcache, __half* vcache, int nseq, int T, int H, int t_max, int pos0,
                   int slot_stride, cudaStream_t s) {
  CC_REQUIRE(pos0 + T <= t_max, CC_ESHAPE, "kv scatter: %d + %d positions exceed cache length %d", pos0, T, t_max);
  const long long n = static_cast<long long>(nseq) * T * H * 8;
  const int grid = static_cast<int>(std::min<long long>((n + 255) / 256, 148LL * 16));
  CC_CUDA(launch_pdl(kv_scatter_kernel, dim3(grid), dim3(256), 0, s, qkv, kcache, vcache, nseq, T, H, t_max, pos0,
                     slot_stride));
  return CC_OK;
}

}  // namespace cc
