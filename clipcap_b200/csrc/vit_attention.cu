// tcgen05 attention for the CLIP ViT-L/14 tower: softmax(Q K^T / sqrt(64)) V with S = 257 tokens (CLS + 16x16 patches),
// head dim 64, no mask. Replaces the attention inside OpenAI CLIP's ResidualAttentionBlock (nn.MultiheadAttention), the
// arithmetic behind CLIPModel.forward (reference clipcap/encoders/clip.py:112-129).
//
// 257 = 1 + 256, so the problem is cut into a tensor-core part and a rank-one part:
//   * the 256 patch queries x 256 patch keys of one (image, head) run on tcgen05 as two 128-query tiles —
//     S = Q K^T (128 x 256 x 64, both operands K-major in shared memory, accumulator in 256 TMEM columns), the
//     probabilities go back to TMEM as packed fp16 pairs over the columns S occupied and feed the second MMA as its A
//     operand straight from TMEM, O = P V (128 x 64 x 256; V is the MN-major B operand exactly as TMA lands its [key][64]
//     rows, so nothing is transposed), accumulator in the upper, already consumed half of the S columns;
//   * the CLS key contributes one extra score per query (a 64-long dot product on the CUDA cores, folded into the row
//     max / sum) and one rank-one update p0 * v0 added to O in the epilogue;
//   * the CLS query row (one per image and head) is done by a spare warp on the CUDA cores, reading K and V from the
//     shared-memory tiles that are there anyway.
// All 257 keys are visible at once, so the softmax is a plain two-pass one (row max, then exp2 with the scale folded in);
// thread == query row == TMEM lane, no shuffles.
//
// The kernel is persistent (one CTA per SM, 640 threads) and MUFU-bound by design (128 x 257 exp2 per tile):
//   warp 0        TMA producer: Q (2 tiles), K, V of the next (image, head) into the other of two 96 KB stages
//   warps 1, 2    one MMA issuer per query tile ("stream"): S_t, then - once its softmax group has written P_t - O_t
//   warp 3        CLS query row
//   warps 4..11   softmax group of tile 0, warps 12..19 of tile 1: TWO threads per query row (warp = lane quadrant x
//                 column half; the halves exchange row maximum / sum through shared memory), so every scheduler holds
//                 four arithmetic warps instead of two — one thread per row left the kernel bound by dependent-instruction
//                 latency at 50 % issue utilisation
// The two streams only share the stage buffers, so they drift half a period apart and one group's exp2 phase overlaps the
// other's MMAs, TMEM round trips and epilogue stores.
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"

namespace cc {
namespace {

constexpr int VA_THREADS = 640;  // producer, 2 MMA issuers, CLS-row merger, 16 softmax warps (two threads per query row)
constexpr int VA_S = 257;
constexpr int VA_Q_BYTES = 256 * 64 * 2;  // both 128-query tiles
constexpr int VA_K_BYTES = 256 * 64 * 2;
constexpr int VA_V_BYTES = 256 * 64 * 2;
constexpr int VA_C_BYTES = 512;  // CLS token rows of q, k, v (3 x 128 B, padded)
constexpr int VA_STAGE = VA_Q_BYTES + VA_K_BYTES + VA_V_BYTES;
constexpr int VA_PART = 72;       // floats per partial result of the CLS query row: 64 dims, warp max, warp sum (padded)
constexpr int VA_NSW = 16;        // softmax warps: 2 tiles x 4 lane quadrants x 2 column halves
constexpr int VA_XCH = 2 * 2 * 128 * 3 * 4;  // per (tile, half, row): row maximum, row sum, half of the CLS-key score
constexpr int VA_SCRATCH = 2 * VA_NSW * VA_PART * 4 + VA_XCH;  // per stage one partial per softmax warp + the exchange
constexpr int VA_SMEM = 2 * VA_STAGE + 2 * VA_C_BYTES + VA_SCRATCH + 256 + 1024;
constexpr int VA_TMEM_COLS = 512;  // 256 per stream: S [0,256) -> P [0,128) + O [128,192)
constexpr int VA_EMPTY_ARRIVALS = 2 + VA_NSW + 1;  // both MMA streams (commit), the softmax warps, CLS warp

// Where token `tok` of {q,k,v} of (image b, head h) lives in the [rows, ld] fp16 matrix the TMA descriptor covers:
//   row = b * row_b + h * row_h + row_w[which] + tok,   column = col_w[which] + h * col_h
// packed row-major QKV ([B*S, 3d]):  row_b = S, row_h = 0, row_w = {0,0,0},       col_w = {0, d, 2d}, col_h = 64
// head-major QKV ([B*3*H*S, 64]):    row_b = 3HS, row_h = S, row_w = {0, HS, 2HS}, col_w = {0,0,0},   col_h = 0
struct VaArgs {
  const __half* base;
  long long ld;
  __half* o;  // [B*S, ldo] row-major, head h at column h*64
  long long ldo;
  int row_b, row_h, col_h;
  int row_w[3], col_w[3];
  float scale_log2;
};
__device__ __forceinline__ int va_row(const VaArgs& a, int b, int h, int which) {
  return b * a.row_b + h * a.row_h + a.row_w[which];
}
__device__ __forceinline__ int va_col(const VaArgs& a, int h, int which) { return a.col_w[which] + h * a.col_h; }

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __half2* hp = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(hp[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t r;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(r) : "r"(addr));
  return r;
}
// Blackwell packed fp32 pairs (FFMA2 / FADD2) and the three-input maximum (FMNMX3): the softmax threads are bound by
// instruction issue (one thread per query row, ~4.5 instructions per score), so every instruction that handles two scores
// at once counts.
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// byte offset of 16-byte chunk c of row r inside a [rows][64 fp16] tile with the 128-byte TMA swizzle
__device__ __forceinline__ uint32_t sw128_off(int r, int c) { return r * 128 + ((c ^ (r & 7)) << 4); }

__global__ void __launch_bounds__(VA_THREADS, 1)
vit_attn_kernel(const __grid_constant__ CUtensorMap map, const __grid_constant__ CUtensorMap map_o, VaArgs a, int H,
                int n_items) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto sQ = [&](int s) { return base + s * VA_STAGE; };
  auto sK = [&](int s) { return base + s * VA_STAGE + VA_Q_BYTES; };
  auto sV = [&](int s) { return base + s * VA_STAGE + VA_Q_BYTES + VA_K_BYTES; };
  auto sC = [&](int s, int which) { return base + 2 * VA_STAGE + s * VA_C_BYTES + which * 128; };  // CLS q/k/v rows
  const uint32_t sP = base + 2 * VA_STAGE + 2 * VA_C_BYTES;
  const uint32_t bars = sP + VA_SCRATCH;
  auto full_qk = [&](int s) { return bars + 8u * s; };
  auto full_v = [&](int s) { return bars + 16u + 8u * s; };
  auto empty = [&](int s) { return bars + 32u + 8u * s; };
  auto s_full = [&](int t) { return bars + 48u + 8u * t; };
  auto p_full = [&](int t) { return bars + 64u + 8u * t; };
  auto o_full = [&](int t) { return bars + 80u + 8u * t; };
  auto t_free = [&](int t) { return bars + 96u + 8u * t; };
  const uint32_t tmem_slot = bars + 112u;
  auto cls_full = [&](int s) { return bars + 128u + 8u * s; };  // the 8 partials of the CLS query row of stage s are written
  float* sPf = reinterpret_cast<float*>(smem_raw + (sP - smem_u32(smem_raw)));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map);
    tma_prefetch_desc(&map_o);
    for (int i = 0; i < 2; ++i) {
      mbar_init(full_qk(i), 1);
      mbar_init(full_v(i), 1);
      mbar_init(empty(i), VA_EMPTY_ARRIVALS);
      mbar_init(s_full(i), 1);
      mbar_init(p_full(i), 8);
      mbar_init(o_full(i), 1);
      mbar_init(t_free(i), 8);
      mbar_init(cls_full(i), VA_NSW);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, VA_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        const int s = it & 1;
        const int b = w / H, h = w - b * H;
        // first patch row (token 1) of q, k, v
        const int rq = va_row(a, b, h, 0) + 1, rk = va_row(a, b, h, 1) + 1, rv = va_row(a, b, h, 2) + 1;
        const int cq = va_col(a, h, 0), ck = va_col(a, h, 1), cv = va_col(a, h, 2);
        mbar_wait(empty(s), ((it >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(full_qk(s), VA_Q_BYTES + VA_K_BYTES + 256);
        bulk_load(sC(s, 0), a.base + static_cast<long long>(rq - 1) * a.ld + cq, 128, full_qk(s));
        bulk_load(sC(s, 1), a.base + static_cast<long long>(rk - 1) * a.ld + ck, 128, full_qk(s));
        tma_load_2d(&map, full_qk(s), sQ(s), cq, rq);
        tma_load_2d(&map, full_qk(s), sQ(s) + 128 * 128, cq, rq + 128);
        tma_load_2d(&map, full_qk(s), sK(s), ck, rk);
        tma_load_2d(&map, full_qk(s), sK(s) + 128 * 128, ck, rk + 128);
        mbar_arrive_expect_tx(full_v(s), VA_V_BYTES + 128);
        bulk_load(sC(s, 2), a.base + static_cast<long long>(rv - 1) * a.ld + cv, 128, full_v(s));
        tma_load_2d(&map, full_v(s), sV(s), cv, rv);
        tma_load_2d(&map, full_v(s), sV(s) + 128 * 128, cv, rv + 128);
      }
    }
  } else if (warp <= 2) {
    // ------------------------------------------------------------ MMA issuer of query tile t
    if (lane == 0) {
      const int t = warp - 1;
      const uint32_t tm = tmem_base + 256u * t;
      constexpr uint32_t idesc_s = umma_idesc_f16(128, 256);
      constexpr uint32_t idesc_o = umma_idesc_f16(128, 64) | (1u << 16);  // bit 16: B is MN-major
      int it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        const int s = it & 1;
        const uint32_t ph_s = (it >> 1) & 1u, ph = it & 1u;
        // S_t[128 x 256] = Q_t K^T over the 64 head dims (4 MMAs of K = 16)
        mbar_wait(full_qk(s), ph_s);
        mbar_wait(t_free(t), ph ^ 1u);  // the group has read O_t of the previous item out of these columns
        tc_fence_after();
        const uint64_t dq = umma_desc_kmajor_sw128(sQ(s) + t * 128 * 128), dk = umma_desc_kmajor_sw128(sK(s));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(tm, dq + 2u * k, dk + 2u * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(s_full(t));
        // O_t[128 x 64] = P_t V over the 256 keys (16 MMAs of K = 16): A = P from TMEM (8 columns per step), B = V rows as
        // they lie in shared memory (MN-major, 16 keys = 2 KB per step)
        mbar_wait(full_v(s), ph_s);
        mbar_wait(p_full(t), ph);
        tc_fence_after();
        const uint64_t dv = umma_desc_kmajor_sw128(sV(s));
#pragma unroll
        for (int k = 0; k < 16; ++k)  // P of keys 0..127 at columns [0,64), of keys 128..255 at [128,192); O at [192,256)
          umma_f16_ts(tm + 192, tm + (k < 8 ? 8 * k : 128 + 8 * (k - 8)), dv + static_cast<uint64_t>(k) * (2048 >> 4), idesc_o,
                      k != 0 ? 1u : 0u);
        umma_commit(o_full(t));
        umma_commit(empty(s));
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------ CLS query row (one per image and head): combine
    // The row's 256 patch keys are scored by the softmax warps (thread == key, 32 keys per warp, see below): each warp
    // leaves a partial result (64 dims, its local maximum and sum); this warp adds the CLS key itself and merges the eight
    // partials the way a split-key softmax is merged. (One warp doing the whole row — 2.8 k instructions per item — was
    // the slowest role of the CTA and set the pace of the whole kernel.)
    int it = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      const int s = it & 1;
      const uint32_t ph_s = (it >> 1) & 1u;
      const int b = w / H, h = w - b * H;
      const long long row0 = static_cast<long long>(b) * VA_S;  // CLS row of this image in the output
      mbar_wait(full_qk(s), ph_s);
      // q_cls . k_cls: lane handles dims 2*lane, 2*lane + 1
      const uint32_t qu = lds32(sC(s, 0) + 4 * lane), ku = lds32(sC(s, 1) + 4 * lane);
      const float2 qf = __half22float2(*reinterpret_cast<const __half2*>(&qu));
      const float2 kf = __half22float2(*reinterpret_cast<const __half2*>(&ku));
      float sc0 = qf.x * kf.x + qf.y * kf.y;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sc0 += __shfl_xor_sync(0xffffffffu, sc0, o);
      mbar_wait(full_v(s), ph_s);
      const uint32_t v0u = lds32(sC(s, 2) + 4 * lane);
      const float2 v0f = __half22float2(*reinterpret_cast<const __half2*>(&v0u));
      mbar_wait(cls_full(s), ph_s);
      const float* part = sPf + s * VA_NSW * VA_PART;
      float mx = sc0;
#pragma unroll
      for (int g = 0; g < VA_NSW; ++g) mx = fmaxf(mx, part[g * VA_PART + 64]);
      const float pc = fast_exp2((sc0 - mx) * a.scale_log2);
      float o0 = pc * v0f.x, o1 = pc * v0f.y, sum = pc;
#pragma unroll
      for (int g = 0; g < VA_NSW; ++g) {
        const float f = fast_exp2((part[g * VA_PART + 64] - mx) * a.scale_log2);
        const float2 pv = *reinterpret_cast<const float2*>(part + g * VA_PART + 2 * lane);
        o0 += f * pv.x;
        o1 += f * pv.y;
        sum += f * part[g * VA_PART + 65];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty(s));
      const float inv = 1.f / sum;
      *reinterpret_cast<uint32_t*>(a.o + row0 * a.ldo + h * 64 + 2 * lane) = pack_half2(o0 * inv, o1 * inv);
    }
  } else {
    // ------------------------------------------------------------ softmax + epilogue of tile t: two threads per query row
    const int t = (warp - 4) >> 3;
    const int quad = warp & 3;         // TMEM lane quadrant (== warp % 4)
    const int hh = ((warp - 4) >> 2) & 1;  // column half of S (keys 128*hh ..) / dim half of O (32*hh ..)
    const int r = quad * 32 + lane;    // row inside the 128-query tile == TMEM lane
    const uint32_t t_row = tmem_base + 256u * t + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t s_cols = t_row + 128u * hh;        // this thread's 128 columns of S
    const uint32_t p_cols = t_row + 128u * hh;        // its 64 columns of P (fp16 pairs) start where its S columns start
    float* xch = sPf + 2 * VA_NSW * VA_PART;          // [tile][half][row][3]
    float* mine = xch + ((t * 2 + hh) * 128 + r) * 3;
    const float* other = xch + ((t * 2 + (hh ^ 1)) * 128 + r) * 3;
    const uint32_t pair_bar = 1 + t * 4 + quad;       // named barrier of the two warps that share these 32 rows
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory"); };
    // this thread's half (32 dims) of the score of its query row against the CLS key of the item in stage s
    auto cls_score_half = [&](int s) {
      float s0 = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float qf[8], kf[8];
        unpack8(lds128(sQ(s) + t * 128 * 128 + sw128_off(r, 4 * hh + c)), qf);
        unpack8(lds128(sC(s, 1) + 16 * (4 * hh + c)), kf);
#pragma unroll
        for (int i = 0; i < 8; ++i) s0 += qf[i] * kf[i];
      }
      return s0;
    };
    int it = 0;
    float s0h = 0.f;
    if (static_cast<int>(blockIdx.x) < n_items) {
      mbar_wait(full_qk(0), 0);
      s0h = cls_score_half(0);
    }
    bool stagger = (t == 1);  // start the two streams half a period apart so their exp2 phases interleave
    int release_stage = -1;   // stage whose Q buffer holds this warp pair's in-flight output rows (issuer warp only)
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      const int s = it & 1;
      const uint32_t ph_s = (it >> 1) & 1u, ph = it & 1u;
      const int b = w / H, h = w - b * H;
      const long long row0 = static_cast<long long>(b) * VA_S;

      mbar_wait(s_full(t), ph);
      if (stagger) {
        __nanosleep(2000);
        stagger = false;
      }
      tc_fence_after();
      // pass 1: maximum of this thread's 128 scores (four independent FMNMX3 chains, TMEM loads double-buffered)
      float mloc;
      {
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
        uint32_t sa[32], sb[32];
        auto fold = [&](const uint32_t(&v)[32]) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            m0 = fmax3(m0, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
            m1 = fmax3(m1, __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            m2 = fmax3(m2, __uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
            m3 = fmax3(m3, __uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
          }
        };
        tmem_ld_x32(s_cols, sa);
#pragma unroll 1
        for (int c = 0; c < 4; c += 2) {
          tmem_ld_wait();
          tmem_ld_x32(s_cols + 32 * (c + 1), sb);
          fold(sa);
          tmem_ld_wait();
          if (c + 2 < 4) tmem_ld_x32(s_cols + 32 * (c + 2), sa);
          fold(sb);
        }
        mloc = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      }
      mine[0] = mloc;
      mine[2] = s0h;
      pair_sync();  // both halves' maxima and CLS-key partial scores are visible
      const float s0 = hh == 0 ? s0h + other[2] : other[2] + s0h;  // same operand order in both threads: identical bits
      const float mx = fmaxf(fmaxf(mloc, other[0]), s0);
      const float ms = mx * a.scale_log2;
      if (release_stage >= 0) {
        if (lane == 0) {
          bulk_wait_read<0>();
          mbar_arrive(empty(release_stage));  // the output rows staged in that stage have left
        }
        release_stage = -1;
      }
      // pass 2: p = exp2((s - max) * scale * log2 e) for this thread's 128 scores, partial row sum, P -> TMEM as fp16 pairs
      // over S columns this thread has already consumed
      float sloc;
      {
        unsigned long long acc0 = pack_f32x2(0.f, 0.f), acc1 = acc0;
        const unsigned long long sc2 = pack_f32x2(a.scale_log2, a.scale_log2), nms2 = pack_f32x2(-ms, -ms);
        uint32_t sa[32], sb[32];
        auto chunk = [&](const uint32_t(&v)[32], int c) {
          uint32_t pr[16];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float x0, x1, x2, x3;
            unpack_f32x2(ffma2(pack_f32x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), sc2, nms2), x0, x1);
            unpack_f32x2(ffma2(pack_f32x2(__uint_as_float(v[2 * j + 2]), __uint_as_float(v[2 * j + 3])), sc2, nms2), x2, x3);
            const float p0 = fast_exp2(x0), p1 = fast_exp2(x1), p2 = fast_exp2(x2), p3 = fast_exp2(x3);
            acc0 = fadd2(acc0, pack_f32x2(p0, p1));
            acc1 = fadd2(acc1, pack_f32x2(p2, p3));
            pr[j] = pack_half2(p0, p1);
            pr[j + 1] = pack_half2(p2, p3);
          }
          tmem_st_x16(p_cols + 16 * c, pr);
        };
        tmem_ld_x32(s_cols, sa);
#pragma unroll 1
        for (int c = 0; c < 4; c += 2) {
          tmem_ld_wait();
          tmem_ld_x32(s_cols + 32 * (c + 1), sb);
          chunk(sa, c);
          tmem_ld_wait();
          if (c + 2 < 4) tmem_ld_x32(s_cols + 32 * (c + 2), sa);
          chunk(sb, c + 1);
        }
        float a0, a1, a2, a3;
        unpack_f32x2(acc0, a0, a1);
        unpack_f32x2(acc1, a2, a3);
        sloc = (a0 + a1) + (a2 + a3);
      }
      mine[1] = sloc;
      const float pc = fast_exp2(s0 * a.scale_log2 - ms);  // probability of the CLS key
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(t));

      // While the P V MMA runs: this warp's share of the CLS QUERY row — 16 patch keys per warp, two lanes per key
      // (32 dims each). Score against q_cls, softmax statistics local to the warp's keys, and the warp's partial sum of
      // p * V (lane owns two head dims); warp 3 merges the sixteen partials.
      {
        const int key0 = t * 128 + quad * 32 + hh * 16;
        const int key = key0 + (lane >> 1), dh = lane & 1;
        float sq = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float qf[8], kf[8];
          unpack8(lds128(sC(s, 0) + 16 * (4 * dh + c)), qf);
          unpack8(lds128(sK(s) + sw128_off(key, 4 * dh + c)), kf);
#pragma unroll
          for (int i = 0; i < 8; ++i) sq += qf[i] * kf[i];
        }
        sq += __shfl_xor_sync(0xffffffffu, sq, 1);
        float wm = sq;
#pragma unroll
        for (int o = 16; o > 1; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
        const float pq = fast_exp2((sq - wm) * a.scale_log2);
        float ws = dh == 0 ? pq : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ws += __shfl_xor_sync(0xffffffffu, ws, o);
        mbar_wait(full_v(s), ph_s);
        const uint32_t vrow0 = sV(s) + (lane & 3) * 4;
        const int c16 = lane >> 2;
        float po0 = 0.f, po1 = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float pj = __shfl_sync(0xffffffffu, pq, 2 * j);
          const uint32_t u = lds32(vrow0 + sw128_off(key0 + j, c16));
          const float2 vf = __half22float2(*reinterpret_cast<const __half2*>(&u));
          po0 += pj * vf.x;
          po1 += pj * vf.y;
        }
        float* part = sPf + (s * VA_NSW + (warp - 4)) * VA_PART;
        *reinterpret_cast<float2*>(part + 2 * lane) = make_float2(po0, po1);
        if (lane == 0) {
          part[64] = wm;
          part[65] = ws;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(cls_full(s));
      }
      // ... and the next item's CLS-key score (its stage has been loading all along)
      float s0h_next = 0.f;
      if (w + static_cast<int>(gridDim.x) < n_items) {
        mbar_wait(full_qk(s ^ 1), ((it + 1) >> 1) & 1u);
        s0h_next = cls_score_half(s ^ 1);
      }
      pair_sync();  // both halves' row sums are visible (and every read of the exchange slots above is done)
      const float inv = 1.f / ((hh == 0 ? sloc + other[1] : other[1] + sloc) + pc);

      mbar_wait(o_full(t), ph);
      tc_fence_after();
      uint32_t orr[32];  // this thread's 32 head dims of its output row
      tmem_ld_x32(t_row + 192 + 32 * hh, orr);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_free(t));  // the stream's TMEM columns may take the next S
      // The output rows leave through shared memory and one TMA store per warp PAIR (32 rows x 128 B): the staging area is
      // this tile's Q buffer, dead since the S MMA and the CLS-key scores; the stage is released once the bulk store has
      // read it.
      const uint32_t stg = sQ(s) + t * 128 * 128 + quad * 32 * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float vf[8];
        unpack8(lds128(sC(s, 2) + 16 * (4 * hh + c)), vf);  // CLS value row (o_full implies full_v of this stage)
        uint32_t op[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          op[i] = pack_half2((__uint_as_float(orr[8 * c + 2 * i]) + pc * vf[2 * i]) * inv,
                             (__uint_as_float(orr[8 * c + 2 * i + 1]) + pc * vf[2 * i + 1]) * inv);
        st_shared_v4(stg + sw128_off(lane, 4 * hh + c), op[0], op[1], op[2], op[3]);
      }
      fence_proxy_async_smem();
      pair_sync();  // both halves of the 32 staged rows are written (and this warp is done with the stage's Q / K / V / CLS rows)
      if (hh == 0) {
        if (lane == 0) {
          tma_store_2d(&map_o, stg, h * 64, static_cast<int>(row0) + 1 + 128 * t + 32 * quad);
          bulk_commit();
        }
        release_stage = s;  // released (above / after the loop) once the bulk store has read the staging rows
      } else if (lane == 0) {
        mbar_arrive(empty(s));  // this warp no longer touches the stage; the pending store is the issuer warp's to wait for
      }
      s0h = s0h_next;
    }
  }

  if (warp >= 4 && lane == 0) bulk_wait<0>();  // output stores complete before the CTA exits (no stage left to release:
                                               // the producer is done)
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, VA_TMEM_COLS);
  }
}

}  // namespace

namespace {
int va_launch(const __half* base, uint64_t rows, uint64_t cols, int64_t ld, VaArgs a, int B, int H, cudaStream_t s) {
  CC_REQUIRE((reinterpret_cast<uintptr_t>(a.o) & 15) == 0, CC_EALIGN, "vit attention: output not 16-byte aligned");
  CC_REQUIRE(H <= 65535 && B <= 65535, CC_ESHAPE, "vit attention: grid too large (H=%d B=%d)", H, B);
  CC_OPT_IN_SMEM(vit_attn_kernel, VA_SMEM);
  // Engines call this with the same buffer for every layer, so the last descriptor is cached.
  struct Cached {
    const __half* base = nullptr;
    int64_t ld = 0;
    uint64_t rows = 0, cols = 0;
    const __half* o = nullptr;
    int64_t ldo = 0;
    int B = 0, H = 0;
    CUtensorMap map, map_o;
  };
  static thread_local Cached cache;
  if (cache.o != a.o || cache.ldo != a.ldo || cache.B != B || cache.H != H) {
    CC_TRY(tma_map_f16_sw128(&cache.map_o, a.o, static_cast<uint64_t>(B) * VA_S, static_cast<uint64_t>(H) * 64, a.ldo, 32));
    cache.o = a.o;
    cache.ldo = a.ldo;
    cache.B = B;
    cache.H = H;
  }
  if (cache.base != base || cache.ld != ld || cache.rows != rows || cache.cols != cols) {
    CC_TRY(tma_map_f16_sw128(&cache.map, base, rows, cols, ld, 128));
    cache.base = base;
    cache.ld = ld;
    cache.rows = rows;
    cache.cols = cols;
  }
  const int n_items = B * H;
  const int grid = n_items < num_sms() ? n_items : num_sms();
  CC_CUDA(launch_pdl(vit_attn_kernel, dim3(grid), dim3(VA_THREADS), VA_SMEM, s, cache.map, cache.map_o, a, H, n_items));
  return CC_OK;
}
}  // namespace

bool vit_attention_fits(int S, int hd, bool causal, int64_t ld, int64_t ldo, const __half* q, const __half* k,
                        const __half* v) {
  if (S != VA_S || hd != 64 || causal) return false;
  if (ld % 8 != 0 || ldo % 8 != 0) return false;
  const long long kc = k - q, vc = v - q;
  if (kc < 0 || vc < 0 || kc % 8 != 0 || vc % 8 != 0 || kc >= ld || vc >= ld) return false;
  if ((reinterpret_cast<uintptr_t>(q) & 15) != 0) return false;
  return true;
}

int vit_attention_run(const __half* q, const __half* k, const __half* v, int64_t ld, __half* o, int64_t ldo, int B, int S,
                      int H, float scale, cudaStream_t s) {
  CC_REQUIRE(vit_attention_fits(S, 64, false, ld, ldo, q, k, v), CC_ESHAPE, "vit attention: unsupported shape");
  const long long kc = k - q, vc = v - q;
  const long long cols = (kc > vc ? kc : vc) + static_cast<long long>(H) * 64;
  CC_REQUIRE(cols <= ld, CC_ESHAPE, "vit attention: q/k/v column blocks exceed the row stride");
  VaArgs a{};
  a.base = q;
  a.ld = ld;
  a.o = o;
  a.ldo = ldo;
  a.row_b = S;
  a.row_h = 0;
  a.col_h = 64;
  a.col_w[1] = static_cast<int>(kc);
  a.col_w[2] = static_cast<int>(vc);
  a.scale_log2 = scale * 1.4426950408889634f;
  return va_launch(q, static_cast<uint64_t>(B) * S, static_cast<uint64_t>(cols), ld, a, B, H, s);
}

int vit_attention_heads_run(const __half* qkvh, __half* o, int64_t ldo, int B, int S, int H, float scale,
                            cudaStream_t s) {
  CC_REQUIRE(S == VA_S && ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(qkvh) & 15) == 0, CC_ESHAPE,
             "vit attention (head-major): unsupported shape S=%d", S);
  VaArgs a{};
  a.base = qkvh;
  a.ld = 64;
  a.o = o;
  a.ldo = ldo;
  a.row_b = 3 * H * S;
  a.row_h = S;
  a.col_h = 0;
  a.row_w[1] = H * S;
  a.row_w[2] = 2 * H * S;
  a.scale_log2 = scale * 1.4426950408889634f;
  return va_launch(qkvh, static_cast<uint64_t>(B) * 3 * H * S, 64, 64, a, B, H, s);
}

}  // namespace cc
