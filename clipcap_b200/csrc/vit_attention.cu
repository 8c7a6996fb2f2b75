// tcgen05 attention for the CLIP ViT-L/14 tower: softmax(Q K^T / sqrt(64)) V with S = 257 tokens (CLS + 16x16 patches),
// head dim 64, no mask. Replaces the attention inside OpenAI CLIP's ResidualAttentionBlock (nn.MultiheadAttention), the
// arithmetic behind CLIPModel.forward (reference clipcap/encoders/clip.py:112-129).
//
// 257 = 1 + 256, so the problem is cut into a tensor-core part and a rank-one part:
//   * one CTA per (image, head, 128-query tile): the 256 patch queries x 256 patch keys run on tcgen05 —
//     S = Q K^T (128 x 256 x 64, both operands K-major in shared memory, accumulator in TMEM columns 0..255), the
//     probabilities go back to TMEM as packed fp16 pairs over the columns S occupied (columns 0..127) and feed the second
//     MMA as its A operand straight from TMEM, O = P V (128 x 64 x 256; V is the MN-major B operand exactly as TMA lands
//     its [key][64] rows, so nothing is transposed), accumulator in TMEM columns 128..191;
//   * the CLS key contributes one extra score per query (a 64-long dot product on the CUDA cores, folded into the row
//     max / sum) and one rank-one update p0 * v0 added to O in the epilogue;
//   * the CLS query row (one row per image and head) is done by a spare warp of the tile-0 CTA on the CUDA cores, reading
//     K and V from the shared-memory tiles that are there anyway.
// All 257 keys are visible at once, so the softmax is a plain two-pass one (row max, then exp2 with the scale folded in);
// thread == query row == TMEM lane, no shuffles. Two CTAs fit per SM (80 KB shared memory, 256 TMEM columns each), so one
// CTA's exp2 phase (the MUFU-bound part: 128 x 257 exp2 per tile) overlaps the other's loads and MMAs.
#include "common.h"
#include "ptx.cuh"

namespace cc {
namespace {

constexpr int VA_THREADS = 224;  // warp 0 TMA, warp 1 MMA, warps 2..5 softmax (thread == query row), warp 6 CLS query
constexpr int VA_S = 257;
constexpr int VA_Q_BYTES = 128 * 64 * 2;
constexpr int VA_K_BYTES = 256 * 64 * 2;
constexpr int VA_V_BYTES = 256 * 64 * 2;
constexpr int VA_SCRATCH = 2048;  // CLS-query probabilities (257 floats)
constexpr int VA_SMEM = VA_Q_BYTES + VA_K_BYTES + VA_V_BYTES + VA_SCRATCH + 64 + 1024;
constexpr int VA_TMEM_COLS = 256;

struct VaArgs {
  const __half* q;
  const __half* k;
  const __half* v;
  long long ld;
  __half* o;
  long long ldo;
  int kcol, vcol;  // column offsets of k and v relative to q inside the row
  float scale_log2;
};

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
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t r;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(r) : "r"(addr));
  return r;
}
// byte offset of 16-byte chunk c of row r inside a [rows][64 fp16] tile with the 128-byte TMA swizzle
__device__ __forceinline__ uint32_t sw128_off(int r, int c) { return r * 128 + ((c ^ (r & 7)) << 4); }

__global__ void __launch_bounds__(VA_THREADS, 2)
vit_attn_kernel(const __grid_constant__ CUtensorMap map, VaArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sK = sQ + VA_Q_BYTES, sV = sK + VA_K_BYTES, sP = sV + VA_V_BYTES;
  const uint32_t bars = sP + VA_SCRATCH;
  const uint32_t bar_qk = bars, bar_v = bars + 8, bar_s = bars + 16, bar_p = bars + 24, bar_o = bars + 32;
  const uint32_t tmem_slot = bars + 40;
  float* sPf = reinterpret_cast<float*>(smem_raw + (sP - smem_u32(smem_raw)));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const long long row0 = static_cast<long long>(b) * VA_S;  // CLS row of this image

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map);
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 4);
    mbar_init(bar_o, 1);
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
    if (lane == 0) {
      const int r1 = static_cast<int>(row0) + 1;  // first patch row
      mbar_arrive_expect_tx(bar_qk, VA_Q_BYTES + VA_K_BYTES);
      tma_load_2d(&map, bar_qk, sQ, h * 64, r1 + 128 * t);
      tma_load_2d(&map, bar_qk, sK, a.kcol + h * 64, r1);
      tma_load_2d(&map, bar_qk, sK + 128 * 128, a.kcol + h * 64, r1 + 128);
      mbar_arrive_expect_tx(bar_v, VA_V_BYTES);
      tma_load_2d(&map, bar_v, sV, a.vcol + h * 64, r1);
      tma_load_2d(&map, bar_v, sV + 128 * 128, a.vcol + h * 64, r1 + 128);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // S[128 x 256] = Q K^T over the 64 head dims (4 MMAs of K = 16)
      mbar_wait(bar_qk, 0);
      tc_fence_after();
      constexpr uint32_t idesc_s = umma_idesc_f16(128, 256);
      const uint64_t dq = umma_desc_kmajor_sw128(sQ), dk = umma_desc_kmajor_sw128(sK);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base, dq + 2u * k, dk + 2u * k, idesc_s, k != 0 ? 1u : 0u);
      umma_commit(bar_s);
      // O[128 x 64] = P V over the 256 keys (16 MMAs of K = 16): A = P from TMEM (8 columns per step), B = V rows as they
      // lie in shared memory (MN-major, 16 keys = 2 KB per step)
      mbar_wait(bar_v, 0);
      mbar_wait(bar_p, 0);
      tc_fence_after();
      constexpr uint32_t idesc_o = umma_idesc_f16(128, 64) | (1u << 16);  // bit 16: B is MN-major
      const uint64_t dv = umma_desc_kmajor_sw128(sV);
#pragma unroll
      for (int k = 0; k < 16; ++k)
        umma_f16_ts(tmem_base + 128, tmem_base + 8 * k, dv + static_cast<uint64_t>(k) * (2048 >> 4), idesc_o,
                    k != 0 ? 1u : 0u);
      umma_commit(bar_o);
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------ softmax + epilogue, thread == query row
    const int quad = warp & 3;
    const int r = quad * 32 + lane;  // row inside the 128-query tile == TMEM lane
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const __half* k0 = a.k + row0 * a.ld + h * 64;  // CLS key / value rows of this image and head
    const __half* v0 = a.v + row0 * a.ld + h * 64;

    // score against the CLS key while the S MMA runs
    mbar_wait(bar_qk, 0);
    float s0 = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float qf[8], kf[8];
      unpack8(lds128(sQ + sw128_off(r, c)), qf);
      unpack8(__ldg(reinterpret_cast<const uint4*>(k0) + c), kf);
#pragma unroll
      for (int i = 0; i < 8; ++i) s0 += qf[i] * kf[i];
    }

    mbar_wait(bar_s, 0);
    tc_fence_after();
    float mx = s0;
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      uint32_t sr[32];
      tmem_ld_x32(t_row + 32 * c, sr);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(sr[j]));
    }
    const float ms = mx * a.scale_log2;
    float sum = 0.f;
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      uint32_t sr[32];
      tmem_ld_x32(t_row + 32 * c, sr);
      tmem_ld_wait();
      uint32_t pr[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float p0 = fast_exp2(__uint_as_float(sr[2 * j]) * a.scale_log2 - ms);
        const float p1 = fast_exp2(__uint_as_float(sr[2 * j + 1]) * a.scale_log2 - ms);
        sum += p0 + p1;
        pr[j] = pack_half2(p0, p1);
      }
      tmem_st_x16(t_row + 16 * c, pr);  // P chunk c lands on S columns this thread has already consumed
    }
    const float pc = fast_exp2(s0 * a.scale_log2 - ms);  // probability of the CLS key
    sum += pc;
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_p);

    mbar_wait(bar_o, 0);
    tc_fence_after();
    const float inv = 1.f / sum;
    __half* orow = a.o + (row0 + 1 + 128 * t + r) * a.ldo + h * 64;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t orr[32];
      tmem_ld_x32(t_row + 128 + 32 * half, orr);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float vf[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(v0) + half * 4 + c), vf);
        uint4 out;
        uint32_t* op = reinterpret_cast<uint32_t*>(&out);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          op[i] = pack_half2((__uint_as_float(orr[8 * c + 2 * i]) + pc * vf[2 * i]) * inv,
                             (__uint_as_float(orr[8 * c + 2 * i + 1]) + pc * vf[2 * i + 1]) * inv);
        *reinterpret_cast<uint4*>(orow + half * 32 + c * 8) = out;
      }
    }
  } else if (t == 0) {
    // ------------------------------------------------------------ CLS query row (one per image and head), CUDA cores
    const __half* q0 = a.q + row0 * a.ld + h * 64;
    const __half* k0 = a.k + row0 * a.ld + h * 64;
    const __half* v0 = a.v + row0 * a.ld + h * 64;
    float qf[64];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float tmp[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(q0) + c), tmp);
#pragma unroll
      for (int i = 0; i < 8; ++i) qf[8 * c + i] = tmp[i];
    }
    mbar_wait(bar_qk, 0);
    // scores: lane handles patch keys lane, lane + 32, ... (8 each); lane 0 also the CLS key
    float sc[8];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int kr = lane + 32 * i;
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float kf[8];
        unpack8(lds128(sK + sw128_off(kr, c)), kf);
#pragma unroll
        for (int e = 0; e < 8; ++e) d += qf[8 * c + e] * kf[e];
      }
      sc[i] = d;
      mx = fmaxf(mx, d);
    }
    float sc0 = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float kf[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(k0) + c), kf);
#pragma unroll
      for (int e = 0; e < 8; ++e) sc0 += qf[8 * c + e] * kf[e];
    }
    mx = fmaxf(mx, sc0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float ms = mx * a.scale_log2;
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float p = fast_exp2(sc[i] * a.scale_log2 - ms);
      sum += p;
      sPf[lane + 32 * i] = p;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float pc = fast_exp2(sc0 * a.scale_log2 - ms);
    sum += pc;
    __syncwarp();
    // O: lane owns head dims 2*lane, 2*lane + 1
    mbar_wait(bar_v, 0);
    const float2 v0f = __half22float2(*reinterpret_cast<const __half2*>(v0 + 2 * lane));
    float o0 = pc * v0f.x, o1 = pc * v0f.y;
    const int c16 = lane >> 2;
    const uint32_t within = (lane & 3) * 4;
#pragma unroll 8
    for (int kr = 0; kr < 256; ++kr) {
      const float p = sPf[kr];
      const uint32_t u = lds32(sV + sw128_off(kr, c16) + within);
      const float2 vf = __half22float2(*reinterpret_cast<const __half2*>(&u));
      o0 += p * vf.x;
      o1 += p * vf.y;
    }
    const float inv = 1.f / sum;
    *reinterpret_cast<uint32_t*>(a.o + row0 * a.ldo + h * 64 + 2 * lane) = pack_half2(o0 * inv, o1 * inv);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, VA_TMEM_COLS);
  }
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
  CC_REQUIRE(H <= 65535 && B <= 65535, CC_ESHAPE, "vit attention: grid too large (H=%d B=%d)", H, B);
  static bool configured = false;
  if (!configured) {
    CC_CUDA(cudaFuncSetAttribute(vit_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VA_SMEM));
    configured = true;
  }
  // One descriptor covers q, k and v: they are column blocks of the same [B*S, ld] matrix. Engines call this with the
  // same buffer for every layer, so the last descriptor is cached.
  struct Cached {
    const __half* q = nullptr;
    int64_t ld = 0;
    long long rows = 0, cols = 0;
    CUtensorMap map;
  };
  static thread_local Cached cache;
  const long long rows = static_cast<long long>(B) * S;
  const long long kc = k - q, vc = v - q;
  const long long cols = (kc > vc ? kc : vc) + static_cast<long long>(H) * 64;
  CC_REQUIRE(cols <= ld, CC_ESHAPE, "vit attention: q/k/v column blocks exceed the row stride");
  if (cache.q != q || cache.ld != ld || cache.rows != rows || cache.cols != cols) {
    CC_TRY(tma_map_f16_sw128(&cache.map, q, rows, cols, ld, 128));
    cache.q = q;
    cache.ld = ld;
    cache.rows = rows;
    cache.cols = cols;
  }
  VaArgs a{q, k, v, static_cast<long long>(ld), o, static_cast<long long>(ldo), static_cast<int>(kc),
           static_cast<int>(vc), scale * 1.4426950408889634f};
  CC_CUDA(launch_pdl(vit_attn_kernel, dim3(2, H, B), dim3(VA_THREADS), VA_SMEM, s, cache.map, a));
  return CC_OK;
}

}  // namespace cc
