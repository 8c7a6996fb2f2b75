// Warp-level tensor-core matrix-vector attention (one query row over a block of keys) for the decode / class-row kernels
// of attention.cu. Device code only.
#pragma once
#include "ptx.cuh"

namespace cc {
namespace {

// Cache rows are stored with their eight 16-byte chunks rotated: chunk c (dims 8c .. 8c+7) of position t sits at chunk
// c ^ (t & 7) of the 128-byte row. Eight consecutive rows read at one logical chunk then fall into eight different bank
// groups of shared memory, so decode_attn_mma_kernel can feed ldmatrix straight from a plain bulk copy of the rows.
__device__ __forceinline__ int kv_chunk(int t, int c) { return c ^ (t & 7); }

// ---- the two products with the KEYS in the M dimension of the MMA (what the single-query kernels use) ----
// With the query in row 0 of A (mv_block in attention.cu, still used where several queries share the keys), 15 of 16 MMA
// rows are padding. Transposed, the padding moves to the 8-wide N dimension: half the MMAs, half the accumulator registers
// (80 instead of 126 per thread in the greedy kernel) and fewer instructions around them —
//   s^T[keys x 8] = K[16 keys x 16 dims] q^T        A = K rows through ldmatrix.x4, B = q broadcast to all 8 columns
//   o^T[dims x 8] = V^T[16 dims x 16 keys] p^T      A = V through ldmatrix.x4.trans, B = p broadcast (re-packed by shuffles)
// Every column of an accumulator holds the same number; lane (g, t) = (lane >> 2, lane & 3) reads keys / dims g and g + 8 of
// each 16-row tile from registers [0] and [2]. Measured (ncu, greedy kernel at T = 43, 4096 pairs): 692 warp instructions per
// pair against 1432 in the FMA kernel, warp-MMA pipe 9 % busy — the kernels are bound by bytes in flight, then by issue.
struct Mv2State {
  float acc[4][4];  // tile n: [0] = dim 16 n + g, [2] = dim 16 n + 8 + g (unnormalised)
  float mx, lsum;   // lsum: this lane's keys only until mv2_finish
};

__device__ __forceinline__ void mv2_init(Mv2State& st) {
#pragma unroll
  for (int n = 0; n < 4; ++n)
#pragma unroll
    for (int i = 0; i < 4; ++i) st.acc[n][i] = 0.f;
  st.mx = -INFINITY;
  st.lsum = 0.f;
}

// B fragments of q for the four 16-dim steps: word (lane & 3) of each 16-byte chunk, every lane
__device__ __forceinline__ void mv2_load_q(const __half* qrow, int lane, uint32_t (&qb)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) qb[i] = reinterpret_cast<const uint32_t*>(qrow)[4 * i + (lane & 3)];
}

// One block of up to 16 MT keys (MT = 4: 64, MT = 2: 32) with an online-softmax update; arguments as mv_block.
template <int MT>
__device__ __forceinline__ void mv2_block(const uint32_t (&qb)[8], uint32_t kbuf, uint32_t vbuf, int t0, int nvalid,
                                          uint32_t zero16, float scale_log2, int lane, Mv2State& st) {
  const int g = lane >> 2, t4 = lane & 3, l7 = lane & 7;
  const int mt = (nvalid + 15) >> 4;  // 16-key tiles
  float sc[MT][4];
#pragma unroll
  for (int m = 0; m < MT; ++m) {
#pragma unroll
    for (int i = 0; i < 4; ++i) sc[m][i] = 0.f;
    if (m < mt) {  // warp-uniform; no shuffles inside (tiles beyond mt are masked below and cost two ex2 of -inf)
      const int r = 16 * m + l7 + ((lane >> 3) & 1) * 8;  // matrices: keys 0-7 / 8-15 at chunk 2 ks, then at chunk 2 ks + 1
      const bool real = r < nvalid;
      const uint32_t row = kbuf + r * 128;
      const int rot = (t0 + r) & 7;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t a[4];
        ldmatrix_x4(a, real ? row + (((2 * ks + (lane >> 4)) ^ rot) << 4) : zero16);
        mma_16816(sc[m], a, qb[2 * ks], qb[2 * ks + 1]);
      }
    }
  }
  float bm = -INFINITY;
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    sc[m][0] = 16 * m + g < nvalid ? sc[m][0] : -INFINITY;
    sc[m][2] = 16 * m + 8 + g < nvalid ? sc[m][2] : -INFINITY;
    bm = fmaxf(bm, fmaxf(sc[m][0], sc[m][2]));
  }
  bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 4));
  bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 8));
  bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 16));
  const float nm = fmaxf(st.mx, bm);  // finite: the block holds at least one valid key
  const float corr = fast_exp2((st.mx - nm) * scale_log2);  // first block: exp2(-inf) = 0 over zero accumulators
  st.mx = nm;
  st.lsum *= corr;
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    st.acc[n][0] *= corr;
    st.acc[n][2] *= corr;
  }
  const float nms = nm * scale_log2;
  uint32_t pb[MT][2];
  const int s0 = 8 * t4 + t4, s1 = s0 + 4;  // lanes (g = 2 t4, t4) and (g = 2 t4 + 1, t4)
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    const float plo = fast_exp2(sc[m][0] * scale_log2 - nms);  // exp2(-inf) = 0 for masked keys
    const float phi = fast_exp2(sc[m][2] * scale_log2 - nms);
    st.lsum += plo + phi;
    // B fragment of p for this key step: keys 2 t, 2 t + 1 (from the lanes with g = 2 t, 2 t + 1) and the same + 8
    const float a0 = __shfl_sync(0xffffffffu, plo, s0), a1 = __shfl_sync(0xffffffffu, plo, s1);
    const float b0 = __shfl_sync(0xffffffffu, phi, s0), b1 = __shfl_sync(0xffffffffu, phi, s1);
    pb[m][0] = pack_half2(a0, a1);
    pb[m][1] = pack_half2(b0, b1);
  }
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    if (m < mt) {
      const int r = 16 * m + l7 + (lane >> 4) * 8;  // matrices: keys 0-7 at chunks 2 n, 2 n + 1, then keys 8-15
      const bool real = r < nvalid;
      const uint32_t row = vbuf + r * 128;
      const int rot = (t0 + r) & 7;
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        uint32_t a[4];
        ldmatrix_x4_trans(a, real ? row + (((2 * n + ((lane >> 3) & 1)) ^ rot) << 4) : zero16);
        mma_16816(st.acc[n], a, pb[m][0], pb[m][1]);
      }
    }
  }
}

// row sum over all keys of the warp (the t replicas of a lane hold the same numbers: reduce over g only)
__device__ __forceinline__ void mv2_finish(Mv2State& st) {
  st.lsum += __shfl_xor_sync(0xffffffffu, st.lsum, 4);
  st.lsum += __shfl_xor_sync(0xffffffffu, st.lsum, 8);
  st.lsum += __shfl_xor_sync(0xffffffffu, st.lsum, 16);
}

// this lane's two output dims of tile n == lane & 3: dim 16 n + g ([0]) and 16 n + 8 + g ([2]), scaled
__device__ __forceinline__ void mv2_mine(const Mv2State& st, int lane, float scale, float& lo, float& hi) {
  const int t4 = lane & 3;
  lo = (t4 == 0 ? st.acc[0][0] : t4 == 1 ? st.acc[1][0] : t4 == 2 ? st.acc[2][0] : st.acc[3][0]) * scale;
  hi = (t4 == 0 ? st.acc[0][2] : t4 == 1 ? st.acc[1][2] : t4 == 2 ? st.acc[2][2] : st.acc[3][2]) * scale;
}

// normalised fp16 row -> global: every lane stores one 32-bit pair (even g: dims 16 t + g, + 1; odd g: 16 t + 8 + g - 1, + g)
__device__ __forceinline__ void mv2_store_pairs(float lo, float hi, __half* orow, int lane) {
  const int g = lane >> 2, t4 = lane & 3;
  const float plo = __shfl_xor_sync(0xffffffffu, lo, 4), phi = __shfl_xor_sync(0xffffffffu, hi, 4);  // partner g ^ 1
  uint32_t* o32 = reinterpret_cast<uint32_t*>(orow);
  if ((g & 1) == 0) o32[(16 * t4 + g) >> 1] = pack_half2(lo, plo);
  else o32[(16 * t4 + 8 + g - 1) >> 1] = pack_half2(phi, hi);
}

__device__ __forceinline__ void mv2_store(Mv2State& st, __half* orow, int lane) {
  mv2_finish(st);
  float lo, hi;
  mv2_mine(st, lane, 1.f / st.lsum, lo, hi);
  mv2_store_pairs(lo, hi, orow, lane);
}

}  // namespace
}  // namespace cc
