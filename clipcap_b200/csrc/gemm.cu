// tcgen05 GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T  with a fused epilogue.
//
// Both operands are fp16, K-major ("TN"): activations are row-major [M,K]; weights are stored [N,K] exactly as
// torch.nn.Linear keeps them (GPT-2's Conv1D [K,N] weights are transposed once at pack time). Accumulation is fp32 in
// tensor memory.
//
// Structure (one persistent CTA per SM, 320 threads, warp-specialised):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D boxes {64 x 128} of A and {64 x BN} of W into a ring of
//               128B-swizzled shared-memory stages, completion counted on `full` mbarriers.
//   warp 1      MMA issuer: one thread issues tcgen05.mma.kind::f16 (M=128, N=BN, K=16) x 4 per stage into one of two
//               TMEM accumulator buffers; tcgen05.commit releases the smem stage (`empty`) and, after the last k-block,
//               publishes the accumulator (`tmem_full`).
//   warps 2..9  epilogue: two warps per TMEM lane quadrant, each owning half of the tile's columns. tcgen05.ld
//               (thread == output row) -> bias / activation in registers -> 64-byte-swizzled per-warp staging buffer
//               in shared memory -> one TMA store (fp16 outputs) or one TMA reduce-add (the fp32 residual update
//               h += acc + bias happens in L2; h is never read by the SM) per 32-row x 64-byte chunk. fp32 logits
//               and the fused argmax use direct stores / atomics. `tmem_empty` hands the accumulator back as soon as
//               it has been read, so the epilogue of tile i overlaps the main loop of tile i+1.
// Tiles are walked n-fastest so the A row-panel and the whole W stay L2 resident.
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"

namespace cc {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 320;  // TMA warp + MMA warp + 8 epilogue warps
constexpr int EPI_WARPS = 8;
constexpr int STG_WARP_BYTES = 4096;  // per epilogue warp: two 32-row x 64-byte staging buffers

// CG = 1: one CTA computes a 128 x BN tile. CG = 2 (cta_group::2): a pair of CTAs on neighbouring SMs computes a 256 x BN
// tile; each CTA stages its own 128 rows of A and HALF of the W tile (BN/2 rows), the MMA reads the other half from the
// peer's shared memory, so per SM both the L2->SM operand traffic and the shared-memory reads drop by a third.
template <int BN, int CG = 1>
struct GemmCfg {
  static constexpr int kStageA = BM * BK * 2;
  static constexpr int kStageB = (BN / CG) * BK * 2;
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kStages = (BN >= 256 && CG == 1) ? 4 : 6;
  static constexpr int kTmemCols = 2 * BN;  // two accumulator buffers (power of two >= 64)
  static constexpr int kBarBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kStaging = EPI_WARPS * STG_WARP_BYTES;
  static constexpr int kSplit = BN >= 64 ? 2 : 1;  // column halves handled by different epilogue warps
  static constexpr int kSmem = kStages * kStage + kStaging + kBarBytes + 1024;  // + alignment slack
};

struct GemmArgs {
  int M, N, K;
  const float* bias;
  void* out;
  long long ldc;
  int splits;        // split-K factor (EPI_PARTIAL_F32 only, else 1)
  int kb_per_split;  // k-blocks per split
  int split_rows;    // row offset of split s inside the partial-sum matrix: s * split_rows
  int heads_S, heads_H;  // EPI_F16_HEADS: tokens per image, heads
  int row_base;          // first row of A / out this launch covers (decode row groups); rows row_base .. row_base + M - 1
};

// x * sigmoid(k x) with the two MUFU ops (ex2, rcp) at approximate precision: relative error ~2^-22, far below the
// fp16 rounding of the result. k = 1.702 is CLIP's QuickGELU; gelu_new(x) = x * sigmoid(2u), u = sqrt(2/pi)(x + 0.044715 x^3),
// because 0.5 (1 + tanh u) == sigmoid(2u); tanh(x) = 2 sigmoid(2x) - 1.
__device__ __forceinline__ float fast_sigmoid_l2(float y_log2e) {  // sigmoid(y) given y * log2(e)
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-y_log2e));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return r;
}
__device__ __forceinline__ float act_quickgelu(float x) { return x * fast_sigmoid_l2(x * (1.702f * 1.4426950408889634f)); }
__device__ __forceinline__ float act_gelu_new(float x) {
  const float u2 = (2.f * 0.7978845608028654f * 1.4426950408889634f) * (x + 0.044715f * x * x * x);
  return x * fast_sigmoid_l2(u2);
}
__device__ __forceinline__ float act_tanh(float x) {
  return 2.f * fast_sigmoid_l2(x * (2.f * 1.4426950408889634f)) - 1.f;
}

// Exact GELU 0.5 x (1 + erf(x / sqrt 2)) through erfc(|z|) = t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) exp(-z^2), t = 1 / (1 + p |z|)
// (Abramowitz & Stegun 7.1.26, |error| <= 1.5e-7: three orders below the fp16 rounding of the result) — one rcp, one ex2 and
// six FMAs instead of erff's branchy ~40 instructions, which made the fc1 epilogue the longest kernel of the Swin blocks.
__device__ __forceinline__ float act_gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z * 1.4426950408889634f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float erfc_abs = p * t * e;
  return 0.5f * x * (x >= 0.f ? 2.f - erfc_abs : erfc_abs);
}

template <int EPI>
__device__ __forceinline__ float apply_act(float x) {
  if constexpr (EPI == EPI_F16_RELU) return fmaxf(x, 0.f);
  if constexpr (EPI == EPI_F16_QUICKGELU) return act_quickgelu(x);
  if constexpr (EPI == EPI_F16_GELU_NEW) return act_gelu_new(x);
  if constexpr (EPI == EPI_F16_TANH) return act_tanh(x);
  if constexpr (EPI == EPI_F16_GELU_ERF) return act_gelu_erf(x);
  return x;
}

__device__ __forceinline__ uint32_t float_order_key(float x) {
  const uint32_t b = __float_as_uint(x);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

template <int EPI>
struct EpiTraits {
  static constexpr bool kTma =
      EPI <= EPI_F16_TANH || EPI == EPI_F16_GELU_ERF || EPI == EPI_RESID_F32 || EPI == EPI_PARTIAL_F32 ||
      EPI == EPI_F16_HEADS;
  static constexpr bool kF32 = EPI == EPI_RESID_F32 || EPI == EPI_PARTIAL_F32;
};

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <int BN, int EPI, int CG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_c, GemmArgs args) {
  using Cfg = GemmCfg<BN, CG>;
  constexpr int TM = BM * CG;  // rows of one work item (tile of the CTA / CTA pair)
  const int cta_rank = CG == 2 ? static_cast<int>(cluster_ctarank()) : 0;
  const bool leader = cta_rank == 0;
  const int unit = CG == 2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);       // CTA or pair index
  const int n_units = CG == 2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stg_base = smem_base + Cfg::kStages * Cfg::kStage;  // 1024-byte aligned
  const uint32_t bar_base = stg_base + Cfg::kStaging;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::kStages + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tiles_m = (args.M + TM - 1) / TM;
  const int tiles_n = (args.N + BN - 1) / BN;
  const int num_work = tiles_m * tiles_n * args.splits;  // work item = (tile, k-split), split fastest
  const int num_kb = (args.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if constexpr (EpiTraits<EPI>::kTma) tma_prefetch_desc(&map_c);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(full_bar(s), 1);  // CG == 2: the leader's barrier; its expect_tx covers both CTAs' tiles
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4 * Cfg::kSplit * CG);  // one arrive per active epilogue warp (of both CTAs)
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (CG == 2) {
      tmem_alloc_cg2(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();  // the peer's barriers are initialised before anyone signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // Everything above touched only this CTA's shared / tensor memory. Weights (operand B) are never written by a kernel
  // of the stream, so the producer puts the first stages' weight tiles in flight before waiting for the predecessor
  // grid; activations (operand A) and all outputs are only touched after pdl_wait().
  pdl_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int pre = 0;  // k-blocks of the first work item whose B tile is already in flight
      if (CG == 1 && unit < num_work) {
        const int w = unit;
        const int tile = w / args.splits, split = w - tile * args.splits;
        const int n0 = (tile % tiles_n) * BN;
        const int kb0 = split * args.kb_per_split;
        const int kb1 = min(num_kb, kb0 + args.kb_per_split);
        pre = min(Cfg::kStages, kb1 - kb0);
        for (int i = 0; i < pre; ++i) {
          mbar_arrive_expect_tx(full_bar(i), Cfg::kStage);
          tma_load_2d(&map_b, full_bar(i), smem_base + i * Cfg::kStage + Cfg::kStageA, (kb0 + i) * BK, n0);
        }
      }
      pdl_wait();
      int stage = 0;
      uint32_t phase = 0;
      for (int w = unit; w < num_work; w += n_units) {
        const int tile = w / args.splits, split = w - tile * args.splits;
        const int m0 = args.row_base + (tile / tiles_n) * TM + cta_rank * BM;
        const int n0 = (tile % tiles_n) * BN + cta_rank * (BN / CG);  // CG == 2: this CTA stages its half of the W tile
        const int kb0 = split * args.kb_per_split;
        const int kb1 = min(num_kb, kb0 + args.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          const uint32_t sa = smem_base + stage * Cfg::kStage;
          const uint32_t sb = sa + Cfg::kStageA;
          if constexpr (CG == 2) {
            // Both CTAs' tiles complete (complete_tx) on the LEADER's full barrier, where the MMA is issued; only the
            // leader arrives, with the byte count of both. (A per-stage remote mbarrier.arrive.release.cluster from the
            // peer costs ~1400 cycles and throttles the whole pipeline: measured.)
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t lbar = mapa_cluster(full_bar(stage), 0);
            if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::kStage);
            tma_load_2d_cg2(&map_a, lbar, sa, kb * BK, m0);
            tma_load_2d_cg2(&map_b, lbar, sb, kb * BK, n0);
          } else if (pre > 0) {  // stage is fresh and its B tile + expect_tx were issued above
            --pre;
            tma_load_2d(&map_a, full_bar(stage), sa, kb * BK, m0);
          } else {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_arrive_expect_tx(full_bar(stage), Cfg::kStage);
            tma_load_2d(&map_a, full_bar(stage), sa, kb * BK, m0);
            tma_load_2d(&map_b, full_bar(stage), sb, kb * BK, n0);
          }
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    } else {
      pdl_wait();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (CG == 2: the leader CTA issues for the pair)
    pdl_wait();
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = umma_idesc_f16(TM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = unit; w < num_work; w += n_units, ++it) {
        const int as = it & 1;
        const int split = w % args.splits;
        const int nkb = min(num_kb, (split + 1) * args.kb_per_split) - split * args.kb_per_split;
        mbar_wait(tempty_bar(as), ((it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::kStage;
          const uint32_t sb = sa + Cfg::kStageA;
          const uint64_t da = umma_desc_kmajor_sw128(sa);
          const uint64_t db = umma_desc_kmajor_sw128(sb);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 elements (32 B) along K inside the 128 B swizzle atom: +2 in the (addr >> 4) field
            if constexpr (CG == 2) umma_f16_ss_cg2(d_tmem, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_f16_ss(d_tmem, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if constexpr (CG == 2) umma_commit_cg2_mc(empty_bar(stage), 0x3);  // frees the stage in both CTAs
          else umma_commit(empty_bar(stage));
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if constexpr (CG == 2) umma_commit_cg2_mc(tfull_bar(as), 0x3);  // both CTAs' epilogues read their own TMEM half
        else umma_commit(tfull_bar(as));
      }
    }
  } else if ((warp - 2) < 4 * Cfg::kSplit) {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    pdl_wait();
    const int ew = warp - 2;
    const int quad = warp & 3;        // TMEM lane quadrant this warp may access
    const int half_id = ew >> 2;      // which column half of the tile
    constexpr int WCOLS = BN / Cfg::kSplit;  // accumulator columns this warp owns per tile
    const int col_base = half_id * WCOLS;
    const int row_in_tile = quad * 32 + lane;
    int it = 0;

    if constexpr (EpiTraits<EPI>::kTma) {
      // 64 bytes of output per row and chunk: 32 fp16 or 16 fp32 columns.
      constexpr bool kF32 = EpiTraits<EPI>::kF32;
      constexpr int CH = kF32 ? 16 : 32;
      constexpr int NCH = WCOLS / CH;
      const uint32_t stg = stg_base + ew * STG_WARP_BYTES;
      // 64-byte swizzle: 16-byte chunk index ^= (row >> 1) & 3
      const uint32_t row_off = lane * 64;
      const uint32_t sw = (lane >> 1) & 3;
      uint32_t chunk_no = 0;
      for (int w = unit; w < num_work; w += n_units, ++it) {
        const int as = it & 1;
        const int tile = w / args.splits, split = w - tile * args.splits;
        const int m0 = args.row_base + (tile / tiles_n) * TM + cta_rank * BM;
        const int n0 = (tile % tiles_n) * BN;
        const bool rows_live = (m0 + quad * 32) < args.row_base + args.M;  // warp-uniform
        const int out_row = split * args.split_rows + m0 + quad * 32;
        mbar_wait(tfull_bar(as), (it >> 1) & 1u);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN + col_base;

        uint32_t r[2][CH];
        if constexpr (CH == 32) tmem_ld_x32(t_row, r[0]);
        else tmem_ld_x16(t_row, r[0]);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          tmem_ld_wait();
          if (c + 1 < NCH) {
            if constexpr (CH == 32) tmem_ld_x32(t_row + (c + 1) * CH, r[(c + 1) & 1]);
            else tmem_ld_x16(t_row + (c + 1) * CH, r[(c + 1) & 1]);
          } else {
            // accumulator fully read (last load has landed): hand the TMEM buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (CG == 2) mbar_arrive_remote(mapa_cluster(tempty_bar(as), 0));
              else mbar_arrive(tempty_bar(as));
            }
          }
          const int nc = n0 + col_base + c * CH;
          if (nc >= args.N || !rows_live) continue;  // warp-uniform
          const uint32_t(&rc)[CH] = r[c & 1];
          float v[CH];
          if (args.bias != nullptr) {
            if (nc + CH <= args.N) {
#pragma unroll
              for (int j = 0; j < CH; j += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(args.bias + nc + j));
                v[j] = __uint_as_float(rc[j]) + b4.x;
                v[j + 1] = __uint_as_float(rc[j + 1]) + b4.y;
                v[j + 2] = __uint_as_float(rc[j + 2]) + b4.z;
                v[j + 3] = __uint_as_float(rc[j + 3]) + b4.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < CH; ++j)
                v[j] = __uint_as_float(rc[j]) + (nc + j < args.N ? __ldg(args.bias + nc + j) : 0.f);
            }
          } else {
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(rc[j]);
          }
          // staging buffer (chunk_no & 1) is free once the bulk store issued two chunks ago has read it
          const uint32_t buf = stg + (chunk_no & 1u) * 2048u + row_off;
          ++chunk_no;
          if (lane == 0) bulk_wait_read<1>();
          __syncwarp();
          if constexpr (kF32) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              st_shared_v4(buf + ((q ^ sw) << 4), __float_as_uint(v[4 * q]), __float_as_uint(v[4 * q + 1]),
                           __float_as_uint(v[4 * q + 2]), __float_as_uint(v[4 * q + 3]));
          } else {
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = pack_half2(apply_act<EPI>(v[2 * j]), apply_act<EPI>(v[2 * j + 1]));
#pragma unroll
            for (int q = 0; q < 4; ++q)
              st_shared_v4(buf + ((q ^ sw) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
            if constexpr (EPI == EPI_F16_HEADS) {
              // Rows of this 32-token group that run past the end of the image belong to the next image's plane: the
              // TMA store below clips them, and their lanes write their 64 bytes directly (1 group in 8 straddles).
              const int img = out_row / args.heads_S, tok = out_row - img * args.heads_S + lane;
              if (tok >= args.heads_S && out_row + lane < args.row_base + args.M) {
                const int d_model = args.heads_H * 64;
                const int which = nc / d_model, rem = nc - which * d_model;
                const long long plane = (static_cast<long long>(img + 1) * 3 + which) * args.heads_H + (rem >> 6);
                uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(args.out) +
                                                      (plane * args.heads_S + (tok - args.heads_S)) * 64 + (rem & 63));
#pragma unroll
                for (int q = 0; q < 4; ++q) dst[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
              }
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if constexpr (EPI == EPI_RESID_F32) {
              tma_reduce_add_2d(&map_c, buf - row_off, nc, out_row);
            } else if constexpr (EPI == EPI_F16_HEADS) {
              // 32 rows x 32 columns = half a head of {q,k,v} for 32 consecutive tokens; rows past the end of the image
              // are clipped by TMA (their lanes stored them directly above).
              const int d_model = args.heads_H * 64;
              const int which = nc / d_model, rem = nc - which * d_model;
              const int hh = rem >> 6, dim0 = rem & 63;
              const int img = out_row / args.heads_S, tok0 = out_row - img * args.heads_S;
              const int plane = (img * 3 + which) * args.heads_H + hh;
              tma_store_3d(&map_c, buf - row_off, dim0, tok0, plane);
            } else {
              tma_store_2d(&map_c, buf - row_off, nc, out_row);
            }
            bulk_commit();
          }
        }
      }
      if (lane == 0) bulk_wait<0>();  // all global writes of this warp are complete before the CTA exits
    } else {
      // fp32 logits (arbitrary ldc) and fused argmax: direct global stores / atomics, thread == output row
      constexpr int CH = WCOLS < 32 ? WCOLS : 32;
      for (int tile = unit; tile < num_work; tile += n_units, ++it) {  // splits == 1 here
        const int as = it & 1;
        const int m0 = args.row_base + (tile / tiles_n) * TM + cta_rank * BM;
        const int n0 = (tile % tiles_n) * BN;
        const int m = m0 + row_in_tile;
        const bool row_ok = m < args.row_base + args.M;
        mbar_wait(tfull_bar(as), (it >> 1) & 1u);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN + col_base;

        float best = -INFINITY;
        int best_n = 0x7fffffff;

#pragma unroll 1
        for (int c = 0; c < WCOLS / CH; ++c) {
          const int nc = n0 + col_base + c * CH;
          if (nc >= args.N) break;  // warp-uniform
          uint32_t r[CH];
          if constexpr (CH == 32) tmem_ld_x32(t_row + c * CH, r);
          else tmem_ld_x16(t_row + c * CH, r);
          tmem_ld_wait();
          const bool full_chunk = nc + CH <= args.N;

          if constexpr (EPI == EPI_ARGMAX) {
            if (row_ok) {
#pragma unroll
              for (int j = 0; j < CH; ++j) {
                const float v = __uint_as_float(r[j]);
                if (nc + j < args.N && v > best) {  // strict > keeps the lowest index on ties
                  best = v;
                  best_n = nc + j;
                }
              }
            }
          } else {
            static_assert(EPI == EPI_F32 || EPI == EPI_ARGMAX, "direct epilogue handles fp32 stores and argmax only");
            if (row_ok) {
              float* dst = reinterpret_cast<float*>(args.out) + static_cast<long long>(m) * args.ldc + nc;
              if (full_chunk && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                for (int j = 0; j < CH; j += 4) {
                  float4 o = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                         __uint_as_float(r[j + 3]));
                  if (args.bias != nullptr) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(args.bias + nc + j));
                    o.x += b4.x;
                    o.y += b4.y;
                    o.z += b4.z;
                    o.w += b4.w;
                  }
                  *reinterpret_cast<float4*>(dst + j) = o;
                }
              } else {
#pragma unroll
                for (int j = 0; j < CH; ++j)
                  if (nc + j < args.N)
                    dst[j] = __uint_as_float(r[j]) + (args.bias != nullptr ? __ldg(args.bias + nc + j) : 0.f);
              }
            }
          }
        }

        // accumulator fully read: hand the TMEM buffer back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2) mbar_arrive_remote(mapa_cluster(tempty_bar(as), 0));
          else mbar_arrive(tempty_bar(as));
        }

        if constexpr (EPI == EPI_ARGMAX) {
          if (row_ok && best_n != 0x7fffffff) {
            const unsigned long long key = (static_cast<unsigned long long>(float_order_key(best)) << 32) |
                                           static_cast<unsigned long long>(~static_cast<uint32_t>(best_n));
            atomicMax(reinterpret_cast<unsigned long long*>(args.out) + m, key);
          }
        }
      }
    }
  }

  if ((warp - 2) >= 4 * Cfg::kSplit) pdl_wait();  // idle epilogue warps (narrow tiles)
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();  // neither CTA leaves (or frees TMEM) while the pair's MMAs can still touch it
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_cg2(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

}  // namespace

// 2-D fp16 row-major [rows, cols] with row stride ld (elements); box = {64 cols, box_rows}, 128B swizzle.
int tma_map_f16_sw128(CUtensorMap* m, const __half* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  CC_REQUIRE(fn != nullptr, CC_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  CC_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, CC_EALIGN, "GEMM operand base %p not 16-byte aligned",
             (const void*)base);
  CC_REQUIRE((ld * 2) % 16 == 0, CC_EALIGN, "GEMM operand row stride %llu elements is not a multiple of 8",
             (unsigned long long)ld);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CC_REQUIRE(r == CUDA_SUCCESS, CC_ECUDA, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu box=%u",
             (int)r, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows);
  return CC_OK;
}

namespace {
int encode_map(CUtensorMap* m, const __half* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  return tma_map_f16_sw128(m, base, rows, cols, ld, box_rows);
}

// Output map for the TMA epilogue: [rows, cols] of fp16 / fp32 with row stride ldc, box = {64 bytes of columns, 32
// rows}, 64-byte swizzle (matches the per-warp staging layout in the kernel).
int encode_out_map(CUtensorMap* m, void* base, bool f32, uint64_t rows, uint64_t cols, uint64_t ldc) {
  EncodeTiledFn fn = get_encode_fn();
  CC_REQUIRE(fn != nullptr, CC_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  const uint64_t es = f32 ? 4 : 2;
  CC_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, CC_EALIGN, "GEMM output base %p not 16-byte aligned", base);
  CC_REQUIRE((ldc * es) % 16 == 0, CC_EALIGN, "GEMM output row stride %llu elements is not 16-byte aligned",
             (unsigned long long)ldc);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ldc * es};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(64 / es), 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CC_REQUIRE(r == CUDA_SUCCESS, CC_ECUDA, "cuTensorMapEncodeTiled (output) failed (%d) rows=%llu cols=%llu ldc=%llu",
             (int)r, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ldc);
  return CC_OK;
}

template <int BN, int EPI, int CG>
int launch(const GemmPlan& p, int bn_idx, int M, cudaStream_t s, int row0) {
  using Cfg = GemmCfg<BN, CG>;
  auto kern = gemm_tn_kernel<BN, EPI, CG>;
  CC_OPT_IN_SMEM(kern, Cfg::kSmem);
  const int num_kb = (p.K + BK - 1) / BK;
  const int splits = (EPI == EPI_PARTIAL_F32) ? p.splits : 1;
  const int kb_per = (num_kb + splits - 1) / splits;
  GemmArgs a{M,      p.N,    p.K,          p.bias,    p.out, static_cast<long long>(p.ldc),
             splits, kb_per, p.split_rows, p.heads_S, p.heads_H, row0};
  const int tiles = ((M + BM * CG - 1) / (BM * CG)) * ((p.N + BN - 1) / BN) * splits;
  int units = num_sms() / CG;  // persistent: one CTA (pair) per SM (pair)
  if (CG == 2) {
    // not every SM pair can host a cluster (GPC shapes): size the persistent grid to what is co-resident (per device)
    static std::atomic<int> cluster_cache[64];
    int dev = 0;
    CC_CUDA(cudaGetDevice(&dev));
    int max_clusters = cluster_cache[dev & 63].load();
    if (max_clusters <= 0) {
      cudaLaunchConfig_t qc = {};
      qc.gridDim = dim3(device_sms());
      qc.blockDim = dim3(GEMM_THREADS);
      qc.dynamicSmemBytes = Cfg::kSmem;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 2;
      qa[0].val.clusterDim.y = 1;
      qa[0].val.clusterDim.z = 1;
      qc.attrs = qa;
      qc.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &qc) != cudaSuccess || n <= 0) n = device_sms() / 2;
      max_clusters = n;
      cluster_cache[dev & 63].store(n);
      if (getenv("CLIPCAP_B200_VERBOSE")) fprintf(stderr, "clipcap_b200: %d co-resident CTA pairs on device %d\n", n, dev);
    }
    if (units > max_clusters) units = max_clusters;
  }
  const int grid = (tiles < units ? tiles : units) * CG;
  CC_CUDA(launch_pdl_cluster(kern, dim3(grid), dim3(GEMM_THREADS), Cfg::kSmem, s, CG, p.map_a, p.map_b[bn_idx], p.map_c, a));
  return CC_OK;
}

template <int BN, int CG>
int launch_epi(const GemmPlan& p, int bn_idx, int M, cudaStream_t s, int row0) {
  switch (p.epi) {
    case EPI_F16_NONE: return launch<BN, EPI_F16_NONE, CG>(p, bn_idx, M, s, row0);
    case EPI_F16_RELU: return launch<BN, EPI_F16_RELU, CG>(p, bn_idx, M, s, row0);
    case EPI_F16_QUICKGELU: return launch<BN, EPI_F16_QUICKGELU, CG>(p, bn_idx, M, s, row0);
    case EPI_F16_GELU_NEW: return launch<BN, EPI_F16_GELU_NEW, CG>(p, bn_idx, M, s, row0);
    case EPI_F16_TANH: return launch<BN, EPI_F16_TANH, CG>(p, bn_idx, M, s, row0);
    case EPI_F32: return launch<BN, EPI_F32, CG>(p, bn_idx, M, s, row0);
    case EPI_RESID_F32: return launch<BN, EPI_RESID_F32, CG>(p, bn_idx, M, s, row0);
    case EPI_ARGMAX: return launch<BN, EPI_ARGMAX, CG>(p, bn_idx, M, s, row0);
    case EPI_PARTIAL_F32: return launch<BN, EPI_PARTIAL_F32, CG>(p, bn_idx, M, s, row0);
    case EPI_F16_HEADS: return launch<BN, EPI_F16_HEADS, CG>(p, bn_idx, M, s, row0);
    case EPI_F16_GELU_ERF: return launch<BN, EPI_F16_GELU_ERF, CG>(p, bn_idx, M, s, row0);
  }
  set_error("unknown GEMM epilogue %d", p.epi);
  return CC_EINVAL;
}

}  // namespace

namespace {
bool two_cta_enabled() {
  static const bool on = [] {
    const char* e = getenv("CLIPCAP_B200_NO_2CTA");
    return !(e != nullptr && e[0] == '1');
  }();
  return on;
}
// Rough cycle estimate of one persistent CTA's share of the problem: per work item the slower of operand ingest
// (~40 B/clk/SM from L2) and the tensor pipe (128 x BN x 16 MMA = BN/8... 8192 FLOP/clk/SM), plus a fixed fill/drain
// cost; split-K adds the partial-sum store and the consumer's re-read.
double gemm_cost(int M, int N, int K, int bn, int splits, int sms) {
  const int tiles = ((M + BM - 1) / BM) * ((N + bn - 1) / bn) * splits;
  const int waves = (tiles + sms - 1) / sms;
  const double ks = static_cast<double>(K) / splits;
  const double mem = (BM + bn) * ks * 2.0 / 40.0;
  const double mma = 2.0 * BM * bn * ks / 8192.0;
  double c = waves * ((mem > mma ? mem : mma) + 2000.0);
  if (splits > 1) c += waves * (BM * bn * 4.0 / 40.0) + splits * (static_cast<double>(M) * N * 4.0 / (sms * 40.0));
  return c;
}
}  // namespace

int gemm_pick_bn(int M, int N, int K) {
  // Large problems (>= one tile per SM at 128x256): the biggest tile. Small ones (decode, M <= a few hundred rows):
  // whatever keeps the work in the fewest, best balanced waves.
  const int sms = num_sms();
  const int tiles_m = (M + BM - 1) / BM;
  if (tiles_m * ((N + 255) / 256) >= sms) return 256;
  static const int cand[4] = {256, 128, 64, 32};
  int best = 32;
  double best_c = 1e30;
  for (int i = 0; i < 4; ++i) {
    const int bn = cand[i];
    if (bn > 32 && bn / 2 >= N) continue;  // do not use a tile mostly outside N
    const double c = gemm_cost(M, N, K, bn, 1, sms);
    if (c < best_c) {
      best_c = c;
      best = bn;
    }
  }
  return best;
}

void gemm_pick_split(int M, int N, int K, int* bn_out, int* splits_out) {
  // The split factor fixes the order in which an output element's partial sums are added, i.e. the result's bits: it is
  // chosen for the WHOLE device whatever SM budget the caller runs under, so a decode step gives bit-identical results
  // inside an SM partition and outside. The tile width only decides how the work is spread and follows the budget.
  const int num_kb = (K + BK - 1) / BK;
  static const int cand[4] = {256, 128, 64, 32};
  double best_c = 1e30;
  *bn_out = 32;
  *splits_out = 1;
  const int dev_sms = device_sms();
  for (int i = 0; i < 4; ++i) {
    const int bn = cand[i];
    if (bn > 32 && bn / 2 >= N) continue;
    for (int sp = 1; sp <= kMaxSplitK && sp * 2 <= num_kb; sp *= 2) {
      const double c = gemm_cost(M, N, K, bn, sp, dev_sms);
      if (c < best_c) {
        best_c = c;
        *bn_out = bn;
        *splits_out = sp;
      }
    }
  }
  const int sms = num_sms();
  if (sms != dev_sms) {
    best_c = 1e30;
    for (int i = 0; i < 4; ++i) {
      const int bn = cand[i];
      if (bn > 32 && bn / 2 >= N) continue;
      const double c = gemm_cost(M, N, K, bn, *splits_out, sms);
      if (c < best_c) {
        best_c = c;
        *bn_out = bn;
      }
    }
  }
}

int gemm_plan(GemmPlan* p, const __half* a, int64_t lda, int max_rows, const __half* w, int N, int K, int epi,
              const float* bias, void* out, int64_t ldc) {
  CC_REQUIRE(max_rows > 0 && N > 0 && K > 0, CC_ESHAPE, "gemm_plan: bad shape M=%d N=%d K=%d", max_rows, N, K);
  CC_REQUIRE(K % 8 == 0, CC_ESHAPE, "gemm_plan: K=%d must be a multiple of 8 (16-byte rows)", K);
  CC_REQUIRE(epi >= 0 && epi < EPI_COUNT, CC_EINVAL, "gemm_plan: bad epilogue %d", epi);
  p->max_rows = max_rows;
  p->N = N;
  p->K = K;
  p->epi = epi;
  p->bias = bias;
  p->out = out;
  p->ldc = ldc;
  CC_TRY(encode_map(&p->map_a, a, max_rows, K, lda, BM));
  static const int bns[4] = {32, 64, 128, 256};
  for (int i = 0; i < 4; ++i) CC_TRY(encode_map(&p->map_b[i], w, N, K, K, bns[i]));
  // fp16 outputs and the fp32 residual update leave through TMA (rows M..max_rows of `out` are scratch: whole 32-row
  // groups are written); fp32 logits and argmax keys use direct stores and get a copy of map_a as a placeholder.
  CC_REQUIRE(epi != EPI_PARTIAL_F32 && epi != EPI_F16_HEADS, CC_EINVAL,
             "gemm_plan: split-K / head-major plans are built with gemm_plan_partial / gemm_plan_heads");
  if (epi <= EPI_F16_TANH || epi == EPI_F16_GELU_ERF || epi == EPI_RESID_F32)
    CC_TRY(encode_out_map(&p->map_c, out, epi == EPI_RESID_F32, max_rows, N, ldc));
  else
    p->map_c = p->map_a;
  return CC_OK;
}

int gemm_plan_heads(GemmPlan* p, const __half* a, int64_t lda, int max_images, int S, int H, const __half* w, int K,
                    const float* bias, __half* out) {
  CC_REQUIRE(max_images > 0 && S > 0 && H > 0 && K > 0 && K % 8 == 0, CC_ESHAPE,
             "gemm_plan_heads: bad shape images=%d S=%d H=%d K=%d", max_images, S, H, K);
  const int N = 3 * H * 64;
  p->max_rows = max_images * S;
  p->N = N;
  p->K = K;
  p->epi = EPI_F16_HEADS;
  p->bias = bias;
  p->out = out;
  p->ldc = 64;
  p->heads_S = S;
  p->heads_H = H;
  CC_TRY(encode_map(&p->map_a, a, p->max_rows, K, lda, BM));
  static const int bns[4] = {32, 64, 128, 256};
  for (int i = 0; i < 4; ++i) CC_TRY(encode_map(&p->map_b[i], w, N, K, K, bns[i]));
  // 3-D output map: {64 dims, S tokens, planes}, box {32, 32, 1}, 64-byte swizzle (the staging layout of the epilogue)
  EncodeTiledFn fn = get_encode_fn();
  CC_REQUIRE(fn != nullptr, CC_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  CC_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, CC_EALIGN, "gemm_plan_heads: output not 16-byte aligned");
  cuuint64_t dims[3] = {64, static_cast<cuuint64_t>(S), static_cast<cuuint64_t>(max_images) * 3 * H};
  cuuint64_t strides[2] = {64 * 2, static_cast<cuuint64_t>(S) * 64 * 2};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(&p->map_c, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, out, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CC_REQUIRE(r == CUDA_SUCCESS, CC_ECUDA, "cuTensorMapEncodeTiled (head-major output) failed (%d)", (int)r);
  return CC_OK;
}

int gemm_plan_partial(GemmPlan* p, const __half* a, int64_t lda, int max_rows, const __half* w, int N, int K,
                      float* partial, int split_rows, int splits, int bn) {
  CC_REQUIRE(max_rows > 0 && N > 0 && K > 0 && K % 8 == 0 && N % 4 == 0, CC_ESHAPE,
             "gemm_plan_partial: bad shape M=%d N=%d K=%d", max_rows, N, K);
  const int num_kb = (K + BK - 1) / BK;
  CC_REQUIRE(splits >= 1 && splits <= num_kb, CC_EINVAL, "gemm_plan_partial: %d splits over %d k-blocks", splits, num_kb);
  CC_REQUIRE(split_rows % 32 == 0 && split_rows >= max_rows, CC_EINVAL,
             "gemm_plan_partial: split_rows %d must be a multiple of 32 and >= %d", split_rows, max_rows);
  // no empty split: splits = ceil(num_kb / kb_per)
  const int kb_per = (num_kb + splits - 1) / splits;
  splits = (num_kb + kb_per - 1) / kb_per;
  p->max_rows = max_rows;
  p->N = N;
  p->K = K;
  p->epi = EPI_PARTIAL_F32;
  p->bias = nullptr;
  p->out = partial;
  p->ldc = N;
  p->splits = splits;
  p->split_rows = split_rows;
  p->force_bn = bn;
  CC_TRY(encode_map(&p->map_a, a, max_rows, K, lda, BM));
  static const int bns[4] = {32, 64, 128, 256};
  for (int i = 0; i < 4; ++i) CC_TRY(encode_map(&p->map_b[i], w, N, K, K, bns[i]));
  CC_TRY(encode_out_map(&p->map_c, partial, true, static_cast<uint64_t>(splits) * split_rows, N, N));
  return CC_OK;
}

// ------------------------------------------------------------------ live per-launch timing (bench.py roofline)
namespace {
struct ProfRec {
  cudaEvent_t e0, e1;
  double flops;
  int bn;
};
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
int gemm_dispatch(const GemmPlan& p, int bn, int M, cudaStream_t s, int row0);
}  // namespace

void gemm_prof_enable(bool on) {
  for (auto& r : g_prof) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_prof.clear();
  g_prof_on = on;
}

// Sums duration / algorithmic FLOPs (2*M*N*K) over the recorded launches of one tile family: bn = 512 the 256 x 256 CTA-pair
// tile, 256 / 128 / 64 / 32 the single-CTA tiles of that width, 0 every launch, -1 the big tiles (256 and 512: the
// dominant kernel of the bench's roofline figure).
void gemm_prof_read_family(int bn, double* ms, double* flops, long long* n) {
  *ms = 0;
  *flops = 0;
  *n = 0;
  cudaDeviceSynchronize();
  for (auto& r : g_prof) {
    if (bn == -1 ? (r.bn != 256 && r.bn != 512) : (bn != 0 && r.bn != bn)) continue;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.e0, r.e1) != cudaSuccess) continue;
    *ms += t;
    *flops += r.flops;
    *n += 1;
  }
}

void gemm_prof_read(double* ms, double* flops, long long* n) { gemm_prof_read_family(-1, ms, flops, n); }

int gemm_run(const GemmPlan& p, int M, cudaStream_t s, int row0) {
  CC_REQUIRE(M > 0 && row0 >= 0 && row0 + M <= p.max_rows, CC_ESHAPE, "gemm_run: rows %d..%d outside plan (max %d)", row0,
             row0 + M, p.max_rows);
  CC_REQUIRE(row0 % 32 == 0, CC_ESHAPE, "gemm_run: row group must start at a multiple of 32 (got %d)", row0);
  int bn = p.force_bn ? p.force_bn : gemm_pick_bn(M, p.N, p.K);
  // Big problems take the 256 x 256 CTA-pair tile (cta_group::2) when there is at least one such tile per SM pair.
  if (p.force_bn == 0 && bn == 256 && p.epi != EPI_PARTIAL_F32 && two_cta_enabled() &&
      ((M + 255) / 256) * ((p.N + 255) / 256) >= num_sms() / 2)
    bn = 512;
  if (g_prof_on) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(s, &cs);
    if (cs == cudaStreamCaptureStatusNone) {
      ProfRec r;
      r.flops = 2.0 * M * p.N * p.K;
      r.bn = bn;
      CC_CUDA(cudaEventCreate(&r.e0));
      CC_CUDA(cudaEventCreate(&r.e1));
      CC_CUDA(cudaEventRecord(r.e0, s));
      const int st = gemm_dispatch(p, bn, M, s, row0);
      CC_CUDA(cudaEventRecord(r.e1, s));
      g_prof.push_back(r);
      return st;
    }
  }
  return gemm_dispatch(p, bn, M, s, row0);
}

namespace {
int gemm_dispatch(const GemmPlan& p, int bn, int M, cudaStream_t s, int row0) {
  switch (bn) {
    case 32: return launch_epi<32, 1>(p, 0, M, s, row0);
    case 64: return launch_epi<64, 1>(p, 1, M, s, row0);
    case 128: return launch_epi<128, 1>(p, 2, M, s, row0);
    case 256: return launch_epi<256, 1>(p, 3, M, s, row0);
    case 512: return launch_epi<256, 2>(p, 2, M, s, row0);  // CTA pair: each CTA stages 128 rows of the 256-row W tile
  }
  set_error("gemm_run: unsupported BLOCK_N %d", bn);
  return CC_EINVAL;
}
}  // namespace

}  // namespace cc
