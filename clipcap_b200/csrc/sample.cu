// Sampling decode step on the device: the token-selection arithmetic of
//   generate_nucleus_sampling (clipcap/inference/nucleus_sampling.py:33-56): softmax(logits / T) -> the top_k (or all)
//     largest probabilities in descending order -> cumulative sum -> keep up to the first position whose cumulative
//     probability reaches top_p -> renormalise -> torch.multinomial;
//   generate_no_beam (clipcap/inference/no_beam.py:37-62) with clipcap/inference/utils.py: repetition penalty on the
//     tokens seen so far (utils.py:34-38) -> / T -> top_k_top_p_filtering (utils.py:5-32: keep the top_k largest logits,
//     then the smallest descending prefix whose cumulative softmax exceeds top_p) -> the "sentence length penalty"
//     (utils.py:40-49; literally: history tokens whose *logit value* equals the stop-token id get scaled) -> softmax ->
//     torch.multinomial.
// Both keep "the k* most probable tokens" with k* read off the descending cumulative distribution, so no sort is needed:
// one CTA per sequence row holds the row's V exponentials in registers (V / 1024 per thread), finds the cut value by
// bisection on the float bit pattern (the kept mass is monotone in the cut), and draws the token by inverse CDF in index
// order with a counter-based Philox stream (seed, row, step) — the same distribution torch.multinomial samples from,
// not the same random numbers (parity is distribution-level; top_k == 1 is deterministic and exact).
#include "common.h"
#include "decode.h"
#include "ptx.cuh"

namespace cc {
namespace {

constexpr int SMP_THREADS = 1024;
constexpr int SMP_MAX_HISTORY = 1024;

// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3"), one 128-bit block per call.
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// Block-wide reduction; the combination order is fixed, so equal inputs give bit-equal results.
template <class Op>
__device__ __forceinline__ float block_reduce(float v, float* red, Op op, float identity) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    float r = threadIdx.x < SMP_THREADS / 32 ? red[threadIdx.x] : identity;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = op(r, __shfl_xor_sync(0xffffffffu, r, o));
    if (threadIdx.x == 0) red[32] = r;
  }
  __syncthreads();
  return red[32];
}

__device__ __forceinline__ uint32_t order_key(float f) {  // order-preserving integer key of a float
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

struct SampleArgs {
  const float* logits;  // [rows, ldl]
  long long ldl;
  int V;
  int mode;  // CC_GEN_NUCLEUS or CC_GEN_SAMPLE
  float inv_temp, top_p;
  int top_k;
  float rep_penalty, len_penalty_scale;  // len_penalty_scale = sentence_length_factor / desired_sentence_length
  int stop_token;
  const int32_t* prefix_hist;  // [n_prefix] text-prefix tokens (part of every row's history), may be null
  int n_prefix;
  int32_t* tokens;  // [rows, entry_len] generated tokens
  int entry_len, step;
  int32_t* stopped;
  int32_t* lengths;
  const unsigned long long* seed;  // device scalar
};

__global__ void __launch_bounds__(SMP_THREADS) sample_kernel(SampleArgs a) {
  extern __shared__ float ev[];  // [V] logits, then exponentials, of this row
  __shared__ float red[33];
  __shared__ int32_t hist[SMP_MAX_HISTORY];
  __shared__ float scan[SMP_THREADS / 32];
  __shared__ int s_pick, s_last;
  const int row = blockIdx.x, tid = threadIdx.x, V = a.V;
  int32_t* trow = a.tokens + static_cast<long long>(row) * a.entry_len;
  if (a.stopped[row]) {
    if (tid == 0) trow[a.step] = 0;
    return;
  }
  auto fmax_op = [](float p, float q) { return fmaxf(p, q); };
  auto add_op = [](float p, float q) { return p + q; };
  // history = text-prefix tokens + this row's generated tokens (no_beam.py:41-44: `tokens`)
  const int n_hist = min(a.n_prefix + a.step, SMP_MAX_HISTORY);
  for (int i = tid; i < n_hist; i += SMP_THREADS) hist[i] = i < a.n_prefix ? a.prefix_hist[i] : trow[i - a.n_prefix];
  if (tid == 0) {
    s_pick = -1;
    s_last = -1;
  }
  const float* x = a.logits + static_cast<long long>(row) * a.ldl;
  for (int c = tid; c < V; c += SMP_THREADS) ev[c] = x[c];
  __syncthreads();
  // repetition penalty once per distinct history token (gather / scatter semantics of utils.py:34-38), then 1 / T
  if (a.mode == CC_GEN_SAMPLE && a.rep_penalty != 1.0f) {
    for (int i = tid; i < n_hist; i += SMP_THREADS) {
      const int t = hist[i];
      bool first = t >= 0 && t < V;
      for (int k = 0; k < i && first; ++k) first = hist[k] != t;
      if (first) {
        const float l = ev[t];
        ev[t] = l < 0.f ? l * a.rep_penalty : l / a.rep_penalty;
      }
    }
    __syncthreads();
  }
  float mx = -INFINITY;
  for (int c = tid; c < V; c += SMP_THREADS) {
    const float l = ev[c] * a.inv_temp;
    ev[c] = l;
    mx = fmaxf(mx, l);
  }
  mx = block_reduce(mx, red, fmax_op, -INFINITY);

  // Z over all tokens (the nucleus variant softmaxes before its topk: nucleus_sampling.py:45)
  float zall = 0.f;
  for (int c = tid; c < V; c += SMP_THREADS) zall += __expf(ev[c] - mx);
  zall = block_reduce(zall, red, add_op, 0.f);

  // top_k cut by value: the largest t with #{l >= t} >= top_k (utils.py:17-20; the support of topk(top_k))
  if (a.top_k > 0 && a.top_k < V) {
    uint32_t lo = 0u, hi = order_key(mx);  // invariant: count(key >= lo) >= top_k
    while (lo < hi) {
      const uint32_t mid = lo + ((hi - lo + 1u) >> 1);
      float cnt = 0.f;
      for (int c = tid; c < V; c += SMP_THREADS) cnt += order_key(ev[c]) >= mid ? 1.f : 0.f;
      cnt = block_reduce(cnt, red, add_op, 0.f);
      if (cnt >= static_cast<float>(a.top_k)) lo = mid;
      else hi = mid - 1u;
    }
    for (int c = tid; c < V; c += SMP_THREADS)
      if (order_key(ev[c]) < lo) ev[c] = -INFINITY;
    __syncthreads();
  }

  // e = exp(l - max) over the surviving tokens, Z = their sum
  float z = 0.f;
  for (int c = tid; c < V; c += SMP_THREADS) {
    const float e = ev[c] == -INFINITY ? 0.f : __expf(ev[c] - mx);
    ev[c] = e;
    z += e;
  }
  z = block_reduce(z, red, add_op, 0.f);

  // top_p cut. With the tokens in descending order and c_i their cumulative probabilities:
  //   nucleus (nucleus_sampling.py:46-49): idx = first i with c_i >= top_p (searchsorted, left; clipped to top_k - 1),
  //                                        kept = {c_i <= c_idx}; probabilities are relative to ALL tokens
  //   no_beam  (utils.py:22-31):           kept = up to and including the first i with c_i > top_p; probabilities are
  //                                        relative to the tokens that survived top_k
  // Both are "the smallest descending prefix whose mass reaches the target", i.e. the largest cut value e* whose tail
  // mass S(e*) = sum{e_i >= e*} still reaches it: bisection on the bit pattern of e (e >= 0, so the bits are monotone).
  const bool nucleus = a.mode == CC_GEN_NUCLEUS;
  const bool use_p = nucleus ? a.top_p < 1.f : (a.top_p > 0.f && a.top_p < 1.f);
  const float target = a.top_p * (nucleus ? zall : z);
  if (use_p && target <= z) {  // target > z: everything that survived top_k is kept (the clip at nucleus_sampling.py:47)
    uint32_t lo = 1u, hi = 0x3f800000u;  // e in (0, 1]; invariant: S(lo) reaches the target (S(min) = Z)
    while (lo < hi) {
      const uint32_t mid = lo + ((hi - lo + 1u) >> 1);
      const float cutv = __uint_as_float(mid);
      float sacc = 0.f;
      for (int c = tid; c < V; c += SMP_THREADS) sacc += ev[c] >= cutv ? ev[c] : 0.f;
      sacc = block_reduce(sacc, red, add_op, 0.f);
      const bool reaches = nucleus ? sacc >= target : sacc > target;
      if (reaches) lo = mid;
      else hi = mid - 1u;
    }
    const float cutv = __uint_as_float(lo);
    for (int c = tid; c < V; c += SMP_THREADS)
      if (ev[c] < cutv) ev[c] = 0.f;
    __syncthreads();
  }

  // "sentence length penalty" (utils.py:40-49), no_beam only: history tokens whose filtered logit *value* equals the
  // stop-token id are scaled by (current_length / desired_length) * factor. Reproduced literally.
  if (!nucleus && n_hist > 0 && a.len_penalty_scale != 0.f) {
    const float pen = static_cast<float>(a.n_prefix + a.step) * a.len_penalty_scale;
    for (int i = tid; i < n_hist; i += SMP_THREADS) {
      const int t = hist[i];
      bool first = t >= 0 && t < V;
      for (int k = 0; k < i && first; ++k) first = hist[k] != t;
      if (first && ev[t] > 0.f) {
        const float l = __logf(ev[t]) + mx;  // the filtered logit this token carries
        if (l == static_cast<float>(a.stop_token)) ev[t] = __expf(l * pen - mx);
      }
    }
    __syncthreads();
  }

  // Inverse-CDF draw over the kept mass. Any fixed enumeration of the tokens samples the same distribution; here it is
  // thread-major (thread t owns tokens t, t + 1024, ...), which needs a single block scan.
  float mine = 0.f;
  int last_kept = -1;
  for (int c = tid; c < V; c += SMP_THREADS) {
    mine += ev[c];
    if (ev[c] > 0.f) last_kept = c;
  }
  float incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((tid & 31) >= o) incl += t;
  }
  if ((tid & 31) == 31) scan[tid >> 5] = incl;
  if (last_kept >= 0) atomicMax(&s_last, last_kept);
  __syncthreads();
  float warp_off = 0.f, total = 0.f;
  for (int w = 0; w < SMP_THREADS / 32; ++w) {
    if (w < (tid >> 5)) warp_off += scan[w];
    total += scan[w];
  }
  const unsigned long long seed = *a.seed;
  const uint4 rnd = philox4x32(make_uint4(static_cast<uint32_t>(row), static_cast<uint32_t>(a.step), 0u, 0u),
                               make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
  const float u = (static_cast<float>(rnd.x >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0, 1)
  const float goal = u * total;
  const float hi_c = warp_off + incl, lo_c = hi_c - mine;
  if (mine > 0.f && goal >= lo_c && goal < hi_c) {
    float run = lo_c;
    int pick = -1;
    for (int c = tid; c < V; c += SMP_THREADS) {
      if (ev[c] > 0.f) {
        pick = c;  // last kept token of this thread if rounding overshoots
        run += ev[c];
        if (goal < run) break;
      }
    }
    s_pick = pick;
  }
  __syncthreads();
  if (tid == 0) {
    const int tok = s_pick >= 0 ? s_pick : (s_last >= 0 ? s_last : 0);  // goal past the kept mass by rounding: last kept
    if (nucleus) {  // the stop token is part of the output (nucleus_sampling.py:60-68)
      trow[a.step] = tok;
      a.lengths[row] = a.step + 1;
      if (tok == a.stop_token) a.stopped[row] = 1;
    } else {  // no_beam.py:67-73: the stop token ends the caption without being appended
      if (tok == a.stop_token) {
        a.stopped[row] = 1;
        trow[a.step] = 0;
      } else {
        trow[a.step] = tok;
        a.lengths[row] = a.step + 1;
      }
    }
  }
}

}  // namespace

int sample_run(const float* logits, int64_t ldl, int V, int mode, float inv_temp, float top_p, int top_k,
               float rep_penalty, float len_penalty_scale, int stop_token, const int32_t* prefix_hist, int n_prefix,
               int32_t* tokens, int entry_len, int step, int32_t* stopped, int32_t* lengths,
               const unsigned long long* seed, int rows, cudaStream_t s) {
  const size_t smem = static_cast<size_t>(V) * sizeof(float);
  CC_REQUIRE(V > 0 && smem <= 220 * 1024, CC_ESHAPE, "sampling: vocabulary %d does not fit in shared memory", V);
  CC_REQUIRE(mode == CC_GEN_NUCLEUS || mode == CC_GEN_SAMPLE, CC_EINVAL, "sampling: mode %d", mode);
  CC_REQUIRE(n_prefix >= 0 && n_prefix + entry_len <= SMP_MAX_HISTORY, CC_ESHAPE,
             "sampling: history of %d + %d tokens exceeds %d", n_prefix, entry_len, SMP_MAX_HISTORY);
  SampleArgs a{logits, static_cast<long long>(ldl), V, mode, inv_temp, top_p, top_k, rep_penalty, len_penalty_scale,
               stop_token, prefix_hist, n_prefix, tokens, entry_len, step, stopped, lengths, seed};
  CC_OPT_IN_SMEM(sample_kernel, 220 * 1024);
  sample_kernel<<<rows, SMP_THREADS, smem, s>>>(a);
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

}  // namespace cc
