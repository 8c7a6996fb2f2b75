// Token-selection kernels of the decode loops (clipcap/inference/base.py:80-130), batched over images.
//   greedy  : consumes the fused-argmax keys of the LM-head GEMM (EPI_ARGMAX); generate_beam(beam_size=1) semantics.
//   beam    : per-row log-softmax + top-`beam` (row_topk), per-image merge over beam x beam candidates with the
//             reference's length-normalised scores, stopped-beam rule and source reordering (beam_step), final pick.
// All state lives on the device; nothing here synchronises the host.
#include "common.h"
#include "decode.h"
#include "ptx.cuh"

namespace cc {
namespace {


__global__ void gen_reset_kernel(int32_t* stopped, int32_t* lengths, unsigned long long* keys, float* scores, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    stopped[i] = 0;
    lengths[i] = 0;
    keys[i] = 0ull;
    scores[i] = 0.f;  // greedy / sampling modes return zero scores (never a previous beam call's)
  }
}

// generate_beam with beam_size == 1: tokens are appended until the stop token has been emitted (base.py:117-121);
// the returned length counts the stop token (seq_lengths starts at 1 and grows while not stopped, base.py:70,100).
__global__ void greedy_select_kernel(unsigned long long* keys, int32_t* tokens, int entry_len, int step, int32_t* stopped,
                                     int32_t* lengths, int stop_token, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  if (i >= n) return;
  const int tok = static_cast<int>(argmax_key_index(keys[i]));
  keys[i] = 0ull;  // re-armed for the next step's atomicMax
  if (stopped[i]) {
    tokens[i * entry_len + step] = 0;
    return;
  }
  tokens[i * entry_len + step] = tok;
  lengths[i] = step + 1;
  if (tok == stop_token) stopped[i] = 1;
}

struct Cand {
  float v;
  int i;
};
__device__ __forceinline__ bool better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

// One CTA per sequence row: lp = log(softmax(logits / T)) (base.py:83-84) and the row's `beam` best (value, token).
// Stopped rows contribute the single candidate (0, token 0) (base.py:96-97).
//
// One pass over the row (it is read from HBM exactly once, 16-byte streaming loads). Each warp keeps ONE sorted list of
// its BEAM_MAX best (value, token) pairs, spread over lanes 0 .. BEAM_MAX-1, and every lane carries the list's last entry
// as a threshold: an element is looked at again only if it beats that threshold (one compare + one ballot per four
// elements; the insertion — a shuffle-shift of the list — runs a few dozen times per row once the list has warmed up).
// The first version kept a private sorted list per THREAD: with 32 lanes in lockstep some lane inserted at almost every
// step, so the whole warp walked the 8-deep insertion on nearly every element (229 us per 1280 x 50257 logits against
// ~45 us of HBM time). Online softmax per lane (maximum updated per group of four, one ex2 per element), sums reduced in a
// fixed order.
template <int BEAM_MAX, int TOPK_THREADS>
__global__ void __launch_bounds__(TOPK_THREADS)
row_topk_kernel(const float* __restrict__ logits, long long ldl, int V, float inv_temp, int beam,
                const int32_t* __restrict__ stopped, float* __restrict__ out_val, int32_t* __restrict__ out_idx) {
  constexpr int NW = TOPK_THREADS / 32;
  constexpr unsigned FULL = 0xffffffffu;
  __shared__ float s_red[NW];
  __shared__ float s_cv[NW * BEAM_MAX];
  __shared__ int s_ci[NW * BEAM_MAX];
  __shared__ float s_m, s_sum;
  const int row = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* ov = out_val + static_cast<long long>(row) * beam;
  int32_t* oi = out_idx + static_cast<long long>(row) * beam;
  if (stopped != nullptr && stopped[row]) {
    if (tid < beam) {
      ov[tid] = tid == 0 ? 0.f : -INFINITY;
      oi[tid] = tid == 0 ? 0 : tid;  // distinct dummy tokens; never selected ahead of finite candidates
    }
    return;
  }
  const float* x = logits + static_cast<long long>(row) * ldl;
  float lv = -INFINITY;    // lanes 0 .. BEAM_MAX-1: entry `lane` of the warp's sorted list
  int li = 0x7fffffff;
  float tau_v = -INFINITY;  // the list's last entry, in every lane
  int tau_i = 0x7fffffff;
  float m = -INFINITY, sum = 0.f;
  const bool aligned = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  const int n4 = (V + 3) >> 2;  // groups of four; the last one may be partial
  for (int base = 0; base < n4; base += TOPK_THREADS) {  // warp-uniform trip count (ballots inside)
    const int c4 = base + tid;
    float v[4];
    if (c4 < n4) {
      const int c = 4 * c4;
      if (aligned && c + 3 < V) {
        const float4 q = __ldcs(reinterpret_cast<const float4*>(x) + c4);  // streamed: the logits are dead after this kernel
        v[0] = q.x * inv_temp;
        v[1] = q.y * inv_temp;
        v[2] = q.z * inv_temp;
        v[3] = q.w * inv_temp;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = c + j < V ? x[c + j] * inv_temp : -INFINITY;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = -INFINITY;
    }
    const float m4 = fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3]));
    if (m4 > m) {  // rare after the first groups
      sum *= __expf(m - m4);  // exp(-inf) = 0 on the first group
      m = m4;
    }
    if (m4 > -INFINITY) sum += (__expf(v[0] - m) + __expf(v[1] - m)) + (__expf(v[2] - m) + __expf(v[3] - m));
    unsigned hot = __ballot_sync(FULL, m4 >= tau_v && m4 > -INFINITY);
    while (hot != 0u) {
      const int src = __ffs(hot) - 1;
      hot &= hot - 1;
      const int c0 = 4 * __shfl_sync(FULL, c4, src);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float cv = __shfl_sync(FULL, v[j], src);
        const int ci = c0 + j;
        if (better(cv, ci, tau_v, tau_i)) {  // warp-uniform
          const bool b = better(cv, ci, lv, li);       // entries from here on move one place down
          const float pv = __shfl_up_sync(FULL, lv, 1);
          const int pi = __shfl_up_sync(FULL, li, 1);
          const bool pb = __shfl_up_sync(FULL, b ? 1 : 0, 1) != 0 && lane > 0;
          if (b) {
            lv = pb ? pv : cv;
            li = pb ? pi : ci;
          }
          tau_v = __shfl_sync(FULL, lv, BEAM_MAX - 1);
          tau_i = __shfl_sync(FULL, li, BEAM_MAX - 1);
        }
      }
    }
  }
  // block-wide max, then the sums rescaled to it (fixed order)
  float gm = m;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) gm = fmaxf(gm, __shfl_xor_sync(FULL, gm, o));
  if (lane == 0) s_red[warp] = gm;
  if (lane < BEAM_MAX) {
    s_cv[warp * BEAM_MAX + lane] = lv;
    s_ci[warp * BEAM_MAX + lane] = li;
  }
  __syncthreads();
  if (tid == 0) {
    float mm = s_red[0];
    for (int w = 1; w < NW; ++w) mm = fmaxf(mm, s_red[w]);
    s_m = mm;
  }
  __syncthreads();
  gm = s_m;
  sum = m == -INFINITY ? 0.f : sum * __expf(m - gm);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
  __syncthreads();
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  if (tid == 0) {
    float ss = 0.f;
    for (int w = 0; w < NW; ++w) ss += s_red[w];
    s_sum = ss;
  }
  __syncthreads();
  // `beam` rounds of argmax over the warps' lists (NW * BEAM_MAX <= 64 candidates, two per lane of warp 0)
  if (warp == 0) {
    const float total = s_sum;
    float v0 = lane < NW * BEAM_MAX ? s_cv[lane] : -INFINITY, v1 = lane + 32 < NW * BEAM_MAX ? s_cv[lane + 32] : -INFINITY;
    int i0 = lane < NW * BEAM_MAX ? s_ci[lane] : 0x7fffffff, i1 = lane + 32 < NW * BEAM_MAX ? s_ci[lane + 32] : 0x7fffffff;
    for (int r = 0; r < beam; ++r) {
      const bool first = better(v0, i0, v1, i1);
      float bv = first ? v0 : v1;
      int bi = first ? i0 : i1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float v2 = __shfl_xor_sync(FULL, bv, o);
        const int i2 = __shfl_xor_sync(FULL, bi, o);
        if (better(v2, i2, bv, bi)) {
          bv = v2;
          bi = i2;
        }
      }
      if (lane == 0) {
        ov[r] = logf(expf(bv - gm) / total);  // softmax(-1).log() of the reference, not log_softmax
        oi[r] = bi;
      }
      if (i0 == bi) v0 = -INFINITY, i0 = 0x7fffffff;  // token ids are unique: pop the winner
      if (i1 == bi) v1 = -INFINITY, i1 = 0x7fffffff;
    }
  }
}

// Step 0 (base.py:86-94): the image's single prefill row fans out into `beam` beams.
__global__ void beam_init_kernel(const float* __restrict__ cand_val, const int32_t* __restrict__ cand_idx, BeamState st,
                                 int beam, int entry_len, int t_max, int Tp, int stop_token, int n_img) {
  const int b = blockIdx.x;
  if (b >= n_img) return;
  for (int j = threadIdx.x; j < beam; j += blockDim.x) {
    const int row = b * beam + j;
    // candidates of the prefill row were written at row index b (one logits row per image)
    const float v = cand_val[b * beam + j];
    const int tok = cand_idx[b * beam + j];
    st.scores[row] = v;
    st.seq_len[row] = 1.f;
    st.stopped[row] = tok == stop_token ? 1 : 0;
    st.tokens[0][static_cast<long long>(row) * entry_len] = tok;
  }
  for (int i = threadIdx.x; i < beam * Tp; i += blockDim.x) {
    const int j = i / Tp, t = i % Tp;
    st.anc[0][static_cast<long long>(b * beam + j) * t_max + t] = b * beam;  // prefill K,V live in slot b*beam
  }
}

// Steps >= 1 (base.py:96-119). One CTA per image; thread 0 does the (beam^2)-candidate selection, all threads copy state.
// in/out = ping-pong index of tokens/anc. `pos` = cache position written by this step's decode pass.
template <int BEAM_MAX>
__global__ void beam_step_kernel(const float* __restrict__ cand_val, const int32_t* __restrict__ cand_idx, BeamState st,
                                 int in, int beam, int V, int entry_len, int t_max, int step, int pos, int stop_token,
                                 int n_img) {
  __shared__ int s_src[BEAM_MAX];
  __shared__ int s_frozen;
  const int b = blockIdx.x;
  if (b >= n_img) return;
  const int out = in ^ 1;
  const int base = b * beam;
  if (threadIdx.x == 0) {
    bool all = true;
    for (int j = 0; j < beam; ++j) all = all && st.stopped[base + j] != 0;
    s_frozen = all ? 1 : 0;  // the reference left its loop here (base.py:120-121): state is final
    if (all) {
      for (int j = 0; j < beam; ++j) s_src[j] = j;
    } else {
      float len2[BEAM_MAX];
      for (int j = 0; j < beam; ++j) len2[j] = st.seq_len[base + j] + (st.stopped[base + j] ? 0.f : 1.f);
      float n_sc[BEAM_MAX], n_len[BEAM_MAX];
      int n_tok[BEAM_MAX], n_stop[BEAM_MAX], n_src[BEAM_MAX];
      int head[BEAM_MAX];
      for (int j = 0; j < beam; ++j) head[j] = 0;
      for (int r = 0; r < beam; ++r) {
        float bv = -INFINITY;
        long long bflat = 0x7fffffffffffffffLL;
        int bj = -1;
        for (int j = 0; j < beam; ++j) {
          if (head[j] >= beam) continue;
          const int c = (base + j) * beam + head[j];
          const float avg = (st.scores[base + j] + cand_val[c]) / len2[j];  // scores_sum / seq_lengths (base.py:99-101)
          const long long flat = static_cast<long long>(j) * V + cand_idx[c];
          if (bj < 0 || avg > bv || (avg == bv && flat < bflat)) {
            bv = avg;
            bflat = flat;
            bj = j;
          }
        }
        const int c = (base + bj) * beam + head[bj];
        head[bj]++;
        n_src[r] = bj;
        n_tok[r] = cand_idx[c];
        n_len[r] = len2[bj];
        n_sc[r] = bv * len2[bj];  // scores = scores_sum_average * seq_lengths (base.py:113)
        n_stop[r] = (st.stopped[base + bj] != 0 || n_tok[r] == stop_token) ? 1 : 0;
      }
      for (int r = 0; r < beam; ++r) {
        st.scores[base + r] = n_sc[r];
        st.seq_len[base + r] = n_len[r];
        st.stopped[base + r] = n_stop[r];
        s_src[r] = n_src[r];
        st.tokens[out][static_cast<long long>(base + r) * entry_len + step] = n_tok[r];
      }
    }
  }
  __syncthreads();
  const bool frozen = s_frozen != 0;
  // tokens: out[r][0..step-1] = in[src][..] (frozen: the whole row incl. position `step` stays as it was)
  const int ncopy = frozen ? entry_len : step;
  for (int i = threadIdx.x; i < beam * ncopy; i += blockDim.x) {
    const int r = i / ncopy, t = i % ncopy;
    st.tokens[out][static_cast<long long>(base + r) * entry_len + t] =
        st.tokens[in][static_cast<long long>(base + s_src[r]) * entry_len + t];
  }
  // ancestry: positions < pos inherited from the source beam; position pos was written into the source beam's slot
  for (int i = threadIdx.x; i < beam * (pos + 1); i += blockDim.x) {
    const int r = i / (pos + 1), t = i % (pos + 1);
    const int src = base + s_src[r];
    st.anc[out][static_cast<long long>(base + r) * t_max + t] =
        t < pos ? st.anc[in][static_cast<long long>(src) * t_max + t] : src;
  }
}

// base.py:123-128: scores / seq_lengths, best beam, its tokens and length.
__global__ void beam_final_kernel(BeamState st, int cur, int beam, int entry_len, int32_t* __restrict__ tokens,
                                  int32_t* __restrict__ lengths, float* __restrict__ scores, int n_img) {
  __shared__ int s_best;
  const int b = blockIdx.x;
  if (b >= n_img) return;
  const int base = b * beam;
  if (threadIdx.x == 0) {
    int best = 0;
    float bv = st.scores[base] / st.seq_len[base];
    for (int j = 1; j < beam; ++j) {
      const float v = st.scores[base + j] / st.seq_len[base + j];
      if (v > bv) {
        bv = v;
        best = j;
      }
    }
    s_best = best;
    lengths[b] = static_cast<int32_t>(st.seq_len[base + best]);
    scores[b] = bv;
  }
  __syncthreads();
  const int len = static_cast<int>(st.seq_len[base + s_best]);
  for (int t = threadIdx.x; t < entry_len; t += blockDim.x)
    tokens[b * entry_len + t] = t < len ? st.tokens[cur][static_cast<long long>(base + s_best) * entry_len + t] : 0;
}

}  // namespace

int gen_reset_run(int32_t* stopped, int32_t* lengths, unsigned long long* keys, float* scores, int n, cudaStream_t s) {
  gen_reset_kernel<<<(n + 255) / 256, 256, 0, s>>>(stopped, lengths, keys, scores, n);
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

int greedy_select_run(unsigned long long* keys, int32_t* tokens, int entry_len, int step, int32_t* stopped,
                      int32_t* lengths, int stop_token, int n, cudaStream_t s) {
  CC_CUDA(launch_pdl(greedy_select_kernel, dim3((n + 255) / 256), dim3(256), 0, s, keys, tokens, entry_len, step, stopped,
                     lengths, stop_token, n));
  return CC_OK;
}

int row_topk_run(const float* logits, int64_t ldl, int V, float inv_temp, int beam, const int32_t* stopped,
                 float* out_val, int32_t* out_idx, int rows, cudaStream_t s) {
  CC_REQUIRE(beam >= 1 && beam <= kMaxBeam, CC_ESHAPE, "beam size %d outside 1..%d", beam, kMaxBeam);
  // 256-thread CTAs fit 8 to an SM: from 1185 rows on a second, nearly empty wave would double the time, so big beam
  // batches run 128-thread CTAs (16 per SM) and the rows spread evenly.
  if (rows > num_sms() * 8)
    row_topk_kernel<kMaxBeam, 128><<<rows, 128, 0, s>>>(logits, ldl, V, inv_temp, beam, stopped, out_val, out_idx);
  else
    row_topk_kernel<kMaxBeam, 256><<<rows, 256, 0, s>>>(logits, ldl, V, inv_temp, beam, stopped, out_val, out_idx);
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

int beam_init_run(const float* cand_val, const int32_t* cand_idx, const BeamState& st, int beam, int entry_len,
                  int t_max, int Tp, int stop_token, int n_img, cudaStream_t s) {
  beam_init_kernel<<<n_img, 64, 0, s>>>(cand_val, cand_idx, st, beam, entry_len, t_max, Tp, stop_token, n_img);
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

int beam_step_run(const float* cand_val, const int32_t* cand_idx, const BeamState& st, int in, int beam, int V,
                  int entry_len, int t_max, int step, int pos, int stop_token, int n_img, cudaStream_t s) {
  beam_step_kernel<kMaxBeam><<<n_img, 128, 0, s>>>(cand_val, cand_idx, st, in, beam, V, entry_len, t_max, step, pos,
                                                    stop_token, n_img);
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

int beam_final_run(const BeamState& st, int cur, int beam, int entry_len, int32_t* tokens, int32_t* lengths,
                   float* scores, int n_img, cudaStream_t s) {
  beam_final_kernel<<<n_img, 64, 0, s>>>(st, cur, beam, entry_len, tokens, lengths, scores, n_img);
  CC_CUDA(cudaGetLastError());
  return CC_OK;
}

}  // namespace cc
