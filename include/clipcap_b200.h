/*
 * libclipcap_b200 — C ABI of the B200-native ClipCap hot path
 * (image -> CLIP ViT-L/14 embedding -> mapper prefix -> GPT-2 autoregressive decode -> token ids).
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types. Every entry point below names the reference
 * interface (TheoCoombes/ClipCap @ f041e35, paths relative to the reference repo) that it replaces; the Python package
 * `clipcap_b200` binds it with ctypes and mirrors the reference's Python API on top.
 *
 * Conventions
 *  - Every function returns a status (CC_OK == 0, negative on error); cc_last_error() returns a thread-local message.
 *  - All data pointers are DEVICE pointers on the device that was current at *_create time unless stated otherwise
 *    (weights passed to *_create may be host or device pointers).
 *  - Nothing in a forward call synchronises the host: work is enqueued on the caller's stream (a cudaStream_t passed
 *    as void*; NULL = legacy default stream).
 *  - A handle owns its packed fp16 weights, workspaces, KV cache and CUDA graphs, all allocated at create time for
 *    `max_batch`. One call in flight per handle.
 *  - The library refuses to run on anything but compute capability 10.x (CC_EARCH): there is no fallback path.
 */
#ifndef CLIPCAP_B200_H
#define CLIPCAP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum cc_status {
  CC_OK = 0,
  CC_EINVAL = -1, /* bad argument */
  CC_ESHAPE = -2, /* unsupported / inconsistent shape */
  CC_EALIGN = -3, /* pointer or stride alignment */
  CC_ECUDA = -4,  /* CUDA runtime / driver error */
  CC_ENCCL = -5,  /* NCCL missing or a collective failed */
  CC_EARCH = -6,  /* device is not sm_100 */
  CC_ENOMEM = -7
};

enum cc_dtype { CC_F32 = 0, CC_F16 = 1 };

/* A named weight tensor in the reference's state_dict layout (fp32, contiguous, row-major). */
typedef struct cc_tensor {
  const char* name;
  const void* data; /* host or device pointer */
  int32_t dtype;    /* CC_F32 only */
  int32_t ndim;
  int64_t shape[4];
} cc_tensor;

const char* cc_last_error(void);
/* "clipcap_b200 <version> sm_100a" — also proves the library was built for the right arch. */
const char* cc_version(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Stage 1 — CLIP ViT image tower.
 * Replaces: CLIPModel.forward -> clip_model.encode_image (clipcap/encoders/clip.py:112-129, :120), i.e. OpenAI CLIP
 * VisionTransformer: conv1 patch embed (no bias) -> [CLS]+pos -> ln_pre -> L x {x += MHA(ln_1 x); x += c_proj(
 * QuickGELU(c_fc(ln_2 x)))} -> ln_post(x[:,0]) @ proj, optional L2 normalisation (clip.py:122-123).
 * Weight names (OpenAI `clip` state_dict, the dependency the reference loads at clip.py:134-136):
 *   visual.conv1.weight [w,3,p,p]  visual.class_embedding [w]  visual.positional_embedding [T,w]
 *   visual.ln_pre.{weight,bias}  visual.transformer.resblocks.{i}.{ln_1,ln_2}.{weight,bias}
 *   ...resblocks.{i}.attn.in_proj_{weight [3w,w],bias}  ...attn.out_proj.{weight,bias}
 *   ...resblocks.{i}.mlp.{c_fc,c_proj}.{weight,bias}  visual.ln_post.{weight,bias}  visual.proj [w,out]
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct cc_vit_cfg {
  int32_t image_size; /* 224 */
  int32_t patch;      /* 14  */
  int32_t width;      /* 1024 */
  int32_t layers;     /* 24 */
  int32_t heads;      /* 16 (head dim must be 64) */
  int32_t mlp_dim;    /* 4096 */
  int32_t out_dim;    /* 768 */
  float eps;          /* 1e-5 */
} cc_vit_cfg;
typedef struct cc_vit cc_vit;

int cc_vit_create(cc_vit** h, const cc_vit_cfg* cfg, const cc_tensor* weights, int n_weights, int max_batch);
/* pixels: [B,3,S,S] CLIP-normalised, dtype pix_dtype; out: [B,out_dim], dtype out_dtype. */
int cc_vit_forward(cc_vit* h, const void* pixels, int pix_dtype, int B, int normalize, void* out, int out_dtype,
                   void* stream);
/* number of kernel launches the last cc_vit_forward enqueued */
int cc_vit_last_launches(cc_vit* h);
void cc_vit_destroy(cc_vit* h);

/* ------------------------------------------------------------------------------------------------------------------
 * Stage 2 — prefix mapper.
 * Replaces: TransformerMapper.forward (clipcap/model/mapper.py:113-130), TransformerMapperWindowed.forward
 * (mapper.py:133-160) with Transformer/TransformerLayer/MLPTransformer (mapper.py:8-110) and MultiHeadAttention
 * (clipcap/model/attention.py:4-43); plus the upstream-defined MLP mapper (absent from the reference).
 * Weight names are relative to `transformer_mapper.`:
 *   linear.{weight [P*d,E],bias}  prefix_const [K,d]  [pos_embeddings [(W)*P,d]]
 *   transformer.layers.{i}.{norm1,norm2}.{weight,bias}  ...attn.to_queries.weight [d,d]
 *   ...attn.to_keys_values.weight [2d,d]  ...attn.project.{weight,bias}  ...mlp.{fc1,fc2}.{weight,bias}
 * MLP kind: model.0.{weight [K*d/2,E],bias}  model.2.{weight [K*d,K*d/2],bias}
 * ------------------------------------------------------------------------------------------------------------------ */
enum cc_mapper_kind { CC_MAPPER_TRANSFORMER = 0, CC_MAPPER_WINDOWED = 1, CC_MAPPER_MLP = 2 };
typedef struct cc_mapper_cfg {
  int32_t kind;
  int32_t E;       /* encoder_embedding_size */
  int32_t d;       /* lm_embedding_size */
  int32_t P;       /* projection_length */
  int32_t K;       /* prefix_length */
  int32_t H;       /* transformer_attention_heads (d/H in {48,64,96,128}) */
  int32_t L;       /* transformer_layers */
  int32_t W;       /* windowed: window_size + 1 (model.py:28); otherwise 1 */
  int32_t use_pos; /* windowed: use_positional_embeddings */
  float eps;       /* 1e-5 */
} cc_mapper_cfg;
typedef struct cc_mapper cc_mapper;

int cc_mapper_create(cc_mapper** h, const cc_mapper_cfg* cfg, const cc_tensor* weights, int n_weights, int max_batch);
/* emb: [B,E] (or [B,W,E] windowed); prefix: [B,K,d]. */
int cc_mapper_forward(cc_mapper* h, const void* emb, int emb_dtype, int B, void* prefix, int prefix_dtype,
                      void* stream);
/* number of kernel launches the last cc_mapper_forward enqueued */
int cc_mapper_last_launches(cc_mapper* h);
void cc_mapper_destroy(cc_mapper* h);

/* ------------------------------------------------------------------------------------------------------------------
 * Stage 3 — GPT-2 language model and the decode loops.
 * Replaces: model.language_model(inputs_embeds=...) .logits and get_input_embeddings() as used by
 * clipcap/inference/base.py:76,81,117 (HF GPT2LMHeadModel arithmetic), and generate_beam (base.py:55-132; beam_size=1
 * is the reference's greedy).
 * Weight names are relative to `language_model.` (HF GPT-2): transformer.wte.weight [V,d]  transformer.wpe.weight
 *   transformer.h.{i}.{ln_1,ln_2}.{weight,bias}  ...attn.c_attn.{weight [d,3d],bias}  ...attn.c_proj.{weight,bias}
 *   ...mlp.c_fc.{weight [d,4d],bias}  ...mlp.c_proj.{weight [4d,d],bias}  transformer.ln_f.{weight,bias}
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct cc_gpt2_cfg {
  int32_t d;     /* n_embd */
  int32_t L;     /* n_layer */
  int32_t H;     /* n_head (d/H must be 64) */
  int32_t V;     /* vocab_size */
  int32_t n_pos; /* n_positions */
  float eps;     /* 1e-5 */
} cc_gpt2_cfg;
typedef struct cc_gpt2 cc_gpt2;

/* max_seqs: largest B (x beam) ever decoded; max_len: largest prefix + generated length. */
int cc_gpt2_create(cc_gpt2** h, const cc_gpt2_cfg* cfg, const cc_tensor* weights, int n_weights, int max_seqs,
                   int max_len);
/* embeds: [B,T,d] -> logits fp32. all_positions == 0: [B,V] of the last position (what base.py:83 consumes);
 * otherwise [B,T,V]. */
int cc_gpt2_logits(cc_gpt2* h, const void* embeds, int dtype, int B, int T, int all_positions, float* logits,
                   void* stream);
/* out[i,:] = wte[ids[i],:] (get_input_embeddings(), base.py:76,117) */
int cc_gpt2_embed(cc_gpt2* h, const int32_t* ids, int n, void* out, int out_dtype, void* stream);

/* GREEDY / BEAM: generate_beam (clipcap/inference/base.py:55-132).
 * NUCLEUS: generate_nucleus_sampling (clipcap/inference/nucleus_sampling.py:9-75) — top-k / top-p over softmax(logits/T),
 *          the stop token is part of the returned tokens.
 * SAMPLE:  generate_no_beam (clipcap/inference/no_beam.py:10-82) with clipcap/inference/utils.py:5-49 — repetition
 *          penalty, temperature, top_k_top_p_filtering, "sentence length penalty", multinomial; the stop token ends the
 *          caption without being returned.
 * The two sampling modes draw from the same distribution as the reference with a Philox stream keyed by (seed, row,
 * step); results are reproducible for a given seed but not bit-identical to torch.multinomial's generator. */
enum cc_gen_mode { CC_GEN_GREEDY = 0, CC_GEN_BEAM = 1, CC_GEN_NUCLEUS = 2, CC_GEN_SAMPLE = 3 };
typedef struct cc_gen_cfg {
  int32_t mode;
  int32_t beam;         /* beam_size (BEAM) */
  int32_t entry_length; /* tokens to generate (base.py:62) */
  float temperature;    /* <= 0 treated as 1 (base.py:83) */
  int32_t stop_token;   /* tokenizer.encode(eos)[0] (base.py:66) / tokenizer.encode(".")[0] (nucleus_sampling.py:21) */
  /* sampling modes only */
  float top_p;                     /* nucleus_sampling.py:16 / no_beam.py:16 (no_beam: <= 0 disables) */
  int32_t top_k;                   /* 0 = whole vocabulary */
  float repetition_penalty;        /* SAMPLE: no_beam.py:20 (1.0 disables) */
  int32_t desired_sentence_length; /* SAMPLE: no_beam.py:21 */
  float sentence_length_factor;    /* SAMPLE: no_beam.py:22 */
  int32_t n_history;               /* SAMPLE: number of text-prefix tokens that start every row's token history */
  const int32_t* history;          /* SAMPLE: those tokens (HOST pointer, may be NULL when n_history == 0) */
  uint64_t seed;
} cc_gen_cfg;
/* prefix: [B,Tp,d] input embeddings (mapper prefix, optionally followed by text-prefix embeddings, base.py:75-77).
 * tokens: [B,entry_length] int32 — best beam per row (base.py:126-128); lengths: [B] int32 = number of valid tokens
 * (stop token included, base.py:125); scores: [B] fp32 length-normalised log-prob of the returned beam (may be NULL).
 * Row i equals the reference called on sample i alone. */
int cc_generate(cc_gpt2* h, const void* prefix, int dtype, int B, int Tp, const cc_gen_cfg* g, int32_t* tokens,
                int32_t* lengths, float* scores, void* stream);
/* The same call in two halves, for callers that run the tensor-bound prefill and the latency-bound decode loop on
 * different streams / SM partitions (cc_partition_*): cc_generate_prefill runs `h = prefix + wpe`, the prefill of all Tp
 * positions and the selection of the FIRST token (base.py:80-94 at step 0); cc_generate_decode runs the remaining
 * entry_length - 1 steps and copies out the results exactly as cc_generate does. The caller orders the decode stream
 * after the prefill (event) and does not touch the handle in between; results equal cc_generate's. */
int cc_generate_prefill(cc_gpt2* h, const void* prefix, int dtype, int B, int Tp, const cc_gen_cfg* g, void* stream);
int cc_generate_decode(cc_gpt2* h, int B, int Tp, const cc_gen_cfg* g, int32_t* tokens, int32_t* lengths, float* scores,
                       void* stream);
/* Balance between the two halves: the last `blocks` transformer blocks of the prefill, the LM head and the first-token
 * selection move from cc_generate_prefill to the head of cc_generate_decode (same kernels, same order, same results) —
 * for callers whose prefill-side stream is the busier one. 0 (default) restores the split described above. Does not
 * affect cc_generate. */
int cc_gpt2_set_prefill_defer(cc_gpt2* h, int blocks);
/* number of kernel launches (graph nodes) the last cc_generate (or prefill + decode pair) enqueued */
int cc_gpt2_last_launches(cc_gpt2* h);
void cc_gpt2_destroy(cc_gpt2* h);

/* ------------------------------------------------------------------------------------------------------------------
 * Stage 1 (audio) — CLAP audio tower (HTSAT-tiny Swin encoder + audio projection), BASELINE configs[4].
 * Replaces: CLAPModel.forward -> clap_model.get_audio_embedding_from_data (clipcap/encoders/clap.py:105-131; laion_clap
 * is not installed, the arithmetic is transformers' ClapAudioModelWithProjection, modeling_clap.py:1725-1775 with
 * ClapAudioEncoder.forward :814-918). Weight names are that module's state_dict keys (audio_model.audio_encoder.*,
 * audio_projection.linear{1,2}.*). Samples flagged `is_longer` (clips longer than the 10 s window; 4 mel views: the
 * global one and three local crops) go through the feature-fusion patch embedding (mel_conv2d + AFF block,
 * modeling_clap.py:296-344, :238-245); it needs the `patch_embed.mel_conv2d.*` / `patch_embed.fusion_model.*` tensors.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct cc_clap_cfg {
  int32_t num_mel_bins;   /* 64 */
  int32_t spec_size;      /* 256: side of the folded spectrogram image */
  int32_t patch;          /* 4 (size == stride) */
  int32_t embed;          /* 96: channels of stage 0; stage i has embed << i */
  int32_t depths[4];      /* 2 2 6 2 */
  int32_t heads[4];       /* 4 8 16 32 (head dim must be 24) */
  int32_t window;         /* 8 */
  int32_t projection_dim; /* 512 */
  float eps;              /* 1e-5 */
} cc_clap_cfg;
typedef struct cc_clap cc_clap;

int cc_clap_create(cc_clap** h, const cc_clap_cfg* cfg, const cc_tensor* weights, int n_weights, int max_batch);
/* mel: [B, channels, T <= spec_size * spec_size / num_mel_bins, num_mel_bins] log-mel features, dtype mel_dtype (channel 0 =
 * global view; channels 1..3 = local views, read for flagged samples only); is_longer: HOST array of B bytes (non-zero =
 * fuse the local views of that sample) or NULL for none; out: [B, projection_dim]. stop_after_stage >= 0 with dump != NULL copies the fp32 token stream after that
 * stage's blocks ([B * tokens, channels] of the stage) into `dump` and returns without producing `out` (debug / tests). */
int cc_clap_forward(cc_clap* h, const void* mel, int mel_dtype, const unsigned char* is_longer, int B, int channels, int T,
                    int normalize, void* out, int out_dtype, int stop_after_stage, float* dump, void* stream);
int cc_clap_last_launches(cc_clap* h);
void cc_clap_destroy(cc_clap* h);

/* ------------------------------------------------------------------------------------------------------------------
 * Training step of the prefix mapper with the language model frozen.
 * Replaces: ClipCapModel.forward + training_step (clipcap/model/model.py:43-58, 94-113) followed by loss.backward() for
 * ClipCapModelPrefixOnly (model.py:116-123: only transformer_mapper parameters train, the LM stays in eval mode):
 *   tokens < 0 are padding (set to 0, model.py:103-104); logits[:, prefix_length-1:-1] are scored against the tokens with
 *   cross-entropy, ignore_index = 0 (model.py:109-110), mean over the scored tokens.
 * `lm_weights`: `language_model.*` tensors as for cc_gpt2_create. `params` / `grads`: the caller's CURRENT fp32
 * transformer_mapper tensors and same-named fp32 buffers that receive d loss / d param (overwritten, not accumulated) —
 * all DEVICE pointers, names as for cc_mapper_create. grads == NULL: forward + loss only (validation).
 * Padding must be trailing (as the reference's dataloader produces): under the causal mask the scored positions then
 * never see a padded key, which is what HF's attention_mask (model.py:52-56) enforces.
 * loss_scale (> 0, power of two; <= 0 picks 1024): activation gradients travel between GEMMs in fp16 scaled by this
 * factor; parameter gradients are returned unscaled. loss: DEVICE pointer to one fp32.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct cc_train cc_train;
int cc_train_create(cc_train** h, const cc_mapper_cfg* mapper_cfg, const cc_gpt2_cfg* lm_cfg, const cc_tensor* lm_weights,
                    int n_weights, int max_batch, int max_tokens);
int cc_train_step(cc_train* h, const cc_tensor* params, int n_params, const cc_tensor* grads, int n_grads,
                  const void* emb /*[B,E]*/, int emb_dtype, const int32_t* tokens /*[B,Tt]*/, int B, int Tt, float loss_scale,
                  float* loss, void* stream);
/* Number of non-finite elements among the gradients the last cc_train_step with gradients wrote (synchronises `stream`).
 * Activation gradients travel in fp16 under the caller's static loss scale; a non-zero count means that scale overflowed:
 * halve it and repeat the step (cc_op_adamw skips non-finite gradient elements, so the optimiser state is never poisoned). */
int cc_train_last_nonfinite(cc_train* h, void* stream);
int cc_train_last_launches(cc_train* h);
void cc_train_destroy(cc_train* h);
/* torch.optim.AdamW update on flat fp32 device arrays (configure_optimizers, model.py:67-91); step counts from 1. */
int cc_op_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                float weight_decay, int step, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Kernel-level test hooks (used by tests/ and bench.py for per-kernel parity and roofline timing; same kernels the
 * engines launch).
 * ------------------------------------------------------------------------------------------------------------------ */
/* C[M,N] = A[M,K] (fp16, row stride lda) x W[N,K]^T (fp16, row stride K) with epilogue `epi` (see csrc/common.h):
 * 0..4 fp16 out with none/relu/quickgelu/gelu_new/tanh, 5 fp32 out, 6 fp32 in-place residual add, 7 fused argmax
 * (out = uint64 keys[M], must be zeroed), 10 fp16 out with the exact (erf) GELU. bn = 0 picks BLOCK_N heuristically. */
int cc_op_gemm(const void* a, int64_t lda, const void* w, const float* bias, void* out, int64_t ldc, int M, int N, int K,
               int epi, int bn, void* stream);
int cc_op_layernorm(const float* x, int64_t x_ld, const float* gamma, const float* beta, void* y, int64_t y_ld, int rows,
                    int d, float eps, void* stream);
int cc_op_attention(const void* q, const void* k, const void* v, int64_t ld, void* o, int64_t ldo, int B, int S, int H,
                    int hd, int causal, float scale, void* stream);
/* One decode step of attention: q, k, v of the step in qkv[nseq, 3 * H * 64]; caches [slot][head][t_max][64] fp16 whose
 * 128-byte rows keep their eight 16-byte chunks rotated (chunk c of position t at chunk c ^ (t & 7): bank-conflict-free
 * tensor-core operand fetches after a plain bulk copy, csrc/attention.cu kv_chunk); the step's k, v are appended at `pos`. */
int cc_op_decode_attention(const void* qkv, void* kcache, void* vcache, const int32_t* anc, void* o, int nseq, int H,
                           int t_max, int pos, float scale, void* stream);
/* The same step for beam search: rows seq = image * beam + b; positions 0 .. shared_len-1 of every beam of an image live in
 * the slot anc names for position 0 (the image's prefix), later positions wherever anc[seq * t_max + t] says. */
int cc_op_decode_attention_beam(const void* qkv, void* kcache, void* vcache, const int32_t* anc, void* o, int nseq, int H,
                                int t_max, int pos, int beam, int shared_len, float scale, void* stream);
/* Window tiling of a decoded square image (CLIPTransform.tile_image, clipcap/encoders/clip.py:60-82:
 * tensor.unfold(1, p, step).unfold(2, p, step)): image [3, size, size] fp32 on the device ->
 * tiles [tiles_per_axis^2, 3, p, p], tile (ty, tx) = pixels [ty*step, ty*step + p) x [tx*step, tx*step + p), row-major over
 * (ty, tx). step = p without overlap, floor(p * (1 - overlap / 100)) with (clip.py:65-68). */
int cc_op_tile_image(const float* image, int size, int tiles_per_axis, int pixels_per_tile, int step, float* tiles,
                     void* stream);

/* One sampling step (the token-selection kernel of the NUCLEUS / SAMPLE modes) on given logits [rows, V] fp32:
 * tokens[row * entry_length + step] receives the draw; stopped / lengths [rows] int32 are updated; cfg->history is a
 * DEVICE pointer here. Used by the distribution tests. */
int cc_op_sample(const float* logits, int rows, int V, const cc_gen_cfg* cfg, int step, int32_t* tokens,
                 int32_t* stopped, int32_t* lengths, void* stream);

/* Skinny GEMM of the single-image / small-beam decode step (M <= 16 rows, csrc/skinny.cu):
 * out[M,N] = X W^T with W [N,K] fp16 and X = LayerNorm(x32; gamma, beta, eps) when x32 != NULL (fp32 rows, stride ldx)
 * or the fp16 rows x16. epi as for cc_op_gemm: 0 fp16, 3 fp16 gelu_new, 5 fp32, 6 fp32 in-place residual add, 7 argmax keys. */
int cc_op_skinny_gemm(const float* x32, const float* gamma, const float* beta, float eps, const void* x16, int64_t ldx, int M,
                      const void* w, int N, int K, int epi, const float* bias, void* out, int64_t ldc, void* stream);

/* Backward of cc_op_attention (training step): given d_o = d loss / d o, writes dq, dk, dv in the layout of q, k, v
 * (row stride ldd). */
int cc_op_attention_bwd(const void* q, const void* k, const void* v, int64_t ld, const void* d_o, int64_t ldo, void* dq,
                        void* dk, void* dv, int64_t ldd, int B, int S, int H, int hd, int causal, float scale,
                        void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Multi-GPU: the one exchange step of the path. One process per GPU; every rank encodes + maps its own images
 * and writes its [B_local, K, d] prefix block into slot `rank` of a [nranks * B_local, K, d] buffer (pass that slot as
 * cc_mapper_forward's `out`); cc_allgather_prefix completes the buffer on every rank with ONE in-place ncclAllGather over
 * NVLink, enqueued on `stream` (no host sync). The 128-byte id comes from cc_comm_unique_id on rank 0 and reaches the
 * other ranks through the launcher (torch.distributed store, MPI, a file). NCCL is bound with dlopen at first use;
 * CC_ENCCL when it is missing or a call fails. No reference counterpart: its inference is single-device
 * (clipcap/inference/args.py:22-27); the collective is the one BASELINE.json's north_star adds. */
typedef struct cc_comm cc_comm;
int cc_comm_unique_id(void* id_out /* 128 bytes */);
int cc_comm_create(cc_comm** c, const void* unique_id /* 128 bytes */, int rank, int nranks);
int cc_allgather_prefix(cc_comm* c, void* prefix_all, size_t bytes_per_rank, void* stream);
int cc_comm_rank(cc_comm* c);
int cc_comm_nranks(cc_comm* c);
int cc_nccl_version(void); /* NCCL_VERSION_CODE of the library bound, 0 if none */
void cc_comm_destroy(cc_comm* c);

/* ---------------------------------------------------------------------------------------------------------------
 * SM partitions for the two-stage serving pipeline (clipcap_b200/pipeline.py). The path's stages have opposite
 * characters on a B200: image tower + mapper + prefill are tensor- and power-bound, the decode loop is a chain of
 * dependent launches that leaves most SMs idle. cc_partition_create splits the device's SMs (CUDA green contexts) into
 * a large partition (which = 0) and a small one (which = 1, >= small_sms SMs; multiples of 8 follow the GPC hierarchy, other
 * counts use the driver's finer, hierarchy-agnostic split) and creates
 * one non-blocking stream in each; work enqueued on a partition's stream runs on its SMs only, so the decode of batch i
 * overlaps the image tower of batch i+1 without either stalling the other. No reference counterpart (single stream,
 * clipcap/inference/demo.py:30-45). */
typedef struct cc_partition cc_partition;
int cc_partition_create(cc_partition** p, int device, int small_sms);
void* cc_partition_stream(cc_partition* p, int which); /* cudaStream_t owned by the partition */
int cc_partition_sms(cc_partition* p, int which);      /* SMs provisioned for that partition */
void cc_partition_destroy(cc_partition* p);

/* SM budget of the calling process' next launches: the number of SMs the stream it enqueues on may use (0 = the whole
 * device). A caller that splits the GPU into SM partitions (CUDA green contexts: the tensor-bound image tower of batch
 * i+1 on one partition, the latency-bound decode loop of batch i on the other — clipcap_b200/pipeline.py) sets the
 * budget of a partition before enqueuing on its stream; persistent grids, tile shapes and split-K factors are sized for
 * it. Graphs captured by cc_generate are cached per budget. The reference has no counterpart (single stream,
 * clipcap/inference/demo.py:30-45). */
void cc_set_sm_budget(int n_sms);
int cc_get_sm_budget(void);

/* Live per-launch timing of the dominant kernel (the 128x256-tile tcgen05 GEMM): while enabled, every such launch that
 * is not inside a graph capture is bracketed by CUDA events on its own stream. cc_prof_read synchronises the device and
 * returns the summed duration (ms), algorithmic FLOPs (2*M*N*K) and launch count since cc_prof_enable(1). */
void cc_prof_enable(int on);
void cc_prof_read(double* ms, double* flops, long long* n);
/* The same sums for one tile family of the tcgen05 GEMM: bn = 512 the 256 x 256 CTA-pair tile, 256 / 128 / 64 / 32 the
 * single-CTA tiles of that width (the decode and small-batch shapes), 0 every recorded launch. */
void cc_prof_read_family(int bn, double* ms, double* flops, long long* n);

#ifdef __cplusplus
}
#endif
#endif /* CLIPCAP_B200_H */
