"""CPU restatement of the ClipCap hot path (image -> ViT-L/14 -> mapper -> GPT-2 decode) in plain fp32 torch ops.

TEST INFRASTRUCTURE — not product code. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module, and only as the checker. The product path (clipcap_b200) never imports
it and has no CPU fallback.

Parity pinning:
  * mapper, ClipCapModel.forward, generate_beam: pinned against the reference's own modules executed in the build
    container (oracle/ref_runner.py imports /root/reference unmodified) — live in tests/test_oracle_vs_reference.py and
    frozen in tests/golden/*.npz (tests/golden/make_golden.py).
  * GPT-2 arithmetic: third-party `transformers` (unpinned in the reference's requirements.txt:3; 5.5.0 installed) —
    pinned against transformers.GPT2LMHeadModel.
  * CLIP ViT arithmetic: third-party OpenAI `clip` (unpinned git dependency, requirements-clip.txt:1, NOT installed) —
    restated from its published VisionTransformer and pinned against transformers.CLIPVisionModelWithProjection
    (quick_gelu), which is the same network with split q/k/v. The reference holds no test or fixture for the encoder:
    PARITY UNPINNED by the reference for this stage.
  * MLP mapper: absent from the reference; upstream rmokady/CLIP_prefix_caption definition.
    PARITY UNPINNED.

Every function takes a flat dict of fp32 tensors keyed by the reference's state_dict names.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Weights = Dict[str, torch.Tensor]


# ------------------------------------------------------------------------------------------------ configs
@dataclass
class VitCfg:
    image_size: int = 224
    patch: int = 14
    width: int = 1024
    layers: int = 24
    heads: int = 16
    mlp_dim: int = 4096
    out_dim: int = 768
    eps: float = 1e-5


@dataclass
class MapperCfg:
    kind: str = "transformer"  # transformer | windowed | mlp
    E: int = 768   # encoder_embedding_size
    d: int = 1024  # lm_embedding_size
    P: int = 10    # projection_length
    K: int = 40    # prefix_length
    H: int = 8     # transformer_attention_heads
    L: int = 8     # transformer_layers
    W: int = 1     # windowed: window_size + 1 (clipcap/model/model.py:28)
    use_pos: bool = False
    eps: float = 1e-5


@dataclass
class Gpt2Cfg:
    d: int = 1024
    L: int = 24
    H: int = 16
    V: int = 50257
    n_pos: int = 1024
    eps: float = 1e-5


def _ln(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


# ------------------------------------------------------------------------------------------------ stage 1: ViT
def vit_encode(w: Weights, pixels: torch.Tensor, cfg: VitCfg, normalize: bool = False) -> torch.Tensor:
    """clip_model.encode_image behind CLIPModel.forward (clipcap/encoders/clip.py:112-129, call at :120; optional
    normalisation :122-123). OpenAI clip/model.py VisionTransformer semantics, weight names of its state_dict."""
    B = pixels.shape[0]
    x = F.conv2d(pixels.float(), w["visual.conv1.weight"], None, stride=cfg.patch)  # [B, w, g, g], no bias
    x = x.reshape(B, cfg.width, -1).permute(0, 2, 1)                                 # [B, g*g, w], row-major grid
    cls = w["visual.class_embedding"].reshape(1, 1, -1).expand(B, 1, -1)
    x = torch.cat([cls, x], dim=1) + w["visual.positional_embedding"]
    x = _ln(x, w["visual.ln_pre.weight"], w["visual.ln_pre.bias"], cfg.eps)
    hd = cfg.width // cfg.heads
    for l in range(cfg.layers):
        p = f"visual.transformer.resblocks.{l}."
        y = _ln(x, w[p + "ln_1.weight"], w[p + "ln_1.bias"], cfg.eps)
        qkv = y @ w[p + "attn.in_proj_weight"].t() + w[p + "attn.in_proj_bias"]
        q, k, v = [t.reshape(B, -1, cfg.heads, hd).transpose(1, 2) for t in qkv.chunk(3, dim=-1)]
        att = torch.softmax((q @ k.transpose(-1, -2)) * (hd ** -0.5), dim=-1)
        o = (att @ v).transpose(1, 2).reshape(B, -1, cfg.width)
        x = x + o @ w[p + "attn.out_proj.weight"].t() + w[p + "attn.out_proj.bias"]
        y = _ln(x, w[p + "ln_2.weight"], w[p + "ln_2.bias"], cfg.eps)
        y = y @ w[p + "mlp.c_fc.weight"].t() + w[p + "mlp.c_fc.bias"]
        y = y * torch.sigmoid(1.702 * y)  # QuickGELU
        x = x + y @ w[p + "mlp.c_proj.weight"].t() + w[p + "mlp.c_proj.bias"]
    e = _ln(x[:, 0], w["visual.ln_post.weight"], w["visual.ln_post.bias"], cfg.eps) @ w["visual.proj"]
    if normalize:
        e = e / e.norm(dim=-1, keepdim=True)  # clip.py:122-123
    return e


# ------------------------------------------------------------------------------------------------ stage 2: mapper
def _mapper_attention(x, wq, wkv, wp, bp, H):
    """MultiHeadAttention.forward, y=None, mask=None (clipcap/model/attention.py:17-43)."""
    b, n, c = x.shape
    hd = c // H
    q = (x @ wq.t()).reshape(b, n, H, hd)                      # :24 (no bias: TransformerLayer passes bias=False)
    kv = (x @ wkv.t()).reshape(b, n, 2, H, hd)                 # :26-28
    k, v = kv[:, :, 0], kv[:, :, 1]                            # :30
    att = torch.einsum("bnhd,bmhd->bnmh", q, k) * (hd ** -0.5)  # :32
    att = att.softmax(dim=2)                                   # :38 (over keys m)
    out = torch.einsum("bnmh,bmhd->bnhd", att, v).reshape(b, n, c)  # :40
    return out @ wp.t() + bp                                   # :41


def mapper_forward(w: Weights, emb: torch.Tensor, cfg: MapperCfg, relu_mask_fp16: bool = False) -> torch.Tensor:
    """TransformerMapper.forward (clipcap/model/mapper.py:122-130) / TransformerMapperWindowed.forward (:148-160) /
    upstream MLP mapper. Keys are relative to `transformer_mapper.`.
    `relu_mask_fp16` (test instrument, not reference behaviour): the ReLU of the MLP keeps a hidden unit according to the
    pre-activation computed from fp16-ROUNDED operands — the decision a kernel with fp16 GEMM operands takes — while
    values and gradients stay fp32. Used to show that the gradient error behind the ReLU mask comes from units whose
    pre-activation is within operand rounding of zero, not from the arithmetic."""
    B = emb.shape[0]
    if cfg.kind == "mlp":
        h = torch.tanh(emb @ w["model.0.weight"].t() + w["model.0.bias"])
        return (h @ w["model.2.weight"].t() + w["model.2.bias"]).view(B, cfg.K, cfg.d)
    Ptot = cfg.P * (cfg.W if cfg.kind == "windowed" else 1)
    x = (emb @ w["linear.weight"].t() + w["linear.bias"]).view(B, Ptot, -1)        # :123 / :149
    if cfg.kind == "windowed" and cfg.use_pos:
        x = x + w["pos_embeddings"].unsqueeze(0)                                     # :151-153
    x = torch.cat([x, w["prefix_const"].unsqueeze(0).expand(B, -1, -1)], dim=1)    # :125-126
    for l in range(cfg.L):                                                         # Transformer.forward :55-67
        p = f"transformer.layers.{l}."
        y = _ln(x, w[p + "norm1.weight"], w[p + "norm1.bias"], cfg.eps)
        x = x + _mapper_attention(y, w[p + "attn.to_queries.weight"], w[p + "attn.to_keys_values.weight"],
                                  w[p + "attn.project.weight"], w[p + "attn.project.bias"], cfg.H)   # :108
        y = _ln(x, w[p + "norm2.weight"], w[p + "norm2.bias"], cfg.eps)
        pre = y @ w[p + "mlp.fc1.weight"].t() + w[p + "mlp.fc1.bias"]
        if relu_mask_fp16:
            with torch.no_grad():
                keep = (y.half().float() @ w[p + "mlp.fc1.weight"].half().float().t() + w[p + "mlp.fc1.bias"]) > 0
            y = pre * keep
        else:
            y = torch.relu(pre)                                                                       # :82-84
        x = x + y @ w[p + "mlp.fc2.weight"].t() + w[p + "mlp.fc2.bias"]                               # :86, :109
    return x[:, Ptot:]                                                             # :128 / :158


# ------------------------------------------------------------------------------------------------ stage 3: GPT-2
def _gelu_new(x):
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def gpt2_logits(w: Weights, embeds: torch.Tensor, cfg: Gpt2Cfg, last_only: bool = False) -> torch.Tensor:
    """model.language_model(inputs_embeds=embeds).logits (clipcap/inference/base.py:81-82): HF GPT2LMHeadModel forward
    without cache (transformers/models/gpt2/modeling_gpt2.py). Keys relative to `language_model.`."""
    B, T, d = embeds.shape
    hd = d // cfg.H
    x = embeds.float() + w["transformer.wpe.weight"][:T]
    mask = torch.full((T, T), float("-inf")).triu(1)
    for l in range(cfg.L):
        p = f"transformer.h.{l}."
        y = _ln(x, w[p + "ln_1.weight"], w[p + "ln_1.bias"], cfg.eps)
        qkv = y @ w[p + "attn.c_attn.weight"] + w[p + "attn.c_attn.bias"]  # Conv1D: weight is [in, out]
        q, k, v = [t.reshape(B, T, cfg.H, hd).transpose(1, 2) for t in qkv.split(d, dim=-1)]
        att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd) + mask, dim=-1)
        o = (att @ v).transpose(1, 2).reshape(B, T, d)
        x = x + o @ w[p + "attn.c_proj.weight"] + w[p + "attn.c_proj.bias"]
        y = _ln(x, w[p + "ln_2.weight"], w[p + "ln_2.bias"], cfg.eps)
        y = _gelu_new(y @ w[p + "mlp.c_fc.weight"] + w[p + "mlp.c_fc.bias"])
        x = x + y @ w[p + "mlp.c_proj.weight"] + w[p + "mlp.c_proj.bias"]
    if last_only:
        x = x[:, -1:]
    x = _ln(x, w["transformer.ln_f.weight"], w["transformer.ln_f.bias"], cfg.eps)
    return x @ w["transformer.wte.weight"].t()  # tied head, no bias


def generate_beam(w: Weights, cfg: Gpt2Cfg, embeds: torch.Tensor, beam_size: int = 5, entry_length: int = 67,
                  temperature: float = 1.0, stop_token: int = 50256,
                  text_prefix_tokens: Optional[torch.Tensor] = None) -> Tuple[List[int], float, List[dict]]:
    """clipcap/inference/base.py:55-132 for one image (embeds [1, Tp, d]), number_to_generate=1. Returns the best
    beam's token ids (truncated to its length, stop token included), its length-normalised score, and a per-step
    trace (top-1/top-2 logit margin of every live row) used by the margin-aware parity tests. beam_size=1 is the
    reference's greedy decode."""
    assert embeds.shape[0] == 1
    wte = w["transformer.wte.weight"]
    tokens = None
    scores = None
    seq_lengths = torch.ones(beam_size)
    has_stopped = torch.zeros(beam_size, dtype=torch.bool)
    trace: List[dict] = []
    if text_prefix_tokens is not None:
        embeds = torch.cat((embeds, wte[text_prefix_tokens]), dim=1)  # :75-77
    for _ in range(entry_length):
        raw = gpt2_logits(w, embeds, cfg, last_only=True)[:, -1, :]
        top2 = raw.topk(2, -1).values
        trace.append({"margin": (top2[:, 0] - top2[:, 1]).tolist(), "absmax": raw.abs().max().item()})
        logits = raw / (temperature if temperature > 0 else 1.0)     # :83
        logits = logits.softmax(-1).log()                            # :84
        if scores is None:
            scores, next_tokens = logits.topk(beam_size, -1)         # :87
            embeds = embeds.expand(beam_size, *embeds.shape[1:])
            next_tokens, scores = next_tokens.permute(1, 0), scores.squeeze(0)
            tokens = next_tokens
        else:
            logits[has_stopped] = -float("inf")                      # :96
            logits[has_stopped, 0] = 0                               # :97
            scores_sum = scores[:, None] + logits
            seq_lengths[~has_stopped] += 1
            scores_sum_average = scores_sum / seq_lengths[:, None]
            scores_sum_average, next_tokens = scores_sum_average.view(-1).topk(beam_size, -1)
            next_tokens_source = torch.div(next_tokens, scores_sum.shape[1], rounding_mode="trunc")
            seq_lengths = seq_lengths[next_tokens_source]
            next_tokens = (next_tokens % scores_sum.shape[1]).unsqueeze(1)
            tokens = torch.cat((tokens[next_tokens_source], next_tokens), dim=1)
            embeds = embeds[next_tokens_source]
            scores = scores_sum_average * seq_lengths
            has_stopped = has_stopped[next_tokens_source]
        nxt = wte[next_tokens.reshape(-1)].view(embeds.shape[0], 1, -1)  # :117
        embeds = torch.cat((embeds, nxt), dim=1)
        has_stopped = has_stopped + next_tokens.eq(stop_token).reshape(-1)
        if has_stopped.all():
            break
    scores = scores / seq_lengths                                     # :123
    order = scores.argsort(descending=True)
    best = int(order[0])
    length = int(seq_lengths[best])
    return tokens[best, :length].tolist(), float(scores[best]), trace


def generate_greedy_batch(w: Weights, cfg: Gpt2Cfg, prefix: torch.Tensor, entry_length: int, stop_token: int):
    """Row i = generate_beam(beam_size=1) on sample i alone (the reference asserts batch size 1, generate.py:19-20)."""
    out = []
    for i in range(prefix.shape[0]):
        out.append(generate_beam(w, cfg, prefix[i:i + 1], 1, entry_length, 1.0, stop_token))
    return out


def teacher_forced_logits(w: Weights, cfg: Gpt2Cfg, prefix: torch.Tensor, tokens: torch.Tensor) -> torch.Tensor:
    """Last-position logits at every decode step when the sequence is forced along `tokens` [B, n]: returns [B, n, V]
    where [:, s] are the logits that choose token s. One full-sequence forward (causality makes it identical)."""
    wte = w["transformer.wte.weight"]
    Tp = prefix.shape[1]
    emb = torch.cat((prefix.float(), wte[tokens[:, :-1]]), dim=1) if tokens.shape[1] > 1 else prefix.float()
    full = gpt2_logits(w, emb, cfg)
    return full[:, Tp - 1:]


# ------------------------------------------------------------------------------------------------ sampling decode
def nucleus_distribution(logits: torch.Tensor, top_p: Optional[float] = 0.8, top_k: int = 0,
                         temperature: float = 1.0) -> torch.Tensor:
    """The distribution generate_nucleus_sampling draws from (clipcap/inference/nucleus_sampling.py:37-54) for last
    logits [B, V]: softmax(logits / T) -> topk(top_k or V) -> cumsum -> searchsorted(top_p) clipped to top_k - 1 ->
    keep cumulative <= cutoff -> renormalise -> scatter back."""
    logits = logits.float() / (temperature if temperature > 0 else 1.0)
    if top_k == 0:
        top_k = logits.shape[-1]
    if top_p is None:
        top_p = 1.0
    p, idx_sorted = logits.softmax(-1).topk(top_k, dim=-1)
    cum = p.cumsum(-1)
    idx = torch.searchsorted(cum, top_p + torch.zeros(len(p), 1)).clip(max=top_k - 1).reshape(-1)
    cutoffs = cum[torch.arange(len(cum)), idx]
    censored = (cum <= cutoffs[:, None]) * p
    renorm = censored / censored.sum(-1, keepdim=True)
    final = torch.zeros_like(logits)
    final[torch.arange(len(p)).unsqueeze(1).repeat(1, top_k), idx_sorted] = renorm
    return final


def no_beam_distribution(logits: torch.Tensor, history: Optional[torch.Tensor], top_p: float = 0.9, top_k: float = 0.0,
                         temperature: float = 1.0, repetition_penalty: float = 1.2, stop_token: int = 13,
                         desired_sentence_length: int = 50, sentence_length_factor: float = 1.0) -> torch.Tensor:
    """The distribution generate_no_beam draws from (clipcap/inference/no_beam.py:37-62 with utils.py:5-49) for ONE
    row of last logits [V]; `history` = the 1-D `tokens` seen so far (text prefix + generated) or None."""
    logits = logits.float().clone()
    if repetition_penalty != 1.0 and history is not None:            # no_beam.py:41-44, utils.py:34-38
        tok = logits.gather(-1, history)
        tok = torch.where(tok < 0, tok * repetition_penalty, tok / repetition_penalty)
        logits.scatter_(-1, history, tok)
    logits = logits / (temperature if temperature > 0 else 1.0)      # no_beam.py:47
    k = min(int(top_k), logits.size(-1))                             # utils.py:15-20
    if k > 0:
        logits[logits < torch.topk(logits, k)[0][..., -1, None]] = -float("inf")
    if top_p > 0.0:                                                  # utils.py:22-31
        sl, si = torch.sort(logits, descending=True)
        cum = torch.cumsum(sl.softmax(-1), dim=-1)
        rm = cum > top_p
        rm[..., 1:] = rm[..., :-1].clone()
        rm[..., 0] = 0
        logits[si[rm]] = -float("inf")
    if history is not None:                                          # no_beam.py:51-56, utils.py:40-49
        penalty = (history.shape[0] / desired_sentence_length) * sentence_length_factor
        tok = logits.gather(-1, history)
        tok = torch.where(tok == stop_token, tok * penalty, tok)
        logits.scatter_(-1, history, tok)
    return logits.softmax(-1)


def generate_sampling(w: Weights, cfg: Gpt2Cfg, embeds: torch.Tensor, mode: str, pick, entry_length: int = 67,
                      text_prefix_tokens: Optional[torch.Tensor] = None, stop_token: int = 13, **kw) -> List[int]:
    """The loops of generate_nucleus_sampling (mode='nucleus', nucleus_sampling.py:26-73) and generate_no_beam
    (mode='sample', no_beam.py:26-80) for one image and number_to_generate=1, with the random draw abstracted as
    `pick(probabilities [V]) -> token` (torch.multinomial in the reference; argmax in the deterministic tests).
    Returns the generated tokens WITHOUT the text prefix the reference prepends."""
    assert embeds.shape[0] == 1
    wte = w["transformer.wte.weight"]
    hist = None if text_prefix_tokens is None else text_prefix_tokens.reshape(-1).clone()
    if text_prefix_tokens is not None:
        embeds = torch.cat((embeds, wte[text_prefix_tokens.reshape(1, -1)]), dim=1)
    out: List[int] = []
    for _ in range(entry_length):
        raw = gpt2_logits(w, embeds, cfg, last_only=True)[:, -1, :]
        if mode == "nucleus":
            probs = nucleus_distribution(raw, kw.get("top_p", 0.8), kw.get("top_k", 0), kw.get("temperature", 1.0))[0]
        else:
            probs = no_beam_distribution(raw[0], hist, kw.get("top_p", 0.9), kw.get("top_k", 0.0),
                                         kw.get("temperature", 1.0), kw.get("repetition_penalty", 1.2), stop_token,
                                         kw.get("desired_sentence_length", 50), kw.get("sentence_length_factor", 1.0))
        tok = int(pick(probs))
        if mode == "sample" and tok == stop_token:                    # no_beam.py:67-68: stop token not appended
            break
        out.append(tok)
        t = torch.tensor([tok])
        hist = t if hist is None else torch.cat((hist, t))
        embeds = torch.cat((embeds, wte[t].view(1, 1, -1)), dim=1)
        if mode == "nucleus" and tok == stop_token:                   # nucleus_sampling.py:67-68: appended, then stop
            break
    return out


# ------------------------------------------------------------------------------------------------ training step
def training_loss(map_w: Weights, lm_w: Weights, mcfg: MapperCfg, gcfg: Gpt2Cfg, tokens: torch.Tensor,
                  emb: torch.Tensor, relu_mask_fp16: bool = False) -> torch.Tensor:
    """ClipCapModel.training_step (clipcap/model/model.py:94-113) with forward (model.py:43-58): tokens [B, Tt] int64 with
    -1 padding, emb [B, E]. Differentiable in map_w (torch autograd) — the gradient oracle of cc_train_step.
    The padding mask (model.py:52-56) is not applied: with trailing padding and a causal LM no scored position can see a
    padded key, and padded positions are dropped by ignore_index=0 (pinned against the reference, which does apply it)."""
    tokens = tokens.clone()
    mask = tokens.ge(0)                                              # :103
    tokens[~mask] = 0                                                # :104
    token_embeddings = lm_w["transformer.wte.weight"][tokens]        # :45
    prefix = mapper_forward(map_w, emb, mcfg, relu_mask_fp16)        # :46
    inputs = torch.cat((prefix, token_embeddings), dim=1)            # :49
    logits = gpt2_logits(lm_w, inputs, gcfg)                         # :56
    logits = logits[:, mcfg.K - 1:-1]                                # :109
    return F.cross_entropy(logits.reshape(-1, logits.shape[-1]), tokens.flatten(), ignore_index=0)  # :110


def training_loss_and_grads(map_w: Weights, lm_w: Weights, mcfg: MapperCfg, gcfg: Gpt2Cfg, tokens: torch.Tensor,
                            emb: torch.Tensor, relu_mask_fp16: bool = False) -> Tuple[float, Dict[str, torch.Tensor]]:
    """loss.backward() of the step above for ClipCapModelPrefixOnly (model.py:116-123): gradients of every mapper tensor."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in map_w.items()}
    loss = training_loss(leaves, lm_w, mcfg, gcfg, tokens, emb, relu_mask_fp16)
    loss.backward()
    return float(loss.detach()), {k: v.grad.detach() for k, v in leaves.items()}


def adamw_reference(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step):
    """torch.optim.AdamW single-tensor update (torch/optim/adamw.py _single_tensor_adamw, amsgrad=False)."""
    p = p * (1 - lr * weight_decay)
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    return p - (lr / bc1) * (m / denom), m, v


# ------------------------------------------------------------------------------------------------ whole path
def caption_greedy(vit_w: Weights, map_w: Weights, lm_w: Weights, vcfg: VitCfg, mcfg: MapperCfg, gcfg: Gpt2Cfg,
                   pixels: torch.Tensor, entry_length: int, stop_token: int, normalize: bool = False):
    """docs/inference.md:14-34 call sequence with greedy decode, one image at a time like the reference."""
    emb = vit_encode(vit_w, pixels, vcfg, normalize)
    prefix = mapper_forward(map_w, emb, mcfg)
    return emb, prefix, generate_greedy_batch(lm_w, gcfg, prefix, entry_length, stop_token)
