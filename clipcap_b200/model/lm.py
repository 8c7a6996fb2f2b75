"""GPT-2 language-model facade: the object the reference reaches as `model.language_model` (clipcap/model/model.py:19),
used by the decode loops as `language_model(inputs_embeds=...).logits` and `language_model.get_input_embeddings()`
(clipcap/inference/base.py:76,81,117). Parameters carry the HF GPT2LMHeadModel names so reference checkpoints load; the
arithmetic runs in libclipcap_b200 (cc_gpt2_logits / cc_gpt2_embed / cc_generate)."""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Optional

import torch
import torch.nn as nn

from clipcap_b200.engine import Gpt2Engine
from clipcap_b200.model._lazy import EngineModule

GPT2_SIZES = {  # name -> (n_embd, n_layer, n_head); vocab 50257, n_positions 1024
    "gpt2": (768, 12, 12), "gpt2-medium": (1024, 24, 16), "gpt2-large": (1280, 36, 20), "gpt2-xl": (1600, 48, 25),
}


def lm_dims(name: str):
    """(n_embd, n_layer, n_head, vocab, n_positions). 'tiny:<n_embd>:<n_layer>:<n_head>:<vocab>:<n_pos>' builds small
    test models without any download."""
    if name.startswith("tiny:"):
        d, L, H, V, P = [int(x) for x in name.split(":")[1:]]
        return d, L, H, V, P
    base = name.split("/")[-1]
    if base not in GPT2_SIZES:
        raise ValueError(f"clipcap_b200 implements the GPT-2 family ({', '.join(GPT2_SIZES)}); got '{name}'")
    d, L, H = GPT2_SIZES[base]
    return d, L, H, 50257, 1024


class _Conv1D(nn.Module):  # HF Conv1D: weight [in, out]
    def __init__(self, nf, nx):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(nx, nf))
        self.bias = nn.Parameter(torch.zeros(nf))


class _Attn(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.c_attn = _Conv1D(3 * d, d)
        self.c_proj = _Conv1D(d, d)


class _Mlp(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.c_fc = _Conv1D(4 * d, d)
        self.c_proj = _Conv1D(d, 4 * d)


class _Block(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.ln_1 = nn.LayerNorm(d)
        self.attn = _Attn(d)
        self.ln_2 = nn.LayerNorm(d)
        self.mlp = _Mlp(d)


class TokenEmbedding(nn.Module):
    """`get_input_embeddings()` of the reference call sites; lookup runs in cc_gpt2_embed."""

    def __init__(self, owner: "GPT2LM", V: int, d: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(V, d))
        object.__setattr__(self, "_owner", owner)

    @torch.no_grad()
    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        return self._owner._engine_for(1, 1).embed(ids)


class _Transformer(nn.Module):
    def __init__(self, owner, d, L, V, n_pos):
        super().__init__()
        self.wte = TokenEmbedding(owner, V, d)
        self.wpe = nn.Embedding(n_pos, d)
        self.h = nn.ModuleList([_Block(d) for _ in range(L)])
        self.ln_f = nn.LayerNorm(d)


class GPT2LM(EngineModule):
    def __init__(self, name_or_dims="gpt2"):
        super().__init__()
        d, L, H, V, n_pos = lm_dims(name_or_dims) if isinstance(name_or_dims, str) else name_or_dims
        if d % H != 0 or d // H != 64:
            raise ValueError(f"clipcap_b200 GPT-2 kernels need head dim 64 (n_embd {d} / n_head {H})")
        self.n_embd, self.n_layer, self.n_head, self.vocab_size, self.n_positions = d, L, H, V, n_pos
        self.config = SimpleNamespace(n_embd=d, n_layer=L, n_head=H, vocab_size=V, n_positions=n_pos)
        self.transformer = _Transformer(self, d, L, V, n_pos)
        self._init_weights()

    def _init_weights(self):  # HF GPT2PreTrainedModel._init_weights
        with torch.no_grad():
            for n, p in self.named_parameters():
                if n.endswith("bias"):
                    p.zero_()
                elif "ln_" in n:
                    p.fill_(1.0)
                elif n.endswith("c_proj.weight"):
                    p.normal_(0.0, 0.02 / math.sqrt(2 * self.n_layer))
                else:
                    p.normal_(0.0, 0.02)

    # HF checkpoints also carry the tied `lm_head.weight`; accept and ignore it.
    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        state_dict.pop(prefix + "lm_head.weight", None)
        for k in [k for k in state_dict if k.startswith(prefix) and (k.endswith(".attn.bias") or k.endswith(".attn.masked_bias"))]:
            state_dict.pop(k)
        return super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def get_input_embeddings(self):
        return self.transformer.wte

    def _engine_weights(self):
        return {k: v for k, v in self.state_dict().items()}

    def _build_engine(self, weights, capacity, device):
        return Gpt2Engine(weights, self.n_embd, self.n_layer, self.n_head, self.vocab_size, self.n_positions,
                          max_seqs=capacity[0], max_len=capacity[1], device=device)

    def _engine_for(self, seqs: int, length: int) -> Gpt2Engine:
        if length > self.n_positions:
            raise ValueError(f"sequence length {length} exceeds n_positions {self.n_positions}")
        return self._get_engine((max(8, seqs), min(self.n_positions, max(64, length))))

    def decode_engines(self, n: int, seqs: int, length: int):
        """n independent engines over the same weights (each with its own KV cache and workspaces): the serving pipeline
        prefills batch i+1 in one while batch i still decodes in the other. Engine 0 is the module's own."""
        first = self._engine_for(seqs, length)
        key = (self._engine_key, self._engine_cap)
        extra = getattr(self, "_extra_engines", None)
        if extra is None or extra[0] != key or len(extra[1]) < n - 1:
            if extra is not None:
                for e in extra[1]:
                    e.close()
            built = [self._build_engine(self._engine_weights(), self._engine_cap, self._device()) for _ in range(n - 1)]
            object.__setattr__(self, "_extra_engines", (key, built))
        return [first] + list(self._extra_engines[1][:n - 1])

    @torch.no_grad()
    def forward(self, inputs_embeds: Optional[torch.Tensor] = None, attention_mask: Optional[torch.Tensor] = None,
                input_ids: Optional[torch.Tensor] = None, **unused):
        if inputs_embeds is None:
            if input_ids is None:
                raise ValueError("pass inputs_embeds or input_ids")
            inputs_embeds = self.transformer.wte(input_ids)
        if attention_mask is not None and attention_mask.shape[1] > 1:
            # Trailing padding only (what the reference's dataloader produces): under the causal mask a real position
            # never attends to a padded one, so the logits of real positions equal HF's masked forward; the logits at
            # padded positions are not meaningful (the training loss ignores them, model.py:109-110).
            m = attention_mask.to(torch.bool)
            if bool((m[:, 1:] & ~m[:, :-1]).any()):
                raise NotImplementedError("attention masks with padding before real tokens are not supported")
        B, T, _ = inputs_embeds.shape
        logits = self._engine_for(B, T).logits(inputs_embeds, all_positions=True)
        return SimpleNamespace(logits=logits)

    @torch.no_grad()
    def last_logits(self, inputs_embeds: torch.Tensor) -> torch.Tensor:
        """logits[:, -1] without materialising the other positions (what base.py:83 consumes)."""
        B, T, _ = inputs_embeds.shape
        return self._engine_for(B, T).logits(inputs_embeds, all_positions=False)

    @torch.no_grad()
    def generate_tokens(self, prefix: torch.Tensor, mode="greedy", beam=1, entry_length=67, temperature=1.0,
                        stop_token=50256, **sampling):
        B, Tp, _ = prefix.shape
        eng = self._engine_for(B * (beam if mode == "beam" else 1), Tp + entry_length)
        return eng.generate(prefix, mode, beam, entry_length, temperature, stop_token, **sampling)
