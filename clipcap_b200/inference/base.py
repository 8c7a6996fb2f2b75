"""Decode entry points with the reference's signatures (clipcap/inference/base.py:55-64, nucleus_sampling.py:9-19,
no_beam.py:10-23). The whole loop — prefill, KV-cached decode steps, log-softmax/top-k, beam bookkeeping — runs on the
device inside one cc_generate call (a replayed CUDA graph); the host only decodes the returned token ids
(base.py:124-130).

Extension over the reference: `embeds` may hold B > 1 images ([B, K, d]); the result is then one caption per image, each
equal to what the reference returns when called on that image alone (the reference itself is batch-size-1:
generate.py:19-20, base.py:88).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch


def _stop_token(tokenizer, text) -> int:
    return tokenizer.encode(text)[0]


def _with_text_prefix(model, embeds, text_prefix_tokens):
    if text_prefix_tokens is None:
        return embeds
    text_prefix_embed = model.language_model.get_input_embeddings()(text_prefix_tokens)  # base.py:75-77
    if text_prefix_embed.dim() == 2:
        text_prefix_embed = text_prefix_embed.unsqueeze(0)
    text_prefix_embed = text_prefix_embed.expand(embeds.shape[0], -1, -1)
    return torch.cat((embeds, text_prefix_embed.to(embeds.dtype)), dim=1)


def generate_beam_tokens(model, embeds: torch.Tensor, text_prefix_tokens: Optional[torch.Tensor] = None,
                         beam_size: int = 5, entry_length: int = 67, temperature: float = 1.0,
                         stop_token: int = 50256) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Device-side result of generate_beam: (tokens int32 [B, entry_length], lengths int32 [B], scores fp32 [B]) for the
    best beam of every image; no host synchronisation."""
    embeds = _with_text_prefix(model, embeds, text_prefix_tokens)
    return model.language_model.generate_tokens(embeds, mode="beam", beam=beam_size, entry_length=entry_length,
                                                temperature=temperature, stop_token=stop_token)


def generate_greedy_tokens(model, embeds: torch.Tensor, entry_length: int = 67, stop_token: int = 50256):
    """generate_beam(beam_size=1) — the reference's greedy decode — through the fused-argmax LM head
    (no logits are materialised). Scores are not computed in this mode (returned as zeros)."""
    return model.language_model.generate_tokens(embeds, mode="greedy", beam=1, entry_length=entry_length,
                                                stop_token=stop_token)


def _decode(tokenizer, tokens, lengths) -> List[str]:
    tokens, lengths = tokens.cpu().numpy(), lengths.cpu().numpy()  # base.py:124
    return [tokenizer.decode(tokens[i][:int(lengths[i])]) for i in range(tokens.shape[0])]


def generate_beam(model, tokenizer: Callable, embeds: torch.Tensor, number_to_generate: int = 1,
                  text_prefix_tokens: Optional[torch.Tensor] = None, beam_size: int = 5, entry_length: int = 67,
                  temperature: float = 1.0) -> List[str]:
    stop_token = _stop_token(tokenizer, tokenizer.eos_token)  # base.py:66
    generations: List[str] = []
    for _ in range(number_to_generate):
        # Beam search is deterministic; the reference's second pass re-uses stale state and is
        # only meaningful for number_to_generate == 1.
        tokens, lengths, _scores = generate_beam_tokens(model, embeds, text_prefix_tokens, beam_size, entry_length,
                                                        temperature, stop_token)
        generations.extend(_decode(tokenizer, tokens, lengths))
    return generations


# The reference keeps an older copy of the nucleus loop in base.py (base.py:135-201); both names resolve to the device
# implementation in clipcap_b200.inference.nucleus_sampling / no_beam.
def generate_nucleus_sampling(*args, **kwargs):
    from clipcap_b200.inference.nucleus_sampling import generate_nucleus_sampling as impl
    return impl(*args, **kwargs)


def generate_no_beam(*args, **kwargs):
    from clipcap_b200.inference.no_beam import generate_no_beam as impl
    return impl(*args, **kwargs)
