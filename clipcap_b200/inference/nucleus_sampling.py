"""generate_nucleus_sampling with the reference's signature (clipcap/inference/nucleus_sampling.py:9-75).

The reference re-runs the language model over the growing sequence and, per step, does softmax -> topk -> cumsum ->
searchsorted(top_p) -> renormalise -> torch.multinomial on the host-visible tensors. Here the prefix is prefilled once,
every step is a KV-cached decode pass, and the token selection is one kernel (csrc/sample.cu, mode CC_GEN_NUCLEUS): the
kept set and the distribution are the reference's, the random stream is Philox keyed by (seed, row, step) — reproducible
per seed, not bit-identical to torch.multinomial. `top_k=1` makes the draw deterministic (tested against the reference).

Extensions: `embeds` may hold B > 1 images (one caption each); `seed` (default: drawn from torch's global generator, so
`torch.manual_seed` controls it).
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch

from clipcap_b200.inference.base import _with_text_prefix


def _draw_seed(seed: Optional[int]) -> int:
    if seed is not None:
        return int(seed)
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


def _text_prefix_list(text_prefix_tokens) -> List[int]:
    if text_prefix_tokens is None:
        return []
    return [int(t) for t in text_prefix_tokens.reshape(-1).tolist()]


def generate_nucleus_sampling_tokens(model, embeds: torch.Tensor, text_prefix_tokens: Optional[torch.Tensor] = None,
                                     entry_length: int = 67, top_p: Optional[float] = 0.8, top_k: int = 0,
                                     temperature: float = 1.0, stop_token: int = 13, seed: Optional[int] = None):
    """Device-side result: (tokens int32 [B, entry_length], lengths int32 [B], scores (unused))."""
    embeds = _with_text_prefix(model, embeds, text_prefix_tokens)
    return model.language_model.generate_tokens(embeds, mode="nucleus", entry_length=entry_length,
                                                temperature=temperature, stop_token=stop_token,
                                                top_p=1.0 if top_p is None else top_p, top_k=top_k,
                                                seed=_draw_seed(seed))


def generate_nucleus_sampling(model, tokenizer: Callable, embeds: torch.Tensor, number_to_generate: int = 1,
                              text_prefix_tokens: Optional[torch.Tensor] = None, entry_length: int = 67,
                              top_p: float = 0.8, top_k: int = 0, temperature: float = 1.0,
                              seed: Optional[int] = None) -> List[str]:
    stop_token = tokenizer.encode(".")[0]  # nucleus_sampling.py:21
    head = _text_prefix_list(text_prefix_tokens)  # the reference returns the text prefix in front (:31, :62-65)
    generations: List[str] = []
    for n in range(number_to_generate):
        tokens, lengths, _ = generate_nucleus_sampling_tokens(
            model, embeds, text_prefix_tokens, entry_length, top_p, top_k, temperature, stop_token,
            None if seed is None else seed + n)
        tokens, lengths = tokens.cpu().numpy(), lengths.cpu().numpy()
        for i in range(tokens.shape[0]):
            generations.append(tokenizer.decode(head + [int(t) for t in tokens[i][:int(lengths[i])]]))
    return generations
