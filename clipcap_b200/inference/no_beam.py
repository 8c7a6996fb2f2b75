"""generate_no_beam with the reference's signature (clipcap/inference/no_beam.py:10-82): repetition penalty, temperature,
top-k / top-p filtering and the "sentence length penalty" of clipcap/inference/utils.py, then a multinomial draw; the
stop token ends a caption without being returned. Prefill + KV-cached decode + one selection kernel per step
(csrc/sample.cu, mode CC_GEN_SAMPLE); see nucleus_sampling.py for the RNG note."""
from __future__ import annotations

from typing import Callable, List, Optional

import torch

from clipcap_b200.inference.base import _with_text_prefix
from clipcap_b200.inference.nucleus_sampling import _draw_seed, _text_prefix_list


def generate_no_beam_tokens(model, embeds: torch.Tensor, text_prefix_tokens: Optional[torch.Tensor] = None,
                            top_p: float = 0.9, top_k: float = 0.0, entry_length: int = 67, temperature: float = 1.0,
                            repetition_penalty: float = 1.2, desired_sentence_length: int = 50,
                            sentence_length_factor: float = 1.0, stop_token: int = 13, seed: Optional[int] = None):
    """Device-side result: (tokens int32 [B, entry_length], lengths int32 [B], scores (unused))."""
    embeds = _with_text_prefix(model, embeds, text_prefix_tokens)
    return model.language_model.generate_tokens(
        embeds, mode="sample", entry_length=entry_length, temperature=temperature, stop_token=stop_token,
        top_p=top_p, top_k=int(top_k), repetition_penalty=repetition_penalty,
        desired_sentence_length=desired_sentence_length, sentence_length_factor=sentence_length_factor,
        history=_text_prefix_list(text_prefix_tokens), seed=_draw_seed(seed))


def generate_no_beam(model, tokenizer: Callable, embeds: torch.Tensor, number_to_generate: int = 5,
                     text_prefix_tokens: Optional[torch.Tensor] = None, top_p: float = 0.9, top_k: float = 0.0,
                     entry_length: int = 67, temperature: float = 1.0, repetition_penalty: float = 1.2,
                     desired_sentence_length: int = 50, sentence_length_factor: float = 1.0,
                     seed: Optional[int] = None) -> List[str]:
    stop_token = tokenizer.encode(".")[0]  # no_beam.py:24
    head = _text_prefix_list(text_prefix_tokens)  # `tokens` starts as the text prefix (no_beam.py:34)
    generations: List[str] = []
    for n in range(number_to_generate):
        tokens, lengths, _ = generate_no_beam_tokens(
            model, embeds, text_prefix_tokens, top_p, top_k, entry_length, temperature, repetition_penalty,
            desired_sentence_length, sentence_length_factor, stop_token, None if seed is None else seed + n)
        tokens, lengths = tokens.cpu().numpy(), lengths.cpu().numpy()
        for i in range(tokens.shape[0]):
            generations.append(tokenizer.decode(head + [int(t) for t in tokens[i][:int(lengths[i])]]))
    return generations
