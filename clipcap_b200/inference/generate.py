"""`generate` convenience wrapper with the reference's signature and call sequence (clipcap/inference/generate.py:8-44):
BOS (+ text prefix) tokens -> their embeddings, mapper prefix, concatenation, then generate_no_beam with the same
text_prefix_tokens (which, as in the reference, embeds and appends the text prefix a second time: generate.py:30-41 feeding
no_beam.py:27-29)."""
from __future__ import annotations

from typing import Callable, List, Optional

import torch

from clipcap_b200.inference.no_beam import generate_no_beam


def generate(model, tokenizer: Callable, embeddings: torch.Tensor, top_p: float = 0.95, top_k: int = 0,
             temperature: float = 1.0, number_to_generate: int = 5, text_prefix: Optional[str] = None,
             stop_token: Optional[str] = None, seed: Optional[int] = None) -> List[str]:
    batch_size = embeddings.shape[0]
    assert batch_size == 1, "Batch size > 1 support coming soon - for now leave embeddings.shape[0] as 1."  # generate.py:20
    text_prefix = tokenizer.bos_token + text_prefix if text_prefix is not None else tokenizer.bos_token
    text_prefix_tokens = tokenizer.encode(text_prefix, return_tensors="pt").expand(batch_size, -1).to(embeddings.device)
    with torch.no_grad():
        token_embeddings = model.language_model.get_input_embeddings()(text_prefix_tokens)
        prefix_projections = model.transformer_mapper(embeddings)
    inputs_embeds = torch.cat((prefix_projections, token_embeddings.to(prefix_projections.dtype)), dim=1)
    return generate_no_beam(model, tokenizer, inputs_embeds, number_to_generate=number_to_generate,
                            text_prefix_tokens=text_prefix_tokens, top_p=top_p, top_k=top_k, temperature=temperature,
                            seed=seed)
