"""The three logit filters of clipcap/inference/utils.py:5-49 under their reference names. The decode loops apply these
rules on the device inside csrc/sample.cu; the tensor versions here serve callers that import them directly. Like the
reference they edit `logits` in place and return it, and they keep its quirks (the length penalty compares logit values
with the stop token's id, utils.py:46)."""
from __future__ import annotations

import torch


def top_k_top_p_filtering(logits: torch.Tensor, top_k=0, top_p=0.0, filter_value=-float("Inf")):
    """Keep the top_k largest logits (0 = all), then the smallest head of the sorted distribution whose mass exceeds top_p
    (0 = off; the first token past the threshold stays). 1-D logits only."""
    assert logits.dim() == 1
    vocab = logits.size(-1)
    k = min(int(top_k), vocab)
    if k > 0:
        threshold = logits.topk(k).values[-1]
        logits.masked_fill_(logits < threshold, filter_value)
    if top_p > 0.0:
        order = logits.argsort(descending=True)
        mass = logits[order].softmax(-1).cumsum(-1)
        drop = torch.zeros(vocab, dtype=torch.bool, device=logits.device)
        drop[1:] = mass[:-1] > top_p      # shifted by one: the token that crosses top_p is kept
        logits[order[drop]] = filter_value
    return logits


def _edit_at(logits: torch.Tensor, tokens: torch.Tensor, fn) -> torch.Tensor:
    logits.scatter_(-1, tokens, fn(logits.gather(-1, tokens)))
    return logits


def repetition_penalty_apply(logits: torch.Tensor, tokens: torch.Tensor, penalty: float) -> torch.Tensor:
    """Tokens already generated become less likely: negative logits are multiplied by `penalty`, positive ones divided."""
    return _edit_at(logits, tokens, lambda v: torch.where(v < 0, v * penalty, v / penalty))


def sentence_length_penalty_apply(logits: torch.Tensor, tokens: torch.Tensor, stop_token: int, current_length: int,
                                  desired_length: int, length_factor: float) -> torch.Tensor:
    scale = (current_length / desired_length) * length_factor
    return _edit_at(logits, tokens, lambda v: torch.where(v == stop_token, v * scale, v))
