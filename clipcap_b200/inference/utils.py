"""Logit filters with the reference's names and semantics (clipcap/inference/utils.py:5-49). The decode loops apply
the same rules inside csrc/sample.cu; these tensor versions exist for callers that import them directly and run on
whatever device the logits live on (they are not on the captioning hot path)."""
from __future__ import annotations

import torch
import torch.nn.functional as nnf


def top_k_top_p_filtering(logits: torch.Tensor, top_k=0, top_p=0.0, filter_value=-float("Inf")):
    assert logits.dim() == 1  # utils.py:14
    top_k = min(int(top_k), logits.size(-1))
    if top_k > 0:
        kth = torch.topk(logits, top_k)[0][..., -1, None]
        logits[logits < kth] = filter_value
    if top_p > 0.0:
        sorted_logits, sorted_indices = torch.sort(logits, descending=True)
        cumulative = torch.cumsum(nnf.softmax(sorted_logits, dim=-1), dim=-1)
        remove = cumulative > top_p
        remove[..., 1:] = remove[..., :-1].clone()  # keep the first token above the threshold
        remove[..., 0] = 0
        logits[sorted_indices[remove]] = filter_value
    return logits


def repetition_penalty_apply(logits: torch.Tensor, tokens: torch.Tensor, penalty: float) -> torch.Tensor:
    tok = torch.gather(logits, -1, tokens)
    tok = torch.where(tok < 0, tok * penalty, tok / penalty)
    logits.scatter_(-1, tokens, tok)
    return logits


def sentence_length_penalty_apply(logits: torch.Tensor, tokens: torch.Tensor, stop_token: int, current_length: int,
                                  desired_length: int, length_factor: float) -> torch.Tensor:
    penalty = (current_length / desired_length) * length_factor
    tok = torch.gather(logits, -1, tokens)
    tok = torch.where(tok == stop_token, tok * penalty, tok)  # compares logit VALUES with the token id (utils.py:46)
    logits.scatter_(-1, tokens, tok)
    return logits
