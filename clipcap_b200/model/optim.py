"""Optimiser and schedule of the reference's `configure_optimizers` (clipcap/model/model.py:67-91): AdamW (deepspeed
FusedAdam in adam_w_mode when `use_deepspeed_optimisers`, torch.optim.AdamW otherwise) with transformers'
`get_linear_schedule_with_warmup`. The update runs in libclipcap_b200 (cc_op_adamw), one launch per parameter."""
from __future__ import annotations

import torch
from torch.optim import Optimizer
from torch.optim.lr_scheduler import LambdaLR

from clipcap_b200.engine import adamw_update


class FusedAdamW(Optimizer):
    """torch.optim.AdamW semantics (decoupled weight decay, bias correction, no amsgrad) on fp32 CUDA parameters."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01):
        if lr < 0.0 or eps < 0.0 or not (0.0 <= betas[0] < 1.0) or not (0.0 <= betas[1] < 1.0):
            raise ValueError(f"invalid AdamW hyper-parameters lr={lr} betas={betas} eps={eps}")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32:
                    raise RuntimeError("FusedAdamW: fp32 CUDA parameters only (clipcap_b200 has no CPU path)")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                adamw_update(p.data, g, st["exp_avg"], st["exp_avg_sq"], group["lr"], b1, b2, group["eps"],
                             group["weight_decay"], st["step"])
                # The update went through a raw pointer: tell autograd (and EngineModule._param_key, which keys the cached
                # native engines on `_version`) that the parameter changed, as an in-place torch op would have.
                torch.autograd.graph.increment_version(p)
        return loss


def linear_schedule_with_warmup(optimizer: Optimizer, num_warmup_steps: int, num_training_steps: int) -> LambdaLR:
    """transformers.get_linear_schedule_with_warmup (optimization.py): 0 -> lr over the warm-up, then linearly to 0."""

    def lr_lambda(step: int) -> float:
        if step < num_warmup_steps:
            return float(step) / float(max(1, num_warmup_steps))
        return max(0.0, float(num_training_steps - step) / float(max(1, num_training_steps - num_warmup_steps)))

    return LambdaLR(optimizer, lr_lambda)
