"""Shared plumbing of the nn.Module facades: parameters live in ordinary torch Parameters under the reference's names
(so load_state_dict / .to(device) / .eval() behave as they do for the reference modules) and the native engine is
(re)built lazily whenever they change."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn


class EngineModule(nn.Module):
    def __init__(self):
        super().__init__()
        self._engine = None
        self._engine_key = None
        self._engine_cap = ()

    # --- to be provided by subclasses
    def _engine_weights(self) -> Dict[str, torch.Tensor]:
        return {k: v for k, v in self.state_dict().items()}

    def _build_engine(self, weights, capacity, device):
        raise NotImplementedError

    # --- helpers
    def _param_key(self):
        return tuple((p.data_ptr(), p._version, str(p.device)) for p in self.parameters())

    def _device(self) -> torch.device:
        return next(self.parameters()).device

    def _get_engine(self, capacity):
        """`capacity` is a tuple of ints the engine must be able to hold (batch, lengths...)."""
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError(
                f"clipcap_b200: module is on {dev}; move it to a B200 (`.to('cuda')`) — there is no CPU fallback path")
        key = self._param_key()
        eng = self._engine
        if eng is not None and key == self._engine_key and all(c <= m for c, m in zip(capacity, self._engine_cap)):
            return eng
        if eng is not None:
            if key == self._engine_key:  # same weights, larger request: grow, never shrink
                capacity = tuple(max(c, m) for c, m in zip(capacity, self._engine_cap))
            eng.close()
            self._engine = None
        self._engine = self._build_engine(self._engine_weights(), capacity, dev)
        self._engine_key = key
        self._engine_cap = capacity
        return self._engine
