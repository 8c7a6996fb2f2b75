"""Prefix mappers with the reference's constructor signatures and parameter names (clipcap/model/mapper.py:113-160,
clipcap/model/attention.py:4-15), forward executed by libclipcap_b200 (cc_mapper_forward). The torch.nn layers below are
parameter containers only — they give the same default initialisation, in the same RNG order, as the reference — and
are never called."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from clipcap_b200.engine import MapperEngine
from clipcap_b200.model._lazy import EngineModule


class _Attention(nn.Module):  # attention.py:5-15 (bias=False as TransformerLayer builds it, mapper.py:97)
    def __init__(self, dim_self, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.to_queries = nn.Linear(dim_self, dim_self, bias=False)
        self.to_keys_values = nn.Linear(dim_self, dim_self * 2, bias=False)
        self.project = nn.Linear(dim_self, dim_self)


class _MLP(nn.Module):  # mapper.py:70-80
    def __init__(self, in_dim, h_dim):
        super().__init__()
        self.fc1 = nn.Linear(in_dim, h_dim)
        self.fc2 = nn.Linear(h_dim, in_dim)


class _Layer(nn.Module):  # mapper.py:91-99
    def __init__(self, dim_self, num_heads, mlp_ratio):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim_self)
        self.attn = _Attention(dim_self, num_heads)
        self.norm2 = nn.LayerNorm(dim_self)
        self.mlp = _MLP(dim_self, int(dim_self * mlp_ratio))


class _Transformer(nn.Module):  # mapper.py:8-42, enc_dec=False, mlp_ratio=2.
    def __init__(self, dim_self, num_heads, num_layers):
        super().__init__()
        self.layers = nn.ModuleList([_Layer(dim_self, num_heads, 2.0) for _ in range(num_layers)])


class TransformerMapper(EngineModule):
    kind = "transformer"

    def __init__(self, encoder_embedding_size: int, lm_embedding_size: int, prefix_length: int, projection_length: int,
                 num_heads: int = 8, num_layers: int = 8):
        super().__init__()
        if lm_embedding_size % num_heads != 0 or (lm_embedding_size // num_heads) not in (48, 64, 96, 128):
            raise ValueError(f"clipcap_b200 mapper kernels support head dims 48/64/96/128; got lm_embedding_size="
                             f"{lm_embedding_size}, num_heads={num_heads}")
        self.encoder_embedding_size = encoder_embedding_size
        self.lm_embedding_size = lm_embedding_size
        self.prefix_length = prefix_length
        self.projection_length = projection_length
        self.num_heads, self.num_layers = num_heads, num_layers
        self.window_size, self.use_pos = 1, False
        self.transformer = _Transformer(lm_embedding_size, num_heads, num_layers)
        self.linear = nn.Linear(encoder_embedding_size, projection_length * lm_embedding_size)
        self.prefix_const = nn.Parameter(torch.randn(prefix_length, lm_embedding_size), requires_grad=True)

    def _build_engine(self, weights, capacity, device):
        return MapperEngine(weights, kind=self.kind, E=self.encoder_embedding_size, d=self.lm_embedding_size,
                            P=self.projection_length, K=self.prefix_length, H=self.num_heads, L=self.num_layers,
                            W=self.window_size, use_pos=self.use_pos, max_batch=capacity[0], device=device)

    @torch.no_grad()
    def forward(self, x: torch.Tensor, out: Optional[torch.Tensor] = None,
                out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        """mapper.py:122-130. `out` / `out_dtype` (extensions): write the prefix into a caller-owned [B, K, d] buffer —
        the rank's slot of the all-gathered prefix tensor — or return it in another dtype than the embeddings'."""
        return self._get_engine((max(8, x.shape[0]),)).forward(x, out_dtype=out_dtype, out=out)


class TransformerMapperWindowed(TransformerMapper):
    kind = "windowed"

    def __init__(self, encoder_embedding_size: int, lm_embedding_size: int, prefix_length: int, projection_length: int,
                 window_size: int, use_pos_embeddings: bool, num_heads: int = 8, num_layers: int = 8):
        super().__init__(encoder_embedding_size, lm_embedding_size, prefix_length, projection_length, num_heads,
                         num_layers)
        self.window_size = window_size
        self.use_pos = bool(use_pos_embeddings)
        if use_pos_embeddings:
            self.pos_embeddings = nn.Parameter(torch.randn(window_size * projection_length, lm_embedding_size),
                                               requires_grad=True)
        else:
            self.pos_embeddings = None


class MLPMapper(EngineModule):
    """Upstream-defined MLP mapper (rmokady/CLIP_prefix_caption; absent from the reference):
    Linear(E, K*d/2) -> Tanh -> Linear(K*d/2, K*d) -> view [B, K, d]."""

    def __init__(self, encoder_embedding_size: int, lm_embedding_size: int, prefix_length: int):
        super().__init__()
        self.encoder_embedding_size, self.lm_embedding_size = encoder_embedding_size, lm_embedding_size
        self.prefix_length = prefix_length
        hid = prefix_length * lm_embedding_size // 2
        self.model = nn.Sequential(nn.Linear(encoder_embedding_size, hid), nn.Tanh(),
                                   nn.Linear(hid, prefix_length * lm_embedding_size))

    def _build_engine(self, weights, capacity, device):
        return MapperEngine(weights, kind="mlp", E=self.encoder_embedding_size, d=self.lm_embedding_size,
                            K=self.prefix_length, P=1, H=1, L=1, max_batch=capacity[0], device=device)

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self._get_engine((max(8, x.shape[0]),)).forward(x)
