"""Seeded synthetic weights and inputs shared by the oracle, the tests and bench.py (there are no pretrained weights
and no network in this environment). TEST / BENCH INFRASTRUCTURE — the product path only ever sees the
resulting state_dict tensors, exactly as it would see a real checkpoint.

Keys follow the reference's state_dict names: `visual.*` (OpenAI clip), mapper keys relative to `transformer_mapper.`
(clipcap/model/mapper.py), GPT-2 keys relative to `language_model.` (HF GPT2LMHeadModel). Scales follow the
libraries' default initialisers (HF GPT-2 / CLIP: N(0, 0.02)-style; torch.nn.Linear: U(+-1/sqrt(fan_in));
prefix_const ~ N(0,1), mapper.py:120), with LayerNorm affine parameters perturbed so a wrong gamma/beta is visible.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from .restate import Gpt2Cfg, MapperCfg, VitCfg


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def _n(g, *shape, std=0.02):
    return torch.randn(*shape, generator=g) * std


def _u(g, *shape, bound):
    return (torch.rand(*shape, generator=g) * 2 - 1) * bound


def _ln(g, d):
    return 1.0 + _n(g, d, std=0.1), _n(g, d, std=0.05)


def vit_weights(cfg: VitCfg, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = _gen(seed)
    w, T = cfg.width, (cfg.image_size // cfg.patch) ** 2 + 1
    sd = {
        "visual.conv1.weight": _n(g, w, 3, cfg.patch, cfg.patch, std=0.02),
        "visual.class_embedding": _n(g, w, std=w ** -0.5),
        "visual.positional_embedding": _n(g, T, w, std=w ** -0.5),
        "visual.proj": _n(g, w, cfg.out_dim, std=w ** -0.5),
    }
    sd["visual.ln_pre.weight"], sd["visual.ln_pre.bias"] = _ln(g, w)
    sd["visual.ln_post.weight"], sd["visual.ln_post.bias"] = _ln(g, w)
    attn_std = w ** -0.5
    proj_std = (w ** -0.5) * ((2 * cfg.layers) ** -0.5)
    fc_std = (2 * w) ** -0.5
    for l in range(cfg.layers):
        p = f"visual.transformer.resblocks.{l}."
        sd[p + "ln_1.weight"], sd[p + "ln_1.bias"] = _ln(g, w)
        sd[p + "ln_2.weight"], sd[p + "ln_2.bias"] = _ln(g, w)
        sd[p + "attn.in_proj_weight"] = _n(g, 3 * w, w, std=attn_std)
        sd[p + "attn.in_proj_bias"] = _n(g, 3 * w, std=0.02)
        sd[p + "attn.out_proj.weight"] = _n(g, w, w, std=proj_std)
        sd[p + "attn.out_proj.bias"] = _n(g, w, std=0.02)
        sd[p + "mlp.c_fc.weight"] = _n(g, cfg.mlp_dim, w, std=fc_std)
        sd[p + "mlp.c_fc.bias"] = _n(g, cfg.mlp_dim, std=0.02)
        sd[p + "mlp.c_proj.weight"] = _n(g, w, cfg.mlp_dim, std=proj_std)
        sd[p + "mlp.c_proj.bias"] = _n(g, w, std=0.02)
    return sd


def mapper_weights(cfg: MapperCfg, seed: int = 1) -> Dict[str, torch.Tensor]:
    g = _gen(seed)
    d = cfg.d
    if cfg.kind == "mlp":
        hid, out = cfg.K * d // 2, cfg.K * d
        return {
            "model.0.weight": _u(g, hid, cfg.E, bound=cfg.E ** -0.5), "model.0.bias": _u(g, hid, bound=cfg.E ** -0.5),
            "model.2.weight": _u(g, out, hid, bound=hid ** -0.5), "model.2.bias": _u(g, out, bound=hid ** -0.5),
        }
    W = cfg.W if cfg.kind == "windowed" else 1
    sd = {
        "linear.weight": _u(g, cfg.P * d, cfg.E, bound=cfg.E ** -0.5),
        "linear.bias": _u(g, cfg.P * d, bound=cfg.E ** -0.5),
        "prefix_const": _n(g, cfg.K, d, std=1.0),
    }
    if cfg.kind == "windowed" and cfg.use_pos:
        sd["pos_embeddings"] = _n(g, W * cfg.P, d, std=1.0)
    for l in range(cfg.L):
        p = f"transformer.layers.{l}."
        sd[p + "norm1.weight"], sd[p + "norm1.bias"] = _ln(g, d)
        sd[p + "norm2.weight"], sd[p + "norm2.bias"] = _ln(g, d)
        sd[p + "attn.to_queries.weight"] = _u(g, d, d, bound=d ** -0.5)
        sd[p + "attn.to_keys_values.weight"] = _u(g, 2 * d, d, bound=d ** -0.5)
        sd[p + "attn.project.weight"] = _u(g, d, d, bound=d ** -0.5)
        sd[p + "attn.project.bias"] = _u(g, d, bound=d ** -0.5)
        sd[p + "mlp.fc1.weight"] = _u(g, 2 * d, d, bound=d ** -0.5)
        sd[p + "mlp.fc1.bias"] = _u(g, 2 * d, bound=d ** -0.5)
        sd[p + "mlp.fc2.weight"] = _u(g, d, 2 * d, bound=(2 * d) ** -0.5)
        sd[p + "mlp.fc2.bias"] = _u(g, d, bound=(2 * d) ** -0.5)
    return sd


def gpt2_weights(cfg: Gpt2Cfg, seed: int = 2, wte_std: float = 0.02) -> Dict[str, torch.Tensor]:
    """HF GPT-2 init: N(0, 0.02) everywhere, residual projections scaled by 1/sqrt(2L). `wte_std` can be raised to
    sharpen the logit distribution (larger top-1/top-2 margins) for the bit-exact token tests."""
    g = _gen(seed)
    d = cfg.d
    sd = {"transformer.wte.weight": _n(g, cfg.V, d, std=wte_std), "transformer.wpe.weight": _n(g, cfg.n_pos, d, std=0.02)}
    sd["transformer.ln_f.weight"], sd["transformer.ln_f.bias"] = _ln(g, d)
    rs = 0.02 / math.sqrt(2 * cfg.L)
    for l in range(cfg.L):
        p = f"transformer.h.{l}."
        sd[p + "ln_1.weight"], sd[p + "ln_1.bias"] = _ln(g, d)
        sd[p + "ln_2.weight"], sd[p + "ln_2.bias"] = _ln(g, d)
        sd[p + "attn.c_attn.weight"] = _n(g, d, 3 * d, std=0.02)
        sd[p + "attn.c_attn.bias"] = _n(g, 3 * d, std=0.02)
        sd[p + "attn.c_proj.weight"] = _n(g, d, d, std=rs)
        sd[p + "attn.c_proj.bias"] = _n(g, d, std=0.02)
        sd[p + "mlp.c_fc.weight"] = _n(g, d, 4 * d, std=0.02)
        sd[p + "mlp.c_fc.bias"] = _n(g, 4 * d, std=0.02)
        sd[p + "mlp.c_proj.weight"] = _n(g, 4 * d, d, std=rs)
        sd[p + "mlp.c_proj.bias"] = _n(g, d, std=0.02)
    return sd


def pixels(B: int, image_size: int = 224, seed: int = 1234) -> torch.Tensor:
    """CLIP-normalised images are ~unit scale."""
    return torch.randn(B, 3, image_size, image_size, generator=_gen(seed))


def embeddings(B: int, E: int, seed: int = 4321) -> torch.Tensor:
    return torch.randn(B, E, generator=_gen(seed))


def checksum(sd: Dict[str, torch.Tensor]) -> float:
    """Order-independent fingerprint used by the golden fixtures to detect RNG drift between torch builds."""
    return float(sum(t.double().abs().sum().item() for t in sd.values()))
