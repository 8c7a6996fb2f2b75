"""CLIP image encoder behind the reference's `CLIPModel` wrapper (clipcap/encoders/clip.py:105-153).

`CLIPModel` keeps the reference's constructor and forward (window flatten / optional normalisation / unflatten,
clip.py:112-129) around any object exposing `encode_image`. `ViTImageTower` is that object here: it holds the OpenAI-clip
named parameters (`visual.*`) and runs libclipcap_b200's cc_vit_forward. The reference's `import clip; clip.load(...)`
(clip.py:134-136) downloads weights; offline the tower is randomly initialised unless `weights_path` (a state_dict with
`visual.*` keys) is given.
"""
from __future__ import annotations

import math
import os
import warnings
from typing import Callable, Optional, Tuple

import torch
import torch.nn as nn

from clipcap_b200.engine import VitEngine
from clipcap_b200.model._lazy import EngineModule

# variant -> (image_size, patch, width, layers, heads, out_dim); mlp_dim = 4 * width
CLIP_VARIANTS = {
    "ViT-B/32": (224, 32, 768, 12, 12, 512),
    "ViT-B/16": (224, 16, 768, 12, 12, 512),
    "ViT-L/14": (224, 14, 1024, 24, 16, 768),
    "ViT-L/14@336px": (336, 14, 1024, 24, 16, 768),
}
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)  # clip.py:23
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


class _MHA(nn.Module):
    def __init__(self, w):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * w, w))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * w))
        self.out_proj = nn.Linear(w, w)


class _Mlp(nn.Module):
    def __init__(self, w, m):
        super().__init__()
        self.c_fc = nn.Linear(w, m)
        self.c_proj = nn.Linear(m, w)


class _ResBlock(nn.Module):
    def __init__(self, w, m):
        super().__init__()
        self.attn = _MHA(w)
        self.ln_1 = nn.LayerNorm(w)
        self.mlp = _Mlp(w, m)
        self.ln_2 = nn.LayerNorm(w)


class _Blocks(nn.Module):
    def __init__(self, w, m, L):
        super().__init__()
        self.resblocks = nn.ModuleList([_ResBlock(w, m) for _ in range(L)])


class _Visual(nn.Module):
    def __init__(self, image_size, patch, w, L, m, out_dim):
        super().__init__()
        scale = w ** -0.5
        self.conv1 = nn.Conv2d(3, w, kernel_size=patch, stride=patch, bias=False)
        self.class_embedding = nn.Parameter(scale * torch.randn(w))
        self.positional_embedding = nn.Parameter(scale * torch.randn((image_size // patch) ** 2 + 1, w))
        self.ln_pre = nn.LayerNorm(w)
        self.transformer = _Blocks(w, m, L)
        self.ln_post = nn.LayerNorm(w)
        self.proj = nn.Parameter(scale * torch.randn(w, out_dim))


class ViTImageTower(EngineModule):
    """The `clip_model` object of the reference: exposes `encode_image(x[B,3,S,S]) -> [B, out_dim]`."""

    def __init__(self, image_size=224, patch=14, width=1024, layers=24, heads=16, out_dim=768, mlp_dim=None):
        super().__init__()
        if width % heads != 0 or width // heads != 64:
            raise ValueError(f"clipcap_b200 ViT kernels need head dim 64 (width {width} / heads {heads})")
        self.image_size, self.patch, self.width, self.layers, self.heads, self.out_dim = (
            image_size, patch, width, layers, heads, out_dim)
        self.mlp_dim = mlp_dim or 4 * width
        self.visual = _Visual(image_size, patch, width, layers, self.mlp_dim, out_dim)
        with torch.no_grad():  # OpenAI clip initialisation (model.py initialize_parameters)
            proj_std = (width ** -0.5) * ((2 * layers) ** -0.5)
            for blk in self.visual.transformer.resblocks:
                nn.init.normal_(blk.attn.in_proj_weight, std=width ** -0.5)
                nn.init.normal_(blk.attn.out_proj.weight, std=proj_std)
                nn.init.normal_(blk.mlp.c_fc.weight, std=(2 * width) ** -0.5)
                nn.init.normal_(blk.mlp.c_proj.weight, std=proj_std)

    def _build_engine(self, weights, capacity, device):
        return VitEngine(weights, self.image_size, self.patch, self.width, self.layers, self.heads, self.mlp_dim,
                         self.out_dim, max_batch=capacity[0], device=device)

    @torch.no_grad()
    def encode_image(self, x: torch.Tensor, normalize: bool = False) -> torch.Tensor:
        return self._get_engine((max(8, x.shape[0]),)).forward(x, normalize=normalize)

    forward = encode_image


class CLIPModel(nn.Module):
    """clipcap/encoders/clip.py:105-129, same constructor and semantics."""

    def __init__(self, model: nn.Module, normalize_embeddings: bool = False, use_windowed_embeddings: bool = False) -> None:
        super().__init__()
        self.model = model
        self.normalize_embeddings = normalize_embeddings
        self.use_windowed_embeddings = use_windowed_embeddings

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        original_shape = x.shape
        if self.use_windowed_embeddings:
            x = torch.flatten(x, start_dim=0, end_dim=1)  # clip.py:116-118
        if isinstance(self.model, ViTImageTower):
            out = self.model.encode_image(x, normalize=self.normalize_embeddings)  # fused L2 normalisation
        else:
            out = self.model.encode_image(x)
            if self.normalize_embeddings:
                out /= out.norm(dim=-1, keepdim=True)  # clip.py:122-123
        if self.use_windowed_embeddings:
            out = out.view(original_shape[0], original_shape[1], *out.shape[1:])  # clip.py:125-127
        return out


class TensorTransform:
    """Stand-in for CLIPTransform (clip.py:9-103) on decoded images: takes a PIL image, a path or a uint8/float tensor
    [3,H,W] and returns the CLIP-normalised [3,S,S] tensor (bicubic resize of the short side, centre crop). With
    `use_windowed_embeddings` it returns [window_size + 1, 3, S, S] — the global view followed by the window tiles in
    row-major order (clip.py:83-103): centre crop to a square (clip.py:35-47), bilinear resize to a multiple of the tile
    grid (clip.py:49-58), window tiling on the device (cc_op_tile_image = the reference's unfold, clip.py:60-82), then the
    same resize + normalisation per tile. File decoding (PIL) stays with the caller's data loader. What the reference's
    own tile path does as committed cannot be reproduced literally: `image.convert("rgb")` (clip.py:70) raises, and its
    Normalize / flatten act on the tile-x axis instead of the channels (clip.py:91,99); this follows the documented
    intent — one CLIP-normalised image per tile."""

    def __init__(self, image_size: int, use_windowed_embeddings: bool = False, window_size: Optional[int] = 9,
                 window_overlap_percentage: float = 0.0, device="cuda"):
        self.image_size = image_size
        self.use_windowed_embeddings = use_windowed_embeddings
        self.window_size = window_size
        self.window_overlap_percentage = window_overlap_percentage
        self.device = device
        if use_windowed_embeddings:  # clip.py:14-15
            assert math.sqrt(window_size).is_integer(), \
                "`window_size` must be a square number with CLIP, e.g. (3x3) = 9 for tiles of 3 by 3."

    @staticmethod
    def _decode(image) -> torch.Tensor:
        if isinstance(image, str):
            from PIL import Image
            image = Image.open(image)
        if not isinstance(image, torch.Tensor):
            import numpy as np
            image = torch.from_numpy(np.asarray(image.convert("RGB"))).permute(2, 0, 1)
        return image.float() / 255.0 if image.dtype == torch.uint8 else image.float()

    def _clip_view(self, x: torch.Tensor) -> torch.Tensor:
        """[..., 3, H, W] in [0, 1] -> CLIP-normalised [..., 3, S, S]."""
        import torch.nn.functional as F
        S = self.image_size
        lead = x.shape[:-3]
        x = x.reshape(-1, *x.shape[-3:])
        H, W = x.shape[-2:]
        s = S / min(H, W)
        nh, nw = max(S, round(H * s)), max(S, round(W * s))
        x = F.interpolate(x, size=(nh, nw), mode="bicubic", align_corners=False, antialias=True)
        t, l = (nh - S) // 2, (nw - S) // 2
        x = x[..., t:t + S, l:l + S].clamp(0, 1)
        mean = torch.tensor(CLIP_MEAN, device=x.device).view(3, 1, 1)
        std = torch.tensor(CLIP_STD, device=x.device).view(3, 1, 1)
        return ((x - mean) / std).reshape(*lead, 3, S, S)

    def tile_image(self, square: torch.Tensor) -> torch.Tensor:
        """clip.py:60-82 on a decoded square image [3, S, S] (CUDA, fp32): -> [window_size, 3, p, p]."""
        import ctypes as C  # noqa: F401
        from clipcap_b200 import _ffi
        if not square.is_cuda:
            raise RuntimeError("clipcap_b200: window tiling runs on the device (there is no CPU path)")
        size = square.shape[-1]
        n = int(math.sqrt(self.window_size))
        p = size // n
        step = math.floor(p * (1 - self.window_overlap_percentage / 100)) if self.window_overlap_percentage != 0 else p
        square = square.float().contiguous()
        tiles = torch.empty(n * n, 3, p, p, device=square.device, dtype=torch.float32)
        with torch.cuda.device(square.device):
            _ffi.check(_ffi.lib().cc_op_tile_image(square.data_ptr(), size, n, p, step, tiles.data_ptr(),
                                                   _ffi.current_stream_ptr()))
        return tiles

    def __call__(self, image) -> torch.Tensor:
        import torch.nn.functional as F
        x = self._decode(image)
        if not self.use_windowed_embeddings:
            return self._clip_view(x)
        x = x.to(self.device)
        _, H, W = x.shape
        side = min(H, W)                                   # centre crop to a square, clip.py:35-47
        t, l = (H - side) // 2, (W - side) // 2
        sq = x[:, t:t + side, l:l + side]
        n = int(math.sqrt(self.window_size))
        target = math.ceil(side / n) * n                   # ensure_tileable, clip.py:49-58
        if target != side:
            sq = F.interpolate(sq[None], size=(target, target), mode="bilinear", align_corners=False)[0]
        tiles = self.tile_image(sq)
        return torch.cat((self._clip_view(x)[None], self._clip_view(tiles)), dim=0)  # clip.py:95-101


def warn_random_init(what: str, env_var: str, allowed: bool) -> None:
    """An encoder without pretrained weights produces meaningless embeddings: never let that pass silently."""
    if allowed or os.environ.get("CLIPCAP_B200_ALLOW_RANDOM_INIT") == "1":
        return
    warnings.warn(f"clipcap_b200: no pretrained weights given for the {what} encoder (weights_path / ${env_var}); the "
                  "tower is RANDOMLY INITIALISED and its embeddings are meaningless. Pass allow_random_init=True or set "
                  "CLIPCAP_B200_ALLOW_RANDOM_INIT=1 if that is intended (benchmarks, tests).", RuntimeWarning,
                  stacklevel=3)


def get_clip_encoder(encoder_model_variant: str, window_size: Optional[int] = None, normalize_embeddings: bool = False,
                     use_windowed_embeddings: bool = False, window_overlap_percentage: float = 0.0,
                     device: str = "cuda", weights_path: Optional[str] = None,
                     allow_random_init: bool = False) -> Tuple[Callable, Callable]:
    """clip.py:132-153. `weights_path` (or $CLIPCAP_B200_CLIP_WEIGHTS) = torch-saved state_dict with OpenAI `visual.*`
    keys. The reference always loads pretrained weights (`clip.load`); without a path the tower keeps its random
    initialisation (no network here) and says so with a warning unless `allow_random_init` (benchmarks, tests) or
    $CLIPCAP_B200_ALLOW_RANDOM_INIT=1 is set."""
    if encoder_model_variant not in CLIP_VARIANTS:
        raise ValueError(f"clipcap_b200 implements the CLIP ViT towers {sorted(CLIP_VARIANTS)}; "
                         f"got '{encoder_model_variant}'")
    image_size, patch, width, layers, heads, out_dim = CLIP_VARIANTS[encoder_model_variant]
    tower = ViTImageTower(image_size, patch, width, layers, heads, out_dim)
    weights_path = weights_path or os.environ.get("CLIPCAP_B200_CLIP_WEIGHTS")
    if weights_path:
        sd = torch.load(weights_path, map_location="cpu")
        sd = {k: v.float() for k, v in sd.items() if k.startswith("visual.")}
        tower.load_state_dict(sd, strict=True)
    else:
        warn_random_init("CLIP " + encoder_model_variant, "CLIPCAP_B200_CLIP_WEIGHTS", allow_random_init)
    transform = TensorTransform(image_size, use_windowed_embeddings, window_size if window_size else 9,
                                window_overlap_percentage, device)
    model = CLIPModel(tower, normalize_embeddings=normalize_embeddings, use_windowed_embeddings=use_windowed_embeddings)
    model = model.eval()
    model = model.to(device)
    return model, transform
