"""Encoder factories of the drop-in surface — `get_encoder`, `get_encoder_from_config`, `get_encoder_from_model` with the
reference's names, arguments, return value `(encoder module, transform)` and error for an unknown encoder name
(clipcap/encoders/base.py:10-39). Dispatch is a table of builders: the CLIP image towers and the CLAP audio tower."""
from typing import Callable, Dict, Optional, Tuple

from torch.nn import Module

from clipcap_b200.configs import EncoderConfig
from clipcap_b200.encoders.clip import get_clip_encoder

Encoder = Tuple[Module, Callable]


def _build_clip(variant: str, **options) -> Encoder:
    return get_clip_encoder(variant, **options)


def _build_clap(variant: str, normalize_embeddings: bool = False, device: str = "cuda", **_unused) -> Encoder:
    # the reference's get_clap_encoder takes no variant / window options (clap.py:133); HTSAT-tiny is the only tower
    from clipcap_b200.encoders.clap import get_clap_encoder
    return get_clap_encoder(normalize_embeddings=normalize_embeddings, device=device)


_BUILDERS: Dict[str, Callable[..., Encoder]] = {"clip": _build_clip, "clap": _build_clap}


def get_encoder(encoder_model_name: str, encoder_model_variant: str, normalize_embeddings: bool = False,
                window_size: Optional[int] = None, use_windowed_embeddings: bool = False,
                window_overlap_percentage: float = 0.0, device: str = "cuda") -> Encoder:
    builder = _BUILDERS.get(encoder_model_name)
    if builder is None:
        raise ValueError(f"invalid encoder name: '{encoder_model_name}'")
    return builder(encoder_model_variant, normalize_embeddings=normalize_embeddings, device=device,
                   use_windowed_embeddings=use_windowed_embeddings, window_size=window_size,
                   window_overlap_percentage=window_overlap_percentage)


def get_encoder_from_config(config: EncoderConfig, device: str = "cpu") -> Encoder:
    if config.encoder_model_name == "clip":
        # YAML-safe variant names ("ViT-L_14") back to CLIP's ("ViT-L/14"); written back like the reference does (base.py:30)
        config.encoder_model_variant = config.encoder_model_variant.replace("_", "/")
    options = {k: getattr(config, k) for k in ("normalize_embeddings", "use_windowed_embeddings", "window_size",
                                                "window_overlap_percentage")}
    return get_encoder(config.encoder_model_name, config.encoder_model_variant, device=device, **options)


def get_encoder_from_model(model, device: str = "cpu") -> Encoder:
    return get_encoder_from_config(model.config.encoder_config, device=device)
