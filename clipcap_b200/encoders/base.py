"""Encoder factory — same names, arguments and error behaviour as clipcap/encoders/base.py:10-39."""
from typing import Callable, Optional, Tuple, Union

from torch.nn import Module

from clipcap_b200.encoders.clip import get_clip_encoder
from clipcap_b200.encoders.config import EncoderConfig


def get_encoder(encoder_model_name: str, encoder_model_variant: str, normalize_embeddings: bool = False,
                window_size: Optional[int] = None, use_windowed_embeddings: bool = False,
                window_overlap_percentage: float = 0.0, device: str = "cuda") -> Tuple[Module, Callable]:
    kwargs = {"normalize_embeddings": normalize_embeddings, "device": device}
    if encoder_model_name == "clip":
        return get_clip_encoder(encoder_model_variant, use_windowed_embeddings=use_windowed_embeddings,
                                window_size=window_size, window_overlap_percentage=window_overlap_percentage, **kwargs)
    elif encoder_model_name == "clap":
        # the CLAP path is broken in the reference as committed (clap.py:136,152) and is a "next" row.
        raise NotImplementedError("the CLAP audio encoder is not part of the clipcap_b200 hot path yet")
    else:
        raise ValueError(f"invalid encoder name: '{encoder_model_name}'")


def get_encoder_from_config(config: EncoderConfig, device: str = "cpu") -> Tuple[Module, Callable]:
    if config.encoder_model_name == "clip":
        config.encoder_model_variant = config.encoder_model_variant.replace("_", "/")
    return get_encoder(
        config.encoder_model_name, config.encoder_model_variant, normalize_embeddings=config.normalize_embeddings,
        use_windowed_embeddings=config.use_windowed_embeddings, window_size=config.window_size,
        window_overlap_percentage=config.window_overlap_percentage, device=device)


def get_encoder_from_model(model, device: str = "cpu") -> Tuple[Module, Callable]:
    return get_encoder_from_config(model.config.encoder_config, device=device)
