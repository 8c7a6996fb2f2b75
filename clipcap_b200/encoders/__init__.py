"""Encoder side of the drop-in surface (clipcap/encoders/__init__.py): the factories, the config record and the CLIP tower."""
from clipcap_b200.configs import EncoderConfig
from clipcap_b200.encoders.base import get_encoder, get_encoder_from_config, get_encoder_from_model
from clipcap_b200.encoders.clip import CLIPModel, ViTImageTower

__all__ = ["CLIPModel", "EncoderConfig", "ViTImageTower", "get_encoder", "get_encoder_from_config",
           "get_encoder_from_model"]
