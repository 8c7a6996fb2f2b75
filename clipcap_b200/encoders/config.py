"""`clipcap.encoders.config` name (EncoderConfig); the record lives in clipcap_b200/configs.py."""
from clipcap_b200.configs import EncoderConfig  # noqa: F401

__all__ = ["EncoderConfig"]
