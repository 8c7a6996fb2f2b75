"""`clipcap.model.config` names (Config, TrainingConfig); the records live in clipcap_b200/configs.py."""
from clipcap_b200.configs import Config, EncoderConfig, TrainingConfig  # noqa: F401

__all__ = ["Config", "TrainingConfig"]
