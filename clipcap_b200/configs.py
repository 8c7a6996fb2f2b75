"""Configuration records of the drop-in surface.

The YAML a reference training run leaves on disk (train/callback.py:16-18) is `Config.to_dict()`, so the field names and
defaults below are part of the interface (clipcap/model/config.py:7-55, clipcap/encoders/config.py:5-29) and a reference
config file loads unchanged. Everything else — `to_dict` and the argparse constructor — is shared machinery here:
`from_args` copies every field the namespace carries and leaves the ones named in `_NOT_FROM_ARGS` at None, which is what
the reference's hand-written constructors do (nested configs and the embedding size are filled in later by its callers).
"""
from __future__ import annotations

import dataclasses
from argparse import Namespace
from typing import Optional, Tuple


class _Record:
    _NOT_FROM_ARGS: Tuple[str, ...] = ()

    def to_dict(self) -> dict:
        return dataclasses.asdict(self)

    @classmethod
    def from_args(cls, args: Namespace):
        picked = {}
        for f in dataclasses.fields(cls):
            picked[f.name] = None if f.name in cls._NOT_FROM_ARGS else getattr(args, cls._ARG_NAMES.get(f.name, f.name))
        return cls(**picked)

    _ARG_NAMES: dict = {}


@dataclasses.dataclass
class EncoderConfig(_Record):
    encoder_model_name: str = "clip"
    encoder_model_variant: str = "ViT-L/14"
    encoder_embedding_size: Optional[int] = None  # known once the embeddings have been read (the reference's dataloader)
    normalize_embeddings: bool = False
    use_windowed_embeddings: bool = False
    window_size: int = 16                          # 4 x 4 tiles
    window_overlap_percentage: float = 0.0

    _NOT_FROM_ARGS = ("encoder_embedding_size",)


@dataclasses.dataclass
class TrainingConfig(_Record):
    optimizer_lr: float = 2e-5
    use_deepspeed_optimisers: bool = True
    scheduler_warmup_steps: int = 123
    total_steps: int = 123

    _ARG_NAMES = {"use_deepspeed_optimisers": "enable_deepspeed"}  # the CLI flag's name (model/config.py:19-25)


@dataclasses.dataclass
class Config(_Record):
    language_model: str = "gpt2-xl"
    train_language_model: bool = False
    prefix_length: int = 10
    projection_length: int = 10
    transformer_layers: int = 8
    transformer_attention_heads: int = 16
    use_positional_embeddings: bool = True
    encoder_config: Optional[EncoderConfig] = None
    training_config: Optional[TrainingConfig] = None

    _NOT_FROM_ARGS = ("encoder_config", "training_config")
