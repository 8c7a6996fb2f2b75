"""Config / TrainingConfig — same fields and defaults as the reference (clipcap/model/config.py:7-55), so the YAML
written by the reference's training run (train/callback.py:16-18) loads unchanged."""
from argparse import Namespace
from dataclasses import asdict, dataclass
from typing import Optional

from clipcap_b200.encoders.config import EncoderConfig


@dataclass
class TrainingConfig:
    optimizer_lr: float = 2e-5
    use_deepspeed_optimisers: bool = True
    scheduler_warmup_steps: int = 123
    total_steps: int = 123

    def to_dict(self) -> dict:
        return asdict(self)


@dataclass
class Config:
    language_model: str = "gpt2-xl"
    train_language_model: bool = False
    prefix_length: int = 10
    projection_length: int = 10
    transformer_layers: int = 8
    transformer_attention_heads: int = 16
    use_positional_embeddings: bool = True

    encoder_config: Optional[EncoderConfig] = None
    training_config: Optional[TrainingConfig] = None

    def to_dict(self) -> dict:
        return asdict(self)

    @classmethod
    def from_args(cls, args: Namespace) -> "Config":
        return cls(
            language_model=args.language_model,
            train_language_model=args.train_language_model,
            prefix_length=args.prefix_length,
            projection_length=args.projection_length,
            transformer_layers=args.transformer_layers,
            transformer_attention_heads=args.transformer_attention_heads,
            use_positional_embeddings=args.use_positional_embeddings,
            encoder_config=None,
            training_config=None,
        )
