"""ClipCapModel / ClipCapModelPrefixOnly with the reference's attribute names and constructor
(clipcap/model/model.py:13-123): `.language_model`, `.transformer_mapper`, `.config`, `forward(tokens, embeddings,
mask)`. The training-only methods (configure_optimizers / training_step, model.py:67-113) are out of scope
 and raise."""
from __future__ import annotations

import warnings
from typing import Callable

import torch
import torch.nn as nn

from clipcap_b200.model.config import Config, TrainingConfig
from clipcap_b200.model.lm import GPT2LM
from clipcap_b200.model.mapper import TransformerMapper, TransformerMapperWindowed


class IdTokenizer:
    """Offline stand-in when no tokenizer files are cached: captions are returned as space-separated token ids."""
    eos_token = "<|endoftext|>"

    def __init__(self, eos_id: int = 50256):
        self.eos_id = eos_id

    def encode(self, text):
        if text == self.eos_token:
            return [self.eos_id]
        if text == ".":
            return [13]
        return [int(t) for t in text.split()]

    def decode(self, ids):
        return " ".join(str(int(i)) for i in ids)


def get_tokenizer(language_model_name: str, **huggingface_kwargs) -> Callable:
    """model.py:10-11. Uses the HF tokenizer when its files are available locally; there is no network here."""
    if language_model_name.startswith("tiny:"):
        return IdTokenizer(int(language_model_name.split(":")[4]) - 1)
    try:
        from transformers import AutoTokenizer
        return AutoTokenizer.from_pretrained(language_model_name, local_files_only=True, **huggingface_kwargs)
    except Exception as e:  # noqa: BLE001
        warnings.warn(f"tokenizer files for '{language_model_name}' not available offline ({type(e).__name__}); "
                      "captions will be returned as token ids")
        return IdTokenizer()


class ClipCapModel(nn.Module):
    def __init__(self, config: Config):
        super().__init__()
        self.config = config
        self.language_model = GPT2LM(self.config.language_model)
        self.lm_embedding_size = self.language_model.get_input_embeddings().weight.shape[1]
        enc = self.config.encoder_config
        if enc.use_windowed_embeddings:
            self.transformer_mapper = TransformerMapperWindowed(
                encoder_embedding_size=enc.encoder_embedding_size, lm_embedding_size=self.lm_embedding_size,
                prefix_length=self.config.prefix_length, projection_length=self.config.projection_length,
                window_size=(enc.window_size + 1), use_pos_embeddings=self.config.use_positional_embeddings,
                num_heads=self.config.transformer_attention_heads, num_layers=self.config.transformer_layers)
        else:
            self.transformer_mapper = TransformerMapper(
                encoder_embedding_size=enc.encoder_embedding_size, lm_embedding_size=self.lm_embedding_size,
                prefix_length=self.config.prefix_length, projection_length=self.config.projection_length,
                num_heads=self.config.transformer_attention_heads, num_layers=self.config.transformer_layers)

    @torch.no_grad()
    def forward(self, tokens: torch.Tensor, embeddings: torch.Tensor, mask: torch.Tensor):
        """model.py:43-58 — teacher-forced logits over [prefix, tokens]."""
        token_embeddings = self.language_model.get_input_embeddings()(tokens)
        prefix_projections = self.transformer_mapper(embeddings)
        inputs_embeds = torch.cat((prefix_projections.to(token_embeddings.dtype), token_embeddings), dim=1)
        prefix_mask = torch.ones(prefix_projections.shape[:-1], dtype=torch.bool, device=mask.device)
        mask = torch.cat((prefix_mask, mask), dim=1)
        return self.language_model(inputs_embeds=inputs_embeds, attention_mask=mask)

    def set_training_config(self, training_config: TrainingConfig, reinit_optims: bool = False) -> None:
        self.config.training_config = training_config

    def configure_optimizers(self):
        raise NotImplementedError("training is outside the clipcap_b200 inference hot path (SURVEY §8f)")

    def training_step(self, batch, _):
        raise NotImplementedError("training is outside the clipcap_b200 inference hot path (SURVEY §8f)")


class ClipCapModelPrefixOnly(ClipCapModel):
    def parameters(self, recurse: bool = True):
        return self.transformer_mapper.parameters()
