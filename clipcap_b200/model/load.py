"""`load(model_path, config_path, device, from_checkpoint)` — same contract as clipcap/model/load.py:9-43."""
from typing import Callable, Tuple, Union

import torch
import yaml

from clipcap_b200.encoders.config import EncoderConfig
from clipcap_b200.model.config import Config
from clipcap_b200.model.model import ClipCapModel, ClipCapModelPrefixOnly, get_tokenizer


def load(model_path: str, config_path: str, device: str = "cpu",
         from_checkpoint: bool = False) -> Tuple[Union[ClipCapModel, ClipCapModelPrefixOnly], Callable]:
    with open(config_path, "r") as f:
        raw_config = yaml.safe_load(f)

    # Remove old training config data from past training runs (load.py:14-16).
    if from_checkpoint and raw_config["training_config"] is not None:
        raw_config["training_config"] = None

    raw_config["encoder_config"] = EncoderConfig(**raw_config["encoder_config"])
    config = Config(**raw_config)

    model_cls = ClipCapModel if config.train_language_model else ClipCapModelPrefixOnly
    model = model_cls(config)

    state_dict = torch.load(model_path, map_location="cpu")
    if from_checkpoint:
        state_dict = state_dict["state_dict"]
    model.load_state_dict(state_dict, strict=False)

    model = model.eval()
    model = model.to(device)

    tokenizer = get_tokenizer(config.language_model)
    return model, tokenizer
