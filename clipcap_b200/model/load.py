"""`clipcap.load` — builds a model from what a reference training run leaves on disk (clipcap/model/load.py:9-43):
`<prefix>_config.yaml` (= Config.to_dict(), train/callback.py:16-18) and a state_dict (plain, or the "state_dict" entry of
a Lightning checkpoint). Same arguments, same return value `(model, tokenizer)`, same order of effects: stale training
settings of a checkpoint are dropped, keys are loaded non-strictly, the model is put in eval mode on `device`."""
from pathlib import Path
from typing import Callable, Tuple, Union

import torch
import yaml

from clipcap_b200.configs import Config, EncoderConfig
from clipcap_b200.model.model import ClipCapModel, ClipCapModelPrefixOnly, get_tokenizer

Model = Union[ClipCapModel, ClipCapModelPrefixOnly]


def _config_from_yaml(config_path, from_checkpoint: bool) -> Config:
    fields = yaml.safe_load(Path(config_path).read_text())
    if from_checkpoint and fields["training_config"] is not None:
        fields["training_config"] = None  # settings of the run that wrote the checkpoint do not carry over (load.py:14-16)
    fields["encoder_config"] = EncoderConfig(**fields["encoder_config"])
    return Config(**fields)


def _weights(model_path, from_checkpoint: bool) -> dict:
    blob = torch.load(model_path, map_location="cpu")
    return blob["state_dict"] if from_checkpoint else blob


def load(model_path: str, config_path: str, device: str = "cpu", from_checkpoint: bool = False) -> Tuple[Model, Callable]:
    config = _config_from_yaml(config_path, from_checkpoint)
    model = (ClipCapModel if config.train_language_model else ClipCapModelPrefixOnly)(config)
    model.load_state_dict(_weights(model_path, from_checkpoint), strict=False)
    model = model.eval().to(device)
    return model, get_tokenizer(config.language_model)
