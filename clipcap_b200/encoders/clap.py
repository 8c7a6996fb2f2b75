"""CLAP audio encoder behind the reference's `CLAPModel` wrapper (clipcap/encoders/clap.py:105-131).

The reference hands raw 48 kHz waveforms to `laion_clap.CLAP_Module.get_audio_embedding_from_data`, a package that is not
installed here and whose wrapper does not run as committed (clap.py:136 reads `model_id` before assignment, :152 passes an
unknown keyword). What is built here is the tower that call ends in: HTSAT-tiny + the audio projection, with the arithmetic
and parameter names of transformers' `ClapAudioModelWithProjection` (the stand-in SURVEY §8c names), running in
libclipcap_b200's cc_clap_forward. Its input is the log-mel feature tensor `[B, channels, T <= 1024, 64]` (what
`ClapFeatureExtractor` / laion_clap's `get_mel` produce) plus the per-sample `is_longer` flags that select the
feature-fusion patch embedding; the waveform -> mel front end is host-side feature extraction and stays with the caller.
"""
from __future__ import annotations

import os
from typing import Callable, Optional, Tuple

import torch
import torch.nn as nn

from clipcap_b200.engine import ClapEngine
from clipcap_b200.model._lazy import EngineModule

HTSAT_TINY = dict(num_mel_bins=64, spec_size=256, patch=4, embed=96, depths=(2, 2, 6, 2), heads=(4, 8, 16, 32), window=8,
                  projection_dim=512)


class ClapAudioTower(EngineModule):
    """Parameter holder (a `ClapAudioModelWithProjection`, so checkpoints of that class load with load_state_dict) plus
    the native engine. Exposes `get_audio_embedding_from_mel(mel) -> [B, projection_dim]`."""

    def __init__(self, enable_fusion: bool = True, **overrides):
        super().__init__()
        from transformers import ClapAudioConfig, ClapAudioModelWithProjection
        self.arch = dict(HTSAT_TINY, **overrides)
        a = self.arch
        cfg = ClapAudioConfig(enable_fusion=enable_fusion, depths=list(a["depths"]), num_attention_heads=list(a["heads"]),
                              patch_embeds_hidden_size=a["embed"], hidden_size=a["embed"] * 8,
                              projection_dim=a["projection_dim"], window_size=a["window"], num_mel_bins=a["num_mel_bins"],
                              spec_size=a["spec_size"], patch_size=a["patch"], patch_stride=[a["patch"], a["patch"]])
        self.clap = ClapAudioModelWithProjection(cfg)

    def _engine_weights(self):
        return {k[len("clap."):]: v for k, v in self.state_dict().items() if k.startswith("clap.")}

    def _param_key(self):
        return super()._param_key() + tuple((b.data_ptr(), b._version) for b in self.buffers())

    def _build_engine(self, weights, capacity, device):
        return ClapEngine(weights, max_batch=capacity[0], device=device, **self.arch)

    @torch.no_grad()
    def get_audio_embedding_from_mel(self, mel: torch.Tensor, is_longer: Optional[torch.Tensor] = None,
                                     normalize: bool = False) -> torch.Tensor:
        return self._get_engine((max(8, mel.shape[0]),)).forward(mel, normalize=normalize, is_longer=is_longer)

    forward = get_audio_embedding_from_mel


class CLAPModel(nn.Module):
    """clipcap/encoders/clap.py:105-131: same constructor; `x` is the mel feature tensor and stays on the device."""

    def __init__(self, model: nn.Module, normalize_embeddings: bool = False) -> None:
        super().__init__()
        self.model = model
        self.normalize_embeddings = normalize_embeddings

    @torch.no_grad()
    def forward(self, x: torch.Tensor, is_longer: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self.model.get_audio_embedding_from_mel(x, is_longer=is_longer,
                                                       normalize=self.normalize_embeddings)  # fused L2 normalisation


def get_clap_encoder(normalize_embeddings: bool = False, device: str = "cuda",
                     weights_path: Optional[str] = None, allow_random_init: bool = False) -> Tuple[Callable, Callable]:
    """clap.py:133-161. `weights_path` (or $CLIPCAP_B200_CLAP_WEIGHTS): a torch-saved ClapAudioModelWithProjection
    state_dict; without it the tower keeps its random initialisation (no network here) and warns unless
    `allow_random_init` / $CLIPCAP_B200_ALLOW_RANDOM_INIT=1. The returned transform is the identity on mel tensors."""
    tower = ClapAudioTower()
    weights_path = weights_path or os.environ.get("CLIPCAP_B200_CLAP_WEIGHTS")
    if weights_path:
        tower.clap.load_state_dict(torch.load(weights_path, map_location="cpu"), strict=True)
    else:
        from clipcap_b200.encoders.clip import warn_random_init
        warn_random_init("CLAP audio", "CLIPCAP_B200_CLAP_WEIGHTS", allow_random_init)
    model = CLAPModel(tower, normalize_embeddings=normalize_embeddings).eval().to(device)
    return model, (lambda mel: mel)
