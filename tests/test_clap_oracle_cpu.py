"""Pins oracle/restate_clap.py (the CLAP audio tower of BASELINE configs[4]; its CUDA path is checked against it in tests/test_clap_gpu.py) against
transformers.ClapAudioModelWithProjection, the stand-in SURVEY §8c names for the un-installed `laion_clap`."""
import pytest
import torch

from conftest import rel_err
from oracle import restate_clap as RC

TOL = 2e-5


@pytest.mark.parametrize("fusion,longer", [(True, [1, 0, 1]), (True, [0, 0, 0]), (False, [0, 0, 0])])
def test_clap_restatement_matches_hf(fusion, longer):
    cfg = RC.ClapCfg(depths=(1, 2, 2, 1), enable_fusion=fusion)  # shifted and unshifted blocks in the middle stages
    m = RC.hf_clap(cfg, seed=3)
    w = {k: v.detach() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(5)
    B = len(longer)
    mel = torch.randn(B, 4 if fusion else 1, 1001, 64, generator=g)
    is_longer = torch.tensor(longer, dtype=torch.bool).view(B, 1)
    with torch.no_grad():
        want = m(input_features=mel, is_longer=is_longer).audio_embeds
        got = RC.clap_audio_embed(w, mel, is_longer, cfg)
    assert tuple(got.shape) == (B, 512)
    assert rel_err(got, want) < TOL


def test_clap_htsat_tiny_full_depth():
    """The real HTSAT-tiny layout (depths 2/2/6/2, 28.2 M parameters), one fused and one plain sample."""
    cfg = RC.ClapCfg()
    m = RC.hf_clap(cfg, seed=0)
    assert abs(sum(p.numel() for p in m.parameters()) / 1e6 - 28.2) < 0.8
    w = {k: v.detach() for k, v in m.state_dict().items()}
    mel = torch.randn(2, 4, 1001, 64, generator=torch.Generator().manual_seed(1))
    is_longer = torch.tensor([[True], [False]])
    with torch.no_grad():
        want = m(input_features=mel, is_longer=is_longer).audio_embeds
        got = RC.clap_audio_embed(w, mel, is_longer, cfg)
    assert rel_err(got, want) < TOL
