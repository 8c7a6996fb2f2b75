"""Loads the committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the reference)."""
import os

import numpy as np
import pytest
import torch

from oracle import restate as R
from oracle import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

LM_CASES = {
    "tiny_a": ("tiny:128:2:2:1003:64", R.Gpt2Cfg(d=128, L=2, H=2, V=1003, n_pos=64),
               R.MapperCfg(E=64, d=128, P=3, K=5, H=2, L=2)),
    "tiny_b": ("tiny:192:3:3:517:48", R.Gpt2Cfg(d=192, L=3, H=3, V=517, n_pos=48),
               R.MapperCfg(E=72, d=192, P=2, K=4, H=2, L=1)),
}
SAMPLING_CASES = [  # must match tests/golden/make_golden.py
    ("nucleus", dict(top_p=0.8, top_k=0, temperature=0.9)),
    ("nucleus", dict(top_p=0.5, top_k=7, temperature=1.0)),
    ("nucleus", dict(top_p=0.3, top_k=1, temperature=1.0)),
    ("sample", dict(top_p=0.9, top_k=0.0, temperature=1.0, repetition_penalty=1.2)),
    ("sample", dict(top_p=0.0, top_k=5, temperature=0.7, repetition_penalty=1.5)),
    ("sample", dict(top_p=0.6, top_k=1, temperature=1.0, repetition_penalty=5.0)),
]
VIT_CASE = R.VitCfg(image_size=28, patch=14, width=128, layers=2, heads=2, mlp_dim=512, out_dim=64)
ENTRY = 9


def load_lm_case(name):
    spec, gcfg, mcfg = LM_CASES[name]
    g = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    map_w, lm_w = synth.mapper_weights(mcfg), synth.gpt2_weights(gcfg, wte_std=0.1)
    cs = synth.checksum(map_w) + synth.checksum(lm_w)
    if abs(cs - float(g["w_checksum"])) > 1e-6 * abs(cs):
        pytest.skip("seeded weights differ from the ones the fixture was made with (torch RNG drift): regenerate "
                    "with tests/golden/make_golden.py")
    return spec, gcfg, mcfg, map_w, lm_w, {k: g[k] for k in g.files}


def load_vit_case():
    g = np.load(os.path.join(GOLDEN_DIR, "vit_tiny.npz"))
    w = synth.vit_weights(VIT_CASE)
    if abs(synth.checksum(w) - float(g["w_checksum"])) > 1e-6 * synth.checksum(w):
        pytest.skip("seeded weights differ from the fixture's (torch RNG drift)")
    return VIT_CASE, w, {k: g[k] for k in g.files}


def golden_rows(arr):
    return [[int(t) for t in row if t >= 0] for row in arr]


def load_train_case(name):
    """Fixture of the reference's training_step + backward (tests/golden/make_golden.py make_train)."""
    spec, gcfg, mcfg = LM_CASES[name]
    g = np.load(os.path.join(GOLDEN_DIR, f"train_{name}.npz"))
    map_w, lm_w = synth.mapper_weights(mcfg), synth.gpt2_weights(gcfg, wte_std=0.1)
    cs = synth.checksum(map_w) + synth.checksum(lm_w)
    if abs(cs - float(g["w_checksum"])) > 1e-6 * abs(cs):
        pytest.skip("seeded weights differ from the ones the fixture was made with (torch RNG drift)")
    grads = {k[5:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("grad:")}
    norms = {str(n): float(v) for n, v in zip(g["names"], g["norms"])}
    return spec, gcfg, mcfg, map_w, lm_w, torch.from_numpy(g["tokens"]), torch.from_numpy(g["emb"]), float(g["loss"]), grads, norms
