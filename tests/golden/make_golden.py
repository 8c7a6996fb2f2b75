"""Freezes outputs of the REFERENCE (TheoCoombes/ClipCap imported unmodified from /root/reference, plus the third-party
modules it calls: transformers GPT-2 / CLIP) on seeded inputs into tests/golden/*.npz. Run in the build container:

    python tests/golden/make_golden.py

The reference ships no golden vectors of its own, so these are "outputs of the reference itself run here".
Weights are regenerated from seeds at test time (oracle/synth.py); each fixture stores a checksum of them so RNG drift
between torch builds is detected instead of silently comparing against different weights.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_runner as RR  # noqa: E402
from oracle import restate as R  # noqa: E402
from oracle import synth  # noqa: E402

CASES = {
    # name: (lm spec, Gpt2Cfg, MapperCfg, beams)
    "tiny_a": ("tiny:128:2:2:1003:64", R.Gpt2Cfg(d=128, L=2, H=2, V=1003, n_pos=64),
               R.MapperCfg(E=64, d=128, P=3, K=5, H=2, L=2)),
    "tiny_b": ("tiny:192:3:3:517:48", R.Gpt2Cfg(d=192, L=3, H=3, V=517, n_pos=48),
               R.MapperCfg(E=72, d=192, P=2, K=4, H=2, L=1)),
}
SAMPLING_CASES = [  # (mode, kwargs of the reference function)
    ("nucleus", dict(top_p=0.8, top_k=0, temperature=0.9)),
    ("nucleus", dict(top_p=0.5, top_k=7, temperature=1.0)),
    ("nucleus", dict(top_p=0.3, top_k=1, temperature=1.0)),
    ("sample", dict(top_p=0.9, top_k=0.0, temperature=1.0, repetition_penalty=1.2)),
    ("sample", dict(top_p=0.0, top_k=5, temperature=0.7, repetition_penalty=1.5)),
    ("sample", dict(top_p=0.6, top_k=1, temperature=1.0, repetition_penalty=5.0)),
]
VIT = R.VitCfg(image_size=28, patch=14, width=128, layers=2, heads=2, mlp_dim=512, out_dim=64)
ENTRY = 9


def main():
    assert RR.available(), "needs /root/reference"
    torch.manual_seed(0)
    for name, (spec, gcfg, mcfg) in CASES.items():
        map_w, lm_w = synth.mapper_weights(mcfg), synth.gpt2_weights(gcfg, wte_std=0.1)
        model = RR.build_reference_model(spec, mcfg.E, mcfg.K, mcfg.P, mcfg.H, mcfg.L, map_w, lm_w)
        B = 4
        emb = synth.embeddings(B, mcfg.E)
        out = {"emb": emb.numpy(), "w_checksum": np.float64(synth.checksum(map_w) + synth.checksum(lm_w))}
        with torch.no_grad():
            prefix = model.transformer_mapper(emb)                       # reference TransformerMapper
            out["prefix"] = prefix.numpy()
            out["logits"] = model.language_model(inputs_embeds=prefix).logits.numpy()  # HF GPT-2 via the reference model
            free = RR.reference_generate_beam(model, prefix[:1], 1, ENTRY, 1.0, stop_token=gcfg.V + 7)
            stop = free[4]
            out["stop_token"] = np.int64(stop)
            for beam in (1, 3, 5):
                for temp in (1.0, 0.7):
                    rows = [RR.reference_generate_beam(model, prefix[i:i + 1], beam, ENTRY, temp, stop) for i in range(B)]
                    arr = np.full((B, ENTRY), -1, dtype=np.int64)
                    for i, r in enumerate(rows):
                        arr[i, :len(r)] = r
                    out[f"beam{beam}_t{temp}"] = arr
            # sampling loops (nucleus_sampling.py / no_beam.py) with torch.multinomial replaced by RR.rank_cycle_pick
            tp = torch.tensor([[5, 17, 5]])
            for ci, (mode, kw) in enumerate(SAMPLING_CASES):
                for ti, tpt in enumerate((None, tp)):
                    toks, dists = RR.reference_generate_sampling(model, prefix[:1], mode, ENTRY, text_prefix_tokens=tpt, **kw)
                    out[f"samp{ci}_tp{ti}_tokens"] = np.asarray(toks, dtype=np.int64)
                    out[f"samp{ci}_tp{ti}_nkept"] = np.asarray([int((d > 0).sum()) for d in dists], dtype=np.int64)
                    out[f"samp{ci}_tp{ti}_pmax"] = np.asarray([float(d.max()) for d in dists], dtype=np.float64)
            # ClipCapModel.forward (teacher-forced logits, model.py:43-58)
            tokens = torch.randint(0, gcfg.V, (B, 6), generator=torch.Generator().manual_seed(3))
            mask = torch.ones(B, 6, dtype=torch.bool)
            out["fwd_tokens"] = tokens.numpy()
            out["fwd_logits"] = model(tokens, emb, mask).logits.numpy()
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(name, {k: getattr(v, "shape", v) for k, v in out.items()})
    # ViT (HF CLIP stand-in inside the reference's CLIPModel wrapper), with and without normalisation
    vit_w = synth.vit_weights(VIT)
    px = synth.pixels(3, VIT.image_size)
    with torch.no_grad():
        out = {"pixels": px.numpy(), "w_checksum": np.float64(synth.checksum(vit_w)),
               "emb": RR.reference_clip_model(VIT, vit_w, False)(px).numpy(),
               "emb_norm": RR.reference_clip_model(VIT, vit_w, True)(px.clone()).numpy()}
    np.savez_compressed(os.path.join(HERE, "vit_tiny.npz"), **out)
    print("vit_tiny", {k: getattr(v, "shape", v) for k, v in out.items()})


TRAIN_KEEP = ("linear.bias", "prefix_const", "transformer.layers.0.norm1.weight", "transformer.layers.0.norm1.bias",
              "transformer.layers.0.attn.to_queries.weight", "transformer.layers.0.attn.project.bias",
              "transformer.layers.0.norm2.weight", "transformer.layers.0.mlp.fc1.bias",
              "transformer.layers.0.mlp.fc2.bias")  # full tensors; every other gradient is pinned through its L2 norm


def train_batch(gcfg, mcfg, B=4, Tt=7):
    """Seeded caption batch as the reference's dataloader hands it to training_step: -1 padding at the end of some rows,
    one real token with id 0 (which ignore_index=0 also drops — reference quirk, model.py:110)."""
    g = torch.Generator().manual_seed(11)
    tokens = torch.randint(1, gcfg.V, (B, Tt), generator=g)
    tokens[1, Tt - 2:] = -1
    tokens[2, Tt - 4:] = -1
    tokens[3, 1] = 0
    return tokens, synth.embeddings(B, mcfg.E, seed=77)


def make_train():
    """training_step + loss.backward() of the reference's ClipCapModelPrefixOnly (model.py:94-123) -> train_<case>.npz"""
    for name, (spec, gcfg, mcfg) in CASES.items():
        map_w, lm_w = synth.mapper_weights(mcfg), synth.gpt2_weights(gcfg, wte_std=0.1)
        model = RR.build_reference_model(spec, mcfg.E, mcfg.K, mcfg.P, mcfg.H, mcfg.L, map_w, lm_w)
        tokens, emb = train_batch(gcfg, mcfg)
        loss, grads = RR.reference_training_step(model, tokens, emb)
        out = {"tokens": tokens.numpy(), "emb": emb.numpy(), "loss": np.float64(loss),
               "w_checksum": np.float64(synth.checksum(map_w) + synth.checksum(lm_w)),
               "names": np.asarray(sorted(grads)), "norms": np.asarray([float(grads[k].norm()) for k in sorted(grads)])}
        for k in TRAIN_KEEP:
            out["grad:" + k] = grads[k].numpy()
        np.savez_compressed(os.path.join(HERE, f"train_{name}.npz"), **out)
        print("train_" + name, "loss", loss, {k: getattr(v, "shape", v) for k, v in out.items() if k.startswith("grad:")})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "train":
        make_train()
    else:
        main()
        make_train()
