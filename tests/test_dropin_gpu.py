"""The reference's programmatic call sequence (docs/inference.md:14-34) run verbatim against clipcap_b200, checked
against the CPU oracle."""
import pytest
import torch
import yaml

from conftest import rel_err
from helpers import check_tokens_against_oracle
from oracle import restate as R
from oracle import synth

pytestmark = pytest.mark.gpu


def test_docs_inference_sequence(cuda_device, tmp_path):
    import clipcap_b200 as clipcap
    from clipcap_b200.inference.base import generate_beam, generate_beam_tokens
    from clipcap_b200.encoders.clip import ViTImageTower, CLIPModel

    lm_name = "tiny:128:2:2:1003:64"
    gcfg = R.Gpt2Cfg(d=128, L=2, H=2, V=1003, n_pos=64)
    mcfg = R.MapperCfg(E=64, d=128, P=3, K=5, H=2, L=2)
    vcfg = R.VitCfg(image_size=28, patch=14, width=128, layers=2, heads=2, mlp_dim=512, out_dim=64)
    map_w, lm_w, vit_w = synth.mapper_weights(mcfg), synth.gpt2_weights(gcfg, wte_std=0.1), synth.vit_weights(vcfg)

    # what a reference training run leaves on disk: <prefix>_config.yaml (train/callback.py:16-18) + state_dict
    cfg_yaml = {
        "language_model": lm_name, "train_language_model": False, "prefix_length": 5, "projection_length": 3,
        "transformer_layers": 2, "transformer_attention_heads": 2, "use_positional_embeddings": True,
        "encoder_config": {"encoder_model_name": "clip", "encoder_model_variant": "ViT-L_14",
                           "encoder_embedding_size": 64, "normalize_embeddings": False, "use_windowed_embeddings": False,
                           "window_size": 16, "window_overlap_percentage": 0.0},
        "training_config": None,
    }
    (tmp_path / "config.yaml").write_text(yaml.safe_dump(cfg_yaml))
    sd = {f"transformer_mapper.{k}": v for k, v in map_w.items()}
    sd.update({f"language_model.{k}": v for k, v in lm_w.items()})
    sd["language_model.lm_head.weight"] = lm_w["transformer.wte.weight"]
    torch.save(sd, tmp_path / "model.pt")

    model, tokenizer = clipcap.load(str(tmp_path / "model.pt"), str(tmp_path / "config.yaml"), device="cuda")
    assert model.config.encoder_config.encoder_model_variant == "ViT-L_14"

    # get_encoder_from_model would build a full ViT-L/14 with random weights; use the same wrapper around a tiny tower
    tower = ViTImageTower(vcfg.image_size, vcfg.patch, vcfg.width, vcfg.layers, vcfg.heads, vcfg.out_dim, vcfg.mlp_dim)
    tower.load_state_dict(vit_w, strict=True)
    encode_fn = CLIPModel(tower).eval().to("cuda")

    px = synth.pixels(3, vcfg.image_size)
    emb_ref, prefix_ref, oracle = R.caption_greedy(vit_w, map_w, lm_w, vcfg, mcfg, gcfg, px, 8, stop_token=1002)
    embedding = encode_fn(px.to("cuda"))
    assert rel_err(embedding, emb_ref) < 1e-3
    prefix = model.transformer_mapper(embedding)
    assert rel_err(prefix, prefix_ref) < 2e-3
    # language_model facade: logits of the reference's per-step call (base.py:81-83)
    logits = model.language_model(inputs_embeds=prefix).logits
    assert rel_err(logits, R.gpt2_logits(lm_w, prefix_ref, gcfg)) < 2e-3
    tok_emb = model.language_model.get_input_embeddings()(torch.tensor([[1, 2]], device="cuda"))
    assert torch.equal(tok_emb.cpu(), lm_w["transformer.wte.weight"][[1, 2]][None])

    captions = generate_beam(model, tokenizer, prefix, beam_size=1, entry_length=8)
    assert len(captions) == 3
    toks, lens, _ = generate_beam_tokens(model, prefix, beam_size=1, entry_length=8, stop_token=1002)
    exact, _ = check_tokens_against_oracle(toks, lens, oracle, margin_tol=5e-3)
    assert exact >= 2
    for i in range(3):
        if toks[i].tolist()[:lens[i]] == oracle[i][0]:
            assert captions[i] == " ".join(str(t) for t in oracle[i][0])

    # single image, as the reference is called
    one = generate_beam(model, tokenizer, prefix[:1], beam_size=3, entry_length=8)
    assert len(one) == 1 and isinstance(one[0], str)


def test_get_encoder_errors():
    import clipcap_b200 as clipcap
    with pytest.raises(ValueError, match="invalid encoder name"):
        clipcap.get_encoder("nope", "x")
