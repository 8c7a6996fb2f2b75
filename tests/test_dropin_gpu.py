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


@pytest.mark.parametrize("name", ["tiny_a", "tiny_b"])
def test_model_forward_matches_golden(cuda_device, name):
    """ClipCapModel.forward(tokens, embeddings, mask) (clipcap/model/model.py:43-58) on the GPU against the logits the
    UNMODIFIED reference produced for the same weights and inputs (tests/golden/tiny_{a,b}.npz::fwd_logits, frozen by
    tests/golden/make_golden.py), incl. the -1-padded mask handling of training batches."""
    from golden_util import load_lm_case
    from clipcap_b200.encoders.config import EncoderConfig
    from clipcap_b200.model import ClipCapModelPrefixOnly, Config
    spec, gcfg, mcfg, map_w, lm_w, g = load_lm_case(name)
    cfg = Config(language_model=spec, prefix_length=mcfg.K, projection_length=mcfg.P, transformer_layers=mcfg.L,
                 transformer_attention_heads=mcfg.H, encoder_config=EncoderConfig(encoder_embedding_size=mcfg.E))
    model = ClipCapModelPrefixOnly(cfg)
    sd = {f"transformer_mapper.{k}": v for k, v in map_w.items()}
    sd.update({f"language_model.{k}": v for k, v in lm_w.items()})
    model.load_state_dict(sd, strict=True)
    model = model.eval().to(cuda_device)
    emb = torch.from_numpy(g["emb"]).to(cuda_device)
    tokens = torch.from_numpy(g["fwd_tokens"]).to(cuda_device)
    mask = torch.ones_like(tokens, dtype=torch.bool)
    want = torch.from_numpy(g["fwd_logits"])
    out = model(tokens, emb, mask)
    assert tuple(out.logits.shape) == tuple(want.shape) == (tokens.shape[0], mcfg.K + tokens.shape[1], gcfg.V)
    err = rel_err(out.logits, want)
    assert err < 1e-3, err
    # trailing padding (what the reference's dataloader produces): the logits of the real positions do not change
    mask[:, -2:] = False
    out2 = model(tokens, emb, mask)
    keep = mcfg.K + tokens.shape[1] - 2
    assert rel_err(out2.logits[:, :keep], want[:, :keep]) < 1e-3


@pytest.mark.parametrize("size,window_size,overlap", [(224, 9, 0.0), (300, 16, 0.0), (257, 4, 25.0), (96, 9, 10.0)])
def test_window_tiling_equals_reference_unfold(cuda_device, size, window_size, overlap):
    """cc_op_tile_image against the reference's own tiling arithmetic (CLIPTransform.tile_image, clipcap/encoders/clip.py:
    60-82: tensor.unfold(1, p, step).unfold(2, p, step), step from the overlap percentage), bit for bit, and the whole
    windowed transform's output layout [window_size + 1, 3, S, S] (clip.py:95-101) feeding the windowed encoder call."""
    import math
    from clipcap_b200.encoders.clip import TensorTransform
    n = int(math.sqrt(window_size))
    img = torch.rand(3, size, size, generator=torch.Generator().manual_seed(size))
    tr = TensorTransform(28, use_windowed_embeddings=True, window_size=window_size, window_overlap_percentage=overlap,
                         device=cuda_device)
    target = math.ceil(size / n) * n
    sq = img if target == size else torch.nn.functional.interpolate(img[None], size=(target, target), mode="bilinear",
                                                                    align_corners=False)[0]
    p = target // n
    step = math.floor(p * (1 - overlap / 100)) if overlap != 0 else p
    ref = sq.unfold(1, p, step).unfold(2, p, step)[:, :n, :n]          # [3, ny, nx, p, p]   (clip.py:73-80)
    want = ref.permute(1, 2, 0, 3, 4).reshape(n * n, 3, p, p)            # tile-major, channels inside each tile
    got = tr.tile_image(sq.to(cuda_device))
    assert torch.equal(got.cpu(), want)
    out = tr(img)
    assert tuple(out.shape) == (window_size + 1, 3, 28, 28) and out.is_cuda
    plain = TensorTransform(28)
    assert rel_err(out[0], plain(img)) < 1e-5                            # global view first (clip.py:97-99)
    assert rel_err(out[1], plain(want[0])) < 1e-5                        # then the tiles, each a CLIP-normalised image
    with pytest.raises(RuntimeError):
        tr.tile_image(sq)                                                # CPU tensor: no CPU path
    with pytest.raises(AssertionError):
        TensorTransform(28, use_windowed_embeddings=True, window_size=8)  # clip.py:14-15


def test_get_encoder_errors():
    import clipcap_b200 as clipcap
    with pytest.raises(ValueError, match="invalid encoder name"):
        clipcap.get_encoder("nope", "x")


def test_windowed_sequence(cuda_device):
    """use_windowed_embeddings=True end to end (clipcap/encoders/clip.py:112-129 flatten/unflatten + TransformerMapperWindowed,
    mapper.py:133-160, window_size passed as W+1, model.py:28): [B, W+1, 3, S, S] pixels -> [B, W+1, E] -> prefix -> beam."""
    from clipcap_b200.encoders.clip import CLIPModel, ViTImageTower
    from clipcap_b200.encoders.config import EncoderConfig
    from clipcap_b200.inference.base import generate_beam_tokens
    from clipcap_b200.model import ClipCapModelPrefixOnly, Config
    gcfg = R.Gpt2Cfg(d=128, L=2, H=2, V=1003, n_pos=64)
    W = 3  # window_size 2 -> W + 1 = 3 windows per sample (global view + 2 tiles)
    mcfg = R.MapperCfg(kind="windowed", E=64, d=128, P=2, K=5, H=2, L=2, W=W, use_pos=True)
    vcfg = R.VitCfg(image_size=28, patch=14, width=128, layers=2, heads=2, mlp_dim=512, out_dim=64)
    map_w, lm_w, vit_w = synth.mapper_weights(mcfg), synth.gpt2_weights(gcfg, wte_std=0.1), synth.vit_weights(vcfg)
    cfg = Config(language_model="tiny:128:2:2:1003:64", prefix_length=5, projection_length=2, transformer_layers=2,
                 transformer_attention_heads=2, use_positional_embeddings=True,
                 encoder_config=EncoderConfig(encoder_embedding_size=64, use_windowed_embeddings=True, window_size=W - 1))
    model = ClipCapModelPrefixOnly(cfg)
    sd = {f"transformer_mapper.{k}": v for k, v in map_w.items()}
    sd.update({f"language_model.{k}": v for k, v in lm_w.items()})
    model.load_state_dict(sd, strict=True)
    model = model.eval().to("cuda")
    tower = ViTImageTower(vcfg.image_size, vcfg.patch, vcfg.width, vcfg.layers, vcfg.heads, vcfg.out_dim, vcfg.mlp_dim)
    tower.load_state_dict(vit_w, strict=True)
    encode_fn = CLIPModel(tower, normalize_embeddings=True, use_windowed_embeddings=True).eval().to("cuda")

    B = 2
    px = synth.pixels(B * W, vcfg.image_size).view(B, W, 3, vcfg.image_size, vcfg.image_size)
    emb_ref = R.vit_encode(vit_w, px.flatten(0, 1), vcfg, normalize=True).view(B, W, -1)
    prefix_ref = R.mapper_forward(map_w, emb_ref, mcfg)
    emb = encode_fn(px.to("cuda"))
    assert tuple(emb.shape) == (B, W, 64) and rel_err(emb, emb_ref) < 1e-3
    prefix = model.transformer_mapper(emb)
    assert tuple(prefix.shape) == (B, 5, 128) and rel_err(prefix, prefix_ref) < 2e-3
    toks, lens, _ = generate_beam_tokens(model, prefix, beam_size=1, entry_length=8, stop_token=1002)
    oracle = R.generate_greedy_batch(lm_w, gcfg, prefix_ref, 8, 1002)
    exact, _ = check_tokens_against_oracle(toks, lens, oracle, margin_tol=5e-3)
    assert exact >= B - 1


def test_windowed_mapper_reference_size(cuda_device):
    """The reference's default windowed shape: 16 windows + the global view (W+1 = 17), P=10, K=40 -> 210 tokens, d=1024."""
    from clipcap_b200.engine import MapperEngine
    cfg = R.MapperCfg(kind="windowed", E=768, d=1024, P=10, K=40, H=8, L=2, W=17, use_pos=True)
    w = synth.mapper_weights(cfg, 3)
    emb = synth.embeddings(2 * 17, 768, seed=8).view(2, 17, 768)
    ref = R.mapper_forward(w, emb, cfg)
    eng = MapperEngine(w, kind="windowed", E=768, d=1024, P=10, K=40, H=8, L=2, W=17, use_pos=True, max_batch=2,
                       device=cuda_device)
    assert rel_err(eng.forward(emb.to(cuda_device)), ref) < 1e-3
