"""Host-side mirror of the reference interface: config dataclasses, state_dict key names, factories, error behaviour."""
import dataclasses

import pytest
import torch

from oracle import ref_runner as RR


def _small_model():
    from clipcap_b200.encoders import EncoderConfig
    from clipcap_b200.model import ClipCapModelPrefixOnly, Config
    return ClipCapModelPrefixOnly(Config(language_model="tiny:128:2:2:1003:64", prefix_length=5, projection_length=3,
                                         transformer_layers=2, transformer_attention_heads=2,
                                         encoder_config=EncoderConfig(encoder_embedding_size=64)))


def test_top_level_surface():
    import clipcap_b200
    for name in ("get_encoder", "get_encoder_from_model", "load"):  # clipcap/__init__.py:1-2
        assert callable(getattr(clipcap_b200, name))


def test_config_defaults_match_reference_dataclasses():
    from clipcap_b200.encoders import EncoderConfig
    from clipcap_b200.model import Config, TrainingConfig
    c = Config()
    assert (c.language_model, c.prefix_length, c.projection_length, c.transformer_layers,
            c.transformer_attention_heads, c.use_positional_embeddings) == ("gpt2-xl", 10, 10, 8, 16, True)
    e = EncoderConfig()
    assert (e.encoder_model_name, e.encoder_model_variant, e.window_size, e.normalize_embeddings) == \
        ("clip", "ViT-L/14", 16, False)
    assert set(Config(encoder_config=e).to_dict()["encoder_config"]) == {f.name for f in dataclasses.fields(EncoderConfig)}
    if RR.available():
        RR.import_reference()
        from clipcap.encoders.config import EncoderConfig as RE
        from clipcap.model.config import Config as RC, TrainingConfig as RT
        assert Config().to_dict() == RC().to_dict()
        assert EncoderConfig().to_dict() == RE().to_dict()
        assert TrainingConfig().to_dict() == RT().to_dict()


def test_state_dict_keys_match_reference():
    m = _small_model()
    keys = set(m.state_dict())
    assert "transformer_mapper.prefix_const" in keys and "transformer_mapper.linear.weight" in keys
    assert "transformer_mapper.transformer.layers.1.attn.to_keys_values.weight" in keys
    assert "language_model.transformer.h.1.attn.c_attn.weight" in keys
    assert m.state_dict()["language_model.transformer.h.0.attn.c_attn.weight"].shape == (128, 384)  # Conv1D [in, out]
    if RR.available():
        from oracle import restate as R, synth
        ref = RR.build_reference_model("tiny:128:2:2:1003:64", 64, 5, 3, 2, 2,
                                       synth.mapper_weights(R.MapperCfg(E=64, d=128, P=3, K=5, H=2, L=2)),
                                       synth.gpt2_weights(R.Gpt2Cfg(d=128, L=2, H=2, V=1003, n_pos=64)))
        ref_keys = {k for k in ref.state_dict() if not k.endswith("lm_head.weight")
                    and not k.endswith(".attn.bias") and not k.endswith(".attn.masked_bias")}
        assert keys == ref_keys
        # a reference checkpoint loads as is
        missing, unexpected = m.load_state_dict(ref.state_dict(), strict=False)
        assert not missing, missing


def test_mapper_default_init_is_rng_identical_to_reference():
    if not RR.available():
        pytest.skip("needs /root/reference")
    RR.import_reference()
    from clipcap.model.mapper import TransformerMapper as RefMapper
    from clipcap_b200.model import TransformerMapper
    torch.manual_seed(5)
    a = RefMapper(64, 128, 5, 3, 2, 2).state_dict()
    torch.manual_seed(5)
    b = TransformerMapper(64, 128, 5, 3, 2, 2).state_dict()
    assert list(a) == list(b)
    assert all(torch.equal(a[k], b[k]) for k in a)


def test_no_cpu_fallback():
    m = _small_model()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.transformer_mapper(torch.randn(2, 64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.language_model(inputs_embeds=torch.randn(1, 3, 128))


def test_factory_errors_match_reference():
    import clipcap_b200 as clipcap
    with pytest.raises(ValueError, match="invalid encoder name: 'x'"):  # encoders/base.py:25
        clipcap.get_encoder("x", "y")
    with pytest.raises(ValueError):
        clipcap.get_encoder("clip", "RN50")
    from clipcap_b200.model import GPT2LM, TransformerMapper
    with pytest.raises(ValueError):
        GPT2LM("llama-7b")
    with pytest.raises(ValueError):
        TransformerMapper(64, 100, 5, 3, num_heads=3)


def test_product_package_never_imports_the_oracle():
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "clipcap_b200")
    for dp, _, fs in os.walk(root):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dp, f)


def test_training_surface_matches_reference():
    """ClipCapModelPrefixOnly's training-side surface (clipcap/model/model.py:60-123): parameters() = the mapper's,
    train() keeps the language model in eval mode, configure_optimizers() needs a TrainingConfig (same assertion text) and
    returns the Lightning dict; on a CPU model training_step refuses to run (no CPU path)."""
    import pytest
    import torch
    from clipcap_b200.encoders.config import EncoderConfig
    from clipcap_b200.model import ClipCapModel, ClipCapModelPrefixOnly, Config, TrainingConfig
    cfg = Config(language_model="tiny:128:2:2:1003:64", prefix_length=5, projection_length=3, transformer_layers=2,
                 transformer_attention_heads=2, encoder_config=EncoderConfig(encoder_embedding_size=64))
    model = ClipCapModelPrefixOnly(cfg)
    mapper_ids = {id(p) for p in model.transformer_mapper.parameters()}
    assert {id(p) for p in model.parameters()} == mapper_ids and mapper_ids
    model.train()
    assert model.training and model.transformer_mapper.training and not model.language_model.training
    with pytest.raises(AssertionError, match="set_training_config"):
        model.configure_optimizers()
    model.set_training_config(TrainingConfig(optimizer_lr=1e-4, use_deepspeed_optimisers=True, scheduler_warmup_steps=3,
                                             total_steps=10))
    oc = model.configure_optimizers()
    assert set(oc) == {"optimizer", "lr_scheduler"} and oc["lr_scheduler"]["interval"] == "step"
    assert oc["optimizer"].param_groups[0]["weight_decay"] == 0.0  # deepspeed FusedAdam default; torch AdamW: 0.01
    assert {id(p) for g in oc["optimizer"].param_groups for p in g["params"]} == mapper_ids
    lrs = []
    for _ in range(5):
        lrs.append(oc["lr_scheduler"]["scheduler"].get_last_lr()[0])
        oc["lr_scheduler"]["scheduler"].step()
    assert lrs[0] == 0.0 and abs(lrs[3] - 1e-4) < 1e-12 and lrs[4] < lrs[3]
    tokens = torch.tensor([[4, 9, -1]])
    with pytest.raises(RuntimeError, match="no CPU path"):
        model.training_step((tokens, torch.zeros(1, 64)), 0)
    assert tokens.tolist() == [[4, 9, 0]]  # padding rewritten in place before anything else, like the reference (model.py:104)
    with pytest.raises(NotImplementedError):  # fine-tuning the language model is not built
        ClipCapModel(cfg).training_step((tokens, torch.zeros(1, 64)), 0)


def test_logit_filters_equal_the_reference_functions():
    """clipcap_b200.inference.utils (tensor versions of the decode filters) against clipcap/inference/utils.py itself."""
    import pytest
    import torch
    from oracle import ref_runner as RR
    if not RR.available():
        pytest.skip("needs /root/reference (build container only)")
    RR.import_reference()
    import clipcap.inference.utils as U
    from clipcap_b200.inference import utils as V
    g = torch.Generator().manual_seed(0)
    for k, p in [(0, 0.9), (5, 0.0), (7, 0.5), (0, 0.0), (1, 0.3), (200, 0.99)]:
        x = torch.randn(101, generator=g)
        assert torch.equal(V.top_k_top_p_filtering(x.clone(), k, p), U.top_k_top_p_filtering(x.clone(), k, p)), (k, p)
    x, t = torch.randn(50, generator=g), torch.tensor([3, 7, 7, 11])
    assert torch.equal(V.repetition_penalty_apply(x.clone(), t, 1.3), U.repetition_penalty_apply(x.clone(), t, 1.3))
    assert torch.equal(V.sentence_length_penalty_apply(x.clone(), t, 13, 5, 50, 1.0),
                       U.sentence_length_penalty_apply(x.clone(), t, 13, 5, 50, 1.0))


def test_clap_surface_has_no_cpu_path():
    """get_encoder("clap", ...) returns the reference's (CLAPModel, transform) pair; on a CPU device the forward refuses."""
    from clipcap_b200.encoders import get_encoder
    from clipcap_b200.encoders.clap import CLAPModel, ClapAudioTower
    model, transform = get_encoder("clap", "", normalize_embeddings=True, device="cpu")
    assert isinstance(model, CLAPModel) and model.normalize_embeddings and isinstance(model.model, ClapAudioTower)
    assert sum(p.numel() for p in model.model.clap.parameters()) > 28e6          # HTSAT-tiny + projection
    mel = torch.zeros(1, 4, 1001, 64)
    assert transform(mel) is mel
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(mel)
