"""Runs the UNMODIFIED reference (TheoCoombes/ClipCap under /root/reference) on CPU so the restatement in
oracle/restate.py can be pinned against it, and so tests/golden/make_golden.py can freeze its outputs.

TEST INFRASTRUCTURE. Only usable where /root/reference exists (the build container) — `available()` says so; the GPU box
never has it. What has to be stubbed to import the reference offline:
  * `pytorch_lightning` is not installed: a 10-line stand-in module (LightningModule = nn.Module) is injected.
  * `AutoModelForCausalLM.from_pretrained` needs the network: patched to build GPT2LMHeadModel(GPT2Config(...)).
  * tokenizer files need the network: FakeTokenizer returns token ids as text.
Nothing under /root/reference is modified or copied.
"""
from __future__ import annotations

import os
import sys
import types
from typing import Dict, Optional

import torch

REFERENCE_ROOT = os.environ.get("CLIPCAP_REFERENCE_ROOT", "/root/reference")

GPT2_SIZES = {"gpt2": (768, 12, 12), "gpt2-medium": (1024, 24, 16), "gpt2-large": (1280, 36, 20), "gpt2-xl": (1600, 48, 25)}


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "clipcap"))


class FakeTokenizer:
    """eos -> [stop_token]; '.' -> [13]; decode(ids) -> 'id id id' so the returned caption *is* the token ids."""

    def __init__(self, stop_token: int = 50256):
        self.eos_token = "<|endoftext|>"
        self.stop_token = stop_token

    def encode(self, text):
        if text == self.eos_token:
            return [self.stop_token]
        if text == ".":
            return [13]
        return [int(t) for t in text.split()]

    def decode(self, ids):
        return " ".join(str(int(i)) for i in ids)


_imported = None


def import_reference(lm_cfg_override: Optional[dict] = None):
    """Returns the reference `clipcap` package. `lm_cfg_override` = kwargs for GPT2Config when the config names a model
    that is not a stock GPT-2 size (tiny test models: language_model='tiny:<n_embd>:<n_layer>:<n_head>:<vocab>:<n_pos>')."""
    global _imported
    if _imported is not None:
        return _imported
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(torch.nn.Module):
            def save_hyperparameters(self, *a, **k):
                pass

            def log(self, *a, **k):
                pass

        class Callback:
            pass

        pl.LightningModule, pl.Callback, pl.Trainer = LightningModule, Callback, object
        sys.modules["pytorch_lightning"] = pl
    import transformers
    from transformers import GPT2Config, GPT2LMHeadModel

    def fake_from_pretrained(name, *a, **k):
        if name.startswith("tiny:"):
            n_embd, n_layer, n_head, vocab, n_pos = [int(x) for x in name.split(":")[1:]]
            cfg = GPT2Config(n_embd=n_embd, n_layer=n_layer, n_head=n_head, vocab_size=vocab, n_positions=n_pos)
        else:
            n_embd, n_layer, n_head = GPT2_SIZES[name]
            cfg = GPT2Config(n_embd=n_embd, n_layer=n_layer, n_head=n_head)
        return GPT2LMHeadModel(cfg)

    transformers.AutoModelForCausalLM.from_pretrained = staticmethod(fake_from_pretrained)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import clipcap  # noqa: the reference package, unmodified
    import clipcap.inference.base  # noqa
    _imported = clipcap
    return clipcap


def build_reference_model(language_model: str, E: int, K: int, P: int, H: int, L: int, map_w: Dict[str, torch.Tensor],
                          lm_w: Dict[str, torch.Tensor], windowed: bool = False, window_size: int = 16,
                          use_pos: bool = True):
    """ClipCapModelPrefixOnly(config) from the reference with our seeded weights loaded under its own key names."""
    clipcap = import_reference()
    from clipcap.encoders.config import EncoderConfig
    from clipcap.model.config import Config
    enc = EncoderConfig(encoder_model_name="clip", encoder_model_variant="ViT-L/14", encoder_embedding_size=E,
                        use_windowed_embeddings=windowed, window_size=window_size)
    cfg = Config(language_model=language_model, train_language_model=False, prefix_length=K, projection_length=P,
                 transformer_layers=L, transformer_attention_heads=H, use_positional_embeddings=use_pos,
                 encoder_config=enc)
    model = clipcap.model.ClipCapModelPrefixOnly(cfg)
    sd = {f"transformer_mapper.{k}": v for k, v in map_w.items()}
    sd.update({f"language_model.{k}": v for k, v in lm_w.items()})
    sd["language_model.lm_head.weight"] = lm_w["transformer.wte.weight"]
    missing, unexpected = model.load_state_dict(sd, strict=False)
    missing = [m for m in missing if not m.endswith(".attn.bias") and not m.endswith(".attn.masked_bias")]
    assert not missing and not unexpected, (missing, unexpected)
    return model.eval()


def reference_generate_beam(model, embeds: torch.Tensor, beam_size: int, entry_length: int, temperature: float = 1.0,
                            stop_token: int = 50256):
    """clipcap.inference.base.generate_beam on one image; returns the token ids of the best beam."""
    import_reference()
    from clipcap.inference.base import generate_beam
    out = generate_beam(model, FakeTokenizer(stop_token), embeds, number_to_generate=1, beam_size=beam_size,
                        entry_length=entry_length, temperature=temperature)
    return [int(t) for t in out[0].split()] if out[0] else []


def reference_training_step(model, tokens: torch.Tensor, emb: torch.Tensor):
    """The reference's own ClipCapModelPrefixOnly.training_step + loss.backward() (clipcap/model/model.py:94-123).
    Returns (loss, {mapper parameter name: gradient})."""
    import_reference()
    model.train()
    for p in model.language_model.parameters():
        p.grad = None
    for p in model.transformer_mapper.parameters():
        p.grad = None
    loss = model.training_step((tokens.clone(), emb.clone()), 0)
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in model.transformer_mapper.named_parameters()}
    model.eval()
    return float(loss.detach()), grads


def rank_cycle_pick(p: torch.Tensor, step: int) -> int:
    """Deterministic stand-in for torch.multinomial used on both sides of the sampling parity tests: the token of rank
    (step mod 3) among the kept (non-zero) probabilities, so the sequences are not just the arg-max path."""
    p = p.reshape(-1)
    nz = int((p > 0).sum())
    r = step % max(1, min(3, nz))
    return int(p.topk(r + 1).indices[r])


def reference_generate_sampling(model, embeds: torch.Tensor, mode: str, entry_length: int, text_prefix_tokens=None,
                                deterministic: bool = True, **kw):
    """clipcap.inference.nucleus_sampling.generate_nucleus_sampling (mode='nucleus') or
    clipcap.inference.no_beam.generate_no_beam (mode='sample') on one image, number_to_generate=1. The random draw
    (torch.multinomial) is replaced by rank_cycle_pick when `deterministic`, and the distribution of every step is
    captured. Returns (token ids as the reference returns them, i.e. text prefix first; per-step distributions [V])."""
    import_reference()
    import clipcap.inference.no_beam as NB
    import clipcap.inference.nucleus_sampling as NS
    captured = []
    real = torch.multinomial

    def fake_multinomial(p, num_samples=1, *a, **k):
        captured.append(p.detach().reshape(-1).clone())
        if deterministic:
            return torch.tensor([rank_cycle_pick(p, len(captured) - 1)]).reshape(*p.shape[:-1], 1)
        return real(p, num_samples, *a, **k)

    torch.multinomial = fake_multinomial
    try:
        if mode == "nucleus":
            out = NS.generate_nucleus_sampling(model, FakeTokenizer(), embeds, number_to_generate=1,
                                               text_prefix_tokens=text_prefix_tokens, entry_length=entry_length, **kw)
        else:
            out = NB.generate_no_beam(model, FakeTokenizer(), embeds, number_to_generate=1,
                                      text_prefix_tokens=text_prefix_tokens, entry_length=entry_length, **kw)
    finally:
        torch.multinomial = real
    return ([int(t) for t in out[0].split()] if out[0] else []), captured


def hf_clip_vision(vcfg, vit_w: Dict[str, torch.Tensor]):
    """transformers.CLIPVisionModelWithProjection(quick_gelu) loaded with OpenAI-named weights — the stand-in for the
    un-installed `clip` package. Exposes encode_image so the reference CLIPModel wrapper can hold it."""
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    c = CLIPVisionConfig(hidden_size=vcfg.width, intermediate_size=vcfg.mlp_dim, num_hidden_layers=vcfg.layers,
                         num_attention_heads=vcfg.heads, patch_size=vcfg.patch, image_size=vcfg.image_size,
                         projection_dim=vcfg.out_dim, hidden_act="quick_gelu", layer_norm_eps=vcfg.eps)
    m = CLIPVisionModelWithProjection(c).eval()
    w = vcfg.width
    sd = {
        "vision_model.embeddings.class_embedding": vit_w["visual.class_embedding"],
        "vision_model.embeddings.patch_embedding.weight": vit_w["visual.conv1.weight"],
        "vision_model.embeddings.position_embedding.weight": vit_w["visual.positional_embedding"],
        "vision_model.pre_layrnorm.weight": vit_w["visual.ln_pre.weight"],
        "vision_model.pre_layrnorm.bias": vit_w["visual.ln_pre.bias"],
        "vision_model.post_layernorm.weight": vit_w["visual.ln_post.weight"],
        "vision_model.post_layernorm.bias": vit_w["visual.ln_post.bias"],
        "visual_projection.weight": vit_w["visual.proj"].t().contiguous(),
    }
    for l in range(vcfg.layers):
        s, t = f"visual.transformer.resblocks.{l}.", f"vision_model.encoder.layers.{l}."
        wi, bi = vit_w[s + "attn.in_proj_weight"], vit_w[s + "attn.in_proj_bias"]
        for i, n in enumerate(("q_proj", "k_proj", "v_proj")):
            sd[t + f"self_attn.{n}.weight"] = wi[i * w:(i + 1) * w]
            sd[t + f"self_attn.{n}.bias"] = bi[i * w:(i + 1) * w]
        sd[t + "self_attn.out_proj.weight"] = vit_w[s + "attn.out_proj.weight"]
        sd[t + "self_attn.out_proj.bias"] = vit_w[s + "attn.out_proj.bias"]
        sd[t + "layer_norm1.weight"], sd[t + "layer_norm1.bias"] = vit_w[s + "ln_1.weight"], vit_w[s + "ln_1.bias"]
        sd[t + "layer_norm2.weight"], sd[t + "layer_norm2.bias"] = vit_w[s + "ln_2.weight"], vit_w[s + "ln_2.bias"]
        sd[t + "mlp.fc1.weight"], sd[t + "mlp.fc1.bias"] = vit_w[s + "mlp.c_fc.weight"], vit_w[s + "mlp.c_fc.bias"]
        sd[t + "mlp.fc2.weight"], sd[t + "mlp.fc2.bias"] = vit_w[s + "mlp.c_proj.weight"], vit_w[s + "mlp.c_proj.bias"]
    missing, unexpected = m.load_state_dict(sd, strict=False)
    missing = [k for k in missing if "position_ids" not in k]
    assert not missing and not unexpected, (missing, unexpected)

    class Adaptor(torch.nn.Module):
        def __init__(self, inner):
            super().__init__()
            self.inner = inner

        def encode_image(self, x):
            return self.inner(pixel_values=x).image_embeds

    return Adaptor(m).eval()


def reference_clip_model(vcfg, vit_w, normalize: bool = False):
    """The reference's own CLIPModel wrapper (clipcap/encoders/clip.py:105-129) around the HF stand-in."""
    import_reference()
    from clipcap.encoders.clip import CLIPModel
    return CLIPModel(hf_clip_vision(vcfg, vit_w), normalize_embeddings=normalize).eval()
