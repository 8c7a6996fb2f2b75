"""ClipCapModel / ClipCapModelPrefixOnly with the reference's attribute names and constructor
(clipcap/model/model.py:13-123): `.language_model`, `.transformer_mapper`, `.config`, `forward(tokens, embeddings,
mask)`, `configure_optimizers()` and `training_step(batch, idx)`. Training is implemented for ClipCapModelPrefixOnly
(frozen language model, the reference default `train_language_model=False`): the step's forward, loss and backward run in
libclipcap_b200 (cc_train_step) and surface as an ordinary autograd node, so `loss.backward()` + any optimiser — or a
Lightning-style loop — work as they do for the reference."""
from __future__ import annotations

import warnings
from typing import Callable

import torch
import torch.nn as nn

from clipcap_b200.engine import TrainEngine
from clipcap_b200.model.config import Config, TrainingConfig
from clipcap_b200.model.lm import GPT2LM
from clipcap_b200.model.mapper import TransformerMapper, TransformerMapperWindowed


class IdTokenizer:
    """Offline stand-in when no tokenizer files are cached: captions are returned as space-separated token ids."""
    eos_token = "<|endoftext|>"

    def __init__(self, eos_id: int = 50256):
        self.eos_id = eos_id

    def encode(self, text):
        if text == self.eos_token:
            return [self.eos_id]
        if text == ".":
            return [13]
        return [int(t) for t in text.split()]

    def decode(self, ids):
        return " ".join(str(int(i)) for i in ids)


def get_tokenizer(language_model_name: str, **huggingface_kwargs) -> Callable:
    """model.py:10-11. Uses the HF tokenizer when its files are available locally; there is no network here."""
    if language_model_name.startswith("tiny:"):
        return IdTokenizer(int(language_model_name.split(":")[4]) - 1)
    try:
        from transformers import AutoTokenizer
        return AutoTokenizer.from_pretrained(language_model_name, local_files_only=True, **huggingface_kwargs)
    except Exception as e:  # noqa: BLE001
        warnings.warn(f"tokenizer files for '{language_model_name}' not available offline ({type(e).__name__}); "
                      "captions will be returned as token ids")
        return IdTokenizer()


class ClipCapModel(nn.Module):
    def __init__(self, config: Config):
        super().__init__()
        self.config = config
        self.language_model = GPT2LM(self.config.language_model)
        self.lm_embedding_size = self.language_model.get_input_embeddings().weight.shape[1]
        enc = self.config.encoder_config
        if enc.use_windowed_embeddings:
            self.transformer_mapper = TransformerMapperWindowed(
                encoder_embedding_size=enc.encoder_embedding_size, lm_embedding_size=self.lm_embedding_size,
                prefix_length=self.config.prefix_length, projection_length=self.config.projection_length,
                window_size=(enc.window_size + 1), use_pos_embeddings=self.config.use_positional_embeddings,
                num_heads=self.config.transformer_attention_heads, num_layers=self.config.transformer_layers)
        else:
            self.transformer_mapper = TransformerMapper(
                encoder_embedding_size=enc.encoder_embedding_size, lm_embedding_size=self.lm_embedding_size,
                prefix_length=self.config.prefix_length, projection_length=self.config.projection_length,
                num_heads=self.config.transformer_attention_heads, num_layers=self.config.transformer_layers)

    @torch.no_grad()
    def forward(self, tokens: torch.Tensor, embeddings: torch.Tensor, mask: torch.Tensor):
        """model.py:43-58 — teacher-forced logits over [prefix, tokens]."""
        token_embeddings = self.language_model.get_input_embeddings()(tokens)
        prefix_projections = self.transformer_mapper(embeddings)
        inputs_embeds = torch.cat((prefix_projections.to(token_embeddings.dtype), token_embeddings), dim=1)
        prefix_mask = torch.ones(prefix_projections.shape[:-1], dtype=torch.bool, device=mask.device)
        mask = torch.cat((prefix_mask, mask), dim=1)
        return self.language_model(inputs_embeds=inputs_embeds, attention_mask=mask)

    def set_training_config(self, training_config: TrainingConfig, reinit_optims: bool = False) -> None:
        """model.py:60-65."""
        self.config.training_config = training_config
        if reinit_optims:
            self.configure_optimizers()

    def configure_optimizers(self) -> dict:
        """model.py:67-91: AdamW + linear warm-up schedule, stepped every optimiser step."""
        from clipcap_b200.model.optim import FusedAdamW, linear_schedule_with_warmup
        tc = self.config.training_config
        assert tc is not None, "You must first use `set_training_config` before training."
        # deepspeed FusedAdam(adam_w_mode=True) defaults to weight_decay 0, torch.optim.AdamW to 0.01 (model.py:72-77)
        optimizer = FusedAdamW(self.parameters(), lr=tc.optimizer_lr,
                               weight_decay=0.0 if tc.use_deepspeed_optimisers else 0.01)
        scheduler = linear_schedule_with_warmup(optimizer, tc.scheduler_warmup_steps, tc.total_steps)
        return {"optimizer": optimizer, "lr_scheduler": {"scheduler": scheduler, "interval": "step", "frequency": 1}}

    def training_step(self, batch, _=None):
        raise NotImplementedError(
            "clipcap_b200 trains the prefix mapper with the language model frozen (ClipCapModelPrefixOnly, "
            "train_language_model=False); fine-tuning the language model is not implemented")


class _PrefixOnlyStep(torch.autograd.Function):
    """One cc_train_step: the loss comes back as a tensor, the parameter gradients computed by the same call are handed
    to autograd in backward (scaled by the incoming gradient, 1 for a plain loss.backward())."""

    @staticmethod
    def forward(ctx, engine, emb, tokens, loss_scale, names, *params):
        need = any(ctx.needs_input_grad[5:])  # False under torch.no_grad() (validation): forward + loss only
        table = {n: p.detach() for n, p in zip(names, params)}
        grads = {n: torch.empty_like(p, memory_format=torch.contiguous_format) for n, p in table.items()} if need else None
        loss = engine.step(table, emb, tokens, grads, loss_scale)
        ctx.grads = None if grads is None else [grads[n] for n in names]
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.grads is None:
            raise RuntimeError("training step ran without gradients enabled")
        return (None, None, None, None, None) + tuple(g * grad_out for g in ctx.grads)


class ClipCapModelPrefixOnly(ClipCapModel):
    loss_scale = 1024.0  # static scale of the fp16 activation gradients inside cc_train_step (parameter grads unscaled)

    def parameters(self, recurse: bool = True):
        return self.transformer_mapper.parameters()

    def train(self, mode: bool = True):
        """model.py:120-123: the language model stays in eval mode."""
        super().train(mode)
        self.language_model.eval()
        return self

    def _train_engine_for(self, batch: int, n_tokens: int) -> TrainEngine:
        lm, mp = self.language_model, self.transformer_mapper
        dev = next(mp.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError(f"clipcap_b200: model is on {dev}; move it to a B200 (`.to('cuda')`) — there is no CPU path")
        key = tuple((p.data_ptr(), p._version) for p in lm.parameters())
        eng = getattr(self, "_train_engine", None)
        cap = getattr(self, "_train_cap", (0, 0))
        if eng is not None and key == self._train_key and batch <= cap[0] and n_tokens <= cap[1]:
            return eng
        if eng is not None:
            eng.close()
        cap = (max(batch, cap[0]), max(n_tokens, cap[1]))
        eng = TrainEngine(lm.state_dict(), E=mp.encoder_embedding_size, d=mp.lm_embedding_size, P=mp.projection_length,
                          K=mp.prefix_length, H=mp.num_heads, L=mp.num_layers, lm_layers=lm.n_layer, lm_heads=lm.n_head,
                          V=lm.vocab_size, n_pos=lm.n_positions, max_batch=cap[0], max_tokens=cap[1], device=dev,
                          kind=mp.kind, W=mp.window_size, use_pos=mp.use_pos)
        object.__setattr__(self, "_train_engine", eng)
        object.__setattr__(self, "_train_key", key)
        object.__setattr__(self, "_train_cap", cap)
        return eng

    def training_step(self, batch, _=None) -> torch.Tensor:
        """model.py:94-113. `batch` = (tokens [B, Tt] int64 with -1 padding, embeddings [B, E]). Returns the loss; its
        backward fills `.grad` of every transformer_mapper parameter."""
        tokens, embeds = batch
        mask = tokens.ge(0)           # model.py:103
        tokens[~mask] = 0             # model.py:104 (in place, like the reference)
        eng = self._train_engine_for(tokens.shape[0], tokens.shape[1])
        named = [(n, p) for n, p in self.transformer_mapper.named_parameters()]
        for n, p in named:
            if not p.is_contiguous():
                raise RuntimeError(f"parameter transformer_mapper.{n} must be contiguous")
        loss = _PrefixOnlyStep.apply(eng, embeds.float(), tokens, float(self.loss_scale), [n for n, _ in named],
                                     *[p for _, p in named])
        self.log("loss", loss)
        return loss

    def check_overflow(self) -> bool:
        """Dynamic loss scaling hook (call it every few steps; it synchronises): True if the last training step's gradients
        held non-finite values — the static scale of the fp16 activation gradients overflowed. The scale is then halved for
        the following steps; the optimiser has skipped the offending elements (cc_op_adamw), its state is intact."""
        eng = getattr(self, "_train_engine", None)
        if eng is None or eng.last_nonfinite() <= 0:
            return False
        self.loss_scale = max(1.0, float(self.loss_scale) / 2.0)
        return True

    def log(self, name: str, value) -> None:  # Lightning's self.log; keeps the last value without a host sync
        object.__setattr__(self, "_last_logged", (name, value.detach() if torch.is_tensor(value) else value))
