"""Thin Python owners of the C-ABI handles (include/clipcap_b200.h). Tensors in, tensors out; all arithmetic happens in
libclipcap_b200.so on the caller's current CUDA stream. No fallbacks: a missing library or a non-sm_100 device raises."""
from __future__ import annotations

import contextlib
import ctypes as C
import sys
from typing import Dict, Iterable, Optional, Tuple

import torch

from . import _ffi


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"clipcap_b200: {what} must be a CUDA tensor (there is no CPU path); got device {t.device}")


def _named(weights: Dict[str, torch.Tensor], device) -> Iterable[Tuple[str, torch.Tensor]]:
    return [(k, v.detach().to(device=device, dtype=torch.float32).contiguous()) for k, v in weights.items()]


class _Handle:
    _destroy_name = ""

    def __init__(self):
        self._h = C.c_void_p()

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            getattr(_ffi.lib(), self._destroy_name)(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        # At interpreter shutdown objects die in no particular order: a green context destroyed before the torch streams
        # and events that live in it makes torch's own destructors throw ("context is destroyed") and abort the process.
        # The driver reclaims everything at exit anyway, so handles are only destroyed while the interpreter is alive.
        try:
            if sys.is_finalizing():
                return
            self.close()
        except Exception:
            pass


class SmPartition(_Handle):
    """cc_partition_* — the device's SMs split into a large partition (index 0: image tower, mapper, prefill) and a small
    one (index 1: the decode loop), each with its own stream (CUDA green contexts). `with part.on(i):` makes partition i's
    stream the current torch stream and sizes the kernels enqueued inside for its SM count (cc_set_sm_budget)."""
    _destroy_name = "cc_partition_destroy"

    def __init__(self, small_sms: int = 24, device="cuda"):
        super().__init__()
        self.device = torch.device(device)
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _ffi.check(_ffi.lib().cc_partition_create(C.byref(self._h), index, int(small_sms)))
        lib = _ffi.lib()
        self.sms = [lib.cc_partition_sms(self._h, i) for i in (0, 1)]
        self.streams = [torch.cuda.ExternalStream(lib.cc_partition_stream(self._h, i), device=self.device) for i in (0, 1)]

    @contextlib.contextmanager
    def on(self, which: int):
        lib = _ffi.lib()
        before = lib.cc_get_sm_budget()
        lib.cc_set_sm_budget(self.sms[which])
        try:
            with torch.cuda.stream(self.streams[which]):
                yield self.streams[which]
        finally:
            lib.cc_set_sm_budget(before)


class VitEngine(_Handle):
    """cc_vit_* — CLIP ViT image tower. `weights`: OpenAI-clip-named fp32 tensors (`visual.*`)."""
    _destroy_name = "cc_vit_destroy"

    def __init__(self, weights: Dict[str, torch.Tensor], image_size=224, patch=14, width=1024, layers=24, heads=16,
                 mlp_dim=4096, out_dim=768, eps=1e-5, max_batch=256, device="cuda"):
        super().__init__()
        self.device = torch.device(device)
        self.cfg = _ffi.cc_vit_cfg(image_size, patch, width, layers, heads, mlp_dim, out_dim, eps)
        self.max_batch = max_batch
        with torch.cuda.device(self.device):
            arr, keep = _ffi.make_tensor_table(_named(weights, self.device))
            _ffi.check(_ffi.lib().cc_vit_create(C.byref(self._h), C.byref(self.cfg), arr, len(keep), max_batch))
            torch.cuda.synchronize()

    def forward(self, pixels: torch.Tensor, normalize: bool = False, out_dtype: Optional[torch.dtype] = None):
        _require_cuda(pixels, "pixels")
        s = self.cfg.image_size
        if pixels.dim() != 4 or tuple(pixels.shape[1:]) != (3, s, s):
            raise ValueError(f"pixels must be [B, 3, {s}, {s}], got {tuple(pixels.shape)}")
        pixels = pixels.contiguous()
        out = torch.empty(pixels.shape[0], self.cfg.out_dim, device=pixels.device, dtype=out_dtype or pixels.dtype)
        with torch.cuda.device(pixels.device):
            _ffi.check(_ffi.lib().cc_vit_forward(self._h, pixels.data_ptr(), _ffi.torch_dtype_code(pixels),
                                                 pixels.shape[0], int(bool(normalize)), out.data_ptr(),
                                                 _ffi.torch_dtype_code(out), _ffi.current_stream_ptr()))
        return out

    @property
    def last_launches(self) -> int:
        return _ffi.lib().cc_vit_last_launches(self._h)


class ClapEngine(_Handle):
    """cc_clap_* — CLAP audio tower (HTSAT Swin encoder + audio projection). `weights`: tensors under the state_dict keys
    of transformers' ClapAudioModelWithProjection (`audio_model.audio_encoder.*`, `audio_projection.*`)."""
    _destroy_name = "cc_clap_destroy"

    def __init__(self, weights: Dict[str, torch.Tensor], num_mel_bins=64, spec_size=256, patch=4, embed=96,
                 depths=(2, 2, 6, 2), heads=(4, 8, 16, 32), window=8, projection_dim=512, eps=1e-5, max_batch=64,
                 device="cuda"):
        super().__init__()
        if len(depths) != 4 or len(heads) != 4:
            raise ValueError("the CLAP audio tower has 4 stages")
        self.device = torch.device(device)
        self.cfg = _ffi.cc_clap_cfg(num_mel_bins, spec_size, patch, embed, (C.c_int32 * 4)(*depths),
                                    (C.c_int32 * 4)(*heads), window, projection_dim, eps)
        self.max_batch = max_batch
        self.stage_shapes = [((spec_size // patch >> i) ** 2, embed << i) for i in range(4)]  # (tokens, channels)
        wanted = ("audio_model.audio_encoder.", "audio_projection.")
        skip = ("relative_position_index", "num_batches_tracked")
        weights = {k: v for k, v in weights.items() if k.startswith(wanted) and not any(x in k for x in skip)}
        with torch.cuda.device(self.device):
            arr, keep = _ffi.make_tensor_table(_named(weights, self.device))
            _ffi.check(_ffi.lib().cc_clap_create(C.byref(self._h), C.byref(self.cfg), arr, len(keep), max_batch))
            torch.cuda.synchronize()

    def forward(self, mel: torch.Tensor, normalize: bool = False, out_dtype: Optional[torch.dtype] = None,
                stop_after_stage: int = -1, is_longer=None):
        """mel: [B, channels, T, num_mel_bins] fp32 / fp16 (channel 0 = the global view, 1..3 = local views of clips longer
        than the window). `is_longer`: per-sample flags (tensor / sequence, any shape with B elements) selecting the
        feature-fusion patch embedding. Returns [B, projection_dim]; with `stop_after_stage` = s >= 0, the fp32 token stream
        [B, tokens_s, channels_s] after stage s instead."""
        _require_cuda(mel, "mel")
        if mel.dim() != 4 or mel.shape[3] != self.cfg.num_mel_bins:
            raise ValueError(f"mel must be [B, channels, T, {self.cfg.num_mel_bins}], got {tuple(mel.shape)}")
        if mel.dtype not in (torch.float32, torch.float16):
            mel = mel.float()
        mel = mel.contiguous()
        B = mel.shape[0]
        flags = None
        if is_longer is not None:
            host = torch.as_tensor(is_longer).reshape(-1).to("cpu")
            if host.numel() != B:
                raise ValueError(f"is_longer must hold one flag per sample ({B}), got {host.numel()}")
            if bool(host.any()):
                flags = (C.c_ubyte * B)(*[1 if v else 0 for v in host.tolist()])
        out = torch.empty(B, self.cfg.projection_dim, device=mel.device, dtype=out_dtype or mel.dtype)
        dump = None
        if stop_after_stage >= 0:
            tokens, ch = self.stage_shapes[stop_after_stage]
            dump = torch.empty(B, tokens, ch, device=mel.device, dtype=torch.float32)
        with torch.cuda.device(mel.device):
            _ffi.check(_ffi.lib().cc_clap_forward(self._h, mel.data_ptr(), _ffi.torch_dtype_code(mel),
                                                  C.cast(flags, C.c_void_p) if flags is not None else None, B, mel.shape[1],
                                                  mel.shape[2], int(bool(normalize)), out.data_ptr(),
                                                  _ffi.torch_dtype_code(out), stop_after_stage,
                                                  dump.data_ptr() if dump is not None else None,
                                                  _ffi.current_stream_ptr()))
        return dump if dump is not None else out

    @property
    def last_launches(self) -> int:
        return _ffi.lib().cc_clap_last_launches(self._h)


class MapperEngine(_Handle):
    """cc_mapper_* — TransformerMapper / TransformerMapperWindowed / MLP mapper. Keys relative to `transformer_mapper.`"""
    _destroy_name = "cc_mapper_destroy"
    KINDS = {"transformer": _ffi.CC_MAPPER_TRANSFORMER, "windowed": _ffi.CC_MAPPER_WINDOWED, "mlp": _ffi.CC_MAPPER_MLP}

    def __init__(self, weights: Dict[str, torch.Tensor], kind="transformer", E=768, d=1024, P=10, K=40, H=8, L=8, W=1,
                 use_pos=False, eps=1e-5, max_batch=256, device="cuda"):
        super().__init__()
        self.device = torch.device(device)
        self.kind = kind
        self.cfg = _ffi.cc_mapper_cfg(self.KINDS[kind], E, d, P, K, H, L, W, int(bool(use_pos)), eps)
        self.max_batch = max_batch
        with torch.cuda.device(self.device):
            arr, keep = _ffi.make_tensor_table(_named(weights, self.device))
            _ffi.check(_ffi.lib().cc_mapper_create(C.byref(self._h), C.byref(self.cfg), arr, len(keep), max_batch))
            torch.cuda.synchronize()

    def forward(self, emb: torch.Tensor, out_dtype: Optional[torch.dtype] = None,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """`out`: optional contiguous [B, K, d] fp32 / fp16 CUDA tensor to write into (e.g. this rank's slot of the
        all-gathered prefix buffer), else a new tensor of `out_dtype` (default: the embeddings' dtype)."""
        _require_cuda(emb, "embeddings")
        c = self.cfg
        want = (c.W, c.E) if self.kind == "windowed" else (c.E,)
        if tuple(emb.shape[1:]) != want:
            raise ValueError(f"embeddings must be [B, {', '.join(map(str, want))}], got {tuple(emb.shape)}")
        emb = emb.contiguous()
        if out is None:
            out = torch.empty(emb.shape[0], c.K, c.d, device=emb.device, dtype=out_dtype or emb.dtype)
        elif (tuple(out.shape) != (emb.shape[0], c.K, c.d) or not out.is_contiguous() or not out.is_cuda
              or out.dtype not in (torch.float32, torch.float16)):
            raise ValueError(f"out must be a contiguous CUDA fp32 / fp16 tensor of shape {(emb.shape[0], c.K, c.d)}")
        with torch.cuda.device(emb.device):
            _ffi.check(_ffi.lib().cc_mapper_forward(self._h, emb.data_ptr(), _ffi.torch_dtype_code(emb), emb.shape[0],
                                                    out.data_ptr(), _ffi.torch_dtype_code(out),
                                                    _ffi.current_stream_ptr()))
        return out

    @property
    def last_launches(self) -> int:
        return _ffi.lib().cc_mapper_last_launches(self._h)


class Gpt2Engine(_Handle):
    """cc_gpt2_* / cc_generate — GPT-2 forward, embedding lookup and the decode loops. Keys relative to
    `language_model.` (HF GPT2LMHeadModel names)."""
    _destroy_name = "cc_gpt2_destroy"

    def __init__(self, weights: Dict[str, torch.Tensor], d=1024, L=24, H=16, V=50257, n_pos=1024, eps=1e-5,
                 max_seqs=256, max_len=64, device="cuda"):
        super().__init__()
        self.device = torch.device(device)
        self.cfg = _ffi.cc_gpt2_cfg(d, L, H, V, n_pos, eps)
        self.max_seqs, self.max_len = max_seqs, max_len
        with torch.cuda.device(self.device):
            arr, keep = _ffi.make_tensor_table(_named(weights, self.device))
            _ffi.check(_ffi.lib().cc_gpt2_create(C.byref(self._h), C.byref(self.cfg), arr, len(keep), max_seqs, max_len))
            torch.cuda.synchronize()

    def logits(self, embeds: torch.Tensor, all_positions: bool = False) -> torch.Tensor:
        _require_cuda(embeds, "inputs_embeds")
        if embeds.dim() != 3 or embeds.shape[2] != self.cfg.d:
            raise ValueError(f"inputs_embeds must be [B, T, {self.cfg.d}], got {tuple(embeds.shape)}")
        embeds = embeds.contiguous()
        B, T, _ = embeds.shape
        shape = (B, T, self.cfg.V) if all_positions else (B, self.cfg.V)
        out = torch.empty(shape, device=embeds.device, dtype=torch.float32)
        with torch.cuda.device(embeds.device):
            _ffi.check(_ffi.lib().cc_gpt2_logits(self._h, embeds.data_ptr(), _ffi.torch_dtype_code(embeds), B, T,
                                                 int(all_positions), out.data_ptr(), _ffi.current_stream_ptr()))
        return out

    def embed(self, ids: torch.Tensor, dtype: torch.dtype = torch.float32) -> torch.Tensor:
        _require_cuda(ids, "token ids")
        flat = ids.reshape(-1).to(torch.int32).contiguous()
        out = torch.empty(flat.numel(), self.cfg.d, device=ids.device, dtype=dtype)
        if flat.numel():
            with torch.cuda.device(ids.device):
                _ffi.check(_ffi.lib().cc_gpt2_embed(self._h, flat.data_ptr(), flat.numel(), out.data_ptr(),
                                                    _ffi.torch_dtype_code(out), _ffi.current_stream_ptr()))
        return out.view(*ids.shape, self.cfg.d)

    _MODES = {"greedy": _ffi.CC_GEN_GREEDY, "beam": _ffi.CC_GEN_BEAM, "nucleus": _ffi.CC_GEN_NUCLEUS,
              "sample": _ffi.CC_GEN_SAMPLE}

    def generate(self, prefix: torch.Tensor, mode: str = "greedy", beam: int = 1, entry_length: int = 67,
                 temperature: float = 1.0, stop_token: int = 50256, top_p: float = 1.0, top_k: int = 0,
                 repetition_penalty: float = 1.0, desired_sentence_length: int = 50,
                 sentence_length_factor: float = 1.0, history=None, seed: int = 0):
        """-> (tokens int32 [B, entry_length], lengths int32 [B], scores fp32 [B]) on the device, no host sync.
        mode: "greedy" | "beam" (generate_beam), "nucleus" (generate_nucleus_sampling), "sample" (generate_no_beam);
        `history` = text-prefix token ids that start every row's repetition-penalty history ("sample" only)."""
        if mode not in self._MODES:
            raise ValueError(f"unknown decode mode '{mode}'")
        _require_cuda(prefix, "prefix embeddings")
        if prefix.dim() != 3 or prefix.shape[2] != self.cfg.d:
            raise ValueError(f"prefix must be [B, Tp, {self.cfg.d}], got {tuple(prefix.shape)}")
        prefix = prefix.contiguous()
        B, Tp, _ = prefix.shape
        hist = None
        if history is not None and len(history) > 0:
            hist = (C.c_int32 * len(history))(*[int(t) for t in history])
        g = _ffi.cc_gen_cfg(self._MODES[mode], beam, entry_length, float(temperature), stop_token,
                            float(1.0 if top_p is None else top_p), int(top_k), float(repetition_penalty),
                            int(desired_sentence_length), float(sentence_length_factor),
                            0 if hist is None else len(hist), None if hist is None else C.addressof(hist),
                            int(seed) & 0xFFFFFFFFFFFFFFFF)
        tokens = torch.empty(B, entry_length, device=prefix.device, dtype=torch.int32)
        lengths = torch.empty(B, device=prefix.device, dtype=torch.int32)
        scores = torch.empty(B, device=prefix.device, dtype=torch.float32)
        with torch.cuda.device(prefix.device):
            _ffi.check(_ffi.lib().cc_generate(self._h, prefix.data_ptr(), _ffi.torch_dtype_code(prefix), B, Tp,
                                              C.byref(g), tokens.data_ptr(), lengths.data_ptr(), scores.data_ptr(),
                                              _ffi.current_stream_ptr()))
        return tokens, lengths, scores

    def _greedy_or_beam_cfg(self, mode, beam, entry_length, temperature, stop_token):
        if mode not in ("greedy", "beam"):
            raise ValueError("the two-phase generate supports the deterministic modes 'greedy' and 'beam'")
        return _ffi.cc_gen_cfg(self._MODES[mode], beam, entry_length, float(temperature), stop_token, 1.0, 0, 1.0, 50, 1.0,
                               0, None, 0)

    def prefill(self, prefix: torch.Tensor, mode: str = "greedy", beam: int = 1, entry_length: int = 67,
                temperature: float = 1.0, stop_token: int = 50256) -> Tuple[int, int]:
        """cc_generate_prefill on the current stream: prefix -> KV cache + first token (handle state). Returns (B, Tp) to
        hand to `decode`, which may run on another stream once it is ordered after this call."""
        _require_cuda(prefix, "prefix embeddings")
        if prefix.dim() != 3 or prefix.shape[2] != self.cfg.d:
            raise ValueError(f"prefix must be [B, Tp, {self.cfg.d}], got {tuple(prefix.shape)}")
        prefix = prefix.contiguous()
        B, Tp, _ = prefix.shape
        g = self._greedy_or_beam_cfg(mode, beam, entry_length, temperature, stop_token)
        with torch.cuda.device(prefix.device):
            _ffi.check(_ffi.lib().cc_generate_prefill(self._h, prefix.data_ptr(), _ffi.torch_dtype_code(prefix), B, Tp,
                                                      C.byref(g), _ffi.current_stream_ptr()))
        return B, Tp

    def decode(self, B: int, Tp: int, mode: str = "greedy", beam: int = 1, entry_length: int = 67,
               temperature: float = 1.0, stop_token: int = 50256):
        """cc_generate_decode on the current stream: the remaining entry_length - 1 steps of the call `prefill` started.
        -> (tokens, lengths, scores) as `generate`."""
        g = self._greedy_or_beam_cfg(mode, beam, entry_length, temperature, stop_token)
        tokens = torch.empty(B, entry_length, device=self.device, dtype=torch.int32)
        lengths = torch.empty(B, device=self.device, dtype=torch.int32)
        scores = torch.empty(B, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _ffi.check(_ffi.lib().cc_generate_decode(self._h, B, Tp, C.byref(g), tokens.data_ptr(), lengths.data_ptr(),
                                                     scores.data_ptr(), _ffi.current_stream_ptr()))
        return tokens, lengths, scores

    def set_prefill_defer(self, blocks: int) -> None:
        """cc_gpt2_set_prefill_defer: the last `blocks` prefill blocks + the first token run at the head of `decode`."""
        _ffi.check(_ffi.lib().cc_gpt2_set_prefill_defer(self._h, int(blocks)))

    @property
    def last_launches(self) -> int:
        return _ffi.lib().cc_gpt2_last_launches(self._h)


class TrainEngine(_Handle):
    """cc_train_* — one training step of ClipCapModelPrefixOnly (clipcap/model/model.py:94-123): forward, cross-entropy
    loss with ignore_index 0, and the gradients of every transformer_mapper parameter; the language model is frozen.
    `lm_weights`: `language_model.*` tensors (HF GPT-2 names)."""
    _destroy_name = "cc_train_destroy"

    def __init__(self, lm_weights: Dict[str, torch.Tensor], E=768, d=1024, P=10, K=40, H=8, L=8, lm_layers=24,
                 lm_heads=16, V=50257, n_pos=1024, eps=1e-5, max_batch=64, max_tokens=67, device="cuda",
                 kind="transformer", W=1, use_pos=False):
        """kind "windowed" (TransformerMapperWindowed, mapper.py:133-160): embeddings are [B, W, E] (W = window_size + 1)
        and `pos_embeddings` joins the parameters / gradients when `use_pos`."""
        super().__init__()
        self.device = torch.device(device)
        self.kind = kind
        self.mcfg = _ffi.cc_mapper_cfg(MapperEngine.KINDS[kind], E, d, P, K, H, L, W, int(bool(use_pos)), eps)
        self.gcfg = _ffi.cc_gpt2_cfg(d, lm_layers, lm_heads, V, n_pos, eps)
        self.max_batch, self.max_tokens = max_batch, max_tokens
        with torch.cuda.device(self.device):
            arr, keep = _ffi.make_tensor_table(_named(lm_weights, self.device))
            _ffi.check(_ffi.lib().cc_train_create(C.byref(self._h), C.byref(self.mcfg), C.byref(self.gcfg), arr, len(keep),
                                                  max_batch, max_tokens))
            torch.cuda.synchronize()

    def step(self, params: Dict[str, torch.Tensor], emb: torch.Tensor, tokens: torch.Tensor,
             grads: Optional[Dict[str, torch.Tensor]] = None, loss_scale: float = 1024.0) -> torch.Tensor:
        """params / grads: name -> fp32 contiguous CUDA tensor (names relative to `transformer_mapper.`). tokens: [B, Tt]
        integer ids, negative = padding. Returns the loss as a 0-dim fp32 CUDA tensor; `grads` (if given) are overwritten
        with d loss / d param. No host synchronisation."""
        _require_cuda(emb, "embeddings")
        _require_cuda(tokens, "tokens")
        want = (self.mcfg.W, self.mcfg.E) if self.kind == "windowed" else (self.mcfg.E,)
        if tuple(emb.shape[1:]) != want:
            raise ValueError(f"embeddings must be [B, {', '.join(map(str, want))}], got {tuple(emb.shape)}")
        if tokens.dim() != 2 or tokens.shape[0] != emb.shape[0]:
            raise ValueError(f"tokens must be [B, Tt] with B = {emb.shape[0]}, got {tuple(tokens.shape)}")
        for name, t in list(params.items()) + list((grads or {}).items()):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise ValueError(f"'{name}' must be a contiguous fp32 CUDA tensor")
        emb = emb.contiguous()
        tok32 = tokens.to(torch.int32).contiguous()
        loss = torch.empty((), device=emb.device, dtype=torch.float32)
        with torch.cuda.device(emb.device):
            parr, pkeep = _ffi.make_tensor_table(list(params.items()))
            if grads is not None:
                garr, gkeep = _ffi.make_tensor_table(list(grads.items()))
                ng = len(gkeep)
            else:
                garr, ng = None, 0
            _ffi.check(_ffi.lib().cc_train_step(self._h, parr, len(pkeep), garr, ng, emb.data_ptr(),
                                                _ffi.torch_dtype_code(emb), tok32.data_ptr(), emb.shape[0],
                                                tokens.shape[1], float(loss_scale), loss.data_ptr(),
                                                _ffi.current_stream_ptr()))
        return loss

    def last_nonfinite(self) -> int:
        """Non-finite gradient elements of the last step with gradients (host sync): > 0 means the loss scale overflowed."""
        with torch.cuda.device(self.device):
            return _ffi.lib().cc_train_last_nonfinite(self._h, _ffi.current_stream_ptr())

    @property
    def last_launches(self) -> int:
        return _ffi.lib().cc_train_last_launches(self._h)


def adamw_update(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, lr: float, beta1=0.9, beta2=0.999,
                 eps=1e-8, weight_decay=0.01, step=1) -> None:
    """cc_op_adamw — torch.optim.AdamW's update on one flat fp32 CUDA parameter (in place)."""
    for t in (p, g, m, v):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError("adamw_update: contiguous fp32 CUDA tensors only")
    with torch.cuda.device(p.device):
        _ffi.check(_ffi.lib().cc_op_adamw(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), float(lr),
                                          float(beta1), float(beta2), float(eps), float(weight_decay), int(step),
                                          _ffi.current_stream_ptr()))
