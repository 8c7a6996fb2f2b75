"""CPU restatement of the CLAP audio tower (BASELINE configs[4]) in plain fp32 torch ops.

TEST INFRASTRUCTURE: the checker of clipcap_b200's CUDA path for CLAP (`csrc/clap.cu`, `tests/test_clap_gpu.py`); nothing
in the product imports it. PARITY UNPINNED by the reference: its CLAP wrapper does not run as
committed (`clipcap/encoders/clap.py:136,152` — undefined names) and its arithmetic lives in `laion_clap` (unpinned,
`requirements-clap.txt:1`, not installed). The stand-in named by SURVEY §8c is
`transformers.ClapAudioModelWithProjection(ClapAudioConfig(enable_fusion=True))` (5.5.0 installed): HTSAT-tiny, a Swin
transformer over the mel spectrogram folded into a 256 x 256 image, 28.2 M parameters, [B, 4, 1001, 64] -> [B, 512].
Every step below cites the lines of `transformers/models/clap/modeling_clap.py` it follows; `tests/test_clap_oracle_cpu.py`
pins the restatement against that module on seeded weights.

Weights: a flat dict keyed by the HF state_dict names (`audio_model.audio_encoder.*`, `audio_projection.*`).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

Weights = Dict[str, torch.Tensor]
ENC = "audio_model.audio_encoder."


@dataclass
class ClapCfg:
    num_mel_bins: int = 64
    spec_size: int = 256
    patch: int = 4                      # patch_size == patch_stride
    embed: int = 96                     # patch_embeds_hidden_size; stage i has embed * 2^i channels
    depths: Tuple[int, ...] = (2, 2, 6, 2)
    heads: Tuple[int, ...] = (4, 8, 16, 32)
    window: int = 8
    mlp_ratio: float = 4.0
    enable_fusion: bool = True
    aff_r: int = 4
    projection_dim: int = 512
    eps: float = 1e-5

    @property
    def hidden(self) -> int:
        return self.embed * 2 ** (len(self.depths) - 1)


def _bn_eval(x, w, p, eps=1e-5):
    """BatchNorm2d in eval mode over dim 1 with the running statistics stored under prefix p."""
    shape = (1, -1) + (1,) * (x.dim() - 2)
    scale = w[p + "weight"] / torch.sqrt(w[p + "running_var"] + eps)
    return (x - w[p + "running_mean"].view(shape)) * scale.view(shape) + w[p + "bias"].view(shape)


def _ln(x, w, p, eps):
    return F.layer_norm(x, (x.shape[-1],), w[p + "weight"], w[p + "bias"], eps)


def reshape_mel2img(x: torch.Tensor, cfg: ClapCfg) -> torch.Tensor:
    """[B, C, T, F] -> [B, C, 256, 256]  (modeling_clap.py:777-812): bicubic stretch of the time axis to
    spec_size * freq_ratio, then the time axis is cut into freq_ratio pieces that are stacked along frequency."""
    ratio = cfg.spec_size // cfg.num_mel_bins
    spec_w, spec_h = cfg.spec_size * ratio, cfg.spec_size // ratio
    _, _, t, f = x.shape
    assert t <= spec_w and f <= spec_h
    if t < spec_w:
        x = F.interpolate(x, (spec_w, f), mode="bicubic", align_corners=True)
    if f < spec_h:
        x = F.interpolate(x, (x.shape[2], spec_h), mode="bicubic", align_corners=True)
    b, c, t, f = x.shape
    x = x.reshape(b, c * ratio, t // ratio, f).permute(0, 1, 3, 2).contiguous()
    return x.reshape(b, c, f * ratio, t // ratio)


def _aff(w: Weights, p: str, x: torch.Tensor, residual: torch.Tensor) -> torch.Tensor:
    """ClapAudioAFFBlock.forward (modeling_clap.py:238-245): attentional feature fusion of the global and local maps."""
    xa = x + residual
    # local_att: conv1x1 -> BN -> ReLU -> conv1x1 -> BN            (indices 0,1,2,3,4 of the Sequential)
    l = F.conv2d(xa, w[p + "local_att.0.weight"], w[p + "local_att.0.bias"])
    l = torch.relu(_bn_eval(l, w, p + "local_att.1."))
    l = _bn_eval(F.conv2d(l, w[p + "local_att.3.weight"], w[p + "local_att.3.bias"]), w, p + "local_att.4.")
    # global_att: global average pool -> conv1x1 -> BN -> ReLU -> conv1x1 -> BN   (indices 0..5)
    g = xa.mean(dim=(2, 3), keepdim=True)
    g = F.conv2d(g, w[p + "global_att.1.weight"], w[p + "global_att.1.bias"])
    g = torch.relu(_bn_eval(g, w, p + "global_att.2."))
    g = _bn_eval(F.conv2d(g, w[p + "global_att.4.weight"], w[p + "global_att.4.bias"]), w, p + "global_att.5.")
    gate = torch.sigmoid(l + g)
    return 2 * x * gate + 2 * residual * (1 - gate)


def patch_embed(w: Weights, img: torch.Tensor, is_longer: torch.Tensor, cfg: ClapCfg) -> torch.Tensor:
    """ClapAudioPatchEmbed.forward (modeling_clap.py:296-344): 4x4/4 conv on the global channel; for the samples flagged
    `is_longer`, the three local crops go through a 4x12/(4,12) conv, are laid side by side along time, zero-padded to the
    global width and fused by the AFF block. Then flatten to tokens + LayerNorm."""
    p = ENC + "patch_embed."
    if cfg.enable_fusion:
        glob = F.conv2d(img[:, 0:1], w[p + "proj.weight"], w[p + "proj.bias"], stride=cfg.patch)
        idx = torch.where(is_longer.reshape(-1) == 1)[0]
        if len(idx) > 0:
            local = img[idx, 1:].contiguous()
            b, c, hh, ww = local.shape
            local = F.conv2d(local.view(b * c, 1, hh, ww), w[p + "mel_conv2d.weight"], w[p + "mel_conv2d.bias"],
                             stride=(cfg.patch, cfg.patch * 3))
            _, feat, hh, ww = local.shape
            local = local.view(b, c, feat, hh, ww).permute(0, 2, 3, 1, 4).contiguous().flatten(3)
            local = F.pad(local, (0, glob.shape[-1] - local.shape[-1]))
            glob = glob.clone()
            glob[idx] = _aff(w, p + "fusion_model.", glob[idx], local)
        x = glob
    else:
        x = F.conv2d(img, w[p + "proj.weight"], w[p + "proj.bias"], stride=cfg.patch)
    x = x.flatten(2).transpose(1, 2)
    return _ln(x, w, p + "norm.", cfg.eps)


def _rel_index(ws: int) -> torch.Tensor:
    """create_relative_position_index (modeling_clap.py:427-438)."""
    coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing="ij")).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def _shift_mask(h: int, wd: int, ws: int, shift: int) -> torch.Tensor:
    """get_attn_mask (modeling_clap.py:525-550): -100 between tokens that come from different image regions after the roll."""
    img = torch.zeros(1, h, wd, 1)
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, hs, wsl, :] = cnt
            cnt += 1
    win = img.view(1, h // ws, ws, wd // ws, ws, 1).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws)
    m = win.unsqueeze(1) - win.unsqueeze(2)
    return m.masked_fill(m != 0, -100.0).masked_fill(m == 0, 0.0)


def swin_block(w: Weights, p: str, x: torch.Tensor, res: Tuple[int, int], heads: int, shift: int, cfg: ClapCfg):
    """ClapAudioLayer.forward (modeling_clap.py:560-623) with ClapAudioSelfAttention.forward (:374-425)."""
    h, wd = res
    b, _, c = x.shape
    ws = cfg.window
    if min(res) <= ws:  # set_shift_and_window_size (:517-523)
        shift, ws = 0, min(res)
    y = _ln(x, w, p + "layernorm_before.", cfg.eps).view(b, h, wd, c)
    if shift > 0:
        y = torch.roll(y, shifts=(-shift, -shift), dims=(1, 2))
    win = y.view(b, h // ws, ws, wd // ws, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, c)  # window_partition
    a = p + "attention.self."
    hd = c // heads
    q = (win @ w[a + "query.weight"].t() + w[a + "query.bias"]).view(-1, ws * ws, heads, hd).transpose(1, 2)
    k = (win @ w[a + "key.weight"].t() + w[a + "key.bias"]).view(-1, ws * ws, heads, hd).transpose(1, 2)
    v = (win @ w[a + "value.weight"].t() + w[a + "value.bias"]).view(-1, ws * ws, heads, hd).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    bias = w[a + "relative_position_bias_table"][_rel_index(cfg.window).view(-1)]
    s = s + bias.view(ws * ws, ws * ws, -1).permute(2, 0, 1).unsqueeze(0)
    if shift > 0:
        mask = _shift_mask(h, wd, ws, shift)
        nw = mask.shape[0]
        s = (s.view(b, nw, heads, ws * ws, ws * ws) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, ws * ws, ws * ws)
    o = (s.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(-1, ws * ws, c)
    o = o @ w[p + "attention.output.dense.weight"].t() + w[p + "attention.output.dense.bias"]
    o = o.view(b, h // ws, wd // ws, ws, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(b, h, wd, c)            # window_reverse
    if shift > 0:
        o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))
    x = x + o.reshape(b, h * wd, c)
    y = _ln(x, w, p + "layernorm_after.", cfg.eps)
    y = F.gelu(y @ w[p + "intermediate.dense.weight"].t() + w[p + "intermediate.dense.bias"])             # exact (erf) GELU
    return x + y @ w[p + "output.dense.weight"].t() + w[p + "output.dense.bias"]


def patch_merge(w: Weights, p: str, x: torch.Tensor, res: Tuple[int, int], cfg: ClapCfg) -> torch.Tensor:
    """ClapAudioPatchMerging.forward (modeling_clap.py:710-733): 2x2 neighbourhoods -> LayerNorm(4C) -> Linear(4C, 2C)."""
    h, wd = res
    b, _, c = x.shape
    x = x.view(b, h, wd, c)
    x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1).view(b, -1, 4 * c)
    return _ln(x, w, p + "norm.", cfg.eps) @ w[p + "reduction.weight"].t()


def clap_audio_embed(w: Weights, mel: torch.Tensor, is_longer: torch.Tensor, cfg: ClapCfg) -> torch.Tensor:
    """ClapAudioModelWithProjection.forward(...).audio_embeds (modeling_clap.py:1725-1775 -> ClapAudioEncoder.forward
    :814-918 -> ClapProjectionLayer :932-937). mel: [B, 4 (fusion) | 1, T <= 1024, 64]; is_longer: [B, 1] bool."""
    x = _bn_eval(mel.float().transpose(1, 3), w, ENC + "batch_norm.").transpose(1, 3)      # :830-832, over the mel bins
    img = reshape_mel2img(x, cfg)
    frames = img.shape[2]
    x = patch_embed(w, img, is_longer, cfg)
    grid = cfg.spec_size // cfg.patch
    for i, (depth, heads) in enumerate(zip(cfg.depths, cfg.heads)):
        res = (grid // 2 ** i, grid // 2 ** i)
        for j in range(depth):
            x = swin_block(w, f"{ENC}layers.{i}.blocks.{j}.", x, res, heads, 0 if j % 2 == 0 else cfg.window // 2, cfg)
        if i < len(cfg.depths) - 1:
            x = patch_merge(w, f"{ENC}layers.{i}.downsample.", x, res, cfg)
    x = _ln(x, w, ENC + "norm.", cfg.eps)                                                   # :886
    b, _, c = x.shape
    ratio = cfg.spec_size // cfg.num_mel_bins
    fs = frames // 2 ** (len(cfg.depths) - 1) // cfg.patch
    x = x.permute(0, 2, 1).contiguous().reshape(b, c, fs, fs)                               # :890-896
    cf = fs // ratio
    x = x.reshape(b, c, fs // cf, cf, fs).permute(0, 1, 3, 2, 4).contiguous().reshape(b, c, cf, -1)
    latent = x.flatten(2).mean(-1)                                                          # AdaptiveAvgPool1d(1) :907
    hmid = torch.relu(latent @ w["audio_projection.linear1.weight"].t() + w["audio_projection.linear1.bias"])
    return hmid @ w["audio_projection.linear2.weight"].t() + w["audio_projection.linear2.bias"]


def hf_clap(cfg: ClapCfg, seed: int = 0):
    """The stand-in module itself with seeded random weights (and non-trivial BatchNorm statistics), in eval mode."""
    from transformers import ClapAudioConfig, ClapAudioModelWithProjection
    torch.manual_seed(seed)
    c = ClapAudioConfig(enable_fusion=cfg.enable_fusion, depths=list(cfg.depths), num_attention_heads=list(cfg.heads),
                        patch_embeds_hidden_size=cfg.embed, hidden_size=cfg.hidden, projection_dim=cfg.projection_dim,
                        window_size=cfg.window, num_mel_bins=cfg.num_mel_bins, spec_size=cfg.spec_size,
                        aff_block_r=cfg.aff_r)
    m = ClapAudioModelWithProjection(c).eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, buf in m.named_buffers():
            if name.endswith("running_mean"):
                buf.copy_(torch.randn(buf.shape, generator=g) * 0.3)
            elif name.endswith("running_var"):
                buf.copy_(torch.rand(buf.shape, generator=g) + 0.5)
        for name, prm in m.named_parameters():
            if "relative_position_bias_table" in name:
                prm.copy_(torch.randn(prm.shape, generator=g) * 0.2)  # HF initialises the table to zeros
    return m
