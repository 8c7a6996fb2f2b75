"""Sampling decode modes (generate_nucleus_sampling / generate_no_beam) on the GPU.

The random stream cannot match torch.multinomial's, so parity is checked in two ways: (1) the selection kernel's draws
over many Philox streams reproduce the oracle's per-step distribution (same kept set, total-variation distance at the
sampling-noise level); (2) with top_k == 1 the loops are deterministic and the token ids equal the reference's (golden
fixtures made from the reference itself)."""
import ctypes as C

import numpy as np
import pytest
import torch

from golden_util import ENTRY, SAMPLING_CASES, load_lm_case
from oracle import restate as R

pytestmark = pytest.mark.gpu


def _cfg(mode, entry_length, stop_token=13, seed=0, history_dev=None, n_history=0, **kw):
    from clipcap_b200 import _ffi
    return _ffi.cc_gen_cfg(_ffi.CC_GEN_NUCLEUS if mode == "nucleus" else _ffi.CC_GEN_SAMPLE, 1, entry_length,
                           float(kw.get("temperature", 1.0)), stop_token, float(kw.get("top_p", 1.0)),
                           int(kw.get("top_k", 0)), float(kw.get("repetition_penalty", 1.0)),
                           int(kw.get("desired_sentence_length", 50)), float(kw.get("sentence_length_factor", 1.0)),
                           n_history, history_dev, seed)


@pytest.mark.parametrize("ci", range(len(SAMPLING_CASES)))
@pytest.mark.parametrize("V", [1003, 50257])
def test_sample_kernel_distribution(cuda_device, ci, V):
    from clipcap_b200 import _ffi
    lib = _ffi.lib()
    mode, kw = SAMPLING_CASES[ci]
    g = torch.Generator().manual_seed(100 + ci)
    logits1 = torch.randn(V, generator=g) * (3.0 if V < 2000 else 4.0)
    rows, EL, step = 32768, 6, 2
    prefix_hist = torch.tensor([7, 3, 7], dtype=torch.int32)
    gen_so_far = torch.tensor([int(logits1.argmax()), 11], dtype=torch.int32)  # includes the most likely token
    history = torch.cat((prefix_hist, gen_so_far)).long() if mode == "sample" else None
    if mode == "nucleus":
        want = R.nucleus_distribution(logits1[None], kw["top_p"], kw["top_k"], kw["temperature"])[0]
    else:
        want = R.no_beam_distribution(logits1, history, kw["top_p"], kw["top_k"], kw["temperature"],
                                      kw["repetition_penalty"], 13)
        want = want.clone()

    logits = logits1[None].repeat(rows, 1).contiguous().to(cuda_device)
    tokens = torch.zeros(rows, EL, dtype=torch.int32, device=cuda_device)
    tokens[:, :2] = gen_so_far.to(cuda_device)
    stopped = torch.zeros(rows, dtype=torch.int32, device=cuda_device)
    lengths = torch.full((rows,), step, dtype=torch.int32, device=cuda_device)
    hist_dev = prefix_hist.to(cuda_device)
    cfg = _cfg(mode, EL, seed=1234 + ci, history_dev=hist_dev.data_ptr(), n_history=3, **kw)
    _ffi.check(lib.cc_op_sample(logits.data_ptr(), rows, V, C.byref(cfg), step, tokens.data_ptr(), stopped.data_ptr(),
                                lengths.data_ptr(), _ffi.current_stream_ptr()))
    torch.cuda.synchronize()
    drawn = tokens[:, step].cpu().long()
    stop_rows = stopped.cpu().bool()
    if mode == "sample":  # a drawn stop token is not stored: recover it from the flag
        drawn = torch.where(stop_rows, torch.full_like(drawn, 13), drawn)
    freq = torch.bincount(drawn, minlength=V).double() / rows
    kept = want > 0
    assert freq[~kept].sum() == 0, "a token outside the reference's kept set was drawn"
    tv = 0.5 * (freq - want.double()).abs().sum().item()
    k = int(kept.sum())
    noise = 0.5 * np.sqrt(2 * k / (np.pi * rows))  # expected TV of an exact sampler with `rows` draws over k outcomes
    assert tv < 3 * noise + 0.005, (tv, noise, k)
    if k > 1:
        assert (freq > 0).sum() > 1  # really sampling, not arg-max


@pytest.mark.parametrize("name", ["tiny_a", "tiny_b"])
@pytest.mark.parametrize("ci", [2, 5])  # the top_k == 1 cases: deterministic
@pytest.mark.parametrize("ti", [0, 1])
def test_generate_sampling_topk1_matches_reference(cuda_device, name, ci, ti):
    from clipcap_b200.engine import Gpt2Engine
    spec, gcfg, mcfg, map_w, lm_w, g = load_lm_case(name)
    mode, kw = SAMPLING_CASES[ci]
    tp = [5, 17, 5] if ti else []
    want = g[f"samp{ci}_tp{ti}_tokens"].tolist()[len(tp):]
    prefix = torch.from_numpy(g["prefix"])[:1]
    if tp:  # the loops append the text-prefix embeddings to the prompt (nucleus_sampling.py:27-29 / no_beam.py:27-29)
        prefix = torch.cat((prefix, lm_w["transformer.wte.weight"][torch.tensor(tp)][None]), dim=1)
    eng = Gpt2Engine(lm_w, gcfg.d, gcfg.L, gcfg.H, gcfg.V, gcfg.n_pos, max_seqs=8, max_len=prefix.shape[1] + ENTRY,
                     device=cuda_device)
    batch = prefix.repeat(3, 1, 1).to(cuda_device)  # rows are independent: every row must give the same caption
    toks, lens, _ = eng.generate(batch, "nucleus" if mode == "nucleus" else "sample", 1, ENTRY,
                                 kw.get("temperature", 1.0), 13, top_p=kw["top_p"], top_k=int(kw["top_k"]),
                                 repetition_penalty=kw.get("repetition_penalty", 1.0), history=tp, seed=7)
    toks, lens = toks.cpu().tolist(), lens.cpu().tolist()
    for i in range(3):
        assert toks[i][:lens[i]] == want, (i, toks[i], lens[i], want)


def test_sampling_api_and_seed(cuda_device):
    """generate_nucleus_sampling / generate_no_beam / generate with the reference signatures; a seed makes the draw
    reproducible, different seeds explore different captions."""
    from clipcap_b200.inference.no_beam import generate_no_beam
    from clipcap_b200.inference.nucleus_sampling import generate_nucleus_sampling
    from clipcap_b200.model import ClipCapModelPrefixOnly, Config
    from clipcap_b200.encoders.config import EncoderConfig
    from oracle import ref_runner as RR
    spec, gcfg, mcfg, map_w, lm_w, g = load_lm_case("tiny_a")
    cfg = Config(language_model=spec, prefix_length=mcfg.K, projection_length=mcfg.P, transformer_layers=mcfg.L,
                 transformer_attention_heads=mcfg.H, encoder_config=EncoderConfig(encoder_embedding_size=mcfg.E))
    model = ClipCapModelPrefixOnly(cfg)
    sd = {f"transformer_mapper.{k}": v for k, v in map_w.items()}
    sd.update({f"language_model.{k}": v for k, v in lm_w.items()})
    model.load_state_dict(sd, strict=True)
    model = model.eval().to(cuda_device)
    tok = RR.FakeTokenizer()
    prefix = model.transformer_mapper(torch.from_numpy(g["emb"])[:1].to(cuda_device))
    a = generate_nucleus_sampling(model, tok, prefix, number_to_generate=2, entry_length=8, top_p=0.95, temperature=1.5, seed=5)
    b = generate_nucleus_sampling(model, tok, prefix, number_to_generate=2, entry_length=8, top_p=0.95, temperature=1.5, seed=5)
    c = generate_nucleus_sampling(model, tok, prefix, number_to_generate=2, entry_length=8, top_p=0.95, temperature=1.5, seed=99)
    assert a == b and len(a) == 2 and a != c
    tp = torch.tensor([[5, 17, 5]], device=cuda_device)
    d = generate_no_beam(model, tok, prefix, number_to_generate=1, text_prefix_tokens=tp, entry_length=8, seed=3)
    assert d[0].split()[:3] == ["5", "17", "5"]  # the text prefix leads the returned tokens (no_beam.py:34)
    assert "13" not in d[0].split()[3:]          # the stop token is never returned by generate_no_beam
