"""Parity of the three stage engines (through the C ABI) against the CPU oracle on identical seeded inputs/weights.
Tolerance: 1e-3 relative (max|a-b|/max|b|) for stage outputs — BASELINE.json north_star; fp16 operands, fp32 accumulate."""
import pytest
import torch

from conftest import rel_err
from oracle import restate as R
from oracle import synth

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.mark.parametrize("cfg,B", [
    (R.MapperCfg(E=64, d=128, P=3, K=5, H=2, L=2), 4),       # hd 64
    (R.MapperCfg(E=72, d=96, P=2, K=4, H=2, L=1), 3),        # hd 48
    (R.MapperCfg(E=64, d=192, P=4, K=3, H=2, L=2), 2),       # hd 96
    (R.MapperCfg(E=512, d=256, P=1, K=7, H=2, L=3), 5),      # hd 128
    (R.MapperCfg(E=768, d=768, P=10, K=10, H=8, L=8), 2),    # config #1-like (GPT-2-small width)
    (R.MapperCfg(kind="windowed", E=64, d=128, P=2, K=5, H=2, L=2, W=3, use_pos=True), 3),
    (R.MapperCfg(kind="windowed", E=64, d=128, P=2, K=5, H=2, L=2, W=2, use_pos=False), 2),
    (R.MapperCfg(kind="mlp", E=64, d=128, K=4), 3),
])
def test_mapper_matches_oracle(cuda_device, cfg, B):
    from clipcap_b200.engine import MapperEngine
    w = synth.mapper_weights(cfg)
    emb = synth.embeddings(B * (cfg.W if cfg.kind == "windowed" else 1), cfg.E)
    if cfg.kind == "windowed":
        emb = emb.view(B, cfg.W, cfg.E)
    ref = R.mapper_forward(w, emb, cfg)
    eng = MapperEngine(w, kind=cfg.kind, E=cfg.E, d=cfg.d, P=cfg.P, K=cfg.K, H=cfg.H, L=cfg.L, W=cfg.W,
                       use_pos=cfg.use_pos, max_batch=B + 1, device=cuda_device)
    for dt in (torch.float32, torch.float16):
        out = eng.forward(emb.to(cuda_device, dt))
        assert out.dtype == dt and tuple(out.shape) == (B, cfg.K, cfg.d)
        assert rel_err(out, ref) < (TOL if dt == torch.float32 else 2e-3)
    assert eng.last_launches > 0


@pytest.mark.parametrize("cfg,B", [
    (R.VitCfg(image_size=28, patch=14, width=128, layers=2, heads=2, mlp_dim=256, out_dim=64), 3),
    (R.VitCfg(image_size=56, patch=14, width=256, layers=3, heads=4, mlp_dim=512, out_dim=96), 2),
])
@pytest.mark.parametrize("normalize", [False, True])
def test_vit_matches_oracle(cuda_device, cfg, B, normalize):
    from clipcap_b200.engine import VitEngine
    w = synth.vit_weights(cfg)
    px = synth.pixels(B, cfg.image_size)
    ref = R.vit_encode(w, px, cfg, normalize)
    eng = VitEngine(w, cfg.image_size, cfg.patch, cfg.width, cfg.layers, cfg.heads, cfg.mlp_dim, cfg.out_dim,
                    max_batch=B, device=cuda_device)
    out = eng.forward(px.to(cuda_device), normalize=normalize)
    assert rel_err(out, ref) < TOL
    out16 = eng.forward(px.to(cuda_device).half(), normalize=normalize)
    assert out16.dtype == torch.float16 and rel_err(out16, ref) < 3e-3


def test_vit_l14_one_image(cuda_device):
    """Full-size ViT-L/14 on one image (the oracle needs ~1 s of CPU)."""
    from clipcap_b200.engine import VitEngine
    cfg = R.VitCfg()
    w = synth.vit_weights(cfg)
    px = synth.pixels(2, 224)
    ref = R.vit_encode(w, px, cfg)
    eng = VitEngine(w, max_batch=2, device=cuda_device)
    out = eng.forward(px.to(cuda_device))
    assert rel_err(out, ref) < TOL


@pytest.mark.parametrize("cfg,B,T", [
    (R.Gpt2Cfg(d=128, L=2, H=2, V=1003, n_pos=64), 3, 7),
    (R.Gpt2Cfg(d=192, L=3, H=3, V=517, n_pos=32), 2, 1),
    (R.Gpt2Cfg(d=768, L=12, H=12, V=50257, n_pos=1024), 2, 12),  # GPT-2-small
])
def test_gpt2_logits_match_oracle(cuda_device, cfg, B, T):
    from clipcap_b200.engine import Gpt2Engine
    w = synth.gpt2_weights(cfg)
    x = torch.randn(B, T, cfg.d, generator=torch.Generator().manual_seed(7))
    ref = R.gpt2_logits(w, x, cfg)
    eng = Gpt2Engine(w, cfg.d, cfg.L, cfg.H, cfg.V, cfg.n_pos, max_seqs=B, max_len=T + 2, device=cuda_device)
    last = eng.logits(x.to(cuda_device))
    assert rel_err(last, ref[:, -1]) < TOL
    full = eng.logits(x.to(cuda_device), all_positions=True)
    assert rel_err(full, ref) < TOL
    ids = torch.tensor([[0, 5, cfg.V - 1], [1, 2, 3]], device=cuda_device)
    assert torch.equal(eng.embed(ids).cpu(), w["transformer.wte.weight"][ids.cpu()])


# ------------------------------------------------------------------------------------------------ decode loops
from helpers import check_tokens_against_oracle  # noqa: E402


def _lm_setup(cfg, B, Tp, seed=11, wte_std=0.1):
    w = synth.gpt2_weights(cfg, wte_std=wte_std)
    prefix = torch.randn(B, Tp, cfg.d, generator=torch.Generator().manual_seed(seed)) * 0.5
    return w, prefix


@pytest.mark.parametrize("cfg,B,Tp,EL", [
    (R.Gpt2Cfg(d=128, L=2, H=2, V=1003, n_pos=64), 5, 5, 10),
    (R.Gpt2Cfg(d=192, L=3, H=3, V=517, n_pos=48), 3, 1, 12),
])
def test_greedy_matches_oracle(cuda_device, cfg, B, Tp, EL):
    from clipcap_b200.engine import Gpt2Engine
    w, prefix = _lm_setup(cfg, B, Tp)
    stop = cfg.V - 1
    oracle = R.generate_greedy_batch(w, cfg, prefix, EL, stop)
    eng = Gpt2Engine(w, cfg.d, cfg.L, cfg.H, cfg.V, cfg.n_pos, max_seqs=B, max_len=Tp + EL, device=cuda_device)
    for _ in range(2):  # second call replays the captured CUDA graph
        toks, lens, _ = eng.generate(prefix.to(cuda_device), "greedy", 1, EL, 1.0, stop)
        exact, flips = check_tokens_against_oracle(toks, lens, oracle, margin_tol=5e-3)
        assert exact >= B - 1
    assert eng.last_launches > 0


def test_greedy_stop_token(cuda_device):
    """Stop-token bookkeeping of generate_beam(beam_size=1): tokens end with the stop token, length counts it."""
    from clipcap_b200.engine import Gpt2Engine
    cfg, B, Tp, EL = R.Gpt2Cfg(d=128, L=2, H=2, V=1003, n_pos=64), 4, 5, 10
    w, prefix = _lm_setup(cfg, B, Tp)
    free = R.generate_greedy_batch(w, cfg, prefix, EL, stop_token=cfg.V + 7)
    stop = free[0][0][3]  # sample 0 will now stop at (or before) step 3
    oracle = R.generate_greedy_batch(w, cfg, prefix, EL, stop)
    assert len(oracle[0][0]) <= 4
    eng = Gpt2Engine(w, cfg.d, cfg.L, cfg.H, cfg.V, cfg.n_pos, max_seqs=B, max_len=Tp + EL, device=cuda_device)
    toks, lens, _ = eng.generate(prefix.to(cuda_device), "greedy", 1, EL, 1.0, stop)
    exact, _ = check_tokens_against_oracle(toks, lens, oracle, margin_tol=5e-3)
    assert exact >= B - 1
    assert lens[0].item() == len(oracle[0][0])


@pytest.mark.parametrize("beam", [1, 3, 5])
@pytest.mark.parametrize("temperature", [1.0, 0.7])
def test_beam_matches_oracle(cuda_device, beam, temperature):
    from clipcap_b200.engine import Gpt2Engine
    cfg, B, Tp, EL = R.Gpt2Cfg(d=128, L=2, H=2, V=1003, n_pos=64), 4, 5, 9
    w, prefix = _lm_setup(cfg, B, Tp)
    free = R.generate_beam(w, cfg, prefix[:1], beam, EL, temperature, cfg.V + 7)[0]
    stop = free[4]  # make at least one beam of sample 0 terminate early
    eng = Gpt2Engine(w, cfg.d, cfg.L, cfg.H, cfg.V, cfg.n_pos, max_seqs=B * beam, max_len=Tp + EL, device=cuda_device)
    toks, lens, scores = eng.generate(prefix.to(cuda_device), "beam", beam, EL, temperature, stop)
    toks, lens, scores = toks.cpu().tolist(), lens.cpu().tolist(), scores.cpu().tolist()
    n_exact = 0
    for i in range(B):
        otoks, oscore, _ = R.generate_beam(w, cfg, prefix[i:i + 1], beam, EL, temperature, stop)
        got = toks[i][:lens[i]]
        if got == otoks:
            n_exact += 1
            assert abs(scores[i] - oscore) < 2e-3 * max(1.0, abs(oscore))
        else:  # only a near-tie between beams may differ: the returned score must then match the oracle's best
            assert abs(scores[i] - oscore) < 2e-3 * max(1.0, abs(oscore)), (i, got, otoks, scores[i], oscore)
    assert n_exact >= B - 1


@pytest.mark.parametrize("beam", [1, 5])
def test_beam_with_text_prefix_matches_oracle(cuda_device, beam):
    """generate_beam(text_prefix_tokens=...) — the VQA-style prompt of base.py:75-77: the text-prefix embeddings follow the
    mapper prefix, then beam search as usual. Through the drop-in function, against the oracle's generate_beam."""
    from clipcap_b200.encoders.config import EncoderConfig
    from clipcap_b200.inference.base import generate_beam, generate_beam_tokens
    from clipcap_b200.model import ClipCapModelPrefixOnly, Config
    from oracle import ref_runner as RR
    gcfg = R.Gpt2Cfg(d=128, L=2, H=2, V=1003, n_pos=64)
    mcfg = R.MapperCfg(E=64, d=128, P=3, K=5, H=2, L=2)
    map_w, lm_w = synth.mapper_weights(mcfg), synth.gpt2_weights(gcfg, wte_std=0.1)
    cfg = Config(language_model="tiny:128:2:2:1003:64", prefix_length=5, projection_length=3, transformer_layers=2,
                 transformer_attention_heads=2, encoder_config=EncoderConfig(encoder_embedding_size=64))
    model = ClipCapModelPrefixOnly(cfg)
    sd = {f"transformer_mapper.{k}": v for k, v in map_w.items()}
    sd.update({f"language_model.{k}": v for k, v in lm_w.items()})
    model.load_state_dict(sd, strict=True)
    model = model.eval().to(cuda_device)
    emb = synth.embeddings(3, 64, seed=31)
    tp = torch.tensor([[7, 900, 42, 7]])
    prefix_ref = R.mapper_forward(map_w, emb, mcfg)
    stop = 1002
    want = [R.generate_beam(lm_w, gcfg, prefix_ref[i:i + 1], beam, 8, 1.0, stop, text_prefix_tokens=tp) for i in range(3)]
    prefix = model.transformer_mapper(emb.to(cuda_device))
    toks, lens, scores = generate_beam_tokens(model, prefix, tp.to(cuda_device), beam, 8, 1.0, stop)
    ok = 0
    for i in range(3):
        got = toks[i, :int(lens[i])].tolist()
        if got == want[i][0]:
            ok += 1
            if beam > 1:
                assert abs(float(scores[i]) - want[i][1]) < 2e-3 * abs(want[i][1])
        else:  # only a near-tie in the oracle's own scores may flip a beam
            assert min(min(t["margin"]) for t in want[i][2]) < 5e-3, (i, got, want[i][0])
    assert ok >= 2
    text = generate_beam(model, RR.FakeTokenizer(stop), prefix[:1], text_prefix_tokens=tp.to(cuda_device), beam_size=beam,
                         entry_length=8)
    assert text[0] == " ".join(str(t) for t in toks[0, :int(lens[0])].tolist())


def test_generate_wrapper(cuda_device):
    """clipcap.inference.generate (generate.py:8-44): BOS (+ text prefix) embeddings after the mapper prefix, then
    generate_no_beam with the same text_prefix_tokens. With top_k=1 the draw is deterministic: compare with the oracle's
    restatement of that exact call sequence."""
    from clipcap_b200.encoders.config import EncoderConfig
    from clipcap_b200.inference.generate import generate
    from clipcap_b200.model import ClipCapModelPrefixOnly, Config
    gcfg = R.Gpt2Cfg(d=128, L=2, H=2, V=1003, n_pos=128)  # generate() always decodes up to 67 tokens
    mcfg = R.MapperCfg(E=64, d=128, P=3, K=5, H=2, L=2)
    map_w, lm_w = synth.mapper_weights(mcfg), synth.gpt2_weights(gcfg, wte_std=0.1)
    cfg = Config(language_model="tiny:128:2:2:1003:128", prefix_length=5, projection_length=3, transformer_layers=2,
                 transformer_attention_heads=2, encoder_config=EncoderConfig(encoder_embedding_size=64))
    model = ClipCapModelPrefixOnly(cfg)
    sd = {f"transformer_mapper.{k}": v for k, v in map_w.items()}
    sd.update({f"language_model.{k}": v for k, v in lm_w.items()})
    model.load_state_dict(sd, strict=True)
    model = model.eval().to(cuda_device)

    class Tok:  # the slice of the HF tokenizer interface generate() uses
        bos_token = "1001"

        def encode(self, text, return_tensors=None):
            if text == ".":
                return [13]
            ids = [int(t) for t in text.replace("1001", "1001 ").split()]
            return torch.tensor([ids]) if return_tensors == "pt" else ids

        def decode(self, ids):
            return " ".join(str(int(i)) for i in ids)

    emb = synth.embeddings(1, 64, seed=5)
    out = generate(model, Tok(), emb.to(cuda_device), top_p=0.9, top_k=1, temperature=1.0, number_to_generate=1,
                   text_prefix="17 5", seed=1)
    tpt = torch.tensor([[1001, 17, 5]])
    # oracle: prompt = [mapper prefix, wte(BOS + text prefix)], then generate_no_beam appends wte(text prefix tokens) AGAIN
    # (generate.py:30-41 feeding no_beam.py:27-29) and starts its history with them
    prompt = torch.cat((R.mapper_forward(map_w, emb, mcfg), lm_w["transformer.wte.weight"][tpt]), dim=1)
    want = R.generate_sampling(lm_w, gcfg, prompt, "sample", lambda p: int(p.argmax()), entry_length=67,
                               text_prefix_tokens=tpt, stop_token=13, top_p=0.9, top_k=1, temperature=1.0,
                               repetition_penalty=1.2)
    got = [int(t) for t in out[0].split()]
    assert got[:3] == [1001, 17, 5]
    n = min(len(want), len(got) - 3, 20)
    assert n >= 1 and got[3:3 + n] == want[:n], (got, want)


@pytest.mark.parametrize("B,Tp,EL", [
    (37, 5, 6),    # ragged batch: not a multiple of the 32-row epilogue groups / 128-row tiles (tcgen05 decode path)
    (17, 3, 4),    # one row more than the skinny path takes
    (16, 3, 4),    # the most the skinny path takes
    (2, 7, 1),     # entry_length 1: prefill + head only, no decode step
    (3, 1, 63),    # single prefix token, decode up to the last position the model has (1 + 63 = n_positions)
])
def test_greedy_edge_shapes(cuda_device, B, Tp, EL):
    """Ragged and boundary shapes of the decode loop against the oracle (a few rows of it when the batch is large), plus
    batch invariance of every row against the same rows decoded in a smaller call."""
    from clipcap_b200.engine import Gpt2Engine
    cfg = R.Gpt2Cfg(d=128, L=2, H=2, V=1003, n_pos=64)
    w, prefix = _lm_setup(cfg, B, Tp)
    stop = cfg.V - 1
    rows = sorted(set([0, B // 2, B - 1]))
    oracle = R.generate_greedy_batch(w, cfg, prefix[rows], EL, stop)
    eng = Gpt2Engine(w, cfg.d, cfg.L, cfg.H, cfg.V, cfg.n_pos, max_seqs=B, max_len=Tp + EL, device=cuda_device)
    toks, lens, _ = eng.generate(prefix.to(cuda_device), "greedy", 1, EL, 1.0, stop)
    assert tuple(toks.shape) == (B, EL) and int(lens.min()) >= 1 and int(lens.max()) <= EL
    exact, _ = check_tokens_against_oracle(toks[rows], lens[rows], oracle, margin_tol=5e-3)
    assert exact >= len(rows) - 1
    if B > 4:  # the last three rows alone: another kernel path (skinny) or other tiles, same tokens up to near-ties
        t3, l3, _ = eng.generate(prefix[B - 3:].contiguous().to(cuda_device), "greedy", 1, EL, 1.0, stop)
        same = sum(int(torch.equal(t3[i, :l3[i]], toks[B - 3 + i, :lens[B - 3 + i]])) for i in range(3))
        assert same >= 2


def test_entry_length_limits(cuda_device):
    """entry_length up to 128 per call (the handle's state buffers); beyond that, and beyond max_len, a CC_ESHAPE error."""
    from clipcap_b200._ffi import CCError
    from clipcap_b200.engine import Gpt2Engine
    cfg = R.Gpt2Cfg(d=128, L=2, H=2, V=1003, n_pos=256)
    w, prefix = _lm_setup(cfg, 2, 4)
    eng = Gpt2Engine(w, cfg.d, cfg.L, cfg.H, cfg.V, cfg.n_pos, max_seqs=2, max_len=4 + 200, device=cuda_device)
    toks, lens, _ = eng.generate(prefix.to(cuda_device), "greedy", 1, 128, 1.0, cfg.V + 5)  # stop token never produced
    assert tuple(toks.shape) == (2, 128) and lens.tolist() == [128, 128]
    oracle = R.generate_greedy_batch(w, cfg, prefix[:1], 24, cfg.V + 5)
    exact, _ = check_tokens_against_oracle(toks[:1, :24], torch.tensor([24]), oracle, margin_tol=5e-3)
    with pytest.raises(CCError, match="CC_ESHAPE"):
        eng.generate(prefix.to(cuda_device), "greedy", 1, 129, 1.0, 0)


def test_vit_l14_ragged_batch(cuda_device):
    """Full-size ViT-L/14 on 9 images inside a handle created for 16 (a batch that fills neither a 256-row CTA-pair tile nor a
    persistent wave evenly), fp16 pixels, with normalisation; one more image in a second call re-uses the handle."""
    from clipcap_b200.engine import VitEngine
    cfg = R.VitCfg()
    w = synth.vit_weights(cfg)
    px = synth.pixels(9, 224, seed=77)
    ref = R.vit_encode(w, px.half().float(), cfg, normalize=True)
    eng = VitEngine(w, max_batch=16, device=cuda_device)
    out = eng.forward(px.half().to(cuda_device), normalize=True)
    assert out.dtype == torch.float16 and rel_err(out, ref) < 3e-3
    one = eng.forward(px[4:5].to(cuda_device), normalize=True)
    assert rel_err(one, ref[4:5]) < TOL


def test_mapper_ragged_batch(cuda_device):
    """257 samples through a handle created for 300: the last 128-row tile of every GEMM is a single row."""
    from clipcap_b200.engine import MapperEngine
    cfg = R.MapperCfg(E=64, d=128, P=3, K=5, H=2, L=2)
    w = synth.mapper_weights(cfg)
    emb = synth.embeddings(257, cfg.E, seed=9)
    ref = R.mapper_forward(w, emb, cfg)
    eng = MapperEngine(w, E=cfg.E, d=cfg.d, P=cfg.P, K=cfg.K, H=cfg.H, L=cfg.L, max_batch=300, device=cuda_device)
    assert rel_err(eng.forward(emb.to(cuda_device)), ref) < TOL
