"""The serving loop (clipcap_b200.pipeline.CaptionPipeline) and the preprocess EncoderMapper on the GPU: pipelined results
equal the direct stage calls; ragged last batch; pinned-memory requirement."""
import numpy as np
import pytest
import torch

from golden_util import load_lm_case
from oracle import restate as R
from oracle import synth

pytestmark = pytest.mark.gpu


def _build(cuda_device):
    from clipcap_b200.encoders.clip import CLIPModel, ViTImageTower
    from clipcap_b200.encoders.config import EncoderConfig
    from clipcap_b200.model import ClipCapModelPrefixOnly, Config
    spec, gcfg, mcfg, map_w, lm_w, g = load_lm_case("tiny_a")
    vcfg = R.VitCfg(image_size=28, patch=14, width=128, layers=2, heads=2, mlp_dim=512, out_dim=mcfg.E)
    tower = ViTImageTower(vcfg.image_size, vcfg.patch, vcfg.width, vcfg.layers, vcfg.heads, vcfg.out_dim, vcfg.mlp_dim)
    tower.load_state_dict(synth.vit_weights(vcfg), strict=True)
    encode_fn = CLIPModel(tower).eval().to(cuda_device)
    cfg = Config(language_model=spec, prefix_length=mcfg.K, projection_length=mcfg.P, transformer_layers=mcfg.L,
                 transformer_attention_heads=mcfg.H, encoder_config=EncoderConfig(encoder_embedding_size=mcfg.E))
    model = ClipCapModelPrefixOnly(cfg)
    sd = {f"transformer_mapper.{k}": v for k, v in map_w.items()}
    sd.update({f"language_model.{k}": v for k, v in lm_w.items()})
    model.load_state_dict(sd, strict=True)
    return encode_fn, model.eval().to(cuda_device), vcfg, int(g["stop_token"])


def test_pipeline_equals_direct_calls(cuda_device):
    from clipcap_b200.inference.base import generate_greedy_tokens
    from clipcap_b200.pipeline import CaptionPipeline
    encode_fn, model, vcfg, stop = _build(cuda_device)
    B, EL = 4, 7
    batches = [synth.pixels(n, vcfg.image_size, seed=50 + i).pin_memory() for i, n in enumerate((4, 4, 4, 3, 4))]
    pipe = CaptionPipeline(encode_fn, model, B, vcfg.image_size, EL, stop, cuda_device)
    got = [(t.clone(), l.clone()) for t, l in pipe.run(batches)]
    assert len(got) == len(batches)
    for px, (toks, lens) in zip(batches, got):
        prefix = model.transformer_mapper(encode_fn(px.to(cuda_device)))
        want_t, want_l, _ = generate_greedy_tokens(model, prefix, EL, stop)
        assert toks.shape[0] == px.shape[0]
        assert torch.equal(toks, want_t.cpu()) and torch.equal(lens, want_l.cpu())
    with pytest.raises(ValueError):
        list(pipe.run([synth.pixels(2, vcfg.image_size)]))  # pageable host memory is refused
    assert list(pipe.run([])) == []


def test_partitioned_pipeline_equals_sequential(cuda_device):
    """SM-partitioned serving loop (image tower + mapper + prefill on the large partition, decode of the previous batch on
    the small one, two alternating GPT-2 engines): token ids bit-identical to the single-stream loop and to the direct
    calls, ragged batch included, pinned-host and device-resident inputs alike."""
    from clipcap_b200.inference.base import generate_greedy_tokens
    from clipcap_b200.pipeline import CaptionPipeline
    encode_fn, model, vcfg, stop = _build(cuda_device)
    B, EL = 4, 7
    batches = [synth.pixels(n, vcfg.image_size, seed=70 + i).pin_memory() for i, n in enumerate((4, 4, 3, 4, 4, 4, 2))]
    seq = CaptionPipeline(encode_fn, model, B, vcfg.image_size, EL, stop, cuda_device)
    want = [(t.clone(), l.clone()) for t, l in seq.run(batches)]
    part = CaptionPipeline(encode_fn, model, B, vcfg.image_size, EL, stop, cuda_device, partition_sms=32)
    assert part.partition is not None and sum(part.partition.sms) <= 148 and part.partition.sms[1] >= 32
    for resident in (False, True):
        src = [b.to(cuda_device) for b in batches] if resident else batches
        got = [(t.clone(), l.clone()) for t, l in part.run(src, resident=resident)]
        assert len(got) == len(batches)
        for (gt, gl), (wt, wl) in zip(got, want):
            assert torch.equal(gt, wt) and torch.equal(gl, wl)
    px = batches[2]
    prefix = model.transformer_mapper(encode_fn(px.to(cuda_device)))
    t, l, _ = generate_greedy_tokens(model, prefix, EL, stop)
    assert torch.equal(want[2][0], t.cpu()) and torch.equal(want[2][1], l.cpu())
    # trailing prefill blocks moved to the head of the decode loop (cc_gpt2_set_prefill_defer): same ids, greedy and beam;
    # single-batch run (front and decode both on the whole-device stream) included
    for defer in (1,):  # the test model has two blocks
        moved = CaptionPipeline(encode_fn, model, B, vcfg.image_size, EL, stop, cuda_device, partition_sms=32,
                                prefill_defer=defer)
        got = [(t.clone(), l.clone()) for t, l in moved.run(batches)]
        for (gt, gl), (wt, wl) in zip(got, want):
            assert torch.equal(gt, wt) and torch.equal(gl, wl)
        one = [(t.clone(), l.clone()) for t, l in moved.run(batches[:1])]
        assert torch.equal(one[0][0], want[0][0]) and torch.equal(one[0][1], want[0][1])
    # beam mode through the same loop
    from clipcap_b200.inference.base import generate_beam_tokens
    beam_pipe = CaptionPipeline(encode_fn, model, B, vcfg.image_size, EL, stop, cuda_device, partition_sms=32, mode="beam",
                                beam=3, prefill_defer=1)
    got = [(t.clone(), l.clone()) for t, l in beam_pipe.run(batches[:3])]
    for px, (gt, gl) in zip(batches[:3], got):
        prefix = model.transformer_mapper(encode_fn(px.to(cuda_device)))
        wt, wl, _ = generate_beam_tokens(model, prefix, None, 3, EL, 1.0, stop)
        assert torch.equal(gt, wt.cpu()) and torch.equal(gl, wl.cpu())


def test_two_phase_generate_equals_generate(cuda_device):
    """cc_generate_prefill + cc_generate_decode (the halves the partitioned loop runs on different streams) give exactly
    what cc_generate gives, greedy and beam, also when another prefill has run on a second engine in between."""
    from clipcap_b200.engine import Gpt2Engine
    cfg = R.Gpt2Cfg(d=128, L=2, H=2, V=1003, n_pos=64)
    w = synth.gpt2_weights(cfg, wte_std=0.1)
    B, Tp, EL, stop = 5, 5, 9, 1002
    prefix = (torch.randn(B, Tp, cfg.d, generator=torch.Generator().manual_seed(3)) * 0.5).to(cuda_device)
    other = (torch.randn(B, Tp, cfg.d, generator=torch.Generator().manual_seed(4)) * 0.5).to(cuda_device)
    a = Gpt2Engine(w, cfg.d, cfg.L, cfg.H, cfg.V, cfg.n_pos, max_seqs=B * 3, max_len=Tp + EL, device=cuda_device)
    b = Gpt2Engine(w, cfg.d, cfg.L, cfg.H, cfg.V, cfg.n_pos, max_seqs=B * 3, max_len=Tp + EL, device=cuda_device)
    for mode, beam in (("greedy", 1), ("beam", 3)):
        want = a.generate(prefix, mode, beam, EL, 1.0, stop)
        for _ in range(2):  # second round replays the two captured graphs
            n, tp = a.prefill(prefix, mode, beam, EL, 1.0, stop)
            b.prefill(other, mode, beam, EL, 1.0, stop)
            got = a.decode(n, tp, mode, beam, EL, 1.0, stop)
            assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
            if mode == "beam":
                assert torch.equal(got[2], want[2])
            else:
                assert float(got[2].abs().max()) == 0.0  # greedy returns zero scores, never a previous beam call's
        assert a.last_launches > 0
    with pytest.raises(ValueError):
        a.prefill(prefix, "nucleus", 1, EL, 1.0, stop)


def test_pipeline_with_audio_encoder(cuda_device):
    """BASELINE configs[4] through the same serving loop: mel features -> CLAP tower -> mapper (E = 512) -> greedy decode."""
    from clipcap_b200.encoders import get_encoder
    from clipcap_b200.encoders.config import EncoderConfig
    from clipcap_b200.inference.base import generate_greedy_tokens
    from clipcap_b200.model import ClipCapModelPrefixOnly, Config
    from clipcap_b200.pipeline import CaptionPipeline
    spec, gcfg, mcfg, _, lm_w, g = load_lm_case("tiny_a")
    torch.manual_seed(3)
    encode_fn, transform = get_encoder("clap", "", device=cuda_device)
    cfg = Config(language_model=spec, prefix_length=mcfg.K, projection_length=mcfg.P, transformer_layers=mcfg.L,
                 transformer_attention_heads=mcfg.H, encoder_config=EncoderConfig(encoder_model_name="clap",
                                                                                 encoder_embedding_size=512))
    model = ClipCapModelPrefixOnly(cfg)
    model.load_state_dict({f"language_model.{k}": v for k, v in lm_w.items()}, strict=False)
    model = model.eval().to(cuda_device)
    EL, stop = 6, int(g["stop_token"])
    gen = torch.Generator().manual_seed(8)
    batches = [transform(torch.randn(n, 4, 1001, 64, generator=gen)).pin_memory() for n in (3, 3, 2)]
    pipe = CaptionPipeline(encode_fn, model, 3, entry_length=EL, stop_token=stop, device=cuda_device, input_shape=(4, 1001, 64))
    assert pipe.h2d_bytes_per_batch == 3 * 4 * 1001 * 64 * 4
    got = [(t.clone(), l.clone()) for t, l in pipe.run(batches)]
    assert len(got) == 3
    for mel, (toks, lens) in zip(batches, got):
        emb = encode_fn(mel.to(cuda_device))
        assert tuple(emb.shape) == (mel.shape[0], 512)
        want_t, want_l, _ = generate_greedy_tokens(model, model.transformer_mapper(emb), EL, stop)
        assert torch.equal(toks, want_t.cpu()) and torch.equal(lens, want_l.cpu())


def test_encoder_mapper_and_writer(cuda_device, tmp_path):
    from clipcap_b200.preprocess import EncoderMapper, NumpyWriter
    encode_fn, model, vcfg, _ = _build(cuda_device)
    mapper = EncoderMapper(encode_fn, device="cuda")
    writer = NumpyWriter(0, str(tmp_path / "out"), 1)
    refs = []
    for i in range(3):
        px = synth.pixels(2, vcfg.image_size, seed=70 + i)
        out = mapper({"data_tensor": px, "text": [f"a{i}", f"b{i}"]})
        assert out["embeddings"].shape == (2, 64) and out["text"] == [f"a{i}", f"b{i}"]
        refs.append(encode_fn(px.to(cuda_device)).cpu().numpy())
        writer(out)
    writer.flush()
    emb = np.load(tmp_path / "out" / "embeddings" / "embeds_0.npy")
    assert np.array_equal(emb, np.concatenate(refs))
