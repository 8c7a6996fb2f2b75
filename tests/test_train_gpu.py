"""Training step on the GPU (cc_train_step / cc_op_adamw through the C ABI) against the CPU oracle — itself pinned
against the reference's training_step + loss.backward() (tests/test_train_cpu.py) — and against the committed fixtures.
Tolerances: loss 1e-3 relative (fp16 GEMM operands, as the forward stages). Gradients, per tensor: ||a-b|| / ||b|| <= 3e-2
and cosine >= 0.9995; element-wise max|a-b| / max|b| <= 3e-2 as well, except for the tensors directly behind the ReLU mask
in the backward pass (mlp.fc1.{weight,bias}, norm2.{weight,bias}): a hidden unit whose pre-activation is within fp16
operand rounding of zero takes the other branch than in the fp32 oracle, which moves single elements of those gradients by
a whole term (measured at GPT-2-small width: 0.1 max element error at 1.8e-2 norm error, independent of the loss scale)."""
import pytest
import torch

from conftest import rel_err
from golden_util import LM_CASES, load_train_case
from oracle import restate as R
from oracle import synth

pytestmark = pytest.mark.gpu
LOSS_TOL, GRAD_TOL, COS_MIN = 1e-3, 3e-2, 0.9995
RELU_GATED = ("mlp.fc1.weight", "mlp.fc1.bias", "norm2.weight", "norm2.bias")


def _cos(a, b):
    a, b = a.float().cpu().flatten(), b.float().cpu().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30))


def _engine(gcfg, mcfg, lm_w, B, Tt, dev):
    from clipcap_b200.engine import TrainEngine
    return TrainEngine(lm_w, E=mcfg.E, d=mcfg.d, P=mcfg.P, K=mcfg.K, H=mcfg.H, L=mcfg.L, lm_layers=gcfg.L, lm_heads=gcfg.H,
                       V=gcfg.V, n_pos=gcfg.n_pos, max_batch=B, max_tokens=Tt, device=dev,
                       kind=getattr(mcfg, "kind", "transformer"), W=getattr(mcfg, "W", 1),
                       use_pos=getattr(mcfg, "use_pos", False))


def _check_grads(got, want, tol=GRAD_TOL, gated_tol=None, cos_min=COS_MIN):
    """Per tensor: ||a-b|| / ||b|| < tol and cosine > cos_min; element-wise max error < tol except behind the ReLU mask.
    `gated_tol`: separate norm tolerance for the ReLU-gated tensors (tiny batches, where one flipped unit is a visible
    share of the gradient)."""
    assert set(got) == set(want)
    worst = 0.0
    for k in sorted(want):
        a, b = got[k].float().cpu(), want[k].float().cpu()
        e, c = rel_err(a, b), _cos(a, b)
        l2 = float((a - b).norm() / b.norm().clamp_min(1e-30))
        worst = max(worst, l2)
        gated = k.endswith(RELU_GATED)
        lim = gated_tol if (gated and gated_tol is not None) else tol
        assert l2 < lim and c > (cos_min if not gated or gated_tol is None else 0.998), f"{k}: norm err {l2:.3e}, cosine {c:.6f}"
        if not gated:
            assert e < tol, f"{k}: max element err {e:.3e}"
    return worst


@pytest.mark.parametrize("name", list(LM_CASES))
def test_train_step_matches_oracle_and_golden(cuda_device, name):
    spec, gcfg, mcfg, map_w, lm_w, tokens, emb, gold_loss, gold_grads, gold_norms = load_train_case(name)
    want_loss, want = R.training_loss_and_grads(map_w, lm_w, mcfg, gcfg, tokens, emb)
    eng = _engine(gcfg, mcfg, lm_w, tokens.shape[0], tokens.shape[1], cuda_device)
    params = {k: v.to(cuda_device).contiguous() for k, v in map_w.items()}
    grads = {k: torch.full_like(v, float("nan")) for k, v in params.items()}
    loss = eng.step(params, emb.to(cuda_device), tokens.to(cuda_device), grads)
    torch.cuda.synchronize()
    assert abs(float(loss) - want_loss) < LOSS_TOL * abs(want_loss)
    assert abs(float(loss) - gold_loss) < LOSS_TOL * abs(gold_loss)
    _check_grads(grads, want)
    for k, g in gold_grads.items():  # the reference's own gradients
        assert rel_err(grads[k], g) < GRAD_TOL, k
    for k, n in gold_norms.items():
        assert abs(float(grads[k].norm()) - n) < 2e-2 * max(n, 1e-8), k
    # forward-only call (validation): same loss, gradients untouched
    loss2 = eng.step(params, emb.to(cuda_device), tokens.to(cuda_device), None)
    assert float(loss2) == float(loss)
    # a smaller batch on the same handle, different loss scale: same numbers as the oracle on that slice
    w_loss, w = R.training_loss_and_grads(map_w, lm_w, mcfg, gcfg, tokens[:2, :5], emb[:2])
    g2 = {k: torch.empty_like(v) for k, v in params.items()}
    l2 = eng.step(params, emb[:2].to(cuda_device), tokens[:2, :5].to(cuda_device), g2, loss_scale=256.0)
    assert abs(float(l2) - w_loss) < LOSS_TOL * abs(w_loss)
    _check_grads(g2, w)


def test_train_step_gpt2_small_shapes(cuda_device):
    """GPT-2-small width (d=768, 12 heads), mapper head dim 96, 3 LM layers / 2 mapper layers, B=6, Tt=21, real vocab."""
    gcfg = R.Gpt2Cfg(d=768, L=3, H=12, V=50257, n_pos=128)
    mcfg = R.MapperCfg(E=512, d=768, P=4, K=10, H=8, L=2)
    map_w, lm_w = synth.mapper_weights(mcfg, 5), synth.gpt2_weights(gcfg, 6)
    g = torch.Generator().manual_seed(9)
    B, Tt = 6, 21
    tokens = torch.randint(1, gcfg.V, (B, Tt), generator=g)
    for b in range(B):
        tokens[b, Tt - 3 * b:] = -1 if b else tokens[b, Tt:]
    emb = synth.embeddings(B, mcfg.E, seed=3)
    want_loss, want = R.training_loss_and_grads(map_w, lm_w, mcfg, gcfg, tokens, emb)
    eng = _engine(gcfg, mcfg, lm_w, B, Tt, cuda_device)
    params = {k: v.to(cuda_device).contiguous() for k, v in map_w.items()}
    grads = {k: torch.empty_like(v) for k, v in params.items()}
    loss = eng.step(params, emb.to(cuda_device), tokens.to(cuda_device), grads)
    assert abs(float(loss) - want_loss) < LOSS_TOL * abs(want_loss)
    _check_grads(grads, want)
    # The ReLU-mask claim at the width where it shows (see test_relu_mask_explains_the_gated_gradient_error): against the
    # oracle that takes the mask decisions of fp16-rounded operands, the gated tensors are as close as the others.
    _, masked = R.training_loss_and_grads(map_w, lm_w, mcfg, gcfg, tokens, emb, relu_mask_fp16=True)

    def l2(a, b):
        return float((a.float().cpu() - b).norm() / b.norm().clamp_min(1e-30))

    gated = [k for k in want if k.endswith(RELU_GATED)]
    worst_plain = max(l2(grads[k], want[k]) for k in gated)
    worst_masked = max(l2(grads[k], masked[k]) for k in gated)
    worst_other = max(l2(grads[k], want[k]) for k in want if k not in gated)
    print(f"GPT-2-small width: gated tensors vs fp32 oracle {worst_plain:.2e}, vs oracle with fp16-operand mask "
          f"{worst_masked:.2e}; other tensors {worst_other:.2e}")
    # The emulated mask only knows the rounding of the fc1 operands, not the (1e-3-level) differences of the LayerNorm
    # output that feeds them, so some border units still differ: the error must shrink, not vanish.
    assert worst_masked < worst_plain and worst_masked < 2e-2, (worst_plain, worst_masked, worst_other)


@pytest.mark.parametrize("cfg", [(2, 107, 16, 64, 1), (3, 50, 8, 128, 0), (2, 33, 4, 96, 0), (1, 160, 2, 64, 1),
                                 (2, 20, 3, 48, 0), (2, 1, 2, 64, 1), (1, 170, 2, 64, 0)])
def test_attention_backward_kernels(cuda_device, cfg):
    """cc_op_attention_bwd vs torch autograd: the warp-MMA kernel (S <= 160) and, for the last shape, the scalar kernel
    that covers what is left."""
    from clipcap_b200 import _ffi
    B, S, H, hd, causal = cfg
    d = H * hd
    g = torch.Generator().manual_seed(S * 7 + hd)
    qkv = (torch.randn(B * S, 3 * d, generator=g) * 0.7).half()
    d_o = (torch.randn(B * S, d, generator=g) * 0.3).half()
    scale = hd ** -0.5
    x = qkv.float().reshape(B, S, 3, H, hd).requires_grad_(True)
    qq, kk, vv = [x[:, :, i].transpose(1, 2) for i in range(3)]
    att = (qq @ kk.transpose(-1, -2)) * scale
    if causal:
        att = att + torch.full((S, S), float("-inf")).triu(1)
    o = (att.softmax(-1) @ vv).transpose(1, 2).reshape(B * S, d)
    o.backward(d_o.float())
    want = x.grad.reshape(B * S, 3 * d)
    qd, dod = qkv.to(cuda_device), d_o.to(cuda_device)
    out = torch.full((B * S, 3 * d), float("nan"), device=cuda_device, dtype=torch.half)
    p = qd.data_ptr()
    _ffi.check(_ffi.lib().cc_op_attention_bwd(p, p + 2 * d, p + 4 * d, 3 * d, dod.data_ptr(), d, out.data_ptr(),
                                              out.data_ptr() + 2 * d, out.data_ptr() + 4 * d, 3 * d, B, S, H, hd, causal,
                                              scale, _ffi.current_stream_ptr()))
    torch.cuda.synchronize()
    assert rel_err(out, want) < 4e-3, cfg


def test_all_padding_gives_nan_like_torch(cuda_device):
    spec, gcfg, mcfg, map_w, lm_w, tokens, emb, *_ = load_train_case("tiny_a")
    eng = _engine(gcfg, mcfg, lm_w, 4, tokens.shape[1], cuda_device)
    params = {k: v.to(cuda_device).contiguous() for k, v in map_w.items()}
    loss = eng.step(params, emb.to(cuda_device), torch.full_like(tokens, -1).to(cuda_device), None)
    assert torch.isnan(loss).item()  # F.cross_entropy with every target ignored


def test_adamw_kernel_matches_torch(cuda_device):
    from clipcap_b200.engine import adamw_update
    g0 = torch.Generator().manual_seed(5)
    p = torch.randn(100003, generator=g0)
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([ref], lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
    pd, m, v = p.to(cuda_device), torch.zeros_like(p, device=cuda_device), torch.zeros_like(p, device=cuda_device)
    for step in range(1, 5):
        g = torch.randn(100003, generator=g0)
        ref.grad = g.clone()
        opt.step()
        adamw_update(pd, g.to(cuda_device), m, v, 3e-3, 0.9, 0.95, 1e-8, 0.1, step)
        assert rel_err(pd, ref.data) < 1e-6


def test_training_step_api_drop_in(cuda_device):
    """The reference's training loop shape: model.training_step(batch, i) -> loss; loss.backward(); optimizer.step();
    scheduler.step() (what Lightning does with configure_optimizers' dict). Three steps against the oracle + torch.optim.AdamW."""
    from clipcap_b200.encoders.config import EncoderConfig
    from clipcap_b200.model import ClipCapModelPrefixOnly, Config, TrainingConfig
    spec, gcfg, mcfg, map_w, lm_w, tokens, emb, *_ = load_train_case("tiny_a")
    cfg = Config(language_model=spec, prefix_length=mcfg.K, projection_length=mcfg.P, transformer_layers=mcfg.L,
                 transformer_attention_heads=mcfg.H, encoder_config=EncoderConfig(encoder_embedding_size=mcfg.E))
    model = ClipCapModelPrefixOnly(cfg)
    sd = {f"transformer_mapper.{k}": v for k, v in map_w.items()}
    sd.update({f"language_model.{k}": v for k, v in lm_w.items()})
    model.load_state_dict(sd, strict=True)
    model = model.to(cuda_device).train()
    assert not model.language_model.training  # model.py:120-123
    model.set_training_config(TrainingConfig(optimizer_lr=1e-3, use_deepspeed_optimisers=False, scheduler_warmup_steps=2,
                                             total_steps=10))
    oc = model.configure_optimizers()
    opt, sched = oc["optimizer"], oc["lr_scheduler"]["scheduler"]
    # oracle side: autograd on CPU + torch.optim.AdamW + the transformers schedule
    from transformers import get_linear_schedule_with_warmup
    ref_p = {k: torch.nn.Parameter(v.clone()) for k, v in map_w.items()}
    ref_opt = torch.optim.AdamW(list(ref_p.values()), lr=1e-3)
    ref_sched = get_linear_schedule_with_warmup(ref_opt, 2, 10)
    for it in range(3):
        batch = (tokens.clone().to(cuda_device), emb.clone().to(cuda_device))
        loss = model.training_step(batch, it)
        opt.zero_grad()
        loss.backward()
        opt.step()
        sched.step()
        ref_opt.zero_grad()
        ref_loss = R.training_loss(ref_p, lm_w, mcfg, gcfg, tokens, emb)
        ref_loss.backward()
        ref_opt.step()
        ref_sched.step()
        assert abs(float(loss) - float(ref_loss)) < 2e-3 * abs(float(ref_loss)), it
    assert (batch[0] >= 0).all()  # padding was rewritten in place like the reference does (model.py:104)
    # Adam normalises every element's step to ~lr whatever the gradient's size, so elements whose gradient is at noise
    # level may step the other way: compare the overall movement, not single elements.
    for k, p in model.transformer_mapper.named_parameters():
        moved_gpu, moved_ref = p.detach().cpu() - map_w[k], ref_p[k].detach() - map_w[k]
        assert _cos(moved_gpu, moved_ref) > 0.97, k
        assert rel_err(p, ref_p[k]) < 5e-2, k
    # the trained mapper is what inference now uses
    with torch.no_grad():
        prefix = model.transformer_mapper(emb.to(cuda_device))
    want = R.mapper_forward({k: v.detach() for k, v in ref_p.items()}, emb, mcfg)
    assert rel_err(prefix, want) < 3e-3


@pytest.mark.parametrize("use_pos", [True, False])
def test_windowed_mapper_train_step_matches_oracle(cuda_device, use_pos):
    """training_step + backward with TransformerMapperWindowed (clipcap/model/mapper.py:133-160 under model.py:22-32,
    94-113): W = window_size + 1 embeddings per sample, the linear layer shared by the windows, the optional learned
    pos_embeddings (gradient = batch sum of the projected-token gradients). The oracle side of this case is pinned against
    the reference's own training_step + backward (tests/test_train_cpu.py, same seeds and shapes)."""
    gcfg = R.Gpt2Cfg(d=128, L=2, H=2, V=1003, n_pos=64)
    mcfg = R.MapperCfg(kind="windowed", E=64, d=128, P=3, K=5, H=2, L=2, W=3, use_pos=use_pos)
    map_w, lm_w = synth.mapper_weights(mcfg, seed=61), synth.gpt2_weights(gcfg, seed=62, wte_std=0.1)
    emb = synth.embeddings(9, 64, seed=63).view(3, 3, 64)
    tokens = torch.randint(1, gcfg.V, (3, 9), generator=torch.Generator().manual_seed(64))
    tokens[1, 6:] = -1
    tokens[2, 2:] = -1
    want_loss, want = R.training_loss_and_grads(map_w, lm_w, mcfg, gcfg, tokens, emb)
    assert ("pos_embeddings" in want) == use_pos
    eng = _engine(gcfg, mcfg, lm_w, 3, 9, cuda_device)
    params = {k: v.to(cuda_device).contiguous() for k, v in map_w.items()}
    grads = {k: torch.full_like(v, float("nan")) for k, v in params.items()}
    loss = eng.step(params, emb.to(cuda_device), tokens.to(cuda_device), grads)
    assert abs(float(loss) - want_loss) < LOSS_TOL * abs(want_loss)
    _check_grads(grads, want, gated_tol=6e-2)   # 3 samples x 14 tokens: one flipped hidden unit is ~4 % of fc1.bias' gradient
    # forward only (validation): same loss, no gradient buffers touched
    assert abs(float(eng.step(params, emb.to(cuda_device), tokens.to(cuda_device))) - want_loss) < LOSS_TOL * abs(want_loss)
    with pytest.raises(ValueError):
        eng.step(params, emb[:, 0].to(cuda_device), tokens.to(cuda_device), grads)   # [B, E] is not a windowed input


@pytest.mark.parametrize("name", list(LM_CASES))
def test_relu_mask_explains_the_gated_gradient_error(cuda_device, name):
    """The tensors directly behind the ReLU mask (mlp.fc1.*, norm2.*) carry 1-4e-2 gradient error against the fp32 oracle
    while every other tensor is at 2-6e-3. Claim: a hidden unit whose pre-activation is within fp16 operand rounding of
    zero takes the other branch. Proof: give the ORACLE the mask decisions of fp16-rounded operands (values and gradients
    still fp32, oracle/restate.py relu_mask_fp16) — the gated tensors then agree as well as all the others."""
    spec, gcfg, mcfg, map_w, lm_w, tokens, emb, *_ = load_train_case(name)
    eng = _engine(gcfg, mcfg, lm_w, tokens.shape[0], tokens.shape[1], cuda_device)
    params = {k: v.to(cuda_device).contiguous() for k, v in map_w.items()}
    grads = {k: torch.empty_like(v) for k, v in params.items()}
    eng.step(params, emb.to(cuda_device), tokens.to(cuda_device), grads)
    _, plain = R.training_loss_and_grads(map_w, lm_w, mcfg, gcfg, tokens, emb)
    _, masked = R.training_loss_and_grads(map_w, lm_w, mcfg, gcfg, tokens, emb, relu_mask_fp16=True)

    def l2(a, b):
        return float((a.float().cpu() - b).norm() / b.norm().clamp_min(1e-30))

    gated = [k for k in plain if k.endswith(RELU_GATED)]
    worst_plain = max(l2(grads[k], plain[k]) for k in gated)
    worst_masked = max(l2(grads[k], masked[k]) for k in gated)
    worst_other = max(l2(grads[k], plain[k]) for k in plain if k not in gated)
    print(f"{name}: gated tensors vs fp32 oracle {worst_plain:.2e}, vs oracle with fp16-operand mask {worst_masked:.2e}; "
          f"other tensors {worst_other:.2e}")
    assert worst_masked < 8e-3, worst_masked
    assert worst_other < 8e-3, worst_other


def test_loss_scale_overflow_is_detected_and_does_not_poison_adamw(cuda_device):
    """A loss scale far too large overflows the fp16 activation gradients: cc_train_step reports the non-finite gradient
    elements, cc_op_adamw leaves those elements (parameter and moments) untouched, and the module's check_overflow()
    halves the scale. With a sane scale the count is zero."""
    from clipcap_b200.engine import adamw_update
    spec, gcfg, mcfg, map_w, lm_w, tokens, emb, *_ = load_train_case("tiny_a")
    eng = _engine(gcfg, mcfg, lm_w, tokens.shape[0], tokens.shape[1], cuda_device)
    params = {k: v.to(cuda_device).contiguous() for k, v in map_w.items()}
    grads = {k: torch.zeros_like(v) for k, v in params.items()}
    eng.step(params, emb.to(cuda_device), tokens.to(cuda_device), grads, loss_scale=1024.0)
    assert eng.last_nonfinite() == 0
    eng.step(params, emb.to(cuda_device), tokens.to(cuda_device), grads, loss_scale=1e30)
    bad = eng.last_nonfinite()
    assert bad > 0 and bad == sum(int((~torch.isfinite(g)).sum()) for g in grads.values())
    # the optimiser skips exactly the non-finite elements
    p = torch.ones(8, device=cuda_device)
    g = torch.tensor([0.5, float("nan"), float("inf"), -0.5, 0.0, float("-inf"), 1.0, 2.0], device=cuda_device)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    adamw_update(p, g, m, v, 1e-2, 0.9, 0.999, 1e-8, 0.1, 1)
    skipped = ~torch.isfinite(g)
    assert torch.equal(p[skipped], torch.ones(3, device=cuda_device)) and float(m[skipped].abs().max()) == 0.0
    assert torch.isfinite(p).all() and torch.isfinite(m).all() and torch.isfinite(v).all() and (p[~skipped] != 1.0).all()


def test_windowed_model_training_step_api(cuda_device):
    """ClipCapModelPrefixOnly with use_windowed_embeddings (model.py:22-32): training_step -> backward -> optimizer step
    through the module API, loss against the oracle."""
    from clipcap_b200.encoders.config import EncoderConfig
    from clipcap_b200.model import ClipCapModelPrefixOnly, Config, TrainingConfig
    gcfg = R.Gpt2Cfg(d=128, L=2, H=2, V=1003, n_pos=64)
    mcfg = R.MapperCfg(kind="windowed", E=64, d=128, P=3, K=5, H=2, L=2, W=3, use_pos=True)
    map_w, lm_w = synth.mapper_weights(mcfg, seed=61), synth.gpt2_weights(gcfg, seed=62, wte_std=0.1)
    cfg = Config(language_model="tiny:128:2:2:1003:64", prefix_length=5, projection_length=3, transformer_layers=2,
                 transformer_attention_heads=2, use_positional_embeddings=True,
                 encoder_config=EncoderConfig(encoder_embedding_size=64, use_windowed_embeddings=True, window_size=2))
    model = ClipCapModelPrefixOnly(cfg)
    sd = {f"transformer_mapper.{k}": v for k, v in map_w.items()}
    sd.update({f"language_model.{k}": v for k, v in lm_w.items()})
    model.load_state_dict(sd, strict=True)
    model = model.to(cuda_device).train()
    model.set_training_config(TrainingConfig(optimizer_lr=1e-3, use_deepspeed_optimisers=False, scheduler_warmup_steps=0,
                                             total_steps=4))
    opt = model.configure_optimizers()["optimizer"]
    emb = synth.embeddings(9, 64, seed=63).view(3, 3, 64)
    tokens = torch.randint(1, gcfg.V, (3, 9), generator=torch.Generator().manual_seed(64))
    want_loss, _ = R.training_loss_and_grads(map_w, lm_w, mcfg, gcfg, tokens, emb)
    loss = model.training_step((tokens.clone().to(cuda_device), emb.to(cuda_device)), 0)
    assert abs(float(loss) - want_loss) < LOSS_TOL * abs(want_loss)
    opt.zero_grad()
    loss.backward()
    assert model.transformer_mapper.pos_embeddings.grad is not None
    opt.step()
    loss2 = model.training_step((tokens.clone().to(cuda_device), emb.to(cuda_device)), 1)
    assert float(loss2) < float(loss)   # one AdamW step on the same batch lowers its loss


def test_optimizer_step_invalidates_cached_engines(cuda_device):
    """forward -> optimizer.step -> forward: the mapper engine built BEFORE the step (a validation pass, a caption callback)
    must not serve stale weights afterwards. FusedAdamW updates through a raw pointer, so it bumps the parameters'
    version counters itself (EngineModule keys its cached engine on them)."""
    from clipcap_b200.encoders.config import EncoderConfig
    from clipcap_b200.model import ClipCapModelPrefixOnly, Config, TrainingConfig
    spec, gcfg, mcfg, map_w, lm_w, tokens, emb, *_ = load_train_case("tiny_a")
    cfg = Config(language_model=spec, prefix_length=mcfg.K, projection_length=mcfg.P, transformer_layers=mcfg.L,
                 transformer_attention_heads=mcfg.H, encoder_config=EncoderConfig(encoder_embedding_size=mcfg.E))
    model = ClipCapModelPrefixOnly(cfg)
    sd = {f"transformer_mapper.{k}": v for k, v in map_w.items()}
    sd.update({f"language_model.{k}": v for k, v in lm_w.items()})
    model.load_state_dict(sd, strict=True)
    model = model.to(cuda_device).train()
    model.set_training_config(TrainingConfig(optimizer_lr=5e-2, use_deepspeed_optimisers=False, scheduler_warmup_steps=0,
                                             total_steps=10))
    opt = model.configure_optimizers()["optimizer"]
    with torch.no_grad():
        before = model.transformer_mapper(emb.to(cuda_device)).clone()   # builds and caches the mapper engine
    assert rel_err(before, R.mapper_forward(map_w, emb, mcfg)) < 3e-3
    versions = [p._version for p in model.transformer_mapper.parameters()]
    loss = model.training_step((tokens.clone().to(cuda_device), emb.clone().to(cuda_device)), 0)
    opt.zero_grad()
    loss.backward()
    opt.step()
    assert all(p._version > v for p, v in zip(model.transformer_mapper.parameters(), versions))
    with torch.no_grad():
        after = model.transformer_mapper(emb.to(cuda_device))
    new_w = {k: p.detach().cpu() for k, p in model.transformer_mapper.named_parameters()}
    assert rel_err(after, before) > 1e-2                                  # the step really moved the output
    assert rel_err(after, R.mapper_forward(new_w, emb, mcfg)) < 3e-3      # and it is the updated weights' output


@pytest.mark.parametrize("B,Tt", [(1, 1), (5, 9), (2, 59)])
def test_train_step_edge_shapes(cuda_device, B, Tt):
    """One sample with one token; a batch with a fully padded row and a row of token-0 only; captions that fill the model's
    positions (K + Tt = n_positions). Loss and every gradient against the oracle."""
    spec, gcfg, mcfg, map_w, lm_w, *_ = load_train_case("tiny_a")   # K = 5, n_pos = 64
    g = torch.Generator().manual_seed(B * 100 + Tt)
    tokens = torch.randint(1, gcfg.V, (B, Tt), generator=g)
    if B >= 5:
        tokens[1, :] = -1                 # nothing to score in this row
        tokens[2, :] = 0                  # real tokens with id 0: ignored as well (ignore_index = 0)
        tokens[3, Tt // 2:] = -1
    emb = synth.embeddings(B, mcfg.E, seed=Tt)
    want_loss, want = R.training_loss_and_grads(map_w, lm_w, mcfg, gcfg, tokens, emb)
    eng = _engine(gcfg, mcfg, lm_w, B, Tt, cuda_device)
    params = {k: v.to(cuda_device).contiguous() for k, v in map_w.items()}
    grads = {k: torch.full_like(v, float("nan")) for k, v in params.items()}
    loss = eng.step(params, emb.to(cuda_device), tokens.to(cuda_device), grads)
    assert abs(float(loss) - want_loss) < LOSS_TOL * abs(want_loss)
    _check_grads(grads, want)
