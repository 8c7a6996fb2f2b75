"""Pins the training-step oracle (oracle/restate.py training_loss / training_loss_and_grads / adamw_reference): against
the committed fixtures of the reference's own training_step + loss.backward() (tests/golden/train_*.npz), against the
live reference where /root/reference exists, and against torch.optim.AdamW."""
import pytest
import torch

from conftest import rel_err
from golden_util import LM_CASES, load_train_case
from oracle import ref_runner as RR
from oracle import restate as R
from oracle import synth

TOL = 5e-5  # fp32 re-association between the restatement's autograd graph and the reference's


@pytest.mark.parametrize("name", list(LM_CASES))
def test_training_step_matches_golden(name):
    spec, gcfg, mcfg, map_w, lm_w, tokens, emb, loss, grads, norms = load_train_case(name)
    got_loss, got = R.training_loss_and_grads(map_w, lm_w, mcfg, gcfg, tokens, emb)
    assert abs(got_loss - loss) < TOL * abs(loss)
    assert set(got) == set(norms)
    for k, g in grads.items():
        assert rel_err(got[k], g) < TOL, k
    for k, n in norms.items():
        assert abs(float(got[k].norm()) - n) < 1e-4 * max(n, 1e-8), k


def test_padding_and_zero_tokens_are_ignored():
    """model.py:103-110: -1 padding becomes token 0 and ignore_index=0 drops it — and any real token with id 0."""
    spec, gcfg, mcfg, map_w, lm_w, tokens, emb, loss, _, _ = load_train_case("tiny_a")
    t2 = tokens.clone()
    t2[t2 < 0] = 0
    assert abs(float(R.training_loss(map_w, lm_w, mcfg, gcfg, t2, emb)) - loss) < TOL * loss
    t3 = tokens.clone()
    t3[0, -1] = -1  # one fewer scored token changes the mean
    assert abs(float(R.training_loss(map_w, lm_w, mcfg, gcfg, t3, emb)) - loss) > 1e-4


@pytest.mark.skipif(not RR.available(), reason="needs /root/reference (build container only)")
def test_training_restatement_vs_live_reference():
    spec, gcfg, mcfg, map_w, lm_w, tokens, emb, _, _, _ = load_train_case("tiny_b")
    model = RR.build_reference_model(spec, mcfg.E, mcfg.K, mcfg.P, mcfg.H, mcfg.L, map_w, lm_w)
    tokens = tokens.clone()
    tokens[0, 3:] = -1  # a different padding pattern than the fixture
    ref_loss, ref_grads = RR.reference_training_step(model, tokens, emb)
    loss, grads = R.training_loss_and_grads(map_w, lm_w, mcfg, gcfg, tokens, emb)
    assert abs(loss - ref_loss) < TOL * abs(ref_loss)
    assert set(grads) == set(ref_grads)
    for k in grads:
        assert rel_err(grads[k], ref_grads[k]) < TOL, k


def test_adamw_restatement_matches_torch():
    g0 = torch.Generator().manual_seed(5)
    p = torch.randn(1000, generator=g0)
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([ref], lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 6):
        g = torch.randn(1000, generator=g0)
        ref.grad = g.clone()
        opt.step()
        p, m, v = R.adamw_reference(p, g, m, v, 3e-3, 0.9, 0.95, 1e-8, 0.1, step)
        assert rel_err(p, ref.data) < 1e-6


def test_schedule_matches_transformers():
    from transformers import get_linear_schedule_with_warmup
    from clipcap_b200.model.optim import linear_schedule_with_warmup
    a = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=2e-5)
    b = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=2e-5)
    sa, sb = get_linear_schedule_with_warmup(a, 5, 40), linear_schedule_with_warmup(b, 5, 40)
    for _ in range(45):
        assert abs(sa.get_last_lr()[0] - sb.get_last_lr()[0]) < 1e-12
        a.step(); b.step(); sa.step(); sb.step()


@pytest.mark.skipif(not RR.available(), reason="needs /root/reference (build container only)")
@pytest.mark.parametrize("use_pos", [True, False])
def test_windowed_mapper_training_restatement_vs_live_reference(use_pos):
    """SURVEY §8f rank 3, not built on the GPU yet: training_step + backward with TransformerMapperWindowed
    (clipcap/model/mapper.py:133-160) — the oracle side of that row, pinned against the reference itself."""
    gcfg = R.Gpt2Cfg(d=128, L=2, H=2, V=1003, n_pos=64)
    mcfg = R.MapperCfg(kind="windowed", E=64, d=128, P=3, K=5, H=2, L=2, W=3, use_pos=use_pos)
    map_w, lm_w = synth.mapper_weights(mcfg, seed=61), synth.gpt2_weights(gcfg, seed=62, wte_std=0.1)
    model = RR.build_reference_model("tiny:128:2:2:1003:64", 64, 5, 3, 2, 2, map_w, lm_w, windowed=True, window_size=2,
                                     use_pos=use_pos)
    emb = synth.embeddings(9, 64, seed=63).view(3, 3, 64)
    g = torch.Generator().manual_seed(64)
    tokens = torch.randint(1, gcfg.V, (3, 9), generator=g)
    tokens[1, 6:] = -1
    tokens[2, 2:] = -1
    ref_loss, ref_grads = RR.reference_training_step(model, tokens, emb)
    loss, grads = R.training_loss_and_grads(map_w, lm_w, mcfg, gcfg, tokens, emb)
    assert abs(loss - ref_loss) < TOL * abs(ref_loss)
    assert set(grads) == set(ref_grads)
    for k in grads:
        assert rel_err(grads[k], ref_grads[k]) < TOL, k
