"""Error behaviour of the C ABI on a device: every misuse is a status + message (no crash, no silent fallback), and the
handle stays usable afterwards."""
import pytest
import torch

from oracle import restate as R
from oracle import synth

pytestmark = pytest.mark.gpu


def test_capacity_and_shape_errors(cuda_device):
    from clipcap_b200._ffi import CCError
    from clipcap_b200.engine import Gpt2Engine, MapperEngine, VitEngine
    gcfg = R.Gpt2Cfg(d=128, L=2, H=2, V=1003, n_pos=64)
    mcfg = R.MapperCfg(E=64, d=128, P=3, K=5, H=2, L=2)
    vcfg = R.VitCfg(image_size=28, patch=14, width=128, layers=2, heads=2, mlp_dim=256, out_dim=64)
    lm = Gpt2Engine(synth.gpt2_weights(gcfg), gcfg.d, gcfg.L, gcfg.H, gcfg.V, gcfg.n_pos, max_seqs=6, max_len=20,
                    device=cuda_device)
    mapper = MapperEngine(synth.mapper_weights(mcfg), E=64, d=128, P=3, K=5, H=2, L=2, max_batch=4, device=cuda_device)
    vit = VitEngine(synth.vit_weights(vcfg), vcfg.image_size, vcfg.patch, vcfg.width, vcfg.layers, vcfg.heads, vcfg.mlp_dim,
                    vcfg.out_dim, max_batch=2, device=cuda_device)
    prefix = torch.randn(2, 5, 128, device=cuda_device)
    good = lm.generate(prefix, "greedy", 1, 8, 1.0, 1002)[0].clone()

    with pytest.raises(CCError, match="CC_ESHAPE"):          # more sequences than the handle was created for
        lm.generate(torch.randn(7, 5, 128, device=cuda_device), "greedy", 1, 8, 1.0, 1002)
    with pytest.raises(CCError, match="CC_ESHAPE"):          # 2 images x 5 beams > max_seqs 6
        lm.generate(prefix, "beam", 5, 8, 1.0, 1002)
    with pytest.raises(CCError, match="CC_ESHAPE"):          # prefix + generated positions exceed max_len
        lm.generate(prefix, "greedy", 1, 17, 1.0, 1002)
    with pytest.raises(CCError, match="CC_ESHAPE"):          # beam size beyond the kernel's limit
        lm.generate(prefix[:1], "beam", 9, 8, 1.0, 1002)
    with pytest.raises(ValueError):                          # unknown decode mode (host side)
        lm.generate(prefix, "viterbi", 1, 8, 1.0, 1002)
    with pytest.raises(ValueError):                          # wrong embedding width
        lm.generate(torch.randn(2, 5, 64, device=cuda_device), "greedy", 1, 8, 1.0, 1002)
    with pytest.raises(RuntimeError, match="no CPU path"):   # CPU tensor: there is no fallback
        lm.generate(prefix.cpu(), "greedy", 1, 8, 1.0, 1002)
    with pytest.raises(TypeError):                           # fp64 at the boundary
        mapper.forward(torch.randn(2, 64, device=cuda_device, dtype=torch.float64))
    with pytest.raises(CCError, match="CC_ESHAPE"):
        mapper.forward(torch.randn(5, 64, device=cuda_device))
    with pytest.raises(ValueError):
        mapper.forward(torch.randn(2, 65, device=cuda_device))
    with pytest.raises(CCError, match="CC_ESHAPE"):
        vit.forward(torch.randn(3, 3, 28, 28, device=cuda_device))
    with pytest.raises(ValueError):
        vit.forward(torch.randn(2, 3, 32, 32, device=cuda_device))
    # the handle still works and gives the same answer
    again = lm.generate(prefix, "greedy", 1, 8, 1.0, 1002)[0]
    assert torch.equal(good, again)


def test_create_errors(cuda_device):
    from clipcap_b200._ffi import CCError
    from clipcap_b200.engine import Gpt2Engine, MapperEngine, TrainEngine
    gcfg = R.Gpt2Cfg(d=128, L=2, H=2, V=1003, n_pos=64)
    w = synth.gpt2_weights(gcfg)
    with pytest.raises(CCError, match="CC_ESHAPE"):          # head dim 32
        Gpt2Engine(w, 128, 2, 4, 1003, 64, max_seqs=2, max_len=8, device=cuda_device)
    with pytest.raises(CCError, match="CC_ESHAPE"):          # max_len beyond n_positions
        Gpt2Engine(w, 128, 2, 2, 1003, 64, max_seqs=2, max_len=65, device=cuda_device)
    missing = {k: v for k, v in w.items() if k != "transformer.h.1.mlp.c_fc.bias"}
    with pytest.raises(CCError, match="c_fc.bias"):          # a missing tensor is named
        Gpt2Engine(missing, 128, 2, 2, 1003, 64, max_seqs=2, max_len=8, device=cuda_device)
    bad = dict(w)
    bad["transformer.ln_f.weight"] = torch.ones(64)
    with pytest.raises(CCError, match="ln_f.weight"):        # wrong element count is named
        Gpt2Engine(bad, 128, 2, 2, 1003, 64, max_seqs=2, max_len=8, device=cuda_device)
    mw = synth.mapper_weights(R.MapperCfg(E=64, d=128, P=3, K=5, H=2, L=2))
    with pytest.raises(CCError, match="CC_ESHAPE"):          # mapper head dim 32
        MapperEngine(mw, E=64, d=128, P=3, K=5, H=4, L=2, max_batch=2, device=cuda_device)
    with pytest.raises(CCError, match="CC_ESHAPE"):          # training: prefix + tokens beyond n_positions
        TrainEngine(w, E=64, d=128, P=3, K=5, H=2, L=2, lm_layers=2, lm_heads=2, V=1003, n_pos=64, max_batch=2,
                    max_tokens=60, device=cuda_device)
