"""Developer view: per-tensor gradient errors of cc_train_step vs the CPU oracle at GPT-2-small width."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import torch
from oracle import restate as R, synth
from clipcap_b200.engine import TrainEngine
dev = torch.device("cuda:0")
gcfg = R.Gpt2Cfg(d=768, L=3, H=12, V=50257, n_pos=128)
mcfg = R.MapperCfg(E=512, d=768, P=4, K=10, H=8, L=2)
map_w, lm_w = synth.mapper_weights(mcfg, 5), synth.gpt2_weights(gcfg, 6)
g = torch.Generator().manual_seed(9)
B, Tt = 6, 21
tokens = torch.randint(1, gcfg.V, (B, Tt), generator=g)
for b in range(1, B): tokens[b, Tt - 3 * b:] = -1
emb = synth.embeddings(B, mcfg.E, seed=3)
want_loss, want = R.training_loss_and_grads(map_w, lm_w, mcfg, gcfg, tokens, emb)
eng = TrainEngine(lm_w, E=mcfg.E, d=mcfg.d, P=mcfg.P, K=mcfg.K, H=mcfg.H, L=mcfg.L, lm_layers=gcfg.L, lm_heads=gcfg.H, V=gcfg.V, n_pos=gcfg.n_pos, max_batch=B, max_tokens=Tt, device=dev)
params = {k: v.to(dev).contiguous() for k, v in map_w.items()}
for ls in (1024.0, 65536.0, 16.0):
    grads = {k: torch.empty_like(v) for k, v in params.items()}
    loss = eng.step(params, emb.to(dev), tokens.to(dev), grads, loss_scale=ls)
    print(f"loss_scale {ls}: loss {float(loss):.6f} want {want_loss:.6f}")
    for k in sorted(want):
        a, b = grads[k].cpu().flatten(), want[k].flatten()
        e = float((a - b).abs().max() / b.abs().max()); c = float(torch.dot(a, b) / (a.norm() * b.norm()))
        rn = float((a - b).norm() / b.norm())
        print(f"   {k:55s} max-rel {e:.3e}  l2-rel {rn:.3e}  cos {c:.6f}")
