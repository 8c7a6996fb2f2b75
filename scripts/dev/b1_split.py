"""Developer timing of the single-image path (configs[0]-like and the reference's default call: beam 5, entry_length 67)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from clipcap_b200.engine import Gpt2Engine, MapperEngine, VitEngine
from oracle import restate as R, synth
dev = torch.device("cuda:0")
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
vit = VitEngine(synth.vit_weights(R.VitCfg()), max_batch=8, device=dev)
px = synth.pixels(1, 224).to(dev)
print("ViT-L/14 B=1 ms", t(lambda: vit.forward(px)))
for name, g, mc in (("gpt2-small", R.Gpt2Cfg(d=768, L=12, H=12), R.MapperCfg(E=768, d=768, P=10, K=10, H=8, L=8)),
                    ("gpt2-medium", R.Gpt2Cfg(), R.MapperCfg(E=768, d=1024, P=10, K=40, H=8, L=8))):
    mp = MapperEngine(synth.mapper_weights(mc, 3), E=768, d=mc.d, P=mc.P, K=mc.K, H=8, L=8, max_batch=8, device=dev)
    lm = Gpt2Engine(synth.gpt2_weights(g, 4), g.d, g.L, g.H, g.V, g.n_pos, max_seqs=8, max_len=mc.K + 67, device=dev)
    emb = torch.randn(1, 768, device=dev)
    print(name, "mapper B=1 ms", t(lambda: mp.forward(emb)))
    prefix = mp.forward(emb)
    for mode, beam, el in (("greedy", 1, 1), ("greedy", 1, 20), ("greedy", 1, 67), ("beam", 5, 20), ("beam", 5, 67)):
        print(name, mode, beam, "EL", el, "ms", t(lambda: lm.generate(prefix, mode, beam, el, 1.0, 50256), 5))
    del lm, mp
