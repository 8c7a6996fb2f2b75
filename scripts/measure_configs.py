"""Per-GPU measurements of the BASELINE.json configs that bench.py does not time (bench.py = configs[1]):
  configs[0]  B=1, TransformerMapper(K=P=10) + GPT-2-small, 20-token greedy  -> latency
  configs[2]  ViT-L/14 encode only, 128 images per GPU                        -> images/s
  configs[3]  B=256 per GPU, beam=5, entry_length=20, GPT-2-medium             -> captions/s
Synthetic seeded weights/pixels, CUDA events, 3 warm-up + N timed calls. Prints one JSON line per config.
GPU box only:  python scripts/measure_configs.py [--out profiles/xxx.json]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from clipcap_b200.engine import Gpt2Engine, MapperEngine, VitEngine
from oracle import restate as R
from oracle import synth


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    evs[0].record()
    for i in range(n):
        fn()
        evs[i + 1].record()
    torch.cuda.synchronize()
    per = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(n))
    return evs[0].elapsed_time(evs[n]) / n, per[len(per) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    lines = []
    state = bench.synthetic_state()
    vcfg = R.VitCfg()
    vit = VitEngine(state["vit"], vcfg.image_size, vcfg.patch, vcfg.width, vcfg.layers, vcfg.heads, vcfg.mlp_dim,
                    vcfg.out_dim, max_batch=256, device=dev)

    # ---- configs[2]: ViT-L/14 encode only, 128 images per GPU (1024 over 8 GPUs)
    px = synth.pixels(128, 224).to(dev)
    ms, p50 = timed(lambda: vit.forward(px), a.iters)
    flops = bench.flops_per_caption()["vit"] * 128
    lines.append({"config": "configs[2] ViT-L/14 encode only, 128 images per GPU", "images_per_s": 128 / (ms * 1e-3),
                  "ms": ms, "p50_ms": p50, "tflops": flops / (ms * 1e-3) / 1e12})
    px256 = synth.pixels(256, 224).to(dev)
    ms, p50 = timed(lambda: vit.forward(px256), a.iters)
    lines.append({"config": "ViT-L/14 encode only, 256 images per call", "images_per_s": 256 / (ms * 1e-3), "ms": ms,
                  "p50_ms": p50, "tflops": bench.flops_per_caption()["vit"] * 256 / (ms * 1e-3) / 1e12})

    # ---- configs[3]: per-GPU shard of bs=2048 over 8 GPUs = 256 images, beam 5
    mcfg = R.MapperCfg(E=768, d=1024, P=10, K=40, H=8, L=8)
    mapper = MapperEngine(state["mapper"], E=768, d=1024, P=10, K=40, H=8, L=8, max_batch=256, device=dev)
    g = R.Gpt2Cfg()
    lm = Gpt2Engine(state["lm"], g.d, g.L, g.H, g.V, g.n_pos, max_seqs=256 * 5, max_len=40 + 20, device=dev)

    def beam_step():
        emb = vit.forward(px256)
        prefix = mapper.forward(emb)
        return lm.generate(prefix, "beam", 5, 20, 1.0, 50256)

    ms, p50 = timed(beam_step, a.iters)
    prefix = mapper.forward(vit.forward(px256))
    ms_gen, _ = timed(lambda: lm.generate(prefix, "beam", 5, 20, 1.0, 50256), a.iters)
    lines.append({"config": "configs[3] per-GPU shard: B=256, beam=5, entry_length=20, GPT-2-medium",
                  "captions_per_s": 256 / (ms * 1e-3), "ms": ms, "p50_ms": p50, "generate_ms": ms_gen})
    ms_g, _ = timed(lambda: lm.generate(prefix, "greedy", 1, 20, 1.0, 50256), a.iters)
    lines.append({"config": "configs[1] generate stage only: B=256 greedy", "generate_ms": ms_g})
    del lm, mapper
    torch.cuda.empty_cache()

    # ---- configs[0]: B=1, GPT-2-small, TransformerMapper P=K=10 (the MLP mapper is absent from the reference)
    mc = R.MapperCfg(E=768, d=768, P=10, K=10, H=8, L=8)
    gs = R.Gpt2Cfg(d=768, L=12, H=12)
    mw, lw = synth.mapper_weights(mc, 11), synth.gpt2_weights(gs, 12)
    mapper1 = MapperEngine(mw, E=768, d=768, P=10, K=10, H=8, L=8, max_batch=8, device=dev)
    lm1 = Gpt2Engine(lw, gs.d, gs.L, gs.H, gs.V, gs.n_pos, max_seqs=8, max_len=10 + 20, device=dev)
    px1 = synth.pixels(1, 224).to(dev)

    def one():
        return lm1.generate(mapper1.forward(vit.forward(px1)), "greedy", 1, 20, 1.0, 50256)

    ms, p50 = timed(one, 30)
    lines.append({"config": "configs[0] B=1, TransformerMapper(K=P=10) + GPT-2-small, 20-token greedy",
                  "latency_ms": ms, "p50_ms": p50, "captions_per_s": 1e3 / ms})
    for ln in lines:
        print(json.dumps(ln))
    if a.out:
        with open(a.out, "w") as f:
            for ln in lines:
                f.write(json.dumps(ln) + "\n")


if __name__ == "__main__":
    main()
