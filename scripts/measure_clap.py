"""BASELINE configs[4]: CLAP audio tower + TransformerMapper(E=512) + GPT-2-medium, bs=128 synthetic 10 s mel
spectrograms ([128, 4, 1001, 64], no clip longer than the window), 20-token greedy, one B200. Random-init weights of the
named architectures from the product's own parameter containers. CUDA events, 3 warm-up + N timed calls, one JSON line
per measurement. Also times the reference arithmetic of the tower on the host cores (transformers'
ClapAudioModelWithProjection, the stand-in SURVEY §8c names) on a bounded sample.
GPU box only:  python scripts/measure_clap.py [--out profiles/xxx.json]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from clipcap_b200.encoders.clap import ClapAudioTower
from clipcap_b200.encoders.config import EncoderConfig
from clipcap_b200.engine import Gpt2Engine, MapperEngine
from clipcap_b200.model import ClipCapModelPrefixOnly, Config


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
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--cpu-clips", type=int, default=4)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    B = a.batch
    torch.manual_seed(0)
    tower = ClapAudioTower().eval()
    cfg = Config(language_model="gpt2-medium", prefix_length=40, projection_length=10, transformer_layers=8,
                 transformer_attention_heads=8, encoder_config=EncoderConfig(encoder_embedding_size=512))
    model = ClipCapModelPrefixOnly(cfg)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    mapper_w = {k[len("transformer_mapper."):]: v for k, v in sd.items() if k.startswith("transformer_mapper.")}
    lm_w = {k[len("language_model."):]: v for k, v in sd.items() if k.startswith("language_model.")}
    mel_host = torch.randn(B, 4, 1001, 64, generator=torch.Generator().manual_seed(1))
    lines = []

    # host baseline: the stand-in module itself on all cores, bounded sample
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_cpu = min(a.cpu_clips, B)
    with torch.no_grad():
        is_longer = torch.zeros(n_cpu, 1, dtype=torch.bool)
        tower.clap(input_features=mel_host[:n_cpu], is_longer=is_longer)
        t0 = time.perf_counter()
        want = tower.clap(input_features=mel_host[:n_cpu], is_longer=is_longer).audio_embeds
        cpu_s = time.perf_counter() - t0
    lines.append({"config": "CLAP audio tower on the host (transformers ClapAudioModelWithProjection, fp32)",
                  "clips_per_s": n_cpu / cpu_s, "cores": cores, "sample": f"{n_cpu} clips, one call"})

    tower = tower.to(dev)
    mel = mel_host.to(dev)
    eng = tower._get_engine((B,))
    got = eng.forward(mel[:n_cpu]).float().cpu()
    rel = float((got - want).norm() / want.norm())
    ms, p50 = timed(lambda: eng.forward(mel), a.iters)
    lines.append({"config": f"CLAP audio tower only, {B} clips per call", "clips_per_s": B / (ms * 1e-3), "ms": ms,
                  "p50_ms": p50, "launches": eng.last_launches, "rel_err_vs_host_module": rel,
                  "tflops": 11.8e9 * B / (ms * 1e-3) / 1e12})

    one = torch.zeros(B, dtype=torch.bool)
    one[B // 2] = True   # ClapFeatureExtractor flags one random sample when no clip of the batch is longer than the window
    ms1, _ = timed(lambda: eng.forward(mel, is_longer=one), a.iters)
    every = torch.ones(B, dtype=torch.bool)
    msa, _ = timed(lambda: eng.forward(mel, is_longer=every), a.iters)
    lines.append({"config": f"CLAP audio tower, {B} clips, feature fusion on 1 / on all {B} samples", "ms_one": ms1, "ms_all": msa,
                  "clips_per_s_one": B / (ms1 * 1e-3), "clips_per_s_all": B / (msa * 1e-3)})

    mapper = MapperEngine(mapper_w, E=512, d=1024, P=10, K=40, H=8, L=8, max_batch=B, device=dev)
    lm = Gpt2Engine(lm_w, 1024, 24, 16, 50257, 1024, max_seqs=B, max_len=40 + 20, device=dev)

    def step():
        return lm.generate(mapper.forward(eng.forward(mel)), "greedy", 1, 20, 1.0, 50256)

    ms, p50 = timed(step, a.iters)
    lines.append({"config": f"configs[4] CLAP + TransformerMapper(E=512) + GPT-2-medium, bs={B}, 20-token greedy",
                  "captions_per_s": B / (ms * 1e-3), "ms": ms, "p50_ms": p50})

    pinned = mel_host.pin_memory()

    def step_e2e():
        toks = lm.generate(mapper.forward(eng.forward(pinned.to(dev, non_blocking=True))), "greedy", 1, 20, 1.0, 50256)[0]
        return toks.cpu()

    ms, p50 = timed(step_e2e, a.iters)
    lines.append({"config": f"configs[4] end to end from pinned host mel ({pinned.numel() * 4 / 1e6:.0f} MB H2D per step) to host tokens",
                  "captions_per_s": B / (ms * 1e-3), "ms": ms, "p50_ms": p50})
    for ln in lines:
        print(json.dumps(ln))
    if a.out:
        with open(a.out, "w") as f:
            for ln in lines:
                f.write(json.dumps(ln) + "\n")


if __name__ == "__main__":
    main()
