"""Per-stage error of the CLAP CUDA path against the oracle + a timing of the full tower (development aid)."""
import os
import sys
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _ROOT)
sys.path.insert(0, os.path.join(_ROOT, "tests"))
import torch
from oracle import restate_clap as RC
from test_clap_gpu import _weights, _oracle_stages, _engine
from conftest import rel_err

cfg = RC.ClapCfg()
w = _weights(cfg, 0)
mel = torch.randn(2, 4, 1001, 64, generator=torch.Generator().manual_seed(1))
with torch.no_grad():
    stages = _oracle_stages(w, mel, cfg)
    want = RC.clap_audio_embed(w, mel, torch.zeros(2, 1, dtype=torch.bool), cfg)
eng = _engine(w, cfg, 64, "cuda:0")
for s, ref in enumerate(stages):
    got = eng.forward(mel.cuda(), stop_after_stage=s).cpu()
    print("stage", s, "rel", rel_err(got, ref), "finite", bool(torch.isfinite(got).all()), flush=True)
got = eng.forward(mel.cuda()).cpu()
print("embed rel", rel_err(got, want), "launches", eng.last_launches, flush=True)
big = torch.randn(64, 4, 1001, 64, device="cuda")
for _ in range(3):
    eng.forward(big)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    eng.forward(big)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"B=64: {ms:.2f} ms per forward, {64 / ms * 1e3:.0f} clips/s", flush=True)
