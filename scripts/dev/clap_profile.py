"""Two B=64 forwards of the CLAP tower for an ncu launch list (development aid)."""
import sys
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import torch
from oracle import restate_clap as RC
from test_clap_gpu import _weights, _engine

cfg = RC.ClapCfg()
eng = _engine(_weights(cfg, 0), cfg, 64, "cuda:0")
big = torch.randn(64, 1, 1001, 64, device="cuda")
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(2):
    eng.forward(big)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
