"""Two forwards of the CLAP tower at 128 clips (the second with one fused sample) for an ncu launch list."""
import os
import sys
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _ROOT)
sys.path.insert(0, os.path.join(_ROOT, "tests"))
import torch
from oracle import restate_clap as RC
from test_clap_gpu import _weights, _engine

cfg = RC.ClapCfg()
eng = _engine(_weights(cfg, 0), cfg, 128, "cuda:0")
big = torch.randn(128, 4, 1001, 64, device="cuda")
flags = torch.zeros(128, dtype=torch.bool)
flags[5] = True
eng.forward(big, is_longer=flags)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.forward(big)
eng.forward(big, is_longer=flags)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
