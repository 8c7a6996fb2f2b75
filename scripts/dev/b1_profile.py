"""Developer profile: one single-image greedy generate (GPT-2-medium, 20 tokens) between cudaProfilerStart/Stop."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from clipcap_b200.engine import Gpt2Engine
from oracle import restate as R, synth
dev = torch.device("cuda:0")
g = R.Gpt2Cfg()
M = int(os.environ.get("M", "1"))
lm = Gpt2Engine(synth.gpt2_weights(g, 4), g.d, g.L, g.H, g.V, g.n_pos, max_seqs=16, max_len=60, device=dev)
prefix = torch.randn(M, 40, 1024, device=dev)
for _ in range(2): lm.generate(prefix, "greedy", 1, 20, 1.0, 50256)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
lm.generate(prefix, "greedy", 1, 20, 1.0, 50256)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
