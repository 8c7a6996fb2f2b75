"""Developer profile: one beam-5 generate call at the configs[3] per-GPU shard (256 images) between cudaProfilerStart/Stop."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from clipcap_b200.engine import Gpt2Engine
from oracle import restate as R
dev = torch.device("cuda:0")
state = bench.synthetic_state()
g = R.Gpt2Cfg()
B = int(os.environ.get("B", "256"))
lm = Gpt2Engine(state["lm"], g.d, g.L, g.H, g.V, g.n_pos, max_seqs=B * 5, max_len=60, device=dev)
prefix = torch.randn(B, 40, 1024, device=dev)
for _ in range(2): lm.generate(prefix, "beam", 5, 20, 1.0, 50256)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
lm.generate(prefix, "beam", 5, 20, 1.0, 50256)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(3): lm.generate(prefix, "beam", 5, 20, 1.0, 50256)
e1.record(); torch.cuda.synchronize(); print("beam generate ms", e0.elapsed_time(e1) / 3)
