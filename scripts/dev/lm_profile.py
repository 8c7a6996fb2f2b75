"""Developer profile (GPU box): one cc_generate at the bench shape between cudaProfilerStart/Stop (use with
`ncu --profile-from-start off`); CLIPCAP_B200_SM_BUDGET sizes the launches for a partition."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from clipcap_b200.engine import Gpt2Engine
from oracle import restate as R
dev = torch.device("cuda:0")
state = bench.synthetic_state()
g = R.Gpt2Cfg()
B = int(os.environ.get("B", "256")); EL = int(os.environ.get("EL", "3"))
lm = Gpt2Engine(state["lm"], g.d, g.L, g.H, g.V, g.n_pos, max_seqs=B, max_len=60, device=dev)
prefix = torch.randn(B, 40, 1024, device=dev)
for _ in range(2): lm.generate(prefix, "greedy", 1, EL, 1.0, 50256)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
lm.generate(prefix, "greedy", 1, EL, 1.0, 50256)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
