"""Developer timing: cc_generate at the bench shape for several entry lengths (separates prefill from the decode steps)."""
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
lm = Gpt2Engine(state["lm"], g.d, g.L, g.H, g.V, g.n_pos, max_seqs=B, max_len=40 + 20, device=dev)
prefix = torch.randn(B, 40, 1024, device=dev)
def t(fn, n=5):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
res = {}
for el in (1, 2, 11, 20):
    res[el] = t(lambda: lm.generate(prefix, "greedy", 1, el, 1.0, 50256))
    print(f"generate EL={el}: {res[el]:.3f} ms")
print(f"prefill+head {res[1]:.3f} ms; decode step (avg over 19) {(res[20]-res[1])/19*1e3:.1f} us; first step {(res[2]-res[1])*1e3:.1f} us")
