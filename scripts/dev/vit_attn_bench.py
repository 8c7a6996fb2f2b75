"""Developer timing (GPU box): the tcgen05 ViT attention kernel at the bench shape (B=256 images, 16 heads, S=257) on the
head-major layout the engine uses (via a 1-layer VitEngine is overkill: cc_op_attention on packed QKV is the same kernel)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from clipcap_b200 import _ffi
lib = _ffi.lib()
dev = "cuda"
B, S, H = int(os.environ.get("B", "256")), 257, 16
d = H * 64
S_ = lambda: torch.cuda.current_stream().cuda_stream
qkv = (torch.randn(B * S, 3 * d, device=dev) * 0.5).half()
o = torch.empty(B * S, d, device=dev, dtype=torch.half)
def run():
    _ffi.check(lib.cc_op_attention(qkv.data_ptr(), qkv[:, d:].data_ptr(), qkv[:, 2 * d:].data_ptr(), 3 * d, o.data_ptr(), d,
                                   B, S, H, 64, 0, 0.125, S_()))
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
n = 20
e0.record()
for _ in range(n): run()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / n * 1e3
fl = 4 * S * S * 64 * B * H
print(f"vit attention B={B}: {us:.1f} us  {fl / us / 1e6:.0f} TFLOP/s  {(B * S * 4 * d * 2) / us / 1e3:.0f} GB/s")
q, k, v = [t.view(B, S, H, 64)[:2].transpose(1, 2).float() for t in qkv.split(d, dim=1)]
ref = (torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v).transpose(1, 2).reshape(2 * S, d)
print("rel err", ((o[:2 * S].float() - ref).abs().max() / ref.abs().max()).item())
