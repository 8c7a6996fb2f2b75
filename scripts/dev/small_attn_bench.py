"""Developer timing (GPU box): the short-sequence attention at the mapper / prefill shapes of the bench, tcgen05 vs mma.sync
(CLIPCAP_B200_NO_TC_SMALL_ATTN=1 selects the latter)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from clipcap_b200 import _ffi
lib = _ffi.lib()
dev = "cuda"
S_ = lambda: torch.cuda.current_stream().cuda_stream
for (B, S, H, hd, causal) in [(256, 50, 8, 128, 0), (256, 40, 16, 64, 1), (64, 107, 16, 64, 1), (64, 67, 16, 64, 1), (256, 80, 8, 128, 0),
                              (3, 128, 4, 64, 1), (2, 33, 2, 128, 0), (255, 20, 16, 64, 1)]:
    d = H * hd
    qkv = (torch.randn(B * S, 3 * d, device=dev) * 0.5).half()
    o = torch.zeros(B * S, d, device=dev, dtype=torch.half)
    def run():
        _ffi.check(lib.cc_op_attention(qkv.data_ptr(), qkv[:, d:].data_ptr(), qkv[:, 2 * d:].data_ptr(), 3 * d, o.data_ptr(), d,
                                       B, S, H, hd, causal, hd ** -0.5, S_()))
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    q, k, v = [t.view(B, S, H, hd).transpose(1, 2).float() for t in qkv.split(d, dim=1)]
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=bool(causal)).transpose(1, 2).reshape(B * S, d)
    err = ((o.float() - ref).abs().max() / ref.abs().max()).item()
    print(f"B={B} S={S} H={H} hd={hd} causal={causal}: {us:.1f} us  rel err {err:.2e}  ({B * S * 4 * d * 2 / us / 1e3:.0f} GB/s)", flush=True)
