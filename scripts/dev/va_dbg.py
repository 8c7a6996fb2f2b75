import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from clipcap_b200 import _ffi
h = C.CDLL(_ffi.LIB_PATH)
for n in ("cc_last_error", "cc_op_attention"):
    fn = getattr(h, n); fn.restype, fn.argtypes = _ffi.PROTOTYPES[n]
dev = "cuda"
B, Sq, H, hd = 256, 257, 16, 64
d = H * hd
qkv = torch.randn(B * Sq, 3 * d, device=dev).half(); o = torch.zeros(B * Sq, d, device=dev, dtype=torch.half)
S = lambda: torch.cuda.current_stream().cuda_stream
def run():
    st = h.cc_op_attention(qkv.data_ptr(), qkv.data_ptr() + d * 2, qkv.data_ptr() + 4 * d, 3 * d, o.data_ptr(), d, B, Sq, H, hd, 0, hd ** -0.5, S())
    assert st == 0, h.cc_last_error()
run(); run()
