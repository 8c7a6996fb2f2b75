import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clipcap_b200 import _ffi
h = C.CDLL(_ffi.LIB_PATH)
for n in ("cc_last_error", "cc_op_attention"):
    fn = getattr(h, n); fn.restype, fn.argtypes = _ffi.PROTOTYPES[n]
libc = C.CDLL(None)
dev = "cuda"
B, Sq, H, hd = 256, 257, 16, 64
d = H * hd
qkv = torch.randn(B * Sq, 3 * d, device=dev).half(); o = torch.zeros(B * Sq, d, device=dev, dtype=torch.half)
S = lambda: torch.cuda.current_stream().cuda_stream
def run():
    st = h.cc_op_attention(qkv.data_ptr(), qkv.data_ptr() + d * 2, qkv.data_ptr() + 4 * d, 3 * d, o.data_ptr(), d, B, Sq, H, hd, 0, hd ** -0.5, S())
    assert st == 0, h.cc_last_error()
for dbg in [0, 1, 2, 3, 4, 8, 16, 32, 63, 59]:
    libc.setenv(b"CLIPCAP_VA_DBG", str(dbg).encode(), 1)
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); torch.cuda.synchronize(); e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    print(f"dbg={dbg:3d}: {e0.elapsed_time(e1) * 100:.1f} us")
