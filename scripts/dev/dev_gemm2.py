import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from clipcap_b200 import _ffi
h = C.CDLL(_ffi.LIB_PATH)
for n in ("cc_last_error", "cc_op_gemm"):
    fn = getattr(h, n); fn.restype, fn.argtypes = _ffi.PROTOTYPES[n]
dev = "cuda"; S = lambda: torch.cuda.current_stream().cuda_stream
def ck(st):
    if st != 0: raise RuntimeError(h.cc_last_error().decode())
def timeit(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (M, N, K) in [(65792, 4096, 1024), (65792, 1024, 4096), (65792, 3072, 1024)]:
    a = torch.randn(M, K, device=dev).half(); w = torch.randn(N, K, device=dev).half(); out = torch.zeros(M, N, device=dev, dtype=torch.half)
    for bn in (256, 512):
        ms = timeit(lambda: ck(h.cc_op_gemm(a.data_ptr(), K, w.data_ptr(), None, out.data_ptr(), N, M, N, K, 0, bn, S())))
        print(f"time gemm M={M} N={N} K={K} bn={bn}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s")
