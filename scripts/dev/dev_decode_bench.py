"""Developer timing of the decode-path kernels at the benchmark shape (GPU box only): rotates over 24 layers' worth of
weights / KV cache so every launch streams from HBM like the real step."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from clipcap_b200 import _ffi

h = C.CDLL(_ffi.LIB_PATH)
for n in ("cc_last_error", "cc_op_gemm", "cc_op_decode_attention"):
    fn = getattr(h, n); fn.restype, fn.argtypes = _ffi.PROTOTYPES[n]
dev = "cuda"
S = lambda: torch.cuda.current_stream().cuda_stream
def ck(st):
    if st != 0: raise RuntimeError(h.cc_last_error().decode())
L, nseq, H, t_max, d = 24, 256, 16, 64, 1024
kc = torch.randn(L, nseq, H, t_max, 64, device=dev).half(); vc = torch.randn(L, nseq, H, t_max, 64, device=dev).half()
qkv = torch.randn(nseq, 3 * d, device=dev).half(); o = torch.zeros(nseq, d, device=dev, dtype=torch.half)
def timeit(fn, n=5):
    # capture the launches into a CUDA graph so the host (ctypes + plan encode) is out of the measurement
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
        for _ in range(n): g.replay()
        e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
for pos in (41, 50, 58):
    def f():
        for l in range(L):
            ck(h.cc_op_decode_attention(qkv.data_ptr(), kc[l].data_ptr(), vc[l].data_ptr(), None, o.data_ptr(), nseq, H, t_max, pos, 0.125, S()))
    ms = timeit(f)
    by = nseq * H * (pos + 1) * 64 * 2 * 2
    print(f"decode_attn pos={pos}: {ms / L * 1000:.1f} us/launch  {by / (ms / L) / 1e6:.0f} GB/s")
for (N, K, epi) in [(3072, 1024, 0), (4096, 1024, 3), (1024, 1024, 5), (1024, 4096, 5)]:
    a = torch.randn(nseq, K, device=dev).half(); w = torch.randn(L, N, K, device=dev).half()
    out = torch.zeros(nseq, N, device=dev, dtype=torch.float if epi == 5 else torch.half)
    bias = torch.zeros(N, device=dev)
    def f():
        for l in range(L):
            ck(h.cc_op_gemm(a.data_ptr(), K, w[l].data_ptr(), bias.data_ptr(), out.data_ptr(), N, nseq, N, K, epi, 0, S()))
    ms = timeit(f)
    print(f"gemm M=256 N={N} K={K} epi={epi}: {ms / L * 1000:.1f} us/launch  weights {N * K * 2 / (ms / L) / 1e6:.0f} GB/s")
