"""Developer smoke check of the kernel-level hooks against torch (GPU box only)."""
import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from clipcap_b200 import _ffi

h = C.CDLL(_ffi.LIB_PATH)
for n in ("cc_last_error", "cc_version", "cc_op_gemm", "cc_op_layernorm", "cc_op_attention", "cc_op_decode_attention"):
    fn = getattr(h, n); fn.restype, fn.argtypes = _ffi.PROTOTYPES[n]
print(h.cc_version().decode(), torch.cuda.get_device_name(0))
dev = "cuda"
def ck(st):
    if st != 0: raise RuntimeError(h.cc_last_error().decode())
S = lambda: torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)
bad = 0
def rel(a, b): return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6)).item()

# ---- GEMM
for (M, N, K) in [(128, 128, 64), (256, 3072, 1024), (300, 1000, 520), (1, 50257, 768), (12800, 3072, 1024), (257, 768, 1024), (65792, 1024, 4096)]:
    a = (torch.randn(M, K, device=dev) * 0.5).half(); w = (torch.randn(N, K, device=dev) * 0.05).half(); bias = torch.randn(N, device=dev)
    ref = a.float() @ w.float().t() + bias
    for bn in (0, 32, 64, 128, 256, 512):
        for epi in (_ffi.EPI_F16_NONE, _ffi.EPI_F32, _ffi.EPI_RESID_F32, _ffi.EPI_F16_GELU_NEW, _ffi.EPI_ARGMAX):
            if N % 8 and epi not in (_ffi.EPI_F32, _ffi.EPI_ARGMAX): continue  # TMA epilogues need 16-byte row strides
            if M * N > 5e7 and (bn not in (0, 256, 512) or epi not in (_ffi.EPI_F16_NONE, _ffi.EPI_RESID_F32)): continue
            if epi == _ffi.EPI_F16_NONE or epi == _ffi.EPI_F16_GELU_NEW:
                out = torch.zeros(M, N, device=dev, dtype=torch.half); r = ref if epi == _ffi.EPI_F16_NONE else torch.nn.functional.gelu(ref, approximate="tanh")
            elif epi == _ffi.EPI_F32:
                out = torch.zeros(M, N, device=dev); r = ref
            elif epi == _ffi.EPI_RESID_F32:
                out = torch.randn(M, N, device=dev); r = ref + out
            else:
                out = torch.zeros(M, device=dev, dtype=torch.int64); r = None
            b_ = None if epi == _ffi.EPI_ARGMAX else bias.data_ptr()
            ck(h.cc_op_gemm(a.data_ptr(), K, w.data_ptr(), b_, out.data_ptr(), N, M, N, K, epi, bn, S()))
            torch.cuda.synchronize()
            if epi == _ffi.EPI_ARGMAX:
                idx = (~out) & 0xFFFFFFFF
                r2 = a.float() @ w.float().t()
                got = r2.gather(1, idx.view(-1, 1)).squeeze(1); best = r2.max(1).values
                e = ((best - got).abs().max() / best.abs().max()).item()
            else:
                e = rel(out, r)
            ok = e < 2e-3
            bad += (not ok)
            print(f"gemm M={M} N={N} K={K} bn={bn} epi={epi} rel={e:.2e} {'ok' if ok else 'FAIL'}")

# ---- LayerNorm
for (rows, d) in [(7, 768), (1000, 1024), (33, 64), (5, 2048)]:
    x = torch.randn(rows, d, device=dev) * 3 + 1; g = torch.randn(d, device=dev); b = torch.randn(d, device=dev)
    y = torch.zeros(rows, d, device=dev, dtype=torch.half)
    ck(h.cc_op_layernorm(x.data_ptr(), d, g.data_ptr(), b.data_ptr(), y.data_ptr(), d, rows, d, 1e-5, S()))
    e = rel(y, torch.nn.functional.layer_norm(x, (d,), g, b, 1e-5)); ok = e < 2e-3; bad += (not ok)
    print(f"layernorm rows={rows} d={d} rel={e:.2e} {'ok' if ok else 'FAIL'}")

# ---- attention
for (B, Sq, H, hd, causal) in [(2, 257, 16, 64, 0), (3, 50, 8, 128, 0), (2, 20, 8, 96, 0), (2, 20, 16, 48, 0), (4, 40, 16, 64, 1), (1, 1, 2, 64, 1), (2, 130, 2, 64, 1)]:
    d = H * hd
    qkv = (torch.randn(B * Sq, 3 * d, device=dev)).half()
    o = torch.zeros(B * Sq, d, device=dev, dtype=torch.half)
    esz = 2
    ck(h.cc_op_attention(qkv.data_ptr(), qkv.data_ptr() + d * esz, qkv.data_ptr() + 2 * d * esz, 3 * d, o.data_ptr(), d, B, Sq, H, hd, causal, hd ** -0.5, S()))
    q, k, v = [t.view(B, Sq, H, hd).transpose(1, 2).float() for t in qkv.split(d, dim=1)]
    r = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=bool(causal)).transpose(1, 2).reshape(B * Sq, d)
    e = rel(o, r); ok = e < 3e-3; bad += (not ok)
    print(f"attention B={B} S={Sq} H={H} hd={hd} causal={causal} rel={e:.2e} {'ok' if ok else 'FAIL'}")

if os.environ.get("ATTN_TIME", "1") == "1":
    B, Sq, H, hd = 256, 257, 16, 64
    d = H * hd
    qkv = torch.randn(B * Sq, 3 * d, device=dev).half(); o = torch.zeros(B * Sq, d, device=dev, dtype=torch.half)
    def run(): ck(h.cc_op_attention(qkv.data_ptr(), qkv.data_ptr() + d * 2, qkv.data_ptr() + 4 * d, 3 * d, o.data_ptr(), d, B, Sq, H, hd, 0, hd ** -0.5, S()))
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); torch.cuda.synchronize(); e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 10
    q, k, v = [t.view(B, Sq, H, hd).transpose(1, 2)[:4].float() for t in qkv.split(d, dim=1)]
    r = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(4 * Sq, d)
    e = rel(o[:4 * Sq], r); ok = e < 3e-3; bad += (not ok)
    print(f"time attention ViT B=256 S=257 H=16: {ms*1000:.1f} us  {4*Sq*Sq*hd*B*H/ms/1e9:.1f} TFLOP/s rel={e:.2e} {'ok' if ok else 'FAIL'}")

# ---- decode attention
nseq, H, t_max = 5, 4, 32
d = H * 64
kc = torch.randn(nseq, H, t_max, 64, device=dev).half(); vc = torch.randn(nseq, H, t_max, 64, device=dev).half()
for pos in (0, 1, 7, 31):
    qkv = torch.randn(nseq, 3 * d, device=dev).half(); o = torch.zeros(nseq, d, device=dev, dtype=torch.half)
    ck(h.cc_op_decode_attention(qkv.data_ptr(), kc.data_ptr(), vc.data_ptr(), None, o.data_ptr(), nseq, H, t_max, pos, 0.125, S()))
    torch.cuda.synchronize()
    q = qkv[:, :d].view(nseq, H, 1, 64).float()
    assert torch.equal(kc[:, :, pos], qkv[:, d:2 * d].view(nseq, H, 64)) and torch.equal(vc[:, :, pos], qkv[:, 2 * d:].view(nseq, H, 64))
    r = torch.nn.functional.scaled_dot_product_attention(q, kc[:, :, :pos + 1].float(), vc[:, :, :pos + 1].float()).reshape(nseq, d)
    e = rel(o, r); ok = e < 3e-3; bad += (not ok)
    print(f"decode_attention pos={pos} rel={e:.2e} {'ok' if ok else 'FAIL'}")

# ---- GEMM timing (device events)
def timeit(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (M, N, K) in [(65792, 3072, 1024), (65792, 1024, 1024), (65792, 4096, 1024), (65792, 1024, 4096), (12800, 3072, 1024), (256, 3072, 1024), (256, 50264, 1024)]:
    a = torch.randn(M, K, device=dev).half(); w = torch.randn(N, K, device=dev).half(); out = torch.zeros(M, N, device=dev, dtype=torch.half)
    for bn in (128, 256, 512):
        ms = timeit(lambda: ck(h.cc_op_gemm(a.data_ptr(), K, w.data_ptr(), None, out.data_ptr(), N, M, N, K, 0, bn, S())))
        print(f"time gemm M={M} N={N} K={K} bn={bn}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s")
    if M > 1000:
        bias = torch.randn(N, device=dev); h32 = torch.zeros(M, N, device=dev)
        for epi, nm, o in ((_ffi.EPI_F16_QUICKGELU, "quickgelu", out), (_ffi.EPI_F16_GELU_NEW, "gelu_new", out), (_ffi.EPI_RESID_F32, "resid_f32", h32)):
            ms = timeit(lambda: ck(h.cc_op_gemm(a.data_ptr(), K, w.data_ptr(), bias.data_ptr(), o.data_ptr(), N, M, N, K, epi, 512, S())))
            print(f"time gemm M={M} N={N} K={K} bn=512 epi={nm}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s")
    ms = timeit(lambda: torch.matmul(a, w.t()))
    print(f"time torch(cuBLAS) M={M} N={N} K={K}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s")
print("FAILURES:", bad)
sys.exit(1 if bad else 0)
