"""Developer timing (GPU box): the greedy decode attention kernel inside an SM partition of DEC_SMS SMs (green context)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from clipcap_b200 import _ffi
from clipcap_b200.engine import SmPartition
lib = _ffi.lib()
dev = torch.device("cuda:0")
L, nseq, H, t_max, d = 24, 256, 16, 64, 1024
kc = torch.randn(L, nseq, H, t_max, 64, device=dev).half(); vc = torch.randn(L, nseq, H, t_max, 64, device=dev).half()
qkv = torch.randn(nseq, 3 * d, device=dev).half(); o = torch.zeros(nseq, d, device=dev, dtype=torch.half)
torch.cuda.synchronize()
import contextlib
class _Whole:
    sms = (148, 148)
    @contextlib.contextmanager
    def on(self, which):
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            yield st
for sms in [int(x) for x in os.environ.get("DEC_SMS", "24,32").split(",")]:
    part = SmPartition(sms, dev) if sms > 0 else _Whole()  # 0 = the whole device
    with part.on(1) as st:
        for pos in (41, 58):
            def f():
                for l in range(L):
                    _ffi.check(lib.cc_op_decode_attention(qkv.data_ptr(), kc[l].data_ptr(), vc[l].data_ptr(), None, o.data_ptr(), nseq, H,
                                                          t_max, pos, 0.125, st.cuda_stream))
            f(); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                f()
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(st)
            for _ in range(5): g.replay()
            e1.record(st); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 5 / L * 1e3
            print(f"{part.sms[1]} SMs, pos={pos}: {us:.1f} us per launch, {nseq * H * (pos + 1) * 256 / us / 1e3:.0f} GB/s")
    del part
