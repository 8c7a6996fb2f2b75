"""Feasibility probe (GPU box only): split the B200 into two SM partitions with CUDA green contexts and run the
tensor-bound stages (ViT + mapper) of one batch on the large partition while the latency-bound decode of the previous batch
runs on the small one. Prints stage times sequential / partitioned / overlapped."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from cuda.bindings import driver as drv
import bench
from clipcap_b200 import _ffi
from clipcap_b200.engine import Gpt2Engine, MapperEngine, VitEngine
from oracle import restate as R

lib = _ffi.lib()
dev = torch.device("cuda:0")
torch.zeros(1, device=dev)
B = int(os.environ.get("B", "256"))
DEC = int(os.environ.get("DEC_SMS", "24"))


def chk(r):
    assert r[0] == drv.CUresult.CUDA_SUCCESS, r[0]
    return r[1:] if len(r) > 2 else r[1]


cudev = chk(drv.cuDeviceGet(0))
res = chk(drv.cuDeviceGetDevResource(cudev, drv.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
print("device SMs", res.sm.smCount)
groups, nb, remaining = chk(drv.cuDevSmResourceSplitByCount(1, res, 0, DEC))
print("split:", [g.sm.smCount for g in groups], "remaining", remaining.sm.smCount)
n_dec, n_big = groups[0].sm.smCount, remaining.sm.smCount
desc_b = chk(drv.cuDevResourceGenerateDesc([groups[0]], 1))
desc_a = chk(drv.cuDevResourceGenerateDesc([remaining], 1))
flag = drv.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM
gctx_a = chk(drv.cuGreenCtxCreate(desc_a, cudev, flag))
gctx_b = chk(drv.cuGreenCtxCreate(desc_b, cudev, flag))
nonblock = drv.CUstream_flags.CU_STREAM_NON_BLOCKING
sa = chk(drv.cuGreenCtxStreamCreate(gctx_a, nonblock, 0))
sb = chk(drv.cuGreenCtxStreamCreate(gctx_b, nonblock, 0))
stream_a, stream_b = torch.cuda.ExternalStream(int(sa)), torch.cuda.ExternalStream(int(sb))
print("streams", hex(int(sa)), hex(int(sb)))

state = bench.synthetic_state()
g = R.Gpt2Cfg()
vit = VitEngine(state["vit"], max_batch=B, device=dev)
mapper = MapperEngine(state["mapper"], E=768, d=1024, P=10, K=40, H=8, L=8, max_batch=B, device=dev)
lm = Gpt2Engine(state["lm"], g.d, g.L, g.H, g.V, g.n_pos, max_seqs=B, max_len=60, device=dev)
lib.cc_set_sm_budget(n_dec)
lm_small = Gpt2Engine(state["lm"], g.d, g.L, g.H, g.V, g.n_pos, max_seqs=B, max_len=60, device=dev)
lib.cc_set_sm_budget(0)
px = torch.randn(B, 3, 224, 224, device=dev)


def front(budget):
    lib.cc_set_sm_budget(budget)
    emb = vit.forward(px)
    out = mapper.forward(emb)
    lib.cc_set_sm_budget(0)
    return out


def back(eng, prefix, budget, el=20):
    lib.cc_set_sm_budget(budget)
    out = eng.generate(prefix, "greedy", 1, el, 1.0, 50256)
    lib.cc_set_sm_budget(0)
    return out


def timed(fn, stream, n=5):
    with torch.cuda.stream(stream):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


main = torch.cuda.Stream()
prefix = front(0)
toks_ref, _, _ = back(lm, prefix, 0)
torch.cuda.synchronize()
print(f"sequential, whole device: front {timed(lambda: front(0), main):.2f} ms, generate {timed(lambda: back(lm, prefix, 0), main):.2f} ms, "
      f"prefill only {timed(lambda: back(lm, prefix, 0, 1), main):.2f} ms")
try:
    t_a = timed(lambda: front(n_big), stream_a)
    print(f"partition A ({n_big} SMs) alone: front {t_a:.2f} ms")
    t_b = timed(lambda: back(lm_small, prefix, n_dec), stream_b)
    t_b1 = timed(lambda: back(lm_small, prefix, n_dec, 1), stream_b)
    print(f"partition B ({n_dec} SMs) alone: generate {t_b:.2f} ms (prefill only {t_b1:.2f} ms)")
    with torch.cuda.stream(stream_b):
        toks_b, _, _ = back(lm_small, prefix, n_dec)
    torch.cuda.synchronize()
    print("tokens identical to the whole-device run:", bool(torch.equal(toks_b, toks_ref)))
    # overlapped: front of batch i+1 on A while generate of batch i on B
    n = 6
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ea0, ea1 = torch.cuda.Event(True), torch.cuda.Event(True)
    eb0, eb1 = torch.cuda.Event(True), torch.cuda.Event(True)
    with torch.cuda.stream(stream_a):
        ea0.record()
    with torch.cuda.stream(stream_b):
        eb0.record()
    for i in range(n):
        with torch.cuda.stream(stream_a):
            front(n_big)
        with torch.cuda.stream(stream_b):
            back(lm_small, prefix, n_dec)
    with torch.cuda.stream(stream_a):
        ea1.record()
    with torch.cuda.stream(stream_b):
        eb1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / n
    print(f"overlapped: A {ea0.elapsed_time(ea1) / n:.2f} ms/iter, B {eb0.elapsed_time(eb1) / n:.2f} ms/iter, wall {wall:.2f} ms/iter")
    with torch.cuda.stream(stream_b):
        toks_c, _, _ = back(lm_small, prefix, n_dec)
    torch.cuda.synchronize()
    print("tokens identical after the overlapped run:", bool(torch.equal(toks_c, toks_ref)))
except Exception as e:  # noqa: BLE001
    import traceback
    traceback.print_exc()
    print("probe failed:", e)
