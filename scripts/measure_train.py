"""Training-step throughput at the reference's training shape (GPT-2-medium, TransformerMapper L=8 K=40 P=10 H=8, batch 64,
67 caption tokens): cc_train_step forward + backward + AdamW, CUDA events; optionally the reference algorithm (oracle
port, torch autograd on the host cores) on a small batch beside it.  GPU box:  python scripts/measure_train.py [--cpu 2]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from clipcap_b200.engine import TrainEngine, adamw_update
from oracle import restate as R
from oracle import synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--tokens", type=int, default=67)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--cpu", type=int, default=0, help="samples for the CPU (oracle port) training step; 0 = skip")
    ap.add_argument("--out", default=None)
    ap.add_argument("--profile", action="store_true", help="one step between cudaProfilerStart/Stop (ncu --profile-from-start off)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    state = bench.synthetic_state()
    mcfg, gcfg = R.MapperCfg(E=768, d=1024, P=10, K=40, H=8, L=8), R.Gpt2Cfg()
    B, Tt = a.batch, a.tokens
    eng = TrainEngine(state["lm"], E=768, d=1024, P=10, K=40, H=8, L=8, lm_layers=24, lm_heads=16, max_batch=B,
                      max_tokens=Tt, device=dev)
    params = {k: v.to(dev).contiguous() for k, v in state["mapper"].items()}
    grads = {k: torch.empty_like(v) for k, v in params.items()}
    m = {k: torch.zeros_like(v) for k, v in params.items()}
    v2 = {k: torch.zeros_like(v) for k, v in params.items()}
    g = torch.Generator().manual_seed(1)
    tokens = torch.randint(1, gcfg.V, (B, Tt), generator=g)
    lens = torch.randint(Tt // 3, Tt + 1, (B,), generator=g)
    for b in range(B):
        tokens[b, int(lens[b]):] = -1
    emb = synth.embeddings(B, 768, seed=2)
    tok_d, emb_d = tokens.to(dev), emb.to(dev)
    step_no = [0]

    def step():
        loss = eng.step(params, emb_d, tok_d, grads)
        step_no[0] += 1
        for k in params:
            adamw_update(params[k], grads[k], m[k], v2[k], 2e-5, 0.9, 0.999, 1e-8, 0.0, step_no[0])
        return loss

    def fwd_only():
        return eng.step(params, emb_d, tok_d, None)

    def timed(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, float(out)

    if a.profile:
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        step()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return
    ms, loss = timed(step, a.iters)
    ms_f, _ = timed(fwd_only, a.iters)
    d, L = 1024, 24
    rows_l, rows_m, rows_sel = B * (40 + Tt), B * 50, B * Tt
    lm_gemm = 2 * rows_l * L * 12 * d * d
    head = 2 * rows_sel * gcfg.V * d
    map_gemm = 2 * rows_m * 8 * 8 * d * d + 2 * B * 768 * 10 * d
    flops = 2 * (lm_gemm + head) + 3 * map_gemm  # LM: forward + dgrad; mapper: forward + dgrad + wgrad
    line = {"workload": f"ClipCapModelPrefixOnly training step: GPT-2-medium frozen, mapper L=8 K=40 P=10 H=8, batch {B}, {Tt} caption tokens",
            "ms_per_step": ms, "samples_per_s": B / (ms * 1e-3), "forward_loss_only_ms": ms_f, "loss": loss,
            "gemm_tflops": flops / (ms * 1e-3) / 1e12, "launches": eng.last_launches,
            "mem_gb": torch.cuda.max_memory_allocated() / 1e9}
    if a.cpu > 0:
        torch.set_num_threads(os.cpu_count() or 1)
        n = a.cpu
        t0 = time.perf_counter()
        cpu_loss, _ = R.training_loss_and_grads(state["mapper"], state["lm"], mcfg, gcfg, tokens[:n], emb[:n])
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"samples_per_s": n / dt, "cores": os.cpu_count(), "kind": "port",
                                "sample": f"{n} samples of the same batch, torch autograd fp32 (reference algorithm)",
                                "loss": cpu_loss}
    print(json.dumps(line))
    if a.out:
        with open(a.out, "w") as f:
            f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
