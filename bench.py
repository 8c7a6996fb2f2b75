#!/usr/bin/env python
"""bench.py — captions/sec of the ClipCap hot path (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W            # this repo (sm_100a kernels behind the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the box's host cores (oracle port)

A step = one pass of the hot path over one batch of synthetic input:
  pixels [B,3,224,224] -> CLIP ViT-L/14 -> TransformerMapper (8 layers, K=40, P=10, 8 heads) -> GPT-2-medium greedy
  decode of 20 tokens (generate_beam(beam_size=1) semantics) -> token ids                      (BASELINE configs[1], B=256)
`value` is timed with the pixels already resident in HBM; `e2e` is the same calls with pinned HOST pixels copied in and
the token ids read back inside the timed region. N > 1: one process per GPU (torchrun), every rank runs its own B-image
shard (weak scaling) and the prefix embeddings are all-gathered over NCCL.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "captions/sec (224x224 bs=256, 20-tok greedy)"
UNIT = "captions/s"
ENTRY_LENGTH = 20
STOP_TOKEN = 50256
WORKLOAD = dict(workload="configs[1]: ViT-L/14 -> TransformerMapper(L=8,K=40,P=10,H=8) -> GPT-2-medium, 224x224, "
                         "bs=256 per GPU, 20-token greedy decode",
                batch_per_gpu=256, entry_length=ENTRY_LENGTH, lm="gpt2-medium", encoder="ViT-L/14",
                prefix_length=40, projection_length=10, mapper_layers=8, mapper_heads=8)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-captions", type=int, default=12, help="captions timed for the cpu_baseline sample")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ roofline helpers
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def vit_gemm_flops(B):
    """Algorithmic FLOPs of the four GEMMs of one ViT-L/14 block over B images, x24 layers."""
    rows = B * 257
    per_layer = 2 * rows * 1024 * 3072 + 2 * rows * 1024 * 1024 + 2 * 2 * rows * 1024 * 4096
    return 24 * per_layer


def flops_per_caption():
    d, L, V, K, T0 = 1024, 24, 50257, 40, 40
    vit = 2 * 256 * 588 * 1024 + 24 * (2 * 257 * 1024 * 3072 + 4 * 257 * 257 * 1024 + 2 * 257 * 1024 * 1024 +
                                        4 * 257 * 1024 * 4096) + 2 * 1024 * 768
    S = 50
    mapper = 2 * 768 * 10 * d + 8 * (16 * S * d * d + 4 * S * S * d)
    blk = L * (12 * d * d + 13 * d)
    prefill = K * 2 * blk + 2 * V * d + 2 * L * K * K * d
    decode = sum(2 * blk + 2 * V * d + 4 * L * (T0 + s) * d for s in range(1, ENTRY_LENGTH))
    return dict(vit=vit, mapper=mapper, prefill=prefill, decode=decode, total=vit + mapper + prefill + decode)


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU while the timed region runs (NVML)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.05)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join()
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ CPU side
def cpu_reference_setup(state):
    """Oracle-port view of the same weights (fp32, CPU)."""
    from oracle import restate as R
    vit_w = {k: v.detach().float().cpu() for k, v in state["vit"].items()}
    map_w = {k: v.detach().float().cpu() for k, v in state["mapper"].items()}
    lm_w = {k: v.detach().float().cpu() for k, v in state["lm"].items()}
    return R, vit_w, map_w, lm_w, R.VitCfg(), R.MapperCfg(E=768, d=1024, P=10, K=40, H=8, L=8), R.Gpt2Cfg()


def cpu_caption(R, vit_w, map_w, lm_w, vcfg, mcfg, gcfg, pixels_one):
    """One caption exactly as the reference produces it: batch size 1, no KV cache, generate_beam(beam_size=1)."""
    import torch
    with torch.no_grad():
        _, _, out = R.caption_greedy(vit_w, map_w, lm_w, vcfg, mcfg, gcfg, pixels_one, ENTRY_LENGTH, STOP_TOKEN)
    return out[0][0]


def synthetic_state(seed=0):
    """Random-init weights of the named architectures (no checkpoints / network here): the product package's own
    parameter containers under their default initialisers (OpenAI-clip init for the ViT tower, torch defaults +
    prefix_const ~ N(0,1) for the mapper as in the reference, HF GPT-2 init), seeded. The oracle is not involved."""
    import torch
    from clipcap_b200.encoders.clip import ViTImageTower
    from clipcap_b200.encoders.config import EncoderConfig
    from clipcap_b200.model import ClipCapModelPrefixOnly, Config
    torch.manual_seed(seed)
    tower = ViTImageTower()
    cfg = Config(language_model="gpt2-medium", prefix_length=40, projection_length=10, transformer_layers=8,
                 transformer_attention_heads=8, encoder_config=EncoderConfig(encoder_embedding_size=768))
    model = ClipCapModelPrefixOnly(cfg)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    return {"vit": {k: v.detach().clone() for k, v in tower.state_dict().items()},
            "mapper": {k[len("transformer_mapper."):]: v for k, v in sd.items() if k.startswith("transformer_mapper.")},
            "lm": {k[len("language_model."):]: v for k, v in sd.items() if k.startswith("language_model.")}}


def synthetic_pixels(B, seed):
    """CLIP-normalised images are roughly unit-scale noise for throughput purposes: seeded N(0,1), fp32 [B,3,224,224]."""
    import torch
    return torch.randn(B, 3, 224, 224, generator=torch.Generator().manual_seed(seed))


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the Python reference itself cannot travel to the
    GPU box) on all host cores. One step = one caption (the reference is batch-size-1)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    state = synthetic_state()
    R, vit_w, map_w, lm_w, vcfg, mcfg, gcfg = cpu_reference_setup(state)
    px = synthetic_pixels(max(1, min(4, args.steps + args.warmup)), 1234)
    for i in range(args.warmup):
        cpu_caption(R, vit_w, map_w, lm_w, vcfg, mcfg, gcfg, px[i % px.shape[0]:i % px.shape[0] + 1])
    times = []
    t0 = time.perf_counter()
    for i in range(args.steps):
        t1 = time.perf_counter()
        cpu_caption(R, vit_w, map_w, lm_w, vcfg, mcfg, gcfg, px[i % px.shape[0]:i % px.shape[0] + 1])
        times.append(time.perf_counter() - t1)
    total = time.perf_counter() - t0
    value = args.steps / total
    sample = f"{args.steps} captions, one image per step (reference is batch-size-1), fp32, no KV cache"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "p50_ms": 1e3 * statistics.median(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(WORKLOAD, note="CPU path: batch of 1 per step"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------ B200 side
def run_b200(args):
    import torch
    import torch.distributed as dist

    # stdout must carry exactly ONE JSON line. Native libraries write there too (NCCL prints its version banner on
    # communicator creation), so file descriptor 1 points at stderr while the run lasts and the result line is written to
    # the saved descriptor at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from clipcap_b200 import _ffi
    from clipcap_b200.encoders.clip import CLIPModel, ViTImageTower
    from clipcap_b200.encoders.config import EncoderConfig
    from clipcap_b200.distributed import caption_step
    from clipcap_b200.model import ClipCapModelPrefixOnly, Config
    from clipcap_b200.pipeline import CaptionPipeline
    _ffi.lib()

    B = args.batch
    state = synthetic_state()
    tower = ViTImageTower()
    tower.load_state_dict(state["vit"], strict=True)
    encode_fn = CLIPModel(tower).eval().to(dev)
    cfg = Config(language_model="gpt2-medium", prefix_length=40, projection_length=10, transformer_layers=8,
                 transformer_attention_heads=8, encoder_config=EncoderConfig(encoder_embedding_size=768))
    model = ClipCapModelPrefixOnly(cfg)
    sd = {f"transformer_mapper.{k}": v for k, v in state["mapper"].items()}
    sd.update({f"language_model.{k}": v for k, v in state["lm"].items()})
    model.load_state_dict(sd, strict=True)
    model = model.eval().to(dev)

    px_host = synthetic_pixels(B, 1234 + rank).pin_memory()  # fp32, as the reference's preprocess produces
    px_dev = px_host.to(dev, non_blocking=True)
    prefix_all = torch.empty(world * B, 40, 1024, device=dev, dtype=torch.float32) if world > 1 else None
    tok_host = torch.empty(B, ENTRY_LENGTH, dtype=torch.int32).pin_memory()
    len_host = torch.empty(B, dtype=torch.int32).pin_memory()

    def step(pixels):
        # cc_vit_forward -> cc_mapper_forward -> [prefix all-gather over NVLink, SURVEY §8e] -> cc_generate
        toks, lens, _ = caption_step(encode_fn, model, pixels, ENTRY_LENGTH, STOP_TOKEN, prefix_all)
        return toks, lens, None

    # End to end through the public serving loop (clipcap_b200.pipeline.CaptionPipeline): every step copies its pinned
    # host pixels to the device and its token ids back; the copy of step i+1 overlaps the compute of step i.
    pipe = CaptionPipeline(encode_fn, model, B, 224, ENTRY_LENGTH, STOP_TOKEN, dev, prefix_all=prefix_all)

    def run_e2e(steps):
        marks = []
        for toks_h, lens_h in pipe.run(px_host for _ in range(steps)):
            marks.append(time.perf_counter())
            tok_host.copy_(toks_h)  # the caller consumes the ids
            len_host.copy_(lens_h)
        return marks

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        barrier()
        evs[0].record()
        for i in range(steps):
            fn()
            evs[i + 1].record()
        barrier()
        per = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
        total = evs[0].elapsed_time(evs[steps])
        if world > 1:
            t = torch.tensor([total], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total = t.item()
        return total, per

    for _ in range(max(3, args.warmup)):
        step(px_dev)
    barrier()
    vit_eng, map_eng, lm_eng = tower._engine, model.transformer_mapper._engine, model.language_model._engine
    launches_per_step = vit_eng.last_launches + map_eng.last_launches + lm_eng.last_launches + 1 + 3  # + embed, 3 copies

    sampler = ClockSampler(local_rank)
    sampler.start()
    total_ms, per = timed(lambda: step(px_dev), args.steps)
    clocks = sampler.stop()
    run_e2e(2)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    t_start = time.perf_counter()
    marks = run_e2e(args.steps)
    ev1.record()
    barrier()
    e2e_ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = t.item()
    e2e_per = [1e3 * (b - a) for a, b in zip([t_start] + marks[:-1], marks)]

    # ---- roofline of the dominant kernel (tcgen05 GEMM, ViT block shapes), CUDA events around each launch
    peaks = measured_peaks()
    roof = None
    if rank == 0:
        lib = _ffi.lib()
        if hasattr(lib, "cc_prof_enable"):
            import ctypes as C
            lib.cc_prof_enable.argtypes = [C.c_int]
            lib.cc_prof_read.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
            lib.cc_prof_enable(1)
            for _ in range(2):
                encode_fn(px_dev)
            torch.cuda.synchronize()
            ms, fl, n = C.c_double(), C.c_double(), C.c_longlong()
            lib.cc_prof_read(C.byref(ms), C.byref(fl), C.byref(n))
            lib.cc_prof_enable(0)
            if n.value > 0 and ms.value > 0:
                ach = fl.value / (ms.value * 1e-3) / 1e12
                traffic = None
                tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
                if os.path.exists(tp):
                    with open(tp) as f:
                        traffic = json.load(f).get("traffic_bytes_per_launch")
                roof = {"bound": "tensor", "achieved": ach, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                        "frac": ach / peaks["tf_sustained"], "traffic": traffic,
                        "kernel": "gemm_tn_kernel<256,*,2> (tcgen05 cta_group::2, 256x256 CTA-pair tile) over the ViT-L/14 block GEMMs",
                        "launches_timed": n.value, "avg_launch_ms": ms.value / n.value,
                        "flops_per_launch": fl.value / n.value, "peak_source": peaks["source"] + ", sustained bf16"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    captions = B * world
    value = captions * args.steps / (total_ms * 1e-3)
    fpc = flops_per_caption()
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": total_ms / args.steps, "p50_ms": statistics.median(per), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": dict(WORKLOAD, global_batch=captions, pixels_dtype="f32",
                       l2="per-step working set (154 MB pixels, >1 GB activations, 1.4 GB weights) exceeds the 126 MB L2",
                       weights="seeded random init (no checkpoints offline)"),
        "clocks": clocks,
        "e2e": {"value": captions * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                "p50_ms": statistics.median(e2e_per), "h2d_bytes_per_step": pipe.h2d_bytes_per_batch,
                "d2h_bytes_per_step": pipe.d2h_bytes_per_batch,
                "api": "clipcap_b200.pipeline.CaptionPipeline (pinned host pixels in, token ids out, double-buffered H2D)"},
        "gpu_launches": launches_per_step * args.steps,
        "launches_per_step": launches_per_step,
        "model_tflops": value * fpc["total"] / 1e12,
        "roofline": roof,
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        R, vit_w, map_w, lm_w, vcfg, mcfg, gcfg = cpu_reference_setup(state)
        n = args.cpu_captions
        cpu_caption(R, vit_w, map_w, lm_w, vcfg, mcfg, gcfg, px_host[:1])  # warm-up
        t0 = time.perf_counter()
        cpu_tokens = [cpu_caption(R, vit_w, map_w, lm_w, vcfg, mcfg, gcfg, px_host[i:i + 1]) for i in range(n)]
        dt = time.perf_counter() - t0
        gpu_tokens = tok_host[:n].tolist()
        agree = sum(int(gpu_tokens[i][:len(cpu_tokens[i])] == cpu_tokens[i]) for i in range(n))
        out["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"first {n} images of the same batch, one at a time (reference is batch-size-1), "
                                         f"fp32, no KV cache; {agree}/{n} captions token-identical to the GPU run"}
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
