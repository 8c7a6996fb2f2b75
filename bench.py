#!/usr/bin/env python
"""bench.py — captions/sec of the ClipCap hot path (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W            # this repo (sm_100a kernels behind the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the box's host cores (oracle port)
  python bench.py --workload vit_only|beam5 ...            # BASELINE configs[2] / configs[3] per-GPU shards (extra lines)

A step = one pass of the hot path over one batch of synthetic input:
  pixels [B,3,224,224] -> CLIP ViT-L/14 -> TransformerMapper (8 layers, K=40, P=10, 8 heads) -> GPT-2-medium greedy
  decode of 20 tokens (generate_beam(beam_size=1) semantics) -> token ids                      (BASELINE configs[1], B=256)
Both numbers go through the public serving loop, clipcap_b200.pipeline.CaptionPipeline: `value` with the pixels already
resident in HBM, `e2e` with pinned HOST pixels copied in and the token ids read back inside the timed region. The loop
splits the GPU into two SM partitions (CUDA green contexts): image tower + mapper + prefill of batch i+1 on the large one,
the decode steps of batch i on the small one (--partition-sms 0 turns that off; results are bit-identical either way).
The timed region covers K whole steps including the drain of the last batch's decode.
N > 1: one process per GPU (torchrun), every rank runs its own B-image shard (weak scaling) and the prefix embeddings are
all-gathered over NCCL. Prints ONE JSON line on rank 0.
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

ENTRY_LENGTH = 20
STOP_TOKEN = 50256
DEFAULT_PARTITION_SMS = 32
MODEL_KEYS = dict(entry_length=ENTRY_LENGTH, lm="gpt2-medium", encoder="ViT-L/14", prefix_length=40, projection_length=10,
                  mapper_layers=8, mapper_heads=8)
WORKLOADS = {
    "caption": dict(
        metric="captions/sec (224x224 bs=256, 20-tok greedy)", unit="captions/s", batch=256, mode="greedy", beam=1,
        partition_sms=DEFAULT_PARTITION_SMS, prefill_defer=3,
        workload="configs[1]: ViT-L/14 -> TransformerMapper(L=8,K=40,P=10,H=8) -> GPT-2-medium, 224x224, bs=256 per GPU, "
                 "20-token greedy decode"),
    "vit_only": dict(
        metric="images/sec (ViT-L/14 encode-only, bs=1024 over 8 GPUs)", unit="images/s", batch=128, mode=None, beam=1,
        partition_sms=0,
        workload="configs[2]: ViT-L/14 encode-only, 224x224, bs=1024 synthetic images sharded over 8 GPUs = 128 per GPU"),
    "beam5": dict(
        metric="captions/sec (224x224 bs=2048 over 8 GPUs, beam=5, 20 tokens)", unit="captions/s", batch=256, mode="beam",
        beam=5, partition_sms=0,
        workload="configs[3]: ViT-L/14 -> TransformerMapper(L=8,K=40,P=10,H=8) -> GPT-2-medium, bs=2048 over 8 GPUs = 256 "
                 "per GPU, beam=5, 20 tokens, NCCL prefix all-gather"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="caption", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="images per GPU (default: the workload's)")
    ap.add_argument("--partition-sms", type=int, default=int(os.environ.get("CLIPCAP_B200_PARTITION_SMS", "-1")),
                    help="SMs of the decode partition (0 = one stream, no SM partitioning; default: the workload's — 32 for "
                         "the greedy caption step, 0 for beam-5, whose 1280-row decode is throughput-bound like the rest)")
    ap.add_argument("--prefill-defer", type=int, default=int(os.environ.get("CLIPCAP_B200_PREFILL_DEFER", "-1")),
                    help="with SM partitions: trailing GPT-2 prefill blocks that run on the decode partition (balance "
                         "between the two partitions; default: the workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gather", action="store_true", help="N > 1: skip the prefix all-gather (attribution runs)")
    ap.add_argument("--cpu-captions", type=int, default=12, help="captions timed for the cpu_baseline sample")
    return ap.parse_args()


def workload_config(name, batch, world):
    w = WORKLOADS[name]
    return dict(MODEL_KEYS, workload=w["workload"], batch_per_gpu=batch, global_batch=batch * world)


# ------------------------------------------------------------------------------------------------ roofline helpers
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def algorithmic_work(beam=1, Tp=40, EL=ENTRY_LENGTH):
    """Per-image algorithmic work of each stage: FLOPs, and for the decode steps the HBM bytes
    (weights once per step for the whole batch are returned separately)."""
    d, L, V, K = 1024, 24, 50257, Tp
    vit = 2 * 256 * 588 * 1024 + 24 * (2 * 257 * 1024 * 3072 + 4 * 257 * 257 * 1024 + 2 * 257 * 1024 * 1024 +
                                        4 * 257 * 1024 * 4096) + 2 * 1024 * 768
    S = 50
    mapper = 2 * 768 * 10 * d + 8 * (16 * S * d * d + 4 * S * S * d)
    blk = L * (12 * d * d + 13 * d)
    prefill = K * 2 * blk + 2 * V * d + 2 * L * K * K * d
    decode = beam * sum(2 * blk + 2 * V * d + 4 * L * (Tp + s) * d for s in range(1, EL))
    kv_pos = L * 2 * d * 2  # bytes of K,V per sequence position (fp16)
    # unique cache bytes read per image over the steps: the prefix once per image, generated positions once per beam
    decode_kv_bytes = sum((Tp + beam * s) * kv_pos for s in range(1, EL))
    weight_bytes_per_step = 2 * (blk + V * d)  # fp16 blocks + tied head, shared by the batch
    return dict(vit=vit, mapper=mapper, prefill=prefill, decode=decode, decode_kv_bytes=decode_kv_bytes,
                decode_weight_bytes_per_step=weight_bytes_per_step, total=vit + mapper + prefill + decode)


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


def cpu_caption(R, vit_w, map_w, lm_w, vcfg, mcfg, gcfg, pixels_one, beam=1):
    """One caption exactly as the reference produces it: batch size 1, no KV cache, generate_beam(beam_size=beam)."""
    import torch
    with torch.no_grad():
        if beam == 1:
            _, _, out = R.caption_greedy(vit_w, map_w, lm_w, vcfg, mcfg, gcfg, pixels_one, ENTRY_LENGTH, STOP_TOKEN)
            return out[0][0]
        emb = R.vit_encode(vit_w, pixels_one, vcfg)
        prefix = R.mapper_forward(map_w, emb, mcfg)
        return R.generate_beam(lm_w, gcfg, prefix[0:1], beam, ENTRY_LENGTH, 1.0, STOP_TOKEN)[0]


def cpu_encode(R, vit_w, vcfg, pixels):
    import torch
    with torch.no_grad():
        return R.vit_encode(vit_w, pixels, vcfg)


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
    GPU box) on all host cores. One step = one unit (the reference is batch-size-1): a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    w = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    state = synthetic_state()
    R, vit_w, map_w, lm_w, vcfg, mcfg, gcfg = cpu_reference_setup(state)
    px = synthetic_pixels(max(1, min(4, args.steps + args.warmup)), 1234)

    def one(i):
        p = px[i % px.shape[0]:i % px.shape[0] + 1]
        if w["mode"] is None:
            cpu_encode(R, vit_w, vcfg, p)
        else:
            cpu_caption(R, vit_w, map_w, lm_w, vcfg, mcfg, gcfg, p, w["beam"])

    for i in range(args.warmup):
        one(i)
    times = []
    t0 = time.perf_counter()
    for i in range(args.steps):
        t1 = time.perf_counter()
        one(i)
        times.append(time.perf_counter() - t1)
    total = time.perf_counter() - t0
    value = args.steps / total
    batch = args.batch or w["batch"]
    sample = (f"{args.steps} units, one image per step (the reference is batch-size-1), fp32, no KV cache, "
              f"{cores} host threads")
    print(json.dumps({
        "impl": "reference", "metric": w["metric"], "value": value, "unit": w["unit"], "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "p50_ms": 1e3 * statistics.median(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args.workload, batch, max(1, args.gpus)),
        "reference_note": "CPU path: one image per step (bounded sample of the same workload)",
        "cpu_baseline": {"value": value, "unit": w["unit"], "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": w["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------ B200 side
def stage_report(trace, B, beam, peaks, world_rank_note="", defer_frac=0.0):
    """Mean live stage times (CUDA events recorded inside the timed steps) against each stage's roofline. `defer_frac`: the
    share of the GPT-2 prefill blocks that the serving loop runs at the head of the decode interval (prefill_defer / L): that
    share of the prefill FLOPs is charged to the decode stage's roofline time instead of the prefill stage's."""
    work = algorithmic_work(beam=beam)

    def mean_ms(a, b):
        vals = [r[a].elapsed_time(r[b]) for r in trace if a in r and b in r]
        return sum(vals) / len(vals) if vals else None

    t = {"vit": mean_ms("front0", "vit"), "mapper": mean_ms("vit", "mapper"), "prefill": mean_ms("mapper", "prefill"),
         "decode": mean_ms("dec0", "dec1")}
    stages, roof_ms = {}, 0.0
    moved_ms = 0.0
    for name in ("vit", "mapper", "prefill"):
        fl = work[name] * B
        if name == "prefill" and defer_frac > 0:
            moved_ms = fl * defer_frac / (peaks["tf_sustained"] * 1e12) * 1e3
            fl *= 1.0 - defer_frac
        ideal = fl / (peaks["tf_sustained"] * 1e12) * 1e3
        roof_ms += ideal
        if t[name]:
            ach = fl / (t[name] * 1e-3) / 1e12
            stages[name] = {"bound": "tensor", "flops": fl, "ms": t[name], "achieved": ach, "unit": "TFLOP/s",
                            "peak": peaks["tf_sustained"], "frac": ach / peaks["tf_sustained"], "roofline_ms": ideal}
    by = (ENTRY_LENGTH - 1) * work["decode_weight_bytes_per_step"] + B * work["decode_kv_bytes"]
    fl = work["decode"] * B
    ideal_hbm = by / (peaks["hbm"] * 1e9) * 1e3
    ideal_tc = fl / (peaks["tf_sustained"] * 1e12) * 1e3
    ideal = max(ideal_hbm, ideal_tc) + moved_ms
    roof_ms += ideal
    if t["decode"]:
        gbs = by / (t["decode"] * 1e-3) / 1e9
        stages["decode"] = {"bound": "hbm" if ideal_hbm >= ideal_tc else "tensor", "bytes": by, "flops": fl,
                            "ms": t["decode"], "achieved": gbs, "unit": "GB/s", "peak": peaks["hbm"],
                            "frac": ideal / t["decode"], "roofline_ms": ideal,
                            "ms_per_decode_step": t["decode"] / (ENTRY_LENGTH - 1)}
    front = mean_ms("front0", "prefill")
    lat = sorted(r["front0"].elapsed_time(r["dec1"]) for r in trace if "front0" in r and "dec1" in r)
    latency = lat[len(lat) // 2] if lat else None  # one batch through the loop: first kernel of its image tower -> last decode step
    return stages, roof_ms, front, latency


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
    from clipcap_b200.model import ClipCapModelPrefixOnly, Config
    from clipcap_b200.pipeline import CaptionPipeline
    lib = _ffi.lib()

    w = WORKLOADS[args.workload]
    B = args.batch or w["batch"]
    beam = w["beam"]
    vit_only = w["mode"] is None
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
    gather = world > 1 and not args.no_gather and not vit_only
    comm = None
    if gather:  # the C-ABI collective: cc_comm_create + in-place cc_allgather_prefix (NCCL over NVLink)
        from clipcap_b200.distributed import PrefixComm
        comm = PrefixComm.from_process_group(dev)
    tok_host = torch.empty(B, ENTRY_LENGTH, dtype=torch.int32).pin_memory()
    len_host = torch.empty(B, dtype=torch.int32).pin_memory()
    emb_host = torch.empty(B, 768, dtype=torch.float32).pin_memory()
    partition_sms = 0 if vit_only else (w["partition_sms"] if args.partition_sms < 0 else args.partition_sms)
    pipe, partition_note = None, None
    prefill_defer = w.get("prefill_defer", 0) if args.prefill_defer < 0 else args.prefill_defer
    if not vit_only:
        try:
            pipe = CaptionPipeline(encode_fn, model, B, 224, ENTRY_LENGTH, STOP_TOKEN, dev, comm=comm,
                                   prefix_dtype=torch.float16, partition_sms=partition_sms, mode=w["mode"], beam=beam,
                                   prefill_defer=prefill_defer)
        except Exception as e:  # noqa: BLE001 — no green contexts on this driver: same kernels on one stream
            partition_note = f"SM partitioning unavailable ({type(e).__name__}: {e}); single stream"
            partition_sms = 0
            pipe = CaptionPipeline(encode_fn, model, B, 224, ENTRY_LENGTH, STOP_TOKEN, dev, comm=comm,
                                   prefix_dtype=torch.float16, partition_sms=0, mode=w["mode"], beam=beam)

    # ---- the two step loops. Each returns the host time stamps at which a batch's result was handed out.
    if vit_only:
        px_stage = [torch.empty_like(px_dev) for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)

        def run_value(steps):
            marks = []
            for _ in range(steps):
                encode_fn(px_dev)
                marks.append(time.perf_counter())
            return marks

        def run_e2e(steps):
            # pinned host pixels -> device (copy stream, double-buffered) -> encode -> embeddings back to the host
            marks, compute = [], torch.cuda.current_stream(dev)
            copied = [torch.cuda.Event() for _ in range(2)]
            used = [torch.cuda.Event() for _ in range(2)]
            with torch.cuda.stream(copy_stream):
                px_stage[0].copy_(px_host, non_blocking=True)
                copied[0].record(copy_stream)
            for i in range(steps):
                s = i & 1
                if i + 1 < steps:
                    with torch.cuda.stream(copy_stream):
                        if i >= 1:
                            copy_stream.wait_event(used[s ^ 1])
                        px_stage[s ^ 1].copy_(px_host, non_blocking=True)
                        copied[s ^ 1].record(copy_stream)
                compute.wait_event(copied[s])
                emb = encode_fn(px_stage[s])
                used[s].record(compute)
                emb_host.copy_(emb, non_blocking=True)
                marks.append(time.perf_counter())
            return marks
        h2d, d2h = px_host.numel() * 4, emb_host.numel() * 4
        api = "clipcap_b200.encoders.CLIPModel.forward (pinned host pixels in, embeddings out, double-buffered H2D)"
    else:
        def run_value(steps):
            marks = []
            for toks_h, lens_h in pipe.run((px_dev for _ in range(steps)), resident=True):
                marks.append(time.perf_counter())
            return marks

        def run_e2e(steps):
            marks = []
            for toks_h, lens_h in pipe.run(px_host for _ in range(steps)):
                marks.append(time.perf_counter())
                tok_host.copy_(toks_h)  # the caller consumes the ids
                len_host.copy_(lens_h)
            return marks
        h2d, d2h = pipe.h2d_bytes_per_batch, pipe.d2h_bytes_per_batch
        api = ("clipcap_b200.pipeline.CaptionPipeline (pinned host pixels in, token ids out, double-buffered H2D"
               + (f", SM partitions {pipe.partition.sms[0]} + {pipe.partition.sms[1]}" if pipe.partition else "") + ")")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps between barrier + synchronize on both sides, CUDA events around the whole region (the region ends
        after a device synchronize, so every stream — both SM partitions, the copy stream — is inside it)."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        t0 = time.perf_counter()
        marks = fn(steps)
        torch.cuda.synchronize()
        ev1.record()
        barrier()
        mine = ev0.elapsed_time(ev1)
        per = [1e3 * (b - a) for a, b in zip([t0] + marks[:-1], marks)]
        return mine, per

    run_value(max(3, args.warmup))
    barrier()
    launches_per_step = tower._engine.last_launches
    if not vit_only:
        launches_per_step += model.transformer_mapper._engine.last_launches + pipe._engines()[0].last_launches + 1 + 3
    if pipe is not None:
        pipe.trace = []
    sampler = ClockSampler(local_rank)
    sampler.start()
    my_ms, per = timed(run_value, args.steps)
    clocks = sampler.stop()
    trace = pipe.trace if pipe is not None else []
    if pipe is not None:
        pipe.trace = None
    run_e2e(2)
    my_e2e_ms, e2e_per = timed(run_e2e, args.steps)

    total_ms, e2e_ms, per_rank = my_ms, my_e2e_ms, None
    if world > 1:
        t = torch.tensor([my_ms, my_e2e_ms, float(clocks["sm_mhz"] or 0)], device=dev)
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        total_ms = max(x[0].item() for x in allt)   # the step of the job is the slowest rank's
        e2e_ms = max(x[1].item() for x in allt)
        per_rank = [{"rank": r, "ms_per_step": allt[r][0].item() / args.steps,
                     "e2e_ms_per_step": allt[r][1].item() / args.steps, "sm_mhz": allt[r][2].item()}
                    for r in range(world)]

    # ---- multi-GPU exactness: every rank decodes rank 0's batch once more; the ids must equal rank 0's bit for bit
    multi_gpu_exact = None
    if world > 1 and not vit_only:
        px0 = synthetic_pixels(B, 1234).to(dev)
        toks0 = [t.clone() for t, _ in pipe.run([px0], resident=True)][0].to(dev)
        ref = toks0.clone()
        dist.broadcast(ref, src=0)
        ok = torch.tensor([1 if torch.equal(ref, toks0) else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        multi_gpu_exact = bool(ok.item())

    # ---- roofline of the dominant kernel (tcgen05 GEMM, ViT block shapes), CUDA events around each launch
    peaks = measured_peaks()
    roof = None
    if rank == 0:
        import ctypes as C
        lib.cc_prof_enable(1)
        for _ in range(2):
            emb_p = encode_fn(px_dev)
            if not vit_only:
                model.transformer_mapper(emb_p)
        torch.cuda.synchronize()
        ms, fl, n = C.c_double(), C.c_double(), C.c_longlong()
        lib.cc_prof_read(C.byref(ms), C.byref(fl), C.byref(n))
        families = {}
        for bn, label in ((512, "pair_256x256"), (256, "128x256"), (128, "128x128"), (64, "128x64"), (32, "128x32")):
            fm, ff, fn = C.c_double(), C.c_double(), C.c_longlong()
            lib.cc_prof_read_family(bn, C.byref(fm), C.byref(ff), C.byref(fn))
            if fn.value > 0 and fm.value > 0:
                families[label] = {"launches": fn.value, "ms": fm.value, "tflops": ff.value / (fm.value * 1e-3) / 1e12}
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
                    "flops_per_launch": fl.value / n.value, "peak_source": peaks["source"] + ", sustained bf16",
                    "timed": "two image-tower (+ mapper) passes on the whole device right after the timed steps (same process)",
                    "gemm_families": families}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    units = B * world
    ms_per_step = total_ms / args.steps
    value = units * args.steps / (total_ms * 1e-3)
    work = algorithmic_work(beam=beam)
    if roof is not None:
        if vit_only:
            ideal = work["vit"] * B / (peaks["tf_sustained"] * 1e12) * 1e3
            roof["stages"] = {"vit": {"bound": "tensor", "flops": work["vit"] * B, "ms": ms_per_step,
                                      "achieved": work["vit"] * B / (ms_per_step * 1e-3) / 1e12, "unit": "TFLOP/s",
                                      "peak": peaks["tf_sustained"], "roofline_ms": ideal,
                                      "frac": ideal / ms_per_step}}
            roof["step_roofline_ms"], roof["step_frac"] = ideal, ideal / ms_per_step
        else:
            defer = pipe.prefill_defer if pipe is not None and pipe.partition is not None else 0
            stages, roof_ms, front_ms, latency_ms = stage_report(trace, B, beam, peaks, defer_frac=defer / 24.0)
            roof["stages"] = stages
            roof["step_roofline_ms"] = roof_ms           # sum of the stages' roofline times
            roof["step_frac"] = roof_ms / ms_per_step     # whole step against its roofline
            roof["front_ms"] = front_ms                   # image tower + mapper + prefill of one batch (large partition)
            roof["batch_latency_p50_ms"] = latency_ms     # one batch, image tower start -> last decode step (CUDA events)
            roof["stages_note"] = ("stage ms = mean of CUDA-event intervals recorded inside the timed steps on the "
                                   "stage's own stream; with SM partitions the decode stage of batch i overlaps the "
                                   "other stages of batch i+1, so the stage times add up to more than ms_per_step; "
                                   "prefill blocks moved to the decode partition count in the decode stage")
    cfg_out = dict(workload_config(args.workload, B, world), pixels_dtype="f32",
                   l2="per-step working set (154 MB pixels, >1 GB activations, 1.4 GB weights) exceeds the 126 MB L2",
                   weights="seeded random init (no checkpoints offline)",
                   partition=(None if pipe is None or pipe.partition is None else
                              {"front_sms": pipe.partition.sms[0], "decode_sms": pipe.partition.sms[1],
                               "prefill_blocks_on_decode_partition": pipe.prefill_defer,
                               "ends": "first front / last decode loop of a run on the whole device"}),
                   timed_region="K steps back to back through the serving loop, incl. the drain of the last decode")
    if partition_note:
        cfg_out["partition_note"] = partition_note
    out = {
        "metric": w["metric"], "value": value, "unit": w["unit"], "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "p50_ms": statistics.median(per),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": cfg_out, "clocks": clocks,
        "e2e": {"value": units * args.steps / (e2e_ms * 1e-3), "unit": w["unit"], "ms_per_step": e2e_ms / args.steps,
                "p50_ms": statistics.median(e2e_per), "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "api": api},
        "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
        "model_tflops": value * (work["vit"] if vit_only else work["total"]) / 1e12,
        "roofline": roof,
    }
    if roof is not None and roof.get("batch_latency_p50_ms") is not None:
        # BASELINE.json's metric also names the p50 latency: one batch through the serving loop (image tower start -> last
        # decode step), median over the timed batches. `p50_ms` above is the steady-state period between finished batches.
        out["latency_p50_ms"] = roof["batch_latency_p50_ms"]
    if per_rank is not None:
        out["per_rank"] = per_rank
    if multi_gpu_exact is not None:
        out["multi_gpu_exact"] = multi_gpu_exact
    if world > 1:
        out["collective"] = ("none (--no-gather)" if not gather else
                             f"in-place ncclAllGather of the fp16 prefix [{B},40,1024] per rank ({B * 40 * 1024 * 2} bytes) through "
                             f"cc_allgather_prefix on a side stream, NCCL {lib.cc_nccl_version()}")
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        R, vit_w, map_w, lm_w, vcfg, mcfg, gcfg = cpu_reference_setup(state)
        n = args.cpu_captions if not vit_only else 8
        if vit_only:
            cpu_encode(R, vit_w, vcfg, px_host[:1])
            t0 = time.perf_counter()
            for i in range(n):
                cpu_encode(R, vit_w, vcfg, px_host[i:i + 1])
            dt = time.perf_counter() - t0
            sample = f"first {n} images of the same batch, one at a time, fp32"
        else:
            cpu_caption(R, vit_w, map_w, lm_w, vcfg, mcfg, gcfg, px_host[:1], beam)  # warm-up
            t0 = time.perf_counter()
            cpu_tokens = [cpu_caption(R, vit_w, map_w, lm_w, vcfg, mcfg, gcfg, px_host[i:i + 1], beam) for i in range(n)]
            dt = time.perf_counter() - t0
            gpu_tokens = tok_host[:n].tolist()
            agree = sum(int(gpu_tokens[i][:len(cpu_tokens[i])] == list(cpu_tokens[i])) for i in range(n))
            sample = (f"first {n} images of the same batch, one at a time (reference is batch-size-1), fp32, no KV cache; "
                      f"{agree}/{n} captions token-identical to the GPU run")
        out["cpu_baseline"] = {"value": n / dt, "unit": w["unit"], "cores": cores, "kind": "port", "sample": sample}
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
