"""Developer timing (GPU box): CaptionPipeline sequential vs SM-partitioned at the bench shape; token equality between modes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from clipcap_b200.encoders.clip import CLIPModel, ViTImageTower
from clipcap_b200.encoders.config import EncoderConfig
from clipcap_b200.model import ClipCapModelPrefixOnly, Config
from clipcap_b200.pipeline import CaptionPipeline

dev = torch.device("cuda:0")
B = int(os.environ.get("B", "256"))
K = int(os.environ.get("STEPS", "10"))
state = bench.synthetic_state()
tower = ViTImageTower(); tower.load_state_dict(state["vit"], strict=True)
encode_fn = CLIPModel(tower).eval().to(dev)
cfg = Config(language_model="gpt2-medium", prefix_length=40, projection_length=10, transformer_layers=8,
             transformer_attention_heads=8, encoder_config=EncoderConfig(encoder_embedding_size=768))
model = ClipCapModelPrefixOnly(cfg)
sd = {f"transformer_mapper.{k}": v for k, v in state["mapper"].items()}
sd.update({f"language_model.{k}": v for k, v in state["lm"].items()})
model.load_state_dict(sd, strict=True); model = model.eval().to(dev)
px = [bench.synthetic_pixels(B, 1234 + i).to(dev) for i in range(2)]
pxh = [p.cpu().pin_memory() for p in px]
ref = None
for sms in [int(x) for x in os.environ.get("PART", "0,24,32,40").split(",")]:
    pipe = CaptionPipeline(encode_fn, model, B, 224, 20, 50256, dev, partition_sms=sms,
                           prefill_defer=int(os.environ.get("DEFER", "0")))
    outs = [(t.clone(), l.clone()) for t, l in pipe.run((px[i & 1] for i in range(4)), resident=True)]
    torch.cuda.synchronize()
    if ref is None:
        ref = outs
    same = all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) for a, b in zip(outs, ref))
    for name, src, res in (("resident", px, True), ("host", pxh, False)):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        n = sum(1 for _ in pipe.run((src[i & 1] for i in range(K)), resident=res))
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3 / K
        print(f"partition_sms={sms} ({'off' if pipe.partition is None else pipe.partition.sms}) {name}: {dt:.2f} ms/step = {B / dt * 1e3:.0f} captions/s; "
              f"tokens identical to sequential: {same}", flush=True)
    del pipe
