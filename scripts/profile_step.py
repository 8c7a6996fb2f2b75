"""One step of the bench workload between cudaProfilerStart/Stop (use with `ncu --profile-from-start off`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from clipcap_b200.encoders.clip import CLIPModel, ViTImageTower
from clipcap_b200.encoders.config import EncoderConfig
from clipcap_b200.inference.base import generate_greedy_tokens
from clipcap_b200.model import ClipCapModelPrefixOnly, Config
from oracle import synth

B = int(os.environ.get("B", "256"))
stage = os.environ.get("STAGE", "all")
dev = torch.device("cuda:0")
state = bench.synthetic_state()
tower = ViTImageTower(); tower.load_state_dict(state["vit"], strict=True)
encode_fn = CLIPModel(tower).eval().to(dev)
cfg = Config(language_model="gpt2-medium", prefix_length=40, projection_length=10, transformer_layers=8,
             transformer_attention_heads=8, encoder_config=EncoderConfig(encoder_embedding_size=768))
model = ClipCapModelPrefixOnly(cfg)
sd = {f"transformer_mapper.{k}": v for k, v in state["mapper"].items()}
sd.update({f"language_model.{k}": v for k, v in state["lm"].items()})
model.load_state_dict(sd, strict=True); model = model.eval().to(dev)
px = synth.pixels(B, 224).to(dev)

def step():
    emb = encode_fn(px)
    prefix = model.transformer_mapper(emb)
    return generate_greedy_tokens(model, prefix, 20, 50256)

for _ in range(3): step()
torch.cuda.synchronize()
emb = encode_fn(px); prefix = model.transformer_mapper(emb); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
if stage == "all": step()
elif stage == "vit": encode_fn(px)
elif stage == "mapper": model.transformer_mapper(emb)
elif stage == "lm": generate_greedy_tokens(model, prefix, 20, 50256)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
# stage timings with events (not under the profiler's influence when run standalone)
def t(fn, n=5):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
if os.environ.get("TIMES", "1") == "1":
    print("vit ms", t(lambda: encode_fn(px)))
    print("mapper ms", t(lambda: model.transformer_mapper(emb)))
    print("generate ms", t(lambda: generate_greedy_tokens(model, prefix, 20, 50256)))
