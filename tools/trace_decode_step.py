"""CUPTI trace (torch.profiler / Kineto) of the decode step's CUDA-graph replay on the bench configuration: per-launch device
durations in launch order, so the per-kernel split of the step is evidence rather than a model.
    python tools/trace_decode_step.py [--bs 32] [--layers 32] [--mode skinny|chain] [--steps 4]
Writes gpurun_out/decode_trace_<mode>_bs<bs>.txt (one layer's launch sequence + per-class totals)."""
import argparse, json, os, sys, tempfile
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench as BN  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--bs", type=int, default=32)
ap.add_argument("--layers", type=int, default=32)
ap.add_argument("--mode", default="skinny")
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--ctx", type=int, default=1150)
ap.add_argument("--backbone", default="llama")
a = ap.parse_args()
os.environ["CRAB_DECODE_MODE"] = a.mode
from crab_b200 import ops
from crab_b200.engine import BeatsConfig, ClipConfig, CrabConfig, CrabEngine, DecoderConfig, QformerConfig
from crab_b200.models.unified_arch import decoder_manifest
dev = torch.device("cuda:0")
b = dict(BN.BACKBONES[a.backbone]); b["layers"] = a.layers; b["vocab"] = b["base_vocab"] + 17
dcfg = DecoderConfig(hidden=b["hidden"], inter=b["inter"], layers=a.layers, heads=b["heads"], kv_heads=b["kv_heads"], head_dim=b["head_dim"],
                     vocab=b["vocab"], rope_theta=b["rope_theta"], qkv_bias=b["qkv_bias"])
cfg = CrabConfig(decoder=dcfg, max_ctx=1280)
eng = CrabEngine(BN.LazySynthSD(decoder_manifest(dcfg), 42, dev), cfg, dev, load_encoders=False)
emb = torch.randn(a.bs, 64, b["hidden"], device=dev).to(torch.bfloat16)
eng.prefill(emb)
eng.cur_len = a.ctx
eng.begin_decode(a.bs, use_graph=True)
for _ in range(3):
    eng.decode_step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
eng.cur_len = a.ctx; eng.past_dev.fill_(a.ctx); eng.len_dev.fill_(a.ctx + 1)
e0.record()
for _ in range(16):
    eng.decode_step()
e1.record(); torch.cuda.synchronize()
step_ms = e0.elapsed_time(e1) / 16
eng.cur_len = a.ctx; eng.past_dev.fill_(a.ctx); eng.len_dev.fill_(a.ctx + 1)
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(a.steps):
        eng.decode_step()
    torch.cuda.synchronize()
tmp = tempfile.mktemp(suffix=".json")
prof.export_chrome_trace(tmp)
ev = [e for e in json.load(open(tmp))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
per = len(ev) // a.steps
out = [f"decode step trace: mode {a.mode} backbone {a.backbone} bs {a.bs} layers {a.layers} ctx {a.ctx}: {per} launches/step; graph step (CUDA events, not traced) {step_ms*1e3:.1f} us"]
step = ev[per:2 * per]   # second replay
t0 = step[0]["ts"]
out.append("launch sequence of layer 1 (start us rel., duration us, gap to previous end, kernel):")
prev_end = None
names = [e["name"][:60] for e in step]
per_layer = (per - 5) // a.layers if a.layers else per
for i, e in enumerate(step[: 1 + 2 * per_layer + 1]):
    gap = (e["ts"] - prev_end) if prev_end is not None else 0.0
    out.append(f"  {i:3d} {e['ts'] - t0:9.2f} {e['dur']:8.2f} {gap:7.2f}  {e['name'][:90]}")
    prev_end = e["ts"] + e["dur"]
tot = {}
for e in ev:
    k = e["name"].split("(")[0][:70]
    d = tot.setdefault(k, [0, 0.0]); d[0] += 1; d[1] += e["dur"]
busy = sum(v[1] for v in tot.values()) / a.steps
out.append(f"per-kernel totals per step (launches, us) — sum of kernel durations {busy:.1f} us, idle {(step_ms*1e3 - busy):.1f} us:")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    out.append(f"  {v[0] // a.steps:5d} {v[1] / a.steps:10.1f}  {k}")
# positional split of the weight-streaming launches (qkv / o / gate-up / down / head)
sk = [e for e in step if "gemm_skinny" in e["name"]]
if len(sk) == 4 * a.layers + 1:
    lab = ["qkv", "o", "gate_up", "down"]
    agg = {k: 0.0 for k in lab + ["lm_head"]}
    for i, e in enumerate(sk[:-1]):
        agg[lab[i % 4]] += e["dur"]
    agg["lm_head"] = sk[-1]["dur"]
    out.append("weight-streaming launches by position (us per step, us per launch): " + ", ".join(f"{k} {v:.1f} ({v / (a.layers if k != 'lm_head' else 1):.1f})" for k, v in agg.items()))
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
p = ROOT / "gpurun_out" / f"decode_trace_{a.mode}_{a.backbone}_bs{a.bs}.txt"
p.write_text("\n".join(out) + "\n")
print("\n".join(out))
