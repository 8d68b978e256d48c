"""In-kernel timeline of the decode step's five launches per layer (qkv | attention | o | gate/up | down) inside the CUDA-graph
replay: every CTA stamps %globaltimer at a handful of points (crab_debug_trace in include/crab_b200.h), ~1 us of overhead per
CTA.  Answers "where do the microseconds between two dependent weight streams go" (ramp, dependency wait, first operand, last
MMA, split-K exchange, exit).
    python tools/trace_skinny.py [--bs 32] [--layers 8] [--ctx 1150] [--layer 4]
Writes gpurun_out/skinny_timeline_bs<bs>.txt"""
import argparse, ctypes as C, os, sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench as BN  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--bs", type=int, default=32)
ap.add_argument("--layers", type=int, default=8)
ap.add_argument("--ctx", type=int, default=1150)
ap.add_argument("--layer", type=int, default=4)
ap.add_argument("--backbone", default="llama")
ap.add_argument("--tag", default="")
a = ap.parse_args()
os.environ.setdefault("CRAB_DECODE_MODE", "skinny")
from crab_b200 import lib, ops
from crab_b200.engine import CrabConfig, CrabEngine, DecoderConfig
from crab_b200.models.unified_arch import decoder_manifest
dev = torch.device("cuda:0")
b = dict(BN.BACKBONES[a.backbone]); b["layers"] = a.layers; b["vocab"] = b["base_vocab"] + 17
dcfg = DecoderConfig(hidden=b["hidden"], inter=b["inter"], layers=a.layers, heads=b["heads"], kv_heads=b["kv_heads"], head_dim=b["head_dim"],
                     vocab=b["vocab"], rope_theta=b["rope_theta"], qkv_bias=b["qkv_bias"])
eng = CrabEngine(BN.LazySynthSD(decoder_manifest(dcfg), 42, dev), CrabConfig(decoder=dcfg, max_ctx=1280), dev, load_encoders=False)
emb = torch.randn(a.bs, 64, b["hidden"], device=dev).to(torch.bfloat16)
eng.prefill(emb)
eng.cur_len = a.ctx
L = lib.load()
L.crab_debug_trace.argtypes = [C.c_void_p, C.c_int, C.c_int]
L.crab_debug_trace.restype = C.c_int
NS, NC = 4 * (5 * a.layers + 1) + 8, 2048
buf = torch.zeros(NS * NC * 16, dtype=torch.int64, device=dev)
assert L.crab_debug_trace(C.c_void_p(buf.data_ptr()), NS, NC) == 0
eng.begin_decode(a.bs, use_graph=True)          # eager warm-up + capture: the graph's launches keep their slots
L.crab_debug_trace(None, 0, 0)
for _ in range(3):
    eng.decode_step()
torch.cuda.synchronize()
buf.zero_()
eng.cur_len = a.ctx; eng.past_dev.fill_(a.ctx); eng.len_dev.fill_(a.ctx + 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.decode_step(); e1.record(); torch.cuda.synchronize()
t = buf.view(NS, NC, 16).cpu().numpy().astype(np.int64)
used = [i for i in range(NS) if (t[i, :, 0] > 0).any()]
out = [f"in-kernel timeline: backbone {a.backbone} bs {a.bs} layers {a.layers} ctx {a.ctx} mode {os.environ['CRAB_DECODE_MODE']} "
       f"splits {os.environ.get('CRAB_SKINNY_SPLITS', 'default')} pdl {os.environ.get('CRAB_PDL', 'default')}: "
       f"{len(used)} traced launches in the replay, traced step {e0.elapsed_time(e1) * 1e3:.1f} us"]
PL = len(used) // a.layers          # traced launches per layer: 5, or 4 when the attention runs in the (untraced) flash kernel
assert PL in (4, 5) and len(used) - PL * a.layers in (0, 1), (len(used), a.layers)   # + lm_head unless its grid exceeds a trace slot
names = ["qkv", "attn", "o", "gate_up", "down"] if PL == 5 else ["qkv", "o", "gate_up", "down"]
lab = {0: "entry", 1: "w-ring armed", 2: "dep wait over", 3: "first operand", 4: "last MMA / loop end", 5: "flag seen/raised", 6: "split-K landed", 7: "exit", 8: "ring free (all ranks)", 9: "partials sent", 10: "epilogue loop done", 11: "producer thread done", 12: "MMA thread done", 13: "finish: start", 14: "stats: ticket taken", 15: "stats: partials re-read"}
base = used[PL * a.layer]
origin = t[base, :, 0][t[base, :, 0] > 0].min()
out.append(f"layer {a.layer}: us relative to the first CTA of its qkv launch entering; per stamp: min / median / max over CTAs (count)")
prev_exit = None
for j in range(PL + 1):
    if PL * a.layer + j >= len(used):
        break
    s = used[PL * a.layer + j]
    nm = names[j % PL] + (" (next layer)" if j == PL else "")
    m = t[s]
    n = int((m[:, 0] > 0).sum())
    out.append(f"  {nm}: {n} CTAs")
    for k in (0, 1, 2, 3, 12, 11, 4, 5, 8, 9, 6, 10, 7):
        v = m[:, k][m[:, k] > 0]
        if v.size:
            r = (v - origin) / 1e3
            out.append(f"      {lab[k]:22s} {r.min():8.2f} {np.median(r):8.2f} {r.max():8.2f}   ({v.size})")
    tiles = {"qkv": -(-(b["heads"] + 2 * b["kv_heads"]) * b["head_dim"] // 128), "gate_up": -(-2 * b["inter"] // 128), "down": -(-b["hidden"] // 128)}.get(names[j % PL])
    ns = 0
    if tiles:
        for S_ in (8, 4, 2, 1):
            rest = n - tiles * S_
            if rest > 0 and rest % S_ == 0 and rest // S_ <= 8:
                ns = rest
                break
    if ns:
        out.append(f"      statistics clusters (first {ns} CTAs, cluster size {S_}):")
        for k in (0, 2, 3, 4, 8, 9, 6, 14, 15, 5, 7):
            v = m[:ns, k][m[:ns, k] > 0]
            if v.size:
                r = (v - origin) / 1e3
                out.append(f"        {lab[k]:22s} {r.min():8.2f} {np.median(r):8.2f} {r.max():8.2f}   ({v.size})")
    ex = (m[:, 7][m[:, 7] > 0] - origin) / 1e3
    if prev_exit is not None:
        out.append(f"      -> last exit of previous launch to last exit of this one: {ex.max() - prev_exit:.2f} us")
    prev_exit = ex.max()
per_layer = []
for l in range(1, a.layers - 1):
    s0, s1 = used[PL * l], used[PL * (l + 1)]
    per_layer.append((t[s1, :, 0][t[s1, :, 0] > 0].min() - t[s0, :, 0][t[s0, :, 0] > 0].min()) / 1e3)
out.append(f"layer period (qkv entry to next qkv entry), layers 1..{a.layers - 2}: " + " ".join(f"{x:.1f}" for x in per_layer))
# critical path: last exit of each launch, averaged over the middle layers
acc = {n: [] for n in names}
for l in range(1, a.layers - 1):
    prev = None
    for j in range(PL + 1):
        s = used[PL * l + j]
        ex = t[s, :, 7][t[s, :, 7] > 0].max() / 1e3
        if prev is not None:
            acc[names[j % PL]].append(ex - prev)
        prev = ex
out.append("last-exit to last-exit per launch, mean over the middle layers (us): " + ", ".join(f"{k} {np.mean(v):.1f}" for k, v in acc.items() if v))
txt = "\n".join(out)
print(txt)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / f"skinny_timeline_bs{a.bs}{a.tag}.txt").write_text(txt + "\n")
