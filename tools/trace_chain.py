"""Per-role globaltimer trace of one crab_decode_chain launch (diagnostics): where each CTA's MMA issuer and epilogue spend
their time.  Usage: python tools/trace_chain.py [gu|o|d|q|all] [cluster]"""
import os, sys, math
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.argv = [sys.argv[0]] + sys.argv[1:]
which = sys.argv[1] if len(sys.argv) > 1 else "gu"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4
from crab_b200 import ops
dev = torch.device("cuda:0"); ops.init(0)
D, F, nq, nk, B = 4096, 11008, 4096, 4096, 32
def rnd(*shape, scale=1.0): return (scale * torch.randn(*shape, device=dev)).to(torch.bfloat16)
def make(N, K, kext, linears, gamma, swiglu=False):
    w = torch.zeros((N, K + kext), device=dev, dtype=torch.bfloat16); w[:, :K] = rnd(N, K, scale=1 / math.sqrt(K))
    pk = ops.pack_skinny_weight(w, k=K + kext, swiglu=swiglu)
    st = ops.pack_chain_stats(rnd(11 * linears, K, scale=1 / math.sqrt(K)), (1 + 0.1 * torch.randn(K, device=dev)) if gamma else None)
    return pk, st
wo, so = make(D, nq, 32, 1, False); wgu, sgu = make(2 * F, D, 64, 2, True, swiglu=True); wd, sd_ = make(D, F, 32, 1, False); wq, sq = make(nq + 2 * nk, D, 96, 3, True)
at = torch.zeros((B, nq + 32), device=dev, dtype=torch.bfloat16); at[:, :nq] = rnd(B, nq)
x = rnd(B, D); hh = rnd(B, F); qkv = torch.empty((B, nq + 2 * nk), device=dev, dtype=torch.bfloat16)
z = {k: torch.zeros((32, 128), device=dev, dtype=torch.bfloat16) for k in "o gu d q".split()}
rs = {k: torch.zeros(32, device=dev, dtype=torch.float32) for k in "gu q".split()}
cnt = torch.zeros(288, dtype=torch.int32, device=dev)
P = {"o": lambda: ops.ChainPhase(at, wo, x, k=nq, z=at[:, nq:], kext=32, residual=x),
     "gu": lambda: ops.ChainPhase(x, wgu, hh, k=D, z=z["gu"], kext=64, stats=sgu, stats_linears=2, norm=True, eps=1e-6, lora_scale=2.0, rstd=rs["gu"], act=ops.ACT_SWIGLU),
     "d": lambda: ops.ChainPhase(hh, wd, x, k=F, z=z["d"], kext=32, stats=sd_, stats_linears=1, lora_scale=2.0, residual=x),
     "q": lambda: ops.ChainPhase(x, wq, qkv, k=D, z=z["q"], kext=96, stats=sq, stats_linears=3, norm=True, eps=1e-6, lora_scale=2.0, rstd=rs["q"])}
keys = ["o", "gu", "d", "q"] if which == "all" else [which]
phases = [P[k]() for k in keys]
nct = ops.decode_chain_max_clusters(S) * S
trace = torch.zeros((nct, 32, 16), dtype=torch.int64, device=dev)
for _ in range(3): ops.decode_chain(phases, B, cnt, S)
torch.cuda.synchronize()
os.environ["CRAB_CHAIN_TRACE"] = hex(trace.data_ptr())
ops.decode_chain(phases, B, cnt, S)
torch.cuda.synchronize()
os.environ.pop("CRAB_CHAIN_TRACE")
t = trace.cpu()
valid = t[:, :, 8] > 0
t0 = int(t[:, :, 8][valid].min())
print(f"{which} cluster {S}: {nct} CTAs; all times us relative to the first MMA-item start")
def us(v): return (v - t0) / 1e3
for cta in [0, 1, 2, 3, 4, 64, nct - 1]:
    print(f"CTA {cta}")
    for it in range(32):
        r = t[cta, it]
        if r[8] == 0: break
        print(f"  item {it} (phase*1000+idx {int(r[7])}, {int(r[12])} kb): MMA start {us(r[8]):7.2f} tempty-wait {(r[9]-r[8])/1e3:5.2f} issue-span {(r[10]-r[9])/1e3:6.2f} (waiting for data {r[11]/1e3:6.2f})"
              f" | EPI idle-until {us(r[0]):7.2f} tfull@ {us(r[1]):7.2f} sent +{(r[2]-r[1])/1e3:5.2f} pfull +{(r[3]-r[2])/1e3:5.2f} math+stores +{(r[4]-r[3])/1e3:5.2f} publish +{(r[5]-r[4])/1e3:5.2f}")
end = t[:, :, 5].max()
print(f"last epilogue done at {us(int(end)):.2f} us")
# aggregate over CTAs: mean per-item durations
import statistics
spans = {"tempty": [], "issue": [], "datawait": [], "sent": [], "pfull": [], "math": [], "publish": [], "epi_total": []}
for cta in range(nct):
    for it in range(32):
        r = t[cta, it]
        if r[8] == 0 or r[5] == 0: continue
        spans["tempty"].append((r[9]-r[8])/1e3); spans["issue"].append((r[10]-r[9])/1e3); spans["datawait"].append(r[11]/1e3)
        spans["sent"].append((r[2]-r[1])/1e3); spans["pfull"].append((r[3]-r[2])/1e3); spans["math"].append((r[4]-r[3])/1e3)
        spans["publish"].append((r[5]-r[4])/1e3); spans["epi_total"].append((r[5]-r[1])/1e3)
for k, v in spans.items():
    v = [float(a) for a in v]
    print(f"  {k:10s} mean {statistics.mean(v):6.2f}  max {max(v):6.2f}  n {len(v)}")
# per-phase timeline over all CTAs
print("per-phase timeline (us): first MMA-item start | first tfull | median tfull | last MMA issue end | first epilogue done | last epilogue done | items")
phs = {}
for cta in range(nct):
    for it in range(32):
        r = t[cta, it]
        if r[8] == 0: continue
        ph = int(r[7]) // 1000
        d = phs.setdefault(ph, dict(ms=[], me=[], ee=[], tf=[]))
        d["ms"].append(us(int(r[8]))); d["me"].append(us(int(r[10])))
        if r[5] > 0: d["ee"].append(us(int(r[5])))
        if r[1] > 0: d["tf"].append(us(int(r[1])))
for ph in sorted(phs):
    d = phs[ph]
    tf = sorted(d["tf"]) or [0]
    print(f"  phase {ph}: {min(d['ms']):7.2f} | {tf[0]:7.2f} | {tf[len(tf)//2]:7.2f} | {max(d['me']):7.2f} | {min(d['ee']) if d['ee'] else 0:7.2f} | {max(d['ee']) if d['ee'] else 0:7.2f} | {len(d['ms'])}")
