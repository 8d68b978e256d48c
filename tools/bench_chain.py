"""Micro-benchmark of crab_decode_chain on one decoder layer's four linears (LLaMA-7B dims by default): the whole chain and
each phase alone, per cluster size, replayed from a CUDA graph.  Prints us per launch and GB/s of weight bytes.

    python tools/bench_chain.py [--qwen] [--bs 32]
"""
import argparse
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from crab_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--qwen", action="store_true")
ap.add_argument("--bs", type=int, default=32)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--clusters", default="4")
ap.add_argument("--stages", default="0", help="comma list of ring depths (0 = as many as fit)")
ap.add_argument("--debug", default="0", help="comma list of CRAB_CHAIN_DEBUG masks (1 skip publish, 2 skip X loads, 4 skip MMAs)")
ap.add_argument("--sets", default="o,gu,d,q,all")
ap.add_argument("--prefetch", default="24", help="comma list of L2 prefetch distances (blocks)")
args = ap.parse_args()
dev = torch.device("cuda:0")
ops.init(0)
D, F, nq, nk = (3584, 18944, 3584, 512) if args.qwen else (4096, 11008, 4096, 4096)
B = args.bs


def rnd(*shape, scale=1.0):
    return (scale * torch.randn(*shape, device=dev)).to(torch.bfloat16)


def make(N, K, kext, linears, gamma, swiglu=False):
    w = torch.zeros((N, K + kext), device=dev, dtype=torch.bfloat16)
    w[:, :K] = rnd(N, K, scale=1 / math.sqrt(K))
    w[:, K:K + 24 * linears] = rnd(N, 24 * linears, scale=0.05)
    pk = ops.pack_skinny_weight(w, k=K + kext, swiglu=swiglu)
    st = ops.pack_chain_stats(rnd(11 * linears, K, scale=1 / math.sqrt(K)), (1 + 0.1 * torch.randn(K, device=dev)) if gamma else None)
    return pk, st


wo, so = make(D, nq, 32, 1, False)
wgu, sgu = make(2 * F, D, 64, 2, True, swiglu=True)
wd, sd_ = make(D, F, 32, 1, False)
wq, sq = make(nq + 2 * nk, D, 96, 3, True)
at = torch.zeros((B, nq + 32), device=dev, dtype=torch.bfloat16)
at[:, :nq] = rnd(B, nq)
x = rnd(B, D)
hh = rnd(B, F)
qkv = torch.empty((B, nq + 2 * nk), device=dev, dtype=torch.bfloat16)
z = {k: torch.zeros((32, 128), device=dev, dtype=torch.bfloat16) for k in "o gu d q".split()}
rs = {k: torch.zeros(32, device=dev, dtype=torch.float32) for k in "gu q".split()}
cnt = torch.zeros(288, dtype=torch.int32, device=dev)

P = {
    "o": lambda: ops.ChainPhase(at, wo, x, k=nq, z=at[:, nq:], kext=32, residual=x),
    "gu": lambda: ops.ChainPhase(x, wgu, hh, k=D, z=z["gu"], kext=64, stats=sgu, stats_linears=2, norm=True, eps=1e-6, lora_scale=2.0,
                                 rstd=rs["gu"], act=ops.ACT_SWIGLU),
    "d": lambda: ops.ChainPhase(hh, wd, x, k=F, z=z["d"], kext=32, stats=sd_, stats_linears=1, lora_scale=2.0, residual=x),
    "q": lambda: ops.ChainPhase(x, wq, qkv, k=D, z=z["q"], kext=96, stats=sq, stats_linears=3, norm=True, eps=1e-6, lora_scale=2.0,
                                rstd=rs["q"]),
}
# a big buffer written between replays is not needed: one layer's weights (410 MB) exceed the 126 MB L2


def time_graph(fn, reps):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (3 * reps)


import os

ALL = {"o": ["o"], "gu": ["gu"], "d": ["d"], "q": ["q"], "all": ["o", "gu", "d", "q"]}
for S in [int(c) for c in args.clusters.split(",")]:
    for st in args.stages.split(","):
        os.environ["CRAB_CHAIN_STAGES"] = st
        for dbg, pfd in [(d_, p_) for d_ in args.debug.split(",") for p_ in args.prefetch.split(",")]:
            os.environ["CRAB_CHAIN_DEBUG"] = dbg
            pass
            mc = ops.decode_chain_max_clusters(S)
            print(f"cluster {S} stages {st} debug {dbg} prefetch {pfd}: max co-resident clusters {mc} ({mc * S} CTAs)")
            for name in args.sets.split(","):
                keys = ALL[name]
                if int(dbg) and len(keys) > 1:
                    continue  # debug masks break the dependencies between phases
                phases = [P[k]() for k in keys]
                nbytes = sum(ph.weight_bytes() for ph in phases)
                us = time_graph(lambda: ops.decode_chain(phases, B, cnt, S), args.reps)
                print(f"   {name:6s} {us:8.2f} us   {nbytes / us / 1e3:7.0f} GB/s   ({nbytes / 1e6:.1f} MB)")
os.environ["CRAB_CHAIN_DEBUG"] = "0"
