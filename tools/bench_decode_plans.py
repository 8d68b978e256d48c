"""A/B of programmatic-dependent-launch plans over the CUDA-graph decode step (bs 32, LLaMA-7B dims + hyper-LoRA,
context 1086 -> 1213).  One process: the decoder is prefetched once, then each plan re-captures the decode graph and
replays 127 steps from the same state; the generated ids must be identical across plans.

usage: python tools/bench_decode_plans.py [--layers 32] [--plans 0,0 1,1 3,1 ...]
"""
import argparse
import hashlib
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import LazySynthSD  # noqa: E402
from crab_b200 import ops  # noqa: E402
from crab_b200.engine import CrabConfig, CrabEngine, DecoderConfig  # noqa: E402
from crab_b200.models.unified_arch import full_manifest, special_token_ids  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=32)
ap.add_argument("--bs", type=int, default=32)
ap.add_argument("--S", type=int, default=1086)
ap.add_argument("--new", type=int, default=128)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--plans", nargs="*", default=["0,0", "1,1", "3,1", "35,35", "19,17", "31,29", "31,31", "63,63", "0,0"])
args = ap.parse_args()

dev = torch.device("cuda:0")
cfg = CrabConfig(decoder=DecoderConfig(layers=args.layers), max_ctx=(args.S + args.new + 7) // 8 * 8,
                 special_ids=special_token_ids(32000))
man = {k: v for k, v in full_manifest(cfg).items() if k.startswith("model.layers.") or k in
       ("model.embed_tokens.weight", "model.norm.weight", "lm_head.weight")}
t0 = time.time()
eng = CrabEngine(LazySynthSD(man, 42, dev), cfg, dev, load_encoders=False)
torch.cuda.synchronize()
print(f"load {time.time() - t0:.1f}s", flush=True)
g = torch.Generator(device=dev).manual_seed(7)
emb = torch.randn((args.bs, args.S, cfg.decoder.hidden), generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
_, nxt0 = eng.prefill(emb.clone())
nxt0 = nxt0.clone()
torch.cuda.synchronize()

results = []
for plan in args.plans:
    chain, after = (int(v) for v in plan.split(","))
    eng.pdl_chain, eng.pdl_after_attn = chain, after
    eng._graph = None
    best, digest = None, None
    for rep in range(args.reps + 1):  # first repetition = warm-up (captures the graph)
        eng.cur_len = args.S
        eng.next_ids.copy_(nxt0)
        eng.begin_decode(args.bs)
        out = torch.empty((args.bs, args.new - 1), device=dev, dtype=torch.int64)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for s in range(args.new - 1):
            _, nxt = eng.decode_step()
            out[:, s].copy_(nxt)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (args.new - 1)
        if rep > 0:
            best = ms if best is None else min(best, ms)
        digest = hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:12]
    results.append((plan, best, digest))
    print(f"plan chain={chain:2d} after_attn={after:2d}: {best:.3f} ms/step  ids {digest}", flush=True)
ref = results[0][2]
print("ids identical across plans:", all(r[2] == ref for r in results))
