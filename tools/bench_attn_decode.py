"""Decode attention microbench (bs 32, 32 heads, ctx ~1150): the three-kernel sequence vs the fused kernel, in a CUDA graph."""
import math, sys, torch
sys.path.insert(0, ".")
from crab_b200 import ops
ops.init(0)
dev = torch.device("cuda:0")
B, H, KVH, hd, ctx, past = 32, 32, 32, 128, 1280, 1150
nq = H * hd
g = torch.Generator(device=dev).manual_seed(1)
qkv = torch.randn(B, 3 * nq, generator=g, device=dev).to(torch.bfloat16)
caches = [(torch.randn(B, KVH, ctx, hd, generator=g, device=dev).to(torch.bfloat16),
           torch.randn(B, KVH, ctx, hd, generator=g, device=dev).to(torch.bfloat16)) for _ in range(2)]
ra = (torch.randn(11, nq + 32, generator=g, device=dev) / 64).to(torch.bfloat16)
rope = ops.rope_table(ctx, hd, 10000.0, dev)
pd = torch.tensor([past], dtype=torch.int32, device=dev)
ld = torch.tensor([past + 1], dtype=torch.int32, device=dev)
o = torch.zeros(B, nq + 32, dtype=torch.bfloat16, device=dev)
ws = torch.empty(B * KVH * 11, dtype=torch.float32, device=dev)
cnt = torch.zeros(B, dtype=torch.int32, device=dev)
q2 = qkv.clone()

def seq(kc, vc):
    ops.rope_kv_append(q2, rope, kc, vc, B, 1, H, KVH, hd, past=0, past_dev=pd)
    ops.attn_decode(q2, kc, vc, o[:, :nq], B=B, H=H, KVH=KVH, head_dim=hd, scale=hd ** -0.5, len_dev=ld)
    ops.row_norm_loraz(o[:, :nq], ra=ra[:, :nq], groups=1, z=o[:, nq:], scale=2.0)

def attn_only(kc, vc):
    ops.attn_decode(q2, kc, vc, o[:, :nq], B=B, H=H, KVH=KVH, head_dim=hd, scale=hd ** -0.5, len_dev=ld)

def fused(kc, vc, lora=True):
    ops.attn_decode_fused(qkv, rope, kc, vc, o[:, :nq], B=B, H=H, KVH=KVH, head_dim=hd, scale=hd ** -0.5, past_dev=pd,
                          ra=ra[:, :nq] if lora else None, z=o[:, nq:] if lora else None, lora_scale=2.0,
                          lora_ws=ws if lora else None, lora_counters=cnt if lora else None)

def time_graph(fn, n=16, reps=5):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(*caches[0])
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(n):
            fn(*caches[i % 2])
    gr.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e3 / n

kv_bytes = 2 * B * KVH * (past + 1) * hd * 2
for name, fn in [("attn only", attn_only), ("rope+attn+row(o)", seq), ("fused +lora", fused), ("fused no lora", lambda a, b: fused(a, b, False))]:
    us = time_graph(fn)
    print(f"{name:18s}: {us:7.2f} us   {kv_bytes / us / 1e3:7.1f} GB/s (KV bytes only)", flush=True)
