"""A/B: decode-chain microbench with PDL on/off (CRAB_PDL env is read at first use, so run as two processes)."""
import os, sys, math, torch
sys.path.insert(0, ".")
from crab_b200 import ops
ops.init(0)
dev = torch.device("cuda:0")

def time_graph(fn, reps=5):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        n = fn()
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e3 / n

M = 32
print("CRAB_PDL =", os.environ.get("CRAB_PDL", "1"))
for name, N, K in [("qkv", 12288, 4192), ("o", 4096, 4128), ("gateup", 22016, 4160), ("down", 4096, 11040)]:
    nW = max(2, math.ceil((600 << 20) / (N * K * 2)))
    Ws = [ops.pack_skinny_weight(torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.02, swiglu=(name == "gateup")) for _ in range(nW)]
    x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    out = torch.empty(M, N // 2 if name == "gateup" else N, device=dev, dtype=torch.bfloat16)
    act = ops.ACT_SWIGLU if name == "gateup" else ops.ACT_NONE
    def fn():
        for r in range(3):
            for W in Ws:
                ops.gemm_skinny(x, W, act=act, out=out)
        return 3 * nW
    us = time_graph(fn)
    print(f"  skinny-packed {name:7s}: {us:7.2f} us  {(N*K*2)/us/1e3:7.1f} GB/s", flush=True)
    del Ws
# a chain of tiny kernels: launch-to-launch latency
x = torch.randn(32, 4096, device=dev, dtype=torch.bfloat16); g = torch.ones(4096, device=dev)
y = torch.empty_like(x)
def fn2():
    for _ in range(200):
        ops.rmsnorm(x, g, 1e-6, out=y)
    return 200
print(f"  rmsnorm chain: {time_graph(fn2):.2f} us/launch")
