"""Isolated timing of the decode-step kernels (inside CUDA graphs, rotating buffers so nothing is L2-resident)."""
import json, math, sys, torch
sys.path.insert(0, ".")
from crab_b200 import ops

ops.init(0)
dev = torch.device("cuda:0")
res = {}


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
    return sorted(ts)[len(ts) // 2] * 1e3 / n  # us per launch


M = 32
for name, N, K in [("qkv", 12288, 4192), ("o", 4096, 4128), ("gateup", 22016, 4160), ("down", 4096, 11040), ("lm_head", 32024, 4096)]:
    nW = max(2, math.ceil((600 << 20) / (N * K * 2)))
    Ws = [torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.02 for _ in range(nW)]
    x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    out = torch.empty(M, N // 2 if name == "gateup" else N, device=dev, dtype=torch.bfloat16)
    act = ops.ACT_SWIGLU if name == "gateup" else ops.ACT_NONE
    Wp = [ops.pack_skinny_weight(W, swiglu=(name == "gateup")) for W in Ws]
    for splits in [0, 1, 2, 4, 8]:
        def fnp():
            for r in range(3):
                for W in Wp:
                    ops.gemm_skinny(x, W, act=act, out=out, splits=splits)
            return 3 * nW
        us = time_graph(fnp)
        print(f"skinny-PACKED {name:8s} ctas={splits:3d}: {us:8.2f} us  {(N * K * 2 + M * K * 2) / us / 1e3:8.1f} GB/s", flush=True)
        res[f"skinny_packed_{name}_c{splits}"] = dict(us=round(us, 2))
    del Wp
    # the prefill kernel at M=32 for comparison
    def fn2():
        for r in range(3):
            for W in Ws:
                ops.gemm(x, W, act=act, out=out)
        return 3 * nW
    us = time_graph(fn2)
    print(f"tcgen05-gemm(M=32) {name:8s}: {us:8.2f} us  {(N*K*2)/us/1e3:8.1f} GB/s", flush=True)
    del Ws

# decode attention: bs32, 32 heads, ctx 1150
B, H, hd, ctx_max = 32, 32, 128, 1216
caches = [(torch.randn(B, H, ctx_max, hd, device=dev, dtype=torch.bfloat16), torch.randn(B, H, ctx_max, hd, device=dev, dtype=torch.bfloat16)) for _ in range(3)]
q = torch.randn(B, 3 * H * hd, device=dev, dtype=torch.bfloat16)
o = torch.empty(B, H * hd, device=dev, dtype=torch.bfloat16)
for ctx in (1087, 1150, 1213):
    ld = torch.tensor([ctx], dtype=torch.int32, device=dev)
    for nsplit in (1, 2):
        def fn3():
            for r in range(4):
                for kc, vc in caches:
                    ops.attn_decode(q, kc, vc, o, B=B, H=H, KVH=H, head_dim=hd, scale=hd ** -0.5, len_dev=ld, nsplit=nsplit)
            return 12
        us = time_graph(fn3)
        gbs = B * H * ctx * hd * 2 * 2 / us / 1e3
        res[f"attn_decode_ctx{ctx}_s{nsplit}"] = dict(us=round(us, 2), gbs=round(gbs, 1))
        print(f"attn_decode ctx={ctx} nsplit={nsplit}: {us:8.2f} us  {gbs:8.1f} GB/s", flush=True)

# small kernels: latency inside a graph
D, F = 4096, 11008
x = torch.randn(M, D, device=dev, dtype=torch.bfloat16)
gamma = torch.ones(D, device=dev)
buf = torch.zeros(M, D + 96, device=dev, dtype=torch.bfloat16)
ra3 = torch.randn(33, D, device=dev, dtype=torch.bfloat16) * 0.01
hbuf = torch.zeros(M, F + 32, device=dev, dtype=torch.bfloat16)
rad = torch.randn(11, F, device=dev, dtype=torch.bfloat16) * 0.01
def fn4():
    for _ in range(50):
        ops.row_norm_loraz(x, gamma=gamma, eps=1e-6, y=buf[:, :D], ra=ra3, groups=3, z=buf[:, D:], scale=2.0)
    return 50
print(f"row_norm_loraz(D=4096, 3 groups): {time_graph(fn4):.2f} us")
def fn5():
    for _ in range(50):
        ops.row_norm_loraz(hbuf[:, :F], ra=rad, groups=1, z=hbuf[:, F:], scale=2.0)
    return 50
print(f"row_loraz(F=11008, 1 group): {time_graph(fn5):.2f} us")
kc, vc = caches[0]
tab = ops.rope_table(ctx_max, hd, 10000.0, dev)
pd = torch.tensor([1100], dtype=torch.int32, device=dev)
def fn6():
    for _ in range(50):
        ops.rope_kv_append(q, tab, kc, vc, B, 1, H, H, hd, past_dev=pd)
    return 50
print(f"rope_kv_append(decode): {time_graph(fn6):.2f} us")
json.dump(res, open("gpurun_out/bench_decode_kernels.json", "w"), indent=1)
