"""Micro-benchmark of crab_gemm_bf16 on the hot-path shapes (CUDA events, L2 flushed between iterations)."""
import json, sys, torch
sys.path.insert(0, ".")
from crab_b200 import ops

ops.init(0)
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
shapes = [
    ("llama_qkv", 34752, 12288, 4096 + 96), ("llama_o", 34752, 4096, 4096 + 32), ("llama_gateup", 34752, 22016, 4096 + 64),
    ("llama_down", 34752, 4096, 11008 + 32), ("clip_qkv", 65792, 3072, 1024), ("clip_fc1", 65792, 4096, 1024),
    ("clip_fc2", 65792, 1024, 4096), ("beats_fc1", 15360, 3072, 768), ("decode_qkv", 32, 12288, 4192), ("decode_down", 32, 4096, 11040),
    ("sq8192", 8192, 8192, 8192),
]
res = []
for name, M, N, K in shapes:
    a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    w = torch.randn(N, K, device=dev, dtype=torch.bfloat16) / K ** 0.5
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for bn in (64, 128, 256):
        for _ in range(3):
            ops.gemm(a, w, out=out, block_n=bn)
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.gemm(a, w, out=out, block_n=bn); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        tf = 2.0 * M * N * K / t / 1e9
        gbs = (M * K + N * K + M * N) * 2 / t / 1e6
        res.append(dict(name=name, M=M, N=N, K=K, bn=bn, ms=round(t, 4), tflops=round(tf, 1), gbs=round(gbs, 1)))
        print(res[-1], flush=True)
    # cuBLAS yardstick
    for _ in range(3): torch.matmul(a, w.t(), out=out)
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, w.t(), out=out); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    print(dict(name=name, impl="cublas", ms=round(t, 4), tflops=round(2.0 * M * N * K / t / 1e9, 1)), flush=True)
    del a, w, out
json.dump(res, open("gpurun_out/bench_gemm.json", "w"), indent=1)
