"""CTA-pair (cta_group::2) GEMM vs the single-CTA kernel and cuBLAS: correctness on the hot shapes, then timing.
Run as:  CRAB_GEMM_2CTA=1 python tools/check_gemm2.py   (the mode is read once per process)"""
import os, sys, torch
sys.path.insert(0, ".")
from crab_b200 import ops
ops.init(0)
dev = torch.device("cuda:0")
mode = os.environ.get("CRAB_GEMM_2CTA", "0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
torch.manual_seed(0)

def check(M, N, K, **kw):
    a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    w = torch.randn(N, K, device=dev, dtype=torch.bfloat16) / K ** 0.5
    out = ops.gemm(a, w, block_n=256, **kw)
    torch.cuda.synchronize()
    ref = torch.matmul(a, w.t()).float()
    if "bias" in kw: ref = ref + kw["bias"]
    if "residual" in kw: ref = ref + kw["residual"].float()
    err = (out.float() - ref).abs().max().item() / (ref.abs().max().item() + 1e-6)
    print(f"  check M={M} N={N} K={K} {list(kw)}: rel max err {err:.3e}", flush=True)
    assert err < 2e-2, err

print("CRAB_GEMM_2CTA =", mode)
check(4096, 4096, 1024)
check(4096 + 77, 4096 + 264, 1000)                      # ragged M / N / K tails
check(8192, 4096, 4128, bias=torch.randn(4096, device=dev), residual=torch.randn(8192, 4096, device=dev, dtype=torch.bfloat16))
check(34752, 12288, 4192)
for name, M, N, K in [("llama_qkv", 34752, 12288, 4192), ("llama_o", 34752, 4096, 4128), ("llama_gateup", 34752, 22016, 4160),
                      ("llama_down", 34752, 4096, 11040), ("clip_fc1", 65792, 4096, 1024), ("sq8192", 8192, 8192, 8192)]:
    a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    w = torch.randn(N, K, device=dev, dtype=torch.bfloat16) / K ** 0.5
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3): ops.gemm(a, w, out=out, block_n=256)
    ts = []
    for _ in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gemm(a, w, out=out, block_n=256); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    print(f"  {name:14s} 2cta={mode}: {t:8.3f} ms  {2.0 * M * N * K / t / 1e9:7.1f} TFLOP/s", flush=True)
    del a, w, out
