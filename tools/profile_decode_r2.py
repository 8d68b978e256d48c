"""Round-2 decode kernels, one launch each at the bench's shapes (bs 32, LLaMA-7B dims), for
`ncu --set full --clock-control none --import-source on -o gpurun_out/<tag> python tools/profile_decode_r2.py`:
the fused weight-streaming linears (norm as epilogue scale + in-launch statistics cluster) in the order of a decoder layer,
then the persistent chain kernel on the same four linears."""
import math, sys, torch
sys.path.insert(0, ".")
from crab_b200 import ops
ops.init(0)
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, D, F, nq, nk = 32, 4096, 11008, 4096, 4096


def rnd(*shape, scale=1.0):
    return (scale * torch.randn(*shape, device=dev)).to(torch.bfloat16)


def make(N, K, kext, linears, gamma, swiglu=False):
    w = torch.zeros((N, K + kext), device=dev, dtype=torch.bfloat16)
    w[:, :K] = rnd(N, K, scale=1 / math.sqrt(K))
    w[:, K:K + 24 * linears] = rnd(N, 24 * linears, scale=0.05)
    return ops.pack_skinny_weight(w, k=K + kext, swiglu=swiglu), ops.pack_chain_stats(rnd(11 * linears, K, scale=1 / math.sqrt(K)),
                                                                                    (1 + 0.1 * torch.randn(K, device=dev)) if gamma else None)


wq, sq = make(nq + 2 * nk, D, 96, 3, True)
wo, so = make(D, nq, 32, 1, False)
wgu, sgu = make(2 * F, D, 64, 2, True, swiglu=True)
wd, sd_ = make(D, F, 32, 1, False)
at = torch.zeros((B, nq + 32), device=dev, dtype=torch.bfloat16); at[:, :nq] = rnd(B, nq)
x = rnd(B, D); hh = rnd(B, F); qkv = torch.empty((B, nq + 2 * nk), device=dev, dtype=torch.bfloat16)
z = {k: torch.zeros((32, 128), device=dev, dtype=torch.bfloat16) for k in "o gu d q".split()}
rs = {k: torch.zeros(32, device=dev, dtype=torch.float32) for k in "gu q".split()}
fl = lambda: torch.zeros(64, dtype=torch.int32, device=dev)
# fused linears: qkv (norm + 3 LoRA linears, split 2), o (z from the attention kernel, split 4), gate/up (norm + 2, SwiGLU, no split),
# down (1 LoRA linear, split 8)
ops.gemm_skinny(x, wq, out=qkv, z=z["q"], kext=96, stats=sq, stats_linears=3, norm=True, eps=1e-6, lora_scale=2.0, rstd=rs["q"], flags=fl())
ops.gemm_skinny(at, wo, residual=x, out=x, z=at[:, nq:], kext=32, splits=4)
ops.gemm_skinny(x, wgu, act=ops.ACT_SWIGLU, out=hh, z=z["gu"], kext=64, stats=sgu, stats_linears=2, norm=True, eps=1e-6, lora_scale=2.0, rstd=rs["gu"], flags=fl())
ops.gemm_skinny(hh, wd, residual=x, out=x, z=z["d"], kext=32, stats=sd_, stats_linears=1, lora_scale=2.0, flags=fl())
# the persistent chain on the same four linears
cnt = torch.zeros(288, dtype=torch.int32, device=dev)
phases = [ops.ChainPhase(at, wo, x, k=nq, z=at[:, nq:], kext=32, residual=x),
          ops.ChainPhase(x, wgu, hh, k=D, z=z["gu"], kext=64, stats=sgu, stats_linears=2, norm=True, eps=1e-6, lora_scale=2.0, rstd=rs["gu"], act=ops.ACT_SWIGLU),
          ops.ChainPhase(hh, wd, x, k=F, z=z["d"], kext=32, stats=sd_, stats_linears=1, lora_scale=2.0, residual=x),
          ops.ChainPhase(x, wq, qkv, k=D, z=z["q"], kext=96, stats=sq, stats_linears=3, norm=True, eps=1e-6, lora_scale=2.0, rstd=rs["q"])]
ops.decode_chain(phases, B, cnt, 4)
torch.cuda.synchronize()
print("ok")
