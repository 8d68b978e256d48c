"""The decode step's weight-streaming linears in their final round-2 form, one launch each at the bench's shapes (bs 32, LLaMA-7B
dims) and with the engine's launch configuration (K-split 2 / 4 / 1 / 4, statistics item shared by several clusters, flag slots in a
ring, st.async exchange, rolled finish), for
`ncu --set full --clock-control none --import-source on -k regex:gemm_skinny -o gpurun_out/<tag> python tools/profile_decode_r4.py`."""
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
ring = torch.zeros((5, 32), dtype=torch.int32, device=dev)
scratch = torch.zeros(8 * 36 * 32, dtype=torch.float32, device=dev)
fk = lambda k: dict(flags=ring[k], flags_clear=ring[(k - 1) % 5], stats_scratch=scratch)
wh, _ = make(32017, D, 0, 0, True)
logits = torch.zeros((B, 32024), device=dev, dtype=torch.float32)
for _ in range(2):   # second pass = warm instruction / constant caches, as in the step
    ops.gemm_skinny(x, wq, out=qkv, z=z["q"], kext=96, stats=sq, stats_linears=3, norm=True, eps=1e-6, lora_scale=2.0, rstd=rs["q"], **fk(0))
    ops.gemm_skinny(at, wo, residual=x, out=x, z=at[:, nq:], kext=32, splits=4)
    ops.gemm_skinny(x, wgu, act=ops.ACT_SWIGLU, out=hh, z=z["gu"], kext=64, stats=sgu, stats_linears=2, norm=True, eps=1e-6, lora_scale=2.0, rstd=rs["gu"], **fk(1))
    ops.gemm_skinny(hh, wd, residual=x, out=x, z=z["d"], kext=32, stats=sd_, stats_linears=1, lora_scale=2.0, splits=4, **fk(2))
    ops.gemm_skinny(x, wh, out=logits, n=32017, norm=True, eps=1e-6, rstd=rs["q"], **fk(3))
    ops.gemm_skinny(x, wh, out=logits, n=32017, norm=True, eps=1e-6, rstd=rs["q"], **fk(4))
torch.cuda.synchronize()
print("ok")
