"""One launch each of the hd-64 encoder attention shapes (CLIP 256 x 16 x 257, BEATs 320 x 12 x 48 with the gated bias, Q-Former cross
256 x 12 x 32 -> 256 keys) for `ncu --set full -k regex:flash_attn_kernel`: the TMA-staged mma.sync kernel of round 2."""
import math, sys, torch
sys.path.insert(0, ".")
from crab_b200 import ops
ops.init(0)
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
def rnd(*s):
    return torch.randn(*s, generator=g, device=dev).to(torch.bfloat16)
for _ in range(2):
    # CLIP
    n, T, H, hd = 256, 257, 16, 64; D = H * hd
    qkv = rnd(n * T, 3 * D); o = torch.empty(n * T, D, device=dev, dtype=torch.bfloat16)
    ops.flash_attn(qkv, qkv[:, D:], qkv[:, 2 * D:], o, B=n, H=H, KVH=H, Sq=T, Sk=T, head_dim=hd, q_strides=(T * 3 * D, 3 * D, hd),
                   k_strides=(T * 3 * D, 3 * D, hd), v_strides=(T * 3 * D, 3 * D, hd), o_strides=(T * D, D, hd), scale=hd ** -0.5)
    # BEATs with the gated relative-position bias
    n, T, H = 320, 48, 12; D = H * hd
    qkv = rnd(n * T, 3 * D); o = torch.empty(n * T, D, device=dev, dtype=torch.bfloat16)
    gate = 1 + torch.rand(n, H, T, generator=g, device=dev); table = torch.randn(H, T, T, generator=g, device=dev)
    ops.flash_attn(qkv, qkv[:, D:], qkv[:, 2 * D:], o, B=n, H=H, KVH=H, Sq=T, Sk=T, head_dim=hd, q_strides=(T * 3 * D, 3 * D, hd),
                   k_strides=(T * 3 * D, 3 * D, hd), v_strides=(T * 3 * D, 3 * D, hd), o_strides=(T * D, D, hd), scale=hd ** -0.5,
                   gate=gate, bias_table=table)
    # Q-Former cross attention: 32 queries over 256 patch tokens (row 0 = CLS skipped)
    n, nq, Tk, H = 256, 32, 257, 12; D = H * hd
    qc = rnd(n * nq, D); kv = rnd(n * Tk, 2 * D); o = torch.empty(n * nq, D, device=dev, dtype=torch.bfloat16)
    kv0 = kv[1:]
    ops.flash_attn(qc, kv0, kv0[:, D:], o, B=n, H=H, KVH=H, Sq=nq, Sk=Tk - 1, head_dim=hd, q_strides=(nq * D, D, hd),
                   k_strides=(Tk * 2 * D, 2 * D, hd), v_strides=(Tk * 2 * D, 2 * D, hd), o_strides=(nq * D, D, hd), scale=hd ** -0.5)
torch.cuda.synchronize()
print("ok")
