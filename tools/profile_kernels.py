"""Launch each hot kernel once at the bench's shapes (bs32, S=1086, LLaMA-7B dims) so that
`ncu --set full --clock-control none --import-source on -o gpurun_out/<tag>_kernels python tools/profile_kernels.py`
captures exactly one representative launch per kernel."""
import math, sys, torch
sys.path.insert(0, ".")
from crab_b200 import ops
ops.init(0)
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, S, D, F, H, hd = 32, 1086, 4096, 11008, 32, 128
M = B * S
# 1. prefill qkv GEMM (K-extended), the largest share of the step
a = torch.randn(M, D + 96, device=dev, dtype=torch.bfloat16)
w = torch.randn(3 * D, D + 96, device=dev, dtype=torch.bfloat16) * 0.02
qkv = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
ops.gemm(a, w, out=qkv)
# 2. prefill gate/up GEMM with SwiGLU epilogue
wgu = torch.randn(2 * F, D + 64, device=dev, dtype=torch.bfloat16) * 0.02
h = torch.empty(M, F, device=dev, dtype=torch.bfloat16)
ops.gemm(a, wgu, act=ops.ACT_SWIGLU, out=h, k=D + 64)
# 3. causal flash attention prefill
kc = torch.randn(B, H, 1216, hd, device=dev, dtype=torch.bfloat16)
vc = torch.randn(B, H, 1216, hd, device=dev, dtype=torch.bfloat16)
o = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
ops.flash_attn(qkv, kc, vc, o, B=B, H=H, KVH=H, Sq=S, Sk=S, head_dim=hd, q_strides=(S * 3 * D, 3 * D, hd),
               k_strides=(H * 1216 * hd, hd, 1216 * hd), v_strides=(H * 1216 * hd, hd, 1216 * hd), o_strides=(S * D, D, hd),
               scale=hd ** -0.5, causal=True)
# 4. decode: skinny GEMMs (gate/up no split, o-proj 8-way cluster split), decode attention, row norm + LoRA-z
x = torch.randn(B, D + 96, device=dev, dtype=torch.bfloat16)
wgu_p = ops.pack_skinny_weight(wgu, k=D + 64, swiglu=True)
hd_ = torch.empty(B, F, device=dev, dtype=torch.bfloat16)
ops.gemm_skinny(x, wgu_p, act=ops.ACT_SWIGLU, out=hd_)
wo = torch.randn(D, D + 32, device=dev, dtype=torch.bfloat16) * 0.02
wo_p = ops.pack_skinny_weight(wo)
xo = torch.randn(B, D, device=dev, dtype=torch.bfloat16)
ops.gemm_skinny(x, wo_p, residual=xo, out=xo)
ln = torch.tensor([1150], dtype=torch.int32, device=dev)
qd = torch.randn(B, 3 * D, device=dev, dtype=torch.bfloat16)
od = torch.empty(B, D, device=dev, dtype=torch.bfloat16)
ops.attn_decode(qd, kc, vc, od, B=B, H=H, KVH=H, head_dim=hd, scale=hd ** -0.5, len_dev=ln)
# 4b. the decode step's fused kernel: RoPE + KV append + attention + o_proj LoRA pre-pass
rope = ops.rope_table(1216, hd, 10000.0, dev)
pd = torch.tensor([1149], dtype=torch.int32, device=dev)
ra_o = torch.randn(11, D + 32, device=dev, dtype=torch.bfloat16) * 0.02
at = torch.zeros(B, D + 32, device=dev, dtype=torch.bfloat16)
ops.attn_decode_fused(qd, rope, kc, vc, at[:, :D], B=B, H=H, KVH=H, head_dim=hd, scale=hd ** -0.5, past_dev=pd,
                      ra=ra_o[:, :D], z=at[:, D:], lora_scale=2.0, lora_ws=torch.empty(B * H * 11, device=dev),
                      lora_counters=torch.zeros(B, dtype=torch.int32, device=dev))
gamma = torch.ones(D, device=dev)
ra = torch.randn(33, D, device=dev, dtype=torch.bfloat16) * 0.02
ops.row_norm_loraz(xo, gamma=gamma, eps=1e-6, y=x[:, :D], ra=ra, groups=3, z=x[:, D:], scale=2.0)
# 5. front-end: Kaldi fbank for 32 samples x 10 one-second segments, fused uint8 normalise + patchify
from crab_b200.dataset import audio_processor as A
wave = 0.2 * torch.randn(320, 16000, device=dev)
A.preprocess(wave)
frames = torch.randint(0, 256, (256, 224, 224, 3), device=dev, dtype=torch.uint8)
ops.patchify_u8(frames, 14, 592, (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711))
torch.cuda.synchronize()
print("ok")
