import math, sys, torch
sys.path.insert(0, ".")
from crab_b200 import ops
ops.init(0)
dev = torch.device("cuda:0")
def bench(name, B, H, KVH, S, hd, causal):
    q = torch.randn(B * S, (H + 2 * KVH) * hd, device=dev, dtype=torch.bfloat16)
    kc = torch.randn(B, KVH, S, hd, device=dev, dtype=torch.bfloat16)
    vc = torch.randn(B, KVH, S, hd, device=dev, dtype=torch.bfloat16)
    o = torch.empty(B * S, H * hd, device=dev, dtype=torch.bfloat16)
    ld = (H + 2 * KVH) * hd
    def run():
        ops.flash_attn(q, kc, vc, o, B=B, H=H, KVH=KVH, Sq=S, Sk=S, head_dim=hd, q_strides=(S * ld, ld, hd),
                       k_strides=(KVH * S * hd, hd, S * hd), v_strides=(KVH * S * hd, hd, S * hd), o_strides=(S * H * hd, H * hd, hd),
                       scale=hd ** -0.5, causal=causal)
    for _ in range(3): run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[2]
    fl = 4.0 * B * H * S * S * hd * (0.5 if causal else 1.0)
    print(f"{name:28s} {t*1e3:9.1f} us   {fl / t / 1e9:7.1f} TFLOP/s (algorithmic{', causal half' if causal else ''})", flush=True)
bench("llama prefill bs32 S=1086", 32, 32, 32, 1086, 128, True)
bench("qwen prefill bs32 S=1086 GQA", 32, 28, 4, 1086, 128, True)
bench("clip 256 frames N=257", 256, 16, 16, 257, 64, False)
bench("beats 320 segs N=48", 320, 12, 12, 48, 64, False)
# is the hd-128 tcgen05 kernel waiting for its K / V tiles?  Same per-CTA work, cache-resident vs HBM-resident operands
bench("llama prefill bs2 (L2-resident)", 2, 32, 32, 1086, 128, True)
bench("llama prefill bs8", 8, 32, 32, 1086, 128, True)
bench("non-causal bs32 S=1086", 32, 32, 32, 1086, 128, False)
bench("qwen GQA bs8 (L2-resident)", 8, 28, 4, 1086, 128, True)
