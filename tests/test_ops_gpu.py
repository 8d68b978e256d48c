"""Kernel-level parity for the non-GEMM ops: each CUDA kernel vs a plain fp32 torch evaluation of the same op on the
same bf16-rounded inputs (tolerance = a few bf16 ulps of the output scale)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _g(seed):
    return torch.Generator(device="cpu").manual_seed(seed)


def _close(out, ref, tol):
    err = (out.float() - ref.float()).abs().max().item()
    scale = ref.float().abs().max().item() + 1e-6
    assert err / scale < tol, f"rel {err / scale:.3e} abs {err:.3e}"


@pytest.mark.parametrize("rows,cols", [(5, 512), (300, 768), (1000, 1024), (77, 4096), (3, 11008)])
def test_layernorm_rmsnorm(cuda_dev, rows, cols):
    from crab_b200 import ops

    g = _g(rows + cols)
    x = (torch.randn(rows, cols, generator=g) * 2 + 0.3).to(torch.bfloat16).to(cuda_dev)
    w = (1 + 0.1 * torch.randn(cols, generator=g)).to(cuda_dev)
    b = (0.1 * torch.randn(cols, generator=g)).to(cuda_dev)
    y = ops.layernorm(x, w, b, 1e-5)
    _close(y, torch.nn.functional.layer_norm(x.float(), (cols,), w, b, 1e-5), 1e-2)
    y = ops.rmsnorm(x, w, 1e-6)
    xf = x.float()
    ref = w * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)).to(torch.bfloat16).float()
    _close(y, ref, 1e-2)
    # strided output (K-extended activation buffers)
    buf = torch.zeros(rows, cols + 64, dtype=torch.bfloat16, device=cuda_dev)
    ops.rmsnorm(x, w, 1e-6, out=buf[:, :cols])
    assert torch.equal(buf[:, :cols], y) and buf[:, cols:].abs().max().item() == 0


def test_rope_table_and_append(cuda_dev):
    from crab_b200 import ops
    from oracle import crab_oracle as O

    B, S, H, KV, hd, ctx = 2, 37, 4, 2, 128, 64
    tab = ops.rope_table(ctx, hd, 10000.0, cuda_dev)
    cos, sin = O.rope_cos_sin(torch.arange(ctx), hd, 10000.0)
    assert (tab[:, : hd // 2].cpu() - cos[:, : hd // 2]).abs().max() < 2e-6
    assert (tab[:, hd // 2:].cpu() - sin[:, : hd // 2]).abs().max() < 2e-6
    g = _g(1)
    qkv = torch.randn(B * S, (H + 2 * KV) * hd, generator=g).to(torch.bfloat16)
    kc = torch.zeros(B, KV, ctx, hd, dtype=torch.bfloat16, device=cuda_dev)
    vc = torch.zeros_like(kc)
    past = 5
    qkv_d = qkv.to(cuda_dev).clone()
    past_dev = torch.tensor([past], dtype=torch.int32, device=cuda_dev)
    ops.rope_kv_append(qkv_d, tab, kc, vc, B, S, H, KV, hd, past_dev=past_dev)
    x = qkv.float().view(B, S, H + 2 * KV, hd).transpose(1, 2)
    c, s_ = cos[past:past + S], sin[past:past + S]
    q_ref = O.apply_rope(x[:, :H], c, s_)
    k_ref = O.apply_rope(x[:, H:H + KV], c, s_)
    got = qkv_d.float().cpu().view(B, S, H + 2 * KV, hd).transpose(1, 2)
    _close(got[:, :H], q_ref, 1e-2)
    _close(kc[:, :, past:past + S].cpu(), k_ref, 1e-2)
    assert torch.equal(vc[:, :, past:past + S].cpu(), qkv.view(B, S, H + 2 * KV, hd).transpose(1, 2)[:, H + KV:])
    assert kc[:, :, :past].abs().max().item() == 0 and kc[:, :, past + S:].abs().max().item() == 0


def _attn_ref(q, k, v, scale, causal=False, bias=None):
    # q (B,H,Sq,d) k,v (B,KVH,Sk,d)
    B, H, Sq, d = q.shape
    KVH, Sk = k.shape[1], k.shape[2]
    k = k.repeat_interleave(H // KVH, dim=1)
    v = v.repeat_interleave(H // KVH, dim=1)
    s = torch.matmul(q.float(), k.float().transpose(-1, -2)) * scale
    if bias is not None:
        s = s + bias
    if causal:
        m = torch.full((Sq, Sk), float("-inf"), device=q.device).triu(diagonal=Sk - Sq + 1)
        s = s + m
    return torch.matmul(torch.softmax(s, dim=-1), v.float())


@pytest.mark.parametrize("B,H,KVH,Sq,Sk,hd,causal", [
    (2, 16, 16, 257, 257, 64, False),    # CLIP
    (3, 12, 12, 32, 32, 64, False),      # Q-Former self
    (3, 12, 12, 32, 256, 64, False),     # Q-Former cross (video)
    (3, 12, 12, 32, 48, 64, False),      # Q-Former cross (audio)
    (2, 4, 4, 200, 200, 128, True),      # LLaMA prefill
    (2, 8, 2, 150, 150, 128, True),      # Qwen2 GQA prefill
    (1, 2, 2, 70, 133, 128, True),       # chunked prefill (Sk > Sq)
    # head_dim 128 with >= 128 queries runs on the tcgen05 / TMEM kernel (flash_tcgen05.cu)
    (1, 2, 2, 128, 128, 128, True),      # exactly one q tile, two key tiles
    (2, 4, 4, 1086, 1086, 128, True),    # the benchmark's sequence length: 9 q tiles, ragged last tile
    (1, 4, 2, 300, 555, 128, True),      # chunked prefill with GQA (Sk > Sq)
    (2, 2, 2, 257, 257, 128, False),     # non-causal, key tail masked by length only
    (1, 2, 1, 640, 640, 128, False),
])
def test_flash_attn(cuda_dev, B, H, KVH, Sq, Sk, hd, causal):
    from crab_b200 import ops

    g = _g(B * 1000 + Sq + Sk)
    # packed layouts as the engine uses them: q in a [B*Sq, H*hd] buffer, k/v in cache layout [B,KVH,Sk,hd]
    q = torch.randn(B, Sq, H, hd, generator=g).to(torch.bfloat16).to(cuda_dev)
    k = torch.randn(B, KVH, Sk, hd, generator=g).to(torch.bfloat16).to(cuda_dev)
    v = torch.randn(B, KVH, Sk, hd, generator=g).to(torch.bfloat16).to(cuda_dev)
    o = torch.zeros(B, Sq, H, hd, dtype=torch.bfloat16, device=cuda_dev)
    ops.flash_attn(q, k, v, o, B=B, H=H, KVH=KVH, Sq=Sq, Sk=Sk, head_dim=hd,
                   q_strides=(Sq * H * hd, H * hd, hd), k_strides=(KVH * Sk * hd, hd, Sk * hd),
                   v_strides=(KVH * Sk * hd, hd, Sk * hd), o_strides=(Sq * H * hd, H * hd, hd),
                   scale=1 / math.sqrt(hd), causal=causal)
    torch.cuda.synchronize()
    ref = _attn_ref(q.transpose(1, 2), k, v, 1 / math.sqrt(hd), causal).transpose(1, 2)
    _close(o, ref, 2e-2)


@pytest.mark.parametrize("Sq,Sk,causal", [(300, 300, True), (257, 257, False), (130, 333, True)])
def test_flash_attn_tcgen05_ignores_garbage_past_sk(cuda_dev, Sq, Sk, causal):
    """The tcgen05 kernel fetches whole 64-key tiles: cache rows past Sk (here NaN and Inf bit patterns, as in a torch.empty cache)
    must not reach the output — their scores are masked and the kernel zeroes the V rows past Sk in shared memory."""
    from crab_b200 import ops

    B, H, hd, ctx = 2, 4, 128, 384
    g = _g(Sq * 7 + Sk)
    q = torch.randn(B, Sq, H, hd, generator=g).to(torch.bfloat16).to(cuda_dev)
    kc = torch.randn(B, H, ctx, hd, generator=g).to(torch.bfloat16).to(cuda_dev)
    vc = torch.randn(B, H, ctx, hd, generator=g).to(torch.bfloat16).to(cuda_dev)
    kc[:, :, Sk:] = float("nan")
    vc[:, :, Sk:] = float("nan")
    vc[:, :, Sk::2] = float("inf")
    o = torch.zeros(B, Sq, H, hd, dtype=torch.bfloat16, device=cuda_dev)
    ops.flash_attn(q, kc, vc, o, B=B, H=H, KVH=H, Sq=Sq, Sk=Sk, head_dim=hd,
                   q_strides=(Sq * H * hd, H * hd, hd), k_strides=(H * ctx * hd, hd, ctx * hd),
                   v_strides=(H * ctx * hd, hd, ctx * hd), o_strides=(Sq * H * hd, H * hd, hd),
                   scale=1 / math.sqrt(hd), causal=causal)
    torch.cuda.synchronize()
    assert torch.isfinite(o.float()).all()
    ref = _attn_ref(q.transpose(1, 2), kc[:, :, :Sk], vc[:, :, :Sk], 1 / math.sqrt(hd), causal).transpose(1, 2)
    _close(o, ref, 2e-2)


def test_flash_attn_beats_bias(cuda_dev):
    from crab_b200 import ops

    B, H, T, hd = 5, 12, 48, 64
    g = _g(77)
    qkv = torch.randn(B * T, 3 * H * hd, generator=g).to(torch.bfloat16).to(cuda_dev)
    gate = (1 + torch.rand(B, H, T, generator=g)).to(cuda_dev)
    table = torch.randn(H, T, T, generator=g).to(cuda_dev)
    o = torch.zeros(B * T, H * hd, dtype=torch.bfloat16, device=cuda_dev)
    D = H * hd
    ops.flash_attn(qkv, qkv[:, D:], qkv[:, 2 * D:], o, B=B, H=H, KVH=H, Sq=T, Sk=T, head_dim=hd,
                   q_strides=(T * 3 * D, 3 * D, hd), k_strides=(T * 3 * D, 3 * D, hd), v_strides=(T * 3 * D, 3 * D, hd),
                   o_strides=(T * D, D, hd), scale=hd ** -0.5, gate=gate, bias_table=table)
    torch.cuda.synchronize()
    x = qkv.view(B, T, 3, H, hd)
    bias = gate.unsqueeze(-1) * table.unsqueeze(0)
    ref = _attn_ref(x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2), hd ** -0.5,
                    bias=bias).transpose(1, 2).reshape(B * T, D)
    _close(o, ref, 2e-2)


@pytest.mark.parametrize("B,H,KVH,hd,length,nsplit", [(4, 8, 8, 128, 300, 1), (2, 8, 8, 128, 1213, 4),
                                                       (3, 28, 4, 128, 517, 1), (1, 28, 4, 128, 1000, 8),
                                                       (2, 4, 4, 128, 1, 1), (2, 4, 4, 128, 3, 2)])
def test_attn_decode(cuda_dev, B, H, KVH, hd, length, nsplit):
    from crab_b200 import ops

    ctx = 1280
    g = _g(B + H + length)
    q = torch.randn(B, (H + 2 * KVH) * hd, generator=g).to(torch.bfloat16).to(cuda_dev)
    kc = torch.randn(B, KVH, ctx, hd, generator=g).to(torch.bfloat16).to(cuda_dev)
    vc = torch.randn(B, KVH, ctx, hd, generator=g).to(torch.bfloat16).to(cuda_dev)
    o = torch.zeros(B, H * hd, dtype=torch.bfloat16, device=cuda_dev)
    ld = torch.tensor([length], dtype=torch.int32, device=cuda_dev)
    ops.attn_decode(q, kc, vc, o, B=B, H=H, KVH=KVH, head_dim=hd, scale=hd ** -0.5, len_dev=ld, nsplit=nsplit)
    torch.cuda.synchronize()
    ref = _attn_ref(q[:, : H * hd].view(B, 1, H, hd).transpose(1, 2), kc[:, :, :length], vc[:, :, :length], hd ** -0.5)
    _close(o.view(B, H, hd), ref[:, :, 0], 2e-2)


@pytest.mark.parametrize("B,H,KVH,hd,past,nsplit,lora", [(4, 8, 8, 128, 300, 1, True), (32, 32, 32, 128, 1100, 1, True),
                                                          (3, 28, 4, 128, 517, 1, True), (2, 8, 8, 128, 1213, 4, False),
                                                          (2, 4, 4, 128, 0, 1, True), (2, 4, 4, 64, 77, 1, False),
                                                          (1, 28, 4, 128, 1000, 8, False), (1, 32, 32, 128, 150, 10, True),
                                                          (3, 8, 8, 128, 1213, 4, True), (2, 28, 4, 128, 517, 3, True)])
def test_attn_decode_fused_equals_the_three_kernel_sequence(cuda_dev, B, H, KVH, hd, past, nsplit, lora):
    """RoPE + KV append + decode attention (+ o_proj LoRA pre-pass) in one launch vs rope_kv_append -> attn_decode ->
    row_norm_loraz: caches bit-identical, attention output and z within fp32 summation-order noise; repeated launches
    (the arrival counters must come back to zero) give identical results."""
    from crab_b200 import ops

    ctx = 1280
    g = _g(B * 7 + H + past)
    nq = H * hd
    qkv = torch.randn(B, (H + 2 * KVH) * hd, generator=g).to(torch.bfloat16).to(cuda_dev)
    kc = torch.randn(B, KVH, ctx, hd, generator=g).to(torch.bfloat16).to(cuda_dev)
    vc = torch.randn(B, KVH, ctx, hd, generator=g).to(torch.bfloat16).to(cuda_dev)
    ra = (torch.randn(11, nq, generator=g) / math.sqrt(nq)).to(torch.bfloat16).to(cuda_dev)
    rope = ops.rope_table(ctx, hd, 10000.0, cuda_dev)
    pd = torch.tensor([past], dtype=torch.int32, device=cuda_dev)
    ld = torch.tensor([past + 1], dtype=torch.int32, device=cuda_dev)
    # reference sequence
    q1, k1, v1 = qkv.clone(), kc.clone(), vc.clone()
    o1 = torch.zeros(B, nq + 32, dtype=torch.bfloat16, device=cuda_dev)
    ops.rope_kv_append(q1, rope, k1, v1, B, 1, H, KVH, hd, past=0, past_dev=pd)
    ops.attn_decode(q1, k1, v1, o1[:, :nq], B=B, H=H, KVH=KVH, head_dim=hd, scale=hd ** -0.5, len_dev=ld, nsplit=nsplit)
    if lora:
        ops.row_norm_loraz(o1[:, :nq], ra=ra, groups=1, z=o1[:, nq:], scale=2.0)
    # fused
    k2, v2 = kc.clone(), vc.clone()
    o2 = torch.zeros(B, nq + 32, dtype=torch.bfloat16, device=cuda_dev)
    ws = torch.empty(B * (KVH if nsplit == 1 else H) * 11, dtype=torch.float32, device=cuda_dev)
    cnt = torch.zeros(B, dtype=torch.int32, device=cuda_dev)
    for rep in range(2):
        if rep == 1:
            o2.zero_()
        ops.attn_decode_fused(qkv, rope, k2, v2, o2[:, :nq], B=B, H=H, KVH=KVH, head_dim=hd, scale=hd ** -0.5, past_dev=pd,
                              nsplit=nsplit, ra=ra if lora else None, z=o2[:, nq:] if lora else None, lora_scale=2.0,
                              lora_ws=ws if lora else None, lora_counters=cnt if lora else None)
        torch.cuda.synchronize()
        assert torch.equal(k2, k1) and torch.equal(v2, v1)
        _close(o2[:, :nq], o1[:, :nq], 1e-2)  # the new key joins a different partial softmax state: same math, other order
        if rep == 0:
            first = o2.clone()
        assert torch.equal(o2, first)
        assert int(cnt.abs().sum()) == 0
        if lora:
            _close(o2[:, nq:nq + 24], o1[:, nq:nq + 24], 2e-2)
            assert o2[:, nq + 24:].abs().max().item() == 0


@pytest.mark.parametrize("B,H,KVH,past,nsplit,lora", [(32, 28, 4, 1150, 3, True), (16, 32, 8, 77, 2, True), (20, 28, 4, 63, 4, False),
                                                      (16, 28, 4, 0, 3, True), (32, 28, 4, 700, 1, False)])
def test_gqa_decode_split_kv_on_tensor_cores(cuda_dev, B, H, KVH, past, nsplit, lora):
    """crab_attn_decode_fused(gqa_tensor_cores): RoPE + append, the G heads of a kv group as the rows of a flash problem split over
    nsplit key ranges, combine (+ o_proj LoRA pre-pass) — against rope_kv_append -> scalar attn_decode -> row_norm_loraz."""
    from crab_b200 import ops

    hd, ctx = 128, 1280
    g = _g(B * 11 + H + past + nsplit)
    nq = H * hd
    qkv = torch.randn(B, (H + 2 * KVH) * hd, generator=g).to(torch.bfloat16).to(cuda_dev)
    kc = torch.randn(B, KVH, ctx, hd, generator=g).to(torch.bfloat16).to(cuda_dev)
    vc = torch.randn(B, KVH, ctx, hd, generator=g).to(torch.bfloat16).to(cuda_dev)
    ra = (torch.randn(11, nq, generator=g) / math.sqrt(nq)).to(torch.bfloat16).to(cuda_dev)
    rope = ops.rope_table(ctx, hd, 1e6, cuda_dev)
    pd = torch.tensor([past], dtype=torch.int32, device=cuda_dev)
    ld = torch.tensor([past + 1], dtype=torch.int32, device=cuda_dev)
    q1, k1, v1 = qkv.clone(), kc.clone(), vc.clone()
    o1 = torch.zeros(B, nq + 32, dtype=torch.bfloat16, device=cuda_dev)
    ops.rope_kv_append(q1, rope, k1, v1, B, 1, H, KVH, hd, past=0, past_dev=pd)
    ops.attn_decode(q1, k1, v1, o1[:, :nq], B=B, H=H, KVH=KVH, head_dim=hd, scale=hd ** -0.5, len_dev=ld)
    if lora:
        ops.row_norm_loraz(o1[:, :nq], ra=ra, groups=1, z=o1[:, nq:], scale=2.0)
    q2, k2, v2 = qkv.clone(), kc.clone(), vc.clone()
    o2 = torch.zeros(B, nq + 32, dtype=torch.bfloat16, device=cuda_dev)
    ws = torch.empty(B * H * 11, dtype=torch.float32, device=cuda_dev)
    cnt = torch.zeros(B, dtype=torch.int32, device=cuda_dev)
    ops.attn_decode_fused(q2, rope, k2, v2, o2[:, :nq], B=B, H=H, KVH=KVH, head_dim=hd, scale=hd ** -0.5, past_dev=pd, nsplit=nsplit,
                          gqa_tc=True, ra=ra if lora else None, z=o2[:, nq:] if lora else None, lora_scale=2.0,
                          lora_ws=ws if lora else None, lora_counters=cnt if lora else None)
    torch.cuda.synchronize()
    assert torch.equal(k2, k1) and torch.equal(v2, v1) and torch.equal(q2, q1)     # RoPE + append in place, as the separate launch
    _close(o2[:, :nq], o1[:, :nq], 1.5e-2)   # P is rounded to bf16 before PV on the tensor-core path
    assert int(cnt.abs().sum()) == 0
    if lora:
        _close(o2[:, nq:nq + 24], o1[:, nq:nq + 24], 3e-2)


@pytest.mark.parametrize("B,H,KVH,length", [(32, 28, 4, 1150), (16, 32, 8, 77), (20, 8, 4, 64), (16, 28, 4, 1)])
def test_gqa_decode_on_tensor_cores(cuda_dev, B, H, KVH, length):
    """Grouped-query decode as a flash-attention problem (the G heads of a kv group = the query rows; key count read from
    device memory, as inside the decode graph) against the scalar decode kernel and the fp32 reference."""
    from crab_b200 import ops

    hd, ctx, G = 128, 1280, H // KVH
    g = _g(B + H + length)
    ld = (H + 2 * KVH) * hd
    q = torch.randn(B, ld, generator=g).to(torch.bfloat16).to(cuda_dev)
    kc = torch.randn(B, KVH, ctx, hd, generator=g).to(torch.bfloat16).to(cuda_dev)
    vc = torch.randn(B, KVH, ctx, hd, generator=g).to(torch.bfloat16).to(cuda_dev)
    ldv = torch.tensor([length], dtype=torch.int32, device=cuda_dev)
    o1 = torch.zeros(B, H * hd + 32, dtype=torch.bfloat16, device=cuda_dev)
    o2 = torch.zeros_like(o1)
    ops.attn_decode(q, kc, vc, o1[:, : H * hd], B=B, H=H, KVH=KVH, head_dim=hd, scale=hd ** -0.5, len_dev=ldv)
    ops.flash_attn(q, kc, vc, o2, B=B, H=KVH, KVH=KVH, Sq=G, Sk=ctx, head_dim=hd, q_strides=(ld, hd, G * hd),
                   k_strides=(KVH * ctx * hd, hd, ctx * hd), v_strides=(KVH * ctx * hd, hd, ctx * hd),
                   o_strides=(H * hd + 32, hd, G * hd), scale=hd ** -0.5, sk_dev=ldv)
    torch.cuda.synchronize()
    ref = _attn_ref(q[:, : H * hd].view(B, 1, H, hd).transpose(1, 2), kc[:, :, :length], vc[:, :, :length], hd ** -0.5)
    _close(o2[:, : H * hd].view(B, H, hd), ref[:, :, 0], 2e-2)
    _close(o2[:, : H * hd], o1[:, : H * hd], 2e-2)
    assert o2[:, H * hd:].abs().max().item() == 0


def test_gather_cast_patchify_argmax(cuda_dev):
    from crab_b200 import ops

    g = _g(4)
    table = torch.randn(50, 256, generator=g).to(torch.bfloat16).to(cuda_dev)
    idx = torch.randint(0, 50, (20,), generator=g).to(cuda_dev)
    dst = torch.zeros(40, 256, dtype=torch.bfloat16, device=cuda_dev)
    dst_rows = (torch.arange(20) * 2).to(cuda_dev)
    ops.gather_rows(table, dst, 20, 256, src_rows=idx, dst_rows=dst_rows)
    assert torch.equal(dst[::2], table[idx]) and dst[1::2].abs().max().item() == 0
    x = torch.randn(3, 1001, generator=g).to(cuda_dev)
    assert torch.equal(ops.cast_bf16(x), x.to(torch.bfloat16))
    img = torch.randn(2, 3, 56, 56, generator=g).to(cuda_dev)
    p = ops.patchify(img, 14, 592)
    ref = torch.nn.functional.unfold(img, kernel_size=14, stride=14).transpose(1, 2).reshape(2 * 16, 588)
    assert torch.equal(p[:, :588], ref.to(torch.bfloat16)) and p[:, 588:].abs().max().item() == 0
    fb = torch.randn(3, 1, 98, 128, generator=g).to(cuda_dev)
    p = ops.patchify(fb, 16, 256)
    ref = torch.nn.functional.unfold(fb[:, :, :96], kernel_size=16, stride=16).transpose(1, 2).reshape(3 * 48, 256)
    assert torch.equal(p, ref.to(torch.bfloat16))
    logits = torch.randn(7, 32024, generator=g).to(cuda_dev)
    logits[:, 32017:] = 100.0  # padding columns must be ignored
    assert torch.equal(ops.argmax(logits, 32017), logits[:, :32017].argmax(-1))
    big = torch.randn(32, 152088, generator=_g(4)).to(cuda_dev)   # Qwen2 vocabulary: the 8-CTA cluster kernel
    big[3, 777] = big[3, 150001] = 50.0                            # tie: the first index wins
    big[5, 152080] = 60.0                                          # winner in the last, ragged slice
    assert torch.equal(ops.argmax(big, 152081), big[:, :152081].argmax(-1))
    assert int(ops.argmax(big, 152081)[3]) == 777


def test_clip_embed_and_beats_helpers(cuda_dev):
    from crab_b200 import ops

    g = _g(6)
    n, T, D = 3, 17, 1024
    pe = torch.randn(n * (T - 1), D, generator=g).to(torch.bfloat16).to(cuda_dev)
    cls = torch.randn(D, generator=g).to(cuda_dev)
    pos = torch.randn(T, D, generator=g).to(cuda_dev)
    w = (1 + 0.1 * torch.randn(D, generator=g)).to(cuda_dev)
    b = (0.1 * torch.randn(D, generator=g)).to(cuda_dev)
    out = ops.clip_embed_ln(pe, cls, pos, w, b, n, T, D, 1e-5)
    x = torch.cat([cls.expand(n, 1, D), pe.float().view(n, T - 1, D)], dim=1) + pos
    _close(out, torch.nn.functional.layer_norm(x, (D,), w, b, 1e-5).view(n * T, D), 1e-2)
    # BEATs gate
    B, Tt, H = 4, 48, 12
    q = torch.randn(B * Tt, 3 * H * 64, generator=g).to(torch.bfloat16).to(cuda_dev)
    gw = (torch.randn(8, 64, generator=g) / 8).to(cuda_dev)
    gb = (0.1 * torch.randn(8, generator=g)).to(cuda_dev)
    ga = (1 + 0.2 * torch.randn(H, generator=g)).to(cuda_dev)
    gate = ops.beats_gate(q, gw, gb, ga, B, Tt, H)
    qh = q[:, : H * 64].float().view(B, Tt, H, 64).transpose(1, 2)
    gg = (qh @ gw.t() + gb).view(B, H, Tt, 2, 4).sum(-1)
    a_, b_ = torch.sigmoid(gg).chunk(2, dim=-1)
    ref = (a_ * (b_ * ga.view(1, H, 1, 1) - 1.0) + 2.0).squeeze(-1)
    _close(gate, ref, 1e-3)
    # group pack / finish round trip
    Cc, G = 768, 16
    x = torch.randn(B, Tt, Cc, generator=g).to(torch.bfloat16).to(cuda_dev)
    xg = ops.beats_group_pack(x, B, Tt, Cc, G)
    ref = x.view(B, Tt, G, Cc // G).permute(2, 0, 1, 3).reshape(G, B, Tt * (Cc // G))
    assert torch.equal(xg, ref)
    bias = torch.randn(Cc, generator=g).to(cuda_dev)
    y = ops.beats_posconv_finish(x, xg, bias, B, Tt, Cc, G)
    _close(y, (x.float() + torch.nn.functional.gelu(x.float() + bias)).view(B * Tt, Cc), 1e-2)
