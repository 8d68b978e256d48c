"""GEMM kernel parity (tcgen05 path) against a plain fp32 torch matmul of the same bf16-rounded operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, w, bias=None, residual=None, res_scale=1.0, act=0, out_scale=1.0):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias
    if act == 1:
        y = y * torch.sigmoid(1.702 * y)
    elif act == 2:
        y = torch.nn.functional.gelu(y)
    y = y * out_scale
    if residual is not None:
        y = y + res_scale * residual.float()
    return y


def _check(out, ref, tol=1e-2):
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    assert err / scale < tol, f"max err {err} vs scale {scale}"


@pytest.mark.parametrize("bn", [64, 128, 256])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 256), (300, 264, 200), (1000, 1024, 1032), (32, 4096, 4096)])
def test_gemm_plain(cuda_dev, bn, M, N, K):
    from crab_b200 import ops

    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).to(cuda_dev)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(cuda_dev)
    out = ops.gemm(a, w, block_n=bn)
    torch.cuda.synchronize()
    _check(out, _ref(a, w))


def test_gemm_multi_tile_per_cta(cuda_dev):
    """Few CTAs, many tiles: exercises the smem ring wrap-around and the TMEM double buffer phases."""
    from crab_b200 import ops

    g = torch.Generator(device="cpu").manual_seed(5)
    a = torch.randn(1500, 520, generator=g).to(torch.bfloat16).to(cuda_dev)
    w = (torch.randn(776, 520, generator=g) / 23).to(torch.bfloat16).to(cuda_dev)
    for bn in (64, 128, 256):
        out = ops.gemm(a, w, block_n=bn, max_ctas=3)
        torch.cuda.synchronize()
        _check(out, _ref(a, w))


@pytest.mark.parametrize("act", [0, 1, 2])
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
def test_gemm_epilogues(cuda_dev, act, out_dtype):
    from crab_b200 import ops

    M, N, K = 515, 776, 328
    g = torch.Generator(device="cpu").manual_seed(11 + act)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).to(cuda_dev)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(cuda_dev)
    bias = torch.randn(N, generator=g).to(cuda_dev)
    res = torch.randn(M, N, generator=g).to(torch.bfloat16).to(cuda_dev)
    out = ops.gemm(a, w, bias=bias, residual=res, res_scale=2.2, act=act, out_dtype=out_dtype, out_scale=0.5)
    torch.cuda.synchronize()
    assert out.dtype == out_dtype
    _check(out, _ref(a, w, bias, res, 2.2, act, 0.5))


def test_gemm_strided_views(cuda_dev):
    """A and C as column slices of wider buffers (the K-extension layout used for hyper-LoRA)."""
    from crab_b200 import ops

    M, N, K = 260, 136, 192
    g = torch.Generator(device="cpu").manual_seed(3)
    abuf = torch.randn(M, K + 64, generator=g).to(torch.bfloat16).to(cuda_dev)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(cuda_dev)
    cbuf = torch.zeros(M, N + 40, dtype=torch.bfloat16, device=cuda_dev)
    ops.gemm(abuf[:, :K], w, out=cbuf[:, 8:8 + N])
    torch.cuda.synchronize()
    _check(cbuf[:, 8:8 + N], _ref(abuf[:, :K], w))
    assert cbuf[:, :8].abs().max().item() == 0 and cbuf[:, 8 + N:].abs().max().item() == 0


def test_gemm_swiglu(cuda_dev):
    from crab_b200 import ops

    M, F, K = 300, 384, 256  # F = intermediate size; packed weight has 2F rows in [64 gate | 64 up] groups
    g = torch.Generator(device="cpu").manual_seed(9)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).to(cuda_dev)
    wg = (torch.randn(F, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(cuda_dev)
    wu = (torch.randn(F, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(cuda_dev)
    packed = torch.stack([wg.view(F // 64, 64, K), wu.view(F // 64, 64, K)], dim=1).reshape(2 * F, K).contiguous()
    ref = torch.nn.functional.silu(a.float() @ wg.float().t()) * (a.float() @ wu.float().t())
    for bn in (128, 256):
        out = ops.gemm(a, packed, act=ops.ACT_SWIGLU, block_n=bn)
        torch.cuda.synchronize()
        assert out.shape == (M, F)
        _check(out, ref)


@pytest.fixture
def force_pair(cuda_dev):
    """Route every block_n=256 GEMM with >= 2 tiles through the cta_group::2 CTA-pair kernel for the duration of a test."""
    from crab_b200 import ops

    ops.set_gemm_2cta(2)
    yield
    ops.set_gemm_2cta(1)


@pytest.mark.parametrize("M,N,K,max_ctas", [(512, 512, 64, 0), (700, 776, 520, 0), (1500, 776, 520, 6), (4173, 1304, 1000, 0),
                                             (257, 1024, 4192, 0), (16640, 520, 2080, 0)])
def test_gemm_cta_pair_plain(cuda_dev, force_pair, M, N, K, max_ctas):
    """cta_group::2: ragged M / N / K tails, odd tile counts per pair, smem ring wrap-around with few pairs, and (last
    case) a shape the auto policy itself sends to the pair kernel."""
    from crab_b200 import ops

    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).to(cuda_dev)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(cuda_dev)
    out = ops.gemm(a, w, block_n=256, max_ctas=max_ctas)
    torch.cuda.synchronize()
    _check(out, _ref(a, w))
    ops.set_gemm_2cta(0)
    single = ops.gemm(a, w, block_n=256, max_ctas=max_ctas)
    assert torch.equal(out, single)  # same k order, same fp32 accumulation: bit-identical to the single-CTA kernel


@pytest.mark.parametrize("act", [0, 1, 2])
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
def test_gemm_cta_pair_epilogues(cuda_dev, force_pair, act, out_dtype):
    from crab_b200 import ops

    M, N, K = 515, 776, 328
    g = torch.Generator(device="cpu").manual_seed(11 + act)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).to(cuda_dev)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(cuda_dev)
    bias = torch.randn(N, generator=g).to(cuda_dev)
    res = torch.randn(M, N, generator=g).to(torch.bfloat16).to(cuda_dev)
    out = ops.gemm(a, w, bias=bias, residual=res, res_scale=2.2, act=act, out_dtype=out_dtype, out_scale=0.5, block_n=256)
    torch.cuda.synchronize()
    _check(out, _ref(a, w, bias, res, 2.2, act, 0.5))
    # in-place residual (x += ...) and the SwiGLU pair epilogue
    x = res.clone()
    ops.gemm(a, w, residual=x, out=x, block_n=256)
    _check(x, _ref(a, w, None, res, 1.0, 0, 1.0))
    F = 384
    wg = (torch.randn(F, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(cuda_dev)
    wu = (torch.randn(F, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(cuda_dev)
    packed = torch.stack([wg.view(F // 64, 64, K), wu.view(F // 64, 64, K)], dim=1).reshape(2 * F, K).contiguous()
    sw = ops.gemm(a, packed, act=ops.ACT_SWIGLU, block_n=256)
    _check(sw, torch.nn.functional.silu(a.float() @ wg.float().t()) * (a.float() @ wu.float().t()))


def test_gemm_lora_z(cuda_dev):
    from crab_b200 import ops

    M, K, G = 333, 512, 3
    g = torch.Generator(device="cpu").manual_seed(21)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).to(cuda_dev)
    ra = (torch.randn(G * 11, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(cuda_dev)
    out = ops.gemm(a, ra, act=ops.ACT_LORA_Z, out_scale=2.0)
    torch.cuda.synchronize()
    y = (a.float() @ ra.float().t()).view(M, G, 11)
    r = torch.softmax(y[..., :3], dim=-1)
    u = y[..., 3:]
    ref = (2.0 * r.unsqueeze(-1) * u.unsqueeze(-2)).reshape(M, G * 24)
    assert out.shape == (M, G * 24)
    _check(out, ref)


@pytest.mark.parametrize("M", [1, 7, 32])
@pytest.mark.parametrize("N,K,splits", [(256, 512, 0), (4096, 4128, 0), (12288, 4192, 3), (1000, 328, 2), (32024, 1024, 1),
                                        (4096, 11040, 8), (4096, 11040, 5), (22016, 4160, 7), (12288, 4192, 0), (384, 4096, 6)])
def test_gemm_skinny(cuda_dev, M, N, K, splits):
    from crab_b200 import ops

    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    ld = K + 24  # the decode activation buffers are wider than K
    xbuf = torch.randn(M, ld, generator=g).to(torch.bfloat16).to(cuda_dev)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(cuda_dev)
    bias = torch.randn(N, generator=g).to(cuda_dev)
    res = torch.randn(M, N, generator=g).to(torch.bfloat16).to(cuda_dev)
    ref = xbuf[:, :K].float() @ w.float().t()
    out = ops.gemm_skinny(xbuf, w, k=K, splits=splits)
    _check(out, ref)
    out2 = ops.gemm_skinny(xbuf, w, k=K, splits=splits)  # ticket counters must have been reset by the kernel
    assert torch.equal(out, out2)
    out = ops.gemm_skinny(xbuf, w, k=K, bias=bias, residual=res, splits=splits, out_dtype=torch.float32)
    assert out.dtype == torch.float32
    _check(out, ref + bias + res.float())
    # in-place residual (the decoder's x += proj(...)): C aliases the residual
    x_inplace = res.clone()
    ops.gemm_skinny(xbuf, w, k=K, residual=x_inplace, out=x_inplace, splits=splits)
    _check(x_inplace, ref + res.float())
    # streaming layout (pre-tiled, pre-swizzled weights read with 1-D bulk copies): bit-identical to the TMA path
    wp = ops.pack_skinny_weight(w)
    out_p = ops.gemm_skinny(xbuf, wp, bias=bias, residual=res, splits=splits, out_dtype=torch.float32)
    assert torch.equal(out_p, out)


def test_gemm_skinny_swiglu_matches_prefill_kernel(cuda_dev):
    from crab_b200 import ops

    M, F, K = 32, 1024, 512
    g = torch.Generator(device="cpu").manual_seed(2)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).to(cuda_dev)
    wg = (torch.randn(F, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(cuda_dev)
    wu = (torch.randn(F, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(cuda_dev)
    packed = torch.stack([wg.view(F // 64, 64, K), wu.view(F // 64, 64, K)], dim=1).reshape(2 * F, K).contiguous()
    ref = torch.nn.functional.silu(a.float() @ wg.float().t()) * (a.float() @ wu.float().t())
    pk = ops.pack_skinny_weight(packed, swiglu=True)  # prefill layout -> interleaved decode layout
    for splits in (1, 2, 0, 3, 5, 7, 8):  # K-split = cluster size (0 = auto)
        out = ops.gemm_skinny(a, pk, act=ops.ACT_SWIGLU, splits=splits)
        assert out.shape == (M, F)
        _check(out, ref)
    _check(ops.gemm(a, packed, act=ops.ACT_SWIGLU), ref)


@pytest.mark.parametrize("cols,groups,norm", [(4096, 3, True), (4096, 2, True), (4096, 1, False), (11008, 1, False), (256, 3, True),
                                              (3584, 3, True), (18944, 1, False),
                                              (4096, 0, True)])
def test_row_norm_loraz(cuda_dev, cols, groups, norm):
    from crab_b200 import ops

    rows = 9
    g = torch.Generator(device="cpu").manual_seed(cols + groups)
    x = (torch.randn(rows, cols, generator=g) * 1.5).to(torch.bfloat16).to(cuda_dev)
    gamma = (1 + 0.1 * torch.randn(cols, generator=g)).to(cuda_dev)
    ra = (torch.randn(max(groups, 1) * 11, cols, generator=g) / cols ** 0.5).to(torch.bfloat16).to(cuda_dev)
    buf = torch.zeros(rows, cols + 96, dtype=torch.bfloat16, device=cuda_dev)
    if norm:
        ops.row_norm_loraz(x, gamma=gamma, eps=1e-6, y=buf[:, :cols], ra=ra if groups else None, groups=groups,
                           z=buf[:, cols:] if groups else None, scale=2.0)
        y_ref = ops.rmsnorm(x, gamma, 1e-6)
        assert torch.equal(buf[:, :cols], y_ref)  # identical rounding to the prefill norm kernel
        src = y_ref
    else:
        buf[:, :cols] = x
        ops.row_norm_loraz(buf[:, :cols], ra=ra, groups=groups, z=buf[:, cols:], scale=2.0)
        src = x
    if groups:
        t = (src.float() @ ra.float().t()).view(rows, groups, 11)
        r = torch.softmax(t[..., :3], dim=-1)
        ref = (2.0 * r.unsqueeze(-1) * t[..., 3:].unsqueeze(-2)).reshape(rows, groups * 24)
        _check(buf[:, cols:cols + groups * 24], ref)
        assert buf[:, cols + groups * 24:].abs().max().item() == 0
