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
