"""TEST INFRASTRUCTURE — a CPU stand-in for `crab_b200.ops` used ONLY to check host-side plumbing (shapes, strides, weight
packing, index tables) of modules such as crab_b200/seg.py without a GPU.  Each function restates the arithmetic of the C-ABI
kernel it stands for with the kernel's own index formulas, rounds to bf16 where the kernel stores bf16, and enforces the same
argument checks as the C entry point (alignment, N % 8, row strides ...), so a call sequence that passes here is accepted by
the real library.  It is never imported by the product."""
import math

import torch

ACT_NONE, ACT_QUICK_GELU, ACT_GELU = 0, 1, 2
EW_ADD, EW_RELU, EW_GELU, EW_GATE = 0, 1, 2, 3
calls = []


def _al16(t):
    return t.data_ptr() % 16 == 0


def gemm(a, w, *, bias=None, residual=None, res_scale=1.0, out_scale=1.0, act=ACT_NONE, out=None, out_dtype=torch.bfloat16,
         block_n=0, max_ctas=0, k=None, n=None):
    assert a.dim() == 2 and w.dim() == 2 and a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    assert a.stride(1) == 1 and w.stride(1) == 1
    M, K, N = a.shape[0], (k if k is not None else a.shape[1]), (n if n is not None else w.shape[0])
    assert w.shape[1] >= K and a.shape[1] >= K and M > 0 and N > 0 and K > 0
    assert a.stride(0) % 8 == 0 and w.stride(0) % 8 == 0 and a.stride(0) >= K and w.stride(0) >= K, "lda/ldb"
    assert N % 8 == 0, f"N % 8 (N={N})"
    assert K >= 64 and K % 8 == 0, f"keep K a multiple of 8 and at least one k-block on this path (K={K})"
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype)
    assert out.stride(1) == 1 and out.shape[0] == M and out.shape[1] >= N
    assert out.stride(0) % (8 if out.dtype == torch.bfloat16 else 4) == 0, "ldc"
    assert _al16(a) and _al16(w) and _al16(out), "16-byte alignment of A / B / C"
    y = a[:, :K].float() @ w[:N, :K].float().t()
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() >= N and _al16(bias)
        y = y + bias[:N]
    if act == ACT_GELU:
        y = torch.nn.functional.gelu(y)
    elif act == ACT_QUICK_GELU:
        y = y * torch.sigmoid(1.702 * y)
    y = y * out_scale
    if residual is not None:
        assert residual.dtype == torch.bfloat16 and residual.stride(1) == 1 and residual.stride(0) % 8 == 0 and _al16(residual)
        y = y + res_scale * residual[:, :N].float()
    out[:, :N] = y.to(out.dtype)
    calls.append(("gemm", M, N, K))
    return out


def layernorm(x, gamma, beta, eps, out=None):
    assert x.dim() == 2 and x.stride(1) == 1 and x.dtype == torch.bfloat16 and gamma.dtype == torch.float32
    assert x.shape[1] % 8 == 0 and x.stride(0) % 8 == 0
    y = torch.nn.functional.layer_norm(x.float(), (x.shape[1],), gamma, beta, eps).to(torch.bfloat16)
    if out is None:
        return y
    out.copy_(y)
    return out


def gather_rows(src, dst, n, cols, src_rows=None, dst_rows=None):
    assert src.dtype == torch.bfloat16 and dst.dtype == torch.bfloat16 and src.stride(-1) == 1 and dst.stride(-1) == 1
    assert cols % 8 == 0 and src.stride(-2) % 8 == 0 and dst.stride(-2) % 8 == 0
    for r in (src_rows, dst_rows):
        assert r is None or (r.dtype == torch.int64 and r.is_contiguous() and r.numel() >= n)
    si = src_rows[:n] if src_rows is not None else torch.arange(n)
    di = dst_rows[:n] if dst_rows is not None else torch.arange(n)
    assert int(si.max()) < src.shape[0] and int(di.max()) < dst.shape[0]
    dst[di, :cols] = src[si, :cols]


def small_attn(q, k, v, out, heads, head_dim):
    assert head_dim in (16, 32)
    for t in (q, k, v, out):
        assert t.dim() == 2 and t.stride(1) == 1 and t.dtype == torch.bfloat16 and t.shape[1] >= heads * head_dim
    assert k.shape[0] == v.shape[0] and out.shape[0] == q.shape[0]
    sp = lambda t: t[:, : heads * head_dim].float().reshape(t.shape[0], heads, head_dim).transpose(0, 1)  # noqa: E731
    a = torch.softmax(sp(q) @ sp(k).transpose(-1, -2) * head_dim ** -0.5, -1) @ sp(v)
    out[:, : heads * head_dim] = a.transpose(0, 1).reshape(q.shape[0], heads * head_dim).to(torch.bfloat16)
    return out


def elementwise(a, op, b=None, gate=None, out=None):
    assert a.dim() == 2 and a.stride(1) == 1 and a.dtype == torch.bfloat16
    if out is None:
        out = torch.empty((a.shape[0], a.shape[1]), dtype=torch.bfloat16)
    x = a.float()
    if op == EW_ADD:
        assert b is not None and b.dim() == 2 and b.shape[0] in (1, a.shape[0]) and b.shape[1] >= a.shape[1] and b.dtype == torch.bfloat16
        y = x + b[:, : a.shape[1]].float()
    elif op == EW_RELU:
        y = torch.relu(x)
    elif op == EW_GELU:
        y = torch.nn.functional.gelu(x)
    else:
        assert gate is not None and gate.dtype == torch.float32 and gate.numel() >= a.shape[0]
        y = (torch.sigmoid(gate[: a.shape[0]]).unsqueeze(1) + 1.0) * x
    out[:, : a.shape[1]] = y.to(torch.bfloat16)
    return out


def row_mean_f32(x, cols):
    assert x.dim() == 2 and x.dtype == torch.float32 and cols <= x.shape[1]
    return x[:, :cols].mean(1)


def im2col3x3(x, h, w):
    """Same index formula as im2col3x3_kernel: column (tap = ky*3+kx, c), source pixel (y + ky - 1, x + kx - 1), zero outside."""
    assert x.dim() == 2 and x.dtype == torch.bfloat16 and x.shape[0] == h * w
    C = x.shape[1]
    out = torch.zeros((h * w, 9 * C), dtype=torch.bfloat16)
    img = x.reshape(h, w, C)
    for tap in range(9):
        dy, dx = tap // 3 - 1, tap % 3 - 1
        ys, ye = max(0, -dy), min(h, h - dy)
        xs, xe = max(0, -dx), min(w, w - dx)
        blk = torch.zeros((h, w, C), dtype=torch.bfloat16)
        blk[ys:ye, xs:xe] = img[ys + dy:ye + dy, xs + dx:xe + dx]
        out[:, tap * C:(tap + 1) * C] = blk.reshape(h * w, C)
    return out


def bilinear_f32(x, hin, win, hout, wout, channels, out=None, alpha=1.0, beta=0.0, nchw_out=False):
    """Same arithmetic as bilinear_kernel (PyTorch align_corners=False source index, clamped neighbours)."""
    assert x.dim() == 2 and x.dtype == torch.float32 and x.shape[0] == hin * win and x.shape[1] >= channels
    oy = torch.arange(hout, dtype=torch.float32)
    ox = torch.arange(wout, dtype=torch.float32)
    sy = torch.clamp((hin / hout) * (oy + 0.5) - 0.5, min=0.0)
    sx = torch.clamp((win / wout) * (ox + 0.5) - 0.5, min=0.0)
    y0, x0 = sy.long(), sx.long()
    y1, x1 = torch.clamp(y0 + 1, max=hin - 1), torch.clamp(x0 + 1, max=win - 1)
    ly, lx = (sy - y0).view(-1, 1, 1), (sx - x0).view(1, -1, 1)
    img = x[:, :channels].reshape(hin, win, channels)
    val = (1 - ly) * ((1 - lx) * img[y0][:, x0] + lx * img[y0][:, x1]) + ly * ((1 - lx) * img[y1][:, x0] + lx * img[y1][:, x1])
    res = val.permute(2, 0, 1).contiguous() if nchw_out else val.reshape(hout * wout, channels)
    if out is None:
        assert beta == 0.0
        return alpha * res
    out.copy_(beta * out + alpha * res)
    return out
