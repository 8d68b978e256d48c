"""TEST INFRASTRUCTURE — a CPU stand-in for `crab_b200.ops` used ONLY to check host-side plumbing (shapes, strides, weight
packing, index tables) of modules such as crab_b200/seg.py without a GPU.  Each function restates the arithmetic of the C-ABI
kernel it stands for with the kernel's own index formulas, rounds to bf16 where the kernel stores bf16, and enforces the same
argument checks as the C entry point (alignment, N % 8, row strides ...), so a call sequence that passes here is accepted by
the real library.  It is never imported by the product."""
import torch

ACT_NONE, ACT_QUICK_GELU, ACT_GELU, ACT_SWIGLU, ACT_LORA_Z = 0, 1, 2, 3, 4
BF16, F32 = 0, 1
MIN_K = 64          # the seg-head path keeps every GEMM at >= one k-block; engine plumbing tests lower this to 8
EW_ADD, EW_RELU, EW_GELU, EW_GATE = 0, 1, 2, 3
calls = []


def _al16(t):
    return t.data_ptr() % 16 == 0


def gemm(a, w, *, bias=None, residual=None, res_scale=1.0, out_scale=1.0, act=ACT_NONE, out=None, out_dtype=torch.bfloat16,
         block_n=0, max_ctas=0, k=None, n=None):
    """crab_gemm_bf16: out = epilogue(a[M, K] @ w[N, K]^T) with the C entry point's argument checks."""
    assert a.dim() == 2 and w.dim() == 2 and a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    assert a.stride(1) == 1 and w.stride(1) == 1
    M, K, N = a.shape[0], (k if k is not None else a.shape[1]), (n if n is not None else w.shape[0])
    assert w.shape[1] >= K and a.shape[1] >= K and M > 0 and N > 0 and K > 0
    assert a.stride(0) % 8 == 0 and w.stride(0) % 8 == 0 and a.stride(0) >= K and w.stride(0) >= K, "lda/ldb"
    assert K >= MIN_K and K % 8 == 0, f"K={K}: this path keeps K a multiple of 8 and >= {MIN_K}"
    assert _al16(a) and _al16(w), "16-byte alignment of A / B"
    y = a[:, :K].float() @ w[:N, :K].float().t()
    if act in (ACT_SWIGLU, ACT_LORA_Z):
        assert bias is None and residual is None
        y = _lora_z(y, out_scale) if act == ACT_LORA_Z else _swiglu_packed(y, out_scale)
        n_out = y.shape[1]
    else:
        assert N % 8 == 0, f"N % 8 (N={N})"
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
        n_out = N
    if out is None:
        out = torch.empty((M, n_out), dtype=out_dtype)
    assert out.stride(1) == 1 and out.shape[0] == M and out.shape[1] >= n_out and _al16(out)
    assert out.stride(0) % (8 if out.dtype == torch.bfloat16 else 4) == 0, "ldc"
    out[:, :n_out] = y.to(out.dtype)
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


# ======================================================================================================================
# The rest of the library, for running the whole engine (crab_b200/engine.py) on the CPU in plumbing tests
# ======================================================================================================================
_launches = 0


def init(dev=0):
    pass


def set_pdl(mask):
    pass


def set_gemm_2cta(mode):
    pass


def launch_count():
    return _launches


def count_launches(n):
    global _launches
    _launches += n


def _bfr(t):
    return t.to(torch.bfloat16).float()


def _lora_z(y, scale):
    M, N = y.shape
    g = N // 11
    t = y.view(M, g, 11)
    r = torch.softmax(t[..., :3], -1)
    return (scale * r.unsqueeze(-1) * t[..., 3:].unsqueeze(-2)).reshape(M, g * 24)


def _swiglu_packed(y, scale=1.0):
    M, N = y.shape
    t = y.view(M, N // 128, 2, 64)
    return (torch.nn.functional.silu(t[:, :, 0]) * t[:, :, 1] * scale).reshape(M, N // 2)


def rmsnorm(x, gamma, eps, out=None):
    xf = x.float()
    y = _bfr(gamma * _bfr(xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps))).to(torch.bfloat16)
    if out is None:
        return y
    out.copy_(y)
    return out


def rope_table(max_pos, head_dim, theta, device):
    half = head_dim // 2
    inv = (1.0 / (theta ** (torch.arange(0, half, dtype=torch.float64) * 2 / head_dim))).float()
    ang = torch.arange(max_pos, dtype=torch.float32).unsqueeze(1) * inv.unsqueeze(0)
    return torch.cat([torch.cos(ang.double()).float(), torch.sin(ang.double()).float()], 1)


def _rope(x, cs, hd):
    """x (..., hd) float, cs (hd,) = [cos | sin] for one position."""
    half = hd // 2
    c, s = cs[:half], cs[half:]
    lo, hi = x[..., :half], x[..., half:]
    return torch.cat([lo * c - hi * s, hi * c + lo * s], -1)


def rope_kv_append(qkv, cos_sin, k_cache, v_cache, B, S, H, KVH, head_dim, past=0, past_dev=None):
    p0 = int(past_dev[0]) if past_dev is not None else past
    hd = head_dim
    for b in range(B):
        for s_ in range(S):
            row = qkv[b * S + s_]
            cs = cos_sin[p0 + s_]
            q = row[: H * hd].float().view(H, hd)
            k = row[H * hd:(H + KVH) * hd].float().view(KVH, hd)
            v = row[(H + KVH) * hd:(H + 2 * KVH) * hd].view(KVH, hd)
            row[: H * hd] = _rope(q, cs, hd).reshape(-1).to(torch.bfloat16)
            k_cache[b, :, p0 + s_] = _rope(k, cs, hd).to(torch.bfloat16)
            v_cache[b, :, p0 + s_] = v


def flash_attn(q, k, v, out, *, B, H, KVH, Sq, Sk, head_dim, q_strides, k_strides, v_strides, o_strides, scale, causal=False,
               gate=None, bias_table=None, sk_dev=None):
    if sk_dev is not None:
        Sk = int(sk_dev[0])
    hd = head_dim

    def view(t, n_heads, S, st):
        return torch.as_strided(t, (B, n_heads, S, hd), (st[0], st[2], st[1], 1), t.storage_offset())

    qf, kf, vf = view(q, H, Sq, q_strides).float(), view(k, KVH, Sk, k_strides).float(), view(v, KVH, Sk, v_strides).float()
    if KVH != H:
        kf, vf = kf.repeat_interleave(H // KVH, 1), vf.repeat_interleave(H // KVH, 1)
    s = qf @ kf.transpose(-1, -2) * scale
    if bias_table is not None:
        s = s + gate.unsqueeze(-1) * bias_table.unsqueeze(0)
    if causal:
        s = s + torch.full((Sq, Sk), float("-inf")).triu(diagonal=Sk - Sq + 1)
    o = torch.softmax(s, -1) @ vf
    view(out, H, Sq, o_strides).copy_(o.to(torch.bfloat16))
    return out


def _lora_rows(src, ra, groups, scale):
    t = src.float() @ ra[: groups * 11].float().t()
    return _lora_z(t, scale)


def row_norm_loraz(x, *, gamma=None, eps=0.0, y=None, ra=None, groups=0, z=None, scale=1.0):
    src = x
    if gamma is not None:
        y.copy_(rmsnorm(x, gamma, eps))
        src = y
    if groups:
        z[:, : groups * 24] = _lora_rows(src, ra, groups, scale).to(torch.bfloat16)


def attn_decode_fused(qkv, rope, k_cache, v_cache, out, *, B, H, KVH, head_dim, scale, past_dev, nsplit=1, workspace=None,
                      ra=None, z=None, lora_scale=0.0, lora_ws=None, lora_counters=None, gqa_tc=False):
    hd, past = head_dim, int(past_dev[0])
    G = H // KVH
    for b in range(B):
        row = qkv[b]
        cs = rope[past]
        q = _bfr(_rope(row[: H * hd].float().view(H, hd), cs, hd))
        k_cache[b, :, past] = _rope(row[H * hd:(H + KVH) * hd].float().view(KVH, hd), cs, hd).to(torch.bfloat16)
        v_cache[b, :, past] = row[(H + KVH) * hd:(H + 2 * KVH) * hd].view(KVH, hd)
        kk = k_cache[b, :, : past + 1].float().repeat_interleave(G, 0)
        vv = v_cache[b, :, : past + 1].float().repeat_interleave(G, 0)
        a = torch.softmax((q.unsqueeze(1) @ kk.transpose(-1, -2)) * scale, -1) @ vv      # (H, 1, hd)
        out[b, : H * hd] = a.reshape(-1).to(torch.bfloat16)
    if ra is not None:   # unsplit: in the attention launch; split KV: in the combine launch — same arithmetic
        z[:, :24] = _lora_rows(out[:, : H * hd], ra, 1, lora_scale).to(torch.bfloat16)
    return out


class PackedWeight:
    def __init__(self, data, N, K, swiglu=False):
        self.data, self.N, self.K, self.swiglu = data, N, K, swiglu


def pack_skinny_weight(w, k=None, swiglu=False):
    return PackedWeight(w, w.shape[0], k if k is not None else w.shape[1], swiglu)


def _fused_linear(x, w, k, z, kext, stats, L, norm, eps, scale, bias, residual, act, out, n, M):
    """One decode linear with the RMSNorm as an epilogue scale and the hyper-LoRA pre-pass from the statistics rows."""
    xf = x[:M, :k].float()
    rstd = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps) if norm else torch.ones(M, 1)
    if L:
        t = (xf @ stats[: 11 * L].float().t()).view(M, L, 11)
        r = torch.softmax(t[..., :3] * rstd.unsqueeze(-1), -1) * scale
        z[:M, : L * 24] = (r.unsqueeze(-1) * t[..., 3:].unsqueeze(-2)).reshape(M, L * 24).to(torch.bfloat16)
    W = w.data
    y = xf @ W[: w.N, :k].float().t()
    if kext:
        kx = min(kext, z.shape[1])
        y = y + z[:M, :kx].float() @ W[: w.N, k: k + kx].float().t()
    y = y * rstd
    if act == ACT_SWIGLU:
        assert w.swiglu and bias is None and residual is None
        y = _swiglu_packed(y)
    else:
        y = y[:, :n]
        if bias is not None:
            y = y + bias[:n]
        if residual is not None:
            y = y + residual[:M, :n].float()
    out[:M, : y.shape[1]] = y.to(out.dtype)


def gemm_skinny(x, w, *, bias=None, residual=None, act=ACT_NONE, out=None, out_dtype=torch.bfloat16, k=None, n=None, splits=0,
                z=None, kext=0, stats=None, stats_linears=0, norm=False, eps=0.0, lora_scale=1.0, rstd=None, flags=None, tag="",
                prefetch=None, prefetch_bytes=0, stats_scratch=None, flags_clear=None, stats_clusters=0):
    packed = isinstance(w, PackedWeight)
    if packed and (kext or norm or stats_linears):
        assert (not norm and not stats_linears) or flags is not None
        K, N = w.K - kext, (n if n is not None else w.N)
        assert x.shape[0] <= 32 and K % 64 == 0 or MIN_K < 64
        n_out = N // 2 if act == ACT_SWIGLU else N
        if out is None:
            out = torch.empty((x.shape[0], n_out), dtype=out_dtype)
        _fused_linear(x, w, K, z, kext, stats, stats_linears, norm, eps, lora_scale, bias, residual, act, out, N, x.shape[0])
        return out
    W = w.data if packed else w
    K = w.K if packed else (k if k is not None else x.shape[1])
    N = w.N if packed else (n if n is not None else W.shape[0])
    if n is not None:
        N = n
    assert x.shape[0] <= 32 and x.stride(0) % 8 == 0 and x.shape[1] >= K
    y = x[:, :K].float() @ W[:N, :K].float().t()
    if act == ACT_SWIGLU:
        assert packed and w.swiglu and bias is None and residual is None
        y = _swiglu_packed(y)
    else:
        if bias is not None:
            y = y + bias[:N]
        if residual is not None:
            y = y + residual[:, : y.shape[1]].float()
    if out is None:
        out = torch.empty((x.shape[0], y.shape[1]), dtype=out_dtype)
    out[:, : y.shape[1]] = y.to(out.dtype)
    return out


def pack_chain_stats(ra, gamma=None):
    """stands for ops.pack_chain_stats: keeps the dense (gamma-folded) router/A rows; the swizzled stream is a layout detail."""
    d = ra.float() if gamma is None else ra.float() * gamma.float()[None, :]
    return d.to(torch.bfloat16)


class ChainPhase:
    def __init__(self, x, w, out, *, k, z=None, kext=0, stats=None, stats_linears=0, norm=False, eps=0.0, lora_scale=1.0,
                 rstd=None, bias=None, residual=None, act=ACT_NONE, n=None):
        assert isinstance(w, PackedWeight) and w.K == k + kext and x.shape[1] >= k and x.stride(0) % 8 == 0 and k % 64 == 0 or MIN_K < 64
        assert stats_linears == 0 or (stats is not None and kext >= 24 * stats_linears and z is not None)
        assert not norm or rstd is not None
        self.x, self.w, self.out, self.k, self.z, self.kext = x, w, out, k, z, kext
        self.stats, self.stats_linears, self.norm, self.eps, self.lora_scale = stats, stats_linears, norm, eps, lora_scale
        self.rstd, self.bias, self.residual, self.act, self.n = rstd, bias, residual, act, (n if n is not None else w.N)


def decode_chain(phases, M, counters, cluster=0, max_clusters=0, tag="crab_decode_chain"):
    """crab_decode_chain: dependent M <= 32 linears; RMSNorm as rstd in the epilogue (gamma folded into the weights by the
    caller), hyper-LoRA pre-pass from the statistics rows, z' = scale * softmax(rstd * logits) * u un-normalised."""
    assert 1 <= len(phases) <= 4 and M <= 32
    for ph in phases:
        _fused_linear(ph.x, ph.w, ph.k, ph.z, ph.kext, ph.stats, ph.stats_linears, ph.norm, ph.eps, ph.lora_scale, ph.bias,
                      ph.residual, ph.act, ph.out, ph.n, M)


def cross_entropy(logits, V, labels):
    loss = torch.zeros(logits.shape[0])
    ok = labels >= 0
    if ok.any():
        loss[ok] = torch.nn.functional.cross_entropy(logits[ok, :V].float(), labels[ok], reduction="none")
    return loss


def sample_top_k_top_p(logits, V, u, *, temperature=1.0, top_k=0, top_p=1.0, out=None):
    """HF order: temperature -> top-k -> top-p (keep the smallest most-probable set reaching top_p), inverse CDF in index order."""
    ids = []
    for r in range(logits.shape[0]):
        s = logits[r, :V].float() / temperature
        if 0 < top_k < V:
            s = torch.where(s >= s.topk(top_k).values[-1], s, torch.full_like(s, -float("inf")))
        if top_p < 1.0:
            sv, si = s.sort(descending=False)
            remove = sv.softmax(-1).cumsum(-1) <= (1 - top_p)
            remove[-1] = False
            s = s.masked_fill(torch.zeros_like(s, dtype=torch.bool).scatter(0, si, remove), -float("inf"))
        p = (s - s.max()).exp()
        c = p.cumsum(0)
        ids.append(int((c > float(u[r]) * c[-1]).nonzero()[0]))
    r_ = torch.tensor(ids, dtype=torch.int64)
    if out is None:
        return r_
    out.copy_(r_)
    return out


def argmax(logits, V, out=None):
    r = logits[:, :V].argmax(-1)
    if out is None:
        return r
    out.copy_(r)
    return out


def add_scalar_i32(p, v):
    p += v


def patchify(images, patch, ld_out):
    n, c, h, w = images.shape
    cols = torch.nn.functional.unfold(images, patch, stride=patch).transpose(1, 2).reshape(n * (h // patch) * (w // patch), c * patch * patch)
    out = torch.zeros((cols.shape[0], ld_out), dtype=torch.bfloat16)
    out[:, : cols.shape[1]] = cols.to(torch.bfloat16)
    return out


def clip_embed_ln(patch_emb, cls, pos, gamma, beta, n_img, tokens, D, eps):
    x = torch.empty((n_img, tokens, D))
    x[:, 0] = cls + pos[0]
    x[:, 1:] = patch_emb.float().view(n_img, tokens - 1, D) + pos[1:]
    return torch.nn.functional.layer_norm(x, (D,), gamma, beta, eps).reshape(n_img * tokens, D).to(torch.bfloat16)


# ---- BEATs helpers (csrc/elementwise.cu: beats_gate_kernel, beats_group_pack_kernel, beats_posconv_finish_kernel) ----------
def beats_gate(q, grep_w, grep_b, grep_a, B, T, H):
    """gate[b, h, t] = ga * (gb * grep_a[h] - 1) + 2, (ga, gb) = sigmoid of the two 4-output sums of grep_linear(q_head)."""
    x = q[:, : H * 64].float().view(B, T, H, 64)
    y = x @ grep_w.t() + grep_b                                    # (B, T, H, 8)
    ga, gb = torch.sigmoid(y[..., :4].sum(-1)), torch.sigmoid(y[..., 4:].sum(-1))
    return (ga * (gb * grep_a.view(1, 1, H) - 1.0) + 2.0).permute(0, 2, 1).contiguous()


def beats_group_pack(x, B, T, Cc, G):
    cg = Cc // G
    return x.view(B, T, G, cg).permute(2, 0, 1, 3).reshape(G, B, T * cg).contiguous()


def beats_posconv_finish(x, conv_g, bias, B, T, Cc, G):
    cg = Cc // G
    conv = conv_g.float().view(G, B, T, cg).permute(1, 2, 0, 3).reshape(B * T, Cc) + bias
    return (x.float() + torch.nn.functional.gelu(conv)).to(torch.bfloat16)
