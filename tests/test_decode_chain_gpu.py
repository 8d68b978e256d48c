"""crab_decode_chain (csrc/decode_chain.cu) against a plain fp32 torch statement of the same linears: single phases at every
cluster size, the RMSNorm-as-rstd epilogue, the statistics item (sum of squares on the diagonal of x x^T, router/A dots,
fp32 router softmax -> z'), K-extension columns, SwiGLU, bias, in-place residual, ragged N (lm_head), M < 32, and a full
four-phase layer chain (o -> gate/up -> down -> next qkv) at LLaMA-7B and Qwen2-7B widths.  Tolerances are relative L2 against
fp32 math on the same bf16 operands (the kernel accumulates in fp32; differences come from bf16 storage of z' / outputs and from
folding gamma into bf16 weights)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def mk(shape, dev, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (scale * torch.randn(shape, generator=g)).to(dev)


class Lin:
    """One hyper-LoRA linear group sharing an input: `linears` wrapped linears of `n_each` rows (peft_hyper/tuners/lora.py:338-369)."""

    def __init__(self, dev, seed, K, n_each, lora=True, gamma=False, bias=False, swiglu=False, kext=None):
        from crab_b200 import ops

        self.K, self.n_each, self.L = K, list(n_each), len(n_each)
        N = sum(n_each)
        self.N = N
        self.W = mk((N, K), dev, seed, 1 / math.sqrt(K)).to(torch.bfloat16)
        self.gamma = (1.0 + 0.1 * mk((K,), dev, seed + 1)) if gamma else None
        self.bias = 0.05 * mk((N,), dev, seed + 2) if bias else None
        self.lora = lora
        self.kext = (kext if kext is not None else 32 * ((24 * self.L + 31) // 32)) if lora else 0
        self.scale = 2.0
        Wx = torch.zeros((N, K + self.kext), device=dev, dtype=torch.bfloat16)
        Wf = self.W.float() * (self.gamma[None, :] if gamma else 1.0)
        Wx[:, :K] = Wf.to(torch.bfloat16)
        if lora:
            self.RA = mk((11 * self.L, K), dev, seed + 3, 1 / math.sqrt(K)).to(torch.bfloat16)      # rows: 3 route + 8 A per linear
            self.B = [0.05 * mk((n, 24), dev, seed + 10 + i) for i, n in enumerate(n_each)]         # [B0 | B1 | B2] per linear
            r0 = 0
            for i, n in enumerate(n_each):
                Wx[r0:r0 + n, K + 24 * i:K + 24 * (i + 1)] = self.B[i].to(torch.bfloat16)
                r0 += n
            self.stats = ops.pack_chain_stats(self.RA, self.gamma)
        else:
            self.stats = None
        self.swiglu = swiglu
        if swiglu:   # prefill layout: [64 gate | 64 up] row groups; n_each = (F, F) = gate rows then up rows
            F = n_each[0]
            Wg = Wx.view(2, F // 64, 64, K + self.kext).permute(1, 0, 2, 3).reshape(N, K + self.kext).contiguous()
            self.packed = ops.pack_skinny_weight(Wg, k=K + self.kext, swiglu=True)
        else:
            self.packed = ops.pack_skinny_weight(Wx, k=K + self.kext)
        self.Wx = Wx

    def ref(self, x, residual=None):
        """fp32 statement: y = xn W^T + sum_i softmax(xn R^T)_i B_i (A xn) scale (+bias) (+residual), xn = rmsnorm(x) * gamma."""
        x = x.float()
        xn = x
        if self.gamma is not None:
            xn = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6) * self.gamma
        y = xn @ self.W.float().t()
        if self.lora:
            t = (xn @ self.RA.float().t()).view(-1, self.L, 11)
            r = torch.softmax(t[..., :3], -1)
            r0 = 0
            for i, n in enumerate(self.n_each):
                z = (r[:, i, :, None] * t[:, i, None, 3:]).reshape(-1, 24) * self.scale
                y[:, r0:r0 + n] += z @ self.B[i].to(torch.bfloat16).float().t()
                r0 += n
        if self.bias is not None:
            y = y + self.bias
        if self.swiglu:
            F = self.n_each[0]
            y = torch.nn.functional.silu(y[:, :F]) * y[:, F:]
        if residual is not None:
            y = y + residual.float()
        return y

    def phase(self, x, out, zbuf=None, rstd=None, residual=None, external_z=None):
        from crab_b200 import ops

        if external_z is not None:
            return ops.ChainPhase(x, self.packed, out, k=self.K, z=external_z, kext=self.kext, bias=self.bias, residual=residual)
        return ops.ChainPhase(x, self.packed, out, k=self.K, z=zbuf if self.lora else None, kext=self.kext, stats=self.stats,
                              stats_linears=self.L if self.lora else 0, norm=self.gamma is not None, eps=1e-6, lora_scale=self.scale,
                              rstd=rstd, bias=self.bias, residual=residual, act=ops.ACT_SWIGLU if self.swiglu else ops.ACT_NONE)


@pytest.mark.parametrize("cluster", [1, 2, 4, 8])
@pytest.mark.parametrize("M", [32, 5])
def test_single_phase_plain_and_residual(cuda_dev, cluster, M):
    from crab_b200 import ops

    lin = Lin(cuda_dev, 1, 4096, [4096], lora=False)
    x = mk((M, 4096), cuda_dev, 2).to(torch.bfloat16)
    res = mk((M, 4096), cuda_dev, 3).to(torch.bfloat16)
    out = res.clone()
    cnt = torch.zeros(288, dtype=torch.int32, device=cuda_dev)
    ops.decode_chain([lin.phase(x, out, residual=out)], M, cnt, cluster)
    torch.cuda.synchronize()
    e = rel(out, lin.ref(x, res))
    assert e < 4e-3, e
    assert int(cnt.abs().sum()) == 0, "the kernel must leave its counters zero"
    out2 = res.clone()
    ops.decode_chain([lin.phase(x, out2, residual=out2)], M, cnt, cluster)
    assert torch.equal(out, out2), "deterministic (fixed reduction order)"


@pytest.mark.parametrize("cluster", [2, 4, 8])
def test_norm_stats_lora_bias_phase(cuda_dev, cluster):
    """qkv-like: RMSNorm folded (gamma in the weights, rstd in the epilogue), three LoRA linears sharing the input, bias."""
    from crab_b200 import ops

    lin = Lin(cuda_dev, 11, 4096, [4096, 512, 512], lora=True, gamma=True, bias=True, kext=96)
    for M in (32, 7):
        x = mk((M, 4096), cuda_dev, 12, 3.0).to(torch.bfloat16)
        out = torch.empty((M, lin.N), device=cuda_dev, dtype=torch.bfloat16)
        z = torch.zeros((32, 128), device=cuda_dev, dtype=torch.bfloat16)
        rstd = torch.zeros(32, device=cuda_dev, dtype=torch.float32)
        cnt = torch.zeros(288, dtype=torch.int32, device=cuda_dev)
        ops.decode_chain([lin.phase(x, out, zbuf=z, rstd=rstd)], M, cnt, cluster)
        torch.cuda.synchronize()
        r_ref = torch.rsqrt(x.float().pow(2).mean(-1) + 1e-6)
        assert torch.allclose(rstd[:M], r_ref, rtol=1e-5, atol=0), (rstd[:M], r_ref)
        e = rel(out, lin.ref(x))
        print(f"cluster {cluster} M {M}: norm+lora+bias phase rel_l2 {e:.3e}")
        assert e < 6e-3, e


def test_ragged_n_fp32_head_phase(cuda_dev):
    """lm_head-like: N = 32017 (not a multiple of 128 / 8), fp32 output, final norm folded, no LoRA."""
    from crab_b200 import ops

    lin = Lin(cuda_dev, 21, 4096, [32017], lora=False, gamma=True)
    M = 32
    x = mk((M, 4096), cuda_dev, 22, 5.0).to(torch.bfloat16)
    out = torch.full((M, 32024), -7.0, device=cuda_dev, dtype=torch.float32)
    rstd = torch.zeros(32, device=cuda_dev, dtype=torch.float32)
    cnt = torch.zeros(288, dtype=torch.int32, device=cuda_dev)
    ops.decode_chain([ops.ChainPhase(x, lin.packed, out, k=4096, norm=True, eps=1e-6, rstd=rstd, n=32017)], M, cnt, 4)
    torch.cuda.synchronize()
    e = rel(out[:, :32017], lin.ref(x))
    assert e < 4e-3, e
    assert bool((out[:, 32020:] == -7.0).all()), "columns past N (beyond the 4-column vector tail) are untouched"


@pytest.mark.parametrize("dims", [dict(D=4096, F=11008, nq=4096, nk=4096, bias=False), dict(D=3584, F=18944, nq=3584, nk=512, bias=True)])
@pytest.mark.parametrize("cluster", [4, 8])
def test_layer_chain_four_phases(cuda_dev, dims, cluster):
    """o_proj (+residual, z from an earlier kernel) -> RMSNorm + gate/up + SwiGLU -> down (+residual) -> RMSNorm + next qkv (+bias),
    one launch, against the fp32 statement chained through bf16 activations like the kernel's buffers."""
    from crab_b200 import ops

    D, F, nq, nk = dims["D"], dims["F"], dims["nq"], dims["nk"]
    dev = cuda_dev
    o = Lin(dev, 31, nq, [D], lora=True)
    gu = Lin(dev, 32, D, [F, F], lora=True, gamma=True, swiglu=True, kext=64)
    dn = Lin(dev, 33, F, [D], lora=True)
    qkv = Lin(dev, 34, D, [nq, nk, nk], lora=True, gamma=True, bias=dims["bias"], kext=96)
    for M in (32, 3):
        at = torch.zeros((M, nq + 32), device=dev, dtype=torch.bfloat16)
        at[:, :nq] = mk((M, nq), dev, 35).to(torch.bfloat16)
        x0 = mk((M, D), dev, 36, 2.0).to(torch.bfloat16)
        # what the fused decode attention would have written: o_proj's z columns
        t = (at[:, :nq].float() @ o.RA.float().t()).view(M, 1, 11)
        zo = (torch.softmax(t[..., :3], -1)[:, 0, :, None] * t[:, 0, None, 3:]).reshape(M, 24) * o.scale
        at[:, nq:nq + 24] = zo.to(torch.bfloat16)
        x = x0.clone()
        hh = torch.empty((M, F), device=dev, dtype=torch.bfloat16)
        out_qkv = torch.empty((M, qkv.N), device=dev, dtype=torch.bfloat16)
        zb = {k: torch.zeros((32, 128), device=dev, dtype=torch.bfloat16) for k in ("gu", "d", "qkv")}
        rs = {k: torch.zeros(32, device=dev, dtype=torch.float32) for k in ("gu", "qkv")}
        cnt = torch.zeros(288, dtype=torch.int32, device=dev)
        phases = [o.phase(at, x, residual=x, external_z=at[:, nq:]),
                  gu.phase(x, hh, zbuf=zb["gu"], rstd=rs["gu"]),
                  dn.phase(hh, x, zbuf=zb["d"], residual=x),
                  qkv.phase(x, out_qkv, zbuf=zb["qkv"], rstd=rs["qkv"])]
        ops.decode_chain(phases, M, cnt, cluster)
        torch.cuda.synchronize()
        # reference, rounding activations to bf16 where the kernel stores them
        x1 = o.ref(at[:, :nq], x0).to(torch.bfloat16)
        h1 = gu.ref(x1).to(torch.bfloat16)
        x2 = dn.ref(h1, x1).to(torch.bfloat16)
        q2 = qkv.ref(x2)
        e = (rel(x, x2), rel(hh, h1), rel(out_qkv, q2))
        print(f"dims {D}/{F} cluster {cluster} M {M}: x {e[0]:.3e} h {e[1]:.3e} qkv {e[2]:.3e}")
        assert max(e) < 8e-3, e
        assert int(cnt.abs().sum()) == 0
        # replay: bit-identical
        x_b, hh_b, q_b = x0.clone(), torch.empty_like(hh), torch.empty_like(out_qkv)
        phases = [o.phase(at, x_b, residual=x_b, external_z=at[:, nq:]), gu.phase(x_b, hh_b, zbuf=zb["gu"], rstd=rs["gu"]),
                  dn.phase(hh_b, x_b, zbuf=zb["d"], residual=x_b), qkv.phase(x_b, q_b, zbuf=zb["qkv"], rstd=rs["qkv"])]
        ops.decode_chain(phases, M, cnt, cluster)
        assert torch.equal(x_b, x) and torch.equal(hh_b, hh) and torch.equal(q_b, out_qkv)


def test_chain_inside_cuda_graph(cuda_dev):
    """The decode step replays the chain from a CUDA graph: counters self-clean, results identical to the eager launch."""
    from crab_b200 import ops

    lin = Lin(cuda_dev, 41, 4096, [4096], lora=True, gamma=True)
    M = 32
    x = mk((M, 4096), cuda_dev, 42).to(torch.bfloat16)
    out = torch.empty((M, 4096), device=cuda_dev, dtype=torch.bfloat16)
    z = torch.zeros((32, 128), device=cuda_dev, dtype=torch.bfloat16)
    rstd = torch.zeros(32, device=cuda_dev, dtype=torch.float32)
    cnt = torch.zeros(288, dtype=torch.int32, device=cuda_dev)
    ph = [lin.phase(x, out, zbuf=z, rstd=rstd)]
    ops.decode_chain(ph, M, cnt, 4)
    torch.cuda.synchronize()
    eager = out.clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ops.decode_chain(ph, M, cnt, 4)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ops.decode_chain(ph, M, cnt, 4)
    for _ in range(3):
        out.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, eager)


# ---- the same fused linears as ONE crab_gemm_skinny_bf16 launch each (the default decode path) -----------------------------
def _flags(dev):
    return torch.zeros(64, dtype=torch.int32, device=dev)


@pytest.mark.parametrize("M", [32, 6])
def test_fused_skinny_qkv_like(cuda_dev, M):
    """RMSNorm as epilogue scale + three LoRA linears from the in-launch statistics cluster + bias (auto split = 2)."""
    from crab_b200 import ops

    lin = Lin(cuda_dev, 51, 4096, [4096, 4096, 4096], lora=True, gamma=True, bias=True, kext=96)
    x = mk((M, 4096), cuda_dev, 52, 3.0).to(torch.bfloat16)
    out = torch.empty((M, lin.N), device=cuda_dev, dtype=torch.bfloat16)
    z = torch.zeros((32, 128), device=cuda_dev, dtype=torch.bfloat16)
    rstd = torch.zeros(32, device=cuda_dev, dtype=torch.float32)
    fl = _flags(cuda_dev)
    for _ in range(2):   # second launch: the flags must have been left clean
        ops.gemm_skinny(x, lin.packed, bias=lin.bias, out=out, z=z, kext=96, stats=lin.stats, stats_linears=3, norm=True, eps=1e-6,
                        lora_scale=lin.scale, rstd=rstd, flags=fl)
        torch.cuda.synchronize()
        assert int(fl.abs().sum()) == 0
        assert torch.allclose(rstd[:M], torch.rsqrt(x.float().pow(2).mean(-1) + 1e-6), rtol=1e-5, atol=0)
        e = rel(out, lin.ref(x))
        assert e < 6e-3, e


@pytest.mark.parametrize("clusters", [1, 3, 8])
def test_fused_skinny_stats_clusters_and_flag_ring(cuda_dev, clusters):
    """The statistics item shared by several clusters (partials combined through the scratch buffer by the last cluster to arrive)
    gives the one-cluster result up to summation order, and is bit-reproducible; flag slots handed round a ring — every launch
    zeroes its predecessor's slot and leaves its own set."""
    from crab_b200 import ops

    M = 32
    lin = Lin(cuda_dev, 51, 4096, [4096, 4096, 4096], lora=True, gamma=True, bias=True, kext=96)
    x = mk((M, 4096), cuda_dev, 52, 3.0).to(torch.bfloat16)
    z = torch.zeros((32, 128), device=cuda_dev, dtype=torch.bfloat16)
    rstd = torch.zeros(32, device=cuda_dev, dtype=torch.float32)
    ring = torch.zeros((3, 32), device=cuda_dev, dtype=torch.int32)
    scratch = torch.full((8 * 36 * 32,), float("nan"), device=cuda_dev, dtype=torch.float32)
    outs = []
    for it in range(6):
        out = torch.empty((M, lin.N), device=cuda_dev, dtype=torch.bfloat16)
        k = it % 3
        ops.gemm_skinny(x, lin.packed, bias=lin.bias, out=out, z=z, kext=96, stats=lin.stats, stats_linears=3, norm=True, eps=1e-6,
                        lora_scale=lin.scale, rstd=rstd, flags=ring[k], flags_clear=ring[(k - 1) % 3], stats_scratch=scratch,
                        stats_clusters=clusters)
        torch.cuda.synchronize()
        assert int(ring[k, 0]) == 1 and int(ring[(k - 1) % 3, :2].abs().sum()) == 0, ring[:, :2]
        assert int(ring[k, 1]) == (clusters if clusters > 1 else 0)
        assert torch.allclose(rstd[:M], torch.rsqrt(x.float().pow(2).mean(-1) + 1e-6), rtol=1e-5, atol=0)
        e = rel(out, lin.ref(x))
        assert e < 6e-3, e
        outs.append(out)
    assert all(torch.equal(outs[0], o) for o in outs[1:])


def test_debug_trace_stamps_are_ordered(cuda_dev):
    """crab_debug_trace: every CTA of a traced launch stamps entry <= dependency wait <= first operand <= last MMA <= exit, the
    result is the untraced launch's bit for bit, and a disarmed library writes nothing."""
    import ctypes as C
    from crab_b200 import lib, ops

    lin = Lin(cuda_dev, 81, 4096, [4096], lora=False)
    x = mk((32, 4096), cuda_dev, 82, 1.0).to(torch.bfloat16)
    ref = ops.gemm_skinny(x, lin.packed, splits=4)
    L = lib.load()
    L.crab_debug_trace.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.crab_debug_trace.restype = C.c_int
    buf = torch.zeros(2 * 256 * 16, dtype=torch.int64, device=cuda_dev)
    assert L.crab_debug_trace(C.c_void_p(buf.data_ptr()), 2, 256) == 0
    try:
        out = ops.gemm_skinny(x, lin.packed, splits=4)
    finally:
        L.crab_debug_trace(None, 0, 0)
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    t = buf.view(2, 256, 16)[0, :128].cpu()          # 32 tiles x 4 ranks
    assert (t[:, 0] > 0).all() and (t[128:] == 0).all() if t.shape[0] > 128 else True
    for a_, b_ in ((0, 2), (2, 3), (3, 4), (4, 7)):
        assert (t[:, a_] <= t[:, b_]).all(), (a_, b_)
    snap = buf.clone()
    ops.gemm_skinny(x, lin.packed, splits=4)
    torch.cuda.synchronize()
    assert torch.equal(buf, snap)


def test_fused_skinny_gate_up_down_head(cuda_dev):
    """gate/up (SwiGLU, norm, 2 LoRA linears, no K split) -> down (K = 11008, 8-way split, 1 LoRA linear, in-place residual) ->
    final norm + lm_head (ragged N, fp32 out), each one launch, chained through bf16 buffers like the decode step."""
    from crab_b200 import ops

    dev = cuda_dev
    D, F = 4096, 11008
    gu = Lin(dev, 61, D, [F, F], lora=True, gamma=True, swiglu=True, kext=64)
    dn = Lin(dev, 62, F, [D], lora=True)
    hd = Lin(dev, 63, D, [32017], lora=False, gamma=True)
    for M in (32, 2):
        x0 = mk((M, D), dev, 64, 2.0).to(torch.bfloat16)
        x = x0.clone()
        hh = torch.empty((M, F), device=dev, dtype=torch.bfloat16)
        logits = torch.zeros((M, 32024), device=dev, dtype=torch.float32)
        zb = {k: torch.zeros((32, 128), device=dev, dtype=torch.bfloat16) for k in ("gu", "d")}
        rs = {k: torch.zeros(32, device=dev, dtype=torch.float32) for k in ("gu", "h")}
        ops.gemm_skinny(x, gu.packed, act=ops.ACT_SWIGLU, out=hh, z=zb["gu"], kext=64, stats=gu.stats, stats_linears=2, norm=True, eps=1e-6,
                        lora_scale=gu.scale, rstd=rs["gu"], flags=_flags(dev))
        ops.gemm_skinny(hh, dn.packed, residual=x, out=x, z=zb["d"], kext=32, stats=dn.stats, stats_linears=1, lora_scale=dn.scale,
                        flags=_flags(dev))
        ops.gemm_skinny(x, hd.packed, out=logits, n=32017, norm=True, eps=1e-6, rstd=rs["h"], flags=_flags(dev))
        torch.cuda.synchronize()
        h1 = gu.ref(x0).to(torch.bfloat16)
        x1 = dn.ref(h1, x0).to(torch.bfloat16)
        lg = hd.ref(x1)
        e = (rel(hh, h1), rel(x, x1), rel(logits[:, :32017], lg))
        print(f"fused skinny M {M}: h {e[0]:.3e} x {e[1]:.3e} logits {e[2]:.3e}")
        assert max(e) < 8e-3, e


def test_engine_decode_modes_agree(cuda_dev, monkeypatch):
    """The three decode-step organisations (row kernel + GEMM / one fused launch per linear / the persistent chain) run the same
    arithmetic: on the small golden case their teacher-forced logits agree to bf16 round-off."""
    from helpers import engine_cfg, load_golden
    from crab_b200.engine import CrabEngine

    g, case, sd, ocfg, ids, X = load_golden("llama_small")
    emb = g["inputs_embeds"].to(cuda_dev).to(torch.bfloat16)
    ref_ids = g["generated_ids"].to(cuda_dev)
    outs = {}
    for mode in ("rows", "skinny", "chain"):
        monkeypatch.setenv("CRAB_DECODE_MODE", mode)
        eng = CrabEngine(sd, engine_cfg(case, ocfg), cuda_dev, load_encoders=False)
        assert eng.decode_mode == mode
        _, logits = eng.generate_from_embeds(emb.clone(), ref_ids.shape[1], return_logits=True, teacher_tokens=ref_ids)
        outs[mode] = logits.float().cpu()
    e1, e2 = rel(outs["chain"], outs["skinny"]), rel(outs["rows"], outs["skinny"])
    print(f"decode modes: chain vs skinny logits rel_l2 {e1:.3e}; rows vs skinny {e2:.3e}")
    assert e1 < 3e-3      # measured 3.8e-4: same arithmetic, different split-K / reduction orders
    assert e2 < 1.2e-2    # the row-kernel path rounds the normalised row to bf16 (HF's rounding); the fused paths do not
