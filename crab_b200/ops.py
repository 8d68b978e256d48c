"""Tensor-level wrappers over the C ABI: take torch CUDA tensors, pass raw pointers + sizes + the current stream.

PyTorch is only the allocator / stream provider here; every op below is one call into libcrab_b200.so.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import lib as _l

_launches = 0


PDL_GEMM, PDL_ROW, PDL_ROPE, PDL_ATTN, PDL_LIGHT, PDL_ATTN_LATE = 1, 2, 4, 8, 16, 32
_pdl_mask = None


def set_pdl(mask: int) -> None:
    """Programmatic-dependent-launch mask over kernel classes (include/crab_b200.h: crab_set_pdl); read at launch time."""
    global _pdl_mask
    _l.check(_l.load().crab_set_pdl(C.c_int(int(mask) & 63)), "crab_set_pdl")
    _pdl_mask = int(mask) & 63


def set_gemm_2cta(mode: int) -> None:
    """CTA-pair GEMM policy: 0 never, 1 auto, 2 always when the 256-wide tile is used (include/crab_b200.h)."""
    _l.check(_l.load().crab_set_gemm_2cta(C.c_int(int(mode))), "crab_set_gemm_2cta")


def launch_count() -> int:
    """Number of crab_b200 CUDA kernels launched (or replayed from a captured graph) so far in this process."""
    return _launches


def count_launches(n: int) -> None:
    global _launches
    _launches += n


_timer = None  # when a list: every wrapper appends (tag, flops, bytes, start_event, end_event)


def start_kernel_timing() -> None:
    """Bracket every subsequent kernel launch with CUDA events on the launching stream (not during graph capture)."""
    global _timer
    _timer = []


def stop_kernel_timing():
    """-> {tag: dict(launches, ms, flops, bytes)} aggregated over the launches since start_kernel_timing()."""
    global _timer
    rec, _timer = _timer or [], None
    torch.cuda.synchronize()
    out = {}
    for tag, fl, by, e0, e1 in rec:
        d = out.setdefault(tag, dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
        d["launches"] += 1
        d["ms"] += e0.elapsed_time(e1)
        d["flops"] += fl
        d["bytes"] += by
    return out


class _timed:
    __slots__ = ("tag", "fl", "by", "e0")

    def __init__(self, tag, flops=0.0, nbytes=0.0):
        self.tag, self.fl, self.by, self.e0 = tag, flops, nbytes, None

    def __enter__(self):
        if _timer is not None and not torch.cuda.is_current_stream_capturing():
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.e0 is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _timer.append((self.tag, self.fl, self.by, self.e0, e1))
        return False


def auto_block_n(M: int, N: int, act: int, sms: int = 148) -> int:
    """Tile-width heuristic (same rule as the C side's block_n=0): widest tile that still gives every SM work."""
    if act == ACT_LORA_Z:
        return 64
    mt = (M + 127) // 128
    if mt * ((N + 255) // 256) >= sms:
        bn = 256
    elif mt * ((N + 127) // 128) >= sms or N < 128:
        bn = 64 if N <= 64 else 128
    else:
        bn = 64
    if act == ACT_SWIGLU and bn == 64:
        bn = 128
    return bn


ACT_NONE, ACT_QUICK_GELU, ACT_GELU, ACT_SWIGLU, ACT_LORA_Z = 0, 1, 2, 3, 4
BF16, F32 = 0, 1


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _l.CrabError("crab_b200 ops need CUDA tensors (there is no CPU path)")


def init(dev: int = 0) -> None:
    _l.check(_l.load().crab_init(C.c_int(dev)), "crab_init")


def gemm(a: torch.Tensor, w: torch.Tensor, *, bias: Optional[torch.Tensor] = None,
         residual: Optional[torch.Tensor] = None, res_scale: float = 1.0, out_scale: float = 1.0,
         act: int = ACT_NONE, out: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.bfloat16,
         block_n: int = 0, max_ctas: int = 0, k: Optional[int] = None, n: Optional[int] = None) -> torch.Tensor:
    """out[M, N'] = epilogue(a[M, K] @ w[N, K]^T).  `a`/`w` are 2-D bf16 with unit inner stride (row stride free)."""
    _req_cuda(a, w, bias, residual, out)
    assert a.dim() == 2 and w.dim() == 2 and a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    assert a.stride(1) == 1 and w.stride(1) == 1
    M = a.shape[0]
    K = k if k is not None else a.shape[1]
    N = n if n is not None else w.shape[0]
    assert w.shape[1] >= K and a.shape[1] >= K
    if act == ACT_SWIGLU:
        n_out = N // 2
    elif act == ACT_LORA_Z:
        n_out = (N // 11) * 24
    else:
        n_out = N
    if out is None:
        out = torch.empty((M, n_out), device=a.device, dtype=out_dtype)
    assert out.dim() == 2 and out.stride(1) == 1 and out.shape[0] == M and out.shape[1] >= n_out
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() >= N
    if residual is not None:
        assert residual.dtype == torch.bfloat16 and residual.stride(1) == 1
    if block_n == 0:
        block_n = auto_block_n(M, N, act)
    args = _l.GemmArgs(
        A=_ptr(a), B=_ptr(w), C=_ptr(out), bias=_ptr(bias), residual=_ptr(residual),
        M=M, N=N, K=K, lda=a.stride(0), ldb=w.stride(0), ldc=out.stride(0),
        ldr=(residual.stride(0) if residual is not None else 0),
        res_scale=res_scale, out_scale=out_scale, act=act,
        out_dtype=(BF16 if out.dtype == torch.bfloat16 else F32), block_n=block_n, max_ctas=max_ctas,
    )
    with _timed(f"gemm_bf16_tcgen05<{block_n}>", 2.0 * M * N * K, 2.0 * (M * K + N * K) + out.element_size() * M * n_out):
        _l.check(_l.load().crab_gemm_bf16(C.byref(args), _stream()), "crab_gemm_bf16")
    count_launches(1)
    return out


def _i(v):
    return C.c_int(int(v))


def _vp(t):
    return C.c_void_p(None if t is None else t.data_ptr())


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Row-wise LayerNorm of a 2-D bf16 view (row stride free); gamma/beta fp32."""
    _req_cuda(x, gamma, beta, out)
    assert x.dim() == 2 and x.stride(1) == 1 and x.dtype == torch.bfloat16 and gamma.dtype == torch.float32
    if out is None:
        out = torch.empty((x.shape[0], x.shape[1]), device=x.device, dtype=torch.bfloat16)
    with _timed("crab_layernorm"):
        _l.check(_l.load().crab_layernorm(_vp(x), _i(x.stride(0)), _vp(gamma), _vp(beta), _vp(out), _i(out.stride(0)),
                                          _i(x.shape[0]), _i(x.shape[1]), C.c_float(eps), _stream()), "crab_layernorm")
    count_launches(1)
    return out


def rmsnorm(x: torch.Tensor, gamma: torch.Tensor, eps: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req_cuda(x, gamma, out)
    assert x.dim() == 2 and x.stride(1) == 1 and x.dtype == torch.bfloat16 and gamma.dtype == torch.float32
    if out is None:
        out = torch.empty((x.shape[0], x.shape[1]), device=x.device, dtype=torch.bfloat16)
    with _timed("crab_rmsnorm"):
        _l.check(_l.load().crab_rmsnorm(_vp(x), _i(x.stride(0)), _vp(gamma), _vp(out), _i(out.stride(0)), _i(x.shape[0]),
                                        _i(x.shape[1]), C.c_float(eps), _stream()), "crab_rmsnorm")
    count_launches(1)
    return out


def rope_table(max_pos: int, head_dim: int, theta: float, device) -> torch.Tensor:
    t = torch.empty((max_pos, head_dim), device=device, dtype=torch.float32)
    _l.check(_l.load().crab_rope_table(_vp(t), _i(max_pos), _i(head_dim), C.c_double(theta), _stream()), "crab_rope_table")
    count_launches(1)
    return t


def rope_kv_append(qkv: torch.Tensor, cos_sin: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, B: int,
                   S: int, H: int, KVH: int, head_dim: int, past: int = 0, past_dev: Optional[torch.Tensor] = None):
    """qkv (B*S, >= (H+2KVH)*hd) bf16: rotate q in place, write rotated k and v into the caches (B,KVH,ctx_max,hd)."""
    _req_cuda(qkv, cos_sin, k_cache, v_cache, past_dev)
    assert qkv.dim() == 2 and qkv.stride(1) == 1 and k_cache.is_contiguous() and v_cache.is_contiguous()
    assert cos_sin.shape[1] == head_dim and k_cache.shape[1] == KVH and k_cache.shape[3] == head_dim
    ctx_max = k_cache.shape[2]
    assert past_dev is not None or cos_sin.shape[0] >= past + S
    with _timed("crab_rope_kv_append"):
        _l.check(_l.load().crab_rope_kv_append(_vp(qkv), _i(qkv.stride(0)), _vp(cos_sin), _vp(k_cache), _vp(v_cache), _i(B),
                                               _i(S), _i(H), _i(KVH), _i(head_dim), _i(ctx_max), _vp(past_dev), _i(past),
                                               _stream()), "crab_rope_kv_append")
    count_launches(1)


def flash_attn(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, *, B: int, H: int, KVH: int,
               Sq: int, Sk: int, head_dim: int, q_strides, k_strides, v_strides, o_strides, scale: float,
               causal: bool = False, gate: Optional[torch.Tensor] = None, bias_table: Optional[torch.Tensor] = None,
               sk_dev: Optional[torch.Tensor] = None):
    """Strides are (batch, row, head) in elements; q/k/v/out are any bf16 tensors whose data_ptr is the origin."""
    _req_cuda(q, k, v, out, gate, bias_table)
    a = _l.AttnArgs(q=q.data_ptr(), k=k.data_ptr(), v=v.data_ptr(), o=out.data_ptr(),
                    q_bs=q_strides[0], q_rs=q_strides[1], q_hs=q_strides[2],
                    k_bs=k_strides[0], k_rs=k_strides[1], k_hs=k_strides[2],
                    v_bs=v_strides[0], v_rs=v_strides[1], v_hs=v_strides[2],
                    o_bs=o_strides[0], o_rs=o_strides[1], o_hs=o_strides[2],
                    B=B, H=H, KVH=KVH, Sq=Sq, Sk=Sk, head_dim=head_dim, scale=scale, causal=int(causal),
                    gate=_ptr(gate), bias_table=_ptr(bias_table), sk_dev=_ptr(sk_dev))
    # algorithmic work: 4*Sq*Sk*hd per (b, h), the visible half under a causal mask; hd-128 problems with >= 128 queries and no
    # bias run on the tcgen05 / TMEM kernel (flash_tcgen05.cu), the rest on the mma.sync kernel
    tag = ("crab_flash_attn_tcgen05<128>" if (head_dim == 128 and gate is None and Sq >= 128 and sk_dev is None)
           else f"crab_flash_attn<{head_dim}>")
    fl = 4.0 * B * H * Sq * Sk * head_dim * ((Sk - Sq / 2.0) / Sk if causal else 1.0)
    with _timed(tag, fl, 2.0 * (B * H * Sq * head_dim * 2 + 2 * B * KVH * Sk * head_dim)):
        _l.check(_l.load().crab_flash_attn(C.byref(a), _stream()), "crab_flash_attn")
    count_launches(1)
    return out


def attn_decode(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, out: torch.Tensor, *, B: int, H: int,
                KVH: int, head_dim: int, scale: float, length: int = 0, len_dev: Optional[torch.Tensor] = None,
                nsplit: int = 1, workspace: Optional[torch.Tensor] = None):
    _req_cuda(q, k_cache, v_cache, out, len_dev, workspace)
    assert q.dim() == 2 and out.dim() == 2 and q.stride(1) == 1 and out.stride(1) == 1
    ctx_max = k_cache.shape[2]
    if nsplit > 1 and workspace is None:
        workspace = torch.empty(B * H * nsplit * (head_dim + 2), device=q.device, dtype=torch.float32)
    with _timed("crab_attn_decode"):
        _l.check(_l.load().crab_attn_decode(_vp(q), _i(q.stride(0)), _vp(k_cache), _vp(v_cache), _vp(out), _i(out.stride(0)),
                                            _vp(workspace), _i(B), _i(H), _i(KVH), _i(head_dim), _i(ctx_max), _i(nsplit),
                                            _vp(len_dev), _i(length), C.c_float(scale), _stream()), "crab_attn_decode")
    count_launches(2 if nsplit > 1 else 1)
    return out


def attn_decode_fused(qkv: torch.Tensor, rope: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, out: torch.Tensor, *,
                      B: int, H: int, KVH: int, head_dim: int, scale: float, past_dev: torch.Tensor, nsplit: int = 1,
                      workspace: Optional[torch.Tensor] = None, ra: Optional[torch.Tensor] = None,
                      z: Optional[torch.Tensor] = None, lora_scale: float = 0.0, lora_ws: Optional[torch.Tensor] = None,
                      lora_counters: Optional[torch.Tensor] = None, gqa_tc: bool = False):
    """Decode-step attention with RoPE + KV append (+ the o_proj hyper-LoRA pre-pass) fused in: `qkv` is the RAW output
    of the qkv projection; the caches are appended at *past_dev; z (24 columns) is written when `ra` is given."""
    _req_cuda(qkv, rope, k_cache, v_cache, out, past_dev, workspace, ra, z, lora_ws, lora_counters)
    assert qkv.dim() == 2 and out.dim() == 2 and qkv.stride(1) == 1 and out.stride(1) == 1
    assert past_dev.dtype == torch.int32 and rope.dtype == torch.float32
    if nsplit > 1 and workspace is None:
        workspace = torch.empty(B * H * nsplit * (head_dim + 2), device=qkv.device, dtype=torch.float32)
    if ra is not None:
        assert z is not None and lora_ws is not None and lora_counters is not None
        assert lora_ws.dtype == torch.float32 and lora_ws.numel() >= B * (KVH if nsplit == 1 else H) * 11
        assert lora_counters.dtype == torch.int32 and lora_counters.numel() >= B and ra.stride(1) == 1 and z.stride(1) == 1
    a = _l.DecodeFusedArgs(
        qkv=qkv.data_ptr(), ldq=qkv.stride(0), cos_sin=rope.data_ptr(), k_cache=k_cache.data_ptr(), v_cache=v_cache.data_ptr(),
        o=out.data_ptr(), ldo=out.stride(0), workspace=_ptr(workspace), B=B, H=H, KVH=KVH, head_dim=head_dim,
        ctx_max=k_cache.shape[2], nsplit=nsplit, past_dev=past_dev.data_ptr(), scale=scale,
        lora_ra=_ptr(ra), ld_ra=ra.stride(0) if ra is not None else 0, lora_z=_ptr(z), ld_z=z.stride(0) if z is not None else 0,
        lora_scale=lora_scale, lora_ws=_ptr(lora_ws), lora_counters=_ptr(lora_counters), gqa_tensor_cores=1 if gqa_tc else 0, reserved0=0)
    with _timed("crab_attn_decode_fused"):
        _l.check(_l.load().crab_attn_decode_fused(C.byref(a), _stream()), "crab_attn_decode_fused")
    clustered = nsplit > 1 and H == KVH and head_dim == 128 and not gqa_tc and os.environ.get("CRAB_ATTN_CLUSTER", "1") != "0"
    count_launches((2 if (nsplit > 1 and not clustered) else 1) + (1 if gqa_tc else 0))   # cluster split: no combine launch
    return out


def gather_rows(src: torch.Tensor, dst: torch.Tensor, n: int, cols: int, src_rows: Optional[torch.Tensor] = None,
                dst_rows: Optional[torch.Tensor] = None):
    _req_cuda(src, dst, src_rows, dst_rows)
    assert src.dtype == torch.bfloat16 and dst.dtype == torch.bfloat16 and src.stride(-1) == 1 and dst.stride(-1) == 1
    for r in (src_rows, dst_rows):
        assert r is None or (r.dtype == torch.int64 and r.is_contiguous() and r.numel() >= n)
    with _timed("crab_gather_rows"):
        _l.check(_l.load().crab_gather_rows(_vp(src), _i(src.stride(-2)), _vp(src_rows), _vp(dst), _i(dst.stride(-2)),
                                            _vp(dst_rows), _i(n), _i(cols), _stream()), "crab_gather_rows")
    count_launches(1)


def cast_bf16(src: torch.Tensor) -> torch.Tensor:
    _req_cuda(src)
    assert src.dtype == torch.float32 and src.is_contiguous()
    out = torch.empty(src.shape, device=src.device, dtype=torch.bfloat16)
    with _timed("crab_cast_f32_bf16"):
        _l.check(_l.load().crab_cast_f32_bf16(_vp(src), _vp(out), C.c_int64(src.numel()), _stream()), "crab_cast_f32_bf16")
    count_launches(1)
    return out


def patchify(images: torch.Tensor, patch: int, ld_out: int) -> torch.Tensor:
    """fp32 (n, C, H, W) -> bf16 (n * (H//p) * (W//p), ld_out) patch rows, zero-padded beyond C*p*p."""
    _req_cuda(images)
    assert images.dtype == torch.float32 and images.is_contiguous() and images.dim() == 4
    n, c, h, w = images.shape
    rows = n * (h // patch) * (w // patch)
    out = torch.empty((rows, ld_out), device=images.device, dtype=torch.bfloat16)
    with _timed("crab_patchify"):
        _l.check(_l.load().crab_patchify(_vp(images), _vp(out), _i(ld_out), _i(n), _i(c), _i(h), _i(w), _i(patch), _stream()),
                 "crab_patchify")
    count_launches(1)
    return out


def patchify_u8(images: torch.Tensor, patch: int, ld_out: int, mean, std, rescale: float = 1.0 / 255.0) -> torch.Tensor:
    """uint8 (n, H, W, 3) frames -> bf16 patch rows of ((u * rescale - mean) / std), same row/column order as `patchify`."""
    _req_cuda(images)
    if images.dtype != torch.uint8 or images.dim() != 4 or images.shape[-1] != 3 or not images.is_contiguous():
        raise _l.CrabError("patchify_u8 wants a contiguous uint8 (n, H, W, 3) tensor")
    n, h, w, _ = images.shape
    rows = n * (h // patch) * (w // patch)
    out = torch.empty((rows, ld_out), device=images.device, dtype=torch.bfloat16)
    m3, s3 = (C.c_float * 3)(*map(float, mean)), (C.c_float * 3)(*map(float, std))
    with _timed("crab_patchify_u8"):
        _l.check(_l.load().crab_patchify_u8(_vp(images), _vp(out), _i(ld_out), _i(n), _i(h), _i(w), _i(patch), m3, s3,
                                            C.c_float(rescale), _stream()), "crab_patchify_u8")
    count_launches(1)
    return out


def normalize_u8(images: torch.Tensor, mean, std, rescale: float = 1.0 / 255.0) -> torch.Tensor:
    """uint8 (n, H, W, 3) -> fp32 (n, 3, H, W) `pixel_values` ((u * rescale - mean) / std)."""
    _req_cuda(images)
    if images.dtype != torch.uint8 or images.dim() != 4 or images.shape[-1] != 3 or not images.is_contiguous():
        raise _l.CrabError("normalize_u8 wants a contiguous uint8 (n, H, W, 3) tensor")
    n, h, w, _ = images.shape
    out = torch.empty((n, 3, h, w), device=images.device, dtype=torch.float32)
    m3, s3 = (C.c_float * 3)(*map(float, mean)), (C.c_float * 3)(*map(float, std))
    with _timed("crab_normalize_u8"):
        _l.check(_l.load().crab_normalize_u8(_vp(images), _vp(out), _i(n), _i(h), _i(w), m3, s3, C.c_float(rescale), _stream()),
                 "crab_normalize_u8")
    count_launches(1)
    return out


def resample_u8(images: torch.Tensor, out_size: int, axis: int, bounds: torch.Tensor, kk: torch.Tensor, out0: int) -> torch.Tensor:
    """One pass of Pillow's 8-bit resampler on uint8 (n, H, W, 3): axis 1 -> (n, H, out_size, 3), axis 0 -> (n, out_size, W, 3);
    output index j uses coefficient row out0 + j (centre crop folded in)."""
    _req_cuda(images, bounds, kk)
    if images.dtype != torch.uint8 or images.dim() != 4 or images.shape[-1] != 3 or not images.is_contiguous():
        raise _l.CrabError("resample_u8 wants a contiguous uint8 (n, H, W, 3) tensor")
    assert bounds.dtype == torch.int32 and kk.dtype == torch.int32 and bounds.is_contiguous() and kk.is_contiguous()
    n, h, w, _ = images.shape
    oh, ow = (h, out_size) if axis == 1 else (out_size, w)
    assert out0 + out_size <= bounds.shape[0]
    out = torch.empty((n, oh, ow, 3), device=images.device, dtype=torch.uint8)
    with _timed("crab_resample_u8"):
        _l.check(_l.load().crab_resample_u8(_vp(images), _vp(out), _i(n), _i(h), _i(w), _i(oh), _i(ow), _i(axis), _vp(bounds),
                                            _vp(kk), _i(kk.shape[1]), _i(out0), _stream()), "crab_resample_u8")
    count_launches(1)
    return out


def fbank_num_frames(n_samples: int) -> int:
    n = C.c_int(0)
    _l.check(_l.load().crab_fbank_num_frames(_i(n_samples), C.byref(n)), "crab_fbank_num_frames")
    return n.value


def kaldi_fbank(wave: torch.Tensor, window: torch.Tensor, twiddle: torch.Tensor, mel_start: torch.Tensor, mel_off: torch.Tensor,
                mel_w: torch.Tensor, in_scale: float, mean: float, std2: float) -> torch.Tensor:
    """fp32 waveforms (n_seg, n_samples) in [-1, 1] -> fp32 (n_seg, n_frames, n_mel) normalised Kaldi log-mel features."""
    _req_cuda(wave, window, twiddle, mel_start, mel_off, mel_w)
    assert twiddle.dtype == torch.float32 and tuple(twiddle.shape) == (256, 2) and twiddle.is_contiguous()
    if wave.dtype != torch.float32 or wave.dim() != 2 or wave.stride(1) != 1:
        raise _l.CrabError("kaldi_fbank wants fp32 (n_seg, n_samples) waveforms with unit inner stride")
    assert window.dtype == torch.float32 and mel_w.dtype == torch.float32 and mel_start.dtype == torch.int32 and mel_off.dtype == torch.int32
    n_seg, n_samples = wave.shape
    n_mel = mel_start.numel()
    frames = fbank_num_frames(n_samples)
    out = torch.empty((n_seg, frames, n_mel), device=wave.device, dtype=torch.float32)
    with _timed("crab_kaldi_fbank"):
        _l.check(_l.load().crab_kaldi_fbank(_vp(wave), C.c_int64(wave.stride(0)), _i(n_seg), _i(n_samples), _vp(window),
                                            _vp(twiddle), _vp(mel_start), _vp(mel_off), _vp(mel_w), _i(n_mel), C.c_float(in_scale),
                                            C.c_float(mean), C.c_float(std2), _vp(out), _stream()), "crab_kaldi_fbank")
    count_launches(1)
    return out


def clip_embed_ln(patch_emb, cls, pos, gamma, beta, n_img: int, tokens: int, D: int, eps: float) -> torch.Tensor:
    _req_cuda(patch_emb, cls, pos, gamma, beta)
    assert patch_emb.is_contiguous() and patch_emb.dtype == torch.bfloat16
    out = torch.empty((n_img * tokens, D), device=patch_emb.device, dtype=torch.bfloat16)
    with _timed("crab_clip_embed_ln"):
        _l.check(_l.load().crab_clip_embed_ln(_vp(patch_emb), _vp(cls), _vp(pos), _vp(gamma), _vp(beta), _vp(out), _i(n_img),
                                              _i(tokens), _i(D), C.c_float(eps), _stream()), "crab_clip_embed_ln")
    count_launches(1)
    return out


def beats_gate(q: torch.Tensor, grep_w, grep_b, grep_a, B: int, T: int, H: int) -> torch.Tensor:
    _req_cuda(q, grep_w, grep_b, grep_a)
    gate = torch.empty((B, H, T), device=q.device, dtype=torch.float32)
    with _timed("crab_beats_gate"):
        _l.check(_l.load().crab_beats_gate(_vp(q), _i(q.stride(0)), _vp(grep_w), _vp(grep_b), _vp(grep_a), _vp(gate), _i(B),
                                           _i(T), _i(H), _stream()), "crab_beats_gate")
    count_launches(1)
    return gate


def beats_group_pack(x: torch.Tensor, B: int, T: int, Cc: int, G: int) -> torch.Tensor:
    _req_cuda(x)
    assert x.is_contiguous() and x.dtype == torch.bfloat16
    out = torch.empty((G, B, T * (Cc // G)), device=x.device, dtype=torch.bfloat16)
    with _timed("crab_beats_group_pack"):
        _l.check(_l.load().crab_beats_group_pack(_vp(x), _vp(out), _i(B), _i(T), _i(Cc), _i(G), _stream()), "crab_beats_group_pack")
    count_launches(1)
    return out


def beats_posconv_finish(x: torch.Tensor, conv_g: torch.Tensor, bias: torch.Tensor, B: int, T: int, Cc: int, G: int):
    _req_cuda(x, conv_g, bias)
    assert x.is_contiguous() and conv_g.is_contiguous()
    out = torch.empty((B * T, Cc), device=x.device, dtype=torch.bfloat16)
    with _timed("crab_beats_posconv_finish"):
        _l.check(_l.load().crab_beats_posconv_finish(_vp(x), _vp(conv_g), _vp(bias), _vp(out), _i(B), _i(T), _i(Cc), _i(G),
                                                     _stream()), "crab_beats_posconv_finish")
    count_launches(1)
    return out


def argmax(logits: torch.Tensor, V: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req_cuda(logits, out)
    assert logits.dtype == torch.float32 and logits.dim() == 2 and logits.stride(1) == 1
    if out is None:
        out = torch.empty((logits.shape[0],), device=logits.device, dtype=torch.int64)
    with _timed("crab_argmax"):
        _l.check(_l.load().crab_argmax(_vp(logits), _i(logits.stride(0)), _i(logits.shape[0]), _i(V), _vp(out), _stream()),
                 "crab_argmax")
    count_launches(1)
    return out


def sample_top_k_top_p(logits: torch.Tensor, V: int, u: torch.Tensor, *, temperature: float = 1.0, top_k: int = 0, top_p: float = 1.0,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ids[row] drawn from softmax(top-p(top-k(logits[row, :V] / temperature))) by inverse CDF with u[row] in [0, 1)."""
    _req_cuda(logits, u, out)
    assert logits.dtype == torch.float32 and u.dtype == torch.float32 and logits.stride(1) == 1
    rows = logits.shape[0]
    if out is None:
        out = torch.empty(rows, device=logits.device, dtype=torch.int64)
    with _timed("crab_sample_top_k_top_p"):
        _l.check(_l.load().crab_sample_top_k_top_p(_vp(logits), _i(logits.stride(0)), _i(rows), _i(V), C.c_float(temperature), _i(top_k),
                                                   C.c_float(top_p), _vp(u), _vp(out), _stream()), "crab_sample_top_k_top_p")
    count_launches(1)
    return out


def cross_entropy(logits: torch.Tensor, V: int, labels: torch.Tensor) -> torch.Tensor:
    """Per-row loss (fp32) of fp32 logits [rows, >= V] against int64 labels [rows]; rows with a negative label give 0."""
    _req_cuda(logits, labels)
    assert logits.dtype == torch.float32 and labels.dtype == torch.int64 and logits.stride(1) == 1
    rows = logits.shape[0]
    loss = torch.empty(rows, device=logits.device, dtype=torch.float32)
    with _timed("crab_cross_entropy"):
        _l.check(_l.load().crab_cross_entropy(_vp(logits), _i(logits.stride(0)), _i(rows), _i(V), _vp(labels), _vp(loss), _stream()),
                 "crab_cross_entropy")
    count_launches(1)
    return loss


def add_scalar_i32(p: torch.Tensor, v: int):
    _req_cuda(p)
    assert p.dtype == torch.int32
    with _timed("crab_add_scalar_i32"):
        _l.check(_l.load().crab_add_scalar_i32(_vp(p), _i(v), _stream()), "crab_add_scalar_i32")
    count_launches(1)


# ---- decode-step kernels ------------------------------------------------------------------------------------------
class PackedWeight:
    """A decode weight in the streaming layout of crab_pack_skinny_weight (plus its logical shape)."""

    __slots__ = ("data", "N", "K", "swiglu")

    def __init__(self, data: torch.Tensor, N: int, K: int, swiglu: bool = False):
        self.data, self.N, self.K, self.swiglu = data, N, K, swiglu


def pack_skinny_weight(w: torch.Tensor, k: Optional[int] = None, swiglu: bool = False) -> PackedWeight:
    """Row-major bf16 [N, >=K] -> contiguous pre-swizzled 16 KB (tile, k-block) blocks for gemm_skinny.
    swiglu=True: `w` is in the prefill layout ([64 gate | 64 up] row groups); rows are interleaved for the decode kernel."""
    _req_cuda(w)
    assert w.dim() == 2 and w.dtype == torch.bfloat16 and w.stride(1) == 1
    N, K = w.shape[0], (k if k is not None else w.shape[1])
    nbytes = C.c_int64(0)
    _l.check(_l.load().crab_skinny_packed_bytes(_i(N), _i(K), C.byref(nbytes)), "crab_skinny_packed_bytes")
    out = torch.empty(nbytes.value // 2, device=w.device, dtype=torch.bfloat16)
    assert out.data_ptr() % 128 == 0
    _l.check(_l.load().crab_pack_skinny_weight(_vp(w), _i(N), _i(K), _i(w.stride(0)), _vp(out), _i(1 if swiglu else 0), _stream()),
             "crab_pack_skinny_weight")
    count_launches(1)
    return PackedWeight(out, N, K, swiglu)


def gemm_skinny(x: torch.Tensor, w, *, bias: Optional[torch.Tensor] = None,
                residual: Optional[torch.Tensor] = None, act: int = ACT_NONE, out: Optional[torch.Tensor] = None,
                out_dtype: torch.dtype = torch.bfloat16, k: Optional[int] = None, n: Optional[int] = None,
                splits: int = 0, z: Optional[torch.Tensor] = None, kext: int = 0, stats: Optional[torch.Tensor] = None,
                stats_linears: int = 0, norm: bool = False, eps: float = 0.0, lora_scale: float = 1.0,
                rstd: Optional[torch.Tensor] = None, flags: Optional[torch.Tensor] = None, tag: str = "gemm_skinny_tcgen05",
                prefetch: Optional[torch.Tensor] = None, prefetch_bytes: int = 0, stats_scratch: Optional[torch.Tensor] = None,
                flags_clear: Optional[torch.Tensor] = None, stats_clusters: int = 0) -> torch.Tensor:
    """out[M, N'] = epilogue(x[M<=32, K] @ w[N, K]^T): the decode-step weight-streaming GEMM (swap-AB, split-K).
    With z / kext the K-extension columns come from a separate buffer; norm / stats_linears fold the RMSNorm (as an epilogue scale
    over a gamma-folded weight) and the hyper-LoRA router / A pre-pass into the same launch (see include/crab_b200.h)."""
    packed = isinstance(w, PackedWeight)
    _req_cuda(x, w.data if packed else w, bias, residual, out, z, stats, rstd, flags, stats_scratch, flags_clear)
    assert stats_scratch is None or (stats_scratch.dtype == torch.float32 and stats_scratch.numel() >= 8 * 36 * 32)
    assert x.dim() == 2 and x.dtype == torch.bfloat16
    M = x.shape[0]
    if packed:
        K, N = w.K - kext, w.N
        assert x.shape[1] >= K
    else:
        assert w.dim() == 2 and w.dtype == torch.bfloat16 and kext == 0
        K = k if k is not None else x.shape[1]
        N = n if n is not None else w.shape[0]
    if n is not None:
        N = n
    n_out = N // 2 if act == ACT_SWIGLU else N
    if out is None:
        out = torch.empty((M, n_out), device=x.device, dtype=out_dtype)
    if act == ACT_SWIGLU:
        assert packed and w.swiglu, "decode SwiGLU needs pack_skinny_weight(..., swiglu=True)"
    args = _l.SkinnyArgs(X=_ptr(x), W=None if packed else _ptr(w), W_packed=_ptr(w.data) if packed else None, C=_ptr(out),
                         bias=_ptr(bias), residual=_ptr(residual),
                         M=M, N=N, K=K, ldx=x.stride(0), ldw=0 if packed else w.stride(0), ldc=out.stride(0),
                         ldr=(residual.stride(0) if residual is not None else 0), act=act,
                         out_dtype=(BF16 if out.dtype == torch.bfloat16 else F32), splits=splits,
                         Z=_ptr(z), ldz=(z.stride(0) if z is not None else 0), Kext=kext, stats_packed=_ptr(stats),
                         stats_linears=stats_linears, norm=1 if norm else 0, eps=eps, lora_scale=lora_scale, rstd=_ptr(rstd),
                         flags=_ptr(flags), stats_scratch=_ptr(stats_scratch), flags_clear=_ptr(flags_clear),
                         stats_clusters=stats_clusters, reserved0=0, prefetch=_ptr(prefetch),
                         prefetch_bytes=(min(prefetch_bytes or prefetch.numel() * prefetch.element_size(), prefetch.numel() * prefetch.element_size())
                                         if prefetch is not None else 0))
    wb = 2.0 * N * (K + kext) + (stats.numel() * 2.0 if stats is not None else 0.0)
    with _timed(tag, 2.0 * M * N * (K + kext), 2.0 * M * K + wb + out.element_size() * M * n_out):
        _l.check(_l.load().crab_gemm_skinny_bf16(C.byref(args), _stream()), "crab_gemm_skinny_bf16")
    count_launches(1)
    return out


def row_norm_loraz(x: torch.Tensor, *, gamma: Optional[torch.Tensor] = None, eps: float = 0.0,
                   y: Optional[torch.Tensor] = None, ra: Optional[torch.Tensor] = None, groups: int = 0,
                   z: Optional[torch.Tensor] = None, scale: float = 1.0):
    """Per-row (decode) RMSNorm (optional) + hyper-LoRA router/A pre-pass writing the 24*groups z columns."""
    _req_cuda(x, gamma, y, ra, z)
    assert x.dim() == 2 and x.stride(1) == 1
    rows, cols = x.shape
    with _timed("crab_row_norm_loraz"):
        _l.check(_l.load().crab_row_norm_loraz(_vp(x), _i(x.stride(0)), _vp(gamma), C.c_float(eps), _vp(y),
                                               _i(y.stride(0) if y is not None else 0), _vp(ra),
                                               _i(ra.stride(0) if ra is not None else 0), _i(groups), _vp(z),
                                               _i(z.stride(0) if z is not None else 0), C.c_float(scale), _i(rows), _i(cols),
                                               _stream()), "crab_row_norm_loraz")
    count_launches(1)


# ---- decode-step GEMM chain (csrc/decode_chain.cu) ---------------------------------------------------------------------
def pack_chain_stats(ra: torch.Tensor, gamma: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Router/A rows [11 * linears, K] (rows of linear g: 3 lora_route rows, then 8 lora_A rows) -> the stats stream of
    crab_decode_chain: [K / 64] blocks of 40 x 64 bf16 (rows past 11 * linears zero), each block pre-swizzled exactly like a
    weight tile (16-byte chunk c of row r stored at c ^ (r & 7)).  `gamma` (RMSNorm scale) is folded into the columns.
    Pure data movement at load time."""
    _req_cuda(ra, gamma)
    rows, K = ra.shape
    assert rows <= 33 and K % 64 == 0
    dense = torch.zeros((40, K), device=ra.device, dtype=torch.float32)
    dense[:rows] = ra.float() if gamma is None else ra.float() * gamma.float()[None, :]
    t = dense.to(torch.bfloat16).view(40, K // 64, 8, 8).permute(1, 0, 2, 3)                 # [kb, row, chunk, 8]
    r = torch.arange(40, device=ra.device)
    src_chunk = torch.arange(8, device=ra.device)[None, :] ^ (r[:, None] & 7)              # stored chunk cs holds source cs ^ (r & 7)
    out = torch.gather(t, 2, src_chunk[None, :, :, None].expand(K // 64, 40, 8, 8)).contiguous()
    assert out.data_ptr() % 128 == 0
    return out.view(-1)


class ChainPhase:
    """One linear of a crab_decode_chain launch (see include/crab_b200.h).  Tensors are kept alive by the caller."""

    def __init__(self, x: torch.Tensor, w: "PackedWeight", out: torch.Tensor, *, k: int, z: Optional[torch.Tensor] = None, kext: int = 0,
                 stats: Optional[torch.Tensor] = None, stats_linears: int = 0, norm: bool = False, eps: float = 0.0,
                 lora_scale: float = 1.0, rstd: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
                 residual: Optional[torch.Tensor] = None, act: int = ACT_NONE, n: Optional[int] = None):
        _req_cuda(x, w.data, out, z, stats, rstd, bias, residual)
        assert x.dtype == torch.bfloat16 and x.stride(1) == 1 and x.shape[1] >= k
        assert w.K == k + kext, (w.K, k, kext)
        self.x, self.w, self.out, self.k, self.z, self.kext = x, w, out, k, z, kext
        self.stats, self.stats_linears, self.norm, self.eps, self.lora_scale = stats, stats_linears, norm, eps, lora_scale
        self.rstd, self.bias, self.residual, self.act, self.n = rstd, bias, residual, act, (n if n is not None else w.N)

    def fill(self, p: "_l.ChainPhase"):
        p.X, p.ldx, p.K = _ptr(self.x), self.x.stride(0), self.k
        p.Z, p.ldz, p.Kext = _ptr(self.z), (self.z.stride(0) if self.z is not None else 0), self.kext
        p.W_packed = _ptr(self.w.data)
        p.stats_packed, p.stats_linears, p.norm = _ptr(self.stats), self.stats_linears, 1 if self.norm else 0
        p.eps, p.lora_scale = self.eps, self.lora_scale
        p.rstd = _ptr(self.rstd)
        p.C, p.ldc, p.N = _ptr(self.out), self.out.stride(0), self.n
        p.out_dtype = BF16 if self.out.dtype == torch.bfloat16 else F32
        p.act = self.act
        p.bias = _ptr(self.bias)
        p.residual, p.ldr = _ptr(self.residual), (self.residual.stride(0) if self.residual is not None else 0)

    def weight_bytes(self) -> float:
        return 2.0 * self.w.N * self.w.K + (self.stats.numel() * 2.0 if self.stats is not None else 0.0)


def decode_chain(phases, M: int, counters: torch.Tensor, cluster: int = 0, max_clusters: int = 0, tag: str = "crab_decode_chain"):
    """One persistent launch over up to four dependent decode-step linears (M <= 32 rows)."""
    assert 1 <= len(phases) <= 4 and counters.dtype == torch.int32 and counters.numel() >= 288
    _req_cuda(counters)
    a = _l.ChainArgs()
    for i, ph in enumerate(phases):
        ph.fill(a.phase[i])
    a.n_phases, a.M, a.cluster, a.max_clusters, a.counters = len(phases), M, cluster, max_clusters, _ptr(counters)
    nbytes = sum(ph.weight_bytes() for ph in phases)
    with _timed(tag, sum(2.0 * M * ph.w.N * ph.w.K for ph in phases), nbytes):
        _l.check(_l.load().crab_decode_chain(C.byref(a), _stream()), "crab_decode_chain")
    count_launches(1)


def decode_chain_max_clusters(cluster: int) -> int:
    n = C.c_int(0)
    _l.check(_l.load().crab_decode_chain_max_clusters(_i(cluster), C.byref(n)), "crab_decode_chain_max_clusters")
    return n.value


# ---- segmentation-head helpers (csrc/seg.cu) -------------------------------------------------------------------------
EW_ADD, EW_RELU, EW_GELU, EW_GATE = 0, 1, 2, 3


def small_attn(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, heads: int, head_dim: int) -> torch.Tensor:
    """softmax(q k^T / sqrt(hd)) v for bf16 [N, heads*hd] views (row strides free, unit inner stride), hd in {16, 32}."""
    _req_cuda(q, k, v, out)
    for t in (q, k, v, out):
        assert t.dim() == 2 and t.stride(1) == 1 and t.dtype == torch.bfloat16 and t.shape[1] >= heads * head_dim
    assert k.shape[0] == v.shape[0] and out.shape[0] == q.shape[0]
    with _timed("crab_small_attn"):
        _l.check(_l.load().crab_small_attn(_vp(q), _i(q.stride(0)), _vp(k), _i(k.stride(0)), _vp(v), _i(v.stride(0)), _vp(out),
                                           _i(out.stride(0)), _i(q.shape[0]), _i(k.shape[0]), _i(heads), _i(head_dim),
                                           C.c_float(head_dim ** -0.5), _stream()), "crab_small_attn")
    count_launches(1)
    return out


def elementwise(a: torch.Tensor, op: int, b: Optional[torch.Tensor] = None, gate: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bf16 [R, C]: EW_ADD (b is [R, C] or [1, C]), EW_RELU, EW_GELU, EW_GATE ((sigmoid(gate[r]) + 1) * a, gate fp32 [R])."""
    _req_cuda(a, b, gate, out)
    assert a.dim() == 2 and a.stride(1) == 1 and a.dtype == torch.bfloat16
    if out is None:
        out = torch.empty((a.shape[0], a.shape[1]), device=a.device, dtype=torch.bfloat16)
    if b is not None:
        assert b.dim() == 2 and b.stride(1) == 1 and b.dtype == torch.bfloat16 and b.shape[1] >= a.shape[1] and b.shape[0] in (1, a.shape[0])
    if gate is not None:
        assert gate.dtype == torch.float32 and gate.is_contiguous() and gate.numel() >= a.shape[0]
    with _timed("crab_elementwise"):
        _l.check(_l.load().crab_elementwise(_vp(a), _i(a.stride(0)), _vp(b), _i(b.stride(0) if b is not None else 0),
                                            _i(b.shape[0] if b is not None else 0), _vp(gate), _vp(out), _i(out.stride(0)),
                                            _i(a.shape[0]), _i(a.shape[1]), _i(op), _stream()), "crab_elementwise")
    count_launches(1)
    return out


def row_mean_f32(x: torch.Tensor, cols: int) -> torch.Tensor:
    _req_cuda(x)
    assert x.dim() == 2 and x.stride(1) == 1 and x.dtype == torch.float32 and cols <= x.shape[1]
    out = torch.empty((x.shape[0],), device=x.device, dtype=torch.float32)
    with _timed("crab_row_mean_f32"):
        _l.check(_l.load().crab_row_mean_f32(_vp(x), _i(x.stride(0)), _i(x.shape[0]), _i(cols), _vp(out), _stream()),
                 "crab_row_mean_f32")
    count_launches(1)
    return out


def im2col3x3(x: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """token-major bf16 [h*w, C] -> [h*w, 9*C] (columns (ky, kx, c), zero padding)."""
    _req_cuda(x)
    assert x.dim() == 2 and x.stride(1) == 1 and x.dtype == torch.bfloat16 and x.shape[0] == h * w
    c = x.shape[1]
    out = torch.empty((h * w, 9 * c), device=x.device, dtype=torch.bfloat16)
    with _timed("crab_im2col3x3"):
        _l.check(_l.load().crab_im2col3x3(_vp(x), _i(x.stride(0)), _vp(out), _i(h), _i(w), _i(c), _stream()), "crab_im2col3x3")
    count_launches(1)
    return out


def bilinear_f32(x: torch.Tensor, hin: int, win: int, hout: int, wout: int, channels: int, out: Optional[torch.Tensor] = None,
                 alpha: float = 1.0, beta: float = 0.0, nchw_out: bool = False) -> torch.Tensor:
    """token-major fp32 [hin*win, >=channels] -> [hout*wout, channels] (or [channels, hout, wout]); out = beta*out + alpha*interp."""
    _req_cuda(x, out)
    assert x.dim() == 2 and x.stride(1) == 1 and x.dtype == torch.float32 and x.shape[0] == hin * win and x.shape[1] >= channels
    if out is None:
        assert beta == 0.0
        out = torch.empty((channels, hout, wout) if nchw_out else (hout * wout, channels), device=x.device, dtype=torch.float32)
    assert out.dtype == torch.float32 and out.is_contiguous()
    ldo = wout if nchw_out else out.shape[1]
    with _timed("crab_bilinear_f32"):
        _l.check(_l.load().crab_bilinear_f32(_vp(x), _i(x.stride(0)), _i(hin), _i(win), _vp(out), _i(ldo), _i(hout), _i(wout),
                                             _i(channels), C.c_float(alpha), C.c_float(beta), _i(1 if nchw_out else 0), _stream()),
                 "crab_bilinear_f32")
    count_launches(1)
    return out
