"""Tensor-level wrappers over the C ABI: take torch CUDA tensors, pass raw pointers + sizes + the current stream.

PyTorch is only the allocator / stream provider here; every op below is one call into libcrab_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import lib as _l

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
    args = _l.GemmArgs(
        A=_ptr(a), B=_ptr(w), C=_ptr(out), bias=_ptr(bias), residual=_ptr(residual),
        M=M, N=N, K=K, lda=a.stride(0), ldb=w.stride(0), ldc=out.stride(0),
        ldr=(residual.stride(0) if residual is not None else 0),
        res_scale=res_scale, out_scale=out_scale, act=act,
        out_dtype=(BF16 if out.dtype == torch.bfloat16 else F32), block_n=block_n, max_ctas=max_ctas,
    )
    _l.check(_l.load().crab_gemm_bf16(C.byref(args), _stream()), "crab_gemm_bf16")
    return out
