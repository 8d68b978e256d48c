"""ctypes binding of `libcrab_b200.so` (the C ABI declared in include/crab_b200.h).

There is no fallback: if the shared library is missing or a call returns an error code, a Python exception is raised.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "_lib" / "libcrab_b200.so"
_lib = None


class CrabError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("B", C.c_void_p), ("C", C.c_void_p), ("bias", C.c_void_p), ("residual", C.c_void_p),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("lda", C.c_int32), ("ldb", C.c_int32), ("ldc", C.c_int32), ("ldr", C.c_int32),
        ("res_scale", C.c_float), ("out_scale", C.c_float),
        ("act", C.c_int32), ("out_dtype", C.c_int32), ("block_n", C.c_int32), ("max_ctas", C.c_int32),
    ]


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load the shared library (building is `python -m crab_b200.build` / `__graft_entry__.build()`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise CrabError(
            f"{_LIB_PATH} is missing: build it with `python crab_b200/build.py` (nvcc, sm_100a). "
            "crab_b200 has no CPU or PyTorch fallback path."
        )
    lib = C.CDLL(str(_LIB_PATH))
    lib.crab_last_error.restype = C.c_char_p
    lib.crab_last_error.argtypes = []
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().crab_last_error()
        raise CrabError(f"{what} failed with code {rc}: {msg.decode(errors='replace') if msg else ''}")


def exported_symbols() -> list[str]:
    """Names declared in include/crab_b200.h (parsed from the header) — used by the CPU-side ABI test."""
    import re

    hdr = (Path(__file__).resolve().parent.parent / "include" / "crab_b200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(crab_[a-z0-9_]+)\s*\(", hdr)))


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("o", C.c_void_p),
        ("q_bs", C.c_int64), ("q_rs", C.c_int64), ("q_hs", C.c_int64),
        ("k_bs", C.c_int64), ("k_rs", C.c_int64), ("k_hs", C.c_int64),
        ("v_bs", C.c_int64), ("v_rs", C.c_int64), ("v_hs", C.c_int64),
        ("o_bs", C.c_int64), ("o_rs", C.c_int64), ("o_hs", C.c_int64),
        ("B", C.c_int32), ("H", C.c_int32), ("KVH", C.c_int32), ("Sq", C.c_int32), ("Sk", C.c_int32),
        ("head_dim", C.c_int32), ("scale", C.c_float), ("causal", C.c_int32),
        ("gate", C.c_void_p), ("bias_table", C.c_void_p), ("sk_dev", C.c_void_p),
    ]


class SkinnyArgs(C.Structure):
    _fields_ = [
        ("X", C.c_void_p), ("W", C.c_void_p), ("W_packed", C.c_void_p), ("C", C.c_void_p), ("bias", C.c_void_p),
        ("residual", C.c_void_p),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("ldx", C.c_int32), ("ldw", C.c_int32), ("ldc", C.c_int32),
        ("ldr", C.c_int32), ("act", C.c_int32), ("out_dtype", C.c_int32), ("splits", C.c_int32),
        ("Z", C.c_void_p), ("ldz", C.c_int32), ("Kext", C.c_int32),
        ("stats_packed", C.c_void_p), ("stats_linears", C.c_int32), ("norm", C.c_int32),
        ("eps", C.c_float), ("lora_scale", C.c_float),
        ("rstd", C.c_void_p), ("flags", C.c_void_p),
        ("prefetch", C.c_void_p), ("prefetch_bytes", C.c_int64),
        ("stats_scratch", C.c_void_p), ("flags_clear", C.c_void_p), ("stats_clusters", C.c_int32), ("reserved0", C.c_int32),
    ]


class ChainPhase(C.Structure):
    _fields_ = [
        ("X", C.c_void_p), ("ldx", C.c_int32), ("K", C.c_int32),
        ("Z", C.c_void_p), ("ldz", C.c_int32), ("Kext", C.c_int32),
        ("W_packed", C.c_void_p),
        ("stats_packed", C.c_void_p), ("stats_linears", C.c_int32), ("norm", C.c_int32),
        ("eps", C.c_float), ("lora_scale", C.c_float),
        ("rstd", C.c_void_p),
        ("C", C.c_void_p), ("ldc", C.c_int32), ("N", C.c_int32), ("out_dtype", C.c_int32), ("act", C.c_int32),
        ("bias", C.c_void_p),
        ("residual", C.c_void_p), ("ldr", C.c_int32),
    ]


class ChainArgs(C.Structure):
    _fields_ = [
        ("phase", ChainPhase * 4),
        ("n_phases", C.c_int32), ("M", C.c_int32), ("cluster", C.c_int32), ("max_clusters", C.c_int32),
        ("counters", C.c_void_p),
    ]


class DecodeFusedArgs(C.Structure):
    _fields_ = [
        ("qkv", C.c_void_p), ("ldq", C.c_int32), ("cos_sin", C.c_void_p), ("k_cache", C.c_void_p), ("v_cache", C.c_void_p),
        ("o", C.c_void_p), ("ldo", C.c_int32), ("workspace", C.c_void_p),
        ("B", C.c_int32), ("H", C.c_int32), ("KVH", C.c_int32), ("head_dim", C.c_int32), ("ctx_max", C.c_int32),
        ("nsplit", C.c_int32), ("past_dev", C.c_void_p), ("scale", C.c_float),
        ("lora_ra", C.c_void_p), ("ld_ra", C.c_int32), ("lora_z", C.c_void_p), ("ld_z", C.c_int32),
        ("lora_scale", C.c_float), ("lora_ws", C.c_void_p), ("lora_counters", C.c_void_p),
        ("gqa_tensor_cores", C.c_int32), ("reserved0", C.c_int32),
    ]
