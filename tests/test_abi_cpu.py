"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/crab_b200.h
declares, the header is plain C, and the ctypes mirrors of its structs have the C layout.  No compute calls."""
import ctypes
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from crab_b200 import build, lib as L

    if not L.lib_path().exists():
        build.build()
    return L


def test_library_exports_every_declared_symbol(lib):
    handle = lib.load()
    syms = lib.exported_symbols()
    assert len(syms) >= 20 and "crab_gemm_bf16" in syms and "crab_flash_attn" in syms
    for s in syms:
        assert getattr(handle, s) is not None


def test_header_is_plain_c_and_struct_layouts_match(lib, tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "crab_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(crab_gemm_args), offsetof(crab_gemm_args, M),'
                   ' offsetof(crab_gemm_args, max_ctas), sizeof(crab_attn_args), offsetof(crab_attn_args, B),'
                   ' offsetof(crab_attn_args, bias_table));'
                   'printf("%zu %zu %zu %zu\\n", sizeof(crab_decode_fused_args), offsetof(crab_decode_fused_args, B),'
                   ' offsetof(crab_decode_fused_args, scale), offsetof(crab_decode_fused_args, lora_counters));'
                   'printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(crab_chain_phase), offsetof(crab_chain_phase, rstd),'
                   ' offsetof(crab_chain_phase, ldr), sizeof(crab_chain_args), offsetof(crab_chain_args, n_phases),'
                   ' offsetof(crab_chain_args, counters)); return 0;}\n')
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    got = list(map(int, subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()))
    G, A = lib.GemmArgs, lib.AttnArgs
    D = lib.DecodeFusedArgs
    P, CA = lib.ChainPhase, lib.ChainArgs
    assert got == [ctypes.sizeof(G), G.M.offset, G.max_ctas.offset, ctypes.sizeof(A), A.B.offset, A.bias_table.offset,
                   ctypes.sizeof(D), D.B.offset, D.scale.offset, D.lora_counters.offset,
                   ctypes.sizeof(P), P.rstd.offset, P.ldr.offset, ctypes.sizeof(CA), CA.n_phases.offset, CA.counters.offset]


def test_error_reporting_without_gpu(lib):
    import torch

    handle = lib.load()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    rc = handle.crab_init(ctypes.c_int(0))
    assert rc != 0
    assert len(handle.crab_last_error()) > 0
    rc = handle.crab_gemm_bf16(None, None)
    assert rc == -1 and b"null" in handle.crab_last_error()


def test_product_path_fails_loudly_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from crab_b200 import ops
    from crab_b200.engine import CrabConfig, CrabEngine
    from crab_b200.lib import CrabError

    with pytest.raises(CrabError):
        CrabEngine({}, CrabConfig())
    with pytest.raises(CrabError):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under crab_b200/ may import it (tier rule)."""
    for f in (ROOT / "crab_b200").rglob("*.py"):
        txt = f.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt, f
