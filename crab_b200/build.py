"""Build the in-tree sm_100a shared library `crab_b200/_lib/libcrab_b200.so` with nvcc (no torch headers: the
boundary is a plain C ABI).  Object files are cached per source by mtime so re-builds only recompile what changed."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
LIBDIR = ROOT / "_lib"
LIB = LIBDIR / "libcrab_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--use_fast_math",
    "-Xptxas", "-v",
    "-Xcudafe", "--diag_suppress=177",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


DIAG_SOURCES = {"debug_stream.cu"}   # diagnostics for tools/ only: built into libcrab_diag.so, never into the product library


def sources() -> list[Path]:
    return sorted(p for p in CSRC.glob("*.cu") if p.name not in DIAG_SOURCES)


def build_diag() -> Path:
    """tools/bench_stream*.py: the pure-streaming probe kernel, in its own shared library."""
    LIBDIR.mkdir(exist_ok=True)
    out = LIBDIR / "libcrab_diag.so"
    srcs = [CSRC / n for n in sorted(DIAG_SOURCES)] + [CSRC / "host_common.cu"]
    if _stale(out, srcs):
        cmd = [_nvcc(), *[f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")], "-shared", *map(str, srcs), "-o", str(out), "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"diag build failed:\n{r.stdout}\n{r.stderr}")
    return out


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(verbose: bool = False, force: bool = False) -> Path:
    LIBDIR.mkdir(exist_ok=True)
    objdir = LIBDIR / "obj"
    objdir.mkdir(exist_ok=True)
    headers = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [ROOT.parent / "include" / "crab_b200.h"]
    nvcc = _nvcc()
    jobs = []
    for src in sources():
        obj = objdir / (src.stem + ".o")
        if force or _stale(obj, [src] + headers):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = objdir / (src.stem + ".log")
        log.write_text(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    objs = [objdir / (s.stem + ".o") for s in sources()]
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs),
               "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(p)
