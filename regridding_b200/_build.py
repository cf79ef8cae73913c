"""
Builds ``libregrid_b200.so`` (hand-written CUDA for sm_100a behind the C ABI of
``include/regrid_b200.h``) in-tree with nvcc.  ``python -m regridding_b200._build``.
"""

from __future__ import annotations

import pathlib
import shutil
import subprocess
import sys

HERE = pathlib.Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libregrid_b200.so"

SOURCES = ["rg_util.cu", "rg_apply.cu", "rg_apply_bulk.cu", "rg_build2d.cu", "rg_locate.cu", "rg_cons1d.cu", "rg_multilinear2d.cu", "rg_multilinear1d.cu", "rg_interp.cu", "rg_fill.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # every fused multiply-add in the fp64 geometry is explicit (parity with the reference's op order)
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
    "--shared",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return exe


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "regrid_b200.h"]
    return any(p.stat().st_mtime > t for p in deps)


def build(force: bool = False, verbose: bool = False, defines: list[str] | None = None,
          out: pathlib.Path | None = None, replace: dict[str, str] | None = None) -> pathlib.Path:
    """``defines``/``out``/``replace`` (source file substitutions) build a tuning variant next to the product
    library (development only)."""
    if out is None and not force and not needs_build():
        return LIB
    out = out or LIB
    cmd = [nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in (defines or [])], "-ccbin", "/usr/bin/g++", "-o", str(out),
           *[str(CSRC / (replace or {}).get(s, s)) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libregrid_b200.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
