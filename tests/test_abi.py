"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/*.h declares.
CPU only: no compute entry point is launched (argument validation and size queries only)."""

import ctypes
import pathlib
import re

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from regridding_b200 import _build, _lib

    _build.build()
    return _lib.load()


def declared_symbols():
    text = (ROOT / "include" / "regrid_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    names = declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/regrid_b200.h but not exported"


def test_binding_table_matches_header():
    from regridding_b200 import _lib

    assert sorted(list(_lib.SIGNATURES) + _lib.OTHER_SYMBOLS) == declared_symbols()


def test_library_is_sm100a():
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not pathlib.Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", str(ROOT / "regridding_b200" / "libregrid_b200.so")],
                         capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_size_queries_and_argument_errors(lib):
    n = ctypes.c_size_t()
    assert lib.rg_build2d_workspace_bytes(2049, 2049, 2049, 2049, ctypes.byref(n)) == 0
    assert 100e6 < n.value < 2e9
    assert lib.rg_build2d_workspace_bytes(1, 5, 5, 5, ctypes.byref(n)) == -1
    assert b"2x2" in lib.rg_last_error_string()
    assert lib.rg_build2d_workspace_bytes(70000, 70000, 5, 5, ctypes.byref(n)) == -2
    assert lib.rg_csr_workspace_bytes(15_000_000, 4_194_304, ctypes.byref(n)) == 0 and n.value > 60e6
    assert lib.rg_find_indices_2d_workspace_bytes(4096, 4096, 1 << 20, ctypes.byref(n)) == 0
    # null pointers are rejected before any CUDA call
    assert lib.rg_apply_csr(0, None, 4, 10, 10, None, None, None, None, None) == -1
    assert lib.rg_cons1d_batched(0, None, 4, 1, 5, None, None, None, None, None, None, None) == -1
    assert lib.rg_find_indices_1d(0, None, 7, 1, 5, 5, None, None, 0, None) == -1
    assert lib.rg_version() >= 100
