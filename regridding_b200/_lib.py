"""
ctypes binding of ``libregrid_b200.so`` (C ABI declared in ``include/regrid_b200.h``).

There is NO fallback: if the CUDA library is missing or a CUDA device is not
available the product path raises.  PyTorch is used only for device memory,
streams and (elsewhere) ``torch.distributed``.
"""

from __future__ import annotations

import ctypes
import pathlib

_HERE = pathlib.Path(__file__).resolve().parent
LIB_PATH = _HERE / "libregrid_b200.so"

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int
_sz = ctypes.c_size_t
_p_i64 = ctypes.POINTER(ctypes.c_int64)
_p_sz = ctypes.POINTER(ctypes.c_size_t)
_p_i32 = ctypes.POINTER(ctypes.c_int32)

# name -> argtypes; every function returns int except the two noted below.
SIGNATURES: dict[str, list] = {
    "rg_build2d_workspace_bytes": [_i64, _i64, _i64, _i64, _p_sz],
    "rg_build2d_count": [_int, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _sz, _p_i64],
    "rg_build2d_fill": [_int, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _sz,
                        _vp, _i64, _p_i64],
    "rg_build2d_emit": [_int, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _sz, _vp, _i64,
                        _vp, _vp, _vp, _i64],
    "rg_build2d_batched": [_int, _vp, _i64, _i64, _i64, _i64, _i64, ctypes.POINTER(_vp), ctypes.POINTER(_vp),
                           ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp), _vp, _sz, _vp, _i64,
                           ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp), _i64, _vp],
    "rg_build2d_part_count": [_int, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _int, _int, _vp, _sz, _p_i64,
                              _int, _p_i64, _p_i64, _p_sz, _vp],
    "rg_build2d_part_fill": [_int, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _int, _int, _vp, _sz,
                             _vp, _i64],
    "rg_build2d_merge_workspace_bytes": [_i64, _int, _p_sz],
    "rg_build2d_gather_counts": [_int, _vp, _i64, _int, ctypes.POINTER(_vp), _vp],
    "rg_build2d_merge": [_int, _vp, _i64, _int, _vp, ctypes.POINTER(_vp), _p_i64, _vp, _sz, _vp, _p_i64],
    "rg_build2d_merge_emit": [_int, _vp, _i64, _int, _i64, _vp, _sz, _vp, _vp, _vp, _vp, _i64],
    "rg_build2d_stats": [_int, _vp, _i64, _i64, _i64, _i64, _vp, _p_i32],
    "rg_build2d_band": [_int, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _sz,
                        _vp, _i64, _vp, _vp, _vp, _i64, _vp],
    "rg_build2d_band_onewalk": [_int, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _sz,
                                _vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _i64],
    "rg_grid_area": [_int, _vp, _i64, _i64, _vp, _vp, _vp],
    "rg_find_indices_2d_workspace_bytes": [_i64, _i64, _i64, _p_sz],
    "rg_find_indices_2d": [_int, _vp, _i64, _i64, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _sz],
    "rg_multilinear2d_weights": [_int, _vp, _i64, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _int, _vp, _vp, _vp],
    "rg_ell4_apply": [_int, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp],
    "rg_multilinear1d_workspace_bytes": [_i64, _i64, _p_sz],
    "rg_multilinear1d_weights": [_int, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _int, _vp, _vp, _vp, _p_i64, _vp, _sz],
    "rg_sort_triplets_workspace_bytes": [_i64, _p_sz],
    "rg_sort_triplets": [_int, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz],
    "rg_interp_linear_1d": [_int, _vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp],
    "rg_interp_bilinear_2d": [_int, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp],
    "rg_measure_fp64_peak": [_int, _vp, ctypes.POINTER(ctypes.c_double)],
    "rg_fill_gauss_seidel_2d": [_int, _vp, _vp, _i64, _i64, _i64, ctypes.POINTER(_vp), _p_i64, _i64],
    "rg_csr_workspace_bytes": [_i64, _i64, _p_sz],
    "rg_csr_from_coo": [_int, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz],
    "rg_apply_csr": [_int, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp],
    "rg_apply_plan_sizes": [_i64, _i64, _p_i64, _p_i64, _p_i64],
    "rg_apply_plan_build": [_int, _vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _p_i64, _p_i64],
    "rg_apply_plan_slots": [_int, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _vp],
    "rg_apply_planned": [_int, _vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp],
    "rg_cons1d_batched": [_int, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "rg_regrid1d_conservative": [_int, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp],
    "rg_find_indices_1d": [_int, _vp, _int, _i64, _i64, _i64, _vp, _vp, _i64, _vp],
    "rg_cell_length_1d": [_int, _vp, _i64, _i64, _vp, _vp],
    "rg_transpose_conservative": [_int, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
}
OTHER_SYMBOLS = ["rg_last_error_string", "rg_version"]

_lib = None


class RegridB200Error(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load the CUDA library; fail loudly if it has not been built."""
    global _lib
    if _lib is None:
        import os

        override = os.environ.get("REGRID_B200_LIB")  # development: load an alternative build of the same ABI
        path = pathlib.Path(override) if override else LIB_PATH
        if not path.exists():
            raise RegridB200Error(
                f"{path} is missing: build it with `python -m regridding_b200._build` "
                "(there is no CPU fallback)"
            )
        L = ctypes.CDLL(str(path))
        for name, argtypes in SIGNATURES.items():
            f = getattr(L, name)
            f.argtypes = argtypes
            f.restype = ctypes.c_int
        L.rg_last_error_string.restype = ctypes.c_char_p
        L.rg_version.restype = ctypes.c_int
        _lib = L
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc == 0:
        return
    msg = load().rg_last_error_string().decode(errors="replace")
    if rc == -4:
        raise RegridB200Error(f"{what}: {msg}")
    if rc < 0:
        raise ValueError(f"{what}: {msg} (code {rc})")
    raise RegridB200Error(f"{what}: CUDA error {rc}: {msg}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()
