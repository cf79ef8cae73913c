"""
regridding_b200 -- B200-native first-order conservative regridding.

Drop-in for the conservative path of sun-data/regridding: ``weights``,
``regrid_from_weights``, ``regrid`` and ``find_indices`` keep the reference's
signatures, error behaviour and saved-weights layout; the compiled kernels behind them
are hand-written CUDA for sm_100a in ``libregrid_b200.so`` (C ABI:
``include/regrid_b200.h``).  There is no CPU fallback.
"""

from ._find_indices import find_indices
from ._regrid import regrid, regrid_from_weights
from ._weights import weights, weights_packed
from ._packed import PackedWeights
from ._fill import fill
from ._interp import ndarray_linear_interpolation
from ._transposed import transpose_weights, transpose_weights_conservative
from . import _device as device  # device-resident operators (torch CUDA tensors in / out)

__all__ = ["regrid", "weights", "regrid_from_weights", "find_indices", "transpose_weights",
           "transpose_weights_conservative", "weights_packed", "PackedWeights", "fill", "ndarray_linear_interpolation", "device"]
__version__ = "0.1.0"
