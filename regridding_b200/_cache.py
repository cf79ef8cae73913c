"""
Keeps the device-resident form of a saved-weights element next to the host tuple the
reference's API exposes, so ``regrid_from_weights(*weights(...))`` does not upload and
convert the weights again.

* Keyed by the identity of ALL THREE host arrays ``(indices_input, indices_output, values)``:
  ``transpose_weights`` re-uses the values array with the index arrays swapped, which is a
  different matrix and gets its own entry (forward and transposed applies do not evict each other).
* Stale entries: the arrays handed out by ``weights()`` are marked read-only, so an in-place edit
  (``w[0][()][2][bad] = 0``) raises instead of silently leaving the device copy behind; a caller who
  flips ``writeable`` back is caught by a fingerprint of a strided sample of the three arrays that
  is re-checked on every lookup.
* An entry dies with any of its host arrays (weak references); the cache is additionally bounded by
  the bytes of device memory it keeps alive (LRU, ``REGRID_B200_CACHE_BYTES``, default 8 GiB).
"""

from __future__ import annotations

import collections
import os
import weakref

import numpy as np

_entries: "collections.OrderedDict[tuple[int, int, int], tuple]" = collections.OrderedDict()
_SAMPLE = 64


def _limit() -> int:
    return int(os.environ.get("REGRID_B200_CACHE_BYTES", str(8 << 30)))


def _fingerprint(arrays) -> tuple:
    """Cheap content check: sizes plus a strided sample of <= 64 elements of every array."""
    out = []
    for a in arrays:
        a = np.asarray(a)
        n = a.size
        step = max(1, n // _SAMPLE)
        out.append((n, a.reshape(-1)[::step][:_SAMPLE + 1].tobytes(), a.reshape(-1)[-1:].tobytes()))
    return tuple(out)


def _device_bytes(dw) -> int:
    try:
        return int(dw.device_bytes())
    except Exception:  # noqa: BLE001 -- tokens in tests
        return 0


def freeze(element) -> None:
    """Mark the host arrays of a saved-weights element read-only (what ``weights()`` hands out)."""
    for a in element:
        if isinstance(a, np.ndarray):
            a.flags.writeable = False


def remember(element, device_weights) -> None:
    """``element`` = the host tuple ``(indices_input, indices_output, values)``."""
    indices_input, indices_output, values = element
    key = (id(values), id(indices_input), id(indices_output))
    try:
        drop = lambda _r, key=key: _entries.pop(key, None)  # noqa: E731
        refs = (weakref.ref(values, drop), weakref.ref(indices_input, drop), weakref.ref(indices_output, drop))
    except TypeError:  # not weak-referenceable (e.g. a Quantity subclass without __weakref__)
        return
    _entries.pop(key, None)
    _entries[key] = (*refs, device_weights, _fingerprint(element))
    # LRU bound on the device memory the cache keeps alive (the newest entry always stays)
    total = sum(_device_bytes(e[3]) for e in _entries.values())
    while total > _limit() and len(_entries) > 1:
        _, old = _entries.popitem(last=False)
        total -= _device_bytes(old[3])


def lookup(element, device=None):
    indices_input, indices_output, values = element
    key = (id(values), id(indices_input), id(indices_output))
    hit = _entries.get(key)
    if hit is None:
        return None
    rv, ri, ro, dw, fp = hit
    if rv() is not values or ri() is not indices_input or ro() is not indices_output:
        return None
    if fp != _fingerprint(element):   # edited in place after the upload: the device copy is stale
        _entries.pop(key, None)
        return None
    if device is not None and dw.device != device:
        return None
    _entries.move_to_end(key)
    return dw


def clear() -> None:
    _entries.clear()
