"""
Keeps the device-resident form of a saved-weights element next to the host tuple the
reference's API exposes, so ``regrid_from_weights(*weights(...))`` does not upload and
convert the weights again.  Keyed by the identity of the host ``values`` array AND of both
index arrays (``transpose_weights`` re-uses the values array with the index arrays swapped,
which must not hit the forward matrix); an entry dies with the values array (weak
reference), and a stale ``id`` is never trusted.
"""

from __future__ import annotations

import weakref

_entries: dict[int, tuple[weakref.ref, weakref.ref, weakref.ref, object]] = {}


def remember(element, device_weights) -> None:
    """``element`` = the host tuple ``(indices_input, indices_output, values)``."""
    indices_input, indices_output, values = element
    try:
        key = id(values)
        refs = (weakref.ref(values, lambda _r, key=key: _entries.pop(key, None)),
                weakref.ref(indices_input), weakref.ref(indices_output))
    except TypeError:  # not weak-referenceable (e.g. a Quantity subclass without __weakref__)
        return
    _entries[key] = (*refs, device_weights)


def lookup(element, device=None):
    indices_input, indices_output, values = element
    hit = _entries.get(id(values))
    if hit is None:
        return None
    rv, ri, ro, dw = hit
    if rv() is not values or ri() is not indices_input or ro() is not indices_output:
        return None
    if device is not None and dw.device != device:
        return None
    return dw


def clear() -> None:
    _entries.clear()
