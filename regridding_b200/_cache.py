"""
Keeps the device-resident form of a saved-weights element next to the host tuple the
reference's API exposes, so ``regrid_from_weights(*weights(...))`` does not upload and
convert the weights again.  Keyed by the identity of the host ``values`` array; an entry
dies with that array (weak reference), and a stale ``id`` is never trusted.
"""

from __future__ import annotations

import weakref

_entries: dict[int, tuple[weakref.ref, object]] = {}


def remember(values_host, device_weights) -> None:
    try:
        key = id(values_host)
        ref = weakref.ref(values_host, lambda _r, key=key: _entries.pop(key, None))
    except TypeError:  # not weak-referenceable (e.g. a Quantity subclass without __weakref__)
        return
    _entries[key] = (ref, device_weights)


def lookup(values_host, device=None):
    hit = _entries.get(id(values_host))
    if hit is None:
        return None
    ref, dw = hit
    if ref() is not values_host:
        return None
    if device is not None and dw.device != device:
        return None
    return dw


def clear() -> None:
    _entries.clear()
