"""Load (building on demand) the host-emulated twin of libmal_b200.  Tests only."""
from __future__ import annotations

import ctypes

from mal_b200 import _capi
from tests.emu.build_emu import build

_handle = None


def emu():
    global _handle
    if _handle is None:
        _handle = _capi.bind(ctypes.CDLL(build()))
    return _handle
