"""chord_b200 — B200-native (sm_100a) implementation of chord's Nanite-style visibility-buffer path.

The product is ``libchord_b200.so`` (hand-written CUDA behind the C ABI in include/chord_b200.h).  This package is
the thin host side: ctypes bindings (runtime), the mirror of the reference's pass-graph API (renderer), the
camera/uniform fill (camera), ABI struct mirrors (abi) and seeded synthetic scenes (synth).

There is no CPU fallback: loading fails loudly if the CUDA library is missing, and every pass returns an error
if no sm_100 device is present.
"""
from . import abi  # noqa: F401

__all__ = ['abi']
