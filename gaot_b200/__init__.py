"""gaot_b200 - B200-native forward/backward path of GAOT behind the reference's `src/model` operator API.

`gaot_b200.model` mirrors `src.model` of the reference (same class names, constructor arguments,
forward signatures and state-dict keys); every operator runs in `libgaot_b200.so`
(include/gaot_b200.h), hand-written CUDA for sm_100a.  There is no CPU or eager fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
