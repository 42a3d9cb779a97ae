"""cadre_b200 — B200-native (sm_100a) learner hot path of BIT-MCS/Cadre behind the reference's Python agent API.

The device side is libcadre_sm100.so (hand-written CUDA, C ABI in include/cadre_b200.h); this package is the
thin host side. Importing the package does not load the library; the first call that needs it does and raises
`CadreError` if it is missing — there is no CPU / eager fallback.
"""
from ._lib import CadreError, LIB_PATH  # noqa: F401

__all__ = ["CadreError", "LIB_PATH"]
