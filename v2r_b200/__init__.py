"""v2r_b200 -- B200-native implementation of Dewberry/v2r's IDW gridding hot path.

The product is the CUDA library `lib/libv2r_idw.so` behind the C ABI of `include/v2r_idw.h`;
this package is the thin Python host used by the tests and bench.py (ctypes over that ABI, with
the reference's names: FullSolve, ChunkSolve, MakeCoordSpace, GetDimensions, ...).
There is no CPU fallback: importing works anywhere, computing needs the library and a GPU.
"""
from . import _lib  # noqa: F401
from .idw import (ChunkSolve, FullSolve, GetDimensions, IdwContext, Info, MakeCoordSpace, PinnedArray,  # noqa: F401
                  PointSet, PointToPair, exp_range, solve)

__all__ = ["FullSolve", "ChunkSolve", "MakeCoordSpace", "GetDimensions", "PointToPair", "Info", "PointSet",
           "IdwContext", "PinnedArray", "exp_range", "solve"]
