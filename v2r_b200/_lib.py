"""ctypes binding of libv2r_idw.so -- exactly the symbols include/v2r_idw.h declares."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libv2r_idw.so")
CSRC = os.path.join(_HERE, "csrc")

OK, EINVAL, ECUDA, ENCCL, ENOMEM, ESTATE = 0, 1, 2, 3, 4, 5
FP64, FP32 = 0, 1


class Grid(C.Structure):
    """struct v2r_idw_grid"""
    _fields_ = [("x_min", C.c_double), ("x_step", C.c_double), ("y_min", C.c_double), ("y_step", C.c_double),
                ("rows", C.c_int64), ("cols", C.c_int64)]


class IdwError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"v2r_idw error {code}: {msg}")
        self.code = code


_D = C.POINTER(C.c_double)
_I = C.POINTER(C.c_int64)
_P = C.c_void_p

# name -> (restype, argtypes); must list every function of include/v2r_idw.h
SIGNATURES = {
    "v2r_idw_abi_version": (C.c_int, []),
    "v2r_idw_last_error": (C.c_char_p, []),
    "v2r_idw_ctx_last_error": (C.c_char_p, [_P]),
    "v2r_idw_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "v2r_idw_init": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "v2r_idw_init_device": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "v2r_idw_destroy": (None, [_P]),
    "v2r_idw_alloc_pinned": (C.c_int, [C.c_size_t, C.POINTER(_P)]),
    "v2r_idw_free_pinned": (C.c_int, [_P]),
    "v2r_idw_get_dimensions": (C.c_int, [C.c_double] * 6 + [_I, _I]),
    "v2r_idw_point_to_pair": (C.c_int, [C.c_double] * 6 + [_I, _I]),
    "v2r_idw_exp_range": (C.c_int, [C.c_double] * 3 + [_D, C.c_int64, _I]),
    "v2r_idw_make_coord_space": (C.c_int, [C.c_int64, _D, _D, _D] + [C.c_double] * 4 + [_D, _D, _D, _I, _I, _I]),
    "v2r_idw_solve": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int64, C.POINTER(Grid), _P, C.c_int, C.c_int, _P]),
    "v2r_idw_solve_rows": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int64, C.POINTER(Grid), C.c_int64, C.c_int64, _P,
                                     C.c_int, C.c_int, _P]),
    "v2r_idw_stream_begin": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int64, C.POINTER(Grid), _P, C.c_int, C.c_int,
                                       C.c_int64]),
    "v2r_idw_stream_next": (C.c_int, [_P, _P, _I, _I]),
    "v2r_idw_stream_end": (C.c_int, [_P]),
    "v2r_idw_launch_rows": (C.c_int, [C.c_int, _P, _P, _P, _P, _P, _P, C.c_int64, C.POINTER(Grid), C.c_int64,
                                      C.c_int64, _P, C.c_int, C.c_int, _P]),
    "v2r_idw_last_timing": (C.c_int, [_P, _D, _D, _D]),
    "v2r_idw_launch_count": (C.c_int64, []),
    "v2r_idw_describe": (C.c_int, [_P, C.c_int, C.c_int, C.c_char_p, C.c_size_t, _D]),
    "v2r_idw_measure_peaks": (C.c_int, [C.c_int, C.c_double, _D, _D, _D, _D]),
    "v2r_idw_measure_peak": (C.c_int, [C.c_int, C.c_double, C.c_int, _D]),
}

_lib = None


def build(force: bool = False) -> str:
    """Compile lib/libv2r_idw.so with nvcc for sm_100a (csrc/Makefile).  No GPU needed."""
    args = ["make", "-C", CSRC, "-j4", "-s"] + (["-B"] if force else [])
    subprocess.check_call(args)
    return LIB_PATH


def load():
    """Load the library.  Fails loudly if it is missing: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            try:  # a fresh checkout: compile it (nvcc cross-compiles sm_100a without a GPU); never a substitute path
                build()
            except Exception as e:  # noqa: BLE001
                raise ImportError(f"{LIB_PATH} is missing and could not be built ({e}): run "
                                  "`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a).  "
                                  "v2r_b200 has no CPU fallback.") from e
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int):
    if rc != OK:
        raise IdwError(rc, load().v2r_idw_last_error().decode(errors="replace"))
