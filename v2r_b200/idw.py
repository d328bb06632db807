"""Host-side mirror of the reference's IDW interface over the C ABI (include/v2r_idw.h).

Same names and argument meaning as the Go entry points so the parity tests read like the
reference's own tests (features/idw/idw_test.go):

    tools.Info / MakeInfo            tools/utils.go:123-131      -> Info
    tools.GetDimensions              tools/utils.go:133-137      -> GetDimensions
    tools.PointToPair                tools/utils.go:70-72        -> PointToPair
    tools.MakeCoordSpace             tools/utils.go:97-119       -> MakeCoordSpace (input-order, deterministic)
    idw.FullSolve                    features/idw/idw.go:15-65   -> FullSolve
    idw.ChunkSolve                   features/idw/idw_chunk.go:43-105 -> ChunkSolve
    exponent loop                    cmd/idw.go:153-161          -> exp_range + ONE solve() call

Everything numeric happens in libv2r_idw.so on the GPU; numpy only carries buffers.
"""
from __future__ import annotations

import ctypes as C
import math
import time
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import FP32, FP64, Grid, check


@dataclass
class Info:
    """tools.Info{Min,Max,Step}; Info() == tools.MakeInfo() == {+Inf, -Inf, 1.0}."""
    min: float = math.inf
    max: float = -math.inf
    step: float = 1.0


@dataclass
class PointSet:
    """map[tools.OrderedPair]tools.Point flattened: SoA x,y,w plus the (r,c) key of every point."""
    x: np.ndarray
    y: np.ndarray
    w: np.ndarray
    key_r: np.ndarray
    key_c: np.ndarray

    def __len__(self):
        return int(len(self.x))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None and a.size else C.c_void_p(0)


def GetDimensions(xInfo: Info, yInfo: Info):
    r, c = C.c_int64(), C.c_int64()
    check(_lib.load().v2r_idw_get_dimensions(xInfo.min, xInfo.max, xInfo.step, yInfo.min, yInfo.max, yInfo.step,
                                             C.byref(r), C.byref(c)))
    return r.value, c.value


def PointToPair(x: float, y: float, xInfo: Info, yInfo: Info):
    r, c = C.c_int64(), C.c_int64()
    check(_lib.load().v2r_idw_point_to_pair(x, y, xInfo.min, xInfo.step, yInfo.min, yInfo.step, C.byref(r), C.byref(c)))
    return r.value, c.value


def MakeCoordSpace(x, y, w, xInfo: Info, yInfo: Info) -> PointSet:
    x, y, w = _f64(x), _f64(y), _f64(w)
    n = len(x)
    ox, oy, ow = np.empty(n), np.empty(n), np.empty(n)
    kr, kc = np.empty(n, np.int64), np.empty(n, np.int64)
    m = C.c_int64()
    D, I = C.POINTER(C.c_double), C.POINTER(C.c_int64)
    check(_lib.load().v2r_idw_make_coord_space(
        n, x.ctypes.data_as(D), y.ctypes.data_as(D), w.ctypes.data_as(D), xInfo.min, xInfo.step, yInfo.min, yInfo.step,
        ox.ctypes.data_as(D), oy.ctypes.data_as(D), ow.ctypes.data_as(D), kr.ctypes.data_as(I), kc.ctypes.data_as(I),
        C.byref(m)))
    k = m.value
    return PointSet(ox[:k].copy(), oy[:k].copy(), ow[:k].copy(), kr[:k].copy(), kc[:k].copy())


def exp_range(es: float, ee: float, ei: float) -> np.ndarray:
    """cmd/idw.go:153 exponent list (FP64-accumulated); raises like expRange (cmd/idw.go:83-88)."""
    buf = np.empty(4096)
    cnt = C.c_int64()
    check(_lib.load().v2r_idw_exp_range(es, ee, ei, buf.ctypes.data_as(C.POINTER(C.c_double)), len(buf), C.byref(cnt)))
    return buf[:min(cnt.value, len(buf))].copy()


def make_grid(xInfo: Info, yInfo: Info, rows=None, cols=None) -> Grid:
    if rows is None or cols is None:
        rows, cols = GetDimensions(xInfo, yInfo)
    return Grid(xInfo.min, xInfo.step, yInfo.min, yInfo.step, rows, cols)


class PinnedArray:
    """A numpy view of page-locked host memory from v2r_idw_alloc_pinned (what the readers of the C++ / Go hosts fill):
    copies to and from it are asynchronous DMA, and a multi-GPU context copies every device's band straight into it."""

    def __init__(self, shape, dtype=np.float64):
        self._L = _lib.load()
        self._p = C.c_void_p()
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        check(self._L.v2r_idw_alloc_pinned(max(n, 8), C.byref(self._p)))
        self.array = np.ctypeslib.as_array((C.c_byte * max(n, 8)).from_address(self._p.value))[:n].view(dtype).reshape(shape)

    def free(self):
        if self._p:
            self.array = None
            self._L.v2r_idw_free_pinned(self._p)
            self._p = C.c_void_p()

    def __enter__(self):
        return self.array

    def __exit__(self, *a):
        self.free()


class IdwContext:
    """v2r_idw_ctx: one process driving n_gpus devices, or one given device."""

    def __init__(self, n_gpus: int = 1, device: int | None = None):
        self._L = _lib.load()
        self._h = C.c_void_p()
        if device is not None:
            check(self._L.v2r_idw_init_device(device, C.byref(self._h)))
        else:
            check(self._L.v2r_idw_init(n_gpus, C.byref(self._h)))

    def close(self):
        if self._h:
            self._L.v2r_idw_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- solves ---------------------------------------------------------------------------
    def solve_rows(self, ps: PointSet, grid: Grid, row0: int, nrows: int, exps, precision: int = FP64, out=None):
        x, y, w, kr, kc = _f64(ps.x), _f64(ps.y), _f64(ps.w), _i64(ps.key_r), _i64(ps.key_c)
        exps = _f64(np.atleast_1d(exps))
        if out is None:
            out = np.empty((len(exps), nrows, grid.cols))
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.size == len(exps) * nrows * grid.cols
        check(self._L.v2r_idw_solve_rows(self._h, _ptr(x), _ptr(y), _ptr(w), _ptr(kr), _ptr(kc), len(x), C.byref(grid),
                                         row0, nrows, _ptr(exps), len(exps), precision, _ptr(out)))
        return out

    def solve(self, ps: PointSet, grid: Grid, exps, precision: int = FP64, out=None):
        return self.solve_rows(ps, grid, 0, grid.rows, exps, precision, out)

    def stream(self, ps: PointSet, grid: Grid, exps, precision: int = FP64, band_rows: int = 0, pinned: bool = False):
        """Generator over (row0, band[E, nrows, cols]) -- the chunk-write protocol of ChunkSolve.  pinned=True receives
        the bands in ONE page-locked buffer (like the C++ / Go hosts): the yielded view is valid until the next step."""
        x, y, w, kr, kc = _f64(ps.x), _f64(ps.y), _f64(ps.w), _i64(ps.key_r), _i64(ps.key_c)
        exps = _f64(np.atleast_1d(exps))
        check(self._L.v2r_idw_stream_begin(self._h, _ptr(x), _ptr(y), _ptr(w), _ptr(kr), _ptr(kc), len(x),
                                           C.byref(grid), _ptr(exps), len(exps), precision, band_rows))
        pin = None
        try:
            cap = max(band_rows if band_rows > 0 else grid.rows, 1)
            if pinned:
                pin = PinnedArray(len(exps) * cap * max(grid.cols, 1))
            while True:
                buf = pin.array if pin is not None else np.empty(len(exps) * cap * max(grid.cols, 1))
                r0, nr = C.c_int64(), C.c_int64()
                check(self._L.v2r_idw_stream_next(self._h, _ptr(buf), C.byref(r0), C.byref(nr)))
                if nr.value == 0:
                    break
                yield r0.value, buf[:len(exps) * nr.value * grid.cols].reshape(len(exps), nr.value, grid.cols)
        finally:
            rc = self._L.v2r_idw_stream_end(self._h)
            if pin is not None:
                pin.free()
            check(rc)

    def last_error(self) -> str:
        return self._L.v2r_idw_ctx_last_error(self._h).decode(errors="replace")

    def last_timing(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        check(self._L.v2r_idw_last_timing(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"h2d_ms": a.value, "kernel_ms": b.value, "d2h_ms": c.value}


_default_ctx = None


def _ctx(ctx):
    global _default_ctx
    if ctx is not None:
        return ctx
    if _default_ctx is None:
        _default_ctx = IdwContext(1)
    return _default_ctx


def solve(ps: PointSet, xInfo: Info, yInfo: Info, exps, precision: int = FP64, rows=None, cols=None, ctx=None):
    """The whole sweep in ONE call (what cmd/idw.go:153-161 becomes): rasters[E, rows, cols]."""
    return _ctx(ctx).solve(ps, make_grid(xInfo, yInfo, rows, cols), exps, precision)


def _check_outfile(outfile: str):
    # features/idw/idw.go:19-25
    if not outfile.endswith((".asc", ".xlsx", ".tiff", ".tif")):
        raise ValueError("outfile is not supported. Use one of these files: tiff (.tif/.tiff), ascii (.asc), or excel (.xlsx).")
    if outfile.endswith(".xlsx"):
        raise ValueError("excel output is outside the GPU path (dead from the CLI: cmd/idw.go:155 always writes .tiff)")


def _go_v(v: float) -> str:
    """Go's %v of a float64 ("pow%v", features/idw/idw.go:63): shortest round-trip digits in %g style."""
    for prec in range(1, 18):
        s = f"{v:.{prec}g}"
        if float(s) == v or v != v:
            break
    return s


def _send(channel, msg):
    if channel is None:
        return
    (channel.put if hasattr(channel, "put") else channel.append)(msg)


def FullSolve(data: PointSet, outfile: str, xInfo: Info, yInfo: Info, proj: str, pow: float, channel=None, ctx=None,
              precision: int = FP64):
    """idw.FullSolve (features/idw/idw.go:15-65): whole grid in memory, one write, one completion message."""
    from . import raster_io
    start = time.time()
    _check_outfile(outfile)
    rows, cols = GetDimensions(xInfo, yInfo)
    grid = make_grid(xInfo, yInfo, rows, cols)
    z = _ctx(ctx).solve(data, grid, [pow], precision)[0]
    gd = raster_io.GDalInfo(xInfo.min, yInfo.min, xInfo.step, yInfo.step, 7, proj)
    raster_io.WriteGDAL(z, gd, outfile, (0, 0), (rows, cols), (rows, cols), True)
    _send(channel, f"pow{_go_v(pow)} [{rows}X{cols}] completed in {time.time() - start:.6f}s")
    return z


def ChunkSolve(data: PointSet, outfile: str, xInfo: Info, yInfo: Info, chunkR: int, chunkC: int, proj: str, pow: float,
               channel=None, ctx=None, precision: int = FP64):
    """idw.ChunkSolve (features/idw/idw_chunk.go:43-105).  On the GPU the chunk is a write granule:
    bands of chunkR rows stream out of the device while the next band computes, and each band is
    written in chunkC-wide windows at (row0, c0) exactly like the reference's per-chunk WriteGDAL."""
    from . import raster_io
    start = time.time()
    _check_outfile(outfile)
    rows, cols = GetDimensions(xInfo, yInfo)
    if chunkR > rows or chunkC > cols:  # idw_chunk.go:48-53
        chunkR, chunkC = rows // 4, cols // 4
    if chunkR <= 0 or chunkC <= 0:
        raise ValueError(f"chunk size ({chunkR}, {chunkC}) is not usable for a {rows}x{cols} grid")
    grid = make_grid(xInfo, yInfo, rows, cols)
    gd = raster_io.GDalInfo(xInfo.min, yInfo.min, xInfo.step, yInfo.step, 7, proj)
    full = np.empty((rows, cols))
    first = True
    for r0, band in _ctx(ctx).stream(data, grid, [pow], precision, band_rows=chunkR):
        for c0 in range(0, cols, chunkC):
            c1 = min(c0 + chunkC, cols)
            chunk = np.ascontiguousarray(band[0][:, c0:c1])
            raster_io.WriteGDAL(chunk, gd, outfile, (r0, c0), (rows, cols), chunk.shape, first)
            first = False
        full[r0:r0 + band.shape[1]] = band[0]
    _send(channel, f"pow{_go_v(pow)} [{rows}X{cols}] completed in {time.time() - start:.6f}s")
    return full
