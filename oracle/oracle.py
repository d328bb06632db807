"""ctypes front end of the CPU oracle (oracle/idw_oracle.c) plus the reference's file formats.

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from v2r_b200/.  See oracle/idw_oracle.c for the
parity status ("pinned to +-0.5 by the reference goldens; 1e-12 bar pinned by long-double truth").

Reference citations are relative to the v2r repository root.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import sqlite3
import struct
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libidw_oracle.so")
_lib = None

_D = C.POINTER(C.c_double)
_I = C.POINTER(C.c_int64)


def build(force: bool = False) -> str:
    """Compile oracle/libidw_oracle.so with the committed Makefile (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, "idw_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libidw_oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.go_pow.restype = C.c_double
        L.go_pow.argtypes = [C.c_double, C.c_double]
        L.go_log.restype = C.c_double
        L.go_log.argtypes = [C.c_double]
        L.go_exp.restype = C.c_double
        L.go_exp.argtypes = [C.c_double]
        L.go_int.restype = C.c_int64
        L.go_int.argtypes = [C.c_double]
        L.oracle_dist_exp.restype = C.c_double
        L.oracle_dist_exp.argtypes = [C.c_double] * 5
        L.oracle_get_dimensions.restype = None
        L.oracle_get_dimensions.argtypes = [C.c_double] * 6 + [_I, _I]
        L.oracle_point_to_pair.restype = None
        L.oracle_point_to_pair.argtypes = [C.c_double] * 6 + [_I, _I]
        L.oracle_make_coord_space.restype = C.c_int64
        L.oracle_make_coord_space.argtypes = [C.c_int64, _D, _D, _D] + [C.c_double] * 4 + [_D, _D, _D, _I, _I]
        geo = [C.c_double] * 4
        pts = [C.c_int64, _D, _D, _D, _I, _I]
        L.oracle_full_solve.restype = C.c_int
        L.oracle_full_solve.argtypes = pts + geo + [C.c_int64, C.c_int64, C.c_double, _D]
        L.oracle_solve_rows.restype = C.c_int
        L.oracle_solve_rows.argtypes = pts + geo + [C.c_int64, C.c_int64, C.c_int64, C.c_double, _D]
        L.oracle_chunk_solve.restype = C.c_int
        L.oracle_chunk_solve.argtypes = pts + geo + [C.c_int64] * 4 + [C.c_double, C.c_int, _D]
        L.oracle_chunk_solve_rows.restype = C.c_int
        L.oracle_chunk_solve_rows.argtypes = pts + geo + [C.c_int64] * 6 + [C.c_double, C.c_int, _D]
        L.oracle_full_solve_ld.restype = C.c_int
        L.oracle_full_solve_ld.argtypes = pts + geo + [C.c_int64, C.c_int64, C.c_double, _D, _D]
        L.oracle_solve_cells.restype = C.c_int
        L.oracle_solve_cells.argtypes = pts + geo + [C.c_int64, _I, _I, C.c_double, C.c_int, _D]
        L.oracle_solve_cells_mode.restype = C.c_int
        L.oracle_solve_cells_mode.argtypes = pts + geo + [C.c_int64, _I, _I, C.c_double, C.c_int, C.c_int, _D]
        L.oracle_exp_range.restype = C.c_int64
        L.oracle_exp_range.argtypes = [C.c_double] * 3 + [_D, C.c_int64]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_num_threads.argtypes = []
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_D)


def _i(a):
    return a.ctypes.data_as(_I)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


# ------------------------------------------------------------------------------- data model
@dataclass
class Info:
    """tools/utils.go:123-131 Info{Min,Max,Step}; MakeInfo() = {+Inf,-Inf,1.0}."""
    min: float = math.inf
    max: float = -math.inf
    step: float = 1.0


@dataclass
class PointSet:
    """The de-duplicated map[OrderedPair]Point flattened to SoA + keys (tools/utils.go:46-55)."""
    x: np.ndarray
    y: np.ndarray
    w: np.ndarray
    key_r: np.ndarray
    key_c: np.ndarray

    def __len__(self):
        return len(self.x)


def get_dimensions(xinfo: Info, yinfo: Info):
    """tools/utils.go:133-137."""
    r, c = C.c_int64(), C.c_int64()
    lib().oracle_get_dimensions(xinfo.min, xinfo.max, xinfo.step, yinfo.min, yinfo.max, yinfo.step,
                                C.byref(r), C.byref(c))
    return r.value, c.value


def make_coord_space(x, y, w, xinfo: Info, yinfo: Info) -> PointSet:
    """tools/utils.go:97-119 in input order."""
    x, y, w = _f64(x), _f64(y), _f64(w)
    n = len(x)
    ox, oy, ow = np.empty(n), np.empty(n), np.empty(n)
    kr, kc = np.empty(n, np.int64), np.empty(n, np.int64)
    m = lib().oracle_make_coord_space(n, _d(x), _d(y), _d(w), xinfo.min, xinfo.step, yinfo.min,
                                      yinfo.step, _d(ox), _d(oy), _d(ow), _i(kr), _i(kc))
    if m < 0:
        raise MemoryError("oracle_make_coord_space")
    return PointSet(ox[:m].copy(), oy[:m].copy(), ow[:m].copy(), kr[:m].copy(), kc[:m].copy())


def exp_range(es: float, ee: float, ei: float):
    """cmd/idw.go:153: FP64-accumulated exponent list."""
    buf = np.empty(4096)
    n = lib().oracle_exp_range(es, ee, ei, _d(buf), len(buf))
    return buf[:min(n, len(buf))].copy()


def _args(ps: PointSet, xinfo: Info, yinfo: Info):
    x, y, w, kr, kc = _f64(ps.x), _f64(ps.y), _f64(ps.w), _i64(ps.key_r), _i64(ps.key_c)
    keep = (x, y, w, kr, kc)
    return keep, [len(x), _d(x), _d(y), _d(w), _i(kr), _i(kc), xinfo.min, xinfo.step, yinfo.min, yinfo.step]


def full_solve(ps: PointSet, xinfo: Info, yinfo: Info, exp: float, rows=None, cols=None):
    """features/idw/idw.go:15-65 minus the writer: the rows x cols float64 raster."""
    if rows is None:
        rows, cols = get_dimensions(xinfo, yinfo)
    out = np.empty((rows, cols))
    keep, a = _args(ps, xinfo, yinfo)
    rc = lib().oracle_full_solve(*a, rows, cols, exp, _d(out))
    assert rc == 0
    return out


def solve_rows(ps: PointSet, xinfo: Info, yinfo: Info, exp: float, cols: int, row0: int, nrows: int):
    out = np.empty((nrows, cols))
    keep, a = _args(ps, xinfo, yinfo)
    rc = lib().oracle_solve_rows(*a, cols, row0, nrows, exp, _d(out))
    assert rc == 0
    return out


def chunk_solve(ps: PointSet, xinfo: Info, yinfo: Info, chunk_r: int, chunk_c: int, exp: float,
                rows=None, cols=None, nthreads: int = 0):
    """features/idw/idw_chunk.go:43-105 minus the writer."""
    if rows is None:
        rows, cols = get_dimensions(xinfo, yinfo)
    out = np.full((rows, cols), np.nan)
    keep, a = _args(ps, xinfo, yinfo)
    rc = lib().oracle_chunk_solve(*a, rows, cols, chunk_r, chunk_c, exp, nthreads, _d(out))
    assert rc == 0, rc
    return out


def chunk_solve_rows(ps: PointSet, xinfo: Info, yinfo: Info, rows: int, cols: int, row0: int, nrows: int,
                     chunk_r: int, chunk_c: int, exp: float, nthreads: int = 0):
    out = np.full((nrows, cols), np.nan)
    keep, a = _args(ps, xinfo, yinfo)
    rc = lib().oracle_chunk_solve_rows(*a, rows, cols, row0, nrows, chunk_r, chunk_c, exp, nthreads, _d(out))
    assert rc == 0, rc
    return out


ORDER_INPUT, ORDER_REVERSE, SUM_LONG_DOUBLE = 0, 1, 2


def solve_cells(ps: PointSet, xinfo: Info, yinfo: Info, exp: float, rr, cc, nthreads: int = 0, mode: int = ORDER_INPUT):
    """calculateIDW on a list of cells (for spot checks of full-size rasters).  mode: ORDER_INPUT = points in array
    order (the oracle's fixed choice); ORDER_REVERSE = the same float64 arithmetic in reverse order -- another valid
    execution of the reference, whose Go map iteration is randomised per `range`; SUM_LONG_DOUBLE = the reference's
    float64 terms summed in 80-bit long double, the order-independent value its executions scatter around."""
    rr, cc = _i64(rr), _i64(cc)
    out = np.empty(len(rr))
    keep, a = _args(ps, xinfo, yinfo)
    rc = lib().oracle_solve_cells_mode(*a, len(rr), _i(rr), _i(cc), exp, nthreads or num_threads(), mode, _d(out))
    assert rc == 0
    return out


def full_solve_truth(ps: PointSet, xinfo: Info, yinfo: Info, exp: float, rows=None, cols=None):
    """80-bit long double evaluation; returns (raster, scale) with scale = sum|w|h / sum h."""
    if rows is None:
        rows, cols = get_dimensions(xinfo, yinfo)
    out, scale = np.empty((rows, cols)), np.empty((rows, cols))
    keep, a = _args(ps, xinfo, yinfo)
    rc = lib().oracle_full_solve_ld(*a, rows, cols, exp, _d(out), _d(scale))
    assert rc == 0
    return out, scale


def num_threads() -> int:
    return lib().oracle_num_threads()


# --------------------------------------------------------------------------- file formats
def _float_or_zero(tok: str) -> float:
    """p.X, _ = strconv.ParseFloat(tok, 64) (read_data.go:104-106): a token Go cannot parse leaves 0; an out-of-range
    one leaves +-Inf (ErrRange is dropped with the other errors).  Python's float() accepts three spellings Go does
    not -- underscores, "nan(...)" and a signed nan -- and Go accepts hex floats (with a p exponent) that float() does not."""
    t = tok.lower()
    if "_" in t or "(" in t or t in ("+nan", "-nan"):
        return 0.0
    try:
        if t.lstrip("+-").startswith("0x"):
            return float.fromhex(tok) if "p" in t else 0.0
        return float(tok)
    except (ValueError, OverflowError):
        return 0.0


def read_text_data(path: str):
    """tools/processing/read_data.go:17-111 grammar: `POINTS n` + n lines `x y w`; optional `STEP`
    + `sx sy`; `ESTIMATE` + `xmin xmax` + `ymin ymax` (returns at ESTIMATE).  Returns
    (x, y, w, xinfo, yinfo).  A file without ESTIMATE is an error (:89)."""
    xinfo, yinfo = Info(), Info()
    xs, ys, ws = [], [], []
    with open(path) as f:
        lines = f.read().split("\n")
    if lines and lines[-1] == "":
        lines.pop()  # bufio.Scanner does not yield a trailing empty token
    it = iter(lines)
    for line in it:
        head = line.split()[0]  # a blank line panics in the reference (:40); IndexError here
        if head == "POINTS":
            # addPoints (:93-111): Atoi and ParseFloat errors are dropped (0 points / 0.0), and `data = addPoints(sc)` (:42)
            # REPLACES whatever an earlier POINTS section read
            try:
                n = int(line.removeprefix("POINTS ").strip())
            except ValueError:
                n = 0
            xs, ys, ws = [], [], []
            for _ in range(n):
                f3 = next(it).split()
                xs.append(_float_or_zero(f3[0])); ys.append(_float_or_zero(f3[1])); ws.append(_float_or_zero(f3[2]))
        elif head == "STEP":
            f2 = next(it).split()
            xinfo.step, yinfo.step = float(f2[0]), float(f2[1])
        elif head == "ESTIMATE":
            a = next(it).split()
            b = next(it).split()
            xinfo.min, xinfo.max = float(a[0]), float(a[1])
            yinfo.min, yinfo.max = float(b[0]), float(b[1])
            return np.array(xs), np.array(ys), np.array(ws), xinfo, yinfo
    raise ValueError("ESTIMATE not in file")


def read_geopackage(path: str, layer: str, field: str, step_x: float, step_y: float):
    """tools/processing/read_data.go:149-173 + io_gdal.go:161-197: features by fid 1..count,
    geom.X(0), geom.Y(0), FieldAsFloat64(field); bbox -> Info; steps from the flags.  Decodes the
    GeoPackageBinary header by hand (no OGR here).  Returns (x, y, w, proj_wkt, xinfo, yinfo)."""
    con = sqlite3.connect(f"file:{path}?mode=ro", uri=True)
    try:
        tables = [r[0] for r in con.execute("select table_name from gpkg_contents")]
        if layer not in tables:
            raise ValueError(f"layer {layer!r} not in {tables}")
        geom_col, srs_id = con.execute(
            "select column_name, srs_id from gpkg_geometry_columns where table_name=?", (layer,)).fetchone()
        proj = con.execute("select definition from gpkg_spatial_ref_sys where srs_id=?", (srs_id,)).fetchone()[0]
        rows = con.execute(f'select "{geom_col}", "{field}" from "{layer}" order by rowid').fetchall()
    finally:
        con.close()
    xs, ys, ws = [], [], []
    for blob, val in rows:
        assert blob[:2] == b"GP"
        flags = blob[3]
        env = (flags >> 1) & 7
        env_bytes = {0: 0, 1: 32, 2: 48, 3: 48, 4: 64}[env]
        wkb = blob[8 + env_bytes:]
        bo = "<" if wkb[0] == 1 else ">"
        (gtype,) = struct.unpack(bo + "I", wkb[1:5])
        assert gtype % 1000 == 1, "POINT expected"
        x, y = struct.unpack(bo + "dd", wkb[5:21])
        xs.append(x); ys.append(y); ws.append(float(val))
    x, y, w = np.array(xs), np.array(ys), np.array(ws)
    xinfo = Info(float(x.min()), float(x.max()), step_x)
    yinfo = Info(float(y.min()), float(y.max()), step_y)
    return x, y, w, proj, xinfo, yinfo


def round_half_away_int16(a: np.ndarray) -> np.ndarray:
    """gdal_translate -ot Int16 float->int rule (io_gdal.go:115-132 via TransferType): round half
    away from zero, clamp to the Int16 range; NaN -> 0 (GDAL behaviour)."""
    a = np.asarray(a, dtype=np.float64)
    r = np.where(np.isnan(a), 0.0, np.sign(a) * np.floor(np.abs(a) + 0.5))
    return np.clip(r, -32768, 32767).astype(np.int16)


def aaigrid_int16_bytes(raster: np.ndarray, x_min: float, y_min: float, cellsize: float) -> bytes:
    """The byte layout of the reference's golden .asc files (features/idw/idw_test.go:57 ->
    GDAL AAIGrid writer on the south-up geotransform {xMin,xStep,0,yMin,0,+yStep}):
    yllcorner = yMin - rows*cellsize; rows emitted in raster order (row 0 = y_min first)."""
    rows, cols = raster.shape
    q = round_half_away_int16(raster)
    out = [f"ncols        {cols}", f"nrows        {rows}", f"xllcorner    {x_min:.12f}",
           f"yllcorner    {y_min - rows * cellsize:.12f}", f"cellsize     {cellsize:.12f}"]
    for r in range(rows):
        out.append("".join(f" {int(v)}" for v in q[r]))
    return ("\n".join(out) + "\n").encode()


def tolerance_ok(z, z_ref, scale, rel=1e-12):
    """|z - z_ref| <= rel * max(|z_ref|, sum|w|h/sum h); NaN positions equal."""
    z, z_ref, scale = np.asarray(z), np.asarray(z_ref), np.asarray(scale)
    nan_a, nan_b = np.isnan(z), np.isnan(z_ref)
    if not np.array_equal(nan_a, nan_b):
        return False
    ok = ~nan_b
    return bool(np.all(np.abs(z[ok] - z_ref[ok]) <= rel * np.maximum(np.abs(z_ref[ok]), scale[ok])))
