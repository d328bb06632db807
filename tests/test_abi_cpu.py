"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol include/v2r_idw.h
declares, its host helpers agree with the oracle, and compute entry points fail loudly without a GPU
(no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import oracle as O

import v2r_b200 as V
from v2r_b200 import _lib, raster_io, read_data, workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "v2r_idw.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(v2r_idw_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    names = header_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/v2r_idw.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes binding and header disagree"
    assert L.v2r_idw_abi_version() == 2


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.IdwError) as e:
        V.IdwContext(1)
    assert e.value.code == _lib.ECUDA and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "v2r_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"(import|from|include|dlopen|CDLL).*oracle", src), f"{f} reaches into oracle/"


def test_host_helpers_match_oracle(golden_dir):
    x, y, w, proj, xi, yi = read_data.ReadTextData(os.path.join(golden_dir, "idw_in.txt"), 2284)
    ox, oy, ow, oxi, oyi = O.read_text_data(os.path.join(golden_dir, "idw_in.txt"))
    assert np.array_equal(x, ox) and np.array_equal(w, ow) and (xi.min, xi.max, yi.max) == (oxi.min, oxi.max, oyi.max)
    assert V.GetDimensions(xi, yi) == O.get_dimensions(oxi, oyi) == (16, 13)
    rng = np.random.default_rng(0)
    for _ in range(200):
        a, b = rng.uniform(-1e3, 1e3, 2)
        s = rng.choice([0.3, 1.0, 2.0, 100.0])
        i1, i2 = V.Info(a, a + abs(b) * 7, s), V.Info(b, b + abs(a) * 3, s * 1.5)
        assert V.GetDimensions(i1, i2) == O.get_dimensions(O.Info(i1.min, i1.max, i1.step), O.Info(i2.min, i2.max, i2.step))
    assert V.PointToPair(-0.5, -0.9, V.Info(0, 10, 1), V.Info(0, 10, 1)) == (0, 0)      # truncation toward zero
    assert V.PointToPair(3.7, 9.2, V.Info(0, 10, 1), V.Info(0, 10, 2)) == (4, 3)
    for es, ee, ei in ((0.5, 5.0, 0.5), (0.1, 0.75, 0.1), (1.5, 2.0, 0.5), (1.0, 3.0, 0.5)):
        assert np.array_equal(V.exp_range(es, ee, ei), O.exp_range(es, ee, ei))


def test_make_coord_space_matches_oracle_with_collisions():
    rng = np.random.default_rng(1)
    n = 20000
    xi, yi = V.Info(0, 99, 1.0), V.Info(0, 49, 2.0)
    x, y, w = rng.uniform(-3, 102, n), rng.uniform(-3, 103, n), rng.normal(0, 50, n)
    a = V.MakeCoordSpace(x, y, w, xi, yi)
    b = O.make_coord_space(x, y, w, O.Info(0, 99, 1.0), O.Info(0, 49, 2.0))
    assert len(a) == len(b) < n
    for f in ("x", "y", "w", "key_r", "key_c"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert len({(r, c) for r, c in zip(a.key_r, a.key_c)}) == len(a)     # keys unique like a Go map


def test_describe_classifies_sweeps():
    L = _lib.load()
    buf, ops = C.create_string_buffer(256), C.c_double()

    def d(exps, prec=0):
        e = np.asarray(exps, dtype=np.float64)
        _lib.check(L.v2r_idw_describe(C.c_void_p(e.ctypes.data), len(e), prec, buf, 256, C.byref(ops)))
        return buf.value.decode(), ops.value
    assert d([2.0])[0].startswith("fp64/rcp E=1") and d([2.0])[1] in (6.0, 6.5, 7.0)   # paired + staged dy^2: 6.0 issued FP64 slots per pair
    assert "rsqrt" in d([1.0, 2.0, 3.0])[0] and "rcp E=3" in d([2.0, 4.0, 6.0])[0]
    assert "root4 E=9" in d(V.exp_range(0.5, 5.0, 0.5))[0]
    assert "general" in d([1.7])[0] and "general" in d([0.0])[0] and "general" in d([-2.0])[0]
    assert "general" in d([0.1 + 0.1 + 0.1])[0]
    assert d(V.exp_range(0.5, 7.5, 0.5))[0].count("+") == 1          # 14 exponents -> two passes of 7
    assert d([3.0, 1.0, 2.0])[0].count("+") == 2                      # not a progression -> one pass each
    with pytest.raises(_lib.IdwError):
        d([float("inf")])


def test_tiff_window_writes_and_int16_translate(tmp_path):
    rng = np.random.default_rng(2)
    z = rng.normal(0, 40, (16, 13))
    z[3, 4], z[5, 6] = 45.5, -100.1
    gd = raster_io.CreateGDalInfo(-1.0, 0.0, 1.0, 1.0, 7, 'PROJCS["x",AUTHORITY["EPSG","2284"]]')
    p = str(tmp_path / "a.tiff")
    first = True
    for r0 in range(0, 16, 3):          # the ChunkSolve protocol: 3x2 chunks, create on the first one
        for c0 in range(0, 13, 2):
            chunk = z[r0:r0 + 3, c0:c0 + 2]
            raster_io.WriteGDAL(chunk, gd, p, (r0, c0), (16, 13), chunk.shape, first)
            first = False
    back, x0, y0, dx, dy = raster_io.ReadTiff(p)
    assert np.array_equal(back, z) and (x0, y0, dx, dy) == (-1.0, 0.0, 1.0, 1.0)
    raster_io.TransferType(p, str(tmp_path / "a.asc"), "Int16")
    assert open(tmp_path / "a.asc", "rb").read() == O.aaigrid_int16_bytes(z, -1.0, 0.0, 1.0)
    with pytest.raises(ValueError):
        raster_io.WriteGDAL(z, gd, str(tmp_path / "a.png"), (0, 0), (16, 13), (16, 13), True)


def test_workloads_have_the_named_shapes(tmp_path):
    for name, cells in (("c2", 4096), ("c3", 8192)):
        x, y, w, xi, yi, exps = workloads.make(name)
        assert V.GetDimensions(xi, yi) == (cells, cells)
    assert len(workloads.make("c3")[5]) == 9 and workloads.make("c5", n_points=1000)[5].tolist() == [1.0, 1.5, 2.0, 2.5]
    x, y, w, xi, yi, _ = workloads.make("c2", n_points=500)
    workloads.write_txt(str(tmp_path / "p.txt"), x, y, w, xi, yi)
    x2, y2, w2, _, xi2, yi2 = read_data.ReadTextData(str(tmp_path / "p.txt"))
    assert np.array_equal(x, x2) and np.array_equal(w, w2) and (xi2.step, yi2.max) == (100.0, yi.max)
    workloads.write_gpkg(str(tmp_path / "p.gpkg"), x, y, w)
    x3, y3, w3, proj, xi3, yi3 = read_data.ReadGeoPackage(str(tmp_path / "p.gpkg"), "wsels", "elevation", 100.0, 100.0)
    assert np.array_equal(x, x3) and np.array_equal(y, y3) and np.array_equal(w, w3) and xi3.max == x.max()
    x4, y4, w4, p4, _, _ = read_data.ReadGeoPackage(os.path.join(ROOT, "tests", "golden", "input.gpkg"), "wsels", "elevation", 1.0, 2.0)
    assert len(x4) == 14 and p4.startswith("PROJCS[")


# ---- property tests: the C ABI's host helpers (C++, idw_api.cu) against the oracle's (C, idw_oracle.c) -----------------
from hypothesis import given, settings, strategies as st  # noqa: E402

_f = st.one_of(st.floats(allow_nan=True, allow_infinity=True, width=64),
               st.floats(min_value=-1e7, max_value=1e7, allow_nan=False),
               st.sampled_from([0.0, -0.0, 1.0, 0.1, 100.0, 0.30000000000000004, 1e-300, 1e300, 9.3e18, -9.3e18]))
_step = st.one_of(st.sampled_from([1.0, 0.5, 100.0, 0.7, -1.0, 37.5, 1e-3]), st.floats(min_value=1e-6, max_value=1e6))


@settings(max_examples=400, deadline=None, derandomize=True)
@given(x=_f, y=_f, x_min=_f, y_min=_f, sx=_step, sy=_step, x_max=_f, y_max=_f)
def test_point_to_pair_and_dimensions_agree_with_the_oracle_on_any_doubles(x, y, x_min, y_min, sx, sy, x_max, y_max):
    """Go's int(float64) on amd64: truncation, NaN / out-of-range -> INT64_MIN (tools/utils.go:70-72, 133-137)"""
    L = _lib.load()
    r, c, orr, occ = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
    _lib.check(L.v2r_idw_point_to_pair(x, y, x_min, sx, y_min, sy, C.byref(r), C.byref(c)))
    O.lib().oracle_point_to_pair(x, y, x_min, sx, y_min, sy, C.byref(orr), C.byref(occ))
    assert (r.value, c.value) == (orr.value, occ.value)
    _lib.check(L.v2r_idw_get_dimensions(x_min, x_max, sx, y_min, y_max, sy, C.byref(r), C.byref(c)))
    O.lib().oracle_get_dimensions(x_min, x_max, sx, y_min, y_max, sy, C.byref(orr), C.byref(occ))
    assert (r.value, c.value) == (orr.value, occ.value)


@settings(max_examples=60, deadline=None, derandomize=True)
@given(seed=st.integers(0, 2**32 - 1), n=st.integers(0, 400), cells=st.integers(1, 12), step=st.sampled_from([1.0, 0.25, 100.0]),
       special=st.booleans())
def test_make_coord_space_agrees_with_the_oracle_on_dense_collisions(seed, n, cells, step, special):
    """few cells, many points: chains of the pairwise running average (new + old) / 2 in input order, keys outside the grid,
    negative keys, and -- with `special` -- NaN / Inf coordinates whose key is Go's INT64_MIN"""
    rng = np.random.default_rng(seed)
    ext = cells * step
    x, y, w = rng.uniform(-ext, 2 * ext, n), rng.uniform(-ext, 2 * ext, n), rng.normal(0, 1e3, n)
    if special and n >= 4:
        x[0], y[1], x[2], y[3] = np.nan, np.inf, -np.inf, np.nan
    a = V.MakeCoordSpace(x, y, w, V.Info(0.0, ext - 1, step), V.Info(0.0, ext - 1, step))
    b = O.make_coord_space(x, y, w, O.Info(0.0, ext - 1, step), O.Info(0.0, ext - 1, step))
    assert len(a) == len(b)
    for f in ("x", "y", "w", "key_r", "key_c"):
        assert np.array_equal(getattr(a, f), getattr(b, f), equal_nan=f in ("x", "y", "w")), f


@settings(max_examples=200, deadline=None, derandomize=True)
@given(es=st.floats(min_value=-5, max_value=10), span=st.floats(min_value=1e-3, max_value=12), ei=st.floats(min_value=1e-2, max_value=3))
def test_exp_range_agrees_with_the_oracle(es, span, ei):
    ee = es + span
    if not ee > es:
        return
    assert np.array_equal(V.exp_range(es, ee, ei), O.exp_range(es, ee, ei))
