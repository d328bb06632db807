"""Pins the CPU oracle against every golden vector the reference's own tests hold for the IDW path
(SURVEY.md section 8c): features/idw/idw_test.go (step1-1 golden, serial and 3x2 chunks, p = 1.7),
the orphan step2-2 golden (stale step-1 keys on a step-2 grid) and the gpkg reader expectations of
tools/processing/gpkg_test.go."""
import json
import math
import os

import numpy as np
import pytest

from oracle import oracle as O


def _load(golden_dir):
    x, y, w, xinfo, yinfo = O.read_text_data(os.path.join(golden_dir, "idw_in.txt"))
    return x, y, w, xinfo, yinfo


def test_text_reader_matches_fixture(golden_dir):
    x, y, w, xinfo, yinfo = _load(golden_dir)
    assert list(zip(x, y, w)) == [(6, 12, -100.1), (1, 7, 15.2), (4, 3, -6.3), (8, 4, 2.4), (10, 9, 45.5)]
    assert (xinfo.min, xinfo.max, xinfo.step) == (-1, 11, 1.0)   # no STEP section -> default 1.0
    assert (yinfo.min, yinfo.max, yinfo.step) == (0, 15, 1.0)
    assert O.get_dimensions(xinfo, yinfo) == (16, 13)            # idw_correct_step1-1.asc: 16 rows x 13 cols


@pytest.mark.parametrize("mode", ["serial", "conc"])
def test_golden_step1_bytes(golden_dir, mode):
    """features/idw/idw_test.go:37-76: pow 1.7, chunkR=3, chunkC=2, tiff -> Int16 AAIGrid -> byte compare."""
    x, y, w, xinfo, yinfo = _load(golden_dir)
    ps = O.make_coord_space(x, y, w, xinfo, yinfo)
    if mode == "serial":
        z = O.full_solve(ps, xinfo, yinfo, 1.7)
    else:
        z = O.chunk_solve(ps, xinfo, yinfo, 3, 2, 1.7)
    got = O.aaigrid_int16_bytes(z, xinfo.min, yinfo.min, xinfo.step)
    want = open(os.path.join(golden_dir, "idw_correct_step1-1.asc"), "rb").read()
    assert got == want


def test_golden_step2_stale_keys(golden_dir):
    """idw_correct_step2-2.asc = grid at step 2 (8x6) evaluated with the point map keyed at step 1:
    the exact-hit decision is the key table, not geometry (features/idw/idw_utils.go:13-16)."""
    x, y, w, xinfo, yinfo = _load(golden_dir)
    ps = O.make_coord_space(x, y, w, xinfo, yinfo)           # keys at step 1
    x2, y2 = O.Info(xinfo.min, xinfo.max, 2.0), O.Info(yinfo.min, yinfo.max, 2.0)
    assert O.get_dimensions(x2, y2) == (8, 6)
    z = O.full_solve(ps, x2, y2, 1.7)
    got = O.aaigrid_int16_bytes(z, x2.min, y2.min, 2.0)
    want = open(os.path.join(golden_dir, "idw_correct_step2-2.asc"), "rb").read()
    assert got == want


def test_serial_equals_chunked_bitwise(golden_dir):
    x, y, w, xinfo, yinfo = _load(golden_dir)
    ps = O.make_coord_space(x, y, w, xinfo, yinfo)
    for p in (0.5, 1.7, 2.0):
        a = O.full_solve(ps, xinfo, yinfo, p)
        for cr, cc in ((3, 2), (5, 5), (16, 13), (100, 100)):
            b = O.chunk_solve(ps, xinfo, yinfo, cr, cc, p, nthreads=3)
            assert np.array_equal(a, b)


def test_exact_hit_cells_bit_equal(golden_dir):
    x, y, w, xinfo, yinfo = _load(golden_dir)
    ps = O.make_coord_space(x, y, w, xinfo, yinfo)
    z = O.full_solve(ps, xinfo, yinfo, 0.5)
    for r, c, wt in zip(ps.key_r, ps.key_c, ps.w):
        assert z[r, c] == wt


def test_oracle_within_1e12_of_long_double_truth(golden_dir):
    x, y, w, xinfo, yinfo = _load(golden_dir)
    ps = O.make_coord_space(x, y, w, xinfo, yinfo)
    for p in (0.5, 1.0, 1.7, 2.0, 3.0, 4.5):
        z = O.full_solve(ps, xinfo, yinfo, p)
        t, scale = O.full_solve_truth(ps, xinfo, yinfo, p)
        assert O.tolerance_ok(z, t, scale, 1e-13)


def test_gpkg_reader_matches_reference_expectations(golden_dir):
    """tools/processing/gpkg_test.go:71-111: 14 points exact float64, WKT, Info min/max/step."""
    exp = json.load(open(os.path.join(golden_dir, "gpkg_points.json")))
    x, y, w, proj, xinfo, yinfo = O.read_geopackage(os.path.join(golden_dir, "input.gpkg"), "wsels", "elevation", 1.0, 2.0)
    got = [[a, b, c] for a, b, c in zip(x.tolist(), y.tolist(), w.tolist())]
    assert got == exp["points"]
    assert proj == open(os.path.join(golden_dir, "gpkg_proj.txt")).readline().rstrip("\n")
    assert (xinfo.min, xinfo.max, xinfo.step) == (exp["info"]["x"]["min"], exp["info"]["x"]["max"], 1.0)
    assert (yinfo.min, yinfo.max, yinfo.step) == (exp["info"]["y"]["min"], exp["info"]["y"]["max"], 2.0)


# ---- Go math restatement checks -------------------------------------------------------------
def test_go_pow_special_cases():
    L = O.lib()
    assert L.go_pow(0.0, -1.7) == math.inf          # d = 0 off-key -> +Inf -> NaN cell
    assert L.go_pow(0.0, -0.0) == 1.0               # p = 0 -> plain mean even at d = 0
    assert L.go_pow(0.0, 1.5) == 0.0                # negative exponent p: d = 0 -> h = 0
    assert L.go_pow(3.0, 2.0) == 9.0
    assert L.go_pow(-3.0, 2.0) == 9.0               # Pow(dx, 2) with negative dx
    assert L.go_pow(2.25, 0.5) == 1.5               # Sqrt special case
    assert L.go_pow(4.0, -0.5) == 0.5
    assert math.isnan(L.go_pow(math.nan, 2.0))
    assert L.go_pow(math.inf, -2.0) == 0.0


def test_go_pow_square_is_exact_product():
    rng = np.random.default_rng(7)
    L = O.lib()
    for v in rng.uniform(-1e7, 1e7, 2000):
        assert L.go_pow(float(v), 2.0) == float(v) * float(v)


def test_go_pow_close_to_libm():
    rng = np.random.default_rng(8)
    L = O.lib()
    d = rng.uniform(1e-3, 3e6, 4000)
    for p in (0.5, 1.0, 1.5, 1.7, 2.0, 2.5, 3.3, 4.5):
        got = np.array([L.go_pow(float(v), -p) for v in d])
        ref = np.power(d, -p)
        assert np.max(np.abs(got / ref - 1)) < 1e-15 * 4


def test_go_int_truncates_toward_zero():
    L = O.lib()
    assert L.go_int(-0.7) == 0 and L.go_int(-1.2) == -1 and L.go_int(2.99) == 2
    assert L.go_int(math.nan) == -2**63 and L.go_int(1e30) == -2**63


def test_point_to_pair_left_of_origin_lands_in_col0():
    xinfo, yinfo = O.Info(0, 10, 1), O.Info(0, 10, 1)
    ps = O.make_coord_space([-0.5, 3.2], [0.2, -0.9], [1.0, 2.0], xinfo, yinfo)
    assert ps.key_c.tolist() == [0, 3] and ps.key_r.tolist() == [0, 0]


def test_dedupe_pairwise_running_average_input_order():
    xinfo, yinfo = O.Info(0, 10, 1), O.Info(0, 10, 1)
    ps = O.make_coord_space([1.1, 1.2, 5.0, 1.9], [1.1, 1.3, 5.0, 1.8], [10.0, 20.0, 7.0, 40.0], xinfo, yinfo)
    assert len(ps) == 2
    # ((10+20)/2 + 40)/2 = 27.5 with X,Y of the last arrival (tools/utils.go:108-115)
    assert (ps.x[0], ps.y[0], ps.w[0]) == (1.9, 1.8, 27.5)
    assert (ps.x[1], ps.y[1], ps.w[1]) == (5.0, 5.0, 7.0)


def test_exp_range_accumulates_in_fp64():
    e = O.exp_range(0.1, 0.45, 0.1)
    assert e.tolist() == [0.1, 0.2, 0.1 + 0.1 + 0.1, 0.1 + 0.1 + 0.1 + 0.1]
    assert O.exp_range(0.5, 5.0, 0.5).tolist() == [0.5, 1, 1.5, 2, 2.5, 3, 3.5, 4, 4.5]
    assert O.exp_range(1.5, 2.0, 0.5).tolist() == [1.5]     # default --es 1.5 --ei .5, ee = es+ei


def test_edge_semantics():
    xinfo, yinfo = O.Info(0, 3, 1), O.Info(0, 3, 1)
    # stale key: the point sits on node (1,1) geometrically but is keyed elsewhere -> NaN there
    ps = O.PointSet(np.array([1.0, 3.0]), np.array([1.0, 2.0]), np.array([5.0, 9.0]),
                    np.array([7, 2], np.int64), np.array([7, 3], np.int64))
    z = O.full_solve(ps, xinfo, yinfo, 2.0, 4, 4)
    assert math.isnan(z[1, 1]) and z[2, 3] == 9.0 and np.isnan(z).sum() == 1
    z0 = O.full_solve(ps, xinfo, yinfo, 0.0, 4, 4)        # p = 0: plain mean everywhere off-key
    assert z0[1, 1] == 7.0 and z0[0, 0] == 7.0
    zn = O.full_solve(ps, xinfo, yinfo, -1.0, 4, 4)       # p < 0 allowed: d = 0 -> h = 0
    assert zn[1, 1] == 9.0
    empty = O.PointSet(*(np.empty(0) for _ in range(3)), np.empty(0, np.int64), np.empty(0, np.int64))
    assert np.isnan(O.full_solve(empty, xinfo, yinfo, 2.0, 4, 4)).all()   # N = 0 -> 0/0


def test_summation_order_modes_of_the_cell_solver():
    """solve_cells(mode=...): input order, reverse order (another valid execution of the reference: its Go map is
    iterated in a random order) and the same float64 terms summed in long double agree to rounding on a small case,
    exact hits are returned verbatim in every mode, and the default mode is the input-order one."""
    rng = np.random.default_rng(3)
    xi, yi = O.Info(0, 99, 1.0), O.Info(0, 79, 1.0)
    ps = O.make_coord_space(rng.uniform(0, 100, 3000), rng.uniform(0, 80, 3000), rng.normal(0, 50, 3000), xi, yi)
    rr, cc = rng.integers(0, 80, 200), rng.integers(0, 100, 200)
    rr[:5], cc[:5] = ps.key_r[:5], ps.key_c[:5]
    for p in (0.5, 1.7, 2.0, 4.5):
        a = O.solve_cells(ps, xi, yi, p, rr, cc)
        b = O.solve_cells(ps, xi, yi, p, rr, cc, mode=O.ORDER_REVERSE)
        c = O.solve_cells(ps, xi, yi, p, rr, cc, mode=O.SUM_LONG_DOUBLE)
        assert np.array_equal(a, O.solve_cells(ps, xi, yi, p, rr, cc, nthreads=1, mode=O.ORDER_INPUT))
        assert np.array_equal(a[:5], ps.w[:5]) and np.array_equal(b[:5], ps.w[:5]) and np.array_equal(c[:5], ps.w[:5])
        scale = np.maximum(np.abs(c), O.solve_cells(O.PointSet(ps.x, ps.y, np.abs(ps.w), ps.key_r, ps.key_c), xi, yi, p, rr, cc))
        assert np.max(np.abs(a - c) / scale) < 2e-13 and np.max(np.abs(b - c) / scale) < 2e-13
