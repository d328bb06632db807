"""Parity of the CUDA path (through the C ABI) against the CPU oracle, the reference's golden rasters
and size-independent properties.  Runs on a B200 only (`-m gpu`).

Tolerance (SURVEY.md section 8c, FP64 bar of north_star = 1e-12 relative):
    |z_gpu - z_ref| <= 1e-12 * max(|z_ref|, sum|w_i| h_i / sum h_i)
NaN positions must match exactly; exact-hit cells must be bit-equal to the stored weight.
"""
import os
import queue

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402  (checker only)

import v2r_b200 as V  # noqa: E402
from v2r_b200 import _lib, raster_io, read_data, workloads  # noqa: E402
from v2r_b200.idw import make_grid  # noqa: E402

REL = 1e-12


@pytest.fixture(scope="module")
def ctx():
    c = V.IdwContext(1)
    yield c
    c.close()


def o_info(i):
    return O.Info(i.min, i.max, i.step)


def o_ps(ps):
    return O.PointSet(ps.x, ps.y, ps.w, ps.key_r, ps.key_c)


def oracle_ref(ps, xi, yi, p, rows=None, cols=None):
    """(reference raster, tolerance scale sum|w|h/sum h) from the threaded oracle (bit-identical to its
    serial FullSolve, tests/test_oracle_golden.py); the scale is the same IDW run on |w|."""
    if rows is None:
        rows, cols = O.get_dimensions(o_info(xi), o_info(yi))
    cr, cc = max(1, min(8, rows)), max(1, min(64, cols))      # (an oversize chunk makes ChunkSolve quarter the grid)
    ref = O.chunk_solve(o_ps(ps), o_info(xi), o_info(yi), cr, cc, p, rows, cols)
    absw = O.PointSet(ps.x, ps.y, np.abs(ps.w), ps.key_r, ps.key_c)
    scale = O.chunk_solve(absw, o_info(xi), o_info(yi), cr, cc, p, rows, cols)
    return ref, scale


def assert_parity(z, z_ref, scale, rel=REL, what=""):
    z, z_ref, scale = np.asarray(z), np.asarray(z_ref), np.asarray(scale)
    assert z.shape == z_ref.shape
    assert np.array_equal(np.isnan(z), np.isnan(z_ref)), f"{what}: NaN positions differ"
    ok = ~np.isnan(z_ref)
    err = np.abs(z[ok] - z_ref[ok]) / np.maximum(np.maximum(np.abs(z_ref[ok]), scale[ok]), 1e-300)
    assert err.size == 0 or err.max() <= rel, f"{what}: max scaled error {err.max():.3e} > {rel:g}"
    return float(err.max()) if err.size else 0.0


def assert_parity_cells(got, ps, xi, yi, p, rr, cc, scale_ps=None, rel=REL, what=""):
    """Sampled cells of a full-size raster.  With 1e5..2e6 points the reference's OWN result depends on the order in
    which its Go map happens to be iterated by more than 1e-12 for large exponents (c3, p = 4.5: input order and
    reverse order, two valid executions, differ by 1.9e-12): once the running sum is large, thousands of tiny terms
    are rounded away.  So the check is made against the order-independent value -- the reference's float64 terms
    (same DistExp, same w*h) summed in long double -- at 1e-12, and against the oracle's input-order execution with
    that execution's own distance from the exact sum as allowance."""
    exact = O.solve_cells(o_ps(ps), o_info(xi), o_info(yi), float(p), rr, cc, mode=O.SUM_LONG_DOUBLE)
    seq = O.solve_cells(o_ps(ps), o_info(xi), o_info(yi), float(p), rr, cc)
    scale = np.abs(exact) if scale_ps is None else O.solve_cells(o_ps(scale_ps), o_info(xi), o_info(yi), float(p), rr, cc)
    err = assert_parity(got, exact, scale, rel=rel, what=what + " vs exact sum of the reference's terms")
    ok = ~np.isnan(seq)
    s = np.maximum(np.maximum(np.abs(seq[ok]), scale[ok]), 1e-300)
    assert np.all(np.abs(np.asarray(got)[ok] - seq[ok]) / s <= rel + np.abs(seq[ok] - exact[ok]) / s), f"{what}: vs input-order oracle"
    return err


def random_case(seed, n, rows, cols, step=1.0, x0=0.0, y0=0.0, signed=True):
    rng = np.random.default_rng(seed)
    xi = V.Info(x0, x0 + cols * step - 1, step)
    yi = V.Info(y0, y0 + rows * step - 1, step)
    x = rng.uniform(xi.min, xi.min + cols * step, n)
    y = rng.uniform(yi.min, yi.min + rows * step, n)
    w = rng.uniform(-100, 100, n) if signed else rng.uniform(0, 500, n)
    ps = V.MakeCoordSpace(x, y, w, xi, yi)
    return ps, xi, yi


def load_golden(golden_dir):
    x, y, w, proj, xi, yi = read_data.ReadTextData(os.path.join(golden_dir, "idw_in.txt"), 2284)
    return V.MakeCoordSpace(x, y, w, xi, yi), xi, yi, proj


# --------------------------------------------------------------------------- reference goldens
@pytest.mark.parametrize("cs", ["Serial", "Conc"])
def test_golden_step1_through_gpu_path(golden_dir, tmp_path, ctx, cs):
    """features/idw/idw_test.go:14-76 replayed: FullSolve / ChunkSolve(3,2) at pow 1.7 -> tiff ->
    Int16 AAIGrid -> byte compare with idw_correct_step1-1.asc."""
    data, xi, yi, proj = load_golden(golden_dir)
    outfile = str(tmp_path / f"idw_step1-1pow1.7{'chunked' if cs == 'Conc' else ''}")
    channel = queue.Queue()
    if cs == "Conc":
        V.ChunkSolve(data, outfile + ".tiff", xi, yi, 3, 2, proj, 1.7, channel, ctx=ctx)
    else:
        V.FullSolve(data, outfile + ".tiff", xi, yi, proj, 1.7, channel, ctx=ctx)
    assert "completed in" in channel.get_nowait()
    raster_io.TransferType(outfile + ".tiff", outfile + ".asc", "Int16")
    assert raster_io.SameFiles(outfile + ".asc", os.path.join(golden_dir, "idw_correct_step1-1.asc"))


def test_golden_step2_stale_keys_through_gpu_path(golden_dir, tmp_path, ctx):
    """idw_correct_step2-2.asc: step-2 grid, keys made at step 1 (exact hit = key table, not geometry)."""
    data, xi, yi, proj = load_golden(golden_dir)
    x2, y2 = V.Info(xi.min, xi.max, 2.0), V.Info(yi.min, yi.max, 2.0)
    out = str(tmp_path / "s2")
    V.FullSolve(data, out + ".tiff", x2, y2, proj, 1.7, None, ctx=ctx)
    raster_io.TransferType(out + ".tiff", out + ".asc", "Int16")
    assert raster_io.SameFiles(out + ".asc", os.path.join(golden_dir, "idw_correct_step2-2.asc"))


@pytest.mark.parametrize("p", [1.7, 0.5, 2.0, 1.0, 3.0, 1.5, 0.3, 4.5, 6.0])
def test_c1_vs_oracle_and_truth(golden_dir, ctx, p):
    data, xi, yi, _ = load_golden(golden_dir)
    z = V.solve(data, xi, yi, [p], ctx=ctx)[0]
    ref = O.full_solve(o_ps(data), o_info(xi), o_info(yi), p)
    truth, scale = O.full_solve_truth(o_ps(data), o_info(xi), o_info(yi), p)
    assert_parity(z, ref, scale, what=f"c1 p={p} vs oracle")
    assert_parity(z, truth, scale, what=f"c1 p={p} vs long-double truth")
    for r, c, w in zip(data.key_r, data.key_c, data.w):
        assert z[r, c] == w  # exact hit: bit-equal


# --------------------------------------------------------------------------- random cases
CLASSES = [2.0, 4.0, 1.0, 3.0, 5.0, 0.5, 1.5, 2.5, 1.7, 0.3, 0.30000000000000004, 7.25, 0.0, -1.0, -2.0]


@pytest.mark.parametrize("p", CLASSES)
def test_random_1k_points_256_grid(ctx, p):
    ps, xi, yi = random_case(11, 1000, 256, 256, step=3.0, x0=-50.0, y0=1e4)  # 65 k cells x 1 k points
    z = V.solve(ps, xi, yi, [p], ctx=ctx)[0]
    ref, scale = oracle_ref(ps, xi, yi, p)
    assert_parity(z, ref, scale, what=f"p={p}")


@pytest.mark.parametrize("n", [0, 1, 2, 511, 512, 513, 1024, 1025, 1537])
def test_tile_boundaries_and_ragged_grids(ctx, n):
    """point counts around the 512-point TMA tile, grid sizes that are not multiples of the CTA tile"""
    ps, xi, yi = random_case(100 + n, n, 37, 131, step=0.7)
    for p in (2.0, 1.5, 1.7):
        z = V.solve(ps, xi, yi, [p], ctx=ctx)[0]
        ref, scale = oracle_ref(ps, xi, yi, p)
        assert_parity(z, ref, scale, what=f"n={n} p={p}")


def test_sweeps_match_single_exponent_solves(ctx):
    ps, xi, yi = random_case(5, 700, 96, 160, step=2.0)
    for exps in (V.exp_range(0.5, 5.0, 0.5), V.exp_range(1.0, 3.0, 0.5), V.exp_range(2.0, 9.0, 2.0),
                 V.exp_range(1.0, 5.0, 1.0), V.exp_range(0.1, 0.75, 0.1), np.array([3.0, 1.0, 2.0]),
                 V.exp_range(0.5, 7.5, 0.5), V.exp_range(1.1, 3.4, 0.2)):
        zs = V.solve(ps, xi, yi, exps, ctx=ctx)
        assert zs.shape == (len(exps), 96, 160)
        for e, p in enumerate(exps):
            ref, scale = oracle_ref(ps, xi, yi, float(p))
            assert_parity(zs[e], ref, scale, what=f"sweep {exps} e={e}")


def test_edge_semantics_match_reference(ctx):
    xi, yi = V.Info(0, 3, 1), V.Info(0, 3, 1)
    # stale key: the point sits on node (1,1) but is keyed outside the grid -> geometric d = 0 -> NaN
    ps = V.PointSet(np.array([1.0, 3.0]), np.array([1.0, 2.0]), np.array([5.0, 9.0]),
                    np.array([7, 2], np.int64), np.array([7, 3], np.int64))
    for p in (2.0, 1.0, 0.5, 1.7):
        z = V.solve(ps, xi, yi, [p], rows=4, cols=4, ctx=ctx)[0]
        ref = O.full_solve(o_ps(ps), o_info(xi), o_info(yi), p, 4, 4)
        assert np.isnan(z[1, 1]) and np.isnan(z).sum() == 1 and z[2, 3] == 9.0
        assert np.array_equal(np.isnan(z), np.isnan(ref))
    z0 = V.solve(ps, xi, yi, [0.0], rows=4, cols=4, ctx=ctx)[0]      # p = 0: plain mean, even at d = 0
    assert z0[1, 1] == 7.0 and z0[0, 0] == 7.0 and z0[2, 3] == 9.0
    zn = V.solve(ps, xi, yi, [-1.0], rows=4, cols=4, ctx=ctx)[0]     # p < 0: d = 0 -> h = 0
    refn = O.full_solve(o_ps(ps), o_info(xi), o_info(yi), -1.0, 4, 4)
    assert zn[1, 1] == 9.0 and np.allclose(zn, refn, rtol=1e-13)
    empty = V.PointSet(*(np.empty(0) for _ in range(3)), np.empty(0, np.int64), np.empty(0, np.int64))
    assert np.isnan(V.solve(empty, xi, yi, [2.0], rows=4, cols=4, ctx=ctx)).all()  # N = 0 -> 0/0


def test_points_left_of_origin_and_negative_keys(ctx):
    xi, yi = V.Info(0, 20, 1), V.Info(0, 20, 1)
    x = np.array([-0.5, 3.2, -7.0, 25.0])
    y = np.array([0.2, -0.9, 4.0, 30.0])
    w = np.array([1.0, 2.0, 3.0, 4.0])
    ps = V.MakeCoordSpace(x, y, w, xi, yi)
    ops = O.make_coord_space(x, y, w, o_info(xi), o_info(yi))
    assert ps.key_r.tolist() == ops.key_r.tolist() and ps.key_c.tolist() == ops.key_c.tolist()
    z = V.solve(ps, xi, yi, [2.0], ctx=ctx)[0]
    ref = O.full_solve(ops, o_info(xi), o_info(yi), 2.0)
    _, scale = O.full_solve_truth(ops, o_info(xi), o_info(yi), 2.0)
    assert_parity(z, ref, scale)
    assert z[0, 0] == 1.0 and z[0, 3] == 2.0     # truncation toward zero pulls them into row/col 0


def test_row_bands_and_stream_equal_full(ctx):
    ps, xi, yi = random_case(21, 900, 150, 70, step=5.0)
    grid = make_grid(xi, yi)
    exps = [1.0, 1.5, 2.0]
    full = ctx.solve(ps, grid, exps)
    parts = np.concatenate([ctx.solve_rows(ps, grid, r0, min(40, grid.rows - r0), exps) for r0 in range(0, grid.rows, 40)], axis=1)
    assert np.array_equal(full, parts)
    seen = np.full_like(full, np.nan)
    for r0, band in ctx.stream(ps, grid, exps, band_rows=33):
        seen[:, r0:r0 + band.shape[1]] = band
    assert np.array_equal(full, seen)
    t = ctx.last_timing()
    assert t["kernel_ms"] > 0


def test_epsg2284_magnitudes(ctx):
    """coordinates ~1.2e7 ft like the gpkg fixture (tools/processing/gpkg_test.go:14-27)"""
    ps, xi, yi = random_case(31, 800, 120, 120, step=100.0, x0=1.2020401943588393e+07, y0=3.885548111378168e+06, signed=False)
    for p in (2.0, 1.5):
        z = V.solve(ps, xi, yi, [p], ctx=ctx)[0]
        ref, scale = oracle_ref(ps, xi, yi, p)
        assert_parity(z, ref, scale, what=f"epsg2284 p={p}")


def test_gpkg_fixture_end_to_end(golden_dir, ctx):
    x, y, w, proj, xi, yi = read_data.ReadGeoPackage(os.path.join(golden_dir, "input.gpkg"), "wsels", "elevation", 500.0, 500.0)
    assert len(x) == 14 and proj.startswith("PROJCS[")
    ps = V.MakeCoordSpace(x, y, w, xi, yi)
    z = V.solve(ps, xi, yi, [1.5, 2.0], ctx=ctx)
    for e, p in enumerate((1.5, 2.0)):
        ref, scale = oracle_ref(ps, xi, yi, p)
        assert_parity(z[e], ref, scale, what=f"gpkg p={p}")


def test_error_behaviour(ctx):
    ps, xi, yi = random_case(1, 10, 8, 8)
    grid = make_grid(xi, yi)
    with pytest.raises(_lib.IdwError) as e:
        ctx.solve(ps, grid, [float("nan")])
    assert e.value.code == _lib.EINVAL
    with pytest.raises(_lib.IdwError):
        ctx.solve_rows(ps, grid, 4, 10, [2.0], out=np.empty((1, 10, 8)))   # rows outside the grid
    with pytest.raises(_lib.IdwError):
        ctx.solve(ps, grid, np.ones(4097))                                   # too many exponents
    with pytest.raises(ValueError):
        V.FullSolve(ps, "out.png", xi, yi, "", 2.0, ctx=ctx)                # idw.go:19-25 suffix check
    with pytest.raises(_lib.IdwError):
        V.exp_range(1.0, 1.0, 0.5)                                          # cmd/idw.go:86-88
    with pytest.raises(_lib.IdwError):
        V.exp_range(1.0, 2.0, 0.0)                                          # cmd/idw.go:83-85


# --------------------------------------------------------------------------- full-size properties
@pytest.fixture(scope="module")
def c2():
    x, y, w, xi, yi, exps = workloads.make("c2")
    return V.MakeCoordSpace(x, y, w, xi, yi), xi, yi, exps


def test_c2_full_size_spot_check_and_properties(ctx, c2):
    """BASELINE.json config 2 (100 k points -> 4096^2, p = 2) at full size:
    (a) 5000 random cells, the four corners and every 97th exact-hit cell against the oracle,
    (b) constant weights reproduce the constant, (c) linearity in the weights,
    (d) row bands computed separately are bit-identical to the full raster."""
    ps, xi, yi, exps = c2
    grid = make_grid(xi, yi)
    assert (grid.rows, grid.cols) == (4096, 4096)
    z = ctx.solve(ps, grid, exps)[0]
    assert not np.isnan(z).any()
    rng = np.random.default_rng(3)
    rr = np.concatenate([rng.integers(0, grid.rows, 5000), ps.key_r[::97], [0, 0, grid.rows - 1, grid.rows - 1]])
    cc = np.concatenate([rng.integers(0, grid.cols, 5000), ps.key_c[::97], [0, grid.cols - 1, 0, grid.cols - 1]])
    assert_parity_cells(z[rr, cc], ps, xi, yi, 2.0, rr, cc, what="c2 spot check")
    assert np.array_equal(z[ps.key_r, ps.key_c], ps.w)           # all 99.7 k exact-hit cells bit-equal
    # (b) constant field
    const = V.PointSet(ps.x, ps.y, np.full(len(ps), 123.456), ps.key_r, ps.key_c)
    zc = ctx.solve_rows(const, grid, 1000, 64, exps)[0]
    assert np.max(np.abs(zc / 123.456 - 1)) < 1e-13
    # (c) linearity: z(2 w1 - 3 w2) = 2 z(w1) - 3 z(w2) within the tolerance of the combined magnitudes
    w2 = np.random.default_rng(4).uniform(0, 500, len(ps))
    band = (2048, 32)
    z1 = ctx.solve_rows(ps, grid, *band, exps)[0]
    z2 = ctx.solve_rows(V.PointSet(ps.x, ps.y, w2, ps.key_r, ps.key_c), grid, *band, exps)[0]
    z3 = ctx.solve_rows(V.PointSet(ps.x, ps.y, 2 * ps.w - 3 * w2, ps.key_r, ps.key_c), grid, *band, exps)[0]
    assert np.max(np.abs(z3 - (2 * z1 - 3 * z2))) <= 1e-12 * 2500
    # (d) bands
    assert np.array_equal(z[2048:2080], z1)


def test_c3_sweep_full_width_band(ctx):
    """BASELINE.json config 3 (250 k points -> 8192^2, p = 0.5..4.5, one pass): three 32-row bands of the
    full-width grid (first, middle, last rows) against the oracle on 510 sampled cells, all nine exponents."""
    x, y, w, xi, yi, exps = workloads.make("c3")
    ps = V.MakeCoordSpace(x, y, w, xi, yi)
    grid = make_grid(xi, yi)
    assert (grid.rows, grid.cols, len(exps)) == (8192, 8192, 9)
    rng = np.random.default_rng(5)
    for r0 in (0, 4096, grid.rows - 32):                                       # first rows, middle, last rows
        z = ctx.solve_rows(ps, grid, r0, 32, exps)
        rr, cc = r0 + rng.integers(0, 32, 170), rng.integers(0, grid.cols, 170)
        rr[:2], cc[:2] = (r0, r0 + 31), (0, grid.cols - 1)
        for e, p in enumerate(exps):                                           # 510 cells x 9 exponents
            assert_parity_cells(z[e][rr - r0, cc], ps, xi, yi, p, rr, cc, what=f"c3 rows {r0} p={p}")   # w >= 0: scale == |z|


# --------------------------------------------------------------------------- FP32 fast path
REL32 = 2e-5   # stated accuracy of the FP32 path: ~1e-5 relative (north_star); tests allow 2e-5


@pytest.mark.parametrize("p", [2.0, 1.0, 3.0, 4.0, 0.5, 1.5, 1.7, 0.3])
def test_fp32_path_random(ctx, p):
    ps, xi, yi = random_case(41, 2000, 200, 300, step=25.0, signed=False)
    z = V.solve(ps, xi, yi, [p], precision=_lib.FP32, ctx=ctx)[0]
    ref, scale = oracle_ref(ps, xi, yi, p)
    err = assert_parity(z, ref, scale, rel=REL32, what=f"fp32 p={p}")
    assert err > 1e-12 or p == 0.0   # really the FP32 kernels, not the FP64 ones
    assert np.array_equal(z[ps.key_r, ps.key_c], ps.w)   # exact hits stay bit-equal


def test_fp32_path_epsg2284_magnitudes_and_signed_weights(ctx):
    """1.2e7 ft coordinates: a plain float has a 1 ft ulp there; the CTA-local hi/lo split must hold 2e-5"""
    ps, xi, yi = random_case(43, 3000, 160, 160, step=100.0, x0=1.2020401943588393e+07, y0=3.885548111378168e+06)
    for exps in ([2.0], [1.0, 1.5, 2.0, 2.5]):
        z = V.solve(ps, xi, yi, exps, precision=_lib.FP32, ctx=ctx)
        for e, p in enumerate(exps):
            ref, scale = oracle_ref(ps, xi, yi, p)
            assert_parity(z[e], ref, scale, rel=REL32, what=f"fp32 epsg2284 p={p}")


def test_fp32_tile_tails_and_nan(ctx):
    for n in (0, 1, 2, 3, 511, 513, 1025):
        ps, xi, yi = random_case(300 + n, n, 37, 131, step=0.7)
        z = V.solve(ps, xi, yi, [2.0], precision=_lib.FP32, ctx=ctx)[0]
        ref, scale = oracle_ref(ps, xi, yi, 2.0)
        assert_parity(z, ref, scale, rel=REL32, what=f"fp32 n={n}")
    xi, yi = V.Info(0, 3, 1), V.Info(0, 3, 1)
    ps = V.PointSet(np.array([1.0, 3.0]), np.array([1.0, 2.0]), np.array([5.0, 9.0]),
                    np.array([7, 2], np.int64), np.array([7, 3], np.int64))
    z = V.solve(ps, xi, yi, [2.0], rows=4, cols=4, precision=_lib.FP32, ctx=ctx)[0]
    assert np.isnan(z[1, 1]) and np.isnan(z).sum() == 1 and z[2, 3] == 9.0


def test_fp32_c2_full_size_spot_check(ctx, c2):
    ps, xi, yi, exps = c2
    grid = make_grid(xi, yi)
    z = ctx.solve(ps, grid, exps, _lib.FP32)[0]
    rng = np.random.default_rng(6)
    rr, cc = rng.integers(0, grid.rows, 300), rng.integers(0, grid.cols, 300)
    ref = O.solve_cells(o_ps(ps), o_info(xi), o_info(yi), 2.0, rr, cc)
    assert_parity(z[rr, cc], ref, np.abs(ref), rel=REL32, what="fp32 c2 spot check")
    assert np.array_equal(z[ps.key_r, ps.key_c], ps.w)


# --------------------------------------------------------------------------- c4 / c5 shapes, randomized sweep
def test_c4_clustered_full_width_band(ctx):
    """BASELINE.json config 4 (1 M clustered points -> 16384^2, p = 2): five 32-row full-width bands (first
    to last rows, 2100 cells) through the paired-reciprocal kernel and the FP32 path against the oracle, then a pass
    with signed weights; dedupe merges clustered points."""
    x, y, w, xi, yi, exps = workloads.make("c4")
    ps = V.MakeCoordSpace(x, y, w, xi, yi)
    ops = O.make_coord_space(x, y, w, o_info(xi), o_info(yi))
    assert len(ps) == len(ops) < len(x) and np.array_equal(ps.w, ops.w) and np.array_equal(ps.key_c, ops.key_c)
    grid = make_grid(xi, yi)
    assert (grid.rows, grid.cols) == (16384, 16384)
    rng = np.random.default_rng(8)
    bands = (0, 5000, 8192, 12345, grid.rows - 32)                             # first rows ... last rows, 5 x 420 cells
    cells = {}
    for r0 in bands:
        z = ctx.solve_rows(ps, grid, r0, 32, exps)[0]
        z32 = ctx.solve_rows(ps, grid, r0, 32, exps, _lib.FP32)[0]
        rr, cc = r0 + rng.integers(0, 32, 420), rng.integers(0, grid.cols, 420)
        rr[:2], cc[:2] = (r0, r0 + 31), (0, grid.cols - 1)
        assert_parity_cells(z[rr - r0, cc], ps, xi, yi, 2.0, rr, cc, what=f"c4 rows {r0}")
        ref = O.solve_cells(ops, o_info(xi), o_info(yi), 2.0, rr, cc)
        err32 = assert_parity(z32[rr - r0, cc], ref, np.abs(ref), rel=REL32, what=f"c4 rows {r0} fp32")
        assert err32 > 1e-12
        inband = (ps.key_r >= r0) & (ps.key_r < r0 + 32)
        assert inband.sum() > 0 and np.array_equal(z[ps.key_r[inband] - r0, ps.key_c[inband]], ps.w[inband])
        assert np.array_equal(z32[ps.key_r[inband] - r0, ps.key_c[inband]], ps.w[inband])
        cells[r0] = (rr, cc)
    # second pass with SIGNED weights (the cancellation-robust tolerance at full size): w - 250 changes sign across the field
    sg = V.PointSet(ps.x, ps.y, ps.w - 250.0, ps.key_r, ps.key_c)
    ab = V.PointSet(ps.x, ps.y, np.abs(ps.w - 250.0), ps.key_r, ps.key_c)
    for r0 in (bands[0], bands[2], bands[4]):
        rr, cc = cells[r0]
        z = ctx.solve_rows(sg, grid, r0, 32, exps)[0]
        assert (z > 0).any() and (z < 0).any()
        assert_parity_cells(z[rr - r0, cc], sg, xi, yi, 2.0, rr, cc, scale_ps=ab, what=f"c4 signed rows {r0}")


def test_c5_epsg2284_sweep_band(ctx):
    """BASELINE.json config 5 (EPSG:2284 offsets 1.1e7 / 3e6, all 2 M points, 32768 columns, p = 1, 1.5, 2, 2.5):
    three 16-row full-width bands (first, middle, last rows), 210 cells x 4 exponents from one pass."""
    x, y, w, xi, yi, exps = workloads.make("c5")
    ps = V.MakeCoordSpace(x, y, w, xi, yi)
    grid = make_grid(xi, yi)
    assert (grid.rows, grid.cols) == (32768, 32768) and exps.tolist() == [1.0, 1.5, 2.0, 2.5] and len(ps) > 1_800_000
    rng = np.random.default_rng(9)
    for r0 in (0, 20000, grid.rows - 16):
        z = ctx.solve_rows(ps, grid, r0, 16, exps)
        rr, cc = r0 + rng.integers(0, 16, 70), rng.integers(0, grid.cols, 70)
        rr[:2], cc[:2] = (r0, r0 + 15), (0, grid.cols - 1)
        for e, p in enumerate(exps):
            assert_parity_cells(z[e][rr - r0, cc], ps, xi, yi, p, rr, cc, what=f"c5 rows {r0} p={p}")


def test_randomized_grids_points_and_sweeps(ctx):
    """40 random problems: grid shape, step, origin magnitude, point count, signed weights, sweep (es, ee, ei)."""
    rng = np.random.default_rng(2024)
    for it in range(40):
        rows, cols = int(rng.integers(1, 90)), int(rng.integers(1, 200))
        step = float(rng.choice([0.25, 1.0, 3.0, 100.0]))
        x0, y0 = float(rng.choice([0.0, -500.0, 1.2e7])), float(rng.choice([0.0, 3.9e6]))
        n = int(rng.integers(1, 1400))
        xi, yi = V.Info(x0, x0 + cols * step - 1, step), V.Info(y0, y0 + rows * step - 1, step)
        rows, cols = V.GetDimensions(xi, yi)
        if rows < 1 or cols < 1:
            continue
        ps = V.MakeCoordSpace(rng.uniform(x0 - step, x0 + (cols + 1) * step, n), rng.uniform(y0 - step, y0 + (rows + 1) * step, n),
                              rng.normal(0, 100, n), xi, yi)
        es = float(rng.choice([0.5, 1.0, 1.5, 2.0, 0.7, 3.0]))
        ei = float(rng.choice([0.5, 1.0, 0.3, 2.0]))
        exps = V.exp_range(es, es + ei * int(rng.integers(1, 5)) - 1e-9, ei)
        zs = V.solve(ps, xi, yi, exps, ctx=ctx)
        for e, p in enumerate(exps):
            ref, scale = oracle_ref(ps, xi, yi, float(p))
            assert_parity(zs[e], ref, scale, what=f"random #{it} {rows}x{cols} n={n} step={step} p={p}")


def test_stream_state_errors_and_reuse(ctx):
    """call-sequence errors of the streaming API (V2R_IDW_ESTATE) and that a context is reusable afterwards"""
    import ctypes as C
    L = _lib.load()
    ps, xi, yi = random_case(77, 300, 40, 50)
    grid = make_grid(xi, yi)
    h = ctx._h
    r0, nr = C.c_int64(), C.c_int64()
    buf = np.empty(2 * 16 * grid.cols)
    assert L.v2r_idw_stream_next(h, C.c_void_p(buf.ctypes.data), C.byref(r0), C.byref(nr)) == _lib.ESTATE     # no session
    assert L.v2r_idw_stream_end(h) == _lib.ESTATE
    x, y, w = (np.ascontiguousarray(a) for a in (ps.x, ps.y, ps.w))
    kr, kc = np.ascontiguousarray(ps.key_r), np.ascontiguousarray(ps.key_c)
    exps = np.array([2.0, 3.0])
    args = [h] + [C.c_void_p(a.ctypes.data) for a in (x, y, w, kr, kc)] + [len(x), C.byref(grid), C.c_void_p(exps.ctypes.data), 2, 0, 16]
    assert L.v2r_idw_stream_begin(*args) == _lib.OK
    assert L.v2r_idw_stream_begin(*args) == _lib.ESTATE                                                         # already open
    with pytest.raises(_lib.IdwError):
        ctx.solve(ps, grid, [2.0])                                                                              # solve during a session
    got = np.full((2, grid.rows, grid.cols), np.nan)
    while True:
        assert L.v2r_idw_stream_next(h, C.c_void_p(buf.ctypes.data), C.byref(r0), C.byref(nr)) == _lib.OK
        if nr.value == 0:
            break
        assert nr.value <= 16
        got[:, r0.value:r0.value + nr.value] = buf[:2 * nr.value * grid.cols].reshape(2, nr.value, grid.cols)
    assert L.v2r_idw_stream_next(h, C.c_void_p(buf.ctypes.data), C.byref(r0), C.byref(nr)) == _lib.OK and nr.value == 0  # stays finished
    assert L.v2r_idw_stream_end(h) == _lib.OK
    assert np.array_equal(got, ctx.solve(ps, grid, exps))                                                       # context still fine


def test_callable_from_many_threads(ctx):
    """cmd/idw.go:153-161 launches one goroutine per exponent and goroutines migrate between OS threads: the library
    must be callable from any thread (explicit device per call, per-context mutex).  Eight threads hammer ONE context
    and a second context concurrently; every raster must equal the one computed serially."""
    import threading
    ps, xi, yi = random_case(88, 800, 64, 96, step=2.0)
    grid = make_grid(xi, yi)
    exps = [0.5, 1.0, 1.5, 1.7, 2.0, 2.5, 3.0, 4.0]
    want = [ctx.solve(ps, grid, [p])[0] for p in exps]
    got, errs = [None] * len(exps), []
    other = V.IdwContext(1)

    def work(i):
        try:
            c = ctx if i % 2 == 0 else other
            for _ in range(3):
                got[i] = c.solve(ps, grid, [exps[i]])[0]
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=work, args=(i,)) for i in range(len(exps))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    other.close()
    assert not errs, errs
    for i in range(len(exps)):
        assert np.array_equal(got[i], want[i]), exps[i]


# --------------------------------------------------------------------------- boundary hardening (round 2)
def test_stream_begin_has_read_the_points_when_it_returns(ctx):
    """include/v2r_idw.h: the library never keeps a caller pointer after the call returns.  With PINNED point arrays
    the upload is a truly asynchronous DMA, so _stream_begin must wait for it: the arrays are overwritten right after
    the call and the rasters must still be those of the original points."""
    import ctypes as C
    L = _lib.load()
    rng = np.random.default_rng(91)
    xi, yi = V.Info(0, 95, 1.0), V.Info(0, 63, 1.0)
    n = 300_000                                                       # 2.4 MB per array: the DMA takes a while
    ps = _stale(rng.uniform(0, 96, n), rng.uniform(0, 64, n), rng.uniform(-5, 5, n))
    grid = make_grid(xi, yi)
    exps = np.array([2.0, 1.5])
    want = ctx.solve(ps, grid, exps)
    pins = [V.PinnedArray(len(ps), np.float64 if i < 3 else np.int64) for i in range(5)]
    try:
        for buf, src in zip(pins, (ps.x, ps.y, ps.w, ps.key_r, ps.key_c)):
            buf.array[:] = src
        args = [ctx._h] + [C.c_void_p(b.array.ctypes.data) for b in pins] + [len(ps), C.byref(grid), C.c_void_p(exps.ctypes.data), 2, 0, 0]
        assert L.v2r_idw_stream_begin(*args) == _lib.OK
        for b in pins[:3]:
            b.array[:] = np.nan                                        # a reader reusing its buffers for the next file
        pins[3].array[:] = 0
        pins[4].array[:] = 0
        got = np.empty_like(want)
        buf = np.empty(want.size)
        r0, nr = C.c_int64(), C.c_int64()
        while True:
            assert L.v2r_idw_stream_next(ctx._h, C.c_void_p(buf.ctypes.data), C.byref(r0), C.byref(nr)) == _lib.OK
            if nr.value == 0:
                break
            got[:, r0.value:r0.value + nr.value] = buf[:2 * nr.value * grid.cols].reshape(2, nr.value, grid.cols)
        assert L.v2r_idw_stream_end(ctx._h) == _lib.OK
    finally:
        for b in pins:
            b.free()
    assert np.array_equal(got, want)


def _stale(x, y, w):
    """points keyed far outside the grid: every cell goes through the sums (no exact hit)"""
    n = len(x)
    return V.PointSet(np.asarray(x, float), np.asarray(y, float), np.asarray(w, float), np.full(n, 10**9, np.int64), np.arange(n, dtype=np.int64))


@pytest.mark.parametrize("prec", ["fp64", "fp32"])
def test_paired_kernels_leave_no_hole_in_their_domain(ctx, prec):
    """The paired p = 2 kernels need the product of two squared distances to stay a normal number.  Point tiles for
    which that cannot be guaranteed on a CTA's cells run through the one-reciprocal loop: squared distances down to
    1e-160 and up to 1e+160 (FP64; FP32: beyond 2^-60 / 2^58 cells^2) give the reference's finite values, exactly like
    ordinary distances do."""
    xi, yi = V.Info(0, 199, 1.0), V.Info(0, 99, 1.0)
    rng = np.random.default_rng(17)
    n = 1500
    x, y, w = rng.uniform(0, 200, n), rng.uniform(0, 100, n), rng.uniform(1, 9, n)
    if prec == "fp64":
        tiny, far = 1e-80, 1e80
        precision, rel = _lib.FP64, REL
    else:
        tiny, far = 1e-12, 1e12
        precision, rel = _lib.FP32, REL32
    # (a) points a hair off a node in BOTH axes (d2 = 2 tiny^2): a*b underflows; (b) points absurdly far away
    #     (d2 ~ far^2): a*b overflows; (c) both in one 512-point tile together with ordinary points
    x[3], y[3] = 17.0 + 17.0 * 2**-52, 40.0 + 40.0 * 2**-52          # one ulp off node (40, 17): d2 ~ 1e-29
    x[700], y[700], w[700] = tiny, tiny, 123.0                        # a hair off node (0, 0), FP64: d2 = 2e-160
    x[701], y[701] = far, -far
    x[1400], y[1400] = -far, 3.0
    ps = _stale(x, y, w)
    z = V.solve(ps, xi, yi, [2.0], precision=precision, ctx=ctx)[0]
    ref = O.full_solve(o_ps(ps), o_info(xi), o_info(yi), 2.0)
    assert np.isfinite(ref).all() and abs(ref[0, 0] - 123.0) < 1e-12  # that weight dominates its node completely
    assert_parity(z, ref, ref, rel=rel, what=f"paired domain {prec}")
    assert abs(z[0, 0] - 123.0) < (1e-12 if prec == "fp64" else 1e-3)


def test_nan_coordinate_and_coincident_point_without_in_loop_specials(ctx):
    """The general class with positive exponents and the packed FP32 kernel have no special cases in their loops:
    a non-key cell that coincides with a point (reference: Pow(0, -p) = +Inf, Inf/Inf = NaN) and a NaN coordinate
    (reference: every weight NaN) are handled by the poison pass -- same rasters as the oracle, NaN for NaN."""
    xi, yi = V.Info(0, 63, 1.0), V.Info(0, 47, 1.0)
    rng = np.random.default_rng(23)
    n = 700
    x, y, w = rng.uniform(0, 64, n), rng.uniform(0, 48, n), rng.uniform(1, 9, n)
    x[10], y[10] = 20.0, 30.0                                         # exactly on node (30, 20), keyed elsewhere
    x[600], y[600] = 63.0, 0.0
    ps = _stale(x, y, w)
    for p, precision, rel in ((1.7, _lib.FP64, REL), (0.3, _lib.FP64, REL), (2.2, _lib.FP64, REL), (2.0, _lib.FP32, REL32)):
        z = V.solve(ps, xi, yi, [p], precision=precision, ctx=ctx)[0]
        ref = O.full_solve(o_ps(ps), o_info(xi), o_info(yi), p)
        assert np.isnan(ref[30, 20]) and np.isnan(ref[0, 63]) and np.isnan(ref).sum() == 2
        assert_parity(z, ref, np.abs(ref), rel=rel, what=f"coincident p={p}")
    # sweep: p > 0 planes poisoned, p = 0 plane is the plain mean there (Pow(d, -0) = 1), p < 0: h = 0
    zs = V.solve(ps, xi, yi, [1.7, 0.0, -1.0], ctx=ctx)
    for e, p in enumerate((1.7, 0.0, -1.0)):
        ref = O.full_solve(o_ps(ps), o_info(xi), o_info(yi), p)
        assert_parity(zs[e], ref, np.abs(ref), what=f"mixed sweep p={p}")
    x[77] = np.nan
    ps = _stale(x, y, w)
    keyed = V.PointSet(ps.x, ps.y, ps.w, ps.key_r.copy(), ps.key_c.copy())
    keyed.key_r[5], keyed.key_c[5] = 7, 9                              # one real exact hit survives the NaN flood
    for p, precision in ((1.7, _lib.FP64), (2.0, _lib.FP64), (2.0, _lib.FP32), (1.5, _lib.FP64)):
        z = V.solve(keyed, xi, yi, [p], precision=precision, ctx=ctx)[0]
        ref = O.full_solve(o_ps(keyed), o_info(xi), o_info(yi), p)
        assert np.isnan(ref).sum() == ref.size - 1 and ref[7, 9] == w[5]
        assert np.array_equal(np.isnan(z), np.isnan(ref)) and z[7, 9] == w[5], (p, precision)


def test_sweep_longer_than_64_exponents(ctx):
    """cmd/idw.go accepts any --es/--ee/--ei: a sweep of 70 exponents runs as 7 passes of 10 in one call"""
    ps, xi, yi = random_case(61, 300, 24, 40, step=2.0)
    exps = V.exp_range(0.1, 7.05, 0.1)
    assert len(exps) == 70
    zs = V.solve(ps, xi, yi, exps, ctx=ctx)
    for e in (0, 9, 10, 33, 69):
        ref, scale = oracle_ref(ps, xi, yi, float(exps[e]))
        assert_parity(zs[e], ref, scale, what=f"70-sweep e={e}")
