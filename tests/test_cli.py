"""The C++ host (`v2r_b200/bin/v2r idw ...`): same flags as the reference CLI (cmd/idw.go:50-67), readers that
fill SoA buffers, native BigTIFF sink, TransferType.  CPU tests use --plan (no GPU needed); the GPU tests replay
features/idw/idw_test.go through the CLI and compare a gpkg sweep with the oracle."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

from v2r_b200 import raster_io, workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
V2R = os.path.join(ROOT, "v2r_b200", "bin", "v2r")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def run(*args, check=True):
    p = subprocess.run([V2R, *args], capture_output=True, text=True)
    if check:
        assert p.returncode == 0, p.stderr
    return p


def test_plan_txt_defaults_match_reference_cli():
    p = json.loads(run("idw", "-f", os.path.join(GOLDEN, "idw_in.txt"), "--plan", "-e").stdout)
    # defaults: --es 1.5 --ei .5, ee = es+ei -> one exponent 1.5; --sx/--sy only name the file in txt mode
    assert (p["rows"], p["cols"], p["points"], p["exponents"]) == (16, 13, 5, 1)
    assert p["first_outfile"] == "data/idw/idw_in_step100-100pow1.5.tiff" and "root4" in p["kernel"]
    p = json.loads(run("idw", "--file", os.path.join(GOLDEN, "idw_in.txt"), "--plan", "-e", "-c", "--es=0.5", "--ee", "5",
                       "--outPath", "out/").stdout)
    assert p["exponents"] == 9 and p["first_outfile"] == "out/idw_in_step100-100chunkedpow0.5.tiff"


def test_plan_gpkg_reader_matches_fixture():
    p = json.loads(run("idw", "-g", "-f", os.path.join(GOLDEN, "input.gpkg"), "--layer", "wsels", "--field", "elevation",
                       "--sx", "1", "--sy", "2", "--plan", "-e").stdout)
    xi, yi = O.Info(1.2020401943588393e+07, 1.2096825067172093e+07, 1.0), O.Info(3.885548111378168e+06, 3.953694544257536e+06, 2.0)
    assert (p["rows"], p["cols"]) == O.get_dimensions(xi, yi) and p["points"] == 14


def test_cli_errors_like_the_reference():
    f = os.path.join(GOLDEN, "idw_in.txt")
    assert "has no iterations" in run("idw", "-f", f, "--es", "2", "--ee", "1", check=False).stderr     # cmd/idw.go:86-88
    assert "must be > 0" in run("idw", "-f", f, "--ei", "0", check=False).stderr                        # cmd/idw.go:83-85
    assert "\"file\" not set" in run("idw", check=False).stderr
    assert "\"layer\", \"field\" not set" in run("idw", "-g", "-f", f, check=False).stderr              # cmd/idw.go:71-76
    p = run("idw", "-g", "-f", os.path.join(GOLDEN, "input.gpkg"), "--layer", "nope", "--field", "elevation", "--plan", check=False)
    assert p.returncode == 1 and "not in file" in p.stderr                                              # io_gdal.go:134-159


def test_translate_matches_python_sink(tmp_path):
    z = np.random.default_rng(0).normal(0, 50, (16, 13))
    z[2, 3] = 45.5
    gd = raster_io.CreateGDalInfo(-1.0, 0.0, 1.0, 1.0, 7, "EPSG:2284")
    raster_io.WriteGDAL(z, gd, str(tmp_path / "a.tiff"), (0, 0), (16, 13), (16, 13), True)
    run("translate", str(tmp_path / "a.tiff"), str(tmp_path / "a.asc"), "Int16")
    assert open(tmp_path / "a.asc", "rb").read() == O.aaigrid_int16_bytes(z, -1.0, 0.0, 1.0)


def test_translate_rejects_damaged_headers_instead_of_reading_past_them(tmp_path):
    """the BigTIFF reader of the C++ sink only trusts offsets that lie inside the header block it has read"""
    import struct
    z = np.arange(12, dtype=np.float64).reshape(3, 4)
    gd = raster_io.CreateGDalInfo(0.0, 0.0, 1.0, 1.0, 7, "EPSG:2284")
    good = tmp_path / "good.tiff"
    raster_io.WriteGDAL(z, gd, str(good), (0, 0), (3, 4), (3, 4), True)
    blob = bytearray(good.read_bytes())
    for name, patch in (("ifd_far", lambda b: b.__setitem__(slice(8, 16), struct.pack("<Q", 1 << 40))),
                        ("ifd_count", lambda b: b.__setitem__(slice(16, 24), struct.pack("<Q", 1 << 50))),
                        ("short", lambda b: b.__delitem__(slice(10, len(b))))):
        bad = bytearray(blob)
        patch(bad)
        f = tmp_path / f"{name}.tiff"
        f.write_bytes(bytes(bad))
        p = run("translate", str(f), str(tmp_path / "o.asc"), "Int16", check=False)
        assert p.returncode == 1 and "[FATAL]" in p.stderr, (name, p.returncode, p.stderr)
    # the ModelTransformationTag pointing outside the header block
    (n,) = struct.unpack_from("<Q", blob, 16)
    for i in range(n):
        if struct.unpack_from("<H", blob, 24 + 20 * i)[0] == 34264:
            struct.pack_into("<Q", blob, 24 + 20 * i + 12, 1 << 33)
    (tmp_path / "xf.tiff").write_bytes(bytes(blob))
    p = run("translate", str(tmp_path / "xf.tiff"), str(tmp_path / "o.asc"), "Int16", check=False)
    assert p.returncode == 1 and "ModelTransformationTag" in p.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("conc", [False, True])
def test_golden_through_cli(tmp_path, conc):
    """features/idw/idw_test.go: pow 1.7, serial and chunked (3x2), GTiff -> Int16 AAIGrid -> golden bytes"""
    out = str(tmp_path) + "/"
    args = ["idw", "-f", os.path.join(GOLDEN, "idw_in.txt"), "--es", "1.7", "--ei", "1", "--outPath", out, "-e"]
    if conc:
        args += ["-c", "--cy", "3", "--cx", "2"]
    run(*args)
    tiff = out + f"idw_in_step100-100{'chunked' if conc else ''}pow1.7.tiff"
    run("translate", tiff, tiff[:-5] + ".asc", "Int16")
    assert open(tiff[:-5] + ".asc", "rb").read() == open(os.path.join(GOLDEN, "idw_correct_step1-1.asc"), "rb").read()
    z, x0, y0, dx, dy = raster_io.ReadTiff(tiff)            # the Python sink reads what the C++ sink wrote
    assert z.shape == (16, 13) and (x0, y0, dx, dy) == (-1.0, 0.0, 1.0, 1.0)


@pytest.mark.gpu
def test_gpkg_sweep_through_cli_vs_oracle(tmp_path):
    x, y, w, xi, yi, _ = workloads.make("c2", n_points=3000, cells=96)
    gp = str(tmp_path / "pts.gpkg")
    workloads.write_gpkg(gp, x, y, w)
    out = str(tmp_path) + "/"
    run("idw", "-g", "-f", gp, "--layer", "wsels", "--field", "elevation", "--sx", "100", "--sy", "100", "--es", "1", "--ee", "3",
        "--ei", "0.5", "--outPath", out, "-e", "-c", "--cy", "40", "--cx", "40")
    oxi, oyi = O.Info(float(x.min()), float(x.max()), 100.0), O.Info(float(y.min()), float(y.max()), 100.0)
    ops = O.make_coord_space(x, y, w, oxi, oyi)
    for p in (1.0, 1.5, 2.0, 2.5):
        z = raster_io.ReadTiff(out + f"pts_step100-100chunkedpow{p:.1f}.tiff")[0]
        ref = O.chunk_solve(ops, oxi, oyi, 8, 32, p)
        assert z.shape == ref.shape == (96, 96)
        assert np.max(np.abs(z - ref) / np.abs(ref)) < 1e-12          # weights > 0: scale == |z|


@pytest.mark.gpu
def test_txt_100k_points_through_cli(tmp_path):
    x, y, w, xi, yi, _ = workloads.make("c2", n_points=100_000, cells=256)
    f = str(tmp_path / "big.txt")
    workloads.write_txt(f, x, y, w, xi, yi)
    out = str(tmp_path) + "/"
    run("idw", "-f", f, "--es", "2", "--ei", "1", "--outPath", out, "-e", "--fp32")
    z = raster_io.ReadTiff(out + "big_step100-100pow2.0.tiff")[0]
    ops = O.make_coord_space(x, y, w, O.Info(xi.min, xi.max, 100.0), O.Info(yi.min, yi.max, 100.0))
    rr, cc = np.arange(0, 256, 17), np.arange(3, 256, 19)[:16]
    ref = O.solve_cells(ops, O.Info(xi.min, xi.max, 100.0), O.Info(yi.min, yi.max, 100.0), 2.0, rr[:len(cc)], cc)
    assert np.max(np.abs(z[rr[:len(cc)], cc] - ref) / np.abs(ref)) < 2e-5


def test_txt_reader_error_cases(tmp_path):
    """read_data.go:17-111 failure modes: missing file, no ESTIMATE section, blank line, short point line, bad number"""
    def plan(text):
        f = tmp_path / "t.txt"
        f.write_text(text)
        return run("idw", "-f", str(f), "--plan", "-e", check=False)
    assert "no such file" in run("idw", "-f", str(tmp_path / "missing.txt"), "--plan", check=False).stderr
    assert "ESTIMATE not in file" in plan("POINTS 1\n1 2 3\n").stderr                       # read_data.go:89
    assert "blank line" in plan("POINTS 1\n1 2 3\n\nESTIMATE\n0 5\n0 5").stderr              # the reference panics (:40)
    assert "fewer than 3 fields" in plan("POINTS 2\n1 2 3\n4 5\nESTIMATE\n0 5\n0 5").stderr
    assert "invalid syntax" in plan("POINTS 1\n1 2 3\nSTEP\nabc 1\nESTIMATE\n0 5\n0 5").stderr  # ParseFloat error is returned (:47-50)
    ok = plan("POINTS 3\n1 2 3\r\n+4.5 5e0 x\n2 2 9\nSTEP\n0.5 0.25\nESTIMATE\n0 5\n0 5")   # CRLF, '+', exponent; 'x' -> 0 like the dropped error
    p = json.loads(ok.stdout)
    assert ok.returncode == 0 and (p["rows"], p["cols"], p["points_in"]) == (24, 12, 3)      # int((1+5)/0.25), int((1+5)/0.5)


def test_txt_readers_keep_the_dropped_error_semantics_of_addPoints(tmp_path):
    """addPoints (read_data.go:93-111) drops the errors of Atoi and ParseFloat: a count Go cannot parse reads 0 points, a
    number Go cannot parse is 0, an out-of-range number is +-Inf; and `data = addPoints(sc)` (:42) makes a later POINTS
    section REPLACE an earlier one.  The C++ reader, the Python reader and the oracle's reader must agree."""
    from v2r_b200 import read_data
    text = ("POINTS 2\n9 9 9\n8 8 8\n"                          # replaced by the next section
            "POINTS 7\r\n"
            "1 2 3\n"
            "0x1p-2 .5 5.\n"                                      # hex float with exponent, bare fractions
            "1_0 0x10 +nan\n"                                     # Go: syntax errors -> 0 0 0
            "1e999 -1e999 2\n"                                    # ErrRange keeps +-Inf
            "inf -Infinity 4\n"
            "nan(1) 1e 7\n"                                       # syntax errors -> 0 0
            "+4.25E+1 -0.0 1e-320\n"                              # denormal
            "POINTS x\n"                                          # Atoi fails -> 0 points: the point set is now EMPTY ...
            "POINTS 3\n5 6 7\n6 7 8 trailing fields are ignored\n7 8 9\n"   # ... and this section is what remains
            "STEP\n2 4 ignored\nESTIMATE\n0 10 99\n0 20\n")
    f = tmp_path / "odd.txt"
    f.write_text(text)
    p = json.loads(run("idw", "-f", str(f), "--plan", "-e", "--dump-points", "10").stdout)
    assert p["raw"] == [[5.0, 6.0, 7.0], [6.0, 7.0, 8.0], [7.0, 8.0, 9.0]] and p["points_in"] == 3
    assert (p["rows"], p["cols"]) == (5, 5)                      # int((1+20)/4), int((1+10)/2)
    x, y, w, _, xi, yi = read_data.ReadTextData(str(f))
    assert [x.tolist(), y.tolist(), w.tolist()] == [[5.0, 6.0, 7.0], [6.0, 7.0, 8.0], [7.0, 8.0, 9.0]]
    assert (xi.min, xi.max, xi.step, yi.min, yi.max, yi.step) == (0.0, 10.0, 2.0, 0.0, 20.0, 4.0)
    ox, oy, ow, oxi, oyi = O.read_text_data(str(f))
    assert np.array_equal(ox, x) and np.array_equal(ow, w) and (oxi.max, oyi.step) == (10.0, 4.0)
    # the middle section on its own: every spelling, value by value (NaN-aware), C++ vs Python vs the oracle
    f2 = tmp_path / "odd2.txt"
    f2.write_text(text.split("POINTS x\n")[0].split("8 8 8\n")[1] + "ESTIMATE\n0 10\n0 20\n")
    want = [[1.0, 2.0, 3.0], [0.25, 0.5, 5.0], [0.0, 0.0, 0.0], [np.inf, -np.inf, 2.0], [np.inf, -np.inf, 4.0], [0.0, 0.0, 7.0],
            [42.5, -0.0, 1e-320]]
    raw = run("idw", "-f", str(f2), "--plan", "-e", "--dump-points", "10").stdout
    raw = re.sub(r"(?<![\d.e])-0(?=[,\]])", "-0.0", raw)       # printf prints -0.0 as "-0", which json reads as the integer 0
    got = json.loads(raw.replace("-inf", "-Infinity").replace(" inf", " Infinity").replace("[inf", "[Infinity"))["raw"]
    x, y, w = read_data.ReadTextData(str(f2))[:3]
    ox, oy, ow = O.read_text_data(str(f2))[:3]
    for k, (a, b, c) in enumerate(((x, y, w), (ox, oy, ow), tuple(np.array([r[i] for r in got]) for i in range(3)))):
        cols = np.array([a, b, c]).T
        assert np.array_equal(cols, np.array(want)), k
        assert np.array_equal(np.signbit(cols), np.signbit(np.array(want))), k


def test_gpkg_bbox_follows_go_min_max_with_nan_coordinates(tmp_path):
    """ReadGeoPackage folds math.Min / math.Max over the points (read_data.go:158-162): a NaN coordinate makes the bound NaN
    (std::fmin / np.nanmin would skip it); the grid then has no valid size (int(NaN)) and the run stops before any solve."""
    from v2r_b200 import read_data
    x, y, w = np.array([1.0, np.nan, 5.0]), np.array([2.0, 3.0, 4.0]), np.array([1.0, 2.0, 3.0])
    f = str(tmp_path / "nan.gpkg")
    workloads.write_gpkg(f, x, y, w)
    _, _, _, _, xi, yi = read_data.ReadGeoPackage(f, "wsels", "elevation", 1.0, 1.0)
    assert np.isnan(xi.min) and np.isnan(xi.max) and (yi.min, yi.max) == (2.0, 4.0)
    p = run("idw", "-g", "-f", f, "--layer", "wsels", "--field", "elevation", "--sx", "1", "--sy", "1", "--plan", "-e", check=False)
    assert '"cols": -9223372036854775808' in p.stdout and '"rows": 3' in p.stdout       # Go: int(NaN) on amd64


def test_cpp_readers_parse_bit_exactly_and_dedupe_like_the_oracle(tmp_path):
    """the in-place tokenizer (std::from_chars) must give the doubles Python's float() gives for every spelling, and
    MakeCoordSpace through the C++ host must equal the oracle's input-order restatement (tools/utils.go:97-119)"""
    rng = np.random.default_rng(5)
    n = 400
    x, y = rng.uniform(-50, 1050, n), rng.uniform(-20, 520, n)
    w = rng.normal(0, 1e3, n)
    fmts = [repr, lambda v: f"{v:.3f}", lambda v: f"{v:.10e}", lambda v: f"{v:+.6f}", lambda v: f"{v:.17g}", lambda v: str(int(v))]
    lines, want = [], []
    for i in range(n):
        f = fmts[i % len(fmts)]
        toks = [f(float(x[i])), f(float(y[i])), f(float(w[i]))]
        lines.append("  ".join(toks) if i % 7 else "\t".join(toks) + " ")
        want.append([float(t) for t in toks])
    path = tmp_path / "pts.txt"
    path.write_text(f"POINTS {n}\n" + "\n".join(lines) + "\nSTEP\n10 5\nESTIMATE\n0 1000\n0 500\n")
    p = json.loads(run("idw", "-f", str(path), "--plan", "-e", "--dump-points", str(n)).stdout)
    assert p["raw"] == want
    wx, wy, ww = (np.array([r[k] for r in want]) for k in range(3))
    ops = O.make_coord_space(wx, wy, ww, O.Info(0, 1000, 10.0), O.Info(0, 500, 5.0))
    assert p["points"] == len(ops) < n
    keyed = np.array(p["keyed"])
    assert np.array_equal(keyed[:, 0], ops.x) and np.array_equal(keyed[:, 2], ops.w)
    assert np.array_equal(keyed[:, 3].astype(np.int64), ops.key_r) and np.array_equal(keyed[:, 4].astype(np.int64), ops.key_c)
    # the gpkg reader against the 14 exact points of tools/processing/gpkg_test.go:12-27
    exp = json.load(open(os.path.join(GOLDEN, "gpkg_points.json")))
    g = json.loads(run("idw", "-g", "-f", os.path.join(GOLDEN, "input.gpkg"), "--layer", "wsels", "--field", "elevation", "--sx", "1",
                       "--sy", "2", "--plan", "-e", "--dump-points", "14").stdout)
    assert g["raw"] == exp["points"]


def test_gpkg_with_a_primary_key_that_is_not_called_fid(tmp_path):
    """The reference reads through OGR (l.Feature(1..count), io_gdal.go:180-190), which works whatever the FID column
    is called; the SQL readers order by rowid (an INTEGER PRIMARY KEY is an alias of it), not by a column named fid."""
    from v2r_b200 import read_data
    rng = np.random.default_rng(12)
    x, y, w = rng.uniform(0, 900, 50), rng.uniform(0, 400, 50), rng.normal(0, 10, 50)
    f = str(tmp_path / "objectid.gpkg")
    workloads.write_gpkg(f, x, y, w, pk="OBJECTID")
    x2, y2, w2, _, _, _ = read_data.ReadGeoPackage(f, "wsels", "elevation", 10.0, 10.0)
    assert np.array_equal(x, x2) and np.array_equal(y, y2) and np.array_equal(w, w2)
    x3, y3, w3 = O.read_geopackage(f, "wsels", "elevation", 10.0, 10.0)[:3]
    assert np.array_equal(x, x3) and np.array_equal(w, w3)
    g = json.loads(run("idw", "-g", "-f", f, "--layer", "wsels", "--field", "elevation", "--sx", "10", "--sy", "10", "--plan", "-e",
                       "--dump-points", "50").stdout)
    assert g["raw"] == [[float(a), float(b), float(c)] for a, b, c in zip(x, y, w)]
