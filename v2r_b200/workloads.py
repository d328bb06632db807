"""Synthetic point clouds of the shapes BASELINE.json names, numpy only.

    c2  100 k uniform points        -> 4096 x 4096,   p = 2
    c3  250 k uniform points        -> 8192 x 8192,   p = 0.5 .. 4.5 step 0.5 (9 rasters, one pass)
    c4  1 M clustered points        -> 16384 x 16384, p = 2
    c5  2 M gpkg-shaped (EPSG:2284) -> 32768 x 32768, p = 1, 1.5, 2, 2.5

Each returns (x, y, w, xInfo, yInfo, exps); the caller runs MakeCoordSpace on it.  `scale_rows`
multiplies the y extent (weak scaling over row bands: every rank keeps a c-sized band).
"""
from __future__ import annotations

import numpy as np

from .idw import Info

SPECS = {
    "c2": dict(n=100_000, cells=4096, seed=1002, kind="uniform", es=2.0, ee=2.5, ei=0.5, x0=0.0, y0=0.0),
    "c3": dict(n=250_000, cells=8192, seed=1003, kind="uniform", es=0.5, ee=5.0, ei=0.5, x0=0.0, y0=0.0),
    "c4": dict(n=1_000_000, cells=16384, seed=1004, kind="clustered", es=2.0, ee=2.5, ei=0.5, x0=0.0, y0=0.0),
    "c5": dict(n=2_000_000, cells=32768, seed=1005, kind="clustered", es=1.0, ee=3.0, ei=0.5, x0=1.10e7, y0=3.0e6),
}
STEP = 100.0


def _clustered(rng, n, ext_x, ext_y):
    nb = n // 10
    nc = n - nb
    k = 64
    cx, cy = rng.uniform(0, ext_x, k), rng.uniform(0, ext_y, k)
    which = rng.integers(0, k, nc)
    sx, sy = 0.02 * ext_x, 0.02 * ext_y
    x = np.clip(cx[which] + rng.normal(0, sx, nc), 0, ext_x)
    y = np.clip(cy[which] + rng.normal(0, sy, nc), 0, ext_y)
    x = np.concatenate([x, rng.uniform(0, ext_x, nb)])
    y = np.concatenate([y, rng.uniform(0, ext_y, nb)])
    perm = rng.permutation(n)
    return x[perm], y[perm]


def make(name: str, scale_rows: int = 1, n_points: int | None = None, cells: int | None = None):
    s = SPECS[name]
    n = n_points or s["n"]
    cells_x = cells or s["cells"]
    cells_y = cells_x * scale_rows
    rng = np.random.default_rng(s["seed"])
    ext_x = cells_x * STEP - 1.0          # int((1 + ext)/100) == cells  (tools.GetDimensions)
    ext_y = cells_y * STEP - 1.0
    if s["kind"] == "uniform":
        x = rng.uniform(0, ext_x, n)
        y = rng.uniform(0, ext_y, n)
        w = rng.uniform(0, 500, n)
    else:
        x, y = _clustered(rng, n, ext_x, ext_y)
        w = 250 + 200 * np.sin(x / 2e5) * np.cos(y / 2e5) + rng.normal(0, 5, n)
    # pin the bounding box with two corner points so the grid has exactly the named size
    x[0], y[0], x[1], y[1] = 0.0, 0.0, ext_x, ext_y
    x += s["x0"]
    y += s["y0"]
    xInfo = Info(s["x0"], s["x0"] + ext_x, STEP)
    yInfo = Info(s["y0"], s["y0"] + ext_y, STEP)
    exps = []
    e = s["es"]
    while e < s["ee"]:
        exps.append(e)
        e += s["ei"]
    return x, y, w, xInfo, yInfo, np.array(exps)


def write_txt(path: str, x, y, w, xInfo: Info, yInfo: Info):
    """The reference's txt grammar (tools/processing/read_data.go:17-111) incl. a STEP section."""
    with open(path, "w") as f:
        f.write(f"POINTS {len(x)}\n")
        for a, b, c in zip(x, y, w):
            f.write(f"{float(a)!r} {float(b)!r} {float(c)!r}\n")
        f.write(f"STEP\n{float(xInfo.step)!r} {float(yInfo.step)!r}\nESTIMATE\n{float(xInfo.min)!r} {float(xInfo.max)!r}\n{float(yInfo.min)!r} {float(yInfo.max)!r}")


def write_gpkg(path: str, x, y, w, layer="wsels", field="elevation", srs_id=2284, wkt="EPSG:2284", pk="fid"):
    """A GeoPackage with the fixture's schema (features/idw/idw_test_files/input.gpkg); `pk` names the integer primary
    key column (OGR writes "fid", ArcGIS "OBJECTID", `ogr2ogr -lco FID=id` anything)."""
    import os
    import sqlite3
    import struct
    if os.path.exists(path):
        os.remove(path)
    con = sqlite3.connect(path)
    con.executescript(f"""
        PRAGMA application_id = 1196444487; PRAGMA user_version = 10200;
        CREATE TABLE gpkg_spatial_ref_sys (srs_name TEXT NOT NULL, srs_id INTEGER NOT NULL PRIMARY KEY,
            organization TEXT NOT NULL, organization_coordsys_id INTEGER NOT NULL, definition TEXT NOT NULL, description TEXT);
        CREATE TABLE gpkg_contents (table_name TEXT NOT NULL PRIMARY KEY, data_type TEXT NOT NULL, identifier TEXT UNIQUE,
            description TEXT DEFAULT '', last_change DATETIME NOT NULL DEFAULT (strftime('%Y-%m-%dT%H:%M:%fZ','now')),
            min_x DOUBLE, min_y DOUBLE, max_x DOUBLE, max_y DOUBLE, srs_id INTEGER);
        CREATE TABLE gpkg_geometry_columns (table_name TEXT NOT NULL, column_name TEXT NOT NULL,
            geometry_type_name TEXT NOT NULL, srs_id INTEGER NOT NULL, z TINYINT NOT NULL, m TINYINT NOT NULL);
        CREATE TABLE "{layer}" ("{pk}" INTEGER PRIMARY KEY AUTOINCREMENT NOT NULL, "geom" POINT, "uid" MEDIUMINT, "{field}" REAL);
    """)
    con.execute("insert into gpkg_spatial_ref_sys values (?,?,?,?,?,?)", ("srs", srs_id, "EPSG", srs_id, wkt, None))
    con.execute("insert into gpkg_contents (table_name, data_type, identifier, min_x, min_y, max_x, max_y, srs_id) "
                "values (?,?,?,?,?,?,?,?)", (layer, "features", layer, float(np.min(x)), float(np.min(y)),
                                             float(np.max(x)), float(np.max(y)), srs_id))
    con.execute("insert into gpkg_geometry_columns values (?,?,?,?,?,?)", (layer, "geom", "POINT", srs_id, 0, 0))
    head = b"GP\x00\x01" + struct.pack("<i", srs_id) + b"\x01" + struct.pack("<I", 1)
    con.executemany(f'insert into "{layer}" (geom, uid, "{field}") values (?,?,?)',
                    ((head + struct.pack("<dd", float(a), float(b)), i + 1, float(c)) for i, (a, b, c) in enumerate(zip(x, y, w))))
    con.commit()
    con.close()
