"""Point-source readers of the IDW path (tools/processing/read_data.go), returning SoA float64
arrays ready for MakeCoordSpace / the C ABI instead of []tools.Point.

    ReadTextData     read_data.go:17-111   grammar POINTS n / STEP / ESTIMATE
    ReadGeoPackage   read_data.go:149-173 + io_gdal.go:161-197 (OGR replaced by a sequential SQL scan
                     that decodes the GeoPackageBinary header itself; GDAL is not in this image)
"""
from __future__ import annotations

import sqlite3
import struct

import numpy as np

from .idw import Info


def ReadTextData(filepath: str, epsg: int = 2284):
    """Returns (x, y, w, proj, xInfo, yInfo).  proj is "EPSG:<code>" (the reference asks GDAL for the
    WKT of the code, read_data.go:25-34; without GDAL the code itself is carried)."""
    xInfo, yInfo = Info(), Info()
    xs, ys, ws = [], [], []
    with open(filepath) as f:
        lines = f.read().split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    it = iter(lines)
    for line in it:
        fields = line.split()
        if not fields:
            raise ValueError(f"{filepath}: blank line (the reference panics here, read_data.go:40)")
        head = fields[0]
        if head == "POINTS":
            n = int(line[len("POINTS "):].strip())
            for _ in range(n):
                f3 = next(it).split()
                xs.append(float(f3[0])); ys.append(float(f3[1])); ws.append(float(f3[2]))
        elif head == "STEP":
            f2 = next(it).split()
            xInfo.step, yInfo.step = float(f2[0]), float(f2[1])
        elif head == "ESTIMATE":
            a, b = next(it).split(), next(it).split()
            xInfo.min, xInfo.max = float(a[0]), float(a[1])
            yInfo.min, yInfo.max = float(b[0]), float(b[1])
            return np.array(xs), np.array(ys), np.array(ws), f"EPSG:{epsg}", xInfo, yInfo
    raise ValueError("ESTIMATE not in file")


def ReadGeoPackage(filename: str, layer: str, field: str, stepX: float, stepY: float):
    """Returns (x, y, w, proj_wkt, xInfo, yInfo); features in fid order like l.Feature(1..count)."""
    con = sqlite3.connect(f"file:{filename}?mode=ro", uri=True)
    try:
        layers = [r[0] for r in con.execute("select table_name from gpkg_contents where data_type='features'")]
        if layer not in layers:
            raise ValueError(f"layer {layer} not in file. Available layers: {layers}")  # io_gdal.go:134-159
        geom_col, srs_id = con.execute(
            "select column_name, srs_id from gpkg_geometry_columns where table_name=?", (layer,)).fetchone()
        proj = con.execute("select definition from gpkg_spatial_ref_sys where srs_id=?", (srs_id,)).fetchone()[0]
        cols = [r[1] for r in con.execute(f'pragma table_info("{layer}")')]
        if field not in cols:
            raise ValueError(f"field {field} not in layer {layer}")
        rows = con.execute(f'select "{geom_col}", "{field}" from "{layer}" order by rowid').fetchall()
    finally:
        con.close()
    n = len(rows)
    x, y, w = np.empty(n), np.empty(n), np.empty(n)
    for i, (blob, val) in enumerate(rows):
        if blob[:2] != b"GP":
            raise ValueError("not a GeoPackageBinary geometry")
        env = (blob[3] >> 1) & 7
        off = 8 + (0, 32, 48, 48, 64)[env]
        bo = "<" if blob[off] == 1 else ">"
        (gtype,) = struct.unpack_from(bo + "I", blob, off + 1)
        if gtype % 1000 != 1:
            raise ValueError("POINT geometries expected")
        x[i], y[i] = struct.unpack_from(bo + "dd", blob, off + 5)
        w[i] = float(val)
    xInfo = Info(float(x.min()) if n else float("inf"), float(x.max()) if n else float("-inf"), stepX)
    yInfo = Info(float(y.min()) if n else float("inf"), float(y.max()) if n else float("-inf"), stepY)
    return x, y, w, proj, xInfo, yInfo
