"""Point-source readers of the IDW path (tools/processing/read_data.go), returning SoA float64
arrays ready for MakeCoordSpace / the C ABI instead of []tools.Point.

    ReadTextData     read_data.go:17-111   grammar POINTS n / STEP / ESTIMATE
    ReadGeoPackage   read_data.go:149-173 + io_gdal.go:161-197 (OGR replaced by a sequential SQL scan
                     that decodes the GeoPackageBinary header itself; GDAL is not in this image)
"""
from __future__ import annotations

import math
import re
import sqlite3
import struct

import numpy as np

from .idw import Info


def _go_atoi_or_zero(s: str) -> int:
    """strconv.Atoi(strings.TrimSpace(s)) with the error dropped (read_data.go:95): sign + decimal digits or 0."""
    s = s.strip(" \t\n\r\v\f")
    d = s[1:] if s[:1] in ("+", "-") else s
    if not d or not d.isascii() or not d.isdigit():
        return 0
    return int(s)


_GO_FLOAT = re.compile(
    r"[+-]?(?:(?:inf|infinity)|(?:\d+\.?\d*|\.\d+)(?:e[+-]?\d+)?|0x(?:[0-9a-f]+\.?[0-9a-f]*|\.[0-9a-f]+)p[+-]?\d+)|nan",
    re.IGNORECASE | re.ASCII)


def _go_parse_float(tok: str):
    """strconv.ParseFloat(tok, 64) -> (value, ok).  Go's grammar: decimal floats, hex floats WITH a p exponent,
    [+-]inf / infinity and an unsigned nan in any case; no underscores without a base prefix, no "nan(...)".
    An out-of-range decimal gives +-Inf with an error that addPoints drops (read_data.go:104-106): +-Inf is kept."""
    if not _GO_FLOAT.fullmatch(tok):
        return 0.0, False
    if "x" in tok.lower():
        try:
            return float.fromhex(tok), True
        except OverflowError:
            return (-math.inf if tok[0] == "-" else math.inf), True
    return float(tok), True


def _go_minmax(a: np.ndarray):
    """math.Min / math.Max folded over a (read_data.go:158-162): NaN beats every finite value, -Inf (+Inf) beats NaN."""
    if a.size == 0:
        return float("inf"), float("-inf")
    if not np.isnan(a).any():
        return float(a.min()), float(a.max())
    lo = float("-inf") if np.isneginf(a).any() else float("nan")
    hi = float("inf") if np.isposinf(a).any() else float("nan")
    return lo, hi


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
            # numPoints, _ := strconv.Atoi(strings.TrimSpace(strings.TrimPrefix(line, "POINTS ")))  (read_data.go:95)
            n = _go_atoi_or_zero(line[len("POINTS "):] if line.startswith("POINTS ") else line)
            xs, ys, ws = [], [], []      # `data = addPoints(sc)` (read_data.go:42): a later section replaces an earlier one
            for _ in range(n):
                f3 = next(it, "").split()
                if len(f3) < 3:
                    raise ValueError(f"{filepath}: point line has fewer than 3 fields (the reference panics: index out of range)")
                # p.X, _ = strconv.ParseFloat(...): a syntax error leaves 0 (read_data.go:104-106)
                xs.append(_go_parse_float(f3[0])[0]); ys.append(_go_parse_float(f3[1])[0]); ws.append(_go_parse_float(f3[2])[0])
        elif head == "STEP":
            f2 = next(it, "").split()
            if len(f2) < 2:
                raise ValueError(f"{filepath}: STEP needs two numbers")
            xInfo.step, yInfo.step = _strict(f2[0], filepath), _strict(f2[1], filepath)
        elif head == "ESTIMATE":
            for xy in range(2):
                for k, tok in enumerate(next(it, "").split()):      # every field is parsed, the first two are used (:57-76)
                    v = _strict(tok, filepath)
                    if (xy, k) == (0, 0):
                        xInfo.min = v
                    elif (xy, k) == (0, 1):
                        xInfo.max = v
                    elif (xy, k) == (1, 0):
                        yInfo.min = v
                    elif (xy, k) == (1, 1):
                        yInfo.max = v
            return np.array(xs, dtype=np.float64), np.array(ys, dtype=np.float64), np.array(ws, dtype=np.float64), f"EPSG:{epsg}", xInfo, yInfo
    raise ValueError("ESTIMATE not in file")


def _strict(tok: str, filepath: str) -> float:
    v, ok = _go_parse_float(tok)
    if not ok:
        raise ValueError(f'{filepath}: strconv.ParseFloat: parsing "{tok}": invalid syntax')
    return v


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
    (x_lo, x_hi), (y_lo, y_hi) = _go_minmax(x), _go_minmax(y)
    xInfo, yInfo = Info(x_lo, x_hi, stepX), Info(y_lo, y_hi, stepY)
    return x, y, w, proj, xInfo, yInfo
