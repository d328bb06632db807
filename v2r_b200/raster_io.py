"""Raster sink of the IDW path without GDAL (absent from this image).

Mirrors tools/processing/io_gdal.go:
    GDalInfo / CreateGDalInfo   io_gdal.go:13-24
    WriteGDAL                   io_gdal.go:27-74   create-or-update + window write at (offsets.C, offsets.R)
    TransferType                io_gdal.go:115-132 (`gdal_translate -ot Int16` to AAIGrid, used by idw_test.go:57)

The GTiff written here is what GDAL's driver.Create(..., Float64, BIGTIFF=YES) produces in substance:
a little-endian BigTIFF, one Float64 band, uncompressed, with the south-up geotransform
{xMin, xCell, 0, yMin, 0, +yCell} stored as a ModelTransformationTag (GDAL uses that tag whenever
the transform is not north-up) and the CRS as GeoKeys (EPSG code when the WKT carries one).
The pixel block is contiguous row-major, so a window write is one seek per row.
"""
from __future__ import annotations

import os
import re
import struct
from dataclasses import dataclass

import numpy as np

GDT_Int16, GDT_Float64 = 3, 7


@dataclass
class GDalInfo:
    XMin: float
    YMin: float
    XCell: float
    YCell: float
    GDalDataType: int
    Proj: str


def CreateGDalInfo(XMin, YMin, XCell, YCell, GDalDataType, proj) -> GDalInfo:
    return GDalInfo(XMin, YMin, XCell, YCell, GDalDataType, proj)


_DATA_OFFSET = 4096  # pixel block starts here; header, IFD and tag payloads live below it


def _epsg_of(wkt: str):
    m = re.search(r'AUTHORITY\["EPSG","(\d+)"\]\]\s*$', wkt or "")
    return int(m.group(1)) if m else None


def _tiff_create(path: str, info: GDalInfo, rows: int, cols: int):
    tags = []  # (tag, type, count, payload bytes)

    def short(v): return struct.pack("<H", v)
    def long8(v): return struct.pack("<Q", v)

    nbytes = rows * cols * 8
    tags.append((256, 16, 1, long8(cols)))           # ImageWidth (LONG8)
    tags.append((257, 16, 1, long8(rows)))           # ImageLength
    tags.append((258, 3, 1, short(64)))              # BitsPerSample
    tags.append((259, 3, 1, short(1)))               # Compression = none
    tags.append((262, 3, 1, short(1)))               # Photometric = BlackIsZero
    tags.append((273, 16, 1, long8(_DATA_OFFSET)))   # StripOffsets
    tags.append((277, 3, 1, short(1)))               # SamplesPerPixel
    tags.append((278, 16, 1, long8(max(rows, 1))))   # RowsPerStrip = whole image
    tags.append((279, 16, 1, long8(nbytes)))         # StripByteCounts
    tags.append((284, 3, 1, short(1)))               # PlanarConfiguration
    tags.append((339, 3, 1, short(3)))               # SampleFormat = IEEE float
    xf = [info.XCell, 0.0, 0.0, info.XMin, 0.0, info.YCell, 0.0, info.YMin, 0, 0, 0, 0, 0, 0, 0, 1.0]
    tags.append((34264, 12, 16, struct.pack("<16d", *xf)))  # ModelTransformationTag (south-up)
    epsg = _epsg_of(info.Proj)
    keys = [(1024, 0, 1, 1), (1025, 0, 1, 1)]        # GTModelType = projected, GTRasterType = PixelIsArea
    ascii_params = b""
    if epsg is not None:
        keys.append((3072, 0, 1, epsg))              # ProjectedCSTypeGeoKey
    elif info.Proj:
        ascii_params = info.Proj.encode() + b"|\x00"
        keys.append((3072, 0, 1, 32767))             # user-defined; WKT carried as citation
        keys.append((3073, 34737, len(ascii_params) - 1, 0))
    gk = [1, 1, 0, len(keys)] + [v for k in keys for v in k]
    tags.append((34735, 3, len(gk), struct.pack(f"<{len(gk)}H", *gk)))
    if ascii_params:
        tags.append((34737, 2, len(ascii_params), ascii_params))
    tags.sort(key=lambda t: t[0])

    ifd_off = 16
    ifd_size = 8 + 20 * len(tags) + 8
    extra_off = ifd_off + ifd_size
    entries, extra = b"", b""
    for tag, typ, cnt, payload in tags:
        if len(payload) <= 8:
            val = payload.ljust(8, b"\x00")
        else:
            val = struct.pack("<Q", extra_off + len(extra))
            extra += payload + (b"\x00" * (-len(payload) % 8))
        entries += struct.pack("<HHQ", tag, typ, cnt) + val
    head = struct.pack("<2sHHHQ", b"II", 43, 8, 0, ifd_off) + struct.pack("<Q", len(tags)) + entries + struct.pack("<Q", 0) + extra
    if len(head) > _DATA_OFFSET:
        raise ValueError("projection text too long for the reserved TIFF header block")
    with open(path, "wb") as f:
        f.write(head.ljust(_DATA_OFFSET, b"\x00"))
        f.truncate(_DATA_OFFSET + nbytes)


def _tiff_layout(path: str):
    """(rows, cols, data_offset, transform16 or None) of a file written by _tiff_create."""
    with open(path, "rb") as f:
        head = f.read(_DATA_OFFSET)
    bo, ver, _, _, ifd = struct.unpack("<2sHHHQ", head[:16])
    if bo != b"II" or ver != 43:
        raise ValueError(f"{path}: not a little-endian BigTIFF")
    (n,) = struct.unpack("<Q", head[ifd:ifd + 8])
    vals = {}
    for i in range(n):
        tag, typ, cnt, raw = struct.unpack("<HHQ8s", head[ifd + 8 + 20 * i: ifd + 28 + 20 * i])
        vals[tag] = (typ, cnt, raw)
    q = lambda t: struct.unpack("<Q", vals[t][2])[0]
    xf = None
    if 34264 in vals:
        off = q(34264)
        xf = struct.unpack("<16d", head[off:off + 128])
    return q(257), q(256), q(273), xf


def _asc_header(cols, rows, x_min, y_ll, cell):
    return (f"ncols        {cols}\nnrows        {rows}\nxllcorner    {x_min:.12f}\n"
            f"yllcorner    {y_ll:.12f}\ncellsize     {cell:.12f}\n")


def WriteGDAL(unwrappedMatrix, GDINFO: GDalInfo, filename: str, offsets, totalSize, bufferSize, create: bool):
    """io_gdal.go:27-74.  offsets/totalSize/bufferSize are (R, C) pairs like tools.OrderedPair."""
    buf = np.ascontiguousarray(unwrappedMatrix, dtype=np.float64).reshape(bufferSize[0], bufferSize[1])
    is_tiff = filename.endswith((".tif", ".tiff"))
    is_asc = filename.endswith(".asc")
    if not (is_tiff or is_asc):
        raise ValueError("Not a valid output file extension. Only accepts tif and ascii files (.tif/.tiff and .asc).")
    rows, cols = totalSize
    r0, c0 = offsets
    if is_asc:
        # Float AAIGrid, whole raster only (GDAL's AAIGrid driver cannot be updated in place either)
        if (r0, c0) != (0, 0) or tuple(bufferSize) != (rows, cols):
            raise ValueError(".asc output supports only whole-raster writes")
        with open(filename, "w") as f:
            f.write(_asc_header(cols, rows, GDINFO.XMin, GDINFO.YMin - rows * GDINFO.YCell, GDINFO.XCell))
            for r in range(rows):
                f.write("".join(f" {v:.17g}" for v in buf[r]) + "\n")
        return
    if create:
        _tiff_create(filename, GDINFO, rows, cols)
    frows, fcols, data_off, _ = _tiff_layout(filename)
    if r0 < 0 or c0 < 0 or r0 + buf.shape[0] > frows or c0 + buf.shape[1] > fcols:
        raise ValueError("Access window out of range in RasterIO()")
    with open(filename, "r+b") as f:
        if c0 == 0 and buf.shape[1] == fcols:
            f.seek(data_off + r0 * fcols * 8)
            f.write(buf.astype("<f8", copy=False).tobytes())
        else:
            for i in range(buf.shape[0]):
                f.seek(data_off + ((r0 + i) * fcols + c0) * 8)
                f.write(buf[i].astype("<f8", copy=False).tobytes())


def ReadTiff(filename: str):
    """(raster float64 [rows, cols], x_min, y_min, x_cell, y_cell) of a GTiff written by WriteGDAL."""
    rows, cols, off, xf = _tiff_layout(filename)
    z = np.fromfile(filename, dtype="<f8", count=rows * cols, offset=off).reshape(rows, cols)
    return z, xf[3], xf[7], xf[0], xf[5]


def TransferType(src: str, dst: str, outputType: str):
    """io_gdal.go:115-132 for the one use the IDW tests make of it (idw_test.go:57): GTiff -> Int16
    AAIGrid.  gdal_translate rounds half away from zero and clamps; the AAIGrid writer puts
    yllcorner = yMin - rows*cell for the south-up transform and emits rows in raster order."""
    if outputType != "Int16" or not dst.endswith(".asc"):
        raise ValueError("only -ot Int16 to .asc is supported")
    z, x_min, y_min, x_cell, y_cell = ReadTiff(src)
    q = np.where(np.isnan(z), 0.0, np.sign(z) * np.floor(np.abs(z) + 0.5))
    q = np.clip(q, -32768, 32767).astype(np.int64)
    rows, cols = z.shape
    with open(dst, "w") as f:
        f.write(_asc_header(cols, rows, x_min, y_min - rows * y_cell, x_cell))
        for r in range(rows):
            f.write("".join(f" {int(v)}" for v in q[r]) + "\n")


def SameFiles(f1: str, f2: str) -> bool:
    """tools.SameFiles (tools/utils.go:211-219): byte compare."""
    if os.path.getsize(f1) != os.path.getsize(f2):
        return False
    with open(f1, "rb") as a, open(f2, "rb") as b:
        return a.read() == b.read()
