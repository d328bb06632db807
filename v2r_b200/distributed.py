"""Row-band sharding of the IDW grid over one process per GPU (torch.distributed, NCCL over NVLink).

The north_star's multi-GPU shape: the de-duplicated point set is replicated with ONE broadcast from
rank 0, each rank computes a contiguous band of rows (every cell costs the same N points, so equal
row counts balance), and the bands are gathered to rank 0 for the unchanged raster writer
(processing.WriteGDAL with offsets = (row0, 0), tools/processing/io_gdal.go:72-73).

There is no reduction across GPUs and nothing to fuse into the kernel: the only collectives are the
broadcast before and the gather after.  The same logic runs under gloo on CPU tensors in the tests
(with an injected band function), which is how the N > 1 host path is covered without GPUs.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .idw import PointSet


def shard_rows(rows: int, world: int, rank: int, align: int = 32):
    """Contiguous band [row0, row0 + nrows) of `rank`: ceil(rows/world) rounded up to `align` rows
    (whole CTA tiles), the last ranks may get fewer rows or none.  Same split as v2r_idw_init(n)."""
    per = -(-rows // world)
    per = -(-per // align) * align
    a = min(rank * per, rows)
    b = min(a + per, rows)
    return a, b - a


def broadcast_points(ps: PointSet | None, device, group=None):
    """Rank 0 passes the PointSet, the others None.  Returns five device tensors (x, y, w, key_r, key_c), all
    views of ONE packed [5, stride] float64 block (the int64 keys travel as bit patterns), so the whole point set
    is replicated by a single broadcast (5*N*8 bytes over NVLink) after a one-element broadcast of the count."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    n = torch.tensor([len(ps) if rank == 0 else 0], dtype=torch.int64, device=device)
    dist.broadcast(n, 0, group=group)
    n = int(n.item())
    stride = (max(n, 1) + 1) // 2 * 2          # even: every row stays 16-byte aligned for the TMA tile loads
    block = torch.empty((5, stride), dtype=torch.float64, device=device)
    if rank == 0:
        host = np.zeros((5, stride), dtype=np.float64)
        host[0, :n], host[1, :n], host[2, :n] = ps.x, ps.y, ps.w
        host[3, :n] = np.ascontiguousarray(ps.key_r, dtype=np.int64).view(np.float64)
        host[4, :n] = np.ascontiguousarray(ps.key_c, dtype=np.int64).view(np.float64)
        block.copy_(torch.from_numpy(host))
    dist.broadcast(block, 0, group=group)
    return [block[0, :n], block[1, :n], block[2, :n], block[3, :n].view(torch.int64), block[4, :n].view(torch.int64)]


def cuda_band(dev_points, grid, row0, nrows, exps, precision, device_index):
    """This rank's band on its GPU through the C ABI (v2r_idw_launch_rows on torch's current stream)."""
    import torch
    L = _lib.load()
    dx, dy, dw, dkr, dkc = dev_points
    exps = np.ascontiguousarray(exps, dtype=np.float64)
    out = torch.empty((len(exps), nrows, grid.cols), dtype=torch.float64, device=dx.device)
    if nrows > 0 and grid.cols > 0:
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(L.v2r_idw_launch_rows(device_index, C.c_void_p(stream), C.c_void_p(dx.data_ptr()), C.c_void_p(dy.data_ptr()),
                                         C.c_void_p(dw.data_ptr()), C.c_void_p(dkr.data_ptr()), C.c_void_p(dkc.data_ptr()),
                                         dx.numel(), C.byref(grid), row0, nrows, C.c_void_p(exps.ctypes.data), len(exps),
                                         precision, C.c_void_p(out.data_ptr())))
    return out


def gather_bands(band, grid, n_exps: int, group=None):
    """Bands [E, nrows_r, cols] of every rank -> rank 0's full raster [E, rows, cols] (None elsewhere).
    Point-to-point sends (bands can be ragged), posted together so they overlap on NVSwitch."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if rank != 0:
        if band.numel():
            dist.send(band.contiguous(), 0, group=group)
        return None
    full = torch.empty((n_exps, grid.rows, grid.cols), dtype=band.dtype, device=band.device)
    r0, nr = shard_rows(grid.rows, world, 0)
    full[:, r0:r0 + nr] = band
    for src in range(1, world):
        r0, nr = shard_rows(grid.rows, world, src)
        if nr == 0 or grid.cols == 0:
            continue
        buf = torch.empty((n_exps, nr, grid.cols), dtype=band.dtype, device=band.device)
        dist.recv(buf, src, group=group)
        full[:, r0:r0 + nr] = buf
    return full


def solve_distributed(ps: PointSet | None, grid, exps, precision: int = _lib.FP64, device=None, band_fn=None, group=None):
    """Whole sweep over all ranks of the process group: broadcast -> band per rank -> gather to rank 0.
    `band_fn(dev_points, grid, row0, nrows, exps)` defaults to the CUDA path; tests inject a CPU one."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    pts = broadcast_points(ps, device, group)
    row0, nrows = shard_rows(grid.rows, world, rank)
    if band_fn is None:
        band = cuda_band(pts, grid, row0, nrows, exps, precision, device.index)
    else:
        band = band_fn(pts, grid, row0, nrows, exps)
    return gather_bands(band, grid, len(np.atleast_1d(exps)), group)
