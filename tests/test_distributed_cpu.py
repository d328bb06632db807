"""World-size-2 gloo test of the N > 1 host path (row-band sharding, point broadcast, band gather) on
CPU tensors.  The per-band compute is injected (the CPU oracle stands in for the CUDA kernel, which is
test-only use of oracle/), so what is checked is the plumbing: every rank receives rank 0's points,
bands tile the grid exactly once, ragged last bands and empty bands work, rank 0 ends up with the same
raster as an unsharded solve."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from v2r_b200.distributed import shard_rows


def test_shard_rows_tiles_the_grid():
    for rows in (0, 1, 13, 16, 31, 32, 33, 100, 4096, 16384, 32768 + 5):
        for world in (1, 2, 4, 8):
            covered = []
            for r in range(world):
                a, n = shard_rows(rows, world, r)
                assert n >= 0 and (n == 0 or a % 32 == 0)
                covered += list(range(a, a + n))
            assert covered == list(range(rows)), (rows, world)
    assert shard_rows(16384, 8, 3) == (6144, 2048)           # config c4 on 8 GPUs: equal 2048-row bands


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, rows, cols, n, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle as O
    import v2r_b200 as V
    from v2r_b200.distributed import solve_distributed
    from v2r_b200.idw import make_grid
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    xi, yi = V.Info(0, cols - 1, 1.0), V.Info(0, rows - 1, 1.0)
    grid = make_grid(xi, yi, rows, cols)
    exps = np.array([1.5, 2.0])
    ps = None
    if rank == 0:   # only rank 0 owns the points, like the Go host
        rng = np.random.default_rng(3)
        ps = V.MakeCoordSpace(rng.uniform(0, cols, n), rng.uniform(0, rows, n), rng.normal(0, 10, n), xi, yi)

    def band_fn(pts, grid, row0, nrows, exps):
        x, y, w, kr, kc = (t.numpy() for t in pts)
        ops = O.PointSet(x, y, w, kr, kc)
        oi = O.Info(0, cols - 1, 1.0), O.Info(0, rows - 1, 1.0)
        out = np.empty((len(exps), nrows, cols))
        for e, p in enumerate(exps):
            out[e] = O.solve_rows(ops, oi[0], oi[1], float(p), cols, row0, nrows)
        return torch.from_numpy(out)

    full = solve_distributed(ps, grid, exps, device=torch.device("cpu"), band_fn=band_fn)
    if rank == 0:
        ops = O.PointSet(ps.x, ps.y, ps.w, ps.key_r, ps.key_c)
        want = np.stack([O.full_solve(ops, O.Info(0, cols - 1, 1.0), O.Info(0, rows - 1, 1.0), float(p), rows, cols) for p in exps])
        q.put(bool(np.array_equal(full.numpy(), want, equal_nan=True)))
    else:
        assert full is None
    dist.destroy_process_group()


@pytest.mark.parametrize("rows,cols,n", [(70, 19, 50), (20, 9, 7), (64, 8, 0)])
def test_gloo_world2_row_bands(rows, cols, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, rows, cols, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
