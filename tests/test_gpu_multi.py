"""Multi-GPU parity (needs >= 2 B200s; skipped otherwise): the single-process context of the Go host
(v2r_idw_init(N): NCCL broadcast of the points, row bands, copy-out either straight from every device into a pinned
host raster or by an NCCL gather to device 0 for a pageable one) must return exactly the raster of a single-GPU
solve -- every cell is computed by the same kernel on the same inputs, and a cell's sums are defined over point
segments, not over the way the launch was cut."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("n_gpus", [2, 4, 8])
def test_single_process_multi_gpu_equals_single_gpu(n_gpus):
    if _ngpu() < n_gpus:
        pytest.skip(f"needs {n_gpus} GPUs")
    import v2r_b200 as V
    from v2r_b200 import _lib
    from v2r_b200.idw import make_grid
    rng = np.random.default_rng(9)
    xi, yi = V.Info(0, 299, 1.0), V.Info(0, 1099, 1.0)
    ps = V.MakeCoordSpace(rng.uniform(0, 300, 5000), rng.uniform(0, 1100, 5000), rng.normal(0, 30, 5000), xi, yi)
    grid = make_grid(xi, yi)
    exps = [1.0, 1.5, 2.0, 2.5]
    with V.IdwContext(1) as c1:
        want = c1.solve(ps, grid, exps)
        want32 = c1.solve(ps, grid, [2.0], _lib.FP32)
    with V.IdwContext(n_gpus) as cn:
        got = cn.solve(ps, grid, exps)                               # pageable raster: NCCL gather path
        assert np.array_equal(got, want)
        pinned = V.PinnedArray(want.shape)                           # page-locked raster: every device copies its own band
        try:
            cn.solve(ps, grid, exps, out=pinned.array)
            assert np.array_equal(pinned.array, want)
            pinned.array[:] = 0
            for r0, band in cn.stream(ps, grid, exps, band_rows=48, pinned=True):
                pinned.array[:, r0:r0 + band.shape[1]] = band
            assert np.array_equal(pinned.array, want)
        finally:
            pinned.free()
        assert np.array_equal(cn.solve(ps, grid, [2.0], _lib.FP32), want32)
        parts = np.full_like(want, np.nan)
        for r0, band in cn.stream(ps, grid, exps, band_rows=64):     # bands arrive device by device
            parts[:, r0:r0 + band.shape[1]] = band
        assert np.array_equal(parts, want)
        sub = cn.solve_rows(ps, grid, 100, 333, exps)               # ragged sub-range over N devices
        assert np.array_equal(sub, want[:, 100:433])
        assert cn.last_timing()["kernel_ms"] > 0


_TORCHRUN_SCRIPT = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["V2R_ROOT"])
import v2r_b200 as V
from v2r_b200.distributed import solve_distributed
from v2r_b200.idw import make_grid
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
xi, yi = V.Info(0, 199, 1.0), V.Info(0, 611, 1.0)
grid = make_grid(xi, yi)
exps = np.array([0.5, 1.0, 1.5, 2.0, 2.5])
ps = None
if rank == 0:                       # only rank 0 owns the points, like the Go host
    rng = np.random.default_rng(12)
    ps = V.MakeCoordSpace(rng.uniform(0, 200, 4000), rng.uniform(0, 612, 4000), rng.normal(0, 30, 4000), xi, yi)
full = solve_distributed(ps, grid, exps)
if rank == 0:
    with V.IdwContext(device=0) as c:
        want = c.solve(ps, grid, exps)
    assert np.array_equal(full.cpu().numpy(), want), "row-band sharded raster differs from the single-GPU raster"
    print("DIST_OK")
dist.barrier()
dist.destroy_process_group()
'''


def test_torchrun_row_bands_over_nccl(tmp_path):
    """one process per GPU (v2r_b200/distributed.py): NCCL broadcast of the points, v2r_idw_launch_rows per band,
    point-to-point gather to rank 0 -- bit-equal to the single-GPU raster"""
    import os
    import subprocess
    import sys
    n = min(_ngpu(), 4)
    if n < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "dist_check.py"
    script.write_text(_TORCHRUN_SCRIPT)
    env = dict(os.environ, V2R_ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0 and "DIST_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-3000:]
