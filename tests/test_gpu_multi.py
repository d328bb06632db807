"""Multi-GPU parity (needs >= 2 B200s; skipped otherwise): the single-process context of the Go host
(v2r_idw_init(N): NCCL broadcast of the points, row bands, NCCL gather to device 0) must return exactly
the raster of a single-GPU solve -- every cell is computed by the same kernel on the same inputs."""
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
        got = cn.solve(ps, grid, exps)
        assert np.array_equal(got, want)
        assert np.array_equal(cn.solve(ps, grid, [2.0], _lib.FP32), want32)
        parts = np.full_like(want, np.nan)
        for r0, band in cn.stream(ps, grid, exps, band_rows=64):     # bands arrive device by device
            parts[:, r0:r0 + band.shape[1]] = band
        assert np.array_equal(parts, want)
        sub = cn.solve_rows(ps, grid, 100, 333, exps)               # ragged sub-range over N devices
        assert np.array_equal(sub, want[:, 100:433])
        assert cn.last_timing()["kernel_ms"] > 0
