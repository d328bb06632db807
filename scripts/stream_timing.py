"""times the streaming API on c2 with the reference's default chunk height (200 rows) against one whole-grid solve"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import v2r_b200 as V
from v2r_b200 import workloads
from v2r_b200.idw import make_grid

x, y, w, xi, yi, ex = workloads.make("c2")
ps = V.MakeCoordSpace(x, y, w, xi, yi)
g = make_grid(xi, yi)
with V.IdwContext(1) as c:
    full = c.solve(ps, g, ex)
    t = time.perf_counter(); full = c.solve(ps, g, ex); t_full = time.perf_counter() - t
    for band in (200, 32, 1000):
        got = np.empty_like(full)
        t = time.perf_counter()
        nb = 0
        for r0, b in c.stream(ps, g, ex, band_rows=band):
            got[:, r0:r0 + b.shape[1]] = b
            nb += 1
        dt = time.perf_counter() - t
        print(f"stream band_rows={band}: {nb} bands, {dt:.3f} s vs whole-grid {t_full:.3f} s, equal={np.array_equal(got, full)}", c.last_timing())
