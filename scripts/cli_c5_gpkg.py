"""config-5-shaped input through the C++ host: 2 M points in a GeoPackage (EPSG:2284 offsets), `v2r idw -g --layer --field`,
four exponents in one pass; the step is 1600 ft instead of 100 ft (2048^2 grid) so that one GPU finishes in seconds."""
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from oracle import oracle as O
from v2r_b200 import raster_io, workloads

x, y, w, xi, yi, ex = workloads.make("c5")
t = time.perf_counter(); workloads.write_gpkg("/tmp/c5.gpkg", x, y, w); print(f"wrote /tmp/c5.gpkg ({len(x)} points) in {time.perf_counter() - t:.1f} s")
t = time.perf_counter()
p = subprocess.run(["v2r_b200/bin/v2r", "idw", "-g", "-f", "/tmp/c5.gpkg", "--layer", "wsels", "--field", "elevation", "--sx", "1600", "--sy", "1600",
                    "--es", "1", "--ee", "3", "--ei", "0.5", "--outPath", "/tmp/", "-c", "--cy", "256"], capture_output=True, text=True)
print(f"v2r idw -g ... rc={p.returncode} wall {time.perf_counter() - t:.2f} s")
print("\n".join(l for l in p.stderr.splitlines() if any(k in l for k in ("RowsXCols", "GPU:", "Completed", "FATAL", "~50%", "~100%"))))
oxi, oyi = O.Info(float(x.min()), float(x.max()), 1600.0), O.Info(float(y.min()), float(y.max()), 1600.0)
ops = O.make_coord_space(x, y, w, oxi, oyi)
rng = np.random.default_rng(1)
for pw in (1.0, 1.5, 2.0, 2.5):
    z = raster_io.ReadTiff(f"/tmp/c5_step1600-1600chunkedpow{pw:.1f}.tiff")[0]
    rr, cc = rng.integers(0, z.shape[0], 24), rng.integers(0, z.shape[1], 24)
    ref = O.solve_cells(ops, oxi, oyi, pw, rr, cc)
    print(f"pow {pw}: raster {z.shape}, max rel err vs oracle on 24 cells = {np.max(np.abs(z[rr, cc] - ref) / np.abs(ref)):.2e}")
