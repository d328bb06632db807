"""end-to-end through the C++ host: write c2 as the reference's txt grammar, run `v2r idw`, read the GTiff back"""
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from v2r_b200 import raster_io, workloads

x, y, w, xi, yi, ex = workloads.make("c2")
workloads.write_txt("/tmp/c2.txt", x, y, w, xi, yi)
for extra in ([], ["-c"], ["--fp32"]):
    t = time.perf_counter()
    p = subprocess.run(["v2r_b200/bin/v2r", "idw", "-f", "/tmp/c2.txt", "--es", "2", "--ei", "1", "--outPath", "/tmp/"] + extra,
                       capture_output=True, text=True)
    dt = time.perf_counter() - t
    tail = [l for l in p.stderr.splitlines() if "GPU:" in l or "Completed" in l or "FATAL" in l]
    name = "/tmp/c2_step100-100" + ("chunked" if "-c" in extra else "") + "pow2.0.tiff"
    z = raster_io.ReadTiff(name)[0]
    print(f"v2r idw {' '.join(extra) or '(serial)'}: rc={p.returncode} wall {dt:.2f} s, raster {z.shape}, finite={bool(np.isfinite(z).all())}; " + " | ".join(tail))
