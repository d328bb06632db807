"""Records the DRAM traffic of one launch of the dominant kernel in profiles/traffic.json, the file bench.py's
`roofline.traffic` is looked up in (by kernel description + launch shape, so a stale entry can never be reported for
a different kernel or launch).

    python scripts/ncu_traffic.py <ncu --csv metrics log> <rows> <cols> <points> [--precision fp64|fp32] [--exps 2.0]

The csv comes from `scripts/gpu_run.sh ncu` (`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum ... --csv`)."""
import argparse
import csv
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("rows", type=int)
    ap.add_argument("cols", type=int)
    ap.add_argument("points", type=int)
    ap.add_argument("--precision", default="fp64")
    ap.add_argument("--exps", default="2.0")
    a = ap.parse_args()
    import numpy as np
    from v2r_b200 import _lib
    rows = [r for r in csv.reader(l for l in open(a.csv) if l.startswith('"'))]
    hdr = rows[0]
    name, unit, val, kern = hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value"), hdr.index("Kernel Name")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    got = {}
    for r in rows[1:]:
        if r[name] in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
            v = float(r[val].replace(",", ""))
            got[r[name]] = v * scale.get(r[unit], 1)
            got["kernel_name"] = r[kern]
    exps = np.array([float(v) for v in a.exps.split(",")])
    buf, ops = C.create_string_buffer(512), C.c_double()
    _lib.check(_lib.load().v2r_idw_describe(C.c_void_p(exps.ctypes.data), len(exps), 0 if a.precision == "fp64" else 1, buf, 512, C.byref(ops)))
    rec = {"kernel": buf.value.decode(), "rows": a.rows, "cols": a.cols, "points": a.points,
           "dram_bytes": int(got["dram__bytes_read.sum"] + got["dram__bytes_write.sum"]),
           "dram_read": int(got["dram__bytes_read.sum"]), "dram_write": int(got["dram__bytes_write.sum"]),
           "sass_kernel": got.get("kernel_name"), "source": os.path.basename(a.csv)}
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        doc = json.load(open(path))
    except Exception:
        doc = {}
    caps = [c for c in doc.get("captures", []) if [c["kernel"], c["rows"], c["cols"], c["points"]] != [rec["kernel"], rec["rows"], rec["cols"], rec["points"]]]
    caps.append(rec)
    json.dump({"_comment": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel (ncu --metrics ..., "
                           "scripts/gpu_run.sh ncu); bench.py reports an entry only for exactly this kernel and launch shape",
               "captures": caps}, open(path, "w"), indent=1)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
