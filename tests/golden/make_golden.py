"""Copies the reference's own IDW test fixtures into tests/golden/ (run in the build container,
where /root/reference is mounted; the GPU box has no reference).  These are test DATA of
Dewberry/v2r (features/idw/idw_test_files/), not source code:

  idw_in.txt                 input of TestIDW            (features/idw/idw_test.go:17)
  idw_correct_step1-1.asc    golden, p = 1.7, step 1     (features/idw/idw_test.go:40,43,67)
  idw_correct_step2-2.asc    orphan golden (step-2 grid evaluated with step-1 keys)
  input.gpkg, gpkg_proj.txt  gpkg reader fixture         (tools/processing/gpkg_test.go:71-111)
  gpkg_points.json           the 14 expected points of gpkg_test.go:12-27 + Info of :63-69, transcribed
"""
import json
import os
import shutil

SRC = "/root/reference/features/idw/idw_test_files"
DST = os.path.dirname(os.path.abspath(__file__))

for name in ["idw_in.txt", "idw_correct_step1-1.asc", "idw_correct_step2-2.asc", "input.gpkg", "gpkg_proj.txt"]:
    shutil.copyfile(os.path.join(SRC, name), os.path.join(DST, name))
    os.chmod(os.path.join(DST, name), 0o644)

# tools/processing/gpkg_test.go:12-27 (expected points) and :63-69 (expected Info at steps 1.0 / 2.0)
points = [
    [1.203154304903684e+07, 3.953694544257536e+06, 0], [1.2039094668343753e+07, 3.9424537521008854e+06, 0],
    [1.2052295156865465e+07, 3.93148004717589e+06, 0], [1.2076264960092723e+07, 3.917394837044359e+06, 0],
    [1.209133966962186e+07, 3.899012514405147e+06, 0], [1.2020401943588393e+07, 3.9463317049700455e+06, 2.5],
    [1.2025601603362886e+07, 3.9380775540176313e+06, 2.9], [1.203495724450095e+07, 3.92078279952575e+06, 7],
    [1.2055886381444821e+07, 3.898699015764322e+06, 6.3], [1.207451364898982e+07, 3.885548111378168e+06, 2],
    [1.2049260510355683e+07, 3.952539493517972e+06, 3.1], [1.205195120503224e+07, 3.947886314968573e+06, 2.8],
    [1.2096825067172093e+07, 3.9268244968121317e+06, 3.9], [1.2080323349286443e+07, 3.9412021225619004e+06, 4.6]]
info = {"x": {"min": 1.2020401943588393e+07, "max": 1.2096825067172093e+07, "step": 1.0},
        "y": {"min": 3.885548111378168e+06, "max": 3.953694544257536e+06, "step": 2.0}}
with open(os.path.join(DST, "gpkg_points.json"), "w") as f:
    json.dump({"points": points, "info": info, "layer": "wsels", "field": "elevation"}, f, indent=1)
print("golden fixtures written to", DST)
