"""The reference-side Go binding (go/) cannot be compiled in this image (no Go toolchain), so what CAN be checked is:
every C symbol / constant the cgo files use is declared in include/v2r_idw.h and exported by the built library, the
entry points keep the reference's signatures, and the two patches apply cleanly to the reference tree when it is
present (it is not on the GPU box: that part is skipped there)."""
import ctypes
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "go", "features", "idw")
REF = "/root/reference"


def _read(*p):
    return open(os.path.join(*p)).read()


def test_cgo_symbols_exist_in_header_and_library():
    from v2r_b200 import _lib
    header = _read(ROOT, "include", "v2r_idw.h")
    src = _read(GO, "idw_gpu.go")
    used = set(re.findall(r"\bC\.(v2r_idw_\w+|V2R_IDW_\w+)", src))
    assert {"v2r_idw_init", "v2r_idw_stream_begin", "v2r_idw_stream_next", "v2r_idw_stream_end", "v2r_idw_alloc_pinned",
            "v2r_idw_free_pinned", "v2r_idw_ctx_last_error"} <= used
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in used:
        assert re.search(r"\b%s\b" % name, header), f"{name} is not declared in include/v2r_idw.h"
        if name.startswith("v2r_idw_") and name not in ("v2r_idw_ctx", "v2r_idw_grid"):
            assert hasattr(lib, name), f"{name} is not exported by libv2r_idw.so"


def test_entry_points_keep_the_reference_signatures():
    gpu = _read(GO, "idw_entry_gpu.go")
    cpu = _read(GO, "idw_entry_cpu.go")
    # features/idw/idw.go:15 and features/idw/idw_chunk.go:43, character for character
    full = ("func FullSolve(data *map[tools.OrderedPair]tools.Point, outfile string, xInfo tools.Info, yInfo tools.Info, "
            "proj string, pow float64, channel chan string) error {")
    chunk = ("func ChunkSolve(data *map[tools.OrderedPair]tools.Point, outfile string, xInfo tools.Info, yInfo tools.Info, "
             "chunkR int, chunkC int, proj string, pow float64, channel chan string) error {")
    assert full in gpu and chunk in gpu
    if os.path.isdir(REF):
        assert full in _read(REF, "features", "idw", "idw.go") and chunk in _read(REF, "features", "idw", "idw_chunk.go")
    assert gpu.startswith("//") and "//go:build v2rgpu" in gpu and "//go:build !v2rgpu" in cpu
    sig = re.search(r"func SweepSolve\(([^{]*)\{", gpu, re.S).group(1)
    assert re.sub(r"\s+", " ", sig) == re.sub(r"\s+", " ", re.search(r"func SweepSolve\(([^{]*)\{", cpu, re.S).group(1))


@pytest.mark.skipif(not os.path.isdir(REF) or shutil.which("patch") is None, reason="reference tree / patch(1) not present")
def test_patches_apply_to_the_reference(tmp_path):
    for d in ("cmd", "features"):
        shutil.copytree(os.path.join(REF, d), tmp_path / d)
    for p in sorted(os.listdir(os.path.join(ROOT, "go", "patches"))):
        subprocess.check_call(["patch", "-p1", "-s", "-i", os.path.join(ROOT, "go", "patches", p)], cwd=tmp_path)
    for f in ("idw.go", "idw_chunk.go", "idw_utils.go"):
        assert (tmp_path / "features" / "idw" / f).read_text().startswith("//go:build !v2rgpu\n\npackage idw")
    cmd = (tmp_path / "cmd" / "idw.go").read_text()
    assert "idw.SweepSolve(&data, outfiles, xInfo, yInfo, useChunking, idwChunkY, idwChunkX, proj, pows, channel)" in cmd
    assert "go idw.FullSolve" not in cmd and '"gpus"' in cmd and '"fp32"' in cmd
