"""Accuracy of the device arithmetic in v2r_b200/csrc/idw_math.cuh -- the table-driven log2/exp2 of the `general`
exponent class and the seed + one-cubic-step reciprocal / rsqrt / fourth-root of the other classes (with the MUFU
seeds emulated at their documented 20-bit worst case), and the p = 2 shared-reciprocal forms (two pairs, and the four
pairs of rcp_shared4 over their whole guarded domain of squared distances, 2^-240 .. 2^121) -- checked on the CPU: the header is __host__ __device__, so nvcc builds it into a host program that compares 2 M samples over
60 decades with long-double libm, plus every special value the kernels rely on (d2 = 0 -> -inf -> h = +inf -> NaN
cell; saturation to 0 / +inf; NaN passes through)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fast_log2_exp2_accuracy_and_specials(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "check_idw_math")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-I", os.path.join(ROOT, "v2r_b200", "csrc"),
                           "-o", exe, os.path.join(ROOT, "tests", "native", "check_idw_math.cu")])
    out = subprocess.check_output([exe], text=True).split()
    max_abs_log2, max_rel_exp2, max_rel_pow, max_rel_fold, bad = float(out[0]), float(out[1]), float(out[2]), float(out[3]), int(out[4])
    assert bad == 0
    assert max_abs_log2 < 2.2e-14       # |log2| <= 100 here: the series' 5e-15 plus 1 ulp of the result
    assert max_rel_exp2 < 4e-15         # 3-coefficient series on a 512-entry table: 2.2e-15 by design
    assert max_rel_pow < 1e-13          # d2^(-p/2) for p < 10 over 60 decades; the parity bar is 1e-12
    assert max_rel_fold < 1e-13         # folded single-exponent form over the WHOLE normal range (|log2| <= 1022)
