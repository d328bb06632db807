import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def _built():
    """A fresh checkout has no binaries (they are git-ignored): build the library, the CLI and the oracle once."""
    lib = os.path.join(ROOT, "v2r_b200", "lib", "libv2r_idw.so")
    cli = os.path.join(ROOT, "v2r_b200", "bin", "v2r")
    if not (os.path.exists(lib) and os.path.exists(cli)):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "v2r_b200", "csrc"), "-j4", "-s"])
