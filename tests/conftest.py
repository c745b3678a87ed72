"""pytest configuration: registers the `gpu` marker and builds native artefacts once per session."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `pytest -m gpu` on the GPU box)")


@pytest.fixture(scope="session")
def native_build():
    """Build the native artefacts once per session (g++ / nvcc cross-compile; no GPU needed)."""
    san = os.environ.get("B200_NATIVE_SAN", "")          # "thread" or "address": run the suite against instrumented binaries
    subprocess.run(["make", "-C", os.path.join(ROOT, "agent", "native"), "-j4"] + ([f"SAN={san}"] if san else []), check=True, capture_output=True)
    return os.path.join(ROOT, "build", "agent" + (f"-{san}" if san else ""))


@pytest.fixture(scope="session")
def coll_lib():
    lib = os.path.join(ROOT, "coll", "lib", "libb200coll.so")
    if not os.path.exists(lib):
        subprocess.run(["make", "-C", os.path.join(ROOT, "coll")], check=True, capture_output=True)
    return lib
