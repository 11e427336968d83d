import importlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    return oracle_lib.load_oracle()


@pytest.fixture(scope="session")
def blob():
    """The product package (ctypes view of libb200blas.so). Builds on demand here;
    on the GPU box the prebuilt in-tree .so is used."""
    pkg = importlib.import_module("gpu-blas-offload-benchmark_b200")
    if not os.path.exists(pkg.LIB_PATH):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "gpu-blas-offload-benchmark_b200", "build.py"),
                               "--no-harness"])
    pkg.load()
    return pkg


@pytest.fixture(scope="session")
def ctx(blob):
    c = blob.Context()
    yield c
    c.close()
