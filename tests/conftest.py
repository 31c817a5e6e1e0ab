import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _build_oracle():
    odir = os.path.join(ROOT, "oracle")
    so = os.path.join(odir, "liboracle.so")
    srcs = [os.path.join(odir, f) for f in ("mtr_oracle.c", "mtr_oracle.h", "mtr_oracle_chain.cpp", "mtr_oracle_main.c")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", odir, "oracle"])
    return so


@pytest.fixture(scope="session")
def oracle_so():
    return _build_oracle()


@pytest.fixture(scope="session")
def gpu_ctx():
    from mtr_b200 import capi
    ctx = capi.Context(0)       # raises if there is no device: -m gpu tests must not silently pass
    yield ctx
    ctx.close()
