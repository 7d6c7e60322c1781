import ctypes
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def paths():
    from assist_b200.synth import ephem_writer
    return ephem_writer.write_all(os.path.join(ROOT, "data"))


@pytest.fixture(scope="session")
def lib():
    from assist_b200 import _lib
    return _lib.load()


@pytest.fixture(scope="session")
def have_gpu(lib):
    return lib.assist_gpu_device_count() > 0


@pytest.fixture(scope="session")
def ref():
    import refharness as rh
    if not rh.have_ref():
        pytest.skip("oracle/_ref/libassist_ref.so not built (needs /root/reference)")
    return rh.ref_lib()


@pytest.fixture(scope="session", params=["bsp", "440"])
def fmt(request):
    return request.param


@pytest.fixture(scope="session")
def golden():
    out = {}
    for tag in ("bsp", "440"):
        out[tag] = np.load(os.path.join(ROOT, "tests", "golden", "golden_%s.npz" % tag))
    return out


def planets_path(paths, fmt):
    return paths["planets_bsp"] if fmt == "bsp" else paths["de440"]


def relerr(a, b):
    """max over systems/bodies of |a-b| / |b| (vector norms over the last axis)."""
    a = np.asarray(a); b = np.asarray(b)
    num = np.linalg.norm(a - b, axis=-1)
    den = np.linalg.norm(b, axis=-1)
    den = np.where(den == 0, 1.0, den)
    return float(np.max(num / den))
