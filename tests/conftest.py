"""Shared fixtures.  `-m "not gpu"` runs on a CPU-only box; `-m gpu` needs a B200.

The checkers (oracle/) are used here only to CHECK: the product path (raytracerwin_b200) never
imports them.  /root/reference is never read at test time: the reference's compiled hot path
(oracle/_ref/*.so) and its staged assets (assets/_ref/Data) are build outputs that travel with
the repository snapshot, and the committed fixtures under tests/golden/ cover the case where
they are absent.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DATA_DIR = os.path.join(ROOT, "assets", "_ref", "Data")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _have_gpu():
    try:
        import raytracerwin_b200 as rt
        return rt.load_library().rt_gpu_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def rt():
    import raytracerwin_b200 as m
    m.load_library()
    return m


@pytest.fixture(scope="session")
def data_dir():
    if not os.path.isdir(DATA_DIR):
        pytest.skip("assets/_ref/Data not staged (run __graft_entry__.build() where /root/reference exists)")
    return DATA_DIR


@pytest.fixture(scope="session")
def ref():
    from oracle import bindings
    if not bindings.ref_available():
        pytest.skip("oracle/_ref/libref_oracle.so not built")
    return bindings.RefOracle()


@pytest.fixture(scope="session")
def ref_counting():
    """The ray-counting twin of the reference (RayTracerScene.cpp built with -finstrument-functions): same
    results, and ref.render(...)["rays"] = the number of FindIntersectionWithScene calls it made."""
    from oracle import bindings
    if not os.path.exists(bindings.REF_COUNT_LIB):
        pytest.skip("oracle/_ref/libref_oracle_count.so not built")
    return bindings.RefOracle(counting=True)


@pytest.fixture(scope="session")
def port():
    from oracle import bindings
    if not bindings.port_available():
        pytest.skip("oracle/librt_oracle.so not built")
    return bindings.PortOracle()


@pytest.fixture(scope="session")
def gpu(rt):
    ctx = rt.GpuContext(0)
    yield ctx
    ctx.close()


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def same_bits(a, b):
    """Bit-identical, except that NaNs only have to coincide in position (x86 and CUDA give NaNs different
    sign/payload bits; the reference itself produces NaN radiance at a few degenerate hits)."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    na, nb = np.isnan(a), np.isnan(b)
    return a.shape == b.shape and np.array_equal(na, nb) and np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb])
