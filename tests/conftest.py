import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def sky():
    from simple_raytracer_b200 import scenes
    return scenes.procedural_skybox()


@pytest.fixture(scope="session")
def small_sky():
    """64x32 sky for CPU-only tests that do not care about texel detail."""
    from simple_raytracer_b200 import scenes
    return scenes.procedural_skybox(64, 32, seed=11)


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle
    oracle.build()
    return oracle


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_bit_equal(a, b, what=""):
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    assert a.shape == b.shape, (a.shape, b.shape)
    neq = bits(a) != bits(b)
    # NaN payloads may differ between CPU and GPU; both-NaN counts as equal
    neq &= ~(np.isnan(a) & np.isnan(b))
    n = int(neq.sum())
    if n:
        i = np.argwhere(neq)[0]
        raise AssertionError(f"{what}: {n} of {a.size} values differ; first at {tuple(i)}: "
                             f"{a[tuple(i)]!r} vs {b[tuple(i)]!r}")


@pytest.fixture(params=["plain", "wavefront"])
def schedule(request, monkeypatch):
    """Run a parity test under BOTH warp schedules of the analytic / small-model kernel builds (srt_set_schedule through
    the Tracer's SRT_SCHEDULE hook): the automatic choice would give every small test launch the plain schedule and
    only full-size launches the wavefront one."""
    monkeypatch.setenv("SRT_SCHEDULE", request.param)
    return request.param
