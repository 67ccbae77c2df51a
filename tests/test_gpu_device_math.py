"""GPU checks of the range-tested fast forms in device_math.cuh (rcp_sqrt_ inside normalize, sqrt_x2 inside
random_float_normal_x2) against the correctly rounded intrinsics they stand in for -- on EVERY 32-bit pattern, on the
device, and against numpy's correctly rounded float32 sqrt and division on a sample (render.cl itself only says
`normalize` / `sqrt`; the contract is DESIGN.md section 2: correctly rounded IEEE operations)."""
import numpy as np
import pytest

from conftest import assert_bit_equal
from util import make_tracer

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tracer(sky):
    from simple_raytracer_b200 import scenes
    tr = make_tracer(scenes.config1(16, 16), sky)
    yield tr
    tr.close()


def patterns(n, seed):
    rng = np.random.default_rng(seed)
    x = rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32).view(np.float32)
    edge = np.array([0.0, -0.0, 1.0, 2.0 ** -100, np.nextafter(np.float32(2.0 ** -100), np.float32(0)), 2.0 ** 100,
                     np.nextafter(np.float32(2.0 ** 100), np.float32(0)), 2.0 ** -101, 2.0 ** -126, 1e-45, 3.4028235e38, np.inf, -1.0,
                     np.nan, 0.99999994, 1.0000001, 4.0, 2.0, 3.0], np.float32)
    x[:edge.size] = edge
    return x


def test_rcp_sqrt_equals_the_two_intrinsics_on_every_pattern(tracer):
    z = np.zeros(1 << 20, np.float32)
    assert tracer.debug_math("rcp_sqrt_all_patterns", z).sum() == 0


def test_sqrt_x2_equals_the_intrinsic_on_every_pattern(tracer):
    z = np.zeros(1 << 20, np.float32)
    assert tracer.debug_math("sqrt_x2_all_patterns", z).sum() == 0


def test_fast_forms_against_numpy(tracer):
    x = np.abs(patterns(1 << 20, 3))
    y = np.abs(patterns(1 << 20, 4))
    with np.errstate(all="ignore"):
        want = (np.float32(1.0) / np.sqrt(x)).astype(np.float32)  # float32 sqrt and division are correctly rounded
        assert_bit_equal(want, tracer.debug_math("rcp_sqrt", x), "rcp_sqrt_")
        assert_bit_equal(want, tracer.debug_math("rcp_of_sqrt", x), "rcp_(sqrt_())")
        assert_bit_equal(np.sqrt(x), tracer.debug_math("sqrt_x2_lo", x, y), "sqrt_x2 low half")
        assert_bit_equal(np.sqrt(y), tracer.debug_math("sqrt_x2_hi", x, y), "sqrt_x2 high half")
