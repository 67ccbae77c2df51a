"""The oracle's stand-ins for the OpenCL builtins against numpy float64, in ULPs of the float32
result.  Bounds are the ones the OpenCL 2.0 spec allows a conforming device (log 3, cos 4, atan2pi 6,
pow 16 ULP); the kernels are in fact much tighter, and those measured bounds are asserted."""
import numpy as np


def ulp_err(got, ref):
    ref32 = ref.astype(np.float32)
    ulp = np.spacing(np.abs(ref32)).astype(np.float64)
    return np.abs(got.astype(np.float64) - ref) / ulp


def test_log(oracle_lib):
    rng = np.random.default_rng(0)
    r = rng.integers(1, 2 ** 32, size=1_000_000, dtype=np.uint64).astype(np.float32) / np.float32(4294967296.0)
    got = oracle_lib.math("log", r)
    m = np.abs(r - 1.0) > 1e-3  # relative ULPs are meaningless where log -> 0
    assert ulp_err(got[m], np.log(r[m].astype(np.float64))).max() <= 1.0
    assert np.abs(got[~m] - np.log(r[~m].astype(np.float64))).max() < 1e-9
    sp = oracle_lib.math("log", np.array([0.0, 1.0, 2.0 ** -32, 0.5], np.float32))
    assert sp[0] == -np.inf and sp[1] == 0.0
    assert abs(sp[2] + 32 * np.log(2.0)) < 2e-6 and abs(sp[3] + np.log(2.0)) < 1e-7


def test_cos(oracle_lib):
    rng = np.random.default_rng(1)
    u = rng.random(1_000_000).astype(np.float32)
    th = (np.float32(6.28318530717958647692) * u).astype(np.float32)
    got = oracle_lib.math("cos", th)
    ref = np.cos(th.astype(np.float64))
    assert np.abs(got - ref).max() < 1.2e-7
    big = np.abs(ref) > 1e-2
    assert ulp_err(got[big], ref[big]).max() <= 2.0
    assert oracle_lib.math("cos", np.array([0.0], np.float32))[0] == 1.0


def test_atan2pi(oracle_lib):
    rng = np.random.default_rng(2)
    y = rng.standard_normal(500_000).astype(np.float32)
    x = rng.standard_normal(500_000).astype(np.float32)
    got = oracle_lib.math("atan2pi", y, x)
    ref = np.arctan2(y.astype(np.float64), x.astype(np.float64)) / np.pi
    assert np.abs(got - ref).max() < 2e-7
    assert np.all(np.abs(got) <= 1.0)
    sp = oracle_lib.math("atan2pi", np.array([0, 1, -1, 0, 0], np.float32), np.array([0, 0, 0, 1, -1], np.float32))
    assert list(sp) == [0.0, 0.5, -0.5, 0.0, 1.0]


def test_pow(oracle_lib):
    rng = np.random.default_rng(3)
    x = rng.random(500_000).astype(np.float32)
    y = (rng.random(500_000) * 100).astype(np.float32)
    got = oracle_lib.math("pow", x, y)
    ref = np.power(x.astype(np.float64), y.astype(np.float64))
    normal = ref > 1.2e-38
    assert ulp_err(got[normal], ref[normal]).max() <= 0.51
    assert np.abs(got[~normal] - ref[~normal]).max() < 1e-44
    sp = oracle_lib.math("pow", np.array([0, 0.5, 1, 0, 0.25], np.float32), np.array([25, 0, 3, 0, 0.5], np.float32))
    assert list(sp) == [0.0, 1.0, 1.0, 1.0, 0.5]


def test_sqrt_is_ieee(oracle_lib):
    rng = np.random.default_rng(4)
    x = rng.random(100_000).astype(np.float32) * 100
    assert np.array_equal(oracle_lib.math("sqrt", x), np.sqrt(x))


def test_schlick_double(oracle_lib):
    rng = np.random.default_rng(5)
    mu = (0.3 + rng.random(100_000) * 3).astype(np.float32)
    c = rng.random(100_000).astype(np.float32)
    got = oracle_lib.math("schlick", mu, c)
    r0 = ((1.0 - mu.astype(np.float64)) / (1.0 + mu.astype(np.float64))).astype(np.float32)
    r0 = (r0 * r0).astype(np.float32).astype(np.float64)
    ref = r0 + (1.0 - r0) * (1.0 - c.astype(np.float64)) ** 5
    assert ulp_err(got, ref).max() <= 0.51
