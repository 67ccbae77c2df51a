"""CPU check of the claim the CUDA sweep filter rests on: whenever the filter rejects a
(ray, triangle) pair, the reference's own u test (render.cl:250-261) rejects it too.  oracle/filter_check.c restates
the filter's arithmetic operation for operation and throws adversarial pairs at it: rays aimed within ulps of u = 0
and u = 1, rays grazing the triangle's plane (det ~ 0), slivers, sizes 1e-4 .. 1e3, up to 1e6 units from the origin."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")


def run(n, seed, margin_scale, fn="filter_check"):
    so = os.path.join(HERE, "libfilter_check.so")
    src = [os.path.join(HERE, f) for f in ("filter_check.c", "oracle_math.h")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in src):
        subprocess.check_call(["make", "-s", "-C", HERE, "-B", "libfilter_check.so"])
    lib = ctypes.CDLL(so)
    getattr(lib, fn).argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_float, ctypes.c_void_p]
    out = np.zeros(4, np.uint64)
    getattr(lib, fn)(n, seed, margin_scale, out.ctypes.data_as(ctypes.c_void_p))
    return dict(zip(("pairs", "filter_rejects", "violations", "reference_rejects"), (int(v) for v in out)))


def test_filter_never_rejects_what_the_reference_accepts():
    r = run(100_000_000, 2026, 1.0)
    assert r["pairs"] > 99_000_000 and r["filter_rejects"] > 0.3 * r["pairs"]  # the filter does reject
    assert r["violations"] == 0, r


def test_the_check_has_teeth_without_margins_the_claim_fails():
    """Same pairs, margins removed: the pre-multiplied operands' rounding now produces wrong rejects, i.e. the
    adversarial distribution does reach the cases the margins exist for.  (They first appear when the margins are
    cut to ~1 % of their value: the analytic bound is worst-case.)"""
    r = run(100_000_000, 2026, 0.0)
    assert r["violations"] > 1000, r


def test_two_strip_filter_never_rejects_what_the_reference_accepts():
    """The u-and-v filter of small models (tri_filter_sweep_uv) against the reference's u, v, u + v and t tests
    together: a pair the filter drops can never become a hit.  Extra pairs are aimed within ulps of v = 0 and v = 1."""
    r = run(100_000_000, 2027, 1.0, "filter_check_uv")
    assert r["pairs"] > 99_000_000 and r["filter_rejects"] > 0.4 * r["pairs"]
    assert r["violations"] == 0, r
    assert r["filter_rejects"] <= r["reference_rejects"]


def test_two_strip_check_has_teeth():
    r = run(100_000_000, 2027, 0.0, "filter_check_uv")
    assert r["violations"] > 1000, r
