"""Closed-form cases for the four intersection routines (render.cl:180-290) and edge semantics."""
import numpy as np


def test_sphere(oracle_lib):
    hit, t = oracle_lib.intersect("sphere", (0, 0, 5), (0, 0, -1), (0, 0, 0), (1,))
    assert hit and t == 4.0                       # front surface
    hit, t = oracle_lib.intersect("sphere", (0, 0, 0), (0, 0, -1), (0, 0, 0), (2,))
    assert hit and t == 2.0                       # origin inside: far root (render.cl:197-198)
    hit, _ = oracle_lib.intersect("sphere", (0, 0, 5), (0, 0, 1), (0, 0, 0), (1,))
    assert not hit                                # behind the ray
    hit, _ = oracle_lib.intersect("sphere", (0, 3, 5), (0, 0, -1), (0, 0, 0), (1,))
    assert not hit                                # disc < 0
    hit, t = oracle_lib.intersect("sphere", (0, 1, 5), (0, 0, -1), (0, 0, 0), (1,))
    assert hit and t == 5.0                       # tangent: disc == 0 is accepted


def test_plane(oracle_lib):
    hit, t = oracle_lib.intersect("plane", (0, 2, 0), (0, -1, 0), (0, 0, 0), (0, 1, 0))
    assert hit and t == 2.0
    hit, _ = oracle_lib.intersect("plane", (0, 2, 0), (1, 0, 0), (0, 0, 0), (0, 1, 0))
    assert not hit                                # parallel: fabs(denom) == 0
    hit, _ = oracle_lib.intersect("plane", (0, 2, 0), (0, 1, 0), (0, 0, 0), (0, 1, 0))
    assert not hit                                # backwards
    hit, t = oracle_lib.intersect("plane", (0, 0, 0), (0, -1, 0), (0, 0, 0), (0, 1, 0))
    assert hit and t == 0.0                       # t == 0 is accepted (render.cl:215)
    hit, t = oracle_lib.intersect("plane", (0, 2, 0), (0, -1, 0), (0, 0, 0), (0, 2, 0))
    assert hit and t == 2.0                       # un-normalised normal: same t


def test_triangle(oracle_lib):
    a, b, c = (0, 0, 0), (1, 0, 0), (0, 1, 0)
    hit, t = oracle_lib.intersect("triangle", (0.25, 0.25, 1), (0, 0, -1), a, b, c)
    assert hit and t == 1.0
    hit, t = oracle_lib.intersect("triangle", (0.25, 0.25, -1), (0, 0, 1), a, b, c)
    assert hit and t == 1.0                       # no back-face culling
    assert not oracle_lib.intersect("triangle", (0.75, 0.75, 1), (0, 0, -1), a, b, c)[0]   # u + v > 1
    assert not oracle_lib.intersect("triangle", (-0.1, 0.2, 1), (0, 0, -1), a, b, c)[0]    # u < 0
    assert not oracle_lib.intersect("triangle", (0.25, 0.25, 1), (0, 0, 1), a, b, c)[0]    # t < 0
    assert not oracle_lib.intersect("triangle", (0.25, 0.25, 1), (1, 0, 0), a, b, c)[0]    # a == 0 (parallel)
    assert not oracle_lib.intersect("triangle", (0.25, 0.25, 0), (0, 0, -1), a, b, c)[0]   # t == 0 rejected (:270)
    assert oracle_lib.intersect("triangle", (0, 0, 1), (0, 0, -1), a, b, c)[0]             # vertex: u = v = 0 accepted


def test_aabb(oracle_lib):
    lo, hi = (-1, -1, -1), (1, 1, 1)
    assert oracle_lib.intersect("aabb", (0, 0, 5), (0, 0, -1), lo, hi, (np.inf,))[0]
    assert not oracle_lib.intersect("aabb", (0, 0, 5), (0, 0, 1), lo, hi, (np.inf,))[0]
    assert not oracle_lib.intersect("aabb", (0, 0, 5), (0, 0, -1), lo, hi, (3.5,))[0]      # culled by tmax (:319)
    assert oracle_lib.intersect("aabb", (0, 0, 0), (0, 0, -1), lo, hi, (np.inf,))[0]       # origin inside
    assert not oracle_lib.intersect("aabb", (3, 0, 5), (0, 0, -1), lo, hi, (np.inf,))[0]
