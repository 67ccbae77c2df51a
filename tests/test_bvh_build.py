"""The host-side BVH builder of the optional acceleration mode (csrc/bvh_build.hpp), checked on the CPU through a small
C++ harness: every triangle of the model lands in exactly one leaf, leaves hold <= 4 triangles in ascending index
order, boxes enclose their triangles and nest, indices are absolute, depth stays within the kernel's stack."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "bvh_check")


@pytest.fixture(scope="module")
def exe():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-ffp-contract=off",
                           "-I" + os.path.join(ROOT, "simple_raytracer_b200", "csrc"),
                           os.path.join(ROOT, "tests", "cpp", "bvh_check.cpp"), "-o", EXE])
    return EXE


@pytest.mark.parametrize("n,seed,mode", [(33, 1, 0), (1000, 2, 0), (100352, 3, 0), (5000, 4, 1), (20000, 5, 2), (37, 6, 2)])
def test_builder_invariants(exe, n, seed, mode):
    out = subprocess.run([exe, str(n), str(seed), str(mode)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.startswith("ok"), out.stdout[-2000:] + out.stderr[-2000:]
    _, nodes, slots, depth = out.stdout.split()
    assert int(slots) == n and int(depth) <= 48 and int(nodes) >= n // 8
