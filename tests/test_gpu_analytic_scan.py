"""GPU parity of the ANALYTIC builds' closest-hit scan (scenes of spheres and planes only, at most 16 shapes: the constant
table scan of the plain schedule and the two-shapes-per-operation scan of the wavefront schedule, render_kernels.cuh
scan_shapes / scan_pairs) against the reference kernel render.cl:293-378 compiled in-tree (oracle/_ref).

The random scenes of test_gpu_parity.py always carry box models, which selects the model-aware builds; these scenes
have none.  Shape lists are drawn in every arrangement the op table distinguishes -- runs of one type of even and odd
length, strictly alternating types, a single shape, the full 16 -- and with exact duplicates, so that ties between the
two halves of one pair, between neighbouring pairs and across types are decided (lowest array index wins, :306 / :356).
Everything is compared bit for bit."""
import numpy as np
import pytest

from conftest import assert_bit_equal
from util import make_tracer, ref_primary_ids

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(sky, oracle_lib):
    if not oracle_lib.ref_available():
        pytest.skip("oracle/_ref library not present")
    return {"sky": sky, "oracle": oracle_lib}


def analytic_scene(seed, pattern, width=128, height=80, duplicates=False):
    """pattern: string of 'S' / 'P' giving the type of every shape in array order."""
    from simple_raytracer_b200 import scenes
    rng = np.random.default_rng(seed)
    mats = [scenes.material(rng.uniform(0.2, 1.0, 3), smoothness=rng.uniform(0, 1), metallic=rng.choice([0, 0.5, 1.0]),
                            specular=rng.choice([0, 0.3, 1.0]), transmittance=rng.choice([0, 0, 0.6, 1.0]),
                            refraction_index=rng.choice([0.8, 1.0, 1.33, 1.5, 2.4]),
                            emission=rng.uniform(0, 1, 3), emission_strength=rng.choice([0, 0, 0, 4.0]))
            for _ in range(6)]
    shapes = []
    for k, kind in enumerate(pattern):
        if duplicates and k >= 2 and rng.random() < 0.5:
            # the same geometry again under another material: an exact tie between two array indices
            src = [i for i in range(k) if pattern[i] == kind]
            if src:
                twin = shapes[int(rng.choice(src))].copy()
                twin["material"] = int(rng.integers(6))
                shapes.append(twin)
                continue
        if kind == "P":
            n = rng.normal(size=3)
            n /= np.linalg.norm(n)
            shapes.append(scenes.plane(int(rng.integers(6)), -n * rng.uniform(2.5, 5.0), n))
        else:
            shapes.append(scenes.sphere(int(rng.integers(6)), rng.uniform(-2.5, 2.5, 3) + (0, 0, -3), rng.uniform(0.3, 1.4)))
    return scenes.Scene(f"analytic{seed}:{pattern}", width, height, 2, 7, 1, scenes._stack(shapes, scenes.SHAPE),
                        np.zeros(0, scenes.TRIANGLE), scenes._stack(mats, scenes.MATERIAL),
                        scenes.camera_matrix((0, 0.2, 3.0), rng.uniform(-0.3, 0.3), rng.uniform(-0.2, 0.2)))


PATTERNS = ["S", "P", "SS", "PP", "SP", "PS", "SSS", "PPP", "PPPSSSS", "SPSPSPSP", "SSPPSSPP", "PSSSSSP", "SSSSSSSSSSSSSSSS",
            "PPPPPPPPPPPPPPPP", "SPPSSSPPPPSSSSSP", "PSPSPSPSPSPSPSPS"]


def check(ctx, sc, seed):
    oracle = ctx["oracle"]
    tr = make_tracer(sc, ctx["sky"])
    for kw in (dict(num_samples=3), dict(num_samples=1, num_bounces=1), dict(num_samples=1, show_normals=True)):
        rd = sc.render_data(seed, **kw)
        tr.clear_canvas()
        tr.accumulate(rd)
        ref, _ = oracle.render(rd, sc.scene_data, sc.shapes, sc.triangles, sc.materials, ctx["sky"], impl="ref")
        assert_bit_equal(ref, tr.read_canvas(), f"{sc.name} {kw}")
    tr.close()


@pytest.mark.parametrize("k", range(len(PATTERNS)))
def test_shape_arrangements_against_the_reference_kernel(ctx, k, schedule):
    check(ctx, analytic_scene(300 + k, PATTERNS[k]), k)


@pytest.mark.parametrize("seed", range(6))
def test_duplicated_shapes_tie_to_the_lower_index(ctx, seed, schedule):
    rng = np.random.default_rng(seed)
    pattern = "".join(rng.choice(["S", "P"], size=int(rng.integers(4, 17)), p=[0.6, 0.4]))
    sc = analytic_scene(400 + seed, pattern, duplicates=True)
    check(ctx, sc, seed)
    # the winner of every primary ray, by array index (a tie resolved to the wrong twin shows up here even when both
    # twins happen to share a material)
    tr = make_tracer(sc, ctx["sky"])
    rd1 = sc.render_data(0, num_samples=1)
    gi, _ = tr.debug_primary(rd1)
    assert np.array_equal(ref_primary_ids(ctx["oracle"], sc, rd1), gi)
    tr.close()


def test_seventeen_shapes_leave_the_constant_table(ctx, schedule):
    """One shape more than the table holds: the general analytic build (global shape arrays) renders the same values."""
    check(ctx, analytic_scene(500, "SPSSPPSSSPPPSSSSP"), 0)


def test_scene_updates_rebuild_the_scan_table(ctx, schedule):
    """Uploading a different scene into the same handle replaces the table the scan reads (srt_upload_scene)."""
    oracle = ctx["oracle"]
    a, b = analytic_scene(600, "PPSSS"), analytic_scene(601, "SPS")
    tr = make_tracer(a, ctx["sky"])
    for sc in (a, b, a):
        tr.scene_data[:] = sc.scene_data
        tr.update_scene(sc.shapes, sc.triangles, sc.materials)
        rd = sc.render_data(1, num_samples=2)
        tr.clear_canvas()
        tr.accumulate(rd)
        ref, _ = oracle.render(rd, sc.scene_data, sc.shapes, sc.triangles, sc.materials, ctx["sky"], impl="ref")
        assert_bit_equal(ref, tr.read_canvas(), sc.name)
    tr.close()
