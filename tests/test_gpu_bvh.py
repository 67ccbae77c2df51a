"""The OPTIONAL BVH mode (srt_set_accel(SRT_ACCEL_BVH)) -- a labelled extension outside the parity-graded path
(reference README.md:41 lists it as a future plan, render.cl:324 brute-forces).  Acceptance is by
tolerance against this repository's own brute-force path (which IS bit-exact with render.cl): primary-hit ids and t
identical on the committed scenes, converged images within RMSE 1/255, and the default mode untouched."""
import numpy as np
import pytest

from conftest import assert_bit_equal
from simple_raytracer_b200 import scenes
from simple_raytracer_b200.records import concat_records
from util import cuda_canvas, make_tracer, random_scene

pytestmark = pytest.mark.gpu


def _both(sc, sky, fn):
    tr = make_tracer(sc, sky)
    brute = fn(tr)
    tr.set_accel("bvh")
    bvh = fn(tr)
    tr.set_accel("none")
    again = fn(tr)
    return brute, bvh, again


@pytest.mark.parametrize("cfg,w,h", [(3, 480, 270), (5, 480, 270), (1, 400, 300)])
def test_primary_hits_identical_to_brute_force(sky, cfg, w, h):
    sc = scenes.CONFIGS[cfg](w, h)
    rd = sc.render_data(0, num_samples=1)
    (bi, bt), (ai, at), (ci, ct) = _both(sc, sky, lambda tr: tr.debug_primary(rd))
    assert np.array_equal(bi, ai), f"{int((bi != ai).sum())} primary-hit shape ids differ under the BVH"
    assert_bit_equal(bt, at, "primary t under the BVH")  # same exact test on the same operands
    assert np.array_equal(bi, ci) and (bi >= 0).any()


@pytest.mark.parametrize("cfg,w,h,ns,launches", [(3, 480, 270, 4, 4), (5, 240, 136, 4, 2)])
def test_converged_image_within_rmse_of_brute_force(sky, cfg, w, h, ns, launches):
    """Same seeds, same RNG streams: the BVH image must sit within RMSE 1/255 of the brute-force image after `average`
    (SURVEY 8c's converged-image gate); in practice almost every canvas float is identical."""
    sc = scenes.CONFIGS[cfg](w, h)
    brute, bvh, again = _both(sc, sky, lambda tr: (cuda_canvas(tr, sc, launches, num_samples=ns), tr.resolve(launches)))
    assert_bit_equal(brute[0], again[0], "the default mode after set_accel('none')")
    a, b = brute[1][..., 1:].astype(np.float64), bvh[1][..., 1:].astype(np.float64)
    rmse = float(np.sqrt(np.mean((a - b) ** 2)))
    same = float((brute[0].view(np.uint32) == bvh[0].view(np.uint32)).all(axis=-1).mean())
    assert rmse <= 1.0, f"RMSE {rmse:.3f} LSB"
    assert same >= 0.99, f"only {same:.4f} of the pixels have identical canvas floats"


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_soups_match_brute_force(sky, seed):
    """Random triangle soups (long slivers, intersecting triangles) + boxes (<= 32 triangles: brute force inline even
    in BVH mode) in random shape order: primary hits and the show_normals image (which triangle won) identical."""
    sc = random_scene(seed, 160, 96, mesh_tris=700)
    rd = sc.render_data(0, num_samples=1)
    (bi, bt), (ai, at), _ = _both(sc, sky, lambda tr: tr.debug_primary(rd))
    assert np.array_equal(bi, ai)
    assert_bit_equal(bt, at, "t")
    rdn = sc.render_data(0, num_samples=1, show_normals=True)

    def normals(tr):
        tr.clear_canvas()
        tr.accumulate(rdn)
        return tr.read_canvas()
    bn, an, _ = _both(sc, sky, normals)
    assert_bit_equal(bn, an, "show_normals image under the BVH")


def test_equal_t_goes_to_the_lowest_triangle_index(sky):
    """Every triangle of the mesh is present TWICE (second copy with flipped vertex normals).  Both copies give the same
    t bit for bit; the reference loop keeps the first (strict `<`, render.cl:332).  The BVH must too, wherever its
    leaves put the two copies: the show_normals image tells which copy won."""
    v, n, f = scenes.noisy_icosphere(2, seed=9)
    a = scenes.mesh_triangles(v, f, n)
    b = scenes.mesh_triangles(v, f, -n)
    for tris in (concat_records(scenes.TRIANGLE, a, b), concat_records(scenes.TRIANGLE, b, a)):
        shapes = scenes._stack([scenes.plane(0, (0, -1.6, 0), (0, 1, 0)),
                                scenes.model(0, tris, 0, len(tris), scenes.translate((0, 0, -1.0)))], scenes.SHAPE)
        sc = scenes.Scene("dup", 200, 150, 1, 4, 1, shapes, tris, scenes._stack([scenes.material((0.8, 0.8, 0.8))], scenes.MATERIAL),
                          scenes.camera_matrix((0, 0.2, 3.0)))
        rdn = sc.render_data(0, show_normals=True)

        def normals(tr):
            tr.clear_canvas()
            tr.accumulate(rdn)
            return tr.read_canvas()
        bn, an, _ = _both(sc, sky, normals)
        assert_bit_equal(bn, an, "which of two coincident triangles wins")


def test_reupload_rebuilds_and_counters_show_the_saving(sky):
    sc = scenes.config5(320, 180)
    tr = make_tracer(sc, sky)
    rd = sc.render_data(0, num_samples=1)
    brute = tr.accumulate_counted(rd)
    tr.set_accel("bvh")
    bvh = tr.accumulate_counted(rd)
    assert int(bvh[0]["samples"]) == int(brute[0]["samples"]) and int(bvh[0]["aabb_pass"]) == int(brute[0]["aabb_pass"])
    assert int(bvh[0]["tri_tests"]) * 200 < int(brute[0]["tri_tests"])  # 100 352 per ray -> a few dozen
    ids_a, _ = tr.debug_primary(rd)
    # a different scene through the same handle: the hierarchy follows the upload
    sc2 = scenes.config3(320, 180)
    tr.scene_data[:] = sc2.scene_data
    tr.update_scene(sc2.shapes, sc2.triangles, sc2.materials)
    ids_b, t_b = tr.debug_primary(sc2.render_data(0, num_samples=1))
    tr.set_accel("none")
    ids_c, t_c = tr.debug_primary(sc2.render_data(0, num_samples=1))
    assert np.array_equal(ids_b, ids_c) and not np.array_equal(ids_a, ids_b)
    assert_bit_equal(t_b, t_c, "t after re-upload")
