"""Full-size, size-independent properties of the CUDA path (BASELINE.json sizes) and the reference-facing
Tracer protocol and error behaviour, all through the C ABI."""
import numpy as np
import pytest

from conftest import assert_bit_equal
from simple_raytracer_b200 import scenes
from simple_raytracer_b200.tracer import SrtError, Tracer
from util import cuda_canvas, make_tracer, psnr

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["one", "two"])
def sweep_filter(request, monkeypatch):
    """Run a mesh test under BOTH conservative sweep filters (srt_set_sweep_filter through the Tracer's
    SRT_SWEEP_FILTER hook): the one-strip filter big models get and the two-strip filter small models get must each
    reproduce the reference bit for bit on every adversarial mesh, whatever the size-based default would pick."""
    monkeypatch.setenv("SRT_SWEEP_FILTER", request.param)
    return request.param


def test_row_bands_partition_is_bit_identical_at_1080p(sky):
    """Tile sharding invariant at config-2 size: the union of 8 interleaved band renders == the full frame."""
    sc = scenes.config2()
    tr = make_tracer(sc, sky)
    rd = sc.render_data(0)
    tr.accumulate(rd)
    full = tr.read_canvas()
    tr.clear_canvas()
    for b in range(8):
        tr.set_row_bands(8, b, 8)
        tr.accumulate(rd)
    tr.set_row_bands(1, 0, 1)
    assert_bit_equal(full, tr.read_canvas(), "bands")
    # a single band touches only its rows
    tr.clear_canvas()
    tr.set_row_bands(8, 3, 8)
    tr.accumulate(rd)
    tr.set_row_bands(1, 0, 1)
    part = tr.read_canvas()
    rows = (np.arange(sc.height) // 8) % 8 == 3
    assert not part[~rows].any() and np.array_equal(part[rows], full[rows])


def test_accumulation_linearity_and_determinism_full_size(sky):
    sc = scenes.config2()
    tr = make_tracer(sc, sky)
    singles = []
    for k in range(3):
        tr.clear_canvas()
        tr.accumulate(sc.render_data(k))
        singles.append(tr.read_canvas())
    acc = cuda_canvas(tr, sc, 3)
    assert_bit_equal(acc, (singles[0] + singles[1]) + singles[2], "canvas += mean, launch by launch")
    assert_bit_equal(acc, cuda_canvas(tr, sc, 3), "run-to-run determinism")
    # log(0) / 0-length vectors are legal in the reference arithmetic (probability ~2^-32 per draw)
    assert (~np.isfinite(acc)).sum() <= 12


def test_mesh_config_full_size_against_oracle_crop(sky, oracle_lib, sweep_filter):
    """Config 3 at 1080p: the CUDA frame equals the oracle on a centred 240x136 window of the same
    full-size launch (global pixel ids, seeds and aspect preserved)."""
    sc = scenes.config3()
    tr = make_tracer(sc, sky)
    rd = sc.render_data(0, num_samples=2)
    tr.accumulate(rd)
    got = tr.read_canvas()
    x0, y0 = (sc.width - 240) // 2, (sc.height - 136) // 2
    want, _ = oracle_lib.render(rd, sc.scene_data, sc.shapes, sc.triangles, sc.materials, sky,
                                window=(x0, y0, x0 + 240, y0 + 136))
    assert_bit_equal(want[y0:y0 + 136, x0:x0 + 240], got[y0:y0 + 136, x0:x0 + 240], "C3 crop")
    assert (want[y0:y0 + 136, x0:x0 + 240, :3] > 0).any()


def test_stress_mesh_primary_and_crop(sky, oracle_lib, sweep_filter):
    """Config 5 (100 352 triangles): primary ids / t and 1-spp radiance on a crop against the oracle."""
    sc = scenes.config5(480, 270)
    tr = make_tracer(sc, sky)
    rd = sc.render_data(0, num_samples=1)
    x0, y0, w, h = 200, 110, 80, 48
    tr.accumulate(rd)
    got = tr.read_canvas()
    want, _ = oracle_lib.render(rd, sc.scene_data, sc.shapes, sc.triangles, sc.materials, sky,
                                window=(x0, y0, x0 + w, y0 + h))
    assert_bit_equal(want[y0:y0 + h, x0:x0 + w], got[y0:y0 + h, x0:x0 + w], "C5 crop")
    gi, gt = tr.debug_primary(rd)
    assert (gi == 1).sum() > 1000  # the mesh is in view


def test_converged_image_psnr(sky, oracle_lib):
    """Converged-image gate (RMSE <= 1/255, PSNR >= 45 dB) -- trivially met because the canvases are equal,
    but asserted on the resolved images as the contract states it."""
    sc = scenes.config2(320, 180)
    tr = make_tracer(sc, sky)
    canvas = None
    tr.clear_canvas()
    for k in range(8):
        rd = sc.render_data(k)
        tr.accumulate(rd)
        canvas, _ = oracle_lib.render(rd, sc.scene_data, sc.shapes, sc.triangles, sc.materials, sky, canvas)
    a, b = tr.resolve(8), oracle_lib.average(8, canvas)
    rmse = np.sqrt(np.mean((a.astype(float) - b.astype(float)) ** 2)) / 255.0
    assert rmse <= 1.0 / 255.0 and psnr(a, b) >= 45.0


def test_tracer_frame_protocol(sky, oracle_lib):
    """main.cpp:277-290: clear + update_scene, then render(ticks_stopped, pixels) frame after frame."""
    sc = scenes.config1(200, 150)
    tr = Tracer(200, 150, sky)
    assert int(tr.options["num_samples"][0]) == 4 and int(tr.options["num_bounces"][0]) == 10  # tracer.hpp:61-66
    tr.scene_data[:] = sc.scene_data
    tr.clear_canvas()
    tr.update_scene(sc.shapes, sc.triangles, sc.materials)
    pixels = np.zeros(200 * 150 * 4, np.uint8)
    canvas = None
    for tick in range(1, 4):
        tr.options[:] = sc.render_data(tick, num_samples=2)
        tr.render(tick, pixels)
        canvas, _ = oracle_lib.render(tr.options, sc.scene_data, sc.shapes, sc.triangles, sc.materials, sky, canvas)
        assert np.array_equal(pixels.reshape(150, 200, 4), oracle_lib.average(tick, canvas))
    assert (pixels.reshape(-1, 4)[:, 0] == 255).all()
    # scene arrays may be modified right after update_scene (copy-in semantics)
    shapes = sc.shapes.copy()
    tr.update_scene(shapes, sc.triangles, sc.materials)
    shapes["material"] = 0
    tr.clear_canvas()
    tr.options[:] = sc.render_data(1, num_samples=2)
    tr.render(1, pixels)
    c1, _ = oracle_lib.render(tr.options, sc.scene_data, sc.shapes, sc.triangles, sc.materials, sky)
    assert np.array_equal(pixels.reshape(150, 200, 4), oracle_lib.average(1, c1))


def test_checkpoint_resume_canvas(sky):
    sc = scenes.config2(160, 90)
    tr = make_tracer(sc, sky)
    a = cuda_canvas(tr, sc, 2)
    tr.accumulate(sc.render_data(2))
    want = tr.read_canvas()
    tr2 = make_tracer(sc, sky)
    tr2.write_canvas(a)
    tr2.accumulate(sc.render_data(2))
    assert_bit_equal(want, tr2.read_canvas(), "resume")


def test_errors(sky):
    sc = scenes.config1(64, 48)
    tr = Tracer(64, 48, sky)
    with pytest.raises(SrtError, match="no scene"):
        tr.accumulate(sc.render_data(0))
    bad = sc.shapes.copy()
    bad["material"][2] = 99
    with pytest.raises(SrtError, match="material 99 out of range"):
        tr.update_scene(bad, sc.triangles, sc.materials)
    bad = sc.shapes.copy()
    bad["model_num_triangles"][6] = 13
    with pytest.raises(SrtError, match="exceed"):
        tr.update_scene(bad, sc.triangles, sc.materials)
    tr.update_scene(sc.shapes, sc.triangles, sc.materials)
    with pytest.raises(SrtError, match="created 64x48"):
        tr.accumulate(sc.render_data(0, width=32, height=32))
    with pytest.raises(ValueError):
        tr.render(1, np.zeros(10, np.uint8))
    # an empty scene renders pure sky (num_shapes = 0)
    tr.scene_data[:] = sc.scene_data
    tr.update_scene(sc.shapes[:0], sc.triangles[:0], sc.materials)
    tr.clear_canvas()
    tr.accumulate(sc.render_data(0))
    assert (tr.read_canvas()[..., :3] > 0).all()


def test_ragged_sizes_and_sample_counts(sky, oracle_lib):
    """Odd image sizes (not a multiple of the warp or the block), num_samples from 1 to 7, 1 bounce."""
    for (w, h, ns, nb) in [(33, 17, 1, 1), (61, 7, 7, 3), (1, 1, 5, 4), (130, 3, 3, 32)]:
        sc = scenes.config1(w, h)
        rd = sc.render_data(0, num_samples=ns, num_bounces=nb)
        tr = make_tracer(sc, sky)
        tr.accumulate(rd)
        want, _ = oracle_lib.render(rd, sc.scene_data, sc.shapes, sc.triangles, sc.materials, sky)
        assert_bit_equal(want, tr.read_canvas(), f"{w}x{h} ns={ns} nb={nb}")


def test_shared_triangles_between_instances_and_empty_model(sky, oracle_lib, sweep_filter):
    """Several box models share the 12 cube triangles (shape.cpp:76-89); a model with zero triangles is legal."""
    tris = scenes.cube_triangles()
    mats = np.zeros(3, scenes.MATERIAL)
    mats[0] = scenes.material((0.8, 0.8, 0.8))
    mats[1] = scenes.material((0.9, 0.3, 0.2), smoothness=0.7, metallic=0.5)
    mats[2] = scenes.material((1, 1, 1), smoothness=1.0, transmittance=1.0, refraction_index=1.4)
    shapes = np.zeros(6, scenes.SHAPE)
    shapes[0] = scenes.plane(0, (0, -1, 0), (0, 1, 0))
    shapes[1] = scenes.model(1, tris, 0, 12, scenes.translate((-1.5, 0, -1)))
    shapes[2] = scenes.model(2, tris, 0, 12, scenes.translate((1.5, 0.2, -2)) @ scenes.rotate_y(0.5) @ scenes.scale(0.8))
    shapes[3] = scenes.model(0, tris, 0, 1, None)
    shapes[3]["model_num_triangles"] = 0
    shapes[3]["model_bounding_min"], shapes[3]["model_bounding_max"] = (-9, -9, -9), (9, 9, 9)
    shapes[4] = scenes.sphere(1, (0, 0, -3), 1.0)
    shapes[5] = scenes.model(1, tris, 6, 6, scenes.translate((0, 2.5, -2)))  # sub-range of the cube
    sc = scenes.Scene("instances", 160, 120, 2, 6, 1, shapes, tris, mats, scenes.camera_matrix((0, 0.5, 4)))
    tr = make_tracer(sc, sky)
    rd = sc.render_data(0)
    tr.accumulate(rd)
    want, _ = oracle_lib.render(rd, sc.scene_data, sc.shapes, sc.triangles, sc.materials, sky)
    assert_bit_equal(want, tr.read_canvas(), "instances")
    oi, ot = oracle_lib.primary(rd, sc.scene_data, sc.shapes, sc.triangles)
    gi, gt = tr.debug_primary(rd)
    assert np.array_equal(oi, gi) and set(np.unique(gi)) >= {0, 1, 2, 4}
    assert_bit_equal(ot, gt, "t")


def _random_mesh(rng, n_tris, center, extent):
    """n_tris random (some degenerate, some huge, some sharing vertices/edges) triangles around `center`."""
    base = rng.normal(size=(n_tris, 1, 3)) * extent * 0.6 + np.asarray(center)
    pos = (base + rng.normal(size=(n_tris, 3, 3)) * extent * 0.35).astype(np.float32)
    # shared edges / coplanar neighbours: triangle 2k+1 reuses two vertices of triangle 2k
    pos[1::2, 0] = pos[0:n_tris - (n_tris % 2):2, 1][:len(pos[1::2])]
    pos[1::2, 1] = pos[0:n_tris - (n_tris % 2):2, 2][:len(pos[1::2])]
    if n_tris > 4:
        pos[3, 2] = pos[3, 1]            # zero-area triangle
        pos[4] = pos[2]                  # exact duplicate: equal t, the lower index must win
    nrm = rng.normal(size=(n_tris, 3, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    return scenes.triangles_from(pos, nrm)


@pytest.mark.parametrize("sizes", [(33,), (127, 128, 129), (32, 257, 31, 1000, 64), (513, 40, 40, 40, 700, 33)])
def test_models_of_awkward_sizes_tile_boundaries_and_many_instances(sky, oracle_lib, sizes, sweep_filter):
    """Triangle-phase bookkeeping: tile tails (n mod 128), the inline/parked threshold (32 / 33 triangles),
    more distinct big models than a warp sweeps at once, duplicate and degenerate triangles, instances."""
    rng = np.random.default_rng(sum(sizes))
    tri_list, shapes = [], [scenes.plane(0, (0, -2.5, 0), (0, 1, 0)), scenes.sphere(1, (0, -1.0, -2.0), 0.8)]
    first = 0
    for i, n in enumerate(sizes):
        ang = 2 * np.pi * i / len(sizes)
        t = _random_mesh(rng, n, (2.2 * np.cos(ang), 0.4 * np.sin(3 * ang), -3.0 + 1.5 * np.sin(ang)), 1.0)
        tri_list.append(t)
        first += n
    tris = scenes.concat_records(scenes.TRIANGLE, *tri_list) if hasattr(scenes, "concat_records") else None
    if tris is None:
        from simple_raytracer_b200.records import concat_records
        tris = concat_records(scenes.TRIANGLE, *tri_list)
    first = 0
    for i, n in enumerate(sizes):
        shapes.append(scenes.model(1 + i % 3, tris, first, n, scenes.rotate_y(0.1 * i)))
        first += n
    # a second instance of the first mesh, overlapping the others
    shapes.append(scenes.model(2, tris, 0, sizes[0], scenes.translate((0.3, 0.2, 0.4)) @ scenes.scale(1.3)))
    mats = scenes._stack([scenes.material((0.8, 0.8, 0.8)),
                          scenes.material((0.9, 0.5, 0.3), smoothness=0.8, metallic=0.6),
                          scenes.material((1, 1, 1), smoothness=1.0, transmittance=0.9, refraction_index=1.4),
                          scenes.material((0.3, 0.6, 0.9), specular=0.4, smoothness=0.5)], scenes.MATERIAL)
    sc = scenes.Scene("awkward", 192, 108, 2, 6, 2, scenes._stack(shapes, scenes.SHAPE), tris, mats,
                      scenes.camera_matrix((0, 0.2, 3.5)))
    tr = make_tracer(sc, sky)
    rd = sc.render_data(0)
    oi, ot = oracle_lib.primary(rd, sc.scene_data, sc.shapes, sc.triangles)
    gi, gt = tr.debug_primary(rd)
    assert np.array_equal(oi, gi)
    assert_bit_equal(ot, gt, "t")
    want, wcnt = None, None
    tr.clear_canvas()
    gcnt = None
    for k in range(2):
        rdk = sc.render_data(k)
        want, c = oracle_lib.render(rdk, sc.scene_data, sc.shapes, sc.triangles, sc.materials, sky, want)
        wcnt = [int(v) for v in c] if wcnt is None else [a + int(b) for a, b in zip(wcnt, c)]
        gcnt = tr.accumulate_counted(rdk, gcnt)
    assert_bit_equal(want, tr.read_canvas(), f"sizes {sizes} (counted kernel)")
    assert wcnt == [int(gcnt[0][n]) for n in gcnt.dtype.names]
    got = cuda_canvas(tr, sc, 2)
    assert_bit_equal(want, got, f"sizes {sizes}")
    assert len(set(np.unique(gi))) >= min(len(sizes), 3)


def test_rays_through_shared_edges_and_vertices(sky, oracle_lib, sweep_filter):
    """Adversarial for the triangle filter: a regular grid of rays (time = 0 gives every pixel the same jitter)
    against an axis-aligned triangle grid, so many rays pass within rounding distance of shared edges and
    vertices (u, v near 0 or 1), plus the same grid seen almost edge-on (det near 0, grazing rays)."""
    n = 24
    xs = np.linspace(-1.5, 1.5, n + 1, dtype=np.float32)
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    P = np.stack([X, Y, (-3.0 + 0.02 * np.sin(7 * X) * np.cos(5 * Y)).astype(np.float32)], -1)  # a flat AABB never passes (:289)
    a, b, c, d = P[:-1, :-1], P[1:, :-1], P[:-1, 1:], P[1:, 1:]
    pos = np.concatenate([np.stack([a, b, c], -2).reshape(-1, 3, 3), np.stack([b, d, c], -2).reshape(-1, 3, 3)])
    nrm = np.tile(np.array([0, 0, 1], np.float32), (len(pos), 3, 1))
    tris = scenes.triangles_from(pos, nrm)
    edge_on = scenes.translate((0, 1e-4, -1.0)) @ scenes.rotate_x(np.pi / 2 - 1e-5) @ scenes.translate((0, 0, 3.0))
    shapes = scenes._stack([scenes.model(0, tris, 0, len(tris), None),
                            scenes.model(0, tris, 0, len(tris), edge_on)], scenes.SHAPE)
    mats = scenes._stack([scenes.material((0.7, 0.7, 0.7), smoothness=0.3, specular=0.2)], scenes.MATERIAL)
    sc = scenes.Scene("grid", 193, 193, 1, 3, 1, shapes, tris, mats, scenes.camera_matrix((0, 0, 0)))
    tr = make_tracer(sc, sky)
    rd = sc.render_data(0)
    rd["time"] = 0  # time = 0 zeroes every seed: u0 = u1 = hash(0) for all pixels -> a regular grid of rays
    oi, ot = oracle_lib.primary(rd, sc.scene_data, sc.shapes, sc.triangles)
    gi, gt = tr.debug_primary(rd)
    assert np.array_equal(oi, gi) and (gi == 0).sum() > 2000 and (gi == 1).sum() > 2000
    assert_bit_equal(ot, gt, "t")
    tr.accumulate(rd)
    want, _ = oracle_lib.render(rd, sc.scene_data, sc.shapes, sc.triangles, sc.materials, sky)
    assert_bit_equal(want, tr.read_canvas(), "grid")


def _mesh_scene(tris, xf, cam, name, mats=None, extra_shapes=(), w=160, h=96, ns=2, nb=5):
    shapes = list(extra_shapes) + [scenes.model(1, tris, 0, len(tris), xf)]
    mats = mats or [scenes.material((0.8, 0.8, 0.8)),
                    scenes.material((0.9, 0.6, 0.3), smoothness=0.7, metallic=0.4, specular=0.2)]
    return scenes.Scene(name, w, h, ns, nb, 1, scenes._stack(shapes, scenes.SHAPE), tris,
                        scenes._stack(mats, scenes.MATERIAL), cam)


def _check_against_oracle(sc, sky, oracle_lib, what, min_mesh_pixels=200, mesh_shape=None):
    tr = make_tracer(sc, sky)
    rd = sc.render_data(0)
    oi, ot = oracle_lib.primary(rd, sc.scene_data, sc.shapes, sc.triangles)
    gi, gt = tr.debug_primary(rd)
    assert np.array_equal(oi, gi), f"{what}: {int((oi != gi).sum())} primary ids differ"
    assert_bit_equal(ot, gt, f"{what}: t")
    mesh_shape = len(sc.shapes) - 1 if mesh_shape is None else mesh_shape
    assert (gi == mesh_shape).sum() >= min_mesh_pixels, f"{what}: the mesh is not in view"
    want, wc = oracle_lib.render(rd, sc.scene_data, sc.shapes, sc.triangles, sc.materials, sky)
    cnt = tr.accumulate_counted(rd)
    assert_bit_equal(want, tr.read_canvas(), what)
    assert [int(v) for v in wc] == [int(cnt[0][n]) for n in cnt.dtype.names]


@pytest.mark.parametrize("offset", [(0, 0, 0), (1000.0, -2000.0, 512.0), (-3.0e5, 1.0e5, 2.0e5)])
def test_sweep_filter_margins_far_from_the_origin(sky, oracle_lib, offset, sweep_filter):
    """The sweep filter works on pre-multiplied operands (o x d, e2 x v0) whose rounding error grows with the
    distance from the origin; its margins must grow with it.  The same mesh and camera translated far away: the
    reference's own arithmetic degrades there (o - v0 loses bits), and the CUDA path must degrade identically."""
    v, n, f = scenes.noisy_icosphere(3, seed=9)
    tris = scenes.mesh_triangles(v, f, n)
    off = np.asarray(offset, np.float32)
    sc = _mesh_scene(tris, scenes.translate(off + (0, 0, -3)) @ scenes.scale(1.2),
                     scenes.camera_matrix(off + (0, 0.2, 1.5)), f"far{offset}",
                     extra_shapes=[scenes.plane(0, off + (0, -1.5, 0), (0, 1, 0))])
    _check_against_oracle(sc, sky, oracle_lib, f"offset {offset}")


@pytest.mark.parametrize("size", [1e-4, 1e-2, 1e3])
def test_sweep_filter_tiny_and_huge_triangles(sky, oracle_lib, size, sweep_filter):
    """Mesh scale from 1e-4 to 1e3 scene units (the absolute 2e-6 slack of the filter dominates at the small end,
    the relative terms at the large end); the camera distance scales along."""
    v, n, f = scenes.noisy_icosphere(2, seed=4)
    tris = scenes.mesh_triangles((v * size).astype(np.float32), f, n)
    sc = _mesh_scene(tris, scenes.translate((0, 0, -3 * size)), scenes.camera_matrix((0, 0, 0.5 * size)), f"size{size}")
    _check_against_oracle(sc, sky, oracle_lib, f"size {size}")


def test_sweep_filter_undecidable_triangles_overflow_the_pair_ring(sky, oracle_lib, sweep_filter):
    """Hundreds of zero-area triangles (all three vertices equal, or collinear): det = 0, so the filter can decide
    nothing and every one of them survives -- far more survivors per tile than the pair ring holds -- while the
    reference rejects them all (render.cl:253).  Real triangles are interleaved so that hits still have to be
    found in order."""
    rng = np.random.default_rng(5)
    v, n, f = scenes.noisy_icosphere(2, seed=6)
    real = scenes.mesh_triangles(v, f, n)
    junk = np.zeros(900, scenes.TRIANGLE)
    p = rng.normal(size=(900, 1, 3)).astype(np.float32) * 0.5
    junk["v"]["pos"] = np.repeat(p, 3, axis=1)
    junk["v"]["pos"][::3, 1] += (0.3, 0, 0)          # collinear: v0, v0 + a, v0 + 2a
    junk["v"]["pos"][::3, 2] += (0.6, 0, 0)
    junk["v"]["normal"] = (0, 0, 1)
    from simple_raytracer_b200.records import concat_records
    order = rng.permutation(len(real) + len(junk))
    tris = concat_records(scenes.TRIANGLE, real, junk)[order]
    sc = _mesh_scene(tris, scenes.translate((0, 0, -3)), scenes.camera_matrix((0, 0, 0.3)), "junk")
    _check_against_oracle(sc, sky, oracle_lib, "junk triangles")


def test_sweep_filter_nan_and_inf_operands(sky, oracle_lib, sweep_filter):
    """A model containing triangles with NaN / inf coordinates, and an AABB that lets every ray in: the filter must
    pass what it cannot decide, and the exact test then behaves as the reference does."""
    v, n, f = scenes.noisy_icosphere(2, seed=8)
    tris = scenes.mesh_triangles(v, f, n)
    tris["v"]["pos"][5, 1, 0] = np.nan
    tris["v"]["pos"][17, 2] = (np.inf, 0, 0)
    tris["v"]["pos"][40, 0, 1] = -np.inf
    tris["v"]["pos"][41] = 3e38
    sc = _mesh_scene(tris, scenes.translate((0, 0, -3)), scenes.camera_matrix((0, 0, 0.3)), "nan")
    sc.shapes["model_bounding_min"][-1], sc.shapes["model_bounding_max"][-1] = (-50, -50, -50), (50, 50, 50)
    _check_against_oracle(sc, sky, oracle_lib, "nan/inf triangles", min_mesh_pixels=50)


@pytest.mark.skipif(not __import__("oracle").ref_available(), reason="oracle/_ref library not present")
@pytest.mark.parametrize("cfg,launches,ns", [(1, 1, 1), (2, 2, 4), (3, 1, 1)])
def test_full_baseline_size_against_the_reference_kernel(sky, oracle_lib, cfg, launches, ns):
    """BASELINE.json's own sizes (800x600, 1920x1080) against render.cl itself (oracle/_ref on the host cores):
    every float of the full-frame canvas and every byte of the resolved image."""
    sc = scenes.CONFIGS[cfg]()
    tr = make_tracer(sc, sky)
    got = cuda_canvas(tr, sc, launches, num_samples=ns)
    ref = None
    for k in range(launches):
        ref, _ = oracle_lib.render(sc.render_data(k, num_samples=ns), sc.scene_data, sc.shapes, sc.triangles,
                                   sc.materials, sky, ref, impl="ref")
    assert_bit_equal(ref, got, f"C{cfg} full size vs render.cl")
    assert np.array_equal(oracle_lib.average(launches, ref, impl="ref"), tr.resolve(launches))


@pytest.mark.parametrize("cfg,w,h,ns,n", [(1, 160, 120, 2, 5), (2, 240, 136, 4, 19), (3, 120, 68, 3, 4), (5, 48, 27, 1, 3)])
def test_batched_launches_equal_separate_launches(sky, cfg, w, h, ns, n):
    """srt_render_batch: n launches through one persistent kernel (item space launch x pixel x sample) leave the same
    canvas, bit for bit, as n srt_render calls -- also beyond MAX_BATCH (19 > 16), under row bands, and when the
    batch mixes cameras / sample counts so that it has to be split into runs."""
    sc = scenes.CONFIGS[cfg](w, h)
    tr = make_tracer(sc, sky)
    rds = [sc.render_data(k, num_samples=ns) for k in range(n)]
    want = cuda_canvas(tr, sc, n, num_samples=ns)
    tr.clear_canvas()
    tr.accumulate_batch(rds)
    assert_bit_equal(want, tr.read_canvas(), f"C{cfg} batch of {n}")
    # row bands
    tr.set_row_bands(2, 1, 3)
    tr.clear_canvas()
    for rd in rds:
        tr.accumulate(rd)
    part = tr.read_canvas()
    tr.clear_canvas()
    tr.accumulate_batch(rds)
    assert_bit_equal(part, tr.read_canvas(), f"C{cfg} batch under row bands")
    tr.set_row_bands(1, 0, 1)
    # a batch that has to be split: camera moved after two launches, sample count changed for the last one
    mixed = [sc.render_data(k, num_samples=ns) for k in range(min(n, 4))]
    mixed[2]["camera_to_world"][0, 3, 0] += 0.25
    mixed[-1]["num_samples"] = ns + 1
    tr.clear_canvas()
    for rd in mixed:
        tr.accumulate(rd)
    want = tr.read_canvas()
    tr.clear_canvas()
    tr.accumulate_batch(mixed)
    assert_bit_equal(want, tr.read_canvas(), f"C{cfg} mixed batch")
    ms, launches = tr.render_time_ms()
    assert launches > 0 and ms > 0


@pytest.mark.parametrize("key", ["config1", "config2", "config3", "config4", "config5_480x270"])
def test_full_size_canvas_digest_of_the_reference_kernel(sky, key):
    """Full-size canvases and resolved images against SHA-256 digests of what the reference kernel produced for the
    same seeded scene (tests/golden/fullsize_hashes.json): BASELINE sizes incl. 4K, and the 100 352-triangle mesh."""
    import hashlib
    import fullsize_util
    h = fullsize_util.load()
    e = h[key]
    if hashlib.sha256(sky.tobytes()).hexdigest() != h["sky_sha256"]:
        pytest.skip("procedural sky differs on this platform")
    sc, same = fullsize_util.scene_for(e)
    if not same:
        pytest.skip("scene builder produced different bytes on this platform")
    tr = make_tracer(sc, sky)
    canvas = cuda_canvas(tr, sc, e["launches"], num_samples=e["num_samples"])
    assert fullsize_util.canvas_digest(canvas) == e["canvas_sha256"]
    assert hashlib.sha256(tr.resolve(e["launches"]).tobytes()).hexdigest() == e["argb_sha256"]


def test_many_launches_recycle_timing_events_and_stay_deterministic(sky):
    """5 000 launches without ever reading the timings (the library recycles its CUDA events every 4 096 launches),
    interleaved with batches, clears and resolves: no error, and the same canvas as a fresh tracer gives."""
    sc = scenes.config1(64, 48)
    tr = make_tracer(sc, sky)
    rds = [sc.render_data(k, num_samples=1, num_bounces=3) for k in range(8)]
    for i in range(600):
        tr.accumulate_batch(rds)
        if i % 50 == 0:
            tr.resolve(i + 1)
        if i % 97 == 0:
            tr.clear_canvas()
    for i in range(300):
        tr.accumulate(rds[i % 8])
    tr.clear_canvas()
    tr.accumulate_batch(rds)
    got = tr.read_canvas()
    fresh = make_tracer(sc, sky)
    fresh.accumulate_batch(rds)
    assert_bit_equal(fresh.read_canvas(), got, "after 5 100 launches")
    ms, n = tr.render_time_ms()
    assert n > 0 and ms > 0


def _tie_scene(order):
    """1x1 image, one ray (pixel 0 / sample 0 has seed 0 whatever `time` is).  A sphere and a plane whose offset is
    searched (on the CPU, with the oracle) until both are hit at the SAME t bit for bit."""
    import oracle
    mats = scenes._stack([scenes.material((0.9, 0.2, 0.2), emission=(1, 0, 0), emission_strength=2.0),
                          scenes.material((0.2, 0.9, 0.2), emission=(0, 1, 0), emission_strength=3.0)], scenes.MATERIAL)
    cam = scenes.camera_matrix((0.1, 0.2, 5.0))
    sph = scenes.sphere(0, (-3.5, 3.0, 1.1), 1.25)  # on the path of the single ray (jitter 0.0302, 0.1356)

    def scene_of(shapes):
        return scenes.Scene("tie", 1, 1, 1, 1, 1, scenes._stack(shapes, scenes.SHAPE), np.zeros(0, scenes.TRIANGLE), mats, cam)

    def t_of(shape):
        sc = scene_of([shape])
        _, t = oracle.primary(sc.render_data(0), sc.scene_data, sc.shapes, sc.triangles)
        return np.float32(t[0, 0])

    ts = t_of(sph)
    assert np.isfinite(ts)
    normal = (0.0, 0.6, 0.8)
    lo, hi = np.float32(-10.0), np.float32(10.0)  # plane offset along z: t decreases as the plane moves towards the camera
    for _ in range(64):
        mid = np.float32((np.float64(lo) + np.float64(hi)) / 2)
        if t_of(scenes.plane(1, (0, 0, mid), normal)) > ts:
            lo = mid
        else:
            hi = mid
    z = hi
    for _ in range(64):  # walk the neighbouring floats for an exact hit
        tp = t_of(scenes.plane(1, (0, 0, z), normal))
        if tp == ts:
            break
        z = np.nextafter(z, np.float32(10.0 if tp > ts else -10.0), dtype=np.float32)
    else:
        pytest.skip("no plane offset gives a bit-exact tie for this ray")
    pl = scenes.plane(1, (0, 0, z), normal)
    return scene_of([sph, pl] if order == "sphere_first" else [pl, sph])


@pytest.mark.parametrize("order", ["sphere_first", "plane_first"])
def test_exact_tie_between_a_sphere_and_a_plane_goes_to_the_lower_array_index(sky, oracle_lib, order):
    """A bit-exact tie between two shapes of different type goes to the one that comes first in the array
    (render.cl:306,:356: strict `<`) -- whichever order an implementation visits them in.  (A variant of the kernel
    that scanned spheres and planes as two dense lists passed this test and was 2 % slower; it was not kept.)"""
    sc = _tie_scene(order)
    rd = sc.render_data(0)
    oi, ot = oracle_lib.primary(rd, sc.scene_data, sc.shapes, sc.triangles)
    assert oi[0, 0] == 0  # the reference semantics: first in the array wins the tie
    tr = make_tracer(sc, sky)
    tr.accumulate(rd)
    want, _ = oracle_lib.render(rd, sc.scene_data, sc.shapes, sc.triangles, sc.materials, sky)
    got = tr.read_canvas()
    assert_bit_equal(want, got, order)
    # the emission colour tells which shape won: red = sphere material 0, green = plane material 1
    assert (got[0, 0, 0] > 0) == (order == "sphere_first") and (got[0, 0, 1] > 0) == (order == "plane_first")
    if __import__("oracle").ref_available():
        ref, _ = oracle_lib.render(rd, sc.scene_data, sc.shapes, sc.triangles, sc.materials, sky, impl="ref")
        assert_bit_equal(ref, got, order + " vs render.cl")
