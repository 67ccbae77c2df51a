"""Round-2 parity gaps: config 5 at its BASELINE size, the mesh-file data path
(.obj / .stl -> srt_load_* -> srt_model_bounds -> srt_upload_scene -> render) against the reference kernel, the
work-item cap, and the read-back paths (staging / explicitly pinned output).  All through the C ABI."""
import struct

import numpy as np
import pytest

from conftest import assert_bit_equal
from simple_raytracer_b200 import scenes, tracer
from simple_raytracer_b200.records import RENDER_DATA, TRIANGLE
from simple_raytracer_b200.tracer import SrtError, Tracer
from test_mesh_io import write_obj, write_stl
from util import make_tracer, reference_parse

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not __import__("oracle").ref_available(), reason="oracle/_ref library not present")


def _ref_window(oracle, sc, sky, rds, win):
    canvas = None
    for rd in rds:
        canvas, _ = oracle.render(rd, sc.scene_data, sc.shapes, sc.triangles, sc.materials, sky, canvas,
                                  window=win, impl="ref")
    x0, y0, x1, y1 = win
    return canvas[y0:y1, x0:x1]


@needs_ref
def test_config5_at_1080p_16_bounces_crop_against_the_reference_kernel(sky, oracle_lib):
    """BASELINE configs[4] at its own size: 1920x1080, 100 352 triangles, 16 bounces.  The CUDA frame equals the
    reference's render.cl (oracle/_ref) bit for bit on an 80x48 window of the same full-size launch (global pixel ids,
    seeds and aspect preserved) placed where the mesh, its shadow and the sky meet."""
    sc = scenes.config5()
    assert (sc.width, sc.height, sc.num_bounces, len(sc.triangles)) == (1920, 1080, 16, 100352)
    tr = make_tracer(sc, sky)
    rd = sc.render_data(0, num_samples=1)
    tr.accumulate(rd)
    got = tr.read_canvas()
    x0, y0, w, h = 860, 380, 80, 48
    want = _ref_window(oracle_lib, sc, sky, [rd], (x0, y0, x0 + w, y0 + h))
    assert_bit_equal(want, got[y0:y0 + h, x0:x0 + w], "C5 1080p crop vs render.cl")
    ids, _ = tr.debug_primary(rd)
    crop_ids = ids[y0:y0 + h, x0:x0 + w]
    assert (crop_ids == 1).any() and (crop_ids != 1).any()  # the window straddles the mesh's silhouette


@needs_ref
def test_config5_at_1080p_batched_launches(sky, oracle_lib):
    """Multi-launch batches on the 100k-triangle mesh at 1080p: ONE persistent kernel over 3 launches x 2 samples gives
    the canvas of 3 separate launches bit for bit, and both equal the reference kernel's accumulation on a window."""
    sc = scenes.config5()
    tr = make_tracer(sc, sky)
    rds = [sc.render_data(k, num_samples=2) for k in range(3)]
    tr.clear_canvas()
    tr.accumulate_batch(rds)
    batched = tr.read_canvas()
    tr.clear_canvas()
    for rd in rds:
        tr.accumulate(rd)
    separate = tr.read_canvas()
    assert_bit_equal(separate, batched, "C5 1080p: batch of 3 launches vs 3 launches")
    x0, y0, w, h = 1000, 500, 48, 32
    want = _ref_window(oracle_lib, sc, sky, rds, (x0, y0, x0 + w, y0 + h))
    assert_bit_equal(want, batched[y0:y0 + h, x0:x0 + w], "C5 1080p batch crop vs render.cl")
    assert np.array_equal(oracle_lib.average(3, want, impl="ref"), tr.resolve(3)[y0:y0 + h, x0:x0 + w])


@needs_ref
@pytest.mark.parametrize("kind", ["stl", "obj"])
def test_mesh_files_through_the_loaders_against_the_reference_kernel(tmp_path, sky, oracle_lib, kind):
    """The data path BASELINE.json names: a mesh FILE -> srt_load_stl / srt_load_obj (reference parser.cpp:17-135) ->
    Model(triangles, first, count) with srt_model_bounds (shape.cpp:37-58) -> srt_upload_scene -> render.  The canvas
    must equal the reference kernel's on the triangle list parser.cpp would have produced from the same file."""
    if kind == "stl":
        v, f = scenes.displaced_torus(40, 24, seed=21)
        n = None
        tris = scenes.mesh_triangles(v, f, None)
        tris["v"]["normal"] *= np.float32(1.7)  # facet normals are used as stored, not renormalised (:46-50)
        path = tmp_path / "mesh.stl"
        write_stl(path, tris)
        loaded = tracer.load_stl_model(str(path), scenes.cube_triangles())
    else:
        v, n, f = scenes.noisy_icosphere(2, seed=22)
        n = (n * np.float32(0.37)).astype(np.float32)  # unnormalised `vn` lines: the loader normalises (:83)
        tris = None
        path = tmp_path / "mesh.obj"
        write_obj(path, v, n, f)
        loaded = tracer.load_obj_model(str(path), scenes.cube_triangles())
    (first, count), all_tris = loaded
    expect = reference_parse(kind, v, n, f, tris)
    assert first == 12 and count == len(expect)
    assert all_tris[12:].tobytes() == expect.tobytes(), "loader output differs from parser.cpp's semantics"

    # Model(triangles, first, count): identity transform, AABB over the triangles (shape.cpp:37-58), as the "Add model"
    # popup does (interface.cpp:281-301); then a UI edit of the transform with the AABB recomputed (interface.cpp:98-101)
    rec = scenes.model(1, all_tris, first, count)
    rec["model_bounding_min"], rec["model_bounding_max"] = 9.0, -9.0
    ident = tracer.model_bounds(rec, all_tris)
    pos = all_tris["v"]["pos"][first:first + count].reshape(-1, 3)
    assert np.array_equal(ident["model_bounding_min"][:3], pos.min(0)) and np.array_equal(ident["model_bounding_max"][:3], pos.max(0))
    xf = scenes.translate((0.3, 0.1, -2.5)) @ scenes.rotate_y(0.7) @ scenes.rotate_x(0.4) @ scenes.scale((0.9, 1.1, 0.8))
    moved = scenes.model(2, all_tris, first, count, xf)
    moved = tracer.model_bounds(moved, all_tris)
    mats = scenes._stack([scenes.material((0.8, 0.8, 0.8)),
                          scenes.material((0.3, 0.9, 0.4), smoothness=0.9, transmittance=0.9, refraction_index=1.3),
                          scenes.material((0.9, 0.55, 0.25), smoothness=0.8, specular=0.25)], scenes.MATERIAL)
    shapes = scenes._stack([scenes.plane(0, (0, -1.6, 0), (0, 1, 0)), moved,
                            scenes.model(1, all_tris, 0, 12, scenes.translate((-2.2, -0.9, -2.0)) @ scenes.scale(0.6))],
                           scenes.SHAPE)
    sc = scenes.Scene(f"mesh file {kind}", 240, 136, 2, 8, 2, shapes, all_tris, mats, scenes.camera_matrix((0, 0.3, 3.0)))
    tr = make_tracer(sc, sky)
    rds = [sc.render_data(k) for k in range(2)]
    for rd in rds:
        tr.accumulate(rd)
    got = tr.read_canvas()
    want = _ref_window(oracle_lib, sc, sky, rds, (0, 0, sc.width, sc.height))
    assert_bit_equal(want, got, f"{kind} file -> loader -> upload -> render vs render.cl")
    ids, _ = tr.debug_primary(sc.render_data(0, num_samples=1))
    assert (ids == 1).sum() > 200  # the loaded mesh is in view


def test_work_item_cap_and_64_bit_cursor(sky):
    """Items are 32-bit, the cursor they are dealt from is 64-bit: a launch above the cap is refused, a launch right at
    it runs (when the per-sample scratch fits) and every item is rendered exactly once."""
    import torch
    sc = scenes.config2(64, 64)
    tr = make_tracer(sc, sky)
    cap = 0xFFFFFF00
    rd = sc.render_data(0, num_samples=cap // (64 * 64) + 1)
    with pytest.raises(SrtError, match="work items"):
        tr.accumulate(rd)
    # a smaller launch still goes through the same cursor arithmetic: compare num_samples = 4096 (16.8 M items, far more
    # than the resident threads) with the sum of what 4096 launches' worth would be is too slow; instead check the
    # full-cap launch when 80 GB are free, else a 2^28-item launch
    free, _ = torch.cuda.mem_get_info()
    ns = cap // (64 * 64) if free > 90 * 2 ** 30 else (1 << 28) // (64 * 64)
    rd = sc.render_data(0, num_samples=ns, num_bounces=2)
    tr.clear_canvas()
    tr.accumulate(rd)
    a = tr.read_canvas()
    # log(0) is legal in the reference arithmetic (probability 2^-32 per draw, render.cl:152): at 4.3e9 samples x ~20
    # draws a handful of samples are inf / NaN and poison their pixel's mean -- on both sides of the parity contract
    fin = np.isfinite(a[..., :3]).all(axis=-1)
    assert (~fin).sum() <= 200 and a[fin][:, :3].mean() > 0
    # the mean over that many samples is the converged 2-bounce image: a 4096-sample launch agrees to Monte-Carlo noise
    tr.clear_canvas()
    tr.accumulate(sc.render_data(1, num_samples=4096, num_bounces=2))
    b = tr.read_canvas()
    fin &= np.isfinite(b[..., :3]).all(axis=-1)
    assert np.abs(a[fin][:, :3] - b[fin][:, :3]).mean() < 0.15 * b[fin][:, :3].mean()
    cnt = tr.accumulate_counted(sc.render_data(0, num_samples=1 << 12, num_bounces=1))
    assert int(cnt[0]["samples"]) == 64 * 64 * (1 << 12)


def test_read_back_into_fresh_buffers_and_pinned_output(sky):
    """the library never page-locks caller memory on its own.  Resolving into a fresh buffer every
    frame (freed between frames: numpy mmaps >= 32 MiB... and smaller ones may reuse addresses) always returns the
    image; the explicit srt_pin_output path returns the same bytes; unpinning restores the staging path."""
    sc = scenes.config2(1280, 720)
    tr = make_tracer(sc, sky)
    tr.accumulate(sc.render_data(0))
    want = tr.resolve(1).copy()
    assert want[..., 0].min() == 255 and want[..., 1:].any()
    for i in range(6):  # same-sized fresh allocations tend to land on the same address
        out = np.empty((sc.height, sc.width, 4), np.uint8)
        out[:] = i
        got = tr.resolve(1, out)
        assert np.array_equal(got, want)
        del out, got
    pinned = np.zeros((sc.height, sc.width, 4), np.uint8)
    tr.pin_output(pinned)
    assert np.array_equal(tr.resolve(1, pinned), want)
    other = np.zeros_like(pinned)
    assert np.array_equal(tr.resolve(1, other), want)  # a different buffer while one is pinned: staging path
    tr.options[:] = sc.render_data(1)
    tr.render(2, pinned.reshape(-1))  # the reference-facing call lands in the pinned vector as well
    tr.unpin_output()
    again = np.zeros_like(pinned)
    tr.resolve(2, again)
    assert np.array_equal(again, pinned)
    tr.pin_output(pinned)
    tr.close()  # destroy while pinned: unregisters before the buffer goes away


def test_failed_upload_keeps_the_handle_consistent(sky):
    """A scene refused by validation leaves the previous scene in place (nothing was touched yet); the handle keeps
    rendering it bit for bit."""
    sc = scenes.config1(200, 150)
    tr = make_tracer(sc, sky)
    tr.accumulate(sc.render_data(0))
    before = tr.read_canvas()
    bad = sc.shapes.copy()
    bad["material"][3] = 99
    with pytest.raises(SrtError, match="material"):
        tr.update_scene(bad, sc.triangles, sc.materials)
    tr.clear_canvas()
    tr.accumulate(sc.render_data(0))
    assert_bit_equal(before, tr.read_canvas(), "scene after a refused upload")
    rd = np.zeros(1, RENDER_DATA)
    rd[:] = sc.render_data(0)
    rd["width"] = 199
    with pytest.raises(SrtError):
        tr.accumulate(rd)


@pytest.mark.parametrize("cfg,w,h", [(2, 1920, 1080), (3, 1000, 701), (1, 800, 600)])
@pytest.mark.parametrize("pinned", [False, True])
def test_frame_call_equals_render_then_resolve_at_full_size(sky, cfg, w, h, pinned):
    """Tracer.render (srt_render_frame: render + average + blocking read-back in one call, tracer.cpp:103-116) at full
    frame sizes: canvas and ARGB8 image are exactly those of srt_render + srt_resolve -- analytic, small-model and
    dense-sweep kernel builds, an odd height, accumulation over several frames, staged and page-locked outputs."""
    sc = scenes.CONFIGS[cfg](w, h)
    tr = make_tracer(sc, sky)
    out = np.zeros((h, w, 4), np.uint8)
    if pinned:
        tr.pin_output(out)
    tr.clear_canvas()
    for k in range(3):
        tr.options[:] = sc.render_data(k, num_samples=2)
        tr.render(k + 1, out.reshape(-1))
    got_canvas, got_img = tr.read_canvas(), out.copy()
    if pinned:
        tr.unpin_output()
    tr.clear_canvas()
    for k in range(3):
        tr.accumulate(sc.render_data(k, num_samples=2))
    assert_bit_equal(tr.read_canvas(), got_canvas, "canvas after three pipelined frames")
    assert np.array_equal(tr.resolve(3), got_img)
    assert got_img[..., 0].min() == 255 and got_img[..., 1:].any()


def test_more_than_256_bounces_keep_the_plain_schedule(sky, oracle_lib, schedule):
    """The wavefront schedule keeps the bounce count in 8 bits of a hit record; a launch asking for more bounces (a
    closed mirror room keeps paths alive that long) runs the plain schedule instead and still equals the oracle."""
    mats = [scenes.material((0.98, 0.98, 0.98), smoothness=1.0, metallic=1.0),
            scenes.material((1, 1, 1), emission=(1.0, 0.9, 0.8), emission_strength=0.02)]
    shapes = [scenes.plane(0, (-2, 0, 0), (1, 0, 0)), scenes.plane(0, (2, 0, 0), (-1, 0, 0)),
              scenes.plane(0, (0, -2, 0), (0, 1, 0)), scenes.plane(0, (0, 2, 0), (0, -1, 0)),
              scenes.plane(0, (0, 0, -2), (0, 0, 1)), scenes.plane(1, (0, 0, 6), (0, 0, -1)),
              scenes.sphere(0, (0.3, -0.4, 0.0), 0.8)]
    sc = scenes.Scene("mirror room", 48, 32, 1, 300, 1, scenes._stack(shapes, scenes.SHAPE), np.zeros(0, scenes.TRIANGLE),
                      scenes._stack(mats, scenes.MATERIAL), scenes.camera_matrix((0, 0, 5)))
    tr = make_tracer(sc, sky)
    for nb in (256, 257, 300):
        rd = sc.render_data(0, num_bounces=nb)  # (under the forced wavefront schedule 256 still fits, 257 and 300 do not)
        tr.clear_canvas()
        cnt = tr.accumulate_counted(rd)
        want, ocnt = oracle_lib.render(rd, sc.scene_data, sc.shapes, sc.triangles, sc.materials, sky)
        assert_bit_equal(want, tr.read_canvas(), f"{nb} bounces")
        assert int(cnt[0]["bounces"]) == int(ocnt["bounces"]) and int(cnt[0]["bounces"]) > 48 * 32 * 100
