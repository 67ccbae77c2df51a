"""The fused frame of srt_render_frame (Tracer.render): accumulation, `average` and the banded read-back happen inside
the render kernel's own run (srt::FrameOut).  Canvas and image must be those of the separate steps -- srt_render +
srt_resolve, i.e. the reference's three enqueue calls (src/tracer.cpp:103-115) -- bit for bit, and equal to the
reference kernel itself, for every kernel build, sample count, frame shape and output kind."""
import hashlib

import numpy as np
import pytest

from conftest import assert_bit_equal
from simple_raytracer_b200 import scenes
from simple_raytracer_b200.tracer import Tracer
from util import make_tracer, random_scene

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not __import__("oracle").ref_available(), reason="oracle/_ref library not present")


def frames(tr, sc, n, out, ticks0=1, **rd_kw):
    tr.clear_canvas()
    imgs = []
    for k in range(n):
        tr.options[:] = sc.render_data(k, **rd_kw)
        tr.render(ticks0 + k, out.reshape(-1))
        imgs.append(out.copy())
    return tr.read_canvas(), imgs


@pytest.mark.parametrize("cfg,w,h", [(1, 203, 151), (2, 320, 97), (3, 161, 120), (5, 96, 54), (2, 1, 1), (1, 7, 300)])
@pytest.mark.parametrize("ns", [1, 2, 3, 4, 7])
def test_fused_frame_equals_separate_steps(sky, schedule, cfg, w, h, ns):
    """Every frame of a progressive run: image after frame k and canvas at the end, fused == separate, for all kernel
    builds (analytic / small models / dense sweep), both warp schedules, sample counts that are and are not powers of
    two (the `/ num_samples` of render.cl:520 as a multiply or a division), widths and heights that do not divide
    into bands or warps."""
    sc = scenes.CONFIGS[cfg](w, h)
    out = np.zeros((h, w, 4), np.uint8)
    tr = make_tracer(sc, sky)
    tr.set_frame_pipeline("fused")
    got_canvas, got = frames(tr, sc, 3, out, num_samples=ns)
    tr.set_frame_pipeline("separate")
    want_canvas, want = frames(tr, sc, 3, out, num_samples=ns)
    assert_bit_equal(want_canvas, got_canvas, "canvas after three fused frames")
    for k in range(3):
        assert np.array_equal(want[k], got[k]), f"image of frame {k}"
    tr.close()


@needs_ref
@pytest.mark.parametrize("seed", range(6))
def test_fused_frames_of_random_scenes_against_the_reference_kernel(sky, oracle_lib, seed):
    """Random scenes (every material branch, boxes, triangle soups of 0 / 40 / 700 triangles -> all kernel builds)
    through Tracer.render, against render.cl's `render` + `average` compiled as they are (oracle/_ref)."""
    sc = random_scene(100 + seed, width=131, height=77, mesh_tris=(0, 40, 700)[seed % 3])
    tr = make_tracer(sc, sky)
    tr.set_frame_pipeline("fused")
    out = np.zeros(131 * 77 * 4, np.uint8)
    canvas = None
    tr.clear_canvas()
    for tick in range(1, 4):
        tr.options[:] = sc.render_data(tick, num_samples=3, num_bounces=6)
        tr.render(tick, out)
        canvas, _ = oracle_lib.render(tr.options, sc.scene_data, sc.shapes, sc.triangles, sc.materials, sky, canvas,
                                      impl="ref")
        assert np.array_equal(out.reshape(77, 131, 4), oracle_lib.average(tick, canvas, impl="ref")), f"frame {tick}"
    assert_bit_equal(canvas, tr.read_canvas(), "canvas")
    tr.close()


@pytest.mark.parametrize("cfg,w,h", [(1, 203, 151), (2, 320, 97), (3, 161, 120), (2, 1, 1), (2, 1920, 1080)])
@pytest.mark.parametrize("ns", [3, 4])
def test_frame_epilogue_into_the_pinned_vector_equals_separate_steps(sky, cfg, w, h, ns):
    """A frame of more than 2 spp into the page-locked caller vector runs accumulate + average + the store into the
    caller's memory as ONE epilogue kernel (frame_epilogue_kernel, 16-byte stores over PCIe; pixel counts that are
    not a multiple of four take the tail branch).  Images of every frame, the device image and the canvas must equal
    the separate steps; an unpinned vector, and a view at an offset inside the pinned one, keep working."""
    sc = scenes.CONFIGS[cfg](w, h)
    tr = make_tracer(sc, sky)
    big = np.zeros(w * h * 4 + 64, np.uint8)
    tr.pin_output(big)
    out = big[16:16 + w * h * 4].reshape(h, w, 4)  # 16-byte aligned inside the pinned range: the epilogue path
    got_canvas, got = frames(tr, sc, 3, out, num_samples=ns)
    dev_img = tr.read_output()
    odd = big[4:4 + w * h * 4].reshape(h, w, 4)    # not 16-byte aligned: copy-engine path
    _, got_odd = frames(tr, sc, 3, odd, num_samples=ns)
    tr.set_frame_pipeline("separate")
    want_canvas, want = frames(tr, sc, 3, out, num_samples=ns)
    tr.unpin_output()
    assert_bit_equal(want_canvas, got_canvas, "canvas")
    for k in range(3):
        assert np.array_equal(want[k], got[k]) and np.array_equal(want[k], got_odd[k]), f"image of frame {k}"
    assert np.array_equal(dev_img, want[2])
    tr.close()


@pytest.mark.parametrize("pinned", [False, True])
def test_fused_frame_at_1080p_and_4k(sky, pinned):
    """Full BASELINE frame sizes (32 bands of 34 / 68 rows), the caller's vector page-locked or not."""
    for w, h in ((1920, 1080), (3840, 2160)):
        sc = scenes.config2(w, h)
        tr = make_tracer(sc, sky)
        out = np.zeros((h, w, 4), np.uint8)
        if pinned:
            tr.pin_output(out)
        tr.set_frame_pipeline("fused")
        got_canvas, got = frames(tr, sc, 2, out, num_samples=2)
        tr.set_frame_pipeline("separate")
        want_canvas, want = frames(tr, sc, 2, out, num_samples=2)
        assert_bit_equal(want_canvas, got_canvas, f"{w}x{h} canvas")
        assert np.array_equal(want[0], got[0]) and np.array_equal(want[1], got[1])
        tr.close()


def test_fused_frame_is_deterministic_and_leaves_no_state_behind(sky):
    """The completion counters are reset by the lanes that complete them: 40 frames in a row, interleaved with plain
    launches, batches, show_normals frames and a change of num_samples, give the same bytes every time."""
    sc = scenes.config1(257, 129)
    tr = make_tracer(sc, sky)
    tr.set_frame_pipeline("fused")
    out = np.zeros(257 * 129 * 4, np.uint8)
    digests = set()
    for rep in range(8):
        tr.clear_canvas()
        h = hashlib.sha256()
        for k in range(5):
            tr.options[:] = sc.render_data(k, num_samples=(4, 3, 1, 4, 2)[k], show_normals=int(k == 2))
            tr.render(k + 1, out)
            h.update(out.tobytes())
            if k == 1:
                tr.accumulate(sc.render_data(9))
            if k == 3:
                tr.accumulate_batch([sc.render_data(20 + j) for j in range(3)])
        h.update(tr.read_canvas().tobytes())
        digests.add(h.hexdigest())
    assert len(digests) == 1
    tr.close()


def test_banded_and_empty_launches_take_the_separate_steps(sky, oracle_lib):
    """srt_set_row_bands (tile sharding) and zero-bounce launches are outside the fused pass; render() still works."""
    sc = scenes.config1(120, 80)
    tr = make_tracer(sc, sky)
    tr.set_frame_pipeline("fused")
    out = np.zeros(120 * 80 * 4, np.uint8)
    tr.clear_canvas()
    tr.options[:] = sc.render_data(0)
    tr.options["num_bounces"] = 0
    tr.render(1, out)
    assert not tr.read_canvas()[..., :3].any()
    tr.set_row_bands(3, 1, 2)
    tr.options[:] = sc.render_data(0)
    tr.render(1, out)
    got = tr.read_canvas()
    want, _ = oracle_lib.render(sc.render_data(0), sc.scene_data, sc.shapes, sc.triangles, sc.materials, sky)
    rows = (np.arange(80) // 3) % 2 == 1
    assert_bit_equal(want[rows], got[rows], "banded rows")
    assert not got[~rows][..., :3].any()
    tr.close()
