"""srt_render_frame (Tracer.render, reference src/tracer.cpp:103-116) has two executions: the separate steps (render
kernel + accumulate, `average`, copy-engine read-back) and -- for a full frame into the vector the caller page-locked --
ONE epilogue kernel that sums the samples, applies `average` and stores the ARGB8 image straight into the caller's memory
(frame_epilogue_kernel).  Canvas and image must be the same bytes either way, and those of the reference kernel."""
import hashlib

import numpy as np
import pytest

from conftest import assert_bit_equal
from simple_raytracer_b200 import scenes
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


@pytest.mark.parametrize("cfg,w,h", [(1, 203, 151), (2, 320, 97), (3, 161, 120), (5, 96, 54), (2, 1, 1), (1, 7, 300),
                                     (2, 1920, 1080)])
@pytest.mark.parametrize("ns", [1, 2, 3, 4, 7])
def test_frame_epilogue_into_the_pinned_vector_equals_separate_steps(sky, schedule, cfg, w, h, ns):
    """Every frame of a progressive run -- image after frame k, device image and canvas at the end -- epilogue kernel ==
    separate steps, for all kernel builds (analytic / small models / dense sweep), both warp schedules, sample counts
    that are and are not powers of two (the `/ num_samples` of render.cl:520 as a multiply or a division), pixel counts
    that are not a multiple of four (the kernel's tail branch).  A view at a misaligned offset inside the pinned vector
    takes the copy-engine path and gives the same bytes."""
    if (w, h) == (1920, 1080) and ns not in (2, 4):
        pytest.skip("full size once per kind of division")
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


@needs_ref
@pytest.mark.parametrize("seed", range(6))
def test_frames_of_random_scenes_against_the_reference_kernel(sky, oracle_lib, seed):
    """Random scenes (every material branch, boxes, triangle soups of 0 / 40 / 700 triangles -> all kernel builds)
    through Tracer.render into a pinned vector (the epilogue kernel), against render.cl's `render` + `average` compiled
    as they are (oracle/_ref)."""
    sc = random_scene(100 + seed, width=131, height=77, mesh_tris=(0, 40, 700)[seed % 3])
    tr = make_tracer(sc, sky)
    out = np.zeros(131 * 77 * 4, np.uint8)
    tr.pin_output(out)
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


def test_frames_from_fresh_handles_are_deterministic(sky):
    """A fresh handle for every repetition (fresh scratch, fresh canvases), frames interleaved with plain launches,
    batches, show_normals frames and changes of num_samples: the same bytes every time, in both frame pipelines."""
    sc = scenes.config3(161, 120)
    digests = set()
    for rep in range(12):
        tr = make_tracer(sc, sky)
        out = np.zeros(161 * 120 * 4, np.uint8)
        if rep % 2:
            tr.pin_output(out)
        if rep % 4 >= 2:
            tr.set_frame_pipeline("separate")
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
        tr.close()
    assert len(digests) == 1


def test_banded_and_empty_launches_take_the_separate_steps(sky, oracle_lib):
    """srt_set_row_bands (tile sharding) and zero-bounce launches are outside the epilogue path; render() still works."""
    sc = scenes.config1(120, 80)
    tr = make_tracer(sc, sky)
    out = np.zeros(120 * 80 * 4, np.uint8)
    tr.pin_output(out)
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
