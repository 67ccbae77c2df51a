"""Size-independent properties of the path on the oracle: the same ones the GPU tests use at full size."""
import numpy as np
import pytest

from conftest import assert_bit_equal
from simple_raytracer_b200 import scenes


def render(oracle, sc, sky, k=0, canvas=None, **kw):
    rd_kw = {n: kw.pop(n) for n in list(kw) if n in ("num_samples", "num_bounces", "show_normals")}
    return oracle.render(sc.render_data(k, **rd_kw), sc.scene_data, sc.shapes, sc.triangles, sc.materials, sky,
                         canvas, **kw)


@pytest.mark.parametrize("cfg,w,h", [(1, 64, 48), (2, 64, 36), (3, 48, 27)])
def test_window_and_band_partitions_are_bit_identical_to_full_frame(oracle_lib, small_sky, cfg, w, h):
    sc = scenes.CONFIGS[cfg](w, h)
    full, cnt = render(oracle_lib, sc, small_sky, num_samples=2)
    # four windows
    parts = np.zeros_like(full)
    for win in [(0, 0, w // 2, h // 2), (w // 2, 0, w, h // 2), (0, h // 2, w // 2, h), (w // 2, h // 2, w, h)]:
        render(oracle_lib, sc, small_sky, canvas=parts, window=win, num_samples=2)
    assert_bit_equal(full, parts, "windows")
    # three interleaved row bands (tile sharding)
    bands = np.zeros_like(full)
    tot = 0
    for b in range(3):
        _, c = render(oracle_lib, sc, small_sky, canvas=bands, bands=(4, b, 3), num_samples=2)
        tot += int(c["samples"])
    assert_bit_equal(full, bands, "bands")
    assert tot == int(cnt["samples"]) == w * h * 2


def test_accumulation_is_a_sum_of_launch_means(oracle_lib, small_sky):
    sc = scenes.config2(48, 27)
    acc = None
    singles = []
    for k in range(3):
        acc, _ = render(oracle_lib, sc, small_sky, k=k, canvas=acc)
        singles.append(render(oracle_lib, sc, small_sky, k=k)[0])
    assert_bit_equal(acc, (singles[0] + singles[1]) + singles[2], "canvas += mean")
    assert not np.array_equal(singles[0], singles[1])  # time_k changes the seeds


def test_threads_do_not_change_results(oracle_lib, small_sky):
    sc = scenes.config3(40, 24)
    a, ca = render(oracle_lib, sc, small_sky, threads=1)
    b, cb = render(oracle_lib, sc, small_sky, threads=4)
    assert_bit_equal(a, b, "threads")
    assert tuple(ca) == tuple(cb)


def test_show_normals_is_primary_hit_only(oracle_lib, small_sky):
    sc = scenes.config1(64, 48)
    img, cnt = render(oracle_lib, sc, small_sky, show_normals=True, num_samples=1)
    assert int(cnt["bounces"]) == 64 * 48          # one closest_intersection per pixel
    rgb = img[..., :3]
    assert rgb.min() >= 0.0 and rgb.max() <= 1.0   # n * 0.5 + 0.5
    idx, t = oracle_lib.primary(sc.render_data(0, num_samples=1), sc.scene_data, sc.shapes, sc.triangles)
    assert (idx >= 0).all() and np.isfinite(t).all()  # closed room: every primary ray hits
    # the back wall (shape 4, normal +z) must show (0.5, 0.5, 1.0)
    back = idx == 4
    assert back.any() and np.allclose(rgb[back], [0.5, 0.5, 1.0])


def test_last_bounce_adds_emission_but_never_sky(oracle_lib, small_sky):
    """render.cl:415-416: with num_bounces = 1 a path that hits returns emission only; a miss returns sky."""
    sc = scenes.config2(48, 27)
    img, cnt = render(oracle_lib, sc, small_sky, num_samples=1, num_bounces=1)
    idx, _ = oracle_lib.primary(sc.render_data(0, num_samples=1), sc.scene_data, sc.shapes, sc.triangles)
    mats = sc.materials[sc.shapes["material"][np.maximum(idx, 0)]]
    expect = mats["emission"] * mats["emission_strength"][..., None]
    hit = idx >= 0
    assert np.array_equal(img[..., :3][hit], expect[hit].astype(np.float32))
    assert int(cnt["sky"]) == int((~hit).sum()) and (img[..., :3][~hit] > 0).all()


def test_zero_time_zeroes_every_seed(oracle_lib):
    assert oracle_lib.seed(5, 1234, 4, 0) == 0   # time = 0 is a legal (degenerate) stream


def test_average_known_values(oracle_lib):
    c = np.zeros((1, 5, 4), np.float32)
    c[0, :, 0] = [0.0, 1.0, 4.0, np.nan, -1.0]
    c[0, :, 1] = [0.5, 0.5, 0.5, 0.5, 0.5]
    out = oracle_lib.average(2, c)
    assert (out[..., 0] == 255).all()             # A,R,G,B byte order (render.cl:534)

    def aces(x):
        return min(max(x * (2.51 * x + 0.03) / (x * (2.43 * x + 0.59) + 0.14), 0.0), 1.0)
    want_r = [int(np.sqrt(aces(v / 2.0)) * 255.0) for v in (0.0, 1.0, 4.0)]
    assert abs(int(out[0, 0, 1]) - want_r[0]) <= 0 and abs(int(out[0, 1, 1]) - want_r[1]) <= 1
    assert abs(int(out[0, 2, 1]) - want_r[2]) <= 1
    assert out[0, 3, 1] == 0                        # NaN -> 0 (documented hazard ix)
    assert out[0, 4, 1] == int(np.sqrt(aces(-0.5)) * 255.0)
    assert out[0, 0, 3] == 0                        # untouched blue channel: aces(0) = 0


def test_counters_are_consistent(oracle_lib, small_sky):
    sc = scenes.config3(48, 27)
    _, c = render(oracle_lib, sc, small_sky, num_samples=2)
    assert int(c["samples"]) == 48 * 27 * 2
    assert int(c["bounces"]) == int(c["hits"]) + int(c["sky"])
    assert int(c["tri_tests"]) == int(c["aabb_pass"]) * 1280
    assert int(c["bounces"]) <= int(c["samples"]) * sc.num_bounces
