"""The oracle pinned against the reference's OWN kernel source.

oracle/_ref/libref_render_cl.so is /root/reference/src/render.cl compiled by g++ from where it lies
(oracle/ref_build/: an OpenCL-C language shim, one mechanical rewrite of vector literals, an NDRange loop);
see DESIGN.md section 2.  Here it plays the role north_star gives to "render.cl on a CPU OpenCL device":
  * oracle.c (the restatement that adds debug outputs and work counters) must reproduce it BIT FOR BIT --
    canvases, resolved ARGB8 images, primary-hit ids -- on every kind of scene;
  * the `contract` variant of the oracle (every a*b+c expression of render.cl fused, which an OpenCL
    compiler is allowed to do) must stay within the tolerances of SURVEY 8c: that is what the tolerances
    are for, and this test is where they are measured.
The library is built in the authoring container (where /root/reference exists) and travels to the GPU box;
without either the tests skip.
"""
import numpy as np
import pytest

import oracle
from conftest import assert_bit_equal
from simple_raytracer_b200 import scenes
from util import random_scene, ref_primary_ids

pytestmark = pytest.mark.skipif(not oracle.ref_available(), reason="no oracle/_ref library and no /root/reference")


def canvas_of(impl, sc, sky, launches, **kw):
    canvas = None
    for k in range(launches):
        canvas, _ = oracle.render(sc.render_data(k, **kw), sc.scene_data, sc.shapes, sc.triangles, sc.materials,
                                  sky, canvas, impl=impl)
    return canvas


@pytest.mark.parametrize("cfg,w,h,ns,launches", [(1, 200, 150, 2, 2), (2, 240, 136, 4, 2), (3, 120, 68, 2, 2),
                                                 (5, 40, 24, 1, 1)])
def test_oracle_equals_reference_kernel_on_the_baseline_configs(oracle_lib, small_sky, cfg, w, h, ns, launches):
    sc = scenes.CONFIGS[cfg](w, h)
    ref = canvas_of("ref", sc, small_sky, launches, num_samples=ns)
    got = canvas_of("oracle", sc, small_sky, launches, num_samples=ns)
    assert_bit_equal(ref, got, f"C{cfg}: oracle.c vs render.cl")
    assert (ref[..., :3] > 0).any() and not ref[..., 3].any()
    assert np.array_equal(oracle.average(launches, ref, impl="ref"), oracle.average(launches, got))


@pytest.mark.parametrize("seed", range(8))
def test_oracle_equals_reference_kernel_on_random_scenes(oracle_lib, small_sky, seed):
    """Random shape order, box instances, a triangle soup, every material branch."""
    sc = random_scene(seed, mesh_tris=60 if seed % 2 else 0)
    for kw in (dict(num_samples=3), dict(num_samples=1, num_bounces=1), dict(num_samples=1, show_normals=True)):
        rd = sc.render_data(seed, **kw)
        ref, _ = oracle.render(rd, sc.scene_data, sc.shapes, sc.triangles, sc.materials, small_sky, impl="ref")
        got, _ = oracle.render(rd, sc.scene_data, sc.shapes, sc.triangles, sc.materials, small_sky)
        assert_bit_equal(ref, got, f"seed {seed} {kw}")


@pytest.mark.parametrize("cfg,w,h", [(1, 200, 150), (2, 200, 112), (3, 160, 90)])
def test_primary_hit_shape_ids_from_the_reference_kernel(oracle_lib, cfg, w, h):
    """closest_intersection of render.cl itself names the same shape as the oracle's debug output, pixel by pixel."""
    sc = scenes.CONFIGS[cfg](w, h)
    rd = sc.render_data(0, num_samples=1)
    ids = ref_primary_ids(oracle, sc, rd)
    oi, _ = oracle.primary(rd, sc.scene_data, sc.shapes, sc.triangles)
    assert np.array_equal(ids, oi)
    assert len(np.unique(ids)) >= 3


def test_reference_kernel_window_and_bands_match_full_frame(small_sky):
    """The NDRange driver's sub-ranges keep global ids: a window / the union of row bands equals the full launch."""
    sc = scenes.config2(96, 54)
    rd = sc.render_data(0, num_samples=2)
    args = (rd, sc.scene_data, sc.shapes, sc.triangles, sc.materials, small_sky)
    full, _ = oracle.render(*args, impl="ref")
    win, _ = oracle.render(*args, impl="ref", window=(10, 5, 50, 40))
    assert_bit_equal(full[5:40, 10:50], win[5:40, 10:50], "window")
    assert not win[:5].any() and not win[:, :10].any()
    acc = None
    for b in range(3):
        acc, _ = oracle.render(*args, acc, impl="ref", bands=(4, b, 3))
    assert_bit_equal(full, acc, "bands")


def test_average_kernel_of_the_reference_on_special_values():
    """kernel `average` (render.cl:525-535): NaN, inf, negative, huge and tiny canvas values resolve identically."""
    vals = np.array([0.0, -0.0, 1e-30, 0.001, 0.18, 0.5, 1.0, 4.0, 1e6, 3e38, np.inf, -np.inf, np.nan, -1.0, -1e-3],
                    np.float32)
    rng = np.random.default_rng(1)
    canvas = np.zeros((64, 4), np.float32)
    canvas[:, :3] = rng.choice(vals, size=(64, 3))
    canvas[:16, :3] = rng.random((16, 3)) * 3
    for steps in (1, 3, 64):
        a = oracle.average(steps, canvas, impl="ref")
        b = oracle.average(steps, canvas)
        assert np.array_equal(a, b), steps
        assert (a[:, 0] == 255).all()


@pytest.mark.parametrize("cfg,w,h", [(2, 240, 136), (3, 160, 90)])
def test_contraction_sensitivity_stays_within_the_stated_tolerances(oracle_lib, small_sky, cfg, w, h):
    """What an FMA-contracting OpenCL compiler could change (measured, not a gate on the product, which is
    bit-exact against the uncontracted build): 1-spp radiance within 1e-4 relative on >= 97 % of pixels -- ten
    bounces amplify a last-bit difference until a discrete decision flips, so even this legal compiler freedom
    misses the 99 % SURVEY 8c proposed (98.6 % on config 2) -- and converged images within RMSE <= 1/255."""
    sc = scenes.CONFIGS[cfg](w, h)
    one_ref = canvas_of("ref", sc, small_sky, 1, num_samples=1)
    one_con = canvas_of("contract", sc, small_sky, 1, num_samples=1)
    close = np.abs(one_con[..., :3] - one_ref[..., :3]) <= 1e-4 * np.abs(one_ref[..., :3]) + 1e-6
    assert close.all(axis=-1).mean() >= 0.97
    ref = canvas_of("ref", sc, small_sky, 6, num_samples=4)
    con = canvas_of("contract", sc, small_sky, 6, num_samples=4)
    a, b = oracle.average(6, ref, impl="ref").astype(float), oracle.average(6, con, impl="contract").astype(float)
    rmse = np.sqrt(np.mean((a - b) ** 2)) / 255.0
    assert rmse <= 1.0 / 255.0, rmse


def test_ref_library_was_built_from_the_reference_file_as_it_lies():
    """oracle/_ref/render.cl.sha256 (written by the Makefile next to the library) names the file that was compiled;
    where the reference is present it must be that file, byte for byte, and the rewrite must touch nothing but
    vector literals."""
    import hashlib
    import os
    import re
    import subprocess
    here = os.path.dirname(os.path.abspath(oracle.__file__))
    sha_file = os.path.join(here, "_ref", "render.cl.sha256")
    if not os.path.exists(oracle.REFERENCE_KERNEL):
        pytest.skip("reference source not present (GPU box): the prebuilt library is used")
    oracle.build_ref()
    src = open(oracle.REFERENCE_KERNEL, "rb").read()
    assert open(sha_file).read().strip() == hashlib.sha256(src).hexdigest()
    out = subprocess.run(["python3", os.path.join(here, "ref_build", "rewrite_cl.py"), oracle.REFERENCE_KERNEL],
                         capture_output=True, text=True, check=True).stdout
    body = out.split("\n", 1)[1]  # drop the #line directive
    # undo the rewrite, T(T_lit{ ... }) -> (T)( ... ), and compare modulo white space: nothing else may differ
    strip = lambda t: re.sub(r"\s+", "", t)  # noqa: E731
    assert "})" not in strip(src.decode())
    undone = strip(re.sub(r"\b(float2|float3|float4|uchar4)\(\1_lit\{", r"(\1)(", body)).replace("})", ")")
    assert undone == strip(src.decode())


@pytest.mark.parametrize("seed", range(100, 124))
def test_oracle_equals_reference_kernel_on_more_random_scenes(oracle_lib, small_sky, seed):
    sc = random_scene(seed, width=48, height=32, n_spheres=1 + seed % 6, n_planes=seed % 4, n_boxes=seed % 3,
                      mesh_tris=(0, 7, 33, 150)[seed % 4])
    rd = sc.render_data(seed, num_samples=2, num_bounces=1 + seed % 12)
    ref, _ = oracle.render(rd, sc.scene_data, sc.shapes, sc.triangles, sc.materials, small_sky, impl="ref")
    got, _ = oracle.render(rd, sc.scene_data, sc.shapes, sc.triangles, sc.materials, small_sky)
    assert_bit_equal(ref, got, f"seed {seed}")


def test_reference_kernel_result_does_not_depend_on_the_optimisation_level(small_sky, tmp_path):
    """The _ref build is meant to be THE IEEE-754 result of render.cl under the documented builtins, not an artefact
    of g++ -O2: the same source compiled -O0 and -O3 (still -ffp-contract=off) must give the same bits."""
    import ctypes
    import os
    import subprocess
    if not os.path.exists(oracle.REFERENCE_KERNEL):
        pytest.skip("reference source not present")
    here = os.path.dirname(os.path.abspath(oracle.__file__))
    sc = scenes.config3(64, 36)
    rd = sc.render_data(3, num_samples=2)
    want, _ = oracle.render(rd, sc.scene_data, sc.shapes, sc.triangles, sc.materials, small_sky, impl="ref")
    for opt in ("-O0", "-O3"):
        so = str(tmp_path / f"ref{opt}.so")
        define = "-DREF_KERNEL_SOURCE='\"/dev/stdin\"'"
        cmd = (f"python3 {here}/ref_build/rewrite_cl.py {oracle.REFERENCE_KERNEL} | g++ -std=c++20 {opt} -march=x86-64-v3 "
               f"-ffp-contract=off -fno-fast-math -fopenmp -fPIC -Wno-narrowing -w -I{here}/ref_build "
               f"{define} -shared -o {so} {here}/ref_build/ref_driver.cpp -lm")
        subprocess.check_call(["bash", "-c", cmd])
        lib = ctypes.CDLL(so)
        vp, i32 = ctypes.c_void_p, ctypes.c_int
        lib.ref_render.argtypes = [vp] * 7 + [i32] * 10
        canvas = np.zeros((36, 64, 4), np.float32)
        sd = sc.scene_data.copy()
        p = lambda a: a.ctypes.data_as(vp)  # noqa: E731
        lib.ref_render(p(rd), p(sd), p(canvas), p(sc.shapes), p(sc.triangles), p(sc.materials), p(small_sky),
                       small_sky.shape[1], small_sky.shape[0], 0, 0, 64, 36, 1, 0, 1, 0)
        assert_bit_equal(want, canvas, f"g++ {opt}")


def test_hypothesis_random_launch_parameters_oracle_equals_reference_kernel(oracle_lib, small_sky):
    """Property test over the launch parameters themselves: image size (down to 1x1), sample and bounce counts, time
    seeds (incl. 0 and values whose product with 5304 wraps), camera pose, fov, show_normals."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None, derandomize=True)
    @given(w=st.integers(1, 40), h=st.integers(1, 30), ns=st.integers(1, 5), nb=st.integers(0, 9),
           time=st.sampled_from([0, 1, 2, 1000003, 809679, 0x7fffffff, 0xffffffff, 404838]),
           yaw=st.floats(-3.0, 3.0), pitch=st.floats(-1.2, 1.2), fov=st.floats(0.2, 2.5), normals=st.booleans(),
           seed=st.integers(0, 5))
    def check(w, h, ns, nb, time, yaw, pitch, fov, normals, seed):
        sc = random_scene(seed, width=w, height=h, mesh_tris=9 if seed % 2 else 0)
        sc.camera = scenes.camera_matrix((0.1 * seed, 0.3, 3.0), yaw, pitch)
        sc.fov_scale = np.float32(fov)
        rd = sc.render_data(0, num_samples=ns, show_normals=normals)
        rd["num_bounces"] = nb
        rd["time"] = time
        ref, _ = oracle.render(rd, sc.scene_data, sc.shapes, sc.triangles, sc.materials, small_sky, impl="ref")
        got, _ = oracle.render(rd, sc.scene_data, sc.shapes, sc.triangles, sc.materials, small_sky)
        assert_bit_equal(ref, got, f"{w}x{h} ns={ns} nb={nb} time={time}")

    check()
