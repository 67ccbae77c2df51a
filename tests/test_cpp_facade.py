"""include/tracer.hpp: the C++ facade with the reference Tracer's surface, driven by examples/headless.cpp
the way src/main.cpp drives the reference."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "headless_test")


def build_harness():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    lib = os.path.join(ROOT, "simple_raytracer_b200")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "headless.cpp"), "-L" + lib, "-lsrt_b200",
                           "-Wl,-rpath," + lib, "-o", EXE])


def test_facade_compiles_and_links():
    build_harness()
    assert os.path.exists(EXE)


def test_c_header_is_plain_c(tmp_path):
    """include/srt.h must be consumable from C (the cgo / ctypes / JNI side of the boundary)."""
    src = tmp_path / "t.c"
    src.write_text('#include "srt.h"\n_Static_assert(sizeof(srt_shape) == 128 && sizeof(srt_triangle) == 96 && '
                   'sizeof(srt_material) == 64 && sizeof(srt_render_data) == 112 && sizeof(srt_scene_data) == 96, "abi");\n'
                   "int main(void) { return srt_abi_version() == SRT_ABI_VERSION ? 0 : 1; }\n")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-c", str(src),
                           "-o", str(tmp_path / "t.o")])


def _mirror_scene(w, h, frames, mesh_tris=None):
    """The scene examples/headless.cpp builds with the C++ helpers, rebuilt with the Python ones."""
    from simple_raytracer_b200 import scenes
    from simple_raytracer_b200.records import concat_records
    F = np.float32
    mats = [scenes.material((0.8, 0.8, 0.8)),
            scenes.material((1, 1, 1), smoothness=1.0, transmittance=1.0, refraction_index=1.5),
            scenes.material((0.25, 0.4, 0.95), smoothness=0.95, metallic=1.0),
            scenes.material((1, 0.2, 0.15), emission=(1, 0.15, 0.1), emission_strength=5.0)]
    tris = scenes.reference_cube_triangles()
    shapes = [scenes.plane(0, (0, -2, 0), (0, 1, 0)), scenes.sphere(0, (-3.2, 0, -3), 2.0),
              scenes.sphere(1, (0.6, -0.8, -0.5), 1.2), scenes.sphere(2, (3.4, -0.6, -2.6), 1.4),
              scenes.sphere(3, (-0.4, -1.3, -4.6), 0.7), scenes.box_model(2, (1.5, -1.0, 1.0), 2.0)]
    if mesh_tris is not None:
        tris = concat_records(scenes.TRIANGLE, tris, mesh_tris)
        shapes.append(scenes.model(0, tris, 12, len(mesh_tris)))
    sc = scenes.Scene("headless", w, h, 2, 10, frames, scenes._stack(shapes, scenes.SHAPE), tris,
                      scenes._stack(mats, scenes.MATERIAL), scenes.camera_matrix((0, 0.5, 5.5)))
    sc.scene_data["sun_color"] = [1.0, 1.0, F(0xD3) / F(255.0)]
    inv = F(1.0) / np.sqrt(F(2.0))
    sc.scene_data["sun_direction"] = [inv, -inv, 0.0]
    return sc


def _mesh_file(tmp_path):
    from simple_raytracer_b200 import scenes
    from test_mesh_io import write_stl
    v, f = scenes.displaced_torus(20, 12, seed=31, R=0.9, r=0.35)
    v = (v + np.array([-1.2, -0.6, 1.8], np.float32)).astype(np.float32)
    tris = scenes.mesh_triangles(v, f, None)
    path = tmp_path / "torus.stl"
    write_stl(path, tris)
    return path, tris


def test_cpp_scene_helpers_build_the_same_bytes_as_the_python_mirror(tmp_path):
    """include/scene.hpp (Material / Sphere / Plane / Model / Box / Shape constructors, load_stl_model) against
    records.py / scenes.py: the C++ harness dumps its three scene vectors, which must equal the Python-built scene byte
    for byte -- including Box::create_triangle's triangle order and Model's bounding box.  No GPU involved."""
    build_harness()
    path, mesh = _mesh_file(tmp_path)
    prefix = str(tmp_path / "scene")
    subprocess.check_call([EXE, str(tmp_path / "unused.ppm"), "64", "48", "1", str(path)],
                          env=dict(os.environ, SRT_DUMP_SCENE=prefix))
    sc = _mirror_scene(64, 48, 1, mesh)
    assert open(prefix + ".materials", "rb").read() == sc.materials.tobytes()
    assert open(prefix + ".triangles", "rb").read() == sc.triangles.tobytes()
    assert open(prefix + ".shapes", "rb").read() == sc.shapes.tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("with_mesh,png_sky", [(False, False), (True, False), (True, True)])
def test_headless_harness_matches_python_path_and_oracle(tmp_path, oracle_lib, with_mesh, png_sky):
    from simple_raytracer_b200 import tracer as tracer_mod
    from simple_raytracer_b200.tracer import Tracer
    build_harness()
    w, h, frames = 200, 120, 3
    out = tmp_path / "o.ppm"
    mesh = None
    args = [EXE, str(out), str(w), str(h), str(frames)]
    if with_mesh:  # "Add model" through load_stl_model + Model(triangles, first, count)
        path, mesh = _mesh_file(tmp_path)
        args.append(str(path))
    env = dict(os.environ)
    if png_sky:  # Tracer(width, height, "sky.png"): the facade decodes the sky box itself, like the reference's constructor
        from PIL import Image
        yy, xx = np.mgrid[0:48, 0:96]
        img = np.stack([60 + yy * 3, 90 + (xx + yy) % 120, 140 + yy * 2], -1).astype(np.uint8)
        Image.fromarray(img, "RGB").save(tmp_path / "sky.png")
        env["SRT_SKYBOX_PNG"] = str(tmp_path / "sky.png")
    subprocess.check_call(args, env=env)
    data = out.read_bytes()
    header = f"P6 {w} {h} 255\n".encode()
    assert data.startswith(header)
    got = np.frombuffer(data[len(header):], np.uint8).reshape(h, w, 3)

    # the same scene and protocol through the Python mirror
    F = np.float32
    sky = np.ones((32, 64, 4), F)
    v = ((np.arange(32, dtype=F) + F(0.5)) / F(32))[:, None]
    sky[..., 0], sky[..., 1], sky[..., 2] = F(0.25) + F(0.3) * v, F(0.35) + F(0.35) * v, F(0.5) + F(0.45) * v
    if png_sky:
        sky = tracer_mod.load_skybox_png(env["SRT_SKYBOX_PNG"])
    sc = _mirror_scene(w, h, frames, mesh)
    tr = Tracer(w, h, sky)
    tr.scene_data[:] = sc.scene_data
    tr.clear_canvas()
    tr.update_scene(sc.shapes, sc.triangles, sc.materials)
    pixels = np.zeros(w * h * 4, np.uint8)
    canvas = None
    for tick in range(frames):
        tr.options[:] = sc.render_data(tick)
        tr.render(tick + 1, pixels)
        canvas, _ = oracle_lib.render(tr.options, sc.scene_data, sc.shapes, sc.triangles, sc.materials, sky, canvas)
    assert np.array_equal(got, pixels.reshape(h, w, 4)[..., 1:])
    assert np.array_equal(got, oracle_lib.average(frames, canvas)[..., 1:])
