"""STL / OBJ ingest and PPM output (reference src/parser.cpp) -- host code, no GPU needed."""
import os
import struct

import numpy as np
import pytest

from simple_raytracer_b200 import scenes, tracer
from simple_raytracer_b200.records import TRIANGLE


def write_stl(path, tris):
    with open(path, "wb") as f:
        f.write(b"synthetic".ljust(80, b"\0"))
        f.write(struct.pack("<I", len(tris)))
        for t in tris:
            f.write(np.asarray(t["v"]["normal"][0], "<f4").tobytes())
            for j in range(3):
                f.write(np.asarray(t["v"]["pos"][j], "<f4").tobytes())
            f.write(b"\0\0")


def write_obj(path, v, n, f, style="//"):
    with open(path, "w") as fh:
        fh.write("# synthetic\no mesh\n")
        for p in v:
            fh.write("v %.9g %.9g %.9g\n" % tuple(p))
        for p in n:
            fh.write("vn %.9g %.9g %.9g\n" % tuple(p))
        fh.write("s 1\n")
        for a, b, c in f + 1:
            if style == "//":
                fh.write(f"f {a}//{a} {b}//{b} {c}//{c}\n")
            else:
                fh.write(f"f {a}/1/{a} {b}/1/{b} {c}/1/{c}\n")


def test_stl_round_trip(tmp_path):
    v, f = scenes.displaced_torus(12, 10, seed=1)
    tris = scenes.mesh_triangles(v, f, None)
    p = tmp_path / "m.stl"
    write_stl(p, tris)
    pre = scenes.cube_triangles()
    (first, count), allt = tracer.load_stl_model(str(p), pre)
    assert (first, count) == (12, 240) and len(allt) == 252
    assert allt[:12].tobytes() == pre.tobytes()
    got = allt[12:]
    assert np.array_equal(got["v"]["pos"], tris["v"]["pos"])
    # facet normal copied to all three vertices, not re-normalised (parser.cpp:46-50)
    assert np.array_equal(got["v"]["normal"], tris["v"]["normal"])
    assert got.dtype == TRIANGLE and got.itemsize == 96


@pytest.mark.parametrize("style", ["//", "/t/"])
def test_obj_round_trip(tmp_path, style):
    v, n, f = scenes.noisy_icosphere(1, seed=2)
    p = tmp_path / "m.obj"
    write_obj(p, v, n, f, style)
    (first, count), got = tracer.load_obj_model(str(p), np.zeros(0, TRIANGLE))
    assert (first, count) == (0, 80)
    want = scenes.mesh_triangles(v, f, n)
    assert np.array_equal(got["v"]["pos"], want["v"]["pos"])
    assert np.allclose(got["v"]["normal"], want["v"]["normal"], atol=2e-7)  # vn is re-normalised at load (:83)
    assert np.allclose(np.linalg.norm(got["v"]["normal"], axis=-1), 1.0, atol=1e-6)


def test_loaders_reproduce_parser_cpp_byte_for_byte(tmp_path):
    """The loaders' Triangle records equal what reference src/parser.cpp builds from the same files, byte for byte:
    unnormalised STL facet normals kept, unnormalised OBJ `vn` normalised with glm::normalize's operations."""
    from util import reference_parse
    v, f = scenes.displaced_torus(16, 12, seed=4)
    tris = scenes.mesh_triangles(v, f, None)
    tris["v"]["normal"] *= np.float32(2.5)
    write_stl(tmp_path / "a.stl", tris)
    (_, count), got = tracer.load_stl_model(str(tmp_path / "a.stl"), np.zeros(0, TRIANGLE))
    assert count == len(tris) and got.tobytes() == reference_parse("stl", v, None, f, tris).tobytes()
    v, n, f = scenes.noisy_icosphere(2, seed=5)
    n = (n * np.float32(0.3)).astype(np.float32)
    write_obj(tmp_path / "a.obj", v, n, f)
    (_, count), got = tracer.load_obj_model(str(tmp_path / "a.obj"), np.zeros(0, TRIANGLE))
    assert count == len(f) and got.tobytes() == reference_parse("obj", v, n, f, None).tobytes()


def test_stl_header_count_beyond_the_file_is_refused(tmp_path):
    p = tmp_path / "huge.stl"
    p.write_bytes(b"\0" * 80 + struct.pack("<I", 0xFFFFFFF0) + b"\0" * 100)
    assert tracer.load_stl_model(str(p), np.zeros(0, TRIANGLE)) is None


def test_model_bounds_enclose_the_unfused_device_transform():
    """srt_model_bounds uses the device pre-transform's separately rounded products and sums (render.cl:114-120 order),
    so the box is exactly the min / max of the world-space vertices the kernel intersects."""
    v, n, f = scenes.noisy_icosphere(2, seed=6)
    tris = scenes.mesh_triangles(v, f, n)
    xf = (scenes.translate((1.3, -2.1, 3.7)) @ scenes.rotate_y(0.41) @ scenes.rotate_x(1.1) @ scenes.scale((2.3, 0.7, 0.51))).astype(np.float32)
    rec = tracer.model_bounds(scenes.model(0, tris, 0, len(tris), xf), tris)
    p = tris["v"]["pos"].reshape(-1, 3).astype(np.float32)
    w = np.stack([((xf[r, 0] * p[:, 0] + xf[r, 1] * p[:, 1]) + xf[r, 2] * p[:, 2]) + xf[r, 3] * np.float32(1) for r in range(3)], 1)
    assert np.array_equal(rec["model_bounding_min"], w.min(0)) and np.array_equal(rec["model_bounding_max"], w.max(0))


def test_obj_negative_indices_quads_and_missing_normals(tmp_path):
    p = tmp_path / "q.obj"
    p.write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvn 0 0 2\nf -4//-1 -3//-1 -2//-1 -1//-1\nf 1 2 3\n")
    (_, count), t = tracer.load_obj_model(str(p), np.zeros(0, TRIANGLE))
    assert count == 3                                           # quad fan-triangulated + one triangle
    assert np.array_equal(t["v"]["pos"][0], [[0, 0, 0], [1, 0, 0], [1, 1, 0]])
    assert np.array_equal(t["v"]["pos"][1], [[0, 0, 0], [1, 1, 0], [0, 1, 0]])
    assert np.array_equal(t["v"]["normal"][0], [[0, 0, 1]] * 3)  # normalised vn
    assert np.array_equal(t["v"]["normal"][2], [[0, 0, 1]] * 3)  # flat geometric normal


def test_missing_and_malformed_files(tmp_path):
    assert tracer.load_stl_model(str(tmp_path / "nope.stl"), np.zeros(0, TRIANGLE)) is None  # nullopt
    assert tracer.load_obj_model(str(tmp_path / "nope.obj"), np.zeros(0, TRIANGLE)) is None
    p = tmp_path / "short.stl"
    p.write_bytes(b"\0" * 80 + struct.pack("<I", 5) + b"\0" * 60)
    assert tracer.load_stl_model(str(p), np.zeros(0, TRIANGLE)) is None
    q = tmp_path / "bad.obj"
    q.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 9\n")
    assert tracer.load_obj_model(str(q), np.zeros(0, TRIANGLE)) is None


def test_save_ppm(tmp_path):
    px = np.arange(2 * 3 * 4, dtype=np.uint8).reshape(2, 3, 4)
    p = tmp_path / "o.ppm"
    tracer.save_ppm(str(p), px, 3, 2)
    data = p.read_bytes()
    assert data.startswith(b"P6 3 2 255\n")
    assert data[len(b"P6 3 2 255\n"):] == px[..., 1:].tobytes()  # bytes 1..3 of A,R,G,B (parser.cpp:10-14)


def test_model_bounds_matches_scene_builder():
    v, n, f = scenes.noisy_icosphere(2, seed=3)
    tris = scenes.mesh_triangles(v, f, n)
    xf = scenes.translate((1, 2, 3)) @ scenes.rotate_y(0.4) @ scenes.scale((2, 1, 0.5))
    rec = scenes.model(0, tris, 0, len(tris), xf)
    out = tracer.model_bounds(rec, tris)
    assert np.allclose(out["model_bounding_min"], rec["model_bounding_min"], atol=1e-5)
    assert np.allclose(out["model_bounding_max"], rec["model_bounding_max"], atol=1e-5)
    assert (out["model_bounding_max"] > out["model_bounding_min"]).all()
