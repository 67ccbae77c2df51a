"""Known-answer vectors for the integer part of the path (render.cl:143-148, :496) and the record
layouts (shape.hpp / material.hpp / tracer.hpp == render.cl:5-105).  The reference has no tests; these
vectors were derived from its source in SURVEY.md 8c and are implementation independent."""
import numpy as np

from simple_raytracer_b200 import records as R


def test_random_float_kat(oracle_lib):
    got = oracle_lib.random_floats(0, 4)
    assert [(s, h) for s, h, _ in got] == [(0xAC564B05, 0x07BB2FE2), (0x4712A88E, 0x22B6B6BC),
                                           (0x21DD796B, 0x3BF6E0B1), (0x3463E0AC, 0x572F7439)]
    want = [0.030199997, 0.13560049, 0.23423581, 0.34056783]
    assert np.allclose([f for _, _, f in got], want, rtol=0, atol=5e-9)
    assert [h for _, h, _ in oracle_lib.random_floats(1, 4)] == [0xA8BEEA3C, 0x2679C518, 0x97AAF6C6, 0x2A521372]
    assert [h for _, h, _ in oracle_lib.random_floats(0x14B8, 4)] == [0x0F36D52D, 0x56FE22D5, 0xBEFC22FD, 0x639151DD]
    assert [h for _, h, _ in oracle_lib.random_floats(0xDEADBEEF, 4)] == [0x67299972, 0x3779382A, 0x50E322B3, 0x91A0E66E]


def test_random_float_is_hash_times_2_pow_minus_32(oracle_lib):
    for s0 in (0, 1, 12345, 0xFFFFFFFF, 0x80000000):
        for _, h, f in oracle_lib.random_floats(s0, 16):
            assert np.float32(f) == np.float32(h) / np.float32(4294967296.0)
            assert 0.0 <= f <= 1.0  # inclusive: (float)r can round up to 2^32


def test_seed_formula_kat(oracle_lib):
    kat = {(0, 0, 1, 1): 0x0, (0, 1, 1, 1): 0x14B8, (0, 1, 1, 12345): 0x03E71CF8,
           (3, 479999, 4, 12345): 0xE0CA8308, (63, 2073599, 64, 1000): 0x8C6F1140}
    for args, want in kat.items():
        assert oracle_lib.seed(*args) == want
    # pure uint32 wrap-around arithmetic
    for sample, pid, ns, time in [(5, 8294399, 16, 1000067), (31, 123456, 32, 0xFFFFFFFF)]:
        assert oracle_lib.seed(sample, pid, ns, time) == ((sample + pid * ns) * time * 5304) & 0xFFFFFFFF


def test_record_layouts():
    def off(dt, name):
        return dt.fields[name][1]
    assert R.MATERIAL.itemsize == 64
    assert [off(R.MATERIAL, n) for n in ("smoothness", "metallic", "specular", "emission_strength",
                                         "transmittance", "refraction_index", "color", "emission")] == \
        [0, 4, 8, 12, 16, 20, 32, 48]
    assert R.TRIANGLE.itemsize == 96 and R.VERTEX.itemsize == 32 and off(R.VERTEX, "pos") == 16
    assert R.SHAPE.itemsize == 128
    assert off(R.SHAPE, "type") == 0 and off(R.SHAPE, "material") == 4
    assert off(R.SHAPE, "sphere_position") == 16 and off(R.SHAPE, "sphere_radius") == 32
    assert off(R.SHAPE, "plane_position") == 16 and off(R.SHAPE, "plane_normal") == 32
    # Model @16: triangle_index @0, num @4, bmin @16, bmax @32, transform @48
    assert [off(R.SHAPE, n) for n in ("model_triangle_index", "model_num_triangles", "model_bounding_min",
                                      "model_bounding_max", "model_transform")] == [16, 20, 32, 48, 64]
    assert R.RENDER_DATA.itemsize == 112
    assert [off(R.RENDER_DATA, n) for n in ("width", "height", "num_samples", "num_bounces", "aspect_ratio",
                                            "fov_scale", "show_normals", "camera_to_world", "time", "tick")] == \
        [0, 4, 8, 12, 16, 20, 24, 32, 96, 100]
    assert R.SCENE_DATA.itemsize == 96
    assert [off(R.SCENE_DATA, n) for n in ("num_shapes", "sun_focus", "sun_intensity", "horizon_color",
                                           "zenith_color", "ground_color", "sun_color", "sun_direction")] == \
        [0, 4, 8, 16, 32, 48, 64, 80]
