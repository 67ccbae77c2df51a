"""Host-side mirror of the reference's scene / kernel-argument records as numpy dtypes.

Byte layouts follow /root/reference/include/shape.hpp:15-111, include/material.hpp:10-37,
include/tracer.hpp:48-80 (host) == src/render.cl:17-105 (device): every glm::vec3 is
`alignas(cl_float3)` = 16 bytes.  The C declaration of the same records is include/srt.h.
"""
import numpy as np

SHAPE_SPHERE, SHAPE_PLANE, SHAPE_MODEL = 0, 1, 2  # render.cl:63-67

MATERIAL = np.dtype({
    "names": ["smoothness", "metallic", "specular", "emission_strength", "transmittance",
              "refraction_index", "color", "emission"],
    "formats": ["f4", "f4", "f4", "f4", "f4", "f4", ("f4", 3), ("f4", 3)],
    "offsets": [0, 4, 8, 12, 16, 20, 32, 48],
    "itemsize": 64,
})

VERTEX = np.dtype({"names": ["normal", "pos"], "formats": [("f4", 3), ("f4", 3)],
                   "offsets": [0, 16], "itemsize": 32})
TRIANGLE = np.dtype([("v", VERTEX, 3)])  # 96 B

# Shape = {type @0, material @4, union @16}; union members overlaid with explicit offsets.
SHAPE = np.dtype({
    "names": ["type", "material",
              "sphere_position", "sphere_radius",
              "plane_position", "plane_normal",
              "model_triangle_index", "model_num_triangles", "model_bounding_min",
              "model_bounding_max", "model_transform"],
    "formats": ["i4", "i4",
                ("f4", 3), "f4",
                ("f4", 3), ("f4", 3),
                "u4", "u4", ("f4", 3), ("f4", 3), ("f4", (4, 4))],
    "offsets": [0, 4,
                16, 32,
                16, 32,
                16, 20, 32, 48, 64],
    "itemsize": 128,
})

# camera_to_world is column-major (glm::mat4): camera_to_world[c] is column c.
RENDER_DATA = np.dtype({
    "names": ["width", "height", "num_samples", "num_bounces", "aspect_ratio", "fov_scale",
              "show_normals", "camera_to_world", "time", "tick"],
    "formats": ["i4", "i4", "i4", "i4", "f4", "f4", "u1", ("f4", (4, 4)), "u4", "u4"],
    "offsets": [0, 4, 8, 12, 16, 20, 24, 32, 96, 100],
    "itemsize": 112,
})

SCENE_DATA = np.dtype({
    "names": ["num_shapes", "sun_focus", "sun_intensity", "horizon_color", "zenith_color",
              "ground_color", "sun_color", "sun_direction"],
    "formats": ["i4", "f4", "f4", ("f4", 3), ("f4", 3), ("f4", 3), ("f4", 3), ("f4", 3)],
    "offsets": [0, 4, 8, 16, 32, 48, 64, 80],
    "itemsize": 96,
})

# Work counters shared by the oracle (OracleCounters) and the C-ABI (srt_counters).
COUNTERS = np.dtype([("samples", "u8"), ("bounces", "u8"), ("tri_tests", "u8"),
                     ("aabb_pass", "u8"), ("hits", "u8"), ("sky", "u8")])

assert MATERIAL.itemsize == 64 and TRIANGLE.itemsize == 96 and SHAPE.itemsize == 128
assert RENDER_DATA.itemsize == 112 and SCENE_DATA.itemsize == 96


def concat_records(dtype, *arrays):
    """Concatenate record arrays WITHOUT np.concatenate, which re-packs padded struct dtypes."""
    arrays = [np.ascontiguousarray(a, dtype).reshape(-1) for a in arrays]
    out = np.zeros(sum(len(a) for a in arrays), dtype)
    at = 0
    for a in arrays:
        out[at:at + len(a)] = a
        at += len(a)
    return out
