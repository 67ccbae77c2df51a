"""Shared helpers for the parity tests: run the same seeded scene through the oracle and the CUDA path."""
import numpy as np


def oracle_canvas(oracle, scene, sky, launches, **rd_kw):
    canvas, total = None, None
    for k in range(launches):
        canvas, cnt = oracle.render(scene.render_data(k, **rd_kw), scene.scene_data, scene.shapes,
                                    scene.triangles, scene.materials, sky, canvas)
        total = cnt if total is None else tuple(int(a) + int(b) for a, b in zip(total, cnt))
    return canvas, np.array([int(v) for v in total], np.uint64)


def make_tracer(scene, sky, width=None, height=None):
    from simple_raytracer_b200.tracer import Tracer
    t = Tracer(width or scene.width, height or scene.height, sky)
    t.scene_data[:] = scene.scene_data
    t.update_scene(scene.shapes, scene.triangles, scene.materials)
    return t


def cuda_canvas(tracer, scene, launches, counted=False, **rd_kw):
    tracer.clear_canvas()
    cnt = None
    for k in range(launches):
        rd = scene.render_data(k, **rd_kw)
        if counted:
            cnt = tracer.accumulate_counted(rd, cnt)
        else:
            tracer.accumulate(rd)
    canvas = tracer.read_canvas()
    if counted:
        return canvas, np.array([int(cnt[0][n]) for n in cnt.dtype.names], np.uint64)
    return canvas


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return float("inf") if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def ref_primary_ids(oracle, scene, rd, sky_shape=(2, 2)):
    """Primary-hit SHAPE index of every pixel as reported by the reference's own closest_intersection
    (oracle/_ref = render.cl compiled as is).  The kernel only exposes radiance, so the id is carried by it:
    shape i gets its own material i with emission (i + 1, 0, 0), emission_strength 1, the launch runs 1 sample
    with num_bounces = 1 (render.cl:413-416: color = emission, then break) over a black sky with the sun off
    (:463-465 adds 0 on a miss).  canvas.x is then exactly i + 1 for a hit and 0 for a miss."""
    from simple_raytracer_b200 import records as R
    shapes = np.ascontiguousarray(scene.shapes, R.SHAPE).copy()
    n = len(shapes)
    shapes["material"] = np.arange(n)
    mats = np.zeros(max(n, 1), R.MATERIAL)
    mats["emission"][:, 0] = np.arange(1, len(mats) + 1)
    mats["emission_strength"] = 1.0
    mats["refraction_index"] = 1.0
    sd = np.ascontiguousarray(scene.scene_data, R.SCENE_DATA).copy()
    sd["sun_intensity"] = 0.0
    rd = np.ascontiguousarray(rd, R.RENDER_DATA).copy()
    assert int(rd["num_samples"].reshape(-1)[0]) == 1, "the seed of sample 0 depends on num_samples (render.cl:496)"
    rd["num_bounces"], rd["show_normals"] = 1, 0
    black = np.zeros(sky_shape + (4,), np.float32)
    canvas, _ = oracle.render(rd, sd, shapes, scene.triangles, mats, black, impl="ref")
    ids = canvas[..., 0]
    assert np.array_equal(ids, np.rint(ids)) and not canvas[..., 1:3].any()
    return ids.astype(np.int32) - 1


def random_scene(seed, width=96, height=64, n_spheres=5, n_planes=3, n_boxes=2, mesh_tris=0):
    """Seeded scene of random spheres / planes / box models (+ an optional random triangle soup) with random
    materials covering every scatter branch (diffuse, metallic, specular, transparent with both ior < 1 and > 1,
    emissive)."""
    from simple_raytracer_b200 import scenes
    from simple_raytracer_b200.records import concat_records
    rng = np.random.default_rng(seed)
    mats = [scenes.material(rng.uniform(0.2, 1.0, 3), smoothness=rng.uniform(0, 1), metallic=rng.choice([0, 0.5, 1.0]),
                            specular=rng.choice([0, 0.3, 1.0]), transmittance=rng.choice([0, 0, 0.6, 1.0]),
                            refraction_index=rng.choice([0.8, 1.0, 1.33, 1.5, 2.4]),
                            emission=rng.uniform(0, 1, 3), emission_strength=rng.choice([0, 0, 0, 4.0]))
            for _ in range(6)]
    tri_parts = [scenes.cube_triangles()]
    if mesh_tris:
        pos = (rng.normal(size=(mesh_tris, 1, 3)) * 0.8 + rng.normal(size=(mesh_tris, 3, 3)) * 0.4).astype(np.float32)
        nrm = rng.normal(size=(mesh_tris, 3, 3)).astype(np.float32)
        nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
        tri_parts.append(scenes.triangles_from(pos, nrm))
    tris = concat_records(scenes.TRIANGLE, *tri_parts)
    shapes = []
    for _ in range(n_planes):
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        shapes.append(scenes.plane(int(rng.integers(6)), -n * rng.uniform(2.5, 5.0), n))
    for _ in range(n_spheres):
        shapes.append(scenes.sphere(int(rng.integers(6)), rng.uniform(-2.5, 2.5, 3) + (0, 0, -3), rng.uniform(0.3, 1.4)))
    for _ in range(n_boxes):
        xf = scenes.translate(rng.uniform(-2, 2, 3) + (0, 0, -3)) @ scenes.rotate_y(rng.uniform(0, 3)) @ \
            scenes.rotate_x(rng.uniform(0, 3)) @ scenes.scale(rng.uniform(0.3, 1.0, 3))
        shapes.append(scenes.model(int(rng.integers(6)), tris, 0, 12, xf))
    if mesh_tris:
        shapes.append(scenes.model(int(rng.integers(6)), tris, 12, mesh_tris, scenes.translate((0, 0, -3))))
    order = rng.permutation(len(shapes))
    shapes = [shapes[i] for i in order]
    return scenes.Scene(f"random{seed}", width, height, 2, 7, 1, scenes._stack(shapes, scenes.SHAPE), tris,
                        scenes._stack(mats, scenes.MATERIAL),
                        scenes.camera_matrix((0, 0.2, 3.0), rng.uniform(-0.3, 0.3), rng.uniform(-0.2, 0.2)))


def reference_parse(kind, v, n, f, tris):
    """What reference src/parser.cpp produces from a mesh file written by test_mesh_io.write_stl / write_obj,
    restated in numpy.  STL (:17-53): the facet normal is copied unnormalised to all three vertices, positions as
    stored.  OBJ (:55-135): `vn` is glm::normalize'd at load (:83) = v * (1 / sqrt(x*x + y*y + z*z)) in float32."""
    from simple_raytracer_b200 import scenes
    if kind == "stl":
        return tris
    nn = np.asarray(n, np.float32)
    d = (nn[:, 0] * nn[:, 0] + nn[:, 1] * nn[:, 1]) + nn[:, 2] * nn[:, 2]
    inv = (np.float32(1.0) / np.sqrt(d)).astype(np.float32)
    return scenes.mesh_triangles(v, f, (nn * inv[:, None]).astype(np.float32))
