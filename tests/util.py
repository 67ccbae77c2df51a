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
