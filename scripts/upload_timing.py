import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
from simple_raytracer_b200 import scenes
from simple_raytracer_b200.tracer import Tracer
for cfg in (2, 3):
    sc = scenes.CONFIGS[cfg]()
    tr = Tracer(sc.width, sc.height, scenes.procedural_skybox())
    tr.scene_data[:] = sc.scene_data
    tr.update_scene(sc.shapes, sc.triangles, sc.materials)
    tr.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        tr.update_scene(sc.shapes, sc.triangles, sc.materials)
    t1 = time.perf_counter()
    for _ in range(200):
        tr.clear_canvas()
    tr.synchronize()
    t2 = time.perf_counter()
    print(f"config {cfg}: update_scene {(t1 - t0) / 200 * 1e3:.3f} ms, clear {(t2 - t1) / 200 * 1e3:.3f} ms")
