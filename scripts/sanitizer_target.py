"""Small workload for compute-sanitizer (memcheck / racecheck / initcheck): every kernel mode, batches, row bands,
a ring-overflowing model.  usage: compute-sanitizer --tool memcheck python scripts/sanitizer_target.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from simple_raytracer_b200 import scenes  # noqa: E402
from simple_raytracer_b200.tracer import Tracer  # noqa: E402

sky = scenes.procedural_skybox(128, 64)
for cfg, (w, h) in {1: (64, 48), 2: (64, 36), 3: (48, 27), 5: (24, 14)}.items():
    sc = scenes.CONFIGS[cfg](w, h)
    tr = Tracer(w, h, sky)
    tr.scene_data[:] = sc.scene_data
    tr.update_scene(sc.shapes, sc.triangles, sc.materials)
    tr.clear_canvas()
    tr.accumulate(sc.render_data(0, num_samples=2))
    tr.accumulate_batch([sc.render_data(k, num_samples=2) for k in range(3)])
    tr.set_row_bands(1, 1, 3)
    tr.accumulate_batch([sc.render_data(k, num_samples=1) for k in range(2)])
    tr.set_row_bands(1, 0, 1)
    tr.accumulate_counted(sc.render_data(1, num_samples=1))
    tr.debug_primary(sc.render_data(0, num_samples=1))
    out = tr.resolve(6)
    print("config", cfg, "ok", int(out[..., 1:].sum()))
    tr.close()
# zero-area triangles: every pair survives the filter and overflows the pair ring
junk = np.zeros(600, scenes.TRIANGLE)
junk["v"]["pos"] = np.repeat(np.random.default_rng(1).normal(size=(600, 1, 3)).astype(np.float32) * 0.4, 3, axis=1)
junk["v"]["normal"] = (0, 0, 1)
shapes = scenes._stack([scenes.model(0, junk, 0, 600, scenes.translate((0, 0, -3)))], scenes.SHAPE)
mats = scenes._stack([scenes.material((0.8, 0.8, 0.8))], scenes.MATERIAL)
sc = scenes.Scene("junk", 32, 24, 1, 3, 1, shapes, junk, mats, scenes.camera_matrix((0, 0, 0.3)))
tr = Tracer(32, 24, sky)
tr.scene_data[:] = sc.scene_data
tr.update_scene(sc.shapes, sc.triangles, sc.materials)
tr.accumulate(sc.render_data(0))
print("junk ok", float(tr.read_canvas()[..., :3].sum()))
