"""Short, deterministic launch sequence for ncu: `python scripts/profile_target.py <config> [launches] [ns]`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_raytracer_b200 import scenes  # noqa: E402
from simple_raytracer_b200.tracer import Tracer  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sc = scenes.CONFIGS[cfg]()
ns = int(sys.argv[3]) if len(sys.argv) > 3 else sc.num_samples
sky = scenes.procedural_skybox()
tr = Tracer(sc.width, sc.height, sky)
tr.scene_data[:] = sc.scene_data
tr.update_scene(sc.shapes, sc.triangles, sc.materials)
tr.clear_canvas()
for k in range(launches):
    tr.accumulate(sc.render_data(k, num_samples=ns))
tr.resolve(launches)
ms, n = tr.render_time_ms()
print(f"config {cfg}: {n} launches, {ms / n:.3f} ms per launch")
