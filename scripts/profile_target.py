"""Short, deterministic launch sequence for ncu:
    python scripts/profile_target.py <config> [launches] [num_samples] [batch]
`batch` submits the launches as ONE srt_render_batch (one persistent render kernel + one accumulate kernel), which is
what bench.py's timed step runs; without it every launch is its own kernel pair."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_raytracer_b200 import scenes  # noqa: E402
from simple_raytracer_b200.tracer import Tracer  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sc = scenes.CONFIGS[cfg]()
ns = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3] not in ("", "-") else sc.num_samples
batch = len(sys.argv) > 4 and sys.argv[4] == "batch"
sky = scenes.procedural_skybox()
tr = Tracer(sc.width, sc.height, sky)
tr.scene_data[:] = sc.scene_data
tr.update_scene(sc.shapes, sc.triangles, sc.materials)
if os.environ.get("SRT_ACCEL") == "bvh":
    tr.set_accel("bvh")
tr.clear_canvas()
rds = [sc.render_data(k, num_samples=ns) for k in range(launches)]
if batch:
    tr.accumulate_batch(rds[:1])  # warm-up kernel pair (captures skip it)
    tr.accumulate_batch(rds)
else:
    for rd in rds:
        tr.accumulate(rd)
tr.resolve(launches)
ms, n = tr.render_time_ms()
print(f"config {cfg}: {n} launches, {ms / n:.3f} ms per launch{' (batched)' if batch else ''}")
