"""Developer stress test of the dense triangle sweep's tile ring: a FRESH handle per repetition renders the same three
launches of a mesh scene; every canvas must be the same bytes.  (A tile stage that is refilled while its loads are
still in flight shows up as rare missed hits.)   python scripts/sweep_race_hunt.py <config> <w> <h> <ns> <reps>"""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_raytracer_b200 import scenes  # noqa: E402
from simple_raytracer_b200.tracer import Tracer  # noqa: E402

cfg, w, h, ns, reps = (int(x) for x in sys.argv[1:6])
sc = scenes.CONFIGS[cfg](w, h)
sky = scenes.procedural_skybox(256, 128)
seen = {}
for r in range(reps):
    tr = Tracer(w, h, sky)
    tr.scene_data[:] = sc.scene_data
    tr.update_scene(sc.shapes, sc.triangles, sc.materials)
    tr.clear_canvas()
    for k in range(3):
        tr.accumulate(sc.render_data(k, num_samples=ns))
    d = hashlib.sha256(tr.read_canvas().tobytes()).hexdigest()[:12]
    seen[d] = seen.get(d, 0) + 1
    tr.close()
print(f"config {cfg} {w}x{h} ns={ns}: {reps} fresh handles -> {len(seen)} distinct canvases {sorted(seen.values(), reverse=True)[:6]}")
