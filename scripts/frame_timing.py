"""Developer timing of Tracer.render (srt_render_frame) per frame: wall clock and the render kernel's own duration
(CUDA events), separate steps vs the epilogue into the pinned vector (auto), BASELINE config 2 at 1080p (pinned output)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_raytracer_b200 import scenes  # noqa: E402
from simple_raytracer_b200.tracer import Tracer  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ns = int(sys.argv[2]) if len(sys.argv) > 2 else None
w, h = (1920, 1080)
sc = scenes.CONFIGS[cfg](w, h)
sky = scenes.procedural_skybox()
tr = Tracer(w, h, sky)
tr.scene_data[:] = sc.scene_data
tr.update_scene(sc.shapes, sc.triangles, sc.materials)
out = np.zeros(w * h * 4, np.uint8)
for pinned in (True, False):
    if pinned:
        tr.pin_output(out)
    for mode in ("separate", "auto", "separate", "auto"):
        tr.set_frame_pipeline(mode)
        tr.clear_canvas()
        for k in range(4):
            tr.options[:] = sc.render_data(k, num_samples=ns)
            tr.render(k + 1, out)
        tr.render_time_ms()
        n = 32
        t0 = time.perf_counter()
        for k in range(n):
            tr.options[:] = sc.render_data(k, num_samples=ns)
            tr.render(k + 1, out)
        wall = (time.perf_counter() - t0) / n * 1e3
        ms, launches = tr.render_time_ms()
        print(f"config {cfg} ns={ns or sc.num_samples} pinned={pinned} {mode:9s} wall {wall:.3f} ms/frame, kernel(s) {ms / launches:.3f} ms/frame, "
              f"{w * h * (ns or sc.num_samples) / wall / 1e3:.0f} Msamples/s", flush=True)
    if pinned:
        tr.unpin_output()
