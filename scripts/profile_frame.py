"""Short frame sequence through Tracer.render (srt_render_frame) for ncu:
    python scripts/profile_frame.py <config> [frames] [auto|separate|fused]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_raytracer_b200 import scenes  # noqa: E402
from simple_raytracer_b200.tracer import Tracer  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mode = sys.argv[3] if len(sys.argv) > 3 else "auto"
sc = scenes.CONFIGS[cfg]()
tr = Tracer(sc.width, sc.height, scenes.procedural_skybox())
tr.scene_data[:] = sc.scene_data
tr.update_scene(sc.shapes, sc.triangles, sc.materials)
tr.set_frame_pipeline(mode)
out = np.zeros(sc.width * sc.height * 4, np.uint8)
tr.pin_output(out)
tr.clear_canvas()
for k in range(frames):
    tr.options[:] = sc.render_data(k)
    tr.render(k + 1, out)
ms, n = tr.render_time_ms()
print(f"config {cfg} {mode}: {n} frames, {ms / n:.3f} ms of kernels per frame")
