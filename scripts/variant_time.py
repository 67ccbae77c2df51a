"""Developer experiment helper (GPU box): kernel time of configs through the library named by SRT_LIB.
    SRT_LIB=build/variants/libsrt_x.so python scripts/variant_time.py 2 3 5
Prints one line per config: launches are submitted as one batch, best of 3 repetitions, CUDA-event kernel time."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_raytracer_b200 import scenes  # noqa: E402
from simple_raytracer_b200.tracer import Tracer  # noqa: E402

sky = scenes.procedural_skybox()
for cfg in [int(a) for a in sys.argv[1:]] or [2, 3, 5]:
    sc = scenes.CONFIGS[cfg]()
    launches = {1: 1, 2: 16, 3: 8, 4: 2, 5: 1}[cfg]
    tr = Tracer(sc.width, sc.height, sky)
    tr.scene_data[:] = sc.scene_data
    tr.update_scene(sc.shapes, sc.triangles, sc.materials)
    if os.environ.get("SRT_ACCEL") == "bvh":
        tr.set_accel("bvh")
    rds = [sc.render_data(k) for k in range(launches)]
    best = 1e30
    for rep in range(4):
        tr.clear_canvas()
        tr.accumulate_batch(rds)
        ms, n = tr.render_time_ms()
        if rep:
            best = min(best, ms)
    samples = sc.width * sc.height * sc.num_samples * launches
    digest = hashlib.sha256(tr.read_canvas().tobytes()).hexdigest()[:12]  # bit-exact builds agree on this
    print(f"cfg {cfg}  kernel_ms {best:9.3f}  per-launch {best / launches:8.3f}  Msamples/s {samples / (best * 1e-3) / 1e6:9.1f}  canvas {digest}", flush=True)
    tr.close()
