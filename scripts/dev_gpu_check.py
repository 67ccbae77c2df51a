"""Developer scratch run on the GPU box: environment probe + quick timings of every config."""
import json
import os
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_raytracer_b200 import scenes  # noqa: E402
from simple_raytracer_b200.tracer import Tracer  # noqa: E402


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=60).stdout.strip()
    except Exception as e:  # noqa: BLE001
        return f"<{e}>"


def main():
    print("== probe")
    print(sh("nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv"))
    print(sh("nproc; lscpu | grep -E 'Model name|Flags' | cut -c1-200"))
    print("opencl:", sh("ls /etc/OpenCL/vendors 2>&1; ldconfig -p | grep -i -E 'opencl|pocl' ; ls /usr/lib/x86_64-linux-gnu | grep -i -E 'opencl|nvidia-compiler|nvvm' | head"))
    sky = scenes.procedural_skybox()
    which = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 5]
    for cfg in which:
        sc = scenes.CONFIGS[cfg]()
        launches = {1: 1, 2: 16, 3: 8, 4: 2, 5: 1}[cfg]
        ns = sc.num_samples if cfg != 5 else 1
        tr = Tracer(sc.width, sc.height, sky)
        tr.scene_data[:] = sc.scene_data
        tr.update_scene(sc.shapes, sc.triangles, sc.materials)
        if cfg == which[0]:
            print("fp32 peak TFLOP/s, est MHz:", tr.measure_fp32_peak())
        for rep in range(3):
            tr.clear_canvas()
            tr.synchronize()
            t0 = time.perf_counter()
            for k in range(launches):
                tr.accumulate(sc.render_data(k, num_samples=ns))
            tr.synchronize()
            dt = time.perf_counter() - t0
            ms, n = tr.render_time_ms()
        samples = sc.width * sc.height * ns * launches
        cnt = tr.accumulate_counted(sc.render_data(0, num_samples=ns))[0]
        names = cnt.dtype.names
        c = {n_: int(cnt[n_]) for n_ in names}
        print(json.dumps({"cfg": cfg, "name": sc.name, "launches": launches, "ns": ns, "wall_ms": dt * 1e3,
                          "kernel_ms": ms, "Msamples/s": samples / (ms * 1e-3) / 1e6, "counters_1launch": c,
                          "bounces/sample": c["bounces"] / max(c["samples"], 1),
                          "Gtests/s": c["tri_tests"] * launches / (ms * 1e-3) / 1e9}))
        out = tr.resolve(launches)
        os.makedirs("gpurun_out", exist_ok=True)
        from PIL import Image
        Image.fromarray(out[..., 1:]).save(f"gpurun_out/c{cfg}.png")
        tr.close()


if __name__ == "__main__":
    main()
