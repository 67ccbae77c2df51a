"""Regenerates tests/golden/fullsize_hashes.json: SHA-256 of the full-BASELINE-size accumulation canvases and resolved
images produced by the REFERENCE KERNEL (oracle/_ref = /root/reference/src/render.cl compiled by g++), so that the GPU
box can check full-size parity against the reference's own output from a few bytes.  Run from the repo root in the
authoring container: python tests/golden/make_fullsize_hashes.py   (about a minute on 8 cores)."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from simple_raytracer_b200 import scenes  # noqa: E402

# (key, config, launches, num_samples, size): full resolution of the config unless a size is given (config 5 at
# full size would keep the host cores busy for ten minutes; a quarter-resolution frame still sweeps all 100 352 triangles)
CASES = [("config1", 1, 1, 1, None), ("config2", 2, 2, 4, None), ("config3", 3, 1, 1, None), ("config4", 4, 1, 1, None),
         ("config5_480x270", 5, 1, 1, (480, 270))]


def canvas_digest(canvas):
    """NaN payloads may differ between implementations: canonicalise, then hash the bytes."""
    c = np.ascontiguousarray(canvas, np.float32).copy()
    c[np.isnan(c)] = np.float32(np.nan)
    return hashlib.sha256(c.view(np.uint32).tobytes()).hexdigest()


if __name__ == "__main__":
    sky = scenes.procedural_skybox()
    out = {"sky_sha256": hashlib.sha256(sky.tobytes()).hexdigest()}
    for key, cfg, launches, ns, size in CASES:
        sc = scenes.CONFIGS[cfg](*size) if size else scenes.CONFIGS[cfg]()
        canvas = None
        for k in range(launches):
            canvas, _ = oracle.render(sc.render_data(k, num_samples=ns), sc.scene_data, sc.shapes, sc.triangles,
                                      sc.materials, sky, canvas, impl="ref")
        argb = oracle.average(launches, canvas, impl="ref")
        out[key] = {"config": cfg, "width": sc.width, "height": sc.height, "launches": launches, "num_samples": ns,
                               "inputs_sha256": hashlib.sha256(sc.shapes.tobytes() + sc.triangles.tobytes() +
                                                               sc.materials.tobytes() + sc.scene_data.tobytes()).hexdigest(),
                               "canvas_sha256": canvas_digest(canvas), "argb_sha256": hashlib.sha256(argb.tobytes()).hexdigest()}
        print(key, out[key])
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fullsize_hashes.json"), "w"), indent=1)
