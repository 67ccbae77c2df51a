"""Regenerates tests/golden/canvas_golden.npz from the CPU oracle (run from the repo root:
python tests/golden/make_golden.py).  The reference has no fixtures of its own and cannot run here
(no OpenCL runtime), so these are ORACLE outputs: they pin the oracle against regressions and let the
GPU box check the CUDA path without trusting a freshly built oracle.  The fixture is self-contained:
the exact input records (shapes, triangles, materials, SceneData, RenderData per launch, sky texels)
are stored next to the expected canvases, so nothing depends on numpy's SIMD dispatch on the box."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from simple_raytracer_b200 import scenes  # noqa: E402
from simple_raytracer_b200.records import RENDER_DATA, concat_records  # noqa: E402

CASES = {1: (96, 72, 2, 2), 2: (96, 54, 2, 2), 3: (64, 36, 1, 2)}

if __name__ == "__main__":
    sky = scenes.procedural_skybox(256, 128, seed=7)
    out = {"sky": sky}
    for cfg, (w, h, ns, launches) in CASES.items():
        sc = scenes.CONFIGS[cfg](w, h)
        canvas = None
        rds = concat_records(RENDER_DATA, *[sc.render_data(k, num_samples=ns) for k in range(launches)])
        for k in range(launches):
            canvas, _ = oracle.render(rds[k:k + 1], sc.scene_data, sc.shapes, sc.triangles, sc.materials,
                                      sky, canvas, threads=1)
        idx, t = oracle.primary(rds[0:1], sc.scene_data, sc.shapes, sc.triangles)
        out.update({f"c{cfg}_canvas": canvas, f"c{cfg}_argb": oracle.average(launches, canvas),
                    f"c{cfg}_primary_idx": idx, f"c{cfg}_primary_t": t,
                    f"c{cfg}_rd": rds.view(np.uint8), f"c{cfg}_shapes": sc.shapes.view(np.uint8),
                    f"c{cfg}_triangles": sc.triangles.view(np.uint8), f"c{cfg}_materials": sc.materials.view(np.uint8),
                    f"c{cfg}_scene_data": sc.scene_data.view(np.uint8)})
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "canvas_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})
