"""Regenerates tests/golden/canvas_golden.npz by RUNNING THE REFERENCE KERNEL (run from the repo root, in
the authoring container where /root/reference exists: python tests/golden/make_golden.py).

The reference has no fixtures of its own.  The canvases, the resolved ARGB8 images and the primary-hit shape
ids stored here are outputs of oracle/_ref, i.e. of /root/reference/src/render.cl itself compiled by g++
(oracle/ref_build/); only `primary_t`, which the kernel cannot expose, comes from oracle.c (which this script
first checks to be bit-identical to the reference kernel on the same inputs).  The fixture is self-contained:
the exact input records (shapes, triangles, materials, SceneData, RenderData per launch, sky texels) are
stored next to the expected outputs, so the GPU box can check both the oracle and the CUDA path against the
reference's results without /root/reference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from simple_raytracer_b200 import scenes  # noqa: E402
from util import ref_primary_ids  # noqa: E402
from simple_raytracer_b200.records import RENDER_DATA, concat_records  # noqa: E402

CASES = {1: (96, 72, 2, 2), 2: (96, 54, 2, 2), 3: (64, 36, 1, 2)}

if __name__ == "__main__":
    sky = scenes.procedural_skybox(256, 128, seed=7)
    out = {"sky": sky}
    for cfg, (w, h, ns, launches) in CASES.items():
        sc = scenes.CONFIGS[cfg](w, h)
        canvas, check = None, None
        rds = concat_records(RENDER_DATA, *[sc.render_data(k, num_samples=ns) for k in range(launches)])
        for k in range(launches):
            canvas, _ = oracle.render(rds[k:k + 1], sc.scene_data, sc.shapes, sc.triangles, sc.materials,
                                      sky, canvas, threads=1, impl="ref")
            check, _ = oracle.render(rds[k:k + 1], sc.scene_data, sc.shapes, sc.triangles, sc.materials,
                                     sky, check, threads=1)
        assert np.array_equal(canvas.view(np.uint32), check.view(np.uint32)), "oracle.c differs from render.cl"
        prd = sc.render_data(0, num_samples=1)  # the seed of sample 0 depends on num_samples (render.cl:496)
        idx = ref_primary_ids(oracle, sc, prd)
        oidx, t = oracle.primary(prd, sc.scene_data, sc.shapes, sc.triangles)
        assert np.array_equal(idx, oidx)
        out.update({f"c{cfg}_canvas": canvas, f"c{cfg}_argb": oracle.average(launches, canvas, impl="ref"),
                    f"c{cfg}_primary_idx": idx, f"c{cfg}_primary_t": t,
                    f"c{cfg}_rd": rds.view(np.uint8), f"c{cfg}_primary_rd": prd.view(np.uint8), f"c{cfg}_shapes": sc.shapes.view(np.uint8),
                    f"c{cfg}_triangles": sc.triangles.view(np.uint8), f"c{cfg}_materials": sc.materials.view(np.uint8),
                    f"c{cfg}_scene_data": sc.scene_data.view(np.uint8)})
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "canvas_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})
