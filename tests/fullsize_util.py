import hashlib
import json
import os

import numpy as np

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fullsize_hashes.json")


def load():
    return json.load(open(PATH))


def canvas_digest(canvas):
    c = np.ascontiguousarray(canvas, np.float32).copy()
    c[np.isnan(c)] = np.float32(np.nan)
    return hashlib.sha256(c.view(np.uint32).tobytes()).hexdigest()


def scene_for(entry):
    from simple_raytracer_b200 import scenes
    sc = scenes.CONFIGS[entry["config"]](entry["width"], entry["height"])
    digest = hashlib.sha256(sc.shapes.tobytes() + sc.triangles.tobytes() + sc.materials.tobytes() +
                            sc.scene_data.tobytes()).hexdigest()
    return sc, digest == entry["inputs_sha256"]
