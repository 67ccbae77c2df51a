import os

import numpy as np

from simple_raytracer_b200 import records as R

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "canvas_golden.npz")
CONFIGS = (1, 2, 3)


def load(cfg):
    g = np.load(PATH)
    return dict(sky=g["sky"], canvas=g[f"c{cfg}_canvas"], argb=g[f"c{cfg}_argb"],
                primary_idx=g[f"c{cfg}_primary_idx"], primary_t=g[f"c{cfg}_primary_t"],
                rd=g[f"c{cfg}_rd"].view(R.RENDER_DATA), primary_rd=g[f"c{cfg}_primary_rd"].view(R.RENDER_DATA), shapes=g[f"c{cfg}_shapes"].view(R.SHAPE),
                triangles=g[f"c{cfg}_triangles"].view(R.TRIANGLE), materials=g[f"c{cfg}_materials"].view(R.MATERIAL),
                scene_data=g[f"c{cfg}_scene_data"].view(R.SCENE_DATA))
