"""Multi-GPU check, run under torchrun (one rank per GPU):
  * sample sharding (config-4 style), both exchange steps (reduce-scatter + per-rank resolve + gather; reduce to
    root): canvas vs the 1-GPU canvas within 1e-5 relative, image within 1 LSB;
  * tile sharding (config-5 style): ARGB8 image bit-identical to the 1-GPU image.
Prints one JSON line on rank 0; exits non-zero on mismatch."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from simple_raytracer_b200 import distributed as D  # noqa: E402
from simple_raytracer_b200 import scenes  # noqa: E402
from simple_raytracer_b200.tracer import Tracer  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sky = scenes.procedural_skybox(512, 256)
    ok = True
    report = {"world": world}

    def tracer_for(sc):
        t = Tracer(sc.width, sc.height, sky, device=local)
        t.scene_data[:] = sc.scene_data
        t.update_scene(sc.shapes, sc.triangles, sc.materials)
        return t

    # -- sample sharding
    sc = scenes.config2(640, 360, num_samples=4, launches=8)
    tr = tracer_for(sc)
    img = D.render_sample_sharded(tr, sc, rank, world)  # reduce-scatter tail: every rank owns a reduced slice
    canvas = D.gather_reduced_canvas(tr, rank, world)
    img_root = D.render_sample_sharded(tr, sc, rank, world, tail="reduce")  # reduce-to-root tail
    if rank == 0:
        report["tails_agree_lsb"] = int(np.abs(img.astype(int) - img_root.astype(int)).max())
        ok &= report["tails_agree_lsb"] <= 1
        one = tracer_for(sc)
        want_img = D.render_sample_sharded(one, sc, 0, 1)
        want = one.read_canvas()
        fin = np.isfinite(want) & np.isfinite(canvas)
        rel = np.abs(canvas - want)[fin] / np.maximum(np.abs(want)[fin], 1e-3)
        report["sample_sharded_max_rel"] = float(rel.max())
        report["sample_sharded_img_max_lsb"] = int(np.abs(img.astype(int) - want_img.astype(int)).max())
        ok &= rel.max() <= 1e-5 and report["sample_sharded_img_max_lsb"] <= 1
    # -- tile sharding
    sc = scenes.config3(480, 270, num_samples=2, launches=2)
    tr2 = tracer_for(sc)
    img_t = D.render_tile_sharded(tr2, sc, rank, world, band_height=2)
    if rank == 0:
        one = tracer_for(sc)
        want_img = D.render_tile_sharded(one, sc, 0, 1)
        report["tile_sharded_identical"] = bool(np.array_equal(img_t, want_img))
        ok &= report["tile_sharded_identical"]
        report["ok"] = bool(ok)
        print(json.dumps(report))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    return 0 if int(flag.item()) else 1


if __name__ == "__main__":
    sys.exit(main())
