"""Developer diagnostic: why is one `render` launch slower when several ranks run side by side?

    torchrun --nproc-per-node 2 scripts/mgpu_diag.py

Every rank times the same config-2 launches (CUDA events inside the tracer) under four set-ups and prints one
line each: (A) no process group, (B) after init_process_group(nccl), (C) with the per-step canvas reduce,
(D) like A but with the time seeds rank r of N would use in bench.py."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from simple_raytracer_b200 import distributed as D  # noqa: E402
from simple_raytracer_b200 import scenes  # noqa: E402
from simple_raytracer_b200.tracer import Tracer  # noqa: E402

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
sc = scenes.config2()
sky = scenes.procedural_skybox()
tr = Tracer(sc.width, sc.height, sky, device=local)
tr.scene_data[:] = sc.scene_data
tr.update_scene(sc.shapes, sc.triangles, sc.materials)


def run(tag, rds, reduce=False, steps=4):
    for s in range(steps + 1):
        if s == 1:
            torch.cuda.synchronize()
            tr.render_time_ms()
        tr.clear_canvas()
        for rd in rds:
            tr.accumulate(rd)
        if reduce:
            D.reduce_canvas(tr, dst=0)
    torch.cuda.synchronize()
    ms, n = tr.render_time_ms()
    print(json.dumps({"rank": rank, "case": tag, "launch_ms": ms / n, "launches": n}), flush=True)


same = [sc.render_data(k) for k in range(sc.launches)]
mine = [sc.render_data(k * world + rank) for k in range(sc.launches)]
run("A no process group, seeds of N=1", same)
run("D no process group, seeds of rank r of N", mine)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dist.barrier()
    run("B process group up, no collective", same)
    run("C with canvas reduce per step", same, reduce=True)
    run("E with canvas reduce per step, seeds of rank r", mine, reduce=True)
    dist.destroy_process_group()
