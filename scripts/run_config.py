"""Run one BASELINE.json config end to end on N GPUs and print one JSON line (rank 0).

    python scripts/run_config.py --config 3                      # 1 GPU
    torchrun --nproc-per-node 8 scripts/run_config.py --config 5 # 8 GPUs

Sharding: configs 1-4 sample-sharded (launches dealt round-robin, NCCL sum-reduce of the canvases),
config 5 tile-sharded (interleaved 8-row bands, MAX-reduce of the ARGB8 images) -- BASELINE.json configs[3], [4].
Total work is fixed (strong scaling).  Reports Msamples/s, counted algorithmic flops and % of FP32 peak.
--write-ppm saves the image; --mesh-files round-trips the meshes of configs 3 / 5 through .obj / .stl files and the
library's loaders first (the data path BASELINE.json names)."""
import argparse
import json
import os
import struct
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import algorithmic_flops  # noqa: E402
from simple_raytracer_b200 import distributed as D  # noqa: E402
from simple_raytracer_b200 import scenes, tracer  # noqa: E402


def meshes_through_files(cfg, sc):
    """Write the config's mesh as .obj (config 3, smooth normals) or binary .stl (config 5), load it back with
    srt_load_obj / srt_load_stl, and rebuild the scene on the loaded triangles."""
    t = sc.triangles
    with tempfile.TemporaryDirectory() as tmp:
        if cfg == 3:
            path = os.path.join(tmp, "mesh.obj")
            with open(path, "w") as f:
                pos, nrm = t["v"]["pos"].reshape(-1, 3), t["v"]["normal"].reshape(-1, 3)
                for p in pos:
                    f.write("v %.9g %.9g %.9g\n" % tuple(p))
                for n in nrm:
                    f.write("vn %.9g %.9g %.9g\n" % tuple(n))
                for i in range(len(t)):
                    a = 3 * i + 1
                    f.write(f"f {a}//{a} {a + 1}//{a + 1} {a + 2}//{a + 2}\n")
            (first, count), loaded = tracer.load_obj_model(path, np.zeros(0, scenes.TRIANGLE))
        else:
            path = os.path.join(tmp, "mesh.stl")
            with open(path, "wb") as f:
                f.write(b"srt".ljust(80, b"\0") + struct.pack("<I", len(t)))
                for tri in t:
                    f.write(np.asarray(tri["v"]["normal"][0], "<f4").tobytes())
                    f.write(np.asarray(tri["v"]["pos"], "<f4").tobytes())
                    f.write(b"\0\0")
            (first, count), loaded = tracer.load_stl_model(path, np.zeros(0, scenes.TRIANGLE))
    assert count == len(t) and np.array_equal(loaded["v"]["pos"], t["v"]["pos"])
    sc.triangles = loaded
    return sc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True)
    ap.add_argument("--launches", type=int, default=0, help="override the launch count (0 = the config's)")
    ap.add_argument("--write-ppm", default="")
    ap.add_argument("--mesh-files", action="store_true")
    ap.add_argument("--band-height", type=int, default=1, help="rows per interleaved band when tile-sharding")
    ap.add_argument("--sharding", default="", choices=["", "tile", "sample"],
                    help="override the config's sharding (config 5 = tile, the others = sample)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sc = scenes.CONFIGS[args.config]()
    if args.mesh_files and args.config in (3, 5):
        sc = meshes_through_files(args.config, sc)
    launches = args.launches or sc.launches
    sky = scenes.procedural_skybox()
    tr = tracer.Tracer(sc.width, sc.height, sky, device=local)
    tr.scene_data[:] = sc.scene_data
    tr.update_scene(sc.shapes, sc.triangles, sc.materials)
    mode = args.sharding or ("tile" if args.config == 5 else "sample")

    def run():
        if mode == "tile":
            return D.render_tile_sharded(tr, sc, rank, world, band_height=args.band_height, total_launches=launches)
        return D.render_sample_sharded(tr, sc, rank, world, total_launches=launches)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up on one launch (and the batch scratch allocated ahead of the clock), then the timed full run
    tr.reserve_batch(sc.render_data(0), -(-launches // world) if mode == "sample" else launches)
    if mode == "tile":
        D.render_tile_sharded(tr, sc, rank, world, band_height=args.band_height, total_launches=1)
    else:
        D.render_sample_sharded(tr, sc, rank, world, total_launches=world)
    sync()
    tr.render_time_ms()
    t0 = time.perf_counter()
    img = run()
    sync()
    dt = time.perf_counter() - t0
    kernel_ms, n_launch = tr.render_time_ms()
    tmax = torch.tensor([dt, kernel_ms], dtype=torch.float64, device="cuda")
    tmin = tmax.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    # counted work of one launch (rank 0, full frame) for the flop figure
    if rank == 0:
        tr.set_row_bands(1, 0, 1)
        cnt = tr.accumulate_counted(sc.render_data(0))
        counters = {n: int(cnt[0][n]) for n in cnt.dtype.names}
        flops = algorithmic_flops(sc, counters, sc.width * sc.height) * launches
        peak, _ = tr.measure_fp32_peak()
        samples = sc.width * sc.height * sc.num_samples * launches
        secs = float(tmax[0].item())
        out = {"config": args.config, "name": sc.name, "n_gpus": world, "sharding": mode if world > 1 else "none",
               "resolution": f"{sc.width}x{sc.height}", "spp": sc.num_samples * launches, "num_bounces": sc.num_bounces,
               "triangles": int(len(sc.triangles)), "seconds": secs, "max_rank_kernel_ms": float(tmax[1].item()),
               "min_rank_kernel_ms": float(tmin[1].item()), "band_height": args.band_height if mode == "tile" else None,
               "Msamples_per_s": samples / secs / 1e6, "algorithmic_tflops": flops / secs / 1e12,
               "fp32_peak_measured_tflops": peak, "pct_of_fp32_peak": 100.0 * flops / secs / 1e12 / (peak * world),
               "Gtests_per_s": counters["tri_tests"] * launches / secs / 1e9, "mesh_files": bool(args.mesh_files),
               "counters_one_launch": counters}
        print(json.dumps(out))
        if args.write_ppm and img is not None:
            tracer.save_ppm(args.write_ppm, img, sc.width, sc.height)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
