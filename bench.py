#!/usr/bin/env python
"""Headline benchmark of the path-tracing hot path (BASELINE.json metric: Msamples/s = px*spp/s).

    python bench.py --gpus N --steps K --warmup W          # CUDA arm (this repository)
    python bench.py --impl reference --steps K --warmup W  # CPU arm: the reference algorithm on host cores

Workload at N = 1: BASELINE config[1] -- four-sphere material scene + sky box, 1920x1080,
16 launches x num_samples 4 = 64 spp, 10 bounces.  One "step" = one full pass of that workload:
clear canvas, 16 `render` launches (submitted as one batch, srt_render_batch), one `average`.  At N > 1 (torchrun, one rank per GPU) every rank
renders its own 16 launches with distinct time seeds (sample sharding, weak scaling), the float
canvases are sum-reduced to rank 0 over NCCL, and rank 0 runs `average` with 16*N steps.

Keys printed (one JSON line, rank 0): see the driver contract; plus
  roofline      bound = fp32 (the path is FP32-pipe bound: SURVEY 8d / BASELINE.md section 2), achieved =
                counted algorithmic flop per `render` launch / mean launch duration (CUDA events on the
                launching stream), peak = FMA-chain micro-benchmark measured in this run
  cpu_baseline  the reference kernel itself (oracle/_ref = render.cl compiled by g++, OpenMP, all host cores) on
                a bounded sample of the same workload
  e2e           the same metric through the reference-facing Tracer API with host buffers: per step
                update_scene (H2D), clear_canvas, 16 x render(ticks, output) each with its ARGB8 read-back
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from simple_raytracer_b200 import scenes  # noqa: E402

# nominal algorithmic flop costs, SURVEY.md 8d
F_SPHERE, F_PLANE, F_MODEL, F_TRI, F_HIT, F_SKY, F_CAM, F_PIX, F_RESOLVE = 19, 14, 24, 46, 147, 75, 47, 6, 20
METRIC = "Msamples/s (px*spp/s), path-tracing kernel"


def algorithmic_flops(scene, counters, pixels_launches):
    """Counted work -> flop, SURVEY 8d.  counters: dict of the srt_counters fields."""
    t = scene.shapes["type"]
    per_bounce = F_SPHERE * int((t == 0).sum()) + F_PLANE * int((t == 1).sum()) + F_MODEL * int((t == 2).sum())
    return (counters["samples"] * F_CAM + counters["bounces"] * per_bounce + counters["tri_tests"] * F_TRI +
            counters["hits"] * F_HIT + counters["sky"] * F_SKY + pixels_launches * F_PIX)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if (t0 is None or t >= t0) and (t1 is None or t <= t1)] or \
               [r for (_, r) in self.rows]
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower() == "active" for r in rows)]
        pw = [float(r[2]) for r in rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(rows), "reasons": reasons}


def cpu_reference_run(scene, sky, steps, warmup, budget_s=4.0):
    """The reference's CPU path on host cores, all of them.  kind "reference": oracle/_ref, i.e. the
    reference's own src/render.cl compiled by g++ (built in the authoring container, shipped as a .so) with
    the NDRange rows spread over OpenMP threads; kind "port" (only if that library is missing): oracle.c.
    A step is a bounded sample of the workload: `n` full-frame launches, n chosen so a step takes about
    budget_s."""
    import oracle
    oracle.build()
    cores = oracle.max_threads()
    impl, kind = ("ref", "reference") if oracle.build_ref() else ("oracle", "port")
    t0 = time.perf_counter()
    oracle.render(scene.render_data(0), scene.scene_data, scene.shapes, scene.triangles, scene.materials, sky,
                  impl=impl)
    t_launch = time.perf_counter() - t0
    n = int(max(1, min(scene.launches, round(budget_s / max(t_launch, 1e-3)))))
    samples_step = scene.width * scene.height * scene.num_samples * n
    times = []
    for s in range(warmup + steps):
        canvas = None
        t0 = time.perf_counter()
        for k in range(n):
            canvas, _ = oracle.render(scene.render_data(k), scene.scene_data, scene.shapes, scene.triangles,
                                      scene.materials, sky, canvas, impl=impl)
        oracle.average(n, canvas, impl=impl)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    total = sum(times)
    value = samples_step * len(times) / total / 1e6
    sample = f"{scene.width}x{scene.height} full frame, {n} of {scene.launches} launches x {scene.num_samples} spp per step"
    return value, total / len(times) * 1e3, cores, sample, kind


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--config", type=int, default=2, help="BASELINE config index 1..5 (default 2 = configs[1])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    scene = scenes.CONFIGS[args.config]()
    workload = {"workload": f"BASELINE config {args.config}: {scene.name}", "resolution": f"{scene.width}x{scene.height}",
                "spp": scene.spp, "launches": scene.launches, "num_samples": scene.num_samples,
                "num_bounces": scene.num_bounces, "shapes": int(len(scene.shapes)), "triangles": int(len(scene.triangles))}

    if args.impl == "reference":
        if rank != 0:
            return 0
        sky = scenes.procedural_skybox()
        value, ms, cores, sample, kind = cpu_reference_run(scene, sky, args.steps, args.warmup)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload,
            "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": ("the reference's own src/render.cl compiled for the host by g++ -O2 (oracle/_ref: OpenCL-C shim + "
                     "NDRange loop, OpenMP over rows, all host cores); no OpenCL runtime exists in this image"
                     if kind == "reference" else
                     "oracle/_ref is missing: oracle/oracle.c, the scalar C restatement of render.cl, OpenMP over rows")}))
        return 0

    import torch
    import torch.distributed as dist
    from simple_raytracer_b200 import distributed as srt_dist
    from simple_raytracer_b200.tracer import Tracer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl cuda needs a GPU: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sky = scenes.procedural_skybox()
    tr = Tracer(scene.width, scene.height, sky, device=local_rank)
    tr.scene_data[:] = scene.scene_data
    tr.update_scene(scene.shapes, scene.triangles, scene.materials)
    stream = torch.cuda.ExternalStream(tr.stream_handle())
    L = scene.launches
    my_rds = [scene.render_data(k * world + rank) for k in range(L)]  # distinct time seeds per rank
    pixels = scene.width * scene.height
    samples_per_rank_step = pixels * scene.num_samples * L
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def device_step():
        tr.clear_canvas()
        tr.accumulate_batch(my_rds)  # the step's 16 launches, one persistent kernel (srt_render_batch)
        if world > 1:
            srt_dist.reduce_canvas(tr, dst=0)
        if rank == 0:
            tr.resolve_device(L * world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm -----------------------------------------------------------------
    for _ in range(args.warmup):
        device_step()
    barrier()
    tr.render_time_ms()  # drop warm-up launch timings
    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.25)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_region0 = time.perf_counter()
    for s in range(args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()  # L2 flush between timed iterations, outside the event pair
        evs[s][0].record(stream)
        device_step()
        evs[s][1].record(stream)
    barrier()
    t_region1 = time.perf_counter()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    my_ms = sum(step_ms)
    kernel_ms, kernel_launches = tr.render_time_ms()
    clocks = sampler.stop(t_region0, t_region1) if sampler else None
    tmax = torch.tensor([my_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms = float(tmax.item())
    value = samples_per_rank_step * world * args.steps / (total_ms * 1e-3) / 1e6

    # ---- end-to-end arm: the reference-facing Tracer protocol with host buffers -----------------
    output = np.empty(pixels * 4, np.uint8)
    h2d = scene.shapes.nbytes + scene.triangles.nbytes + scene.materials.nbytes + 96 + L * 112
    d2h = L * pixels * 4

    def e2e_step():
        tr.clear_canvas()                                                 # src/main.cpp:277-280
        tr.update_scene(scene.shapes, scene.triangles, scene.materials)
        for i, rd in enumerate(my_rds):
            tr.options[:] = rd                                            # src/main.cpp:283-288
            tr.render(i + 1, output)                                      # src/main.cpp:290

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = samples_per_rank_step * world * args.steps / float(te.item()) / 1e6

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (untimed instrumented pass) ------------------------------
    peak_tf, est_mhz = tr.measure_fp32_peak()
    cnt = None
    for rd in my_rds:
        cnt = tr.accumulate_counted(rd, cnt)
    counters = {n: int(cnt[0][n]) for n in cnt.dtype.names}
    flops_per_launch = algorithmic_flops(scene, counters, pixels * L) / L
    launch_ms = kernel_ms / max(kernel_launches, 1)
    achieved = flops_per_launch / (launch_ms * 1e-3) / 1e12
    nominal = 148 * 128 * 2 * 1.965e9 / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    canvas_bytes = pixels * 32  # canvas RMW per launch: the only HBM-resident traffic of the kernel
    roofline = {"bound": "fp32", "kernel": "srt::render_kernel<COUNT=false, MODE=%s>" % ("BIG_MODELS" if len(scene.triangles) > 32 else ("SMALL_MODELS" if len(scene.triangles) else "ANALYTIC")), "achieved": achieved, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": achieved / peak_tf if peak_tf else None,
                "peak_source": "FMA-chain micro-benchmark measured in this run (srt_measure_fp32_peak); "
                               f"nominal 148 SM x 128 x 2 x 1.965 GHz = {nominal:.1f}",
                "frac_of_nominal": achieved / nominal, "flops_per_launch": flops_per_launch,
                "launch_ms": launch_ms, "launches_timed": int(kernel_launches),
                "kernel_share_of_step": kernel_ms / total_ms if total_ms else None,
                "hbm": {"algorithmic_bytes_per_launch": canvas_bytes,
                        "achieved_gbs": canvas_bytes / (launch_ms * 1e-3) / 1e9,
                        "peak_gbs": peaks.get("hbm_gbs", 6650.0),
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback B200_PROFILING.md"},
                "traffic": None, "counters": counters}
    # static evidence from the committed ncu capture of the same launch (profiles/): DRAM traffic, and the view that
    # actually bounds these kernels -- warp-instruction issue (4 schedulers x 1 inst/clk per SM)
    prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(prof):
        try:
            pj = json.load(open(prof))
            roofline["traffic"] = pj.get(f"config{args.config}")
            inst = pj.get(f"config{args.config}_inst_executed")
            if inst:
                issue_peak = 148 * 4 * (clocks["sm_mhz"] or 1965.0) * 1e6 / 1e9  # G warp-inst/s
                issue_ach = inst / (launch_ms * 1e-3) / 1e9
                roofline["issue"] = {"achieved": issue_ach, "peak": issue_peak, "unit": "G warp-inst/s",
                                     "frac": issue_ach / issue_peak, "warp_inst_per_launch": inst,
                                     "source": pj.get(f"config{args.config}_capture")}
        except (OSError, ValueError):
            pass

    cpu_baseline = None
    if not args.no_cpu_baseline:
        v, ms, cores, sample, kind = cpu_reference_run(scene, sky, steps=2, warmup=0, budget_s=6.0)
        cpu_baseline = {"value": v, "unit": "Msamples/s", "cores": cores, "kind": kind, "sample": sample}

    print(json.dumps({
        "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload, parallelism=f"sample-sharded x{world}" if world > 1 else "single GPU",
                       l2="flushed between timed steps (256 MiB memset, outside the event pairs)",
                       timing="CUDA events per step on the launching stream, summed; max over ranks"),
        "clocks": clocks, "gpu_launches": args.steps * 3,  # per step: render (16 launches batched) + accumulate + average
        "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "api": "Tracer.update_scene + clear_canvas + 16 x Tracer.render(ticks, host_output) per step, wall clock"},
        "roofline": roofline, "cpu_baseline": cpu_baseline,
        "fp32_peak_measured_tflops": peak_tf, "sm_clock_est_mhz": est_mhz}))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
