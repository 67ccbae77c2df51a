"""Host-side sharding logic of simple_raytracer_b200/distributed.py under gloo, world_size 2, on CPU.
The renderer is an oracle-backed stand-in with the Tracer's duck-typed surface."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleRenderer:
    """Same surface as tracer.Tracer, computed by the CPU oracle (tests only)."""

    def __init__(self, scene, sky):
        import oracle
        self.o, self.scene, self.sky = oracle, scene, sky
        self.width, self.height = scene.width, scene.height
        self.canvas = np.zeros((scene.height, scene.width, 4), np.float32)
        self.bands = None
        self.output = np.zeros((scene.height, scene.width, 4), np.uint8)

    def clear_canvas(self):
        self.canvas[:] = 0

    def set_row_bands(self, h, i, n):
        self.bands = (h, i, n) if n > 1 else None

    def accumulate(self, rd):
        s = self.scene
        self.o.render(rd, s.scene_data, s.shapes, s.triangles, s.materials, self.sky, self.canvas,
                      bands=self.bands, threads=1)

    def resolve_device(self, steps):
        self.output[:] = self.o.average(steps, self.canvas)

    def resolve_device_range(self, steps, first, count):
        self.output.reshape(-1, 4)[first:first + count] = self.o.average(steps, self.canvas.reshape(-1, 4)[first:first + count])

    def resolve(self, steps):
        self.resolve_device(steps)
        return self.output

    def read_output(self, output=None):
        if output is not None:
            output[:] = self.output
            return output
        return self.output


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from simple_raytracer_b200 import distributed as D
    from simple_raytracer_b200 import scenes
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sky = scenes.procedural_skybox(64, 32, seed=11)
    sc = scenes.config2(40, 24, num_samples=2, launches=4)
    r = OracleRenderer(sc, sky)

    def reduce_fn(renderer, dst):
        t = torch.from_numpy(renderer.canvas)
        dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM)

    def gather_fn(renderer, rk, dst):
        t = torch.from_numpy(renderer.output)
        dist.reduce(t, dst=dst, op=dist.ReduceOp.MAX)
        return renderer.output if rk == dst else None

    img_s = D.render_sample_sharded(r, sc, rank, world, reduce_fn=reduce_fn)
    canvas_s = r.canvas.copy()
    r2 = OracleRenderer(sc, sky)
    img_t = D.render_tile_sharded(r2, sc, rank, world, band_height=4, gather_fn=gather_fn)
    assert r2.bands is None  # the tracer is left rendering full frames
    # the default exchange step: reduce-scatter, per-rank resolve of the own slice, gather of ARGB8 slices
    r3 = OracleRenderer(sc, sky)
    r3.set_row_bands(4, rank, world)  # sticky bands from an earlier tile-sharded run must not leak into this one
    img_rs = D.render_sample_sharded(r3, sc, rank, world)
    canvas_rs = D.gather_reduced_canvas(r3, rank, world)
    q.put((rank, img_s, canvas_s if rank == 0 else None, img_t, img_rs, canvas_rs))
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_schedules():
    from simple_raytracer_b200 import distributed as D
    assert D.launch_schedule(7, 0, 2) == [0, 2, 4, 6] and D.launch_schedule(7, 1, 2) == [1, 3, 5]
    all_k = sorted(k for r in range(8) for k in D.launch_schedule(64, r, 8))
    assert all_k == list(range(64))
    rows = [D.band_rows(1080, 8, r, 8) for r in range(8)]
    assert sorted(np.concatenate(rows).tolist()) == list(range(1080))
    assert abs(len(rows[0]) - len(rows[7])) <= 8
    assert list(D.band_rows(10, 2, 1, 3)) == [2, 3, 8, 9]


@pytest.mark.timeout(180)
def test_sample_and_tile_sharding_world2_gloo():
    import torch.multiprocessing as mp
    from simple_raytracer_b200 import scenes
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = {}
    for _ in range(2):
        rank, img_s, canvas_s, img_t, img_rs, canvas_rs = q.get(timeout=150)
        results[rank] = (img_s, canvas_s, img_t, img_rs, canvas_rs)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0

    # single-process reference with the same launches
    sky = scenes.procedural_skybox(64, 32, seed=11)
    sc = scenes.config2(40, 24, num_samples=2, launches=4)
    one = OracleRenderer(sc, sky)
    for k in range(4):
        one.accumulate(sc.render_data(k))
    want = one.resolve(4)
    img_s, canvas_s, img_t, img_rs, canvas_rs = results[0]
    assert results[1][0] is None and results[1][2] is None and results[1][3] is None
    # sample sharding: only the FP32 summation order differs (<= 1e-5 relative)
    assert np.allclose(canvas_s, one.canvas, rtol=1e-5, atol=1e-7)
    assert np.abs(img_s.astype(int) - want.astype(int)).max() <= 1
    # tile sharding: bit-identical by construction
    assert np.array_equal(img_t, want)
    # reduce-scatter tail: the same sums as the reduce-to-root tail, slice by slice
    assert np.array_equal(canvas_rs, canvas_s) and np.array_equal(img_rs, img_s)
