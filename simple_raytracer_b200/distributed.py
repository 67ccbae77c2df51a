"""Multi-GPU sharding of the path: one process per GPU, torch.distributed for the plumbing.

The reference is single-device (src/tracer.cpp:13).  Its kernel shards trivially because every
sample's RNG stream depends only on (sample, global pixel id, num_samples, time) (render.cl:488,:496):

* sample sharding (BASELINE config 4): launch k (with its own time_k) goes to rank k % world; each
  rank accumulates a private canvas; ONE sum-reduce of the float canvases at the end, then
  `average` with the total launch count on rank 0.  Differs from 1 GPU by FP32 summation order only.
* tile sharding (BASELINE config 5): rows are dealt in interleaved bands (srt_set_row_bands); global
  pixel ids are preserved, so each pixel is bit-identical to the 1-GPU result; every rank resolves
  its bands and the ARGB8 images are combined with a MAX-reduce (unowned pixels resolve to A=255,
  RGB=0, owned ones are >= that bytewise).

There is no data-path collective inside a launch; the collective is the exchange step that follows
the last launch.  The `renderer` argument is duck-typed (tracer.Tracer on a GPU; the CPU tests pass an
oracle-backed stand-in) so the host logic is covered with gloo, world_size 2.
"""
import numpy as np


def launch_schedule(total_launches, rank, world):
    """Launch indices of this rank: k = rank, rank + world, ... (round-robin)."""
    return list(range(rank, total_launches, world))


def band_rows(height, band_height, band_index, band_count):
    """Rows owned by a rank under interleaved row bands (mirrors srt_set_row_bands)."""
    y = np.arange(height)
    return y[(y // band_height) % band_count == band_index] if band_count > 1 else y


def _dist():
    import torch.distributed as dist
    return dist


def render_sample_sharded(renderer, scene, rank, world, total_launches=None, dst=0, reduce_fn=None):
    """Sample-sharded accumulation.  Returns the ARGB8 image on rank `dst` (None elsewhere)."""
    total = scene.launches if total_launches is None else total_launches
    renderer.clear_canvas()
    _accumulate_all(renderer, [scene.render_data(k) for k in launch_schedule(total, rank, world)])
    if world > 1:
        (reduce_fn or reduce_canvas)(renderer, dst)
    if rank == dst:
        return renderer.resolve(total)
    return None


def render_tile_sharded(renderer, scene, rank, world, band_height=1, total_launches=None, dst=0,
                        gather_fn=None):
    """Tile-sharded accumulation (interleaved row bands).  Returns ARGB8 on rank `dst`.
    band_height = 1 deals single rows round-robin: the cost of a row varies smoothly with y, so every rank gets
    the same load to within a row (config 5 on 8 GPUs: slowest rank 16 % above the fastest with 8-row bands, 3 % with
    single rows);
    a row of 1920 pixels x num_samples items is still far more than a warp needs for coherent first bounces."""
    total = scene.launches if total_launches is None else total_launches
    renderer.set_row_bands(band_height, rank, world)
    renderer.clear_canvas()
    _accumulate_all(renderer, [scene.render_data(k) for k in range(total)])
    renderer.resolve_device(total)
    if world > 1:
        return (gather_fn or gather_output)(renderer, rank, dst)
    return renderer.read_output()


def _accumulate_all(renderer, render_datas):
    """All launches of a rank, batched into shared persistent kernels where the renderer can (srt_render_batch)."""
    if hasattr(renderer, "accumulate_batch"):
        if render_datas:
            renderer.accumulate_batch(render_datas)
    else:
        for rd in render_datas:
            renderer.accumulate(rd)


def canvas_tensor(renderer):
    """torch view (no copy) of the renderer's device canvas via __cuda_array_interface__."""
    import torch
    return torch.as_tensor(renderer.canvas_view(), device="cuda")


def output_tensor(renderer):
    import torch
    return torch.as_tensor(renderer.output_view(), device="cuda")


def reduce_canvas(renderer, dst=0):
    """NCCL sum-reduce of the per-GPU float canvases into rank dst, ordered after the render
    launches on the tracer's own stream."""
    import torch
    dist = _dist()
    stream = torch.cuda.ExternalStream(renderer.stream_handle())
    with torch.cuda.stream(stream):
        dist.reduce(canvas_tensor(renderer), dst=dst, op=dist.ReduceOp.SUM)


def gather_output(renderer, rank, dst=0):
    """Combine the per-rank ARGB8 band images on rank dst (MAX-reduce, see module docstring)."""
    import torch
    dist = _dist()
    stream = torch.cuda.ExternalStream(renderer.stream_handle())
    with torch.cuda.stream(stream):
        out = output_tensor(renderer)
        dist.reduce(out, dst=dst, op=dist.ReduceOp.MAX)
    stream.synchronize()
    return out.cpu().numpy() if rank == dst else None
